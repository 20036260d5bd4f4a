"""`embodied.Agent` surface of the B200-native DreamerV2+ learner.

Drop-in for the reference's agent (embodied/agents/dreamerv2plus/agent.py:15-139
behind tfagent.py:22-134): same constructor, `configs`, `dataset`, `train`,
`report`, `policy`, `save`, `load`; numpy in, numpy / python scalars out; methods
are plain instance attributes so `embodied.Timer.wrap` can re-bind them
(embodied/core/timer.py:28-31).  The work happens in engine.Engine on
hand-written sm_100a kernels; this file is host plumbing only (config ->
Spec, pinned staging, CUDA-graph capture/replay, metric read-back).
"""
import os
import re
import sys
import threading
import queue as queuelib

import numpy as np
import torch

from . import configs as _configs
from . import engine as _engine
from .backend import CudaBackend


def _get(cfg, path):
  node = cfg
  for part in path.split('.'):
    node = node[part] if isinstance(node, dict) or hasattr(node, '__getitem__') else getattr(node, part)
  return node


def _space_shape(space):
  return tuple(getattr(space, 'shape', space))


def spec_from_config(config, obs_space, act_space) -> _engine.Spec:
  """Resolve an embodied.Config (or nested dict) + spaces into engine.Spec,
  rejecting loudly every setting whose kernels are not built."""
  g = lambda p: _get(config, p)
  act = act_space['action'] if isinstance(act_space, dict) else act_space
  discrete = bool(getattr(act, 'discrete', False))
  if g('task_behavior') != 'Greedy' or g('expl_behavior') not in ('None', 'Random'):
    raise NotImplementedError('only task_behavior=Greedy with expl_behavior=None|Random are built')
  if g('critic_type') != 'vfunction':
    raise NotImplementedError('only critic_type=vfunction is built')
  checks = {
      'rssm.act': 'elu', 'rssm.norm': 'layer', 'rssm.initial': 'learned2',
      'rssm.post_layers': 1, 'rssm.gru_layers': 1, 'encoder.act': 'elu',
      'encoder.norm': 'layer', 'reward_head.dist': 'symlog', 'cont_head.dist': 'binary',
      'critic.dist': 'symlog', 'actor_dist_cont': 'normal', 'actor_grad_cont': 'backprop',
      'actor_dist_disc': 'onehot', 'actor_grad_disc': 'reinforce',
      'actor_return': 'gve', 'critic_return': 'gve', 'slow_target': True,
      'slow_target_fraction': 1.0, 'advnorm.impl': 'mean_std', 'retnorm.impl': 'off',
      'scorenorm.impl': 'off', 'wmkl.impl': 'mult', 'actent.impl': 'mult',
      'transform_rewards': 'off', 'train_wm': True, 'actent_norm': True,
  }
  for key, want in checks.items():
    if g(key) != want:
      raise NotImplementedError(f'config {key}={g(key)!r} is not built (only {want!r})')
  if not g('rssm.classes'):
    raise NotImplementedError('continuous (Gaussian) latents are not built')
  # settings the kernels would otherwise ignore silently
  if float(g('priority_correct')) != 0.0:
    raise NotImplementedError('priority_correct != 0 (importance-weighted model loss) is not built')
  scales = g('loss_scales')
  for key in (scales.keys() if hasattr(scales, 'keys') else []):
    if float(scales[key]) != 1.0:
      raise NotImplementedError(f'loss_scales.{key}={scales[key]} is not built (only 1.0)')
  if not {'decoder', 'reward', 'cont'} <= set(g('grad_heads')):
    raise NotImplementedError('grad_heads must contain decoder, reward and cont')
  if float(g('actent.vel')) != float(g('wmkl.vel')):
    raise NotImplementedError('actent.vel must equal wmkl.vel (one controller velocity is built)')
  shapes = {k: _space_shape(v) for k, v in obs_space.items() if not k.startswith('log_')}

  def vec_keys(pattern, excluded, ranks):
    return {k: int(np.prod(s)) for k, s in shapes.items()
            if k not in excluded and re.match(pattern, k) and len(s) in ranks}

  def cnn_keys(pattern, excluded):
    return {k: tuple(int(d) for d in s) for k, s in shapes.items()
            if k not in excluded and re.match(pattern, k) and len(s) == 3}

  enc_ex = ('is_first', 'is_last')
  dec_ex = ('is_first', 'is_last', 'is_terminal', 'reward')
  enc_cnn = cnn_keys(g('encoder.cnn_keys'), enc_ex)
  dec_cnn = cnn_keys(g('decoder.cnn_keys'), dec_ex)
  if enc_cnn != dec_cnn:
    raise NotImplementedError(f'encoder image keys {list(enc_cnn)} must equal decoder image '
                              f'keys {list(dec_cnn)}')
  if enc_cnn:
    for part in ('encoder', 'decoder'):
      if g(f'{part}.cnn') != 'simple':
        raise NotImplementedError(f'{part}.cnn={g(f"{part}.cnn")!r} is not built (only simple)')
    if g('decoder.image_dist') != 'mse':
      raise NotImplementedError('decoder.image_dist must be mse')
    if int(g('encoder.cnn_depth')) != int(g('decoder.cnn_depth')):
      raise NotImplementedError('encoder.cnn_depth must equal decoder.cnn_depth')
  enc = vec_keys(g('encoder.mlp_keys'), enc_ex, (0, 1))
  dec = vec_keys(g('decoder.mlp_keys'), dec_ex, (1,))
  if not (enc or enc_cnn):
    raise NotImplementedError('the encoder matches no observation key')
  if not (dec or dec_cnn):
    raise NotImplementedError('the decoder matches no observation key')
  heads = [g(f'{h}.layers') for h in ('reward_head', 'cont_head', 'actor', 'critic')]
  units = [g(f'{h}.units') for h in ('reward_head', 'cont_head', 'actor', 'critic')]
  if len(set(heads)) != 1 or len(set(units)) != 1:
    raise NotImplementedError('reward / cont / actor / critic heads must share layers and units')
  for h in ('decoder', 'reward_head', 'cont_head', 'actor', 'critic'):
    if g(f'{h}.act') != 'elu' or g(f'{h}.norm') != 'layer':
      raise NotImplementedError(f'{h}: only act=elu, norm=layer are built')
    # settings the kernels hard-code: every head reads [deter, stoch] (A/nets.py:612-626) and its
    # output layer is initialised with outscale 0.1 (A/nets.py:448-450)
    try:
      inputs, outscale = list(g(f'{h}.inputs')), float(g(f'{h}.outscale'))
    except (KeyError, AttributeError, TypeError):
      continue
    if inputs != ['deter', 'stoch']:
      raise NotImplementedError(f'{h}.inputs={inputs} is not built (only [deter, stoch])')
    if outscale != 0.1:
      raise NotImplementedError(f'{h}.outscale={outscale} is not built (only 0.1)')
  opts = [tuple(g(f'{o}.{k}') for k in ('lr', 'eps', 'clip', 'wd', 'warmup'))
          for o in ('model_opt', 'actor_opt', 'critic_opt')]
  if len(set(opts)) != 1 or opts[0][4]:
    raise NotImplementedError('optimizers must share lr/eps/clip/wd and warmup=0')
  return _engine.Spec(
      obs_dims=enc, dec_dims=None if list(enc.items()) == list(dec.items()) else dec,
      act_dim=int(np.prod(_space_shape(act))), discrete=discrete,
      deter=int(g('rssm.deter')), units=int(g('rssm.units')), stoch=int(g('rssm.stoch')),
      classes=int(g('rssm.classes')), prior_layers=int(g('rssm.prior_layers')),
      rssm_unimix=float(g('rssm.unimix')), enc_layers=int(g('encoder.mlp_layers')),
      enc_units=int(g('encoder.mlp_units')), head_layers=int(heads[0]),
      head_units=int(units[0]), dec_layers=int(g('decoder.mlp_layers')),
      dec_units=int(g('decoder.mlp_units')), horizon=int(g('imag_horizon')), discount=float(g('discount')),
      lam=float(g('return_lambda')), kl_balance=float(g('wmkl_balance')),
      wmkl_target=float(g('wmkl.target')), wmkl_min=float(g('wmkl.min')),
      wmkl_max=float(g('wmkl.max')), actent_target=float(g('actent.target')),
      actent_min=float(g('actent.min')), actent_max=float(g('actent.max')),
      ctrl_vel=float(g('wmkl.vel')), actor_minstd=float(g('actor.minstd')),
      actor_maxstd=float(g('actor.maxstd')), actor_unimix=float(g('actor.unimix')),
      lr=float(opts[0][0]), adam_eps=float(opts[0][1]), clip=float(opts[0][2]),
      wd=float(opts[0][3]), slow_every=int(g('slow_target_update')),
      advnorm_decay=float(g('advnorm.decay')), advnorm_max=float(g('advnorm.max')),
      cnn_keys=enc_cnn, cnn_depth=int(g('encoder.cnn_depth')),
      enc_kernels=tuple(int(k) for k in g('encoder.cnn_kernels')),
      dec_kernels=tuple(int(k) for k in g('decoder.cnn_kernels')))


class Batcher:
  """zip() over `batch_size` replay generators with worker threads: the role of
  embodied.Prefetch (embodied/core/prefetch.py:6-67) in Agent.dataset()."""

  def __init__(self, generator, batch_size, workers=8, prefetch=4, keys=None):
    """keys: the batch keys the learner reads; everything else a replay step carries (e.g. a1's
    unused 12 KB camera image per step, `log_*` entries) is dropped BEFORE stacking, so the host
    never assembles or copies it."""
    self._queues = [queuelib.Queue(prefetch) for _ in range(batch_size)]
    self._keys = None if keys is None else set(keys)
    self._running = True
    for w in range(min(workers, batch_size)):
      qs = self._queues[w::workers]
      threading.Thread(target=self._fill, args=(generator, qs), daemon=True).start()

  def _fill(self, generator, queues):
    try:
      its = [generator() for _ in queues]
      while self._running:
        for it, q in zip(its, queues):
          q.put(next(it))
    except Exception as e:      # surface worker errors to the consumer (it may be blocked on any
      for q in queues:          # of this worker's queues); the worker then ends
        q.put(e)

  def __iter__(self):
    return self

  def __next__(self):
    elems = []
    for q in self._queues:
      e = q.get()
      if isinstance(e, Exception):
        raise e
      elems.append(e)
    keys = [k for k in elems[0] if self._keys is None or k in self._keys]
    return {k: np.stack([e[k] for e in elems], 0) for k in keys}


class Agent:

  configs = _configs.CONFIGS

  def __init__(self, obs_space, act_space, step, config, seed=None, _backend_for_tests=None):
    """`_backend_for_tests` is a test seam, never a fallback: the product path always constructs CudaBackend (which raises
    without a GPU / the kernel library); tests/test_reference_loop.py passes the torch-CPU
    primitive simulator to drive the reference's own run loop where there is no GPU."""
    self.config = config
    self.obs_space = obs_space
    self.act_space = act_space['action'] if isinstance(act_space, dict) else act_space
    self.step = step
    self.spec = spec_from_config(config, obs_space, act_space)
    self.be = _backend_for_tests if _backend_for_tests is not None else CudaBackend()  # raises without GPU / .so
    self._on_gpu = self.be.device.type == 'cuda'
    self.seed = int(_get(config, 'seed')) if seed is None else int(seed)
    self._rng = np.random.default_rng(self.seed)      # exploration / action noise draws
    self.world, self.rank, self._allreduce = 1, 0, None
    if torch.distributed.is_available() and torch.distributed.is_initialized():
      self.world = torch.distributed.get_world_size()
      self.rank = torch.distributed.get_rank()
      if self.world > 1:
        self._allreduce = lambda t: torch.distributed.all_reduce(t)
    self._engines = {}
    self._main = None
    self._policies = {}
    self._graphs = {}
    self._calls = {}
    self._slow_updates = -1           # host mirror of the device slow-target counter
    self.use_graph = self._on_gpu
    self._init_params = _engine.init_params(self.spec, seed=self.seed + 1)
    # behaviors.Random (A/behaviors.py:97-120): uniform actions in 'explore' mode, nothing to train
    self.expl_random = str(_get(config, 'expl_behavior')) == 'Random'
    self.expl_noise = float(_get(config, 'expl_noise'))
    self.eval_noise = float(_get(config, 'eval_noise'))

  # -- engines ----------------------------------------------------------------
  def _engine_for(self, B, L):
    key = (B, L)
    if key not in self._engines:
      # every rank draws different noise but starts from identical weights
      eng = _engine.Engine(self.spec, self.be, B, L, seed=self.seed * 1000003 + self.rank,
                           world=self.world, allreduce=self._allreduce, share=self._main)
      if self._main is None:
        eng.load_params(self._init_params)
        self._main = eng
      self._engines[key] = eng
      self._attach_stage(eng)
    return self._engines[key]

  def _attach_stage(self, eng):
    """Pinned host + device staging buffers for one engine's [B, L, ...] inputs."""
    B, L = eng.B, eng.L
    pin = self._on_gpu
    eng.h_stage = {
        k: torch.zeros(B, L, d, pin_memory=pin)
        for k, d in (('obs', self.spec.obs_total), ('action', self.spec.act_dim),
                     ('first', 1), ('reward', 1), ('cont', 1))}
    if self.spec.dec_dims is not None:          # decoder targets differ from the encoder input
      eng.h_stage['dec'] = torch.zeros(B, L, self.spec.dec_total, pin_memory=pin)
    # images travel as uint8 and are scaled / channel-concatenated on the device
    for k, shp in self.spec.cnn_keys.items():
      eng.h_stage[f'img/{k}'] = torch.zeros(B, L, *shp, dtype=torch.uint8, pin_memory=pin)
    eng.d_stage = {k: torch.zeros_like(v, device=self.be.device) for k, v in eng.h_stage.items()}

  def _main_engine(self):
    if self._main is None:
      B = int(_get(self.config, 'batch_size'))
      L = int(_get(self.config, 'replay_chunk'))
      self._engine_for(max(B // self.world, 1), L)
    return self._main

  def _stage(self, eng, data, state):
    """numpy batch [B, L, ...] -> pinned host -> device -> time-major buffers."""
    s = self.spec
    B, L = eng.B, eng.L
    h = eng.h_stage
    off = 0
    for k, d in s.obs_dims.items():
      h['obs'][:, :, off:off + d] = torch.from_numpy(
          np.asarray(data[k], np.float32).reshape(B, L, d))
      off += d
    if s.dec_dims is not None:
      off = 0
      for k, d in s.dec_dims.items():
        h['dec'][:, :, off:off + d] = torch.from_numpy(
            np.asarray(data[k], np.float32).reshape(B, L, d))
        off += d
    for k, shp in s.cnn_keys.items():
      arr = np.asarray(data[k])
      if arr.dtype != np.uint8:
        raise TypeError(f'image key {k!r} must be uint8 (got {arr.dtype}); A/agent.py:128')
      h[f'img/{k}'].copy_(torch.from_numpy(arr.reshape((B, L) + tuple(shp))))
    h['action'].copy_(torch.from_numpy(np.asarray(data['action'], np.float32).reshape(B, L, -1)))
    first = np.asarray(data['is_first'], np.float32).reshape(B, L, 1)
    h['first'].copy_(torch.from_numpy(first))
    h['reward'].copy_(torch.from_numpy(np.asarray(data['reward'], np.float32).reshape(B, L, 1)))
    h['cont'].copy_(torch.from_numpy(
        1.0 - np.asarray(data['is_terminal'], np.float32).reshape(B, L, 1)))
    for k in h:
      eng.d_stage[k].copy_(h[k], non_blocking=True)
    if state is None:
      eng.first0.fill_(1.0)
    else:
      eng.first0.copy_(torch.from_numpy(first[:, 0, 0].copy()), non_blocking=True)
      # the returned state is a snapshot; skip the copy back only when the engine's buffers still
      # hold exactly that snapshot (no other train() call on this engine since)
      if state.get('_version') != (id(eng), getattr(eng, '_state_version', -1)):
        eng.prev_deter.copy_(torch.as_tensor(state['deter']).to(self.be.device))
        eng.prev_stoch.copy_(torch.as_tensor(state['stoch']).to(self.be.device).reshape(B, -1))

  @staticmethod
  def _unstage(eng):
    be, d = eng.be, eng.d_stage
    if eng.spec.obs_dims:
      be.transpose01(d['obs'], eng.obs)
    if eng.spec.dec_dims is not None and eng.spec.dec_total:
      be.transpose01(d['dec'], eng.dec_tgt)
    c0 = 0
    for k, shp in eng.spec.cnn_keys.items():
      be.image_stage(d[f'img/{k}'], eng.img, c0, 1.0 / 255.0)
      c0 += shp[2]
    be.transpose01(d['action'], eng.action)
    be.transpose01(d['first'], eng.first)
    be.transpose01(d['reward'], eng.reward)
    be.transpose01(d['cont'], eng.cont)

  def _run(self, eng, name, fn):
    """Eager on the first call of a launch sequence, CUDA-graph replay after.  With a
    collective hook installed (data parallel) the sequence is captured as graph
    SEGMENTS split at every all-reduce and the collectives run eagerly between the
    segment replays: NCCL never executes inside a captured graph."""
    key = (id(eng), name)
    n = self._calls.get(key, 0)
    self._calls[key] = n + 1
    if not self.use_graph or n == 0:
      fn()
      return
    if key not in self._graphs:
      self._graphs[key] = self._capture(eng, fn)
    for kind, item in self._graphs[key]:
      if kind == 'graph':
        item.replay()
      else:
        self._allreduce(item)

  def close(self):
    """Drop the captured CUDA graphs.  Under data parallelism the graphs hold NCCL kernels of the
    default process group: call this BEFORE torch.distributed.destroy_process_group() (tearing the
    communicator down first left the process hanging at exit on 2 x B200, r02)."""
    torch.cuda.synchronize()
    self._graphs.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()

  def _capture(self, eng, fn):
    main = getattr(eng, 'eng', eng)            # PolicyEngine shares the learner's hooks
    be = self.be
    segs, cur = [], {}
    # above the default priority: launches of a concurrent branch (Engine.launch_step) fill in behind
    side = torch.cuda.Stream(device=be.device, priority=-1)

    def begin():
      cur['g'] = torch.cuda.CUDAGraph()
      cur['n0'] = be.launch_count()
      cur['g'].capture_begin(capture_error_mode='thread_local')

    def end():
      cur['g'].capture_end()
      if be.launch_count() > cur['n0']:
        segs.append(('graph', cur['g']))

    def hook(t):
      end()
      segs.append(('allreduce', t))
      begin()

    saved = main.allreduce
    torch.cuda.synchronize()
    if saved is not None and os.environ.get('DD_GRAPH_NCCL', '1') != '0':
      # data parallel: NCCL all-reduces captured INSIDE the graph (one graph per step, no host
      # round trip between the segments); falls back to split segments if the capture fails
      try:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
          begin()
          fn()
          end()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        return segs
      except Exception as e:     # noqa: BLE001  (any capture failure -> the segmented form)
        print(f'[daydreamer_b200] NCCL-in-graph capture failed ({type(e).__name__}: {e}); '
              'using graph segments with eager collectives', file=sys.stderr)
        try:
          cur['g'].capture_end()
        except Exception:        # noqa: BLE001
          pass
        segs.clear()
        torch.cuda.synchronize()
    overlap = getattr(main, 'overlap', False)
    if saved is not None:
      main.allreduce = hook
      main.overlap = False       # a segment cannot end while a second stream is still forked
    side.wait_stream(torch.cuda.current_stream())
    try:
      with torch.cuda.stream(side):
        begin()
        fn()
        end()
    finally:
      main.allreduce = saved
      if saved is not None:
        main.overlap = overlap
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    return segs

  # -- embodied.Agent surface -------------------------------------------------
  def batch_keys(self):
    """Every batch key train() / report() read."""
    keys = set(self.spec.obs_dims) | set(self.spec.dec_vec) | set(self.spec.cnn_keys)
    keys |= {'action', 'reward', 'is_first', 'is_last', 'is_terminal', 'key', 'prob'}
    keys.add(str(_get(self.config, 'priority')))
    return keys

  def dataset(self, generator):
    """Batches of `batch_size` sequences; under data parallelism `batch_size` is the GLOBAL batch
    (as in the reference, A/tfagent.py:112-116) and every rank draws its own batch_size / world
    sequences from its own replay."""
    per_rank = max(1, int(_get(self.config, 'batch_size')) // self.world)
    return Batcher(generator, per_rank, keys=self.batch_keys())

  def train(self, data, state=None):
    B, L = np.asarray(data['is_first']).shape[:2]
    eng = self._engine_for(B, L)
    self._stage(eng, data, state)

    refresh = eng.slow_refresh_due(self._slow_updates)
    self._slow_updates = (0 if refresh else self._slow_updates) + 1

    def launch():
      self._unstage(eng)
      eng.launch_step(slow_refreshed=refresh)
    # two launch sequences (with / without the slow-critic re-evaluation), each its own graph
    self._run(eng, 'train_refresh' if refresh else 'train', launch)
    metrics = eng.read_metrics()
    if not np.isfinite(metrics['model_loss']):
      raise FloatingPointError('model_loss is not finite')       # A/tfutils.py:207
    outs = {}
    if 'key' in data:                    # prioritized replay hook (A/agent.py:90-92)
      outs = {'key': data['key'], 'priority': self._criterion(eng, data)}
    eng._state_version = getattr(eng, '_state_version', 0) + 1
    state = {'deter': eng.prev_deter.clone(), 'stoch': eng.prev_stoch.clone(),
             '_version': (id(eng), eng._state_version)}
    return outs, state, metrics

  def _criterion(self, eng, data):
    """`criteria[config.priority]` of the reference: a data key or a `<head>_loss`
    of the world model, as a [B, L] array."""
    name = str(_get(self.config, 'priority'))
    if name in data:
      return np.asarray(data[name])
    rows = {f'{k}_loss': v for k, v in eng.loss_rows.items() if k != 'total'}
    rows['kl_loss'] = eng.kl_row
    if name not in rows:
      raise KeyError(f'priority={name!r} is neither a batch key nor one of {sorted(rows)}')
    tm = rows[name].detach().cpu().numpy().reshape(eng.L, eng.B)
    if name == 'kl_loss':
      tm = tm * float(eng.wmkl_scale.cpu().numpy()[0])
    return np.ascontiguousarray(tm.T)

  def report(self, data):
    B, L = np.asarray(data['is_first']).shape[:2]
    eng = self._engine_for(B, L)
    self._stage(eng, data, None)

    def launch():
      self._unstage(eng)
      eng.launch_report()
    self._run(eng, 'report', launch)
    full = eng.read_metrics(report=True)
    out = {k: np.float32(v) for k, v in full.items()}
    if self.spec.cnn_keys and L > _engine.VideoReport.CONTEXT:
      out.update(self._report_videos(data, min(6, B), L))
    return out

  def _report_videos(self, data, B6, L):
    """openl_<key> / task_imag_<key> image summaries (A/agent.py:266-282,
    A/behaviors.py:32-46) through engine.VideoReport."""
    key = ('video', B6, L)
    if key not in self._engines:
      rep = _engine.VideoReport(self._main, B6, L)
      self._attach_stage(rep.ctx)
      self._engines[key] = rep
    rep = self._engines[key]
    C5 = rep.CONTEXT
    sub = {k: np.asarray(v)[:B6, :C5] for k, v in data.items()
           if k in self.spec.obs_dims or k in self.spec.dec_vec or k in self.spec.cnn_keys or
           k in ('action', 'is_first', 'reward', 'is_terminal')}
    self._stage(rep.ctx, sub, None)
    acts = np.asarray(data['action'], np.float32)[:B6, C5:].reshape(B6, L - C5, -1)
    rep.actions.copy_(torch.from_numpy(np.ascontiguousarray(acts.swapaxes(0, 1))))
    self._unstage(rep.ctx)
    rep.launch()
    self.be.sync()
    truth = {k: np.asarray(data[k], np.float32)[:B6] / np.float32(255.0) for k in self.spec.cnn_keys}
    return rep.videos(truth)

  def policy(self, obs, state=None, mode='train'):
    obs = {k: v for k, v in obs.items() if not k.startswith('log_')}
    P = len(np.asarray(obs['is_first']))
    main = self._main_engine()
    if P not in self._policies:
      self._policies[P] = _engine.PolicyEngine(main, P)
    pol = self._policies[P]
    s = self.spec
    if s.obs_dims:
      vec = np.concatenate(
          [np.asarray(obs[k], np.float32).reshape(P, -1) for k in s.obs_dims], -1)
      pol.obs.copy_(torch.from_numpy(vec))
    for k, shp in s.cnn_keys.items():
      pol.img_u8[k].copy_(torch.from_numpy(
          np.asarray(obs[k], np.uint8).reshape((P, 1) + tuple(shp))))
    first = np.asarray(obs['is_first'], np.float32).reshape(P)
    if state is None:
      first = np.ones(P, np.float32)
      pol.prev_action.zero_()
    else:
      # the caller's state is its own copy (two drivers with the same number of environments
      # share this PolicyEngine but not their recurrent state): always load it
      pol.prev_action.copy_(torch.from_numpy(np.asarray(state['action'], np.float32)))
      pol.prev_deter.copy_(state['deter'])
      pol.prev_stoch.copy_(state['stoch'])
    pol.first.copy_(torch.from_numpy(first))
    self._run(pol, 'policy', pol.launch)
    noise = self.expl_noise
    if mode == 'eval':
      action = pol.mean.cpu().numpy().copy()
      noise = self.eval_noise
    elif mode == 'explore' and self.expl_random:
      shape = (P,) + tuple(self.act_space.shape)
      if self.spec.discrete:
        action = np.eye(shape[-1], dtype=np.float32)[self._rng.integers(0, shape[-1], shape[:-1])]
      else:
        action = self._rng.uniform(-1, 1, shape).astype(np.float32)
    else:
      action = pol.action.cpu().numpy().copy()
    if noise and self.spec.discrete:                              # A/tfutils.py:85-93
      probs = noise / action.shape[-1] + (1 - noise) * action
      draw = (probs.cumsum(-1) > self._rng.uniform(size=action.shape[:-1] + (1,))).argmax(-1)
      action = np.eye(action.shape[-1], dtype=np.float32)[draw]
    elif noise:
      action = np.clip(action + noise * self._rng.standard_normal(action.shape), -1, 1)
    action = action.astype(np.float32).reshape((P,) + tuple(self.act_space.shape))
    state = {'deter': pol.prev_deter.clone(), 'stoch': pol.prev_stoch.clone(), 'action': action}
    return {'action': action}, state

  # -- checkpointing (embodied/core/checkpoint.py:20-24,41-60) ----------------
  def save(self):
    eng = self._main_engine()
    self.be.sync()
    data = {f'param/{k}': v for k, v in eng.get_params().items()}
    for gname in ('model', 'actor', 'critic'):
      g = eng.groups[gname]
      for n, _ in g.order:
        data[f'adam_m/{n}'] = g.view(g.m, n).cpu().numpy().copy()
        data[f'adam_v/{n}'] = g.view(g.v, n).cpu().numpy().copy()
    data['ctrl/counters'] = eng.counters.cpu().numpy().copy()
    data['ctrl/wmkl_scale'] = eng.wmkl_scale.cpu().numpy().copy()
    data['ctrl/actent_scale'] = eng.actent_scale.cpu().numpy().copy()
    data['ctrl/norm_state'] = eng.norm_state.cpu().numpy().copy()
    count = int(sum(v.size for k, v in data.items() if k.startswith('param/')))
    print(f'Saving agent with {len(data)} tensors and {count} parameters.')
    return data

  def load(self, data):
    eng = self._main_engine()
    eng.load_params({k[len('param/'):]: v for k, v in data.items() if k.startswith('param/')})
    dev = self.be.device
    for gname in ('model', 'actor', 'critic'):
      g = eng.groups[gname]
      for n, _ in g.order:
        g.view(g.m, n).copy_(torch.as_tensor(data[f'adam_m/{n}']).to(dev))
        g.view(g.v, n).copy_(torch.as_tensor(data[f'adam_v/{n}']).to(dev))
    eng.counters.copy_(torch.as_tensor(data['ctrl/counters']).to(dev))
    eng.wmkl_scale.copy_(torch.as_tensor(data['ctrl/wmkl_scale']).to(dev))
    eng.actent_scale.copy_(torch.as_tensor(data['ctrl/actent_scale']).to(dev))
    eng.norm_state.copy_(torch.as_tensor(data['ctrl/norm_state']).to(dev))
    self._slow_updates = int(data['ctrl/counters'][_engine.C_SLOW])
    self.be.sync()
