"""-m gpu: the embodied.Agent surface (train/report/policy/save/load) on the
CUDA path, CUDA-graph replay vs eager launches, in-kernel Philox noise, and
size-independent properties at the full a1 size."""
import numpy as np
import pytest
import torch

from daydreamer_b200.config import make_config
from tests import util

pytestmark = pytest.mark.gpu


class Space:
  def __init__(self, shape, dtype=np.float32, discrete=False):
    self.shape, self.dtype, self.discrete = tuple(shape), np.dtype(dtype), discrete


def spaces(obs_dim=16, act_dim=12):
  obs = {'vector': Space((obs_dim,)), 'image': Space((64, 64, 3), np.uint8),
         'reward': Space(()), 'is_first': Space((), bool), 'is_last': Space((), bool),
         'is_terminal': Space((), bool)}
  return obs, {'action': Space((act_dim,)), 'reset': Space((), bool, True)}


def small_config(**kw):
  base = {'rssm.deter': 64, 'rssm.units': 64, 'rssm.stoch': 8, 'rssm.classes': 8,
          r'.*\.layers': 2, r'.*\.units': 64, 'encoder.mlp_layers': 2, 'encoder.mlp_units': 64,
          'decoder.mlp_layers': 2, 'decoder.mlp_units': 64, 'imag_horizon': 4,
          'batch_size': 4, 'replay_chunk': 8}
  base.update(kw)
  return make_config('a1', **base)


class FakeSpec:
  def __init__(self, obs_dim, act_dim):
    self.obs_dims, self.act_dim = {'vector': obs_dim}, act_dim


def test_graph_replay_equals_eager():
  from daydreamer_b200.agent import Agent
  obs, act = spaces()
  cfg = small_config()
  a1 = Agent(obs, act, None, cfg)
  a2 = Agent(obs, act, None, cfg)
  a2.use_graph = False
  s1 = s2 = None
  for it in range(4):
    batch = util.make_batch(FakeSpec(16, 12), 4, 8, seed=it)
    _, s1, m1 = a1.train(batch, s1)
    _, s2, m2 = a2.train(batch, s2)
    assert set(m1) == set(m2)
    for k in m1:
      if np.isnan(m1[k]):
        assert np.isnan(m2[k])
      else:
        assert abs(m1[k] - m2[k]) <= 1e-4 * max(1.0, abs(m2[k])), (it, k, m1[k], m2[k])
  assert m1['model_grad_steps'] == 4 and m1['actor_grad_steps'] == 4


def test_surface_policy_report_save_load():
  from daydreamer_b200.agent import Agent
  obs, act = spaces()
  cfg = small_config()
  agent = Agent(obs, act, None, cfg)
  batch = util.make_batch(FakeSpec(16, 12), 4, 8, seed=0)
  state = None
  for _ in range(3):
    outs, state, mets = agent.train(batch, state)
  assert outs == {}
  assert all(np.isscalar(v) or np.ndim(v) == 0 for v in mets.values())
  rep = agent.report(batch)
  assert 'model_loss_mean' in rep and np.isfinite(rep['model_loss_mean'])
  # report must not advance the optimizer
  _, _, mets2 = agent.train(batch, state)
  assert mets2['model_grad_steps'] == 4
  # policy: 2 envs, three steps, all modes
  ob = {'vector': np.random.randn(2, 16).astype(np.float32),
        'image': np.zeros((2, 64, 64, 3), np.uint8), 'reward': np.zeros(2, np.float32),
        'is_first': np.array([True, True]), 'is_last': np.zeros(2, bool),
        'is_terminal': np.zeros(2, bool)}
  pst = None
  for mode in ('train', 'explore', 'eval'):
    out, pst = agent.policy(ob, pst, mode)
    assert out['action'].shape == (2, 12) and out['action'].dtype == np.float32
    assert np.isfinite(out['action']).all()
    ob['is_first'] = np.array([False, False])
  # eval mode is the distribution mean: |a| <= 1 (tanh)
  assert np.abs(out['action']).max() <= 1.0
  # checkpoint round trip through a fresh agent
  data = agent.save()
  import pickle
  data = pickle.loads(pickle.dumps(data))
  other = Agent(obs, act, None, cfg, seed=99)
  other.load(data)
  p1, p2 = agent._main.get_params(), other._main.get_params()
  for k in p1:
    assert np.array_equal(p1[k], p2[k]), k
  assert np.array_equal(agent._main.counters.cpu().numpy(), other._main.counters.cpu().numpy())


def test_rejects_unbuilt_configs_loudly():
  from daydreamer_b200.agent import Agent
  obs, act = spaces()
  with pytest.raises(NotImplementedError):
    Agent(obs, act, None, make_config('a1', critic_type='qfunction'))
  with pytest.raises(NotImplementedError):
    Agent(obs, act, None, make_config('a1', expl_behavior='Explore'))
  with pytest.raises(NotImplementedError):
    Agent(obs, act, None, make_config('a1', **{'encoder.cnn_keys': 'image', 'encoder.cnn': 'resnet',
                                               'decoder.cnn_keys': 'image'}))


def test_a1_full_size_properties():
  """Size-independent properties at B=16, L=64, H=15 (BASELINE.json configs[1])."""
  from daydreamer_b200.agent import Agent
  obs, act = spaces()
  cfg = make_config('a1', **{'rssm.deter': 512, 'rssm.units': 512, 'batch_size': 16,
                            'replay_chunk': 64})
  agent = Agent(obs, act, None, cfg)
  B, L, H = 16, 64, 15
  state = None
  losses = []
  batch = util.make_batch(FakeSpec(16, 12), B, L, seed=0)
  for it in range(6):
    _, state, m = agent.train(batch, state if it else None)
    assert all(np.isfinite(v) or np.isnan(v) for v in m.values())
    losses.append(m['model_loss'])
  assert losses[-1] < losses[0], losses          # same batch: the model fits it
  eng = agent._main
  S, C, D = 32, 32, 512
  st = eng.tfeat[:, :, D:].reshape(H + 1, B * L, S, C)
  assert torch.all(st.sum(-1) == 1.0) and torch.all((st == 0) | (st == 1))   # one-hot
  lg = eng.tlogit[1:].reshape(H, B * L, S, C)
  assert torch.allclose(torch.logsumexp(lg, -1), torch.zeros_like(lg[..., 0]), atol=1e-4)
  w, c = eng.r_weight, eng.r_cont
  assert torch.allclose(w[0], c[0])
  assert torch.allclose(w[1:], w[:-1] * c[1:] * 0.995, rtol=1e-5, atol=1e-7)
  # lambda-return recursion holds on the device buffers
  R, V, r = eng.r_ret, eng.r_value, eng.r_reward
  d = 0.995 * c[1:]
  nxt = torch.cat([R[1:], V[H:]], 0)
  assert torch.allclose(R, r + d * (0.05 * V[1:] + 0.95 * nxt), rtol=1e-4, atol=1e-5)
  assert int(eng.counters[0]) == 6 and int(eng.counters[4]) == 6


@pytest.mark.parametrize('in_graph', [False, True])
def test_graph_capture_around_collectives(in_graph, monkeypatch):
  """Data-parallel launch path on one GPU.  DD_GRAPH_NCCL=0: with a collective hook installed the
  step is captured as graph segments split at every all-reduce and the collectives run eagerly
  between the replays; default: the collectives are captured INSIDE one graph (torch.distributed
  NCCL calls are capturable).  With an identity collective both must equal the single-graph agent."""
  from daydreamer_b200.agent import Agent
  monkeypatch.setenv('DD_GRAPH_NCCL', '1' if in_graph else '0')
  obs, act = spaces()
  cfg = small_config()
  a1 = Agent(obs, act, None, cfg)
  a2 = Agent(obs, act, None, cfg)
  calls = []
  a2._allreduce = lambda t: calls.append(t.numel())
  s1 = s2 = None
  for it in range(5):
    batch = util.make_batch(FakeSpec(16, 12), 4, 8, seed=it)
    if it == 0:
      _, s2, m2 = a2.train(batch, s2)
      a2._main.allreduce = a2._allreduce        # install the hook once the engine exists
    else:
      _, s2, m2 = a2.train(batch, s2)
    _, s1, m1 = a1.train(batch, s1)
    for k in m1:
      if not np.isnan(m1[k]):
        assert abs(m1[k] - m2[k]) <= 1e-4 * max(1.0, abs(m1[k])), (it, k, m1[k], m2[k])
  segs = a2._graphs[(id(a2._main), 'train')]
  kinds = [k for k, _ in segs]
  # collectives per step: kl statistic, model grads, the three normalizer rows (one call), entropy
  # statistic, critic grads, actor grads
  if in_graph:
    assert kinds == ['graph'], kinds
  else:
    assert kinds.count('allreduce') == 6 and kinds.count('graph') >= 6, kinds
    assert len(calls) == 6 * 5                   # 5 steps x 6 collectives (eager, capture replay, replays)


def xarm_like_spaces(act_dim=6):
  """xarm_dummy observation / action spaces (E/envs/robot_interface.py:395-410, :604-608)."""
  obs = {'image': Space((64, 64, 3), np.uint8), 'depth': Space((64, 64, 1), np.uint8),
         'cartesian_position': Space((6,)), 'joint_positions': Space((7,)),
         'gripper_pos': Space((1,)), 'gripper_side': Space((3,)), 'grasped_side': Space((3,)),
         'reward': Space(()), 'is_first': Space((), bool), 'is_last': Space((), bool),
         'is_terminal': Space((), bool)}
  return obs, {'action': Space((act_dim,), np.float32, True), 'reset': Space((), bool, True)}


def xarm_like_batch(B, L, A, seed):
  rng = np.random.default_rng(seed)
  f = np.zeros((B, L), bool)
  f[:, 0] = True
  b = {'image': rng.integers(0, 256, (B, L, 64, 64, 3)).astype(np.uint8),
       'depth': rng.integers(0, 256, (B, L, 64, 64, 1)).astype(np.uint8),
       'action': np.eye(A, dtype=np.float32)[rng.integers(0, A, (B, L))],
       'reward': rng.standard_normal((B, L)).astype(np.float32), 'is_first': f,
       'is_terminal': rng.uniform(size=(B, L)) < 0.05}
  for k, d in (('cartesian_position', 6), ('joint_positions', 7), ('gripper_pos', 1),
               ('gripper_side', 3), ('grasped_side', 3)):
    b[k] = rng.standard_normal((B, L, d)).astype(np.float32)
  b['is_last'] = b['is_terminal'].copy()
  return b


def test_xarm_image_config_surface():
  """BASELINE.json configs[0] shape (xarm: image+depth CNN, proprio MLP, 6-way one-hot
  actions, reinforce actor) through the Agent surface: graph replay == eager, losses
  finite and decreasing on a repeated batch, policy returns one-hot actions."""
  from daydreamer_b200.agent import Agent
  obs, act = xarm_like_spaces()
  cfg = make_config('xarm', **{
      'rssm.deter': 64, 'rssm.units': 64, 'rssm.stoch': 8, 'rssm.classes': 8,
      r'.*\.layers': 2, r'.*\.units': 64, 'encoder.mlp_layers': 2, 'encoder.mlp_units': 64,
      'decoder.mlp_layers': 2, 'decoder.mlp_units': 64, r'.*\.cnn_depth': 8, 'imag_horizon': 3,
      'batch_size': 2, 'replay_chunk': 7})
  a1 = Agent(obs, act, None, cfg)
  a2 = Agent(obs, act, None, cfg)
  a2.use_graph = False
  assert a1.spec.discrete and list(a1.spec.cnn_keys) == ['image', 'depth']
  assert a1.spec.obs_total == 20 and a1.spec.embed == 2 * 2 * 64 + 64
  s1 = s2 = None
  first = None
  for it in range(4):
    batch = xarm_like_batch(2, 7, 6, seed=0)
    _, s1, m1 = a1.train(batch, s1)
    _, s2, m2 = a2.train(batch, s2)
    s1 = s2 = None
    for k in m1:
      if np.isnan(m1[k]):
        assert np.isnan(m2[k]), k
      else:
        assert abs(m1[k] - m2[k]) <= 2e-4 * max(1.0, abs(m2[k])), (it, k, m1[k], m2[k])
    first = first or m1
  for k in ('image_loss_mean', 'depth_loss_mean', 'model_loss'):
    assert np.isfinite(m1[k]) and m1[k] < first[k], (k, first[k], m1[k])
  rep = a1.report(batch)
  assert np.isfinite(rep['image_loss_mean'])
  # image summaries: 5 context + 2 open-loop frames, [truth | model | error] on the height axis
  assert rep['openl_image'].shape == (7, 3 * 64, 2 * 64, 3)
  assert rep['openl_depth'].shape == (7, 3 * 64, 2 * 64, 1)
  assert rep['task_imag_image'].shape == (4, 64, 2 * 64, 3)
  truth = batch['image'][:2].astype(np.float32) / 255.0
  assert np.array_equal(rep['openl_image'][:, :64], truth.transpose(1, 2, 0, 3, 4).reshape(7, 64, 128, 3))
  for k in ('openl_image', 'openl_depth', 'task_imag_image', 'task_imag_depth'):
    assert np.isfinite(rep[k]).all() and rep[k].min() >= 0.0 and rep[k].max() <= 1.0, k
  assert all(np.ndim(v) in (0, 2, 3, 4) for v in rep.values())       # E/core/logger.py:30-33
  ob = {k: v[:, 0] for k, v in xarm_like_batch(3, 1, 6, seed=5).items() if k != 'action'}
  state = None
  for mode in ('train', 'eval', 'explore'):
    if mode == 'explore':
      continue
    out, state = a1.policy(ob, state, mode=mode)
    a = out['action']
    assert a.shape == (3, 6) and np.all(a.sum(-1) == 1.0) and set(np.unique(a)) <= {0.0, 1.0}


def test_ur5_full_size_image_step_properties():
  """BASELINE.json configs[2] at full size (ur5: 64x64x3 image + 19-dim proprio, B=32 L=50
  H=15, cnn_depth 64, deter 512) -- too large for the CPU oracle, so size-independent
  properties: identical images at both ends of the 1600-row batch give identical encoder
  rows and identical decoder pixels (catches 32-bit index overflow in the 1.5M-row
  im2col GEMMs), losses are finite, and a repeated batch lowers the image loss."""
  from daydreamer_b200.agent import Agent
  obs, act = xarm_like_spaces()
  obs.pop('depth')
  obs['joint_positions'] = Space((6,))
  cfg = make_config('ur5', batch_size=32, replay_chunk=50)
  agent = Agent(obs, act, None, cfg)
  assert agent.spec.embed == 2560 and agent.spec.obs_total == 19
  B, L = 32, 50
  batch = xarm_like_batch(B, L, 6, seed=1)
  batch.pop('depth')
  batch['joint_positions'] = batch['joint_positions'][..., :6]
  batch['image'][B - 1, L - 1] = batch['image'][0, 0]
  first = None
  for it in range(3):
    _, _, mets = agent.train(batch, None)
    first = first or mets
    assert all(np.isfinite(mets[k]) for k in ('model_loss', 'image_loss_mean', 'actor_loss',
                                               'extr_critic_loss', 'model_grad_norm'))
  assert mets['image_loss_mean'] < first['image_loss_mean']
  eng = agent._main
  N = B * L
  emb = eng.enc_cnn.H[-1].view(N, -1)
  assert torch.equal(emb[0], emb[N - 1]), 'encoder rows of identical images differ'
  assert float(emb.abs().max()) > 0
  # decoder: force two identical feature rows and compare their reconstructions
  eng.feat[N - 1].copy_(eng.feat[0])
  out = eng.dec_cnn.fwd(eng.feat)
  torch.cuda.synchronize()
  assert torch.equal(out[0], out[N - 1]), 'decoder pixels of identical features differ'
  assert agent.be.launch_count() > 0
