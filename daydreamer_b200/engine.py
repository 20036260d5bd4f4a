"""World-model learner engine: the whole `Agent.train` step of the reference
(/root/reference/embodied/agents/dreamerv2plus/agent.py:67-93) laid out as a
static sequence of kernel launches with hand-derived backward passes.

The engine owns every device buffer (parameters, Adam state, controllers,
saved activations) and only ever calls `self.be.<primitive>(...)`; the product
backend (backend.CudaBackend) turns each call into one sm_100a kernel launch
through the C-ABI in include/dd_kernels.h.  Shapes are fixed at construction
(B sequences x L steps, horizon H) so the launch sequence can be captured into
a CUDA graph and replayed.

Internal layout: time-major everywhere.  Row n = l*B + b of any [N, .] buffer
is sequence b at step l; trajectories are [H+1, N, .].

Reference map (A/ = embodied/agents/dreamerv2plus):
  observe scan ........ A/nets.py:66-117   (obs_step), :119-160 (img_step, GRU)
  KL balance .......... A/nets.py:178-183
  world-model loss .... A/agent.py:165-212
  optimizer ........... A/tfutils.py:180-302
  imagination ......... A/agent.py:234-261
  critic / lambda ..... A/agent.py:398-454
  actor loss .......... A/agent.py:326-381
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import program as PG


# ----------------------------------------------------------------------------
# Spec (sizes + hyper-parameters, resolved from the embodied.Config)
# ----------------------------------------------------------------------------

@dataclasses.dataclass
class Spec:
  obs_dims: Dict[str, int]
  act_dim: int = 12
  discrete: bool = False
  deter: int = 512
  units: int = 512
  stoch: int = 32
  classes: int = 32
  prior_layers: int = 3
  rssm_unimix: float = 0.01
  enc_layers: int = 4
  enc_units: int = 512
  head_layers: int = 4
  head_units: int = 512
  dec_layers: int = 0                 # decoder.mlp_layers / mlp_units when they differ (0 = as heads)
  dec_units: int = 0
  # decoder vector keys when they differ from the encoder's (None = the same keys).  The reference's
  # default `mlp_keys: '.*'` feeds the encoder every rank-0/1 key (reward, is_terminal, counters)
  # but decodes only rank-1 observation keys (A/nets.py:192-201, :241-250).
  dec_dims: Optional[Dict[str, int]] = None
  horizon: int = 15
  discount: float = 0.995
  lam: float = 0.95
  kl_balance: float = 0.8
  wmkl_target: float = 3.5
  wmkl_min: float = 1e-3
  wmkl_max: float = 1.0
  actent_target: float = 0.5
  actent_min: float = 1e-5
  actent_max: float = 1.0
  ctrl_vel: float = 0.1
  ctrl_thres: float = 0.1
  actor_minstd: float = 0.1
  actor_maxstd: float = 1.0
  actor_unimix: float = 0.1
  lr: float = 1e-4
  adam_eps: float = 1e-6
  clip: float = 100.0
  wd: float = 1e-2
  beta1: float = 0.9
  beta2: float = 0.999
  slow_every: int = 100
  advnorm_decay: float = 0.99
  advnorm_max: float = 1e8
  # image keys -> (H, W, C), channel-concatenated for the CNN (A/nets.py:218-222)
  cnn_keys: Dict[str, Tuple[int, int, int]] = dataclasses.field(default_factory=dict)
  cnn_depth: int = 64
  enc_kernels: Tuple[int, ...] = (4, 4, 4, 4)
  dec_kernels: Tuple[int, ...] = (5, 5, 6, 6)

  @property
  def sc(self):
    return self.stoch * self.classes

  def trunk(self, prefix):
    """(layers, units) of the MLP trunk called `prefix`."""
    if prefix == 'dec':
      return (self.dec_layers or self.head_layers, self.dec_units or self.head_units)
    if prefix == 'enc':
      return (self.enc_layers, self.enc_units)
    return (self.head_layers, self.head_units)

  @property
  def img_shape(self):
    if not self.cnn_keys:
      return None
    shapes = [tuple(v) for v in self.cnn_keys.values()]
    if any(x[:2] != shapes[0][:2] for x in shapes):
      raise ValueError(f'image keys must share height/width: {shapes}')
    return shapes[0][:2] + (sum(x[2] for x in shapes),)

  @property
  def enc_conv_shapes(self):
    """[(Hin, Win, Cin, k, Hout, Wout, Cout)] per layer of ImageEncoderSimple: stride 2,
    VALID, depth doubling (A/nets.py:298-305)."""
    out = []
    if not self.cnn_keys:
      return out
    h, w, c = self.img_shape
    depth = self.cnn_depth
    for k in self.enc_kernels:
      ho, wo = (h - k) // 2 + 1, (w - k) // 2 + 1
      if ho < 1 or wo < 1:
        raise ValueError(f'image {self.img_shape} too small for encoder kernels {self.enc_kernels}')
      out.append((h, w, c, k, ho, wo, depth))
      h, w, c, depth = ho, wo, depth, depth * 2
    return out

  @property
  def dec_conv_shapes(self):
    """Same tuples for ImageDecoderSimple: transposed convs, stride 2, VALID, out =
    2 in + k - 2, from a 1x1 map of the features (A/nets.py:316-327, :529-533)."""
    out = []
    if not self.cnn_keys:
      return out
    h, w, c = 1, 1, self.feat
    depth = self.cnn_depth * 2 ** (len(self.dec_kernels) - 2)
    for i, k in enumerate(self.dec_kernels):
      last = i == len(self.dec_kernels) - 1
      co = self.img_shape[2] if last else depth
      ho, wo = h * 2 + k - 2, w * 2 + k - 2
      out.append((h, w, c, k, ho, wo, co))
      h, w, c, depth = ho, wo, co, depth // 2
    if (h, w, c) != tuple(self.img_shape):       # the reference asserts this (A/nets.py:326)
      raise ValueError(f'decoder kernels {self.dec_kernels} produce {(h, w, c)}, '
                       f'not the image shape {self.img_shape}')
    return out

  @property
  def cnn_out(self):
    sh = self.enc_conv_shapes
    return sh[-1][4] * sh[-1][5] * sh[-1][6] if sh else 0

  @property
  def embed(self):
    """Encoder output width: [cnn features | mlp features] (A/nets.py:220-231)."""
    return self.cnn_out + (self.enc_units if self.obs_dims else 0)

  @property
  def feat(self):
    return self.deter + self.sc

  @property
  def obs_total(self):
    return int(sum(self.obs_dims.values()))

  @property
  def dec_vec(self):
    """Vector keys the decoder MLP reconstructs."""
    return self.obs_dims if self.dec_dims is None else self.dec_dims

  @property
  def dec_total(self):
    return int(sum(self.dec_vec.values()))


def param_shapes(spec: Spec) -> Dict[str, Tuple[int, ...]]:
  """Every trainable tensor (+ slow critic), reference names and shapes
  (Linear has no bias when LayerNorm follows, A/nets.py:563)."""
  s = {}

  def dense(prefix, fin, fout):
    s[f'{prefix}/kernel'] = (fin, fout)
    s[f'{prefix}/norm/scale'] = (fout,)
    s[f'{prefix}/norm/bias'] = (fout,)

  def out(prefix, fin, fout):
    s[f'{prefix}/kernel'] = (fin, fout)
    s[f'{prefix}/bias'] = (fout,)

  def mlp(prefix, fin, layers, units):
    for i in range(layers):
      dense(f'{prefix}/dense{i}', fin if i == 0 else units, units)

  D, U, SC, A = spec.deter, spec.units, spec.sc, spec.act_dim
  for i, (_, _, ci, k, _, _, co) in enumerate(spec.enc_conv_shapes):
    s[f'enc/conv{i}/kernel'] = (k, k, ci, co)                 # HWIO (A/nets.py:541)
    s[f'enc/conv{i}/bias'] = (co,)
    s[f'enc/conv{i}/norm/scale'] = (co,)
    s[f'enc/conv{i}/norm/bias'] = (co,)
  if spec.obs_dims:
    mlp('enc', spec.obs_total, spec.enc_layers, spec.enc_units)
  dense('rssm/img_in', SC + A, U)
  dense('rssm/gru_out', D + U, 3 * D)
  for i in range(spec.prior_layers):
    dense(f'rssm/img_out_{i}', D if i == 0 else U, U)
  out('rssm/img_stats', U, SC)
  dense('rssm/obs_out', D + spec.embed, U)
  out('rssm/obs_stats', U, SC)
  s['rssm/initial_deter'] = (D,)
  dshapes = spec.dec_conv_shapes
  for i, (_, _, ci, k, _, _, co) in enumerate(dshapes):
    name = 'dec/conv_out' if i == len(dshapes) - 1 else f'dec/conv{i}'
    s[f'{name}/kernel'] = (k, k, co, ci)                      # [k,k,out,in] (A/nets.py:523)
    s[f'{name}/bias'] = (co,)
    if i < len(dshapes) - 1:
      s[f'{name}/norm/scale'] = (co,)
      s[f'{name}/norm/bias'] = (co,)
  if spec.dec_vec:
    mlp('dec', spec.feat, *spec.trunk('dec'))
    out('dec/dist_out/out', spec.trunk('dec')[1], spec.dec_total)
  for name in ('reward', 'cont'):
    mlp(name, spec.feat, spec.head_layers, spec.head_units)
    out(f'{name}/dist_out/out', spec.head_units, 1)
  mlp('actor', spec.feat, spec.head_layers, spec.head_units)
  out('actor/dist_out/out', spec.head_units, A)
  if not spec.discrete:
    out('actor/dist_out/std', spec.head_units, A)
  for name in ('critic', 'slow_critic'):
    mlp(name, spec.feat, spec.head_layers, spec.head_units)
    out(f'{name}/dist_out/out', spec.head_units, 1)
  return s


GROUP_PREFIXES = {
    'model': ('enc/', 'rssm/', 'dec/', 'reward/', 'cont/'),
    'actor': ('actor/',),
    'critic': ('critic/',),
    'slow_critic': ('slow_critic/',),
}


def init_params(spec: Spec, seed=1) -> Dict[str, np.ndarray]:
  """Reference initialisers: kernels U(+-sqrt(3*outscale/mean(fan_in,fan_out)))
  (A/nets.py:568-569), DistLayer 'out' outscale 0.1 (A/nets.py:448-450),
  LayerNorm scale 1 / bias 0, biases 0, initial_deter 0."""
  rng = np.random.default_rng(seed)
  out = {}
  for name, shape in param_shapes(spec).items():
    if name.endswith('/kernel') and len(shape) == 4:
      # Conv2D U(+-sqrt(3/mean(in,out))) (A/nets.py:542-543); transposed
      # U(+-sqrt(3/(k*k*mean(out,in)))) (A/nets.py:524-525)
      spatial = shape[0] * shape[1] if name.startswith('dec/') else 1
      limit = math.sqrt(3.0 / (spatial * (shape[2] + shape[3]) / 2.0))
      out[name] = rng.uniform(-limit, limit, shape).astype(np.float32)
    elif name.endswith('/kernel'):
      outscale = 0.1 if name.endswith('dist_out/out/kernel') else 1.0
      limit = math.sqrt(3.0 * outscale / ((shape[0] + shape[1]) / 2.0))
      out[name] = rng.uniform(-limit, limit, shape).astype(np.float32)
    elif name.endswith('norm/scale'):
      out[name] = np.ones(shape, np.float32)
    else:
      out[name] = np.zeros(shape, np.float32)
  return out


def _ceil4(n):
  return (n + 3) // 4 * 4


def _strided(buf, off, rows, cols, ld):
  return buf[off:off + (rows - 1) * ld + cols].as_strided((rows, cols), (ld, 1))


def _mview_ok(t):
  """A 2-D view the layer machine's TMA descriptors accept: unit column stride, row stride a
  multiple of 16 bytes, 16-byte aligned base."""
  return (t is not None and t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1) and
          t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0)


class ParamGroup:
  """One optimizer group as flat fp32 buffers [params | grads | m | v]; tensors
  matching the weight-decay pattern 'kernel' (A/tfutils.py:292) come first so
  the fused Adam kernel decays a prefix.

  For the layer machine (csrc/dd_machine.cu) the group also keeps GEMM-operand copies,
  refreshed by prep() after every update: `p_lo` (tf32 remainder of every parameter, same
  offsets) and, per 2-D kernel [K, N], the transposed copy [N, ld >= K] with its remainder (B
  operand of x @ W) and, when N % 4 != 0, a row-padded copy [K, ld >= N] (B operand of dY @ W^T)."""

  def __init__(self, be, names_shapes: List[Tuple[str, Tuple[int, ...]]], opt=True):
    decayed = [(n, s) for n, s in names_shapes if 'kernel' in n]
    rest = [(n, s) for n, s in names_shapes if 'kernel' not in n]
    self.order = decayed + rest
    self.offsets = {}
    off = 0
    self.n_decay = 0
    for i, (n, s) in enumerate(self.order):
      off = (off + 3) // 4 * 4        # 16-byte alignment for vectorised kernels
      self.offsets[n] = (off, s)
      off += int(np.prod(s))
      if i == len(decayed) - 1:
        self.n_decay = (off + 3) // 4 * 4
    self.size = (off + 3) // 4 * 4
    self.machine = hasattr(be, 'program_build')
    # (p, p_lo) as the two planes of one buffer: the machine fetches a weight tile and its
    # remainder with a single 3-D TMA box
    self._pp = be.zeros(2 if self.machine else 1, self.size)
    self.p = self._pp[0]
    self.g = be.zeros(self.size) if opt else None
    self.m = be.zeros(self.size) if opt else None
    self.v = be.zeros(self.size) if opt else None
    self.sumsq = be.zeros(1, dtype=torch.float64) if opt else None
    self.t_off, self.b_off, self.jobs = {}, {}, []
    if self.machine:
      off = 0
      for n, s in self.order:
        if not (n.endswith('/kernel') and len(s) == 2):
          continue
        K, N = s
        src = self.offsets[n][0]
        ldt = _ceil4(K)
        self.t_off[n] = (off, ldt)
        self.jobs.append((src, K, N, off, ldt, 0))
        off += _ceil4(N * ldt)
        if N % 4:
          ldb = _ceil4(N)
          self.b_off[n] = (off, ldb)
          self.jobs.append((src, K, N, off, ldb, 1))
          off += K * ldb
      self.p_lo = self._pp[1]
      self._ptt = be.zeros(2, max(_ceil4(off), 4))
      self.pt, self.pt_lo = self._ptt[0], self._ptt[1]

  def view(self, flat, name):
    off, shape = self.offsets[name]
    return flat[off:off + int(np.prod(shape))].view(*shape)

  def P(self, name):
    return self.view(self.p, name)

  def G(self, name):
    return self.view(self.g, name)

  def prep(self, be):
    """Refresh the machine's operand copies from the current parameters."""
    if not self.machine:
      return
    be.split_lo(self.p, self.p_lo)
    if self.jobs:
      be.transpose_lo(self.p, self.pt, self.pt_lo, self.jobs)

  def PT(self, name):
    """(W^T, its remainder) as [N, K] views: the B operand of x @ W."""
    off, ldt = self.t_off[name]
    K, N = self.offsets[name][1]
    return _strided(self.pt, off, N, K, ldt), _strided(self.pt_lo, off, N, K, ldt)

  def PT2(self, a, b):
    """Transposed copies of two kernels with the same input width stacked as one [Na + Nb, K]
    operand (they are adjacent in `pt` by construction: consecutive kernels, K % 4 == 0)."""
    (oa, la), (ob, lb) = self.t_off[a], self.t_off[b]
    Ka, Na = self.offsets[a][1]
    Kb, Nb = self.offsets[b][1]
    assert Ka == Kb and la == lb and ob == oa + Na * la, 'kernels are not adjacent'
    return _strided(self.pt, oa, Na + Nb, Ka, la), _strided(self.pt_lo, oa, Na + Nb, Ka, la)

  def cat(self, a, b):
    """Two 1-D parameters that lie back to back in the flat buffer, as one vector."""
    (oa, sa), (ob, sb) = self.offsets[a], self.offsets[b]
    assert ob == oa + int(np.prod(sa)), 'parameters are not adjacent'
    return self.p[oa:ob + int(np.prod(sb))]

  def PB(self, name):
    """(W, its remainder) as [K, N] views with a 16-byte row stride: the B operand of dY @ W^T."""
    K, N = self.offsets[name][1]
    if name in self.b_off:
      off, ldb = self.b_off[name]
      return _strided(self.pt, off, K, N, ldb), _strided(self.pt_lo, off, K, N, ldb)
    return self.P(name), self.view(self.p_lo, name)


# ----------------------------------------------------------------------------
# MLP trunk with saved activations
# ----------------------------------------------------------------------------

class Trunk:
  """`layers` x (Linear(no bias) -> LayerNorm(eps 1e-3) -> elu) over a fixed
  number of rows, keeping pre-norm Z, output H and (mean, rstd) per layer for
  the backward pass (A/nets.py:408-416, :566-582).

  Inputs are lists of (X, first weight row[, X_lo]): a third element (a tensor, or None for an
  operand that is exact in tf32 such as the one-hot latent) makes the trunk eligible for the layer
  machine, which runs the whole stack -- and optional output layers -- as ONE program
  (csrc/dd_machine.cu); otherwise every layer is a dd_gemm + dd_ln_fwd pair."""

  def __init__(self, eng, group, prefix, layers, units, rows, param_prefix=None, names=None):
    be = eng.be
    self.eng, self.group, self.prefix = eng, group, prefix
    self.pp = param_prefix or prefix
    self.layers, self.units, self.rows = layers, units, rows
    self.Z = [be.zeros(rows, units) for _ in range(layers)]
    self.H = [be.zeros(rows, units) for _ in range(layers)]
    self.st = [be.zeros(rows, 2) for _ in range(layers)]
    self._names_fn = names
    self.machine = bool(getattr(eng, 'machine', False)) and units % 256 == 0 and rows >= 1
    self.Hlo = [be.zeros(rows, units) for _ in range(layers)] if self.machine else None
    self._progs = {}

  def _names(self, i):
    if self._names_fn is not None:
      return self._names_fn(i)
    b = f'{self.pp}/dense{i}'
    return f'{b}/kernel', f'{b}/norm/scale', f'{b}/norm/bias'

  def _machine_xs(self, xs):
    return (self.machine and 1 <= len(xs) <= 2 and all(len(x) == 3 for x in xs) and
            all(_mview_ok(x[0]) and (x[2] is None or _mview_ok(x[2])) for x in xs))

  def fwd(self, xs, r0, r1, group=None, heads=()):
    """xs: list of (X[r1-r0, k], first weight row[, X_lo]).  heads: (kernel name, bias name, out
    view) output layers applied to the last hidden layer.  Returns H[-1][r0:r1]."""
    be, g = self.eng.be, group or self.group
    if (self._machine_xs(xs) and r1 - r0 <= self.eng.machine_max_rows and
        all(_mview_ok(h[2]) or h[2].shape[1] < 4 for h in heads)):
      key = ('f', r0, r1, id(g), tuple((x[0].data_ptr(), x[1]) for x in xs),
             tuple((h[0], h[2].data_ptr()) for h in heads))
      prog = self._progs.get(key)
      if prog is None:
        layers = []
        for i in range(self.layers):
          kn, sn, bn = self._names(i)
          wt, wtlo = g.PT(kn)
          segs = ([PG.Seg(X, lo, w0) for X, w0, lo in xs] if i == 0 else
                  [PG.Seg(self.H[i - 1][r0:r1], self.Hlo[i - 1][r0:r1], 0)])
          layers.append(PG.Layer(segs, wt, wtlo, self.units, PG.EPI_LN, act=1, gamma=g.P(sn),
                                 beta=g.P(bn), z=self.Z[i][r0:r1], st=self.st[i][r0:r1],
                                 out=self.H[i][r0:r1], out_lo=self.Hlo[i][r0:r1]))
        last = PG.Seg(self.H[-1][r0:r1], self.Hlo[-1][r0:r1], 0)
        for kn, bn, out in heads:
          wt, wtlo = g.PT(kn)
          layers.append(PG.Layer([last], wt, wtlo, out.shape[1], PG.EPI_PLAIN, bias=g.P(bn), out=out))
        prog = self._progs[key] = be.program_build(layers, r1 - r0)
      be.program_run(prog)
      return self.H[-1][r0:r1]
    for i in range(self.layers):
      kn, sn, bn = self._names(i)
      W = g.P(kn)
      Z = self.Z[i][r0:r1]
      if i == 0:
        beta = 0.0
        for x in xs:
          X, w0 = x[0], x[1]
          be.gemm(X, W[w0:w0 + X.shape[1]], Z, beta=beta)
          beta = 1.0
      else:
        be.gemm(self.H[i - 1][r0:r1], W, Z)
      be.ln_fwd(Z, g.P(sn), g.P(bn), self.H[i][r0:r1], self.st[i][r0:r1], 1)
    for kn, bn, out in heads:
      be.gemm(self.H[-1][r0:r1], g.P(kn), out, bias=g.P(bn))
    return self.H[-1][r0:r1]

  def bwd(self, dH, r0, r1, xs, dxs, wgrad, group=None):
    """dH: grad wrt H[-1][r0:r1] (clobbered).  xs as in fwd (needed for the
    first layer's weight grad); dxs: list of (dX, first weight row, beta) input
    grads to produce.  wgrad: accumulate parameter grads into the group."""
    eng, be, g = self.eng, self.eng.be, group or self.group
    n = r1 - r0
    if (self.machine and eng.machine_bwd and n <= eng.machine_max_rows and _mview_ok(dH) and
        all(_mview_ok(d[0]) and PG.eligible_width(d[0].shape[1], PG.EPI_PLAIN) for d in dxs)):
      key = ('b', r0, r1, id(g), dH.data_ptr(), bool(wgrad),
             tuple((d[0].data_ptr(), d[1], float(d[2])) for d in dxs))
      prog = self._progs.get(key)
      dZ = [eng.dz_buf(i, n, self.units) for i in range(self.layers)]
      if prog is None:
        layers = []
        for i in reversed(range(self.layers)):
          kn, sn, bn = self._names(i)
          common = dict(act=1, gamma=g.P(sn), beta=g.P(bn), z=self.Z[i][r0:r1], st=self.st[i][r0:r1],
                        out=dZ[i][0], out_lo=dZ[i][1], dgamma=g.G(sn) if wgrad else None,
                        dbeta=g.G(bn) if wgrad else None)
          if i == self.layers - 1:
            layers.append(PG.Layer([], None, None, self.units, PG.EPI_LNBWD, add=dH, **common))
          else:
            w, wlo = g.PB(self._names(i + 1)[0])
            layers.append(PG.Layer([PG.Seg(dZ[i + 1][0], dZ[i + 1][1], 0)], w, wlo, self.units,
                                   PG.EPI_LNBWD, **common))
        w, wlo = g.PB(self._names(0)[0])
        for dX, w0, beta in dxs:
          layers.append(PG.Layer([PG.Seg(dZ[0][0], dZ[0][1], 0)], w, wlo, dX.shape[1], PG.EPI_PLAIN,
                                 wn0=w0, add=dX if beta else None, out=dX))
        prog = self._progs[key] = be.program_build(layers, n)
      be.program_run(prog)
      if wgrad:
        for i in range(self.layers):
          GW = g.G(self._names(i)[0])
          if i == 0:
            for x in xs:
              X, w0 = x[0], x[1]
              be.gemm(X, dZ[0][0], GW[w0:w0 + X.shape[1]], ta=True, beta=1.0)
          else:
            be.gemm(self.H[i - 1][r0:r1], dZ[i][0], GW, ta=True, beta=1.0)
      return
    tA = dH
    for i in reversed(range(self.layers)):
      kn, sn, bn = self._names(i)
      W = g.P(kn)
      tB = eng.tmp_other(tA, n, self.units)
      be.ln_bwd(tA, self.Z[i][r0:r1], self.st[i][r0:r1], g.P(sn), g.P(bn), tB,
                g.G(sn) if wgrad else None, g.G(bn) if wgrad else None, 1)
      if wgrad:
        GW = g.G(kn)
        if i == 0:
          for x in xs:
            X, w0 = x[0], x[1]
            be.gemm(X, tB, GW[w0:w0 + X.shape[1]], ta=True, beta=1.0)
        else:
          be.gemm(self.H[i - 1][r0:r1], tB, GW, ta=True, beta=1.0)
      if i == 0:
        for dX, w0, beta in dxs:
          be.gemm(tB, W[w0:w0 + dX.shape[1]], dX, tb=True, beta=beta)
      else:
        tA = eng.tmp_other(tB, n, self.units)
        be.gemm(tB, W, tA, tb=True)


# ----------------------------------------------------------------------------
# Image encoder / decoder (conv as im2col | col2im around the dense GEMM)
# ----------------------------------------------------------------------------

class ConvEncoder:
  """ImageEncoderSimple (A/nets.py:291-305): per layer conv(k, stride 2, VALID) +
  bias -> LayerNorm over channels -> elu, on `rows` NHWC images.  Keeps the
  im2col matrix, pre-norm Z, output H and LN stats of every layer for backward."""

  def __init__(self, eng, group, rows, train=True):
    be, s = eng.be, eng.spec
    self.eng, self.group, self.rows = eng, group, rows
    self.shapes = s.enc_conv_shapes
    z = be.zeros
    self.col = [z(rows * ho * wo, k * k * ci) for (_, _, ci, k, ho, wo, _) in self.shapes]
    self.Z = [z(rows * ho * wo, co) for (_, _, _, _, ho, wo, co) in self.shapes]
    self.H = [z(rows, ho, wo, co) for (_, _, _, _, ho, wo, co) in self.shapes]
    self.st = [z(rows * ho * wo, 2) for (_, _, _, _, ho, wo, _) in self.shapes]
    if train:
      self.dZ = [z(rows * ho * wo, co) for (_, _, _, _, ho, wo, co) in self.shapes]
      self.dH = [z(rows, ho, wo, co) for (_, _, _, _, ho, wo, co) in self.shapes[:-1]]
      self.dcol = z(max([c.numel() for c in self.col[1:]] + [1]))

  def _names(self, i):
    b = f'enc/conv{i}'
    return f'{b}/kernel', f'{b}/bias', f'{b}/norm/scale', f'{b}/norm/bias'

  def fwd(self, img):
    be, g = self.eng.be, self.group
    x = img
    for i, (_, _, ci, k, ho, wo, co) in enumerate(self.shapes):
      kn, bn, sn, nbn = self._names(i)
      be.im2col(x, k, 2, self.col[i])
      be.gemm(self.col[i], g.P(kn).view(k * k * ci, co), self.Z[i], bias=g.P(bn))
      be.ln_fwd(self.Z[i], g.P(sn), g.P(nbn), self.H[i].view(-1, co), self.st[i], 1)
      x = self.H[i]
    return self.H[-1].view(self.rows, -1)

  def bwd(self, dOut):
    """dOut [rows, cnn_out]: grad wrt the flattened output.  Accumulates parameter grads."""
    be, g = self.eng.be, self.group
    dH = dOut.view(-1, self.shapes[-1][6])
    for i in reversed(range(len(self.shapes))):
      _, _, ci, k, ho, wo, co = self.shapes[i]
      kn, bn, sn, nbn = self._names(i)
      W = g.P(kn).view(k * k * ci, co)
      dZ = self.dZ[i]
      be.ln_bwd(dH, self.Z[i], self.st[i], g.P(sn), g.P(nbn), dZ, g.G(sn), g.G(nbn), 1)
      be.colsum(dZ, g.G(bn), acc=True)
      be.gemm(self.col[i], dZ, g.G(kn).view(k * k * ci, co), ta=True, beta=1.0)
      if i > 0:
        dcol = self.dcol[:self.col[i].numel()].view(self.col[i].shape)
        be.gemm(dZ, W, dcol, tb=True)
        be.col2im(dcol, k, 2, None, self.dH[i - 1])
        dH = self.dH[i - 1].view(-1, ci)


class ConvDecoder:
  """ImageDecoderSimple (A/nets.py:308-327): features as a 1x1 map -> transposed
  convs (stride 2, VALID) with LayerNorm+elu, last layer plain; returns the
  PRE-sigmoid image (the sigmoid lives in the loss / report kernels)."""

  def __init__(self, eng, group, rows, train=True):
    be, s = eng.be, eng.spec
    self.eng, self.group, self.rows = eng, group, rows
    self.shapes = s.dec_conv_shapes
    z = be.zeros
    self.Z = [z(rows, ho, wo, co) for (_, _, _, _, ho, wo, co) in self.shapes]
    self.H = [z(rows, ho, wo, co) for (_, _, _, _, ho, wo, co) in self.shapes[:-1]]
    self.st = [z(rows * ho * wo, 2) for (_, _, _, _, ho, wo, _) in self.shapes[:-1]]
    self.ycol = z(max(rows * hi * wi * k * k * co for (hi, wi, _, k, _, _, co) in self.shapes))
    if train:
      self.dZ = [z(rows, ho, wo, co) for (_, _, _, _, ho, wo, co) in self.shapes[:-1]]
      self.dH = [z(rows, ho, wo, co) for (_, _, _, _, ho, wo, co) in self.shapes[:-1]]

  def _names(self, i):
    b = 'dec/conv_out' if i == len(self.shapes) - 1 else f'dec/conv{i}'
    return f'{b}/kernel', f'{b}/bias', f'{b}/norm/scale', f'{b}/norm/bias'

  def fwd(self, feat):
    be, g = self.eng.be, self.group
    x = feat
    last = len(self.shapes) - 1
    for i, (hi, wi, ci, k, ho, wo, co) in enumerate(self.shapes):
      kn, bn, sn, nbn = self._names(i)
      ycol = self.ycol[:self.rows * hi * wi * k * k * co].view(self.rows * hi * wi, k * k * co)
      be.gemm(x, g.P(kn).view(k * k * co, ci), ycol, tb=True)
      be.col2im(ycol, k, 2, g.P(bn), self.Z[i])
      if i < last:
        be.ln_fwd(self.Z[i].view(-1, co), g.P(sn), g.P(nbn), self.H[i].view(-1, co),
                  self.st[i], 1)
        x = self.H[i].view(-1, co)
    return self.Z[-1]

  def bwd(self, dZ, feat, dfeat, beta):
    """dZ: grad wrt the pre-sigmoid image; dfeat (+)= grad wrt the features."""
    be, g = self.eng.be, self.group
    for i in reversed(range(len(self.shapes))):
      hi, wi, ci, k, ho, wo, co = self.shapes[i]
      kn, bn, sn, nbn = self._names(i)
      W = g.P(kn).view(k * k * co, ci)
      be.colsum(dZ.view(-1, co), g.G(bn), acc=True)
      dycol = self.ycol[:self.rows * hi * wi * k * k * co].view(self.rows * hi * wi, k * k * co)
      be.im2col(dZ, k, 2, dycol)
      x = feat if i == 0 else self.H[i - 1].view(-1, ci)
      be.gemm(dycol, x, g.G(kn).view(k * k * co, ci), ta=True, beta=1.0)
      if i == 0:
        be.gemm(dycol, W, dfeat, beta=beta)
      else:
        psn, pnbn = self._names(i - 1)[2:]
        dH = self.dH[i - 1].view(-1, ci)
        be.gemm(dycol, W, dH)
        be.ln_bwd(dH, self.Z[i - 1].view(-1, ci), self.st[i - 1], g.P(psn), g.P(pnbn),
                  self.dZ[i - 1].view(-1, ci), g.G(psn), g.G(pnbn), 1)
        dZ = self.dZ[i - 1]


# ----------------------------------------------------------------------------
# Engine
# ----------------------------------------------------------------------------

C_MODEL, C_CRITIC, C_ACTOR, C_SLOW, C_RNG = 0, 1, 2, 3, 4
STREAM_OBS, STREAM_IMG, STREAM_ACT = 1, 2, 3


class Engine:

  def __init__(self, spec: Spec, be, B: int, L: int, seed: int = 0,
               world: int = 1, allreduce=None, share: 'Engine' = None):
    self.spec, self.be = spec, be
    self.B, self.L, self.N, self.H = B, L, B * L, spec.horizon
    self.seed = seed
    self.world = world
    self.allreduce = allreduce            # callable(flat tensor) -> sum over ranks
    # layer machine (csrc/dd_machine.cu): trunks, the prior stack and the rollout's img_in / GRU
    # layers run as fused programs wherever their shapes tile across its 8-CTA clusters
    self.machine = hasattr(be, 'program_build') and os.environ.get('DD_MACHINE', '1') != '0'
    # measured (profiles/r02_machine_bench.txt): a chain of 4 layers is 1.7x faster than dd_gemm + dd_ln
    # launches at 1K rows and still 5 % faster at 16K rows (two interleaved tiles per cluster); beyond
    # that it is unmeasured and the per-layer kernels stay
    self.machine_max_rows = int(os.environ.get('DD_MACHINE_MAXROWS', '32768'))
    # reverse chains (LayerNorm-backward epilogue) of the same trunks; DD_MACHINE_BWD=0 keeps
    # dd_gemm + dd_ln_bwd (r02 A/B at a1 size: 0.15 ms per step slower)
    self.machine_bwd = os.environ.get('DD_MACHINE_BWD', '1') != '0'
    s, N, H = spec, self.N, self.H
    D, U, SC, A, E = s.deter, s.units, s.sc, s.act_dim, s.enc_units
    z = be.zeros
    shapes = param_shapes(spec)
    if share is not None:
      # a second shape (e.g. report() on another batch) on the same learner state
      self.groups = share.groups
      self.counters, self.wmkl_scale = share.counters, share.wmkl_scale
      self.actent_scale, self.norm_state = share.actent_scale, share.norm_state
    else:
      self.groups = {
          g: ParamGroup(be, [(n, sh) for n, sh in shapes.items() if n.startswith(pre)],
                        opt=(g != 'slow_critic'))
          for g, pre in GROUP_PREFIXES.items()}
      # controllers / counters (device-resident so the step is one static graph)
      self.counters = z(8, dtype=torch.int64)
      self.wmkl_scale = z(1)
      self.actent_scale = z(1 if spec.discrete else A)   # A/agent.py:305-307
      self.norm_state = z(3, 3, dtype=torch.float64)     # adv / ret / score
      self.reset_state()
    gm = self.groups['model']
    self.adv = z(2)                                    # (mean, scale) transform
    self.scratch2 = z(2)

    # inputs (time-major)
    self.obs = z(N, s.obs_total)
    # decoder targets: the encoder's vector input unless the decoder has its own key set
    self.dec_tgt = self.obs if s.dec_dims is None else z(N, s.dec_total)
    self.img = z(N, *s.img_shape) if s.cnn_keys else None     # f32 in [0,1], channel-concat
    self.action = z(N, A)
    self.first = z(N)
    self.reward = z(N, 1)
    self.cont = z(N, 1)
    self.prev_deter = z(B, D)
    self.prev_stoch = z(B, SC)
    self.first0 = z(B)            # is_first at t=0, or all ones when no state is carried

    # noise
    self.g_obs = z(self.L, B, SC)
    self.g_img = z(H, N, SC)
    self.eps_act = z(H + 1, N, A)

    # shared temporaries for backward ping-pong (largest row count x widest)
    R = (H + 1) * N
    wide = max(3 * D, SC, U, s.head_units, s.trunk('dec')[1], E)
    self._tmp = [z(R * wide) for _ in range(3)]
    # second pool for the launches of a concurrent branch (launch_step: critic next to actor)
    self.overlap = os.environ.get('DD_OVERLAP', '1') != '0'
    self._tmp_alt = [z(R * max(s.head_units, 1)) for _ in range(3)] if self.overlap else None
    # machine: per-layer dZ (+ remainder) of the trunk being differentiated (trunks run one after
    # another, so one pool serves them all), remainders of the GEMM-operand activations
    umax = max(U, s.head_units, s.trunk('dec')[1], E)
    nl = max(s.head_layers, s.enc_layers, s.prior_layers, s.trunk('dec')[0])
    self._dzp = [z(R * umax) for _ in range(2 * nl)] if self.machine else []
    self._dzp_alt = [z(R * umax) for _ in range(2 * nl)] if self.machine_bwd and self.overlap else []
    self.obs_lo = z(N, s.obs_total) if self.machine and s.obs_total else None
    self.feat_lo = z(N, D) if self.machine else None

    # -- world model buffers -------------------------------------------------
    self.enc = Trunk(self, gm, 'enc', s.enc_layers, E, N) if s.obs_dims else None
    self.enc_cnn = ConvEncoder(self, gm, N) if s.cnn_keys else None
    self.dec_cnn = ConvDecoder(self, gm, N) if s.cnn_keys else None
    self.d_img = z(N, *s.img_shape) if s.cnn_keys else None
    self.d_cnn = z(N, s.cnn_out) if s.cnn_keys else None
    self.init_deter = z(1, D)
    self.init_trunk = Trunk(self, gm, 'rssm/initp', s.prior_layers, U, 1, names=self._prior_names)
    self.init_zs = z(1, SC)
    self.init_logit = z(1, SC)
    self.init_stoch = z(1, SC)
    self.m_deter = z(N, D)          # masked previous deter per step
    self.m_stoch = z(N, SC)
    self.m_act = z(N, A)
    self.z_in = z(N, U)
    self.x_in = z(N, U)
    self.st_in = z(N, 2)
    self.z_gru = z(N, 3 * D)
    self.y_gru = z(N, 3 * D)
    self.st_gru = z(N, 2)
    self.feat = z(N, D + SC)        # [post.deter | post.stoch]
    self.z_obs = z(N, U)
    self.h_obs = z(N, U)
    self.st_obs = z(N, 2)
    self.zs_tmp = z(N, SC)
    self.post_logit = z(N, SC)
    self.post_idx = z(N, s.stoch, dtype=torch.int32)
    self.prior = Trunk(self, gm, 'rssm/priorp', s.prior_layers, U, N, names=self._prior_names)
    self.prior_logit = z(N, SC)
    self.heads = {k: Trunk(self, gm, k, *s.trunk(k), N)
                  for k in (('dec',) if s.dec_vec else ()) + ('reward', 'cont')}
    self.out_dec = z(N, s.dec_total)
    self.out_rew = z(N, 1)
    self.out_con = z(N, 1)
    self.d_dec = z(N, s.dec_total)
    self.d_rew = z(N, 1)
    self.d_con = z(N, 1)
    self.kl_row = z(N)
    self.ent_post = z(N)
    self.ent_prior = z(N)
    self.loss_rows = {k: z(N) for k in list(s.cnn_keys) + list(s.dec_vec) +
                      ['reward', 'cont', 'total']}
    self.kl_sum = z(1)
    self.dfeat = z(N, D + SC)
    self.dpost_logit = z(N, SC)
    self.dprior_logit = z(N, SC)
    self.dzs = z(N, SC)
    self.dz_obs = z(N, U)
    self.dz_gru = z(N, 3 * D)
    self.dz_in = z(N, U)
    self.d_embed = z(N, E)
    self.d_initdeter = z(1, D)
    # fused observe scan (persistent cooperative kernel) workspace
    self.fused_scan = True
    self.scan_partial = z(148 * 16 * 2)
    self.scan_barrier = z(4, dtype=torch.int32)
    # fused BPTT kernel: pre-LN gradients of every step (LN parameter grads are batched after it)
    self.fused_bptt = True
    self.b_dh = z(N, U)
    self.b_dy = z(N, 3 * D)
    self.b_dx = z(N, U)

    # -- imagination buffers ------------------------------------------------
    ga, gc, gs = self.groups['actor'], self.groups['critic'], self.groups['slow_critic']
    T = H + 1
    self.tfeat = z(T, N, D + SC)
    self.tfeat_lo = z(T, N, D) if self.machine else None       # deter part; the latent is one-hot
    self.taction_lo = z(T, N, A) if self.machine else None
    self.i_xin_lo = z(H, N, U) if self.machine else None
    # reverse-rollout program scratch: d(raw prior logits), d(pre-norm GRU gates), d(pre-norm img_in)
    if self.machine:
      self.r_dzs, self.r_dzs_lo = z(N, SC), z(N, SC)
      self.r_dzg, self.r_dzg_lo = z(N, 3 * D), z(N, 3 * D)
      self.r_dzi, self.r_dzi_lo = z(N, U), z(N, U)
    self.tlogit = z(T, N, SC)
    self.taction = z(T, N, A)
    self.tmean = z(T, N, A)
    self.tstd = z(T, N, A)
    self.talogit = z(T, N, A) if s.discrete else None     # unimixed action logits
    self.a_mraw = z(T * N, A)
    self.a_sraw = z(T * N, A)
    self.actor = Trunk(self, ga, 'actor', s.head_layers, s.head_units, T * N)
    self.i_zin = z(H, N, U)
    self.i_xin = z(H, N, U)
    self.i_stin = z(H, N, 2)
    self.i_zgru = z(H, N, 3 * D)
    self.i_ygru = z(H, N, 3 * D)
    self.i_stgru = z(H, N, 2)
    self.i_prior = Trunk(self, gm, 'rssm/priorp', s.prior_layers, U, H * N, names=self._prior_names)
    self.i_zs = z(N, SC)
    self.t_reward = Trunk(self, gm, 'reward', s.head_layers, s.head_units, T * N)
    self.t_cont = Trunk(self, gm, 'cont', s.head_layers, s.head_units, T * N)
    self.t_slow = Trunk(self, gs, 'critic', s.head_layers, s.head_units, T * N,
                        param_prefix='slow_critic')
    self.t_critic = Trunk(self, gc, 'critic', s.head_layers, s.head_units, H * N)
    self.rho = z(T, N)
    self.vraw = z(T, N)
    self.kraw = z(T, N)
    self.vpred = z(H * N, 1)
    self.d_vpred = z(H * N, 1)
    self.r_reward = z(H, N)
    self.r_value = z(T, N)
    self.r_cont = z(T, N)
    self.r_weight = z(T, N)
    self.r_ret = z(H, N)
    self.cont0 = z(N)
    self.closs_row = z(H * N)
    self.score = z(H, N)
    self.vsym = z(H * N, 1)
    self.ent = z(H * N, 1 if s.discrete else A)
    self.ent_sum = z(1 if s.discrete else A)
    self.aloss_row = z(H * N)
    self.d_rho = z(T, N)
    self.d_vraw = z(T, N)
    self.d_kraw = z(T, N)
    self.dtfeat = z(T, N, D + SC)
    self.da = z(H * N, A)
    self.d_mraw = z(H * N, A)
    self.d_sraw = z(H * N, A)

    # metrics: rows of 8 doubles written by `stats`/`balance`
    self.mnames = [
        'kl', 'ent_post', 'ent_prior', 'reward_loss', 'cont_loss', 'total_loss',
        *[f'loss_{k}' for k in list(s.cnn_keys) + list(s.dec_vec)], 'bal_reward', 'bal_cont',
        'imag_reward', 'imag_critic', 'imag_return', 'closs', 'score', 'adv_ctrl',
        'ret_ctrl', 'score_ctrl', 'ent', 'aloss']
    self.mstat = z(len(self.mnames), 8, dtype=torch.float64)
    self.midx = {k: i for i, k in enumerate(self.mnames)}

  # -- small helpers --------------------------------------------------------
  def tmp(self, i, rows, cols):
    return self._tmp[i][:rows * cols].view(rows, cols)

  def tmp_other(self, cur, rows, cols):
    """Ping-pong between shared temporaries 0 and 1 (2 is free for callers)."""
    i = 1 if cur.data_ptr() == self._tmp[0].data_ptr() else 0
    return self.tmp(i, rows, cols)

  def dz_buf(self, i, rows, cols):
    """(dZ, dZ_lo) of layer i of the trunk currently being differentiated."""
    return (self._dzp[2 * i][:rows * cols].view(rows, cols),
            self._dzp[2 * i + 1][:rows * cols].view(rows, cols))

  def prep_weights(self, names=None):
    """Refresh the machine's (w, w_lo) operand copies of the given groups from the parameters."""
    for n in (names or self.groups):
      self.groups[n].prep(self.be)

  def fsegs(self, X, Xlo):
    """[deter | stoch] feature rows as the two A segments of a first trunk layer: the deter part
    with its remainder, the one-hot latent exact.  Without the machine: the plain (X, 0) input."""
    D = self.spec.deter
    if not self.machine or Xlo is None:
      return [(X, 0)]
    return [(X[:, :D], 0, Xlo), (X[:, D:], D, None)]

  def reset_state(self):
    be = self.be
    be.fill_(self.wmkl_scale, 1.0)
    be.fill_(self.actent_scale, 1.0)
    be.fill_(self.norm_state, 0.0)
    self.counters.copy_(torch.tensor([0, 0, 0, -1, 0, 0, 0, 0], dtype=torch.int64))

  def load_params(self, params: Dict[str, np.ndarray]):
    for g in self.groups.values():
      for n, _ in g.order:
        g.P(n).copy_(torch.as_tensor(np.asarray(params[n], np.float32)).to(g.p.device))
    self.prep_weights()

  def get_params(self) -> Dict[str, np.ndarray]:
    out = {}
    for g in self.groups.values():
      for n, _ in g.order:
        out[n] = g.P(n).detach().cpu().numpy().copy()
    return out

  def get_grads(self, group) -> Dict[str, np.ndarray]:
    g = self.groups[group]
    return {n: g.G(n).detach().cpu().numpy().copy() for n, _ in g.order}

  def P(self, name):
    for g, pre in GROUP_PREFIXES.items():
      if name.startswith(pre):
        return self.groups[g].P(name)
    raise KeyError(name)

  def _stat(self, X, key):
    self.be.stats(X, self.mstat[self.midx[key]])

  def _out_fwd(self, g, prefix, X, out):
    self.be.gemm(X, g.P(f'{prefix}/kernel'), out, bias=g.P(f'{prefix}/bias'))

  def _out_bwd(self, g, prefix, X, dOut, dX, wgrad):
    be = self.be
    if wgrad:
      be.gemm(X, dOut, g.G(f'{prefix}/kernel'), ta=True, beta=1.0)
      be.colsum(dOut, g.G(f'{prefix}/bias'), acc=True)
    be.gemm(dOut, g.P(f'{prefix}/kernel'), dX, tb=True)

  def _prior_stats_fwd(self, trunk, r0, r1, X, zs, G, logit, stoch, idx, mode, Xlo=None):
    """img_out_0..k -> img_stats -> unimix logits (+ draw)  (A/nets.py:132-137)."""
    s = self.spec
    xs = [(X, 0, Xlo)] if (self.machine and Xlo is not None) else [(X, 0)]
    trunk.fwd(xs, r0, r1, heads=[('rssm/img_stats/kernel', 'rssm/img_stats/bias', zs)])
    self.be.stats_fwd(zs, G, logit, stoch, idx, s.stoch, s.classes, s.rssm_unimix, mode)

  def _prior_names(self, i):
    b = f'rssm/img_out_{i}'
    return f'{b}/kernel', f'{b}/norm/scale', f'{b}/norm/bias'

  def _prior_trunk_bwd(self, trunk, r0, r1, X, dH, dX, beta, wgrad, Xlo=None):
    """Backward of img_out_* trunk; dH clobbered; dX (+)= grad wrt X."""
    xs = [(X, 0, Xlo)] if (self.machine and Xlo is not None) else [(X, 0)]
    trunk.bwd(dH, r0, r1, xs, [(dX, 0, beta)] if dX is not None else [], wgrad)

  # -- noise ------------------------------------------------------------------
  def draw_noise(self):
    be = self.be
    be.fill_gumbel(self.g_obs, self.seed, STREAM_OBS, self.counters, C_RNG)
    be.fill_gumbel(self.g_img, self.seed, STREAM_IMG, self.counters, C_RNG)
    if self.spec.discrete:       # Gumbel-max draw of the one-hot action
      be.fill_gumbel(self.eps_act, self.seed, STREAM_ACT, self.counters, C_RNG)
    else:
      be.fill_normal(self.eps_act, self.seed, STREAM_ACT, self.counters, C_RNG)
    be.counter_inc(self.counters, C_RNG)

  # ==========================================================================
  # 1. world model: forward
  # ==========================================================================
  def wm_forward(self):
    s, be, gm = self.spec, self.be, self.groups['model']
    B, L, N = self.B, self.L, self.N
    D, U, SC, A, E = s.deter, s.units, s.sc, s.act_dim, s.enc_units
    feat = self.feat

    # MultiEncoder: [cnn features | mlp features] (A/nets.py:212-232); their part of
    # obs_out's pre-activation is one GEMM per encoder output over all steps at once
    Wobs = gm.P('rssm/obs_out/kernel')
    self.embed_parts = []
    if s.cnn_keys:
      self.embed_parts.append(self.enc_cnn.fwd(self.img))
    if s.obs_dims:
      if self.machine and self.enc.machine and _mview_ok(self.obs):
        be.split_lo(self.obs, self.obs_lo)
        self.embed_parts.append(self.enc.fwd([(self.obs, 0, self.obs_lo)], 0, N))
      else:
        self.embed_parts.append(self.enc.fwd([(self.obs, 0)], 0, N))
    off, beta = D, 0.0
    for part in self.embed_parts:
      be.gemm(part, Wobs[off:off + part.shape[1]], self.z_obs, beta=beta)
      off, beta = off + part.shape[1], 1.0

    # learned initial state (A/nets.py:54-62): deter0 = tanh(v), stoch0 = mode
    be.tanh_fwd(gm.P('rssm/initial_deter').view(1, D), self.init_deter)
    self._prior_stats_fwd(self.init_trunk, 0, 1, self.init_deter, self.init_zs,
                          None, self.init_logit, self.init_stoch, None, 1)

    if not self._observe_fused():
      self._observe_steps()
    if self.machine:
      be.split_lo(feat[:, :D], self.feat_lo)

    # prior logits for every step at once (the prior draw is unused by observe)
    self._prior_stats_fwd(self.prior, 0, N, feat[:, :D], self.zs_tmp, None,
                          self.prior_logit, None, None, 2, Xlo=self.feat_lo)

    # heads on the posterior features (A/agent.py:170-176)
    xs = self.fsegs(feat, self.feat_lo)
    if s.cnn_keys:
      self.dec_cnn.fwd(feat)
    for name, out in ((('dec', self.out_dec),) if s.dec_vec else ()) + (
        ('reward', self.out_rew), ('cont', self.out_con)):
      self.heads[name].fwd(xs, 0, N, heads=[(f'{name}/dist_out/out/kernel',
                                             f'{name}/dist_out/out/bias', out)])

  def _observe_fused(self):
    """All L obs_steps in one persistent cooperative kernel (csrc/dd_scan.cu) when
    the backend has it and the shape is eligible; False -> per-step launches."""
    be, s, gm = self.be, self.spec, self.groups['model']
    if not (self.fused_scan and hasattr(be, 'observe_fwd')):
      return False
    P = gm.P
    dims = dict(B=self.B, L=self.L, D=s.deter, U=s.units, S=s.stoch, C=s.classes, A=s.act_dim)
    tensors = dict(
        action=self.action, first=self.first, first0=self.first0, prev_deter=self.prev_deter,
        prev_stoch=self.prev_stoch, init_deter=self.init_deter, init_stoch=self.init_stoch,
        g_obs=self.g_obs, Win=P('rssm/img_in/kernel'), ln_in_s=P('rssm/img_in/norm/scale'),
        ln_in_b=P('rssm/img_in/norm/bias'), Wgru=P('rssm/gru_out/kernel'),
        ln_gru_s=P('rssm/gru_out/norm/scale'), ln_gru_b=P('rssm/gru_out/norm/bias'),
        Wobs=P('rssm/obs_out/kernel'), ln_obs_s=P('rssm/obs_out/norm/scale'),
        ln_obs_b=P('rssm/obs_out/norm/bias'), Wst=P('rssm/obs_stats/kernel'),
        bst=P('rssm/obs_stats/bias'), m_deter=self.m_deter, m_stoch=self.m_stoch,
        m_act=self.m_act, z_in=self.z_in, x_in=self.x_in, st_in=self.st_in, z_gru=self.z_gru,
        y_gru=self.y_gru, st_gru=self.st_gru, feat=self.feat, z_obs=self.z_obs,
        h_obs=self.h_obs, st_obs=self.st_obs, post_logit=self.post_logit,
        post_idx=self.post_idx, partial=self.scan_partial, barrier=self.scan_barrier,
        timing=getattr(self, 'scan_timing', None))
    if not hasattr(self, '_scan_ok'):
      self._scan_ok = (be.observe_fwd_grid(dims, s.rssm_unimix, tensors) or 0) > 0
    if not self._scan_ok:
      return False
    be.observe_fwd(dims, s.rssm_unimix, tensors)
    return True

  def _observe_steps(self):
    """The scan as ~12 launches per step (also the reference semantics of the fused kernel)."""
    s, be, gm = self.spec, self.be, self.groups['model']
    B, L = self.B, self.L
    D, SC = s.deter, s.sc
    feat = self.feat
    Win = gm.P('rssm/img_in/kernel')
    Wgru = gm.P('rssm/gru_out/kernel')
    Wobs = gm.P('rssm/obs_out/kernel')
    for t in range(L):
      r = slice(t * B, (t + 1) * B)
      if t == 0:
        pd, ps, first = self.prev_deter, self.prev_stoch, self.first0
      else:
        q = slice((t - 1) * B, t * B)
        pd, ps, first = feat[q, :D], feat[q, D:], self.first[r]
      be.mask_fwd(pd, ps, self.action[r], first, self.init_deter[0],
                  self.init_stoch[0], self.m_deter[r], self.m_stoch[r], self.m_act[r])
      # img_in
      be.gemm(self.m_stoch[r], Win[:SC], self.z_in[r])
      be.gemm(self.m_act[r], Win[SC:], self.z_in[r], beta=1.0)
      be.ln_fwd(self.z_in[r], gm.P('rssm/img_in/norm/scale'),
                gm.P('rssm/img_in/norm/bias'), self.x_in[r], self.st_in[r], 1)
      # GRU (A/nets.py:149-160)
      be.gemm(self.m_deter[r], Wgru[:D], self.z_gru[r])
      be.gemm(self.x_in[r], Wgru[D:], self.z_gru[r], beta=1.0)
      be.ln_fwd(self.z_gru[r], gm.P('rssm/gru_out/norm/scale'),
                gm.P('rssm/gru_out/norm/bias'), self.y_gru[r], self.st_gru[r], 0)
      be.gru_fwd(self.y_gru[r], self.m_deter[r], feat[r, :D])
      # posterior
      be.gemm(feat[r, :D], Wobs[:D], self.z_obs[r], beta=1.0)
      be.ln_fwd(self.z_obs[r], gm.P('rssm/obs_out/norm/scale'),
                gm.P('rssm/obs_out/norm/bias'), self.h_obs[r], self.st_obs[r], 1)
      self._out_fwd(gm, 'rssm/obs_stats', self.h_obs[r], self.zs_tmp[r])
      be.stats_fwd(self.zs_tmp[r], self.g_obs[t], self.post_logit[r], feat[r, D:],
                   self.post_idx[r], s.stoch, s.classes, s.rssm_unimix, 0)

  # ==========================================================================
  # 1b. world model: losses, backward, optimizer
  # ==========================================================================
  def wm_losses(self, update=True):
    s, be = self.spec, self.be
    N = self.N
    coef = 1.0 / (N * self.world)
    be.kl_fwd(self.post_logit, self.prior_logit, s.stoch, s.classes, s.kl_balance,
              self.kl_row, self.ent_post, self.ent_prior)
    if update:      # AutoAdapt.update precedes the use of the scale (A/tfutils.py:441)
      be.colsum(self.kl_row.view(N, 1), self.kl_sum)
      if self.allreduce is not None:
        self.allreduce(self.kl_sum)
      be.autoadapt(self.wmkl_scale, self.kl_sum, float(N * self.world), s.wmkl_target,
                   s.wmkl_min, s.wmkl_max, s.ctrl_vel, s.ctrl_thres, False)
    tot = self.loss_rows['total']
    be.fill_(tot, 0.0)
    off = 0
    for k, shp in s.cnn_keys.items():       # MSEDist(sigmoid(out), 3, 'sum')
      be.image_loss(self.dec_cnn.Z[-1], self.img, off, off + shp[2], coef, self.loss_rows[k],
                    self.d_img)
      be.add_(tot, self.loss_rows[k])
      off += shp[2]
    off = 0
    for k, d in s.dec_vec.items():
      be.head_loss(0, self.out_dec[:, off:off + d], self.dec_tgt[:, off:off + d], None,
                   coef, self.loss_rows[k], self.d_dec[:, off:off + d])
      be.add_(tot, self.loss_rows[k])
      off += d
    be.head_loss(1, self.out_rew, self.reward, None, coef, self.loss_rows['reward'], self.d_rew)
    be.head_loss(2, self.out_con, self.cont, None, coef, self.loss_rows['cont'], self.d_con)
    be.add_(tot, self.loss_rows['reward'])
    be.add_(tot, self.loss_rows['cont'])
    be.axpy_(tot, self.kl_row, self.wmkl_scale)
    be.kl_bwd(self.post_logit, self.prior_logit, s.stoch, s.classes, s.kl_balance,
              coef, self.wmkl_scale, self.dpost_logit, self.dprior_logit)
    # metrics
    self._stat(self.kl_row, 'kl')
    self._stat(self.ent_post, 'ent_post')
    self._stat(self.ent_prior, 'ent_prior')
    self._stat(self.loss_rows['reward'], 'reward_loss')
    self._stat(self.loss_rows['cont'], 'cont_loss')
    self._stat(tot, 'total_loss')
    for k in list(s.cnn_keys) + list(s.dec_vec):
      self._stat(self.loss_rows[k], f'loss_{k}')
    be.balance(1, self.out_rew, self.loss_rows['reward'], self.reward, 0.1,
               self.mstat[self.midx['bal_reward']])
    be.balance(2, self.out_con, self.loss_rows['cont'], self.cont, 0.5,
               self.mstat[self.midx['bal_cont']])

  def wm_backward(self):
    s, be, gm = self.spec, self.be, self.groups['model']
    B, L, N = self.B, self.L, self.N
    D, U, SC, A, E = s.deter, s.units, s.sc, s.act_dim, s.enc_units
    feat, dfeat = self.feat, self.dfeat
    be.fill_(gm.g, 0.0)
    be.fill_(self.d_initdeter, 0.0)
    hu = s.head_units

    # heads -> dfeat
    beta = 0.0
    if s.cnn_keys:
      self.dec_cnn.bwd(self.d_img, feat, dfeat, beta)
      beta = 1.0
    for name, dOut in (('dec', self.d_dec), ('reward', self.d_rew), ('cont', self.d_con)):
      if name not in self.heads:
        continue
      tr = self.heads[name]
      dH = self.tmp(0, N, tr.units)
      self._out_bwd(gm, f'{name}/dist_out/out', tr.H[-1], dOut, dH, True)
      tr.bwd(dH, 0, N, self.fsegs(feat, self.feat_lo), [(dfeat, 0, beta)], True)
      beta = 1.0

    # prior chain (all steps at once): dprior_logit -> ... -> dfeat[:, :D]
    be.stats_bwd(None, self.dprior_logit, self.prior_logit, self.dzs, s.stoch,
                 s.classes, s.rssm_unimix)
    dH = self.tmp(0, N, U)
    self._out_bwd(gm, 'rssm/img_stats', self.prior.H[-1], self.dzs, dH, True)
    self._prior_trunk_bwd(self.prior, 0, N, feat[:, :D], dH, dfeat[:, :D], 1.0, True,
                          Xlo=self.feat_lo)

    # reverse scan (BPTT)
    Win = gm.P('rssm/img_in/kernel')
    Wgru = gm.P('rssm/gru_out/kernel')
    Wobs = gm.P('rssm/obs_out/kernel')
    Wst = gm.P('rssm/obs_stats/kernel')
    for t in ([] if self._bptt_fused() else reversed(range(L))):
      r = slice(t * B, (t + 1) * B)
      dh = self.tmp(0, B, U)
      dd = self.tmp(1, B, D)
      ds = self.tmp(2, B, SC)
      # posterior
      be.stats_bwd(dfeat[r, D:], self.dpost_logit[r], self.post_logit[r], self.dzs[r],
                   s.stoch, s.classes, s.rssm_unimix)
      be.gemm(self.dzs[r], Wst, dh, tb=True)
      be.ln_bwd(dh, self.z_obs[r], self.st_obs[r], gm.P('rssm/obs_out/norm/scale'),
                gm.P('rssm/obs_out/norm/bias'), self.dz_obs[r],
                gm.G('rssm/obs_out/norm/scale'), gm.G('rssm/obs_out/norm/bias'), 1)
      be.gemm(self.dz_obs[r], Wobs[:D], dfeat[r, :D], tb=True, beta=1.0)
      # GRU
      dy = self.tmp(0, B, 3 * D)
      be.gru_bwd(dfeat[r, :D], self.y_gru[r], self.m_deter[r], dy, dd, False)
      be.ln_bwd(dy, self.z_gru[r], self.st_gru[r], gm.P('rssm/gru_out/norm/scale'),
                gm.P('rssm/gru_out/norm/bias'), self.dz_gru[r],
                gm.G('rssm/gru_out/norm/scale'), gm.G('rssm/gru_out/norm/bias'), 0)
      be.gemm(self.dz_gru[r], Wgru[:D], dd, tb=True, beta=1.0)
      dx = self.tmp(0, B, U)
      be.gemm(self.dz_gru[r], Wgru[D:], dx, tb=True)
      # img_in
      be.ln_bwd(dx, self.z_in[r], self.st_in[r], gm.P('rssm/img_in/norm/scale'),
                gm.P('rssm/img_in/norm/bias'), self.dz_in[r],
                gm.G('rssm/img_in/norm/scale'), gm.G('rssm/img_in/norm/bias'), 1)
      be.gemm(self.dz_in[r], Win[:SC], ds, tb=True)
      if t == 0:
        be.mask_bwd(dd, ds, self.first0, None, None, self.d_initdeter[0])
      else:
        q = slice((t - 1) * B, t * B)
        be.mask_bwd(dd, ds, self.first[r], dfeat[q, :D], dfeat[q, D:], self.d_initdeter[0])

    # weight grads of the scan layers, all steps at once
    be.gemm(self.h_obs, self.dzs, gm.G('rssm/obs_stats/kernel'), ta=True, beta=1.0)
    be.colsum(self.dzs, gm.G('rssm/obs_stats/bias'), acc=True)
    Gobs = gm.G('rssm/obs_out/kernel')
    be.gemm(feat[:, :D], self.dz_obs, Gobs[:D], ta=True, beta=1.0)
    off = D
    for part in self.embed_parts:
      be.gemm(part, self.dz_obs, Gobs[off:off + part.shape[1]], ta=True, beta=1.0)
      off += part.shape[1]
    Ggru = gm.G('rssm/gru_out/kernel')
    be.gemm(self.m_deter, self.dz_gru, Ggru[:D], ta=True, beta=1.0)
    be.gemm(self.x_in, self.dz_gru, Ggru[D:], ta=True, beta=1.0)
    Gin = gm.G('rssm/img_in/kernel')
    be.gemm(self.m_stoch, self.dz_in, Gin[:SC], ta=True, beta=1.0)
    be.gemm(self.m_act, self.dz_in, Gin[SC:], ta=True, beta=1.0)
    # initial_deter = tanh(v)
    be.tanh_bwd(self.d_initdeter, self.init_deter,
                gm.G('rssm/initial_deter').view(1, D), True)
    # encoders
    off = D
    if s.cnn_keys:
      be.gemm(self.dz_obs, Wobs[off:off + s.cnn_out], self.d_cnn, tb=True)
      self.enc_cnn.bwd(self.d_cnn)
      off += s.cnn_out
    if s.obs_dims:
      be.gemm(self.dz_obs, Wobs[off:], self.d_embed, tb=True)
      self.enc.bwd(self.d_embed, 0, N, [(self.obs, 0)], [], True)

  def _bptt_fused(self):
    """All L reverse steps in one persistent cooperative kernel (csrc/dd_scan.cu) when the
    backend has it and the shape is eligible; False -> per-step launches.  The LayerNorm
    parameter gradients of the three scan norms come from the saved pre-LN gradients with one
    batched ln_bwd each."""
    be, s, gm = self.be, self.spec, self.groups['model']
    if not (self.fused_bptt and hasattr(be, 'observe_bwd')):
      return False
    P = gm.P
    dims = dict(B=self.B, L=self.L, D=s.deter, U=s.units, S=s.stoch, C=s.classes)
    tensors = dict(
        first=self.first, first0=self.first0, post_logit=self.post_logit,
        dpost_logit=self.dpost_logit, z_obs=self.z_obs, st_obs=self.st_obs,
        ln_obs_s=P('rssm/obs_out/norm/scale'), ln_obs_b=P('rssm/obs_out/norm/bias'),
        y_gru=self.y_gru, z_gru=self.z_gru, st_gru=self.st_gru,
        ln_gru_s=P('rssm/gru_out/norm/scale'), ln_gru_b=P('rssm/gru_out/norm/bias'),
        m_deter=self.m_deter, z_in=self.z_in, st_in=self.st_in,
        ln_in_s=P('rssm/img_in/norm/scale'), ln_in_b=P('rssm/img_in/norm/bias'),
        Wst=P('rssm/obs_stats/kernel'), Wobs=P('rssm/obs_out/kernel'),
        Wgru=P('rssm/gru_out/kernel'), Win=P('rssm/img_in/kernel'), dfeat=self.dfeat,
        dzs=self.dzs, dz_obs=self.dz_obs, dz_gru=self.dz_gru, dz_in=self.dz_in, dh=self.b_dh,
        dy_gru=self.b_dy, dx=self.b_dx, d_initdeter=self.d_initdeter,
        partial=self.scan_partial, barrier=self.scan_barrier,
        timing=getattr(self, 'bptt_timing', None))
    if not hasattr(self, '_bptt_ok'):
      self._bptt_ok = (be.observe_bwd_grid(dims, s.rssm_unimix, tensors) or 0) > 0
    if not self._bptt_ok:
      return False
    be.observe_bwd(dims, s.rssm_unimix, tensors)
    N, U, D = self.N, s.units, s.deter
    G = gm.G
    for dH, Z, st, name, act, width in (
        (self.b_dh, self.z_obs, self.st_obs, 'rssm/obs_out', 1, U),
        (self.b_dy, self.z_gru, self.st_gru, 'rssm/gru_out', 0, 3 * D),
        (self.b_dx, self.z_in, self.st_in, 'rssm/img_in', 1, U)):
      be.ln_bwd(dH, Z, st, P(f'{name}/norm/scale'), P(f'{name}/norm/bias'), self.tmp(0, N, width),
                G(f'{name}/norm/scale'), G(f'{name}/norm/bias'), act)
    return True

  def opt_step(self, group, ci):
    s, be, g = self.spec, self.be, self.groups[group]
    if self.allreduce is not None:
      self.allreduce(g.g)
    be.sumsq(g.g, g.sumsq)
    be.adam(g.p, g.g, g.m, g.v, g.n_decay, s.lr, s.beta1, s.beta2, s.adam_eps,
            s.wd, s.clip, 1.0, g.sumsq, self.counters, ci)
    g.prep(be)

  # ==========================================================================
  # 2. imagination rollout (A/agent.py:234-261)
  # ==========================================================================
  def _actor_fwd(self, t):
    s, be, ga = self.spec, self.be, self.groups['actor']
    N = self.N
    r0, r1 = t * N, (t + 1) * N
    heads = [('actor/dist_out/out/kernel', 'actor/dist_out/out/bias', self.a_mraw[r0:r1])]
    if not s.discrete:
      heads.append(('actor/dist_out/std/kernel', 'actor/dist_out/std/bias', self.a_sraw[r0:r1]))
    self.actor.fwd(self.fsegs(self.tfeat[t], self.tfeat_lo[t] if self.machine else None), r0, r1,
                   heads=heads)
    if s.discrete:       # OneHotDist over unimix logits, straight-through sample (A/nets.py:480-491)
      be.stats_fwd(self.a_mraw[r0:r1], self.eps_act[t], self.talogit[t], self.taction[t],
                   None, 1, s.act_dim, s.actor_unimix, 0)
      return
    be.actor_sample(self.a_mraw[r0:r1], self.a_sraw[r0:r1], self.eps_act[t],
                    s.actor_minstd, s.actor_maxstd, self.taction[t],
                    self.tmean[t], self.tstd[t])
    if self.machine:
      be.split_lo(self.taction[t], self.taction_lo[t])

  def imagine(self, actions=None):
    """WorldModel.imagine (A/agent.py:234-261): the actor picks every action.  With
    `actions` [H, N, A] given it is RSSM.imagine (A/nets.py:78-86): an open-loop
    rollout of the prior along a recorded action sequence, no actor."""
    s, be, gm = self.spec, self.be, self.groups['model']
    N, H = self.N, self.H
    D, U, SC, A = s.deter, s.units, s.sc, s.act_dim
    tf = self.tfeat
    be.copy_(tf[0], self.feat)
    be.copy_(self.tlogit[0], self.post_logit)
    if self.machine:
      be.copy_(self.tfeat_lo[0], self.feat_lo)
    if actions is None and self._rollout_program():
      return
    Win = gm.P('rssm/img_in/kernel')
    Wgru = gm.P('rssm/gru_out/kernel')
    if actions is None:
      self._actor_fwd(0)
    for t in range(H):
      act = self.taction[t] if actions is None else actions[t]
      if not self._cell_program(t, act, actions is None):
        be.gemm(tf[t][:, D:], Win[:SC], self.i_zin[t])
        be.gemm(act, Win[SC:], self.i_zin[t], beta=1.0)
        be.ln_fwd(self.i_zin[t], gm.P('rssm/img_in/norm/scale'),
                  gm.P('rssm/img_in/norm/bias'), self.i_xin[t], self.i_stin[t], 1)
        be.gemm(tf[t][:, :D], Wgru[:D], self.i_zgru[t])
        be.gemm(self.i_xin[t], Wgru[D:], self.i_zgru[t], beta=1.0)
        be.ln_fwd(self.i_zgru[t], gm.P('rssm/gru_out/norm/scale'),
                  gm.P('rssm/gru_out/norm/bias'), self.i_ygru[t], self.i_stgru[t], 0)
      be.gru_fwd(self.i_ygru[t], tf[t][:, :D], tf[t + 1][:, :D])
      if self.machine:
        be.split_lo(tf[t + 1][:, :D], self.tfeat_lo[t + 1])
      self._prior_stats_fwd(self.i_prior, t * N, (t + 1) * N, tf[t + 1][:, :D],
                            self.i_zs, self.g_img[t], self.tlogit[t + 1],
                            tf[t + 1][:, D:], None, 0,
                            Xlo=self.tfeat_lo[t + 1] if self.machine else None)
      if actions is None:
        self._actor_fwd(t + 1)

  def _rollout_program(self):
    """The whole imagination rollout (A/agent.py:234-261) as ONE machine program: per step
    actor trunk + action head + sample, img_in, GRU (Linear + joint LayerNorm + gates), the prior
    stack and the categorical draw -- 11 layers per step, each 128-row tile of start states running
    through all H steps inside one persistent kernel.  False when a shape does not tile."""
    s, be = self.spec, self.be
    gm, ga = self.groups['model'], self.groups['actor']
    D, U, SC, A, N, H = s.deter, s.units, s.sc, s.act_dim, self.N, self.H
    if not hasattr(self, '_rollout'):
      ok = (self.machine and os.environ.get('DD_ROLLOUT', '1') != '0' and not s.discrete and
            s.classes == 32 and A <= 16 and A % 4 == 0 and D % 256 == 0 and U % 256 == 0 and
            s.head_units % 256 == 0 and SC % 256 == 0 and self.actor.machine and self.i_prior.machine)
      self._rollout = None
      if ok:
        tf, tlo = self.tfeat, self.tfeat_lo
        layers = []

        def actor_layers(t):
          r0, r1 = t * N, (t + 1) * N
          tr = self.actor
          for i in range(tr.layers):
            kn, sn, bn = tr._names(i)
            wt, wtlo = ga.PT(kn)
            segs = ([PG.Seg(tf[t][:, :D], tlo[t], 0), PG.Seg(tf[t][:, D:], None, D)] if i == 0 else
                    [PG.Seg(tr.H[i - 1][r0:r1], tr.Hlo[i - 1][r0:r1], 0)])
            layers.append(PG.Layer(segs, wt, wtlo, tr.units, PG.EPI_LN, act=1, gamma=ga.P(sn),
                                   beta=ga.P(bn), z=tr.Z[i][r0:r1], st=tr.st[i][r0:r1],
                                   out=tr.H[i][r0:r1], out_lo=tr.Hlo[i][r0:r1]))
          wt, wtlo = ga.PT2('actor/dist_out/out/kernel', 'actor/dist_out/std/kernel')
          layers.append(PG.Layer([PG.Seg(tr.H[-1][r0:r1], tr.Hlo[-1][r0:r1], 0)], wt, wtlo, 2 * A,
                                 PG.EPI_ACTOR, bias=ga.cat('actor/dist_out/out/bias', 'actor/dist_out/std/bias'),
                                 out=self.taction[t], out_lo=self.taction_lo[t], x0=self.eps_act[t],
                                 x1=self.tmean[t], x2=self.tstd[t], f0=s.actor_minstd,
                                 f1=s.actor_maxstd, ipar=A))

        actor_layers(0)
        win, winlo = gm.PT('rssm/img_in/kernel')
        wg, wglo = gm.PT('rssm/gru_out/kernel')
        wst, wstlo = gm.PT('rssm/img_stats/kernel')
        pr = self.i_prior
        for t in range(H):
          r0, r1 = t * N, (t + 1) * N
          layers.append(PG.Layer(
              [PG.Seg(tf[t][:, D:], None, 0), PG.Seg(self.taction[t], self.taction_lo[t], SC)], win, winlo,
              U, PG.EPI_LN, act=1, gamma=gm.P('rssm/img_in/norm/scale'), beta=gm.P('rssm/img_in/norm/bias'),
              z=self.i_zin[t], st=self.i_stin[t], out=self.i_xin[t], out_lo=self.i_xin_lo[t]))
          layers.append(PG.Layer(
              [PG.Seg(tf[t][:, :D], tlo[t], 0), PG.Seg(self.i_xin[t], self.i_xin_lo[t], D)], wg, wglo,
              3 * D, PG.EPI_GRU, gamma=gm.P('rssm/gru_out/norm/scale'), beta=gm.P('rssm/gru_out/norm/bias'),
              z=self.i_zgru[t], st=self.i_stgru[t], out=tf[t + 1][:, :D], out_lo=tlo[t + 1],
              x0=tf[t][:, :D], x1=self.i_ygru[t]))
          for i in range(pr.layers):
            kn, sn, bn = pr._names(i)
            wt, wtlo = gm.PT(kn)
            seg = (PG.Seg(tf[t + 1][:, :D], tlo[t + 1], 0) if i == 0 else
                   PG.Seg(pr.H[i - 1][r0:r1], pr.Hlo[i - 1][r0:r1], 0))
            layers.append(PG.Layer([seg], wt, wtlo, U, PG.EPI_LN, act=1, gamma=gm.P(sn), beta=gm.P(bn),
                                   z=pr.Z[i][r0:r1], st=pr.st[i][r0:r1], out=pr.H[i][r0:r1],
                                   out_lo=pr.Hlo[i][r0:r1]))
          layers.append(PG.Layer([PG.Seg(pr.H[-1][r0:r1], pr.Hlo[-1][r0:r1], 0)], wst, wstlo, SC,
                                 PG.EPI_STATS, bias=gm.P('rssm/img_stats/bias'), out=self.tlogit[t + 1],
                                 x0=self.g_img[t], x1=tf[t + 1][:, D:], f0=s.rssm_unimix, ipar=s.classes))
          actor_layers(t + 1)
        self._rollout_layers = layers            # kept for scripts/rollout_prof.py
        self._rollout = be.program_build(layers, N)
    if self._rollout is None:
      return False
    be.program_run(self._rollout)
    return True

  def _reverse_program(self):
    """The reverse pass through the imagined rollout (dynamics backprop of the continuous actor,
    A/agent.py:351-381 under the persistent tape) as ONE machine program: per step the categorical
    straight-through backward, the prior stack's reverse (3 x LayerNorm-backward layers), the GRU +
    joint-LayerNorm backward, and the input gradients of gru_out / img_in accumulated into the
    trajectory gradient -- 9 layers per step, t = H .. 1.  Parameters are constants here (no weight
    gradients).  False when a shape does not tile.  DD_REVERSE=0 keeps the per-op launches (the
    r02 A/B: 81 vs 303 launches and 8.5 vs 9.5 ms for the actor update at a1 size)."""
    s, be, gm = self.spec, self.be, self.groups['model']
    D, U, SC, A, N, H = s.deter, s.units, s.sc, s.act_dim, self.N, self.H
    if not hasattr(self, '_reverse'):
      ok = (self.machine and os.environ.get('DD_REVERSE', '1') != '0' and not s.discrete and
            s.classes == 32 and D % 256 == 0 and U % 256 == 0 and SC % 256 == 0 and A % 4 == 0 and
            A <= 64 and self.i_prior.machine and self.i_prior.layers >= 1)
      self._reverse = None
      if ok:
        pr = self.i_prior
        nl = pr.layers
        dtf = self.dtfeat
        wst, wstlo = gm.PB('rssm/img_stats/kernel')
        wgru, wgrulo = gm.PB('rssm/gru_out/kernel')
        win, winlo = gm.PB('rssm/img_in/kernel')
        dZ = [self.dz_buf(i, N, U) for i in range(nl)]
        layers = []
        for t in range(H, 0, -1):
          k = t - 1
          r0, r1 = k * N, (k + 1) * N
          # d(one-hot sample) -> d(raw prior logits)
          layers.append(PG.Layer([], None, None, SC, PG.EPI_STATSBWD, add=dtf[t][:, D:], out=self.r_dzs,
                                 out_lo=self.r_dzs_lo, x0=self.tlogit[t], f0=s.rssm_unimix, ipar=s.classes))
          # img_stats and the prior stack, reversed
          src, w, wlo = (self.r_dzs, self.r_dzs_lo), wst, wstlo
          for i in reversed(range(nl)):
            kn, sn, bn = pr._names(i)
            layers.append(PG.Layer([PG.Seg(src[0], src[1], 0)], w, wlo, U, PG.EPI_LNBWD, act=1,
                                   gamma=gm.P(sn), beta=gm.P(bn), z=pr.Z[i][r0:r1], st=pr.st[i][r0:r1],
                                   out=dZ[i][0], out_lo=dZ[i][1]))
            src = dZ[i]
            w, wlo = gm.PB(kn)
          # d(new deter) = d(img_out_0 input) + what the heads put there -> GRU + joint LayerNorm backward
          layers.append(PG.Layer([PG.Seg(src[0], src[1], 0)], w, wlo, D, PG.EPI_GRUBWD, add=dtf[t][:, :D],
                                 gamma=gm.P('rssm/gru_out/norm/scale'), beta=gm.P('rssm/gru_out/norm/bias'),
                                 z=self.i_zgru[k], st=self.i_stgru[k], out=self.r_dzg, out_lo=self.r_dzg_lo,
                                 x0=self.tfeat[k][:, :D], x1=self.i_ygru[k], x2=dtf[k][:, :D]))
          gseg = [PG.Seg(self.r_dzg, self.r_dzg_lo, 0)]
          # gru_out input gradients: x part -> img_in's LayerNorm backward, deter part -> trajectory
          layers.append(PG.Layer(gseg, wgru, wgrulo, U, PG.EPI_LNBWD, act=1, wn0=D,
                                 gamma=gm.P('rssm/img_in/norm/scale'), beta=gm.P('rssm/img_in/norm/bias'),
                                 z=self.i_zin[k], st=self.i_stin[k], out=self.r_dzi, out_lo=self.r_dzi_lo))
          if k > 0:
            layers.append(PG.Layer(gseg, wgru, wgrulo, D, PG.EPI_PLAIN, wn0=0, add=dtf[k][:, :D],
                                   out=dtf[k][:, :D]))
          iseg = [PG.Seg(self.r_dzi, self.r_dzi_lo, 0)]
          if k > 0:
            layers.append(PG.Layer(iseg, win, winlo, SC, PG.EPI_PLAIN, wn0=0, add=dtf[k][:, D:],
                                   out=dtf[k][:, D:]))
          layers.append(PG.Layer(iseg, win, winlo, A, PG.EPI_PLAIN, wn0=SC, out=self.da[r0:r1]))
        self._reverse_layers = layers            # kept for tests / profiling scripts
        self._reverse = be.program_build(layers, N)
    if self._reverse is None:
      return False
    be.program_run(self._reverse)
    return True

  def _cell_program(self, t, act, own_action):
    """img_in (Linear + LayerNorm + elu over [stoch, action]) and the GRU's Linear + joint
    LayerNorm of rollout step t as one 2-layer machine program (A/nets.py:129, :153-154).
    False when the shapes do not tile (the caller then issues the per-op launches)."""
    s, be, gm = self.spec, self.be, self.groups['model']
    D, U, SC = s.deter, s.units, s.sc
    if not (self.machine and U % 256 == 0 and (3 * D) % 256 == 0 and D % 4 == 0 and SC % 4 == 0):
      return False
    if not hasattr(self, '_cell_progs'):
      self._cell_progs = {}
    prog = self._cell_progs.get(t)
    if prog is None:
      if own_action:
        act_lo = self.taction_lo[t]
      else:                      # recorded actions (report): remainders computed on the fly
        if not hasattr(self, '_act_lo'):
          self._act_lo = be.zeros(*self.taction.shape)
        act_lo = self._act_lo[t]
      if not (_mview_ok(act) and _mview_ok(act_lo)):
        self._cell_progs[t] = False
        return False
      tf = self.tfeat
      wt, wtlo = gm.PT('rssm/img_in/kernel')
      gt, gtlo = gm.PT('rssm/gru_out/kernel')
      layers = [
          PG.Layer([PG.Seg(tf[t][:, D:], None, 0), PG.Seg(act, act_lo, SC)], wt, wtlo, U, PG.EPI_LN,
                   act=1, gamma=gm.P('rssm/img_in/norm/scale'), beta=gm.P('rssm/img_in/norm/bias'),
                   z=self.i_zin[t], st=self.i_stin[t], out=self.i_xin[t], out_lo=self.i_xin_lo[t]),
          PG.Layer([PG.Seg(tf[t][:, :D], self.tfeat_lo[t], 0), PG.Seg(self.i_xin[t], self.i_xin_lo[t], D)],
                   gt, gtlo, 3 * D, PG.EPI_LN, act=0, gamma=gm.P('rssm/gru_out/norm/scale'),
                   beta=gm.P('rssm/gru_out/norm/bias'), z=self.i_zgru[t], st=self.i_stgru[t],
                   out=self.i_ygru[t])]
      prog = self._cell_progs[t] = (be.program_build(layers, self.N), act_lo, own_action)
    if prog is False:
      return False
    if not prog[2]:
      be.split_lo(act, prog[1])
    be.program_run(prog[0])
    return True

  # ==========================================================================
  # 3. trajectory heads, lambda-return, critic update (A/agent.py:398-454)
  # ==========================================================================
  def _tsegs(self, r0, r1):
    """Trajectory feature rows [r0, r1) of the flattened [T*N, D+SC] rollout as trunk inputs."""
    flat = self.tfeat.view(-1, self.spec.feat)
    lo = self.tfeat_lo.view(-1, self.spec.deter)[r0:r1] if self.machine else None
    return self.fsegs(flat[r0:r1], lo)

  def _traj_head(self, trunk, group, prefix, out, r0, r1):
    trunk.fwd(self._tsegs(r0, r1), r0, r1, group,
              heads=[(f'{prefix}/dist_out/out/kernel', f'{prefix}/dist_out/out/bias',
                      out.view(-1, 1)[r0:r1])])

  def traj_values(self, slow_only=False):
    s, be = self.spec, self.be
    N, H, T = self.N, self.H, self.H + 1
    gm, gs = self.groups['model'], self.groups['slow_critic']
    if not slow_only:
      self._traj_head(self.t_reward, gm, 'reward', self.rho, N, T * N)
      self._traj_head(self.t_cont, gm, 'cont', self.kraw, N, T * N)
    self._traj_head(self.t_slow, gs, 'slow_critic', self.vraw, 0, T * N)
    be.lambda_fwd(self.rho, self.vraw, self.kraw, self.cont0, s.discount, s.lam,
                  self.r_reward, self.r_value, self.r_cont, self.r_weight, self.r_ret)

  def critic_update(self):
    self.critic_grads()
    self.critic_apply()

  def critic_grads(self):
    """Critic loss, its gradients and metrics (A/agent.py:417-442).  Reads the trajectory and the
    lambda-returns only and writes nothing the actor update reads, so launch_step may issue it on a
    second stream next to the reverse rollout (whose 800-row products leave most SMs idle)."""
    s, be, gc = self.spec, self.be, self.groups['critic']
    N, H = self.N, self.H
    M = H * N
    flat = self.tfeat.view(-1, s.feat)[:M]
    h = self.t_critic.fwd(self._tsegs(0, M), 0, M,
                          heads=[('critic/dist_out/out/kernel', 'critic/dist_out/out/bias', self.vpred)])
    coef = 1.0 / (M * self.world)
    be.head_loss(1, self.vpred, self.r_ret.view(M, 1), self.r_weight.view(-1)[:M],
                 coef, self.closs_row, self.d_vpred)
    be.fill_(gc.g, 0.0)
    dH = self.tmp(0, M, s.head_units)
    self._out_bwd(gc, 'critic/dist_out/out', h, self.d_vpred, dH, True)
    self.t_critic.bwd(dH, 0, M, self._tsegs(0, M), [], True)
    # metrics before the update changes anything
    self._stat(self.r_reward, 'imag_reward')
    be.symexp(self.vpred, self.vsym)
    self._stat(self.vsym, 'imag_critic')
    self._stat(self.r_ret, 'imag_return')
    self._stat(self.closs_row, 'closs')

  def critic_apply(self):
    s, be, gc = self.spec, self.be, self.groups['critic']
    self.opt_step('critic', C_CRITIC)
    be.slow_sync(gc.p, self.groups['slow_critic'].p, self.counters, C_SLOW, s.slow_every)
    self.groups['slow_critic'].prep(be)

  # ==========================================================================
  # 4. actor update (A/agent.py:326-381), gradients through the dynamics
  # ==========================================================================
  def actor_update(self, slow_refreshed=True):
    s, be = self.spec, self.be
    gm, ga, gs = self.groups['model'], self.groups['actor'], self.groups['slow_critic']
    N, H, T = self.N, self.H, self.H + 1
    D, U, SC, A = s.deter, s.units, s.sc, s.act_dim
    M = H * N
    hu = s.head_units
    # score with the slow critic: its values are unchanged since traj_values() unless
    # update_slow just copied the critic into it (first step and every slow_every-th)
    if slow_refreshed:
      self.traj_values(slow_only=True)
    be.sub(self.score, self.r_ret, self.r_value[:H])
    self._stat(self.score, 'score')
    # Normalize EMAs: retnorm/scorenorm are impl 'off' (state only), advnorm used.  The three
    # statistics rows are adjacent in mstat: one collective for all of them under data parallelism.
    norms = ((self.score, 'adv_ctrl', s.advnorm_decay), (self.r_ret, 'ret_ctrl', 0.999),
             (self.score, 'score_ctrl', 0.99))
    i0 = self.midx['adv_ctrl']
    assert [self.midx[k] for _, k, _ in norms] == [i0, i0 + 1, i0 + 2]
    for src, key, _ in norms:
      be.stats(src, self.mstat[self.midx[key]])
    if self.allreduce is not None:
      self.allreduce(self.mstat[i0:i0 + 3])
    for i, (_, key, decay) in enumerate(norms):
      be.normalize_update(self.norm_state[i], self.mstat[self.midx[key]], decay, s.advnorm_max,
                          self.adv if i == 0 else self.scratch2)
    if s.discrete:
      return self._actor_update_reinforce()
    # entropy controller (A/agent.py:361-371, A/tfutils.py:464-474)
    be.actor_ent(self.tstd.view(-1, A)[:M], s.actor_minstd, s.actor_maxstd, self.ent)
    be.colsum(self.ent, self.ent_sum)
    if self.allreduce is not None:
      self.allreduce(self.ent_sum)
    be.autoadapt(self.actent_scale, self.ent_sum, float(M * self.world), s.actent_target,
                 s.actent_min, s.actent_max, s.ctrl_vel, s.ctrl_thres, True)
    self._stat(self.ent, 'ent')
    be.actor_loss(self.r_ret, self.r_value, self.adv, self.ent, self.actent_scale,
                  self.r_weight, self.aloss_row)
    self._stat(self.aloss_row, 'aloss')

    coef = 1.0 / (M * self.world)
    be.lambda_bwd(self.rho, self.vraw, self.r_value, self.r_cont, self.r_weight,
                  self.r_ret, self.adv, coef, s.discount, s.lam,
                  self.d_rho, self.d_vraw, self.d_kraw)
    # heads: input grads only (world model and slow critic are constants here)
    dtf = self.dtfeat.view(-1, s.feat)
    flat = self.tfeat.view(-1, s.feat)
    for trunk, g, prefix, dOut, r0, beta in (
        (self.t_slow, gs, 'slow_critic', self.d_vraw, 0, 0.0),
        (self.t_reward, gm, 'reward', self.d_rho, N, 1.0),
        (self.t_cont, gm, 'cont', self.d_kraw, N, 1.0)):
      n = T * N - r0
      dH = self.tmp(0, n, hu)
      self._out_bwd(g, f'{prefix}/dist_out/out', trunk.H[-1][r0:], dOut.view(-1, 1)[r0:], dH, False)
      trunk.bwd(dH, r0, T * N, self._tsegs(r0, T * N), [(dtf[r0:], 0, beta)], False, g)

    # reverse through the rollout: t = H .. 1
    Win = gm.P('rssm/img_in/kernel')
    Wgru = gm.P('rssm/gru_out/kernel')
    for t in ([] if self._reverse_program() else range(H, 0, -1)):
      k = t - 1
      dzs = self.tmp(2, N, SC)
      be.stats_bwd(self.dtfeat[t][:, D:], None, self.tlogit[t], dzs, s.stoch,
                   s.classes, s.rssm_unimix)
      dH = self.tmp(0, N, U)
      self._out_bwd(gm, 'rssm/img_stats', self.i_prior.H[-1][k * N:(k + 1) * N], dzs, dH, False)
      self._prior_trunk_bwd(self.i_prior, k * N, (k + 1) * N, self.tfeat[t][:, :D], dH,
                            self.dtfeat[t][:, :D], 1.0, False,
                            Xlo=self.tfeat_lo[t] if self.machine else None)
      dy = self.tmp(0, N, 3 * D)
      dzg = self.tmp(1, N, 3 * D)
      be.gru_bwd(self.dtfeat[t][:, :D], self.i_ygru[k], self.tfeat[k][:, :D], dy,
                 self.dtfeat[k][:, :D], True)
      be.ln_bwd(dy, self.i_zgru[k], self.i_stgru[k], gm.P('rssm/gru_out/norm/scale'),
                gm.P('rssm/gru_out/norm/bias'), dzg, None, None, 0)
      if k > 0:
        be.gemm(dzg, Wgru[:D], self.dtfeat[k][:, :D], tb=True, beta=1.0)
      dx = self.tmp(0, N, U)
      be.gemm(dzg, Wgru[D:], dx, tb=True)
      dzi = self.tmp(2, N, U)
      be.ln_bwd(dx, self.i_zin[k], self.i_stin[k], gm.P('rssm/img_in/norm/scale'),
                gm.P('rssm/img_in/norm/bias'), dzi, None, None, 1)
      if k > 0:
        be.gemm(dzi, Win[:SC], self.dtfeat[k][:, D:], tb=True, beta=1.0)
      be.gemm(dzi, Win[SC:], self.da[k * N:(k + 1) * N], tb=True)

    # actor network backward over all t < H at once
    be.actor_bwd(self.da, self.eps_act.view(-1, A)[:M], self.tmean.view(-1, A)[:M],
                 self.tstd.view(-1, A)[:M], self.r_weight.view(-1)[:M],
                 self.actent_scale, coef, s.actor_minstd, s.actor_maxstd,
                 self.d_mraw, self.d_sraw)
    be.fill_(ga.g, 0.0)
    h = self.actor.H[-1][:M]
    dH = self.tmp(0, M, hu)
    be.gemm(h, self.d_mraw, ga.G('actor/dist_out/out/kernel'), ta=True, beta=1.0)
    be.colsum(self.d_mraw, ga.G('actor/dist_out/out/bias'), acc=True)
    be.gemm(h, self.d_sraw, ga.G('actor/dist_out/std/kernel'), ta=True, beta=1.0)
    be.colsum(self.d_sraw, ga.G('actor/dist_out/std/bias'), acc=True)
    be.gemm(self.d_mraw, ga.P('actor/dist_out/out/kernel'), dH, tb=True)
    be.gemm(self.d_sraw, ga.P('actor/dist_out/std/kernel'), dH, tb=True, beta=1.0)
    self.actor.bwd(dH, 0, M, self._tsegs(0, M), [], True)
    self.opt_step('actor', C_ACTOR)

  def _actor_update_reinforce(self):
    """Discrete actions: loss = -log pi(a) sg(score) - scale*ent (A/agent.py:357-358,
    :372-379); the score is a constant, so nothing flows back through the rollout."""
    s, be, ga = self.spec, self.be, self.groups['actor']
    N, H, A = self.N, self.H, s.act_dim
    M = H * N
    logit = self.talogit.view(-1, A)[:M]
    be.actor_disc_ent(logit, self.ent)
    be.colsum(self.ent, self.ent_sum)
    if self.allreduce is not None:
      self.allreduce(self.ent_sum)
    be.autoadapt(self.actent_scale, self.ent_sum, float(M * self.world), s.actent_target,
                 s.actent_min, s.actent_max, s.ctrl_vel, s.ctrl_thres, True)
    self._stat(self.ent, 'ent')
    coef = 1.0 / (M * self.world)
    dlogit = self.da                       # [M, A] scratch
    be.actor_disc_loss(logit, self.taction.view(-1, A)[:M], self.r_ret, self.r_value, self.adv,
                       self.ent, self.actent_scale, self.r_weight, coef, self.aloss_row, dlogit)
    self._stat(self.aloss_row, 'aloss')
    be.stats_bwd(None, dlogit, logit, self.d_mraw, 1, A, s.actor_unimix)
    be.fill_(ga.g, 0.0)
    h = self.actor.H[-1][:M]
    dH = self.tmp(0, M, s.head_units)
    self._out_bwd(ga, 'actor/dist_out/out', h, self.d_mraw, dH, True)
    self.actor.bwd(dH, 0, M, self._tsegs(0, M), [], True)
    self.opt_step('actor', C_ACTOR)

  # ==========================================================================
  # the whole step
  # ==========================================================================
  def slow_refresh_due(self, updates):
    """Host mirror of VFunction.update_slow (A/agent.py:444-454): given the value of
    the `updates` counter BEFORE this step, does this step hard-copy the critic?"""
    return updates == -1 or updates >= self.spec.slow_every

  def launch_step(self, draw=True, slow_refreshed=True):
    """Everything between 'inputs are on the device' and 'metrics are on the
    device'.  Pure launch sequence: safe to capture into a CUDA graph.
    slow_refreshed=False (only valid when slow_refresh_due() is False for this
    step) skips re-evaluating the slow critic after the critic update."""
    if draw:
      self.draw_noise()
    self.be.stats_init(self.mstat)
    self.wm_forward()
    self.wm_losses()
    self.wm_backward()
    self.opt_step('model', C_MODEL)
    # carried train state = post[:, -1] (A/agent.py:211)
    s, B, L = self.spec, self.B, self.L
    r = slice((L - 1) * B, L * B)
    self.be.copy_(self.prev_deter, self.feat[r, :s.deter])
    self.be.copy_(self.prev_stoch, self.feat[r, s.deter:])
    self.be.copy_(self.cont0.view(-1, 1), self.cont)
    self.imagine()
    self.traj_values()
    if self.overlap and not slow_refreshed and hasattr(self.be, 'branch'):
      # the slow critic the actor scores with is not refreshed this step: critic and actor updates
      # are independent until their optimizers.  No collective is issued on the second stream (one
      # communicator must not run on two streams at once): the critic's all-reduce follows the join.
      with self.be.branch():
        self._tmp, self._tmp_alt = self._tmp_alt, self._tmp
        self._dzp, self._dzp_alt = self._dzp_alt, self._dzp
        try:
          self.critic_grads()
        finally:
          self._tmp, self._tmp_alt = self._tmp_alt, self._tmp
          self._dzp, self._dzp_alt = self._dzp_alt, self._dzp
      self.actor_update(False)
      self.be.join()
      self.critic_apply()
    else:
      self.critic_update()
      self.actor_update(slow_refreshed)

  def launch_report(self, draw=True):
    """WorldModel.report's scalar part (A/agent.py:266-268): the loss metrics of
    a batch without touching parameters, controllers or the carried state."""
    if draw:
      self.draw_noise()
    self.be.stats_init(self.mstat)
    self.wm_forward()
    self.wm_losses(update=False)

  # -- host side of the metrics --------------------------------------------
  def read_metrics(self, report=False) -> Dict[str, float]:
    s = self.spec
    ms = self.mstat.detach().cpu().numpy()
    cnt = self.counters.detach().cpu().numpy()
    wm = float(self.wmkl_scale.detach().cpu().numpy()[0])
    ae = self.actent_scale.detach().cpu().numpy().astype(np.float64)
    sq = {g: float(self.groups[g].sumsq.detach().cpu().numpy()[0])
          for g in ('model', 'critic', 'actor')}
    row = lambda k: ms[self.midx[k]]
    mean = lambda k: row(k)[0] / row(k)[6]
    std = lambda k: math.sqrt(max(row(k)[1] / row(k)[6] - mean(k) ** 2, 0.0))
    m = {}
    m.update(wmkl_mean=mean('kl'), wmkl_std=std('kl'), wmkl_scale_mean=wm,
             wmkl_scale_std=0.0, kl_loss_mean=wm * mean('kl'), kl_loss_std=wm * std('kl'))
    for k in list(s.cnn_keys) + list(s.dec_vec):
      m[f'{k}_loss_mean'], m[f'{k}_loss_std'] = mean(f'loss_{k}'), std(f'loss_{k}')
    for k in ('reward', 'cont'):
      m[f'{k}_loss_mean'], m[f'{k}_loss_std'] = mean(f'{k}_loss'), std(f'{k}_loss')
      b = row(f'bal_{k}')
      div = lambda a, c: a / c if c > 0 else float('nan')
      n = b[0] + b[1]
      m.update({f'{k}_pos_loss': div(b[2], b[0]), f'{k}_neg_loss': div(b[3], b[1]),
                f'{k}_pos_acc': div(b[4], b[0]), f'{k}_neg_acc': div(b[5], b[1]),
                f'{k}_rate': b[0] / n, f'{k}_avg': b[6] / n, f'{k}_pred': b[7] / n})
    m.update(prior_ent_mean=mean('ent_prior'), post_ent_mean=mean('ent_post'),
             prior_ent_min=row('ent_prior')[2], post_ent_min=row('ent_post')[2],
             model_loss_mean=mean('total_loss'), model_loss_std=std('total_loss'))
    if report:
      return m
    m.update(model_loss=mean('total_loss'),
             model_grad_norm=math.sqrt(sq['model']), model_grad_steps=int(cnt[C_MODEL]),
             extr_critic_loss=mean('closs'),
             extr_critic_grad_norm=math.sqrt(sq['critic']),
             extr_critic_grad_steps=int(cnt[C_CRITIC]),
             extr_imag_reward_mean=mean('imag_reward'), extr_imag_reward_std=std('imag_reward'),
             extr_imag_critic_mean=mean('imag_critic'), extr_imag_critic_std=std('imag_critic'),
             extr_imag_return_mean=mean('imag_return'), extr_imag_return_std=std('imag_return'),
             extr_score_mean=mean('score'), extr_score_std=std('score'),
             extr_score_mag=row('score')[4] / row('score')[6], extr_score_max=row('score')[5],
             actent_mean=mean('ent'), actent_std=std('ent'),
             actent_scale_mean=float(ae.mean()), actent_scale_std=float(ae.std()),
             actor_loss=mean('aloss'),
             actor_grad_norm=math.sqrt(sq['actor']), actor_grad_steps=int(cnt[C_ACTOR]))
    return m


# ----------------------------------------------------------------------------
# Image summaries of report(): open-loop prediction and imagined rollouts
# ----------------------------------------------------------------------------

class VideoReport:
  """The image part of `Agent.report` for configs with cnn keys:
    openl_<key>      WorldModel.report (A/agent.py:266-282): 5 context frames reconstructed
                     from the posterior, the rest predicted open loop along the recorded
                     actions; [truth | model | (model - truth + 1) / 2] stacked on the height axis
    task_imag_<key>  Greedy.report (A/behaviors.py:32-46): decoded imagination rollout of the
                     policy from the last context state
  Both laid out by video_grid (A/tfutils.py:390-392): [T, H, B*W, C], B = min(6, batch).
  Three small engines share the learner's parameters: `ctx` observes the context, `ol` rolls
  the prior along the recorded actions, `im` imagines with the actor."""

  CONTEXT = 5

  def __init__(self, main: Engine, B: int, L: int):
    spec, be = main.spec, main.be
    if L <= self.CONTEXT:
      raise ValueError(f'report() videos need sequences longer than {self.CONTEXT} steps')
    self.main, self.be, self.spec = main, be, spec
    self.B, self.L, self.T2 = B, L, L - self.CONTEXT
    gm = main.groups['model']
    self.ctx = Engine(spec, be, B, self.CONTEXT, seed=main.seed + 11, share=main)
    self.ol = Engine(dataclasses.replace(spec, horizon=self.T2), be, B, 1,
                     seed=main.seed + 12, share=main)
    self.im = Engine(spec, be, B, 1, seed=main.seed + 13, share=main)
    self.actions = be.zeros(self.T2, B, spec.act_dim)         # recorded actions, time-major
    self.dec_ol = ConvDecoder(self.ol, gm, self.T2 * B, train=False)
    self.dec_im = ConvDecoder(self.im, gm, (spec.horizon + 1) * B, train=False)
    img = tuple(spec.img_shape)
    self.recon = be.zeros(self.CONTEXT * B, *img)
    self.openl = be.zeros(self.T2 * B, *img)
    self.imag = be.zeros((spec.horizon + 1) * B, *img)

  def launch(self, draw=True):
    """Inputs: ctx.{obs,img,action,first} (context, time-major), self.actions (steps 5..L-1),
    im.cont0.  Outputs: self.recon / self.openl / self.imag, sigmoid images, time-major."""
    s, be = self.spec, self.be
    ctx, ol, im = self.ctx, self.ol, self.im
    B, C5 = self.B, self.CONTEXT
    if draw:
      for e in (ctx, ol, im):
        e.draw_noise()
    be.fill_(ctx.first0, 1.0)                 # observe(..., state=None): initial state
    ctx.wm_forward()
    be.sigmoid(ctx.dec_cnn.Z[-1], self.recon)
    last = slice((C5 - 1) * B, C5 * B)
    for e in (ol, im):                        # start = context[:, -1]
      be.copy_(e.feat, ctx.feat[last])
      be.copy_(e.post_logit, ctx.post_logit[last])
    ol.imagine(actions=self.actions)
    flat = ol.tfeat.view(-1, s.feat)
    be.sigmoid(self.dec_ol.fwd(flat[B:(self.T2 + 1) * B]), self.openl)
    im.imagine()
    be.sigmoid(self.dec_im.fwd(im.tfeat.view(-1, s.feat)), self.imag)

  def videos(self, truth):
    """truth: {key: float [B, L, H, W, C] in [0,1]} -> dict of [T, H', B*W, C] arrays."""
    s = self.spec
    B, L, C5 = self.B, self.L, self.CONTEXT
    grid = lambda v: v.transpose(1, 2, 0, 3, 4).reshape(v.shape[1], v.shape[2], -1, v.shape[4])
    bt = lambda x, T: x.detach().cpu().numpy().reshape((T, B) + x.shape[1:]).swapaxes(0, 1)
    model = np.concatenate([bt(self.recon, C5), bt(self.openl, self.T2)], 1)     # [B, L, H, W, C]
    imag = bt(self.imag, s.horizon + 1)
    out, c0 = {}, 0
    for k, shp in s.cnn_keys.items():
      m = model[..., c0:c0 + shp[2]]
      t = np.asarray(truth[k], np.float32)
      out[f'openl_{k}'] = grid(np.concatenate([t, m, (m - t + 1) / 2], 2))
      out[f'task_imag_{k}'] = grid(imag[..., c0:c0 + shp[2]])
      c0 += shp[2]
    return out


# ----------------------------------------------------------------------------
# Acting path: one obs_step + actor per call (A/agent.py:42-65)
# ----------------------------------------------------------------------------

STREAM_POL_OBS, STREAM_POL_ACT, C_POLICY = 4, 5, 5


class PolicyEngine:
  """`Agent.policy`: encoder -> one RSSM.obs_step -> actor, for P parallel
  environments, on the learner's live parameters (shares `eng.groups`)."""

  def __init__(self, eng: Engine, P: int):
    self.eng, self.be, self.spec, self.P = eng, eng.be, eng.spec, P
    s, z = eng.spec, eng.be.zeros
    D, U, SC, A, E = s.deter, s.units, s.sc, s.act_dim, s.enc_units
    gm, ga = eng.groups['model'], eng.groups['actor']
    self.obs = z(P, s.obs_total)
    self.img = z(P, *s.img_shape) if s.cnn_keys else None
    self.img_u8 = {k: z(P, 1, *shp, dtype=torch.uint8) for k, shp in s.cnn_keys.items()}
    self.first = z(P)
    self.prev_action = z(P, A)
    self.prev_deter = z(P, D)
    self.prev_stoch = z(P, SC)
    self.enc = Trunk(eng, gm, 'enc', s.enc_layers, E, P) if s.obs_dims else None
    self.enc_cnn = ConvEncoder(eng, gm, P, train=False) if s.cnn_keys else None
    self.init_deter = z(1, D)
    self.init_trunk = Trunk(eng, gm, 'rssm/initp', s.prior_layers, U, 1, names=eng._prior_names)
    self.init_zs, self.init_logit, self.init_stoch = z(1, SC), z(1, SC), z(1, SC)
    self.m_deter, self.m_stoch, self.m_act = z(P, D), z(P, SC), z(P, A)
    self.z_in, self.x_in, self.st_in = z(P, U), z(P, U), z(P, 2)
    self.z_gru, self.y_gru, self.st_gru = z(P, 3 * D), z(P, 3 * D), z(P, 2)
    self.feat = z(P, D + SC)
    self.z_obs, self.h_obs, self.st_obs = z(P, U), z(P, U), z(P, 2)
    self.zs, self.logit = z(P, SC), z(P, SC)
    self.g, self.eps = z(P, SC), z(P, A)
    self.actor = Trunk(eng, ga, 'actor', s.head_layers, s.head_units, P)
    self.mraw, self.sraw = z(P, A), z(P, A)
    self.action, self.mean, self.std = z(P, A), z(P, A), z(P, A)

  def launch(self, draw=True):
    """draw=False: the caller has filled self.g (Gumbel [P, SC]) and self.eps ([P, A]) itself
    (parity tests inject the oracle's noise)."""
    eng, be, s = self.eng, self.be, self.spec
    gm, ga = eng.groups['model'], eng.groups['actor']
    D, SC = s.deter, s.sc
    P = self.P
    if draw:
      be.fill_gumbel(self.g, eng.seed, STREAM_POL_OBS, eng.counters, C_POLICY)
      if s.discrete:
        be.fill_gumbel(self.eps, eng.seed, STREAM_POL_ACT, eng.counters, C_POLICY)
      else:
        be.fill_normal(self.eps, eng.seed, STREAM_POL_ACT, eng.counters, C_POLICY)
      be.counter_inc(eng.counters, C_POLICY)
    parts = []
    c0 = 0
    for k, shp in s.cnn_keys.items():          # uint8 -> [0,1], channel concat
      be.image_stage(self.img_u8[k], self.img, c0, 1.0 / 255.0)
      c0 += shp[2]
    if s.cnn_keys:
      parts.append(self.enc_cnn.fwd(self.img))
    if s.obs_dims:
      parts.append(self.enc.fwd([(self.obs, 0)], 0, P))
    be.tanh_fwd(gm.P('rssm/initial_deter').view(1, D), self.init_deter)
    eng._prior_stats_fwd(self.init_trunk, 0, 1, self.init_deter, self.init_zs, None,
                         self.init_logit, self.init_stoch, None, 1)
    be.mask_fwd(self.prev_deter, self.prev_stoch, self.prev_action, self.first,
                self.init_deter[0], self.init_stoch[0], self.m_deter, self.m_stoch, self.m_act)
    Win, Wgru, Wobs = (gm.P('rssm/img_in/kernel'), gm.P('rssm/gru_out/kernel'),
                       gm.P('rssm/obs_out/kernel'))
    be.gemm(self.m_stoch, Win[:SC], self.z_in)
    be.gemm(self.m_act, Win[SC:], self.z_in, beta=1.0)
    be.ln_fwd(self.z_in, gm.P('rssm/img_in/norm/scale'), gm.P('rssm/img_in/norm/bias'),
              self.x_in, self.st_in, 1)
    be.gemm(self.m_deter, Wgru[:D], self.z_gru)
    be.gemm(self.x_in, Wgru[D:], self.z_gru, beta=1.0)
    be.ln_fwd(self.z_gru, gm.P('rssm/gru_out/norm/scale'), gm.P('rssm/gru_out/norm/bias'),
              self.y_gru, self.st_gru, 0)
    be.gru_fwd(self.y_gru, self.m_deter, self.feat[:, :D])
    be.gemm(self.feat[:, :D], Wobs[:D], self.z_obs)
    off = D
    for part in parts:
      be.gemm(part, Wobs[off:off + part.shape[1]], self.z_obs, beta=1.0)
      off += part.shape[1]
    be.ln_fwd(self.z_obs, gm.P('rssm/obs_out/norm/scale'), gm.P('rssm/obs_out/norm/bias'),
              self.h_obs, self.st_obs, 1)
    eng._out_fwd(gm, 'rssm/obs_stats', self.h_obs, self.zs)
    be.stats_fwd(self.zs, self.g, self.logit, self.feat[:, D:], None, s.stoch, s.classes,
                 s.rssm_unimix, 0)
    h = self.actor.fwd([(self.feat, 0)], 0, P)
    eng._out_fwd(ga, 'actor/dist_out/out', h, self.mraw)
    if s.discrete:      # `action` = sample, `mean` = mode (A/agent.py:52-60)
      be.stats_fwd(self.mraw, self.eps, self.std, self.action, None, 1, s.act_dim,
                   s.actor_unimix, 0)
      be.stats_fwd(self.mraw, None, self.std, self.mean, None, 1, s.act_dim,
                   s.actor_unimix, 1)
    else:
      eng._out_fwd(ga, 'actor/dist_out/std', h, self.sraw)
      be.actor_sample(self.mraw, self.sraw, self.eps, s.actor_minstd, s.actor_maxstd,
                      self.action, self.mean, self.std)
    # carry: latent and the action just taken
    be.copy_(self.prev_deter, self.feat[:, :D])
    be.copy_(self.prev_stoch, self.feat[:, D:])
