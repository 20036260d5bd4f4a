"""A small torch-backed stand-in for the TensorFlow / TensorFlow-Probability / Sonnet API surface that
the reference learner (embodied/agents/dreamerv2plus: agent.py, nets.py, tfutils.py, behaviors.py) uses.

STATUS: prepared, not used by any test (see make_reference_goldens.py).  PURPOSE: test infrastructure only.  TensorFlow cannot be installed in this image, so the reference's
own Python cannot run as is.  With `install()` the reference modules import THESE modules as
`tensorflow`, `tensorflow_probability`, `sonnet` and `ruamel.yaml`, and the reference's code -- its
control flow, its wiring of layers, losses, optimizer, controllers, its formulas -- executes eagerly on
torch CPU tensors.  What is substituted is the primitive layer only (matmul, exp, reductions, autograd
instead of GradientTape).  `tests/golden/make_reference_goldens.py` runs the reference's
`Agent.train` this way and stores inputs and outputs as fixtures; `tests/test_reference_goldens.py`
checks the oracle (oracle/dreamer_oracle.py) against them.  Nothing in the product imports this file.

Random draws: `tf.random.categorical` and `tfd.Normal.sample` take their noise from `NOISE`, a
recorder that draws Gumbel / standard-normal numbers from a seeded numpy generator and remembers
every draw in call order, so the generator script can hand the oracle the very same noise.

Only what the reference calls is implemented; anything else raises AttributeError / NotImplementedError.
"""
import contextlib
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# tensors
# ---------------------------------------------------------------------------------------------


class Shape(list):
  """TensorShape look-alike: slices stay shapes, `shape + [n]` and `shape == (a, b)` work."""

  def __getitem__(self, i):
    r = list.__getitem__(self, i)
    return Shape(r) if isinstance(i, slice) else r

  def __add__(self, other):
    return Shape(list(self) + list(other))

  def __radd__(self, other):
    return Shape(list(other) + list(self))

  def __eq__(self, other):
    try:
      return list(self) == list(other)
    except TypeError:
      return False

  def __ne__(self, other):
    return not self.__eq__(other)

  __hash__ = None

  def as_list(self):
    return list(self)


def _plain(x):
  """torch.Tensor view without the TF-flavoured subclass (the reference patches methods onto it)."""
  if isinstance(x, torch.Tensor):
    return x.as_subclass(torch.Tensor)
  return x


def _wrap(x):
  return x.as_subclass(Tensor) if isinstance(x, torch.Tensor) else x


class Tensor(torch.Tensor):
  """What the reference sees as tf.Tensor.  Immutable like an EagerTensor: `x += y` rebinds."""

  @property
  def shape(self):
    return Shape(torch.Tensor.size(_plain(self)))

  def numpy(self):
    return _plain(self).detach().cpu().numpy()

  def __iadd__(self, other):
    return self + other

  def __isub__(self, other):
    return self - other

  def __imul__(self, other):
    return self * other

  def __itruediv__(self, other):
    return self / other

  def __len__(self):
    return int(torch.Tensor.size(_plain(self), 0))

  def __pow__(self, other):
    return _wrap(torch.pow(_plain(self), _plain(_arg(other, like=self))))

  def __rpow__(self, other):
    return _wrap(torch.pow(_plain(_arg(other, like=self)), _plain(self)))

  def __invert__(self):
    return _wrap(torch.logical_not(_plain(self)) if self.dtype == torch.bool else torch.bitwise_not(_plain(self)))

  def __and__(self, other):
    return _wrap(torch.logical_and(_plain(self), _plain(convert_to_tensor(other))))


class Variable(Tensor):
  """tf.Variable: a leaf tensor with assign / assign_add / assign_sub and a name."""

  def __new__(cls, initial_value=None, trainable=True, name=None, dtype=None, **kw):
    value = _plain(convert_to_tensor(initial_value, dtype)).detach().clone()
    grad = bool(trainable) and value.dtype.is_floating_point
    obj = torch.Tensor._make_subclass(cls, value, grad)
    return obj

  def __init__(self, initial_value=None, trainable=True, name=None, dtype=None, **kw):
    self.trainable = bool(trainable) and self.dtype.is_floating_point
    Variable._COUNT[0] += 1
    self._vname = f'{name or "Variable"}_{Variable._COUNT[0]}:0'

  _COUNT = [0]

  @property
  def name(self):
    return self._vname

  def assign(self, value):
    with torch.no_grad():
      _plain(self).copy_(_plain(convert_to_tensor(value)).to(self.dtype))
    return self

  def assign_add(self, value):
    with torch.no_grad():
      _plain(self).add_(_plain(convert_to_tensor(value)).to(self.dtype))
    return self

  def assign_sub(self, value):
    with torch.no_grad():
      _plain(self).sub_(_plain(convert_to_tensor(value)).to(self.dtype))
    return self

  def value(self):
    return _wrap(_plain(self) + 0)

  def __deepcopy__(self, memo):
    raise TypeError('variables are not copied')


float16, float32, float64 = torch.float16, torch.float32, torch.float64
int32, int64, uint8 = torch.int32, torch.int64, torch.uint8
_pybool = bool


def convert_to_tensor(value, dtype=None):
  if isinstance(value, torch.Tensor):
    t = _plain(value)
  elif isinstance(value, np.ndarray):
    t = torch.from_numpy(np.ascontiguousarray(value))
  elif isinstance(value, (np.floating, np.integer, np.bool_)):
    t = torch.from_numpy(np.asarray(value))
  elif isinstance(value, _pybool):
    t = torch.tensor(value)
  elif isinstance(value, int):
    t = torch.tensor(value, dtype=torch.int32)
  elif isinstance(value, float):
    t = torch.tensor(value, dtype=torch.float32)
  elif isinstance(value, (list, tuple)):
    if len(value) and all(isinstance(v, torch.Tensor) for v in value):
      t = torch.stack([_plain(v) for v in value])
    else:
      arr = np.asarray(value)
      if arr.dtype == np.float64:
        arr = arr.astype(np.float32)
      if arr.dtype == np.int64:
        arr = arr.astype(np.int32)
      t = torch.from_numpy(arr)
  else:
    raise TypeError(f'cannot convert {type(value)} to a tensor')
  if dtype is not None and t.dtype != dtype:
    t = t.to(dtype)
  return _wrap(t)


def _arg(x, like=None):
  """Operand of a binary op: python scalars follow the tensor's dtype (TF's weak typing)."""
  if isinstance(x, torch.Tensor):
    return x
  if like is not None and isinstance(x, (int, float)) and not isinstance(x, _pybool):
    return torch.tensor(x, dtype=like.dtype if like.dtype.is_floating_point or isinstance(x, int) else torch.float32)
  return convert_to_tensor(x)


def _t(x):
  return _plain(convert_to_tensor(x))


def _axes(axis, ndim):
  if axis is None:
    return None
  if isinstance(axis, (int, np.integer)):
    return (int(axis),)
  axis = tuple(int(a) for a in axis)
  return axis


def _reduce(fn):
  def run(x, axis=None, keepdims=False):
    t = _t(x)
    ax = _axes(axis, t.dim())
    if ax is None:
      return _wrap(fn(t))
    if len(ax) == 0:
      return _wrap(t)
    return _wrap(fn(t, dim=ax, keepdim=keepdims))
  return run


def _minmax(fn):
  def run(x, axis=None, keepdims=False):
    t = _t(x)
    ax = _axes(axis, t.dim())
    if ax is None:
      return _wrap(fn(t))
    for a in sorted((a % t.dim() for a in ax), reverse=True):
      t = fn(t, dim=a, keepdim=keepdims).values
    return _wrap(t)
  return run


class _Ns(types.ModuleType):
  pass


# ---------------------------------------------------------------------------------------------
# noise recorder
# ---------------------------------------------------------------------------------------------


class Noise:
  def __init__(self):
    self.reset(0)

  def reset(self, seed):
    self.rng = np.random.default_rng(seed)
    self.draws = []          # (kind, array) in call order

  def gumbel(self, shape):
    u = self.rng.uniform(1e-6, 1 - 1e-6, shape)
    g = (-np.log(-np.log(u))).astype(np.float32)
    self.draws.append(('gumbel', g))
    return torch.from_numpy(g)

  def normal(self, shape):
    e = self.rng.standard_normal(shape).astype(np.float32)
    self.draws.append(('normal', e))
    return torch.from_numpy(e)


NOISE = Noise()

# ---------------------------------------------------------------------------------------------
# tensorflow
# ---------------------------------------------------------------------------------------------


def _build_tf():
  tf = _Ns('tensorflow')
  tf.Tensor, tf.Variable = Tensor, Variable
  tf.float16, tf.float32, tf.float64 = float16, float32, float64
  tf.int32, tf.int64, tf.uint8, tf.bool = int32, int64, uint8, torch.bool
  tf.convert_to_tensor = convert_to_tensor

  class IndexedSlices:
    pass
  tf.IndexedSlices = IndexedSlices

  tf.zeros = lambda shape, dtype=float32: _wrap(torch.zeros(tuple(int(s) for s in (shape if hasattr(shape, '__len__') else [shape])), dtype=dtype))
  tf.ones = lambda shape, dtype=float32: _wrap(torch.ones(tuple(int(s) for s in (shape if hasattr(shape, '__len__') else [shape])), dtype=dtype))
  tf.zeros_like = lambda x, dtype=None: _wrap(torch.zeros_like(_t(x), dtype=dtype))
  tf.ones_like = lambda x, dtype=None: _wrap(torch.ones_like(_t(x), dtype=dtype))
  tf.identity = lambda x: _wrap(_t(x) + 0) if not isinstance(x, torch.Tensor) or True else x
  tf.stop_gradient = lambda x: _wrap(_t(x).detach())
  tf.cast = lambda x, dtype: _wrap(_t(x).to(dtype))
  tf.reshape = lambda x, shape: _wrap(_t(x).reshape(tuple(int(s) for s in shape)))

  def transpose(x, perm=None):
    t = _t(x)
    if perm is None:
      perm = list(reversed(range(t.dim())))
    return _wrap(t.permute(*[int(p) for p in perm]))
  tf.transpose = transpose
  tf.concat = lambda values, axis: _wrap(torch.cat([_t(v) for v in values], dim=int(axis)))
  tf.stack = lambda values, axis=0: _wrap(torch.stack([_t(v) for v in values], dim=int(axis)))
  tf.repeat = lambda x, repeats, axis=None: _wrap(torch.repeat_interleave(_t(x), int(repeats), dim=axis))

  def split(x, num_or_sizes, axis=0):
    t = _t(x)
    if isinstance(num_or_sizes, int):
      return [_wrap(p) for p in torch.chunk(t, num_or_sizes, dim=axis)]
    return [_wrap(p) for p in torch.split(t, [int(s) for s in num_or_sizes], dim=axis)]
  tf.split = split
  tf.where = lambda c, a, b: _wrap(torch.where(_t(c), _t(a) if isinstance(a, torch.Tensor) else torch.tensor(a, dtype=torch.float32), _t(b) if isinstance(b, torch.Tensor) else torch.tensor(b, dtype=torch.float32)))
  tf.sign = lambda x: _wrap(torch.sign(_t(x)))
  tf.abs = lambda x: _wrap(torch.abs(_t(x)))
  tf.tanh = lambda x: _wrap(torch.tanh(_t(x)))
  tf.sqrt = lambda x: _wrap(torch.sqrt(_t(x)))
  tf.maximum = lambda a, b: _wrap(torch.maximum(_t(a), _t(b).to(_t(a).dtype)))
  tf.range = lambda *a: _wrap(torch.arange(*[int(v) for v in a]))
  tf.gather = lambda x, idx, axis=0: _wrap(torch.index_select(_t(x), axis, _t(idx).long()))
  tf.einsum = lambda eq, *ops: _wrap(torch.einsum(eq, *[_t(o) for o in ops]))
  tf.one_hot = lambda idx, depth, dtype=float32: _wrap(F.one_hot(_t(idx).long(), int(depth)).to(dtype))
  tf.reduce_sum = _reduce(torch.sum)
  tf.reduce_all = lambda x, axis=None: _wrap(torch.all(_t(x)))
  tf.clip_by_value = lambda x, lo, hi: _wrap(torch.clamp(_t(x), min=float(lo) if not isinstance(lo, torch.Tensor) else _t(lo), max=float(hi) if not isinstance(hi, torch.Tensor) else _t(hi)))
  tf.argsort = lambda x, axis=-1: _wrap(torch.argsort(_t(x), dim=axis))
  tf.get_current_name_scope = lambda: ''

  def clip_by_global_norm(grads, clip, use_norm=None):
    norm = _t(use_norm) if use_norm is not None else _t(tf.linalg.global_norm(grads))
    scale = float(clip) / torch.clamp(norm, min=float(clip))
    return [_wrap(_t(g) * scale) for g in grads], _wrap(norm)
  tf.clip_by_global_norm = clip_by_global_norm

  def function(fn=None, **kw):
    if fn is None:
      return lambda f: f
    return fn
  tf.function = function

  # ---- math
  m = _Ns('tensorflow.math')
  m.log = lambda x: _wrap(torch.log(_t(x)))
  m.exp = lambda x: _wrap(torch.exp(_t(x)))
  m.tanh = lambda x: _wrap(torch.tanh(_t(x)))
  m.sigmoid = lambda x: _wrap(torch.sigmoid(_t(x)))
  m.rsqrt = lambda x: _wrap(torch.rsqrt(_t(x)))
  m.abs = lambda x: _wrap(torch.abs(_t(x)))
  m.is_finite = lambda x: _wrap(torch.isfinite(_t(x)))
  m.cumprod = lambda x, axis=0: _wrap(torch.cumprod(_t(x), dim=axis))
  m.reduce_mean = _reduce(torch.mean)
  m.reduce_sum = _reduce(torch.sum)
  m.reduce_prod = lambda x, axis=None, keepdims=False: _wrap(torch.prod(_t(x)) if axis is None else torch.prod(_t(x), dim=axis, keepdim=keepdims))
  m.reduce_std = lambda x, axis=None, keepdims=False: _wrap(torch.std(_t(x), unbiased=False) if axis is None else torch.std(_t(x), dim=_axes(axis, 0), unbiased=False, keepdim=keepdims))
  m.reduce_variance = lambda x, axis=None, keepdims=False: _wrap(torch.var(_t(x), unbiased=False) if axis is None else torch.var(_t(x), dim=_axes(axis, 0), unbiased=False, keepdim=keepdims))
  m.reduce_min = _minmax(torch.min)
  m.reduce_max = _minmax(torch.max)
  m.reduce_any = lambda x, axis=None: _wrap(torch.any(_t(x)))
  m.reduce_all = lambda x, axis=None: _wrap(torch.all(_t(x)))
  m.reduce_logsumexp = lambda x, axis=None, keepdims=False: _wrap(torch.logsumexp(_t(x), dim=_axes(axis, 0) if axis is not None else tuple(range(_t(x).dim())), keepdim=keepdims))
  tf.math = m

  # ---- nn
  nn = _Ns('tensorflow.nn')
  nn.sigmoid = m.sigmoid
  nn.tanh = m.tanh
  nn.softmax = lambda x, axis=-1: _wrap(torch.softmax(_t(x), dim=axis))
  nn.log_softmax = lambda x, axis=-1: _wrap(torch.log_softmax(_t(x), dim=axis))
  nn.softplus = lambda x: _wrap(F.softplus(_t(x)))
  nn.elu = lambda x: _wrap(F.elu(_t(x)))
  nn.relu = lambda x: _wrap(F.relu(_t(x)))
  nn.silu = lambda x: _wrap(F.silu(_t(x)))
  nn.swish = nn.silu
  nn.gelu = lambda x, approximate=False: _wrap(F.gelu(_t(x), approximate='tanh' if approximate else 'none'))

  def moments(x, axes, keepdims=False):
    t = _t(x)
    ax = _axes(axes, t.dim())
    return (_wrap(t.mean(dim=ax, keepdim=keepdims)), _wrap(t.var(dim=ax, unbiased=False, keepdim=keepdims)))
  nn.moments = moments
  nn.batch_normalization = lambda x, mean, var, offset, scale, eps: _wrap(
      (_t(x) - _t(mean)) * torch.rsqrt(_t(var) + eps) * _t(scale) + _t(offset))

  def conv2d(x, kernel, strides, padding):
    assert padding == 'VALID' or int(strides) == 1, 'only VALID (or stride-1 SAME) convolutions'
    t, k = _t(x).permute(0, 3, 1, 2), _t(kernel).permute(3, 2, 0, 1)          # NHWC / HWIO -> NCHW / OIHW
    pad = 0 if padding == 'VALID' else (k.shape[-1] - 1) // 2
    return _wrap(F.conv2d(t, k, stride=int(strides), padding=pad).permute(0, 2, 3, 1))
  nn.conv2d = conv2d

  def conv2d_transpose(x, kernel, output_shape, strides, padding):
    assert padding == 'VALID', 'only VALID transposed convolutions'
    t, k = _t(x).permute(0, 3, 1, 2), _t(kernel).permute(3, 2, 0, 1)          # HW(out)(in) -> (in)(out)HW
    y = F.conv_transpose2d(t, k, stride=int(strides)).permute(0, 2, 3, 1)
    assert tuple(y.shape) == tuple(int(s) for s in output_shape), (y.shape, output_shape)
    return _wrap(y)
  nn.conv2d_transpose = conv2d_transpose
  tf.nn = nn

  # ---- nest (dict keys sorted, like tf.nest)
  nest = _Ns('tensorflow.nest')

  def flatten(s):
    if isinstance(s, dict):
      return [v for k in sorted(s) for v in flatten(s[k])]
    if isinstance(s, (list, tuple)):
      return [v for item in s for v in flatten(item)]
    return [s]

  def map_structure(fn, *ss):
    s = ss[0]
    if isinstance(s, dict):
      return type(s)((k, map_structure(fn, *[x[k] for x in ss])) for k in s)
    if isinstance(s, (list, tuple)):
      out = [map_structure(fn, *[x[i] for x in ss]) for i in range(len(s))]
      return type(s)(out) if not hasattr(s, '_fields') else type(s)(*out)
    return fn(*ss)

  def pack_sequence_as(s, flat):
    flat = list(flat)

    def build(node):
      if isinstance(node, dict):
        return type(node)((k, build(node[k])) for k in sorted(node))
      if isinstance(node, (list, tuple)):
        return type(node)(build(item) for item in node)
      return flat.pop(0)
    out = build(s)
    if isinstance(s, dict):                      # restore the caller's key order
      out = type(s)((k, out[k]) for k in s)
    return out

  def assert_same_structure(a, b):
    if isinstance(a, dict):
      assert isinstance(b, dict) and sorted(a) == sorted(b), (a.keys(), b.keys() if isinstance(b, dict) else b)
      for k in a:
        assert_same_structure(a[k], b[k])
    elif isinstance(a, (list, tuple)):
      assert isinstance(b, (list, tuple)) and len(a) == len(b)
      for x, y in zip(a, b):
        assert_same_structure(x, y)
  nest.flatten, nest.map_structure = flatten, map_structure
  nest.pack_sequence_as, nest.assert_same_structure = pack_sequence_as, assert_same_structure
  tf.nest = nest

  # ---- misc namespaces
  linalg = _Ns('tensorflow.linalg')
  linalg.global_norm = lambda ts: _wrap(torch.sqrt(sum((_t(g).double() ** 2).sum() for g in ts)).float())
  tf.linalg = linalg
  debugging = _Ns('tensorflow.debugging')

  def check_numerics(x, message):
    assert _pybool(torch.isfinite(_t(x)).all()), message
    return x
  debugging.check_numerics = check_numerics
  tf.debugging = debugging
  distribute = _Ns('tensorflow.distribute')
  distribute.has_strategy = lambda: False
  distribute.get_replica_context = lambda: None
  tf.distribute = distribute
  random = _Ns('tensorflow.random')

  def categorical(logits, num_samples, seed=None):
    lg = _t(logits).detach()
    assert int(num_samples) == 1
    g = NOISE.gumbel(tuple(lg.shape))
    return _wrap(torch.argmax(lg + g, dim=-1, keepdim=True))
  random.categorical = categorical
  tf.random = random

  class GradientTape:
    """Autograd stands in for the tape: every op is recorded anyway; gradient() differentiates."""

    def __init__(self, persistent=False, **kw):
      self.persistent = persistent

    def __enter__(self):
      return self

    def __exit__(self, *exc):
      return False

    def watch(self, x):
      pass

    def gradient(self, target, sources):
      flat = nest.flatten(sources)
      # the sources themselves (as_subclass makes non-leaf aliases, which have no gradient of their own)
      grads = torch.autograd.grad(_plain(target), list(flat), retain_graph=True, allow_unused=True)
      return nest.pack_sequence_as(sources, [None if g is None else _wrap(g) for g in grads]) \
          if isinstance(sources, (dict,)) else [None if g is None else _wrap(g) for g in grads]
  tf.GradientTape = GradientTape

  keras = _Ns('tensorflow.keras')
  keras.mixed_precision = _Ns('tensorflow.keras.mixed_precision')
  tf.keras = keras
  tf.data = _Ns('tensorflow.data')
  tf.optimizers = _Ns('tensorflow.optimizers')
  return tf


# ---------------------------------------------------------------------------------------------
# tensorflow_probability.distributions
# ---------------------------------------------------------------------------------------------


def _build_tfd(tf):
  tfd = _Ns('tensorflow_probability.distributions')

  class Distribution:
    pass

  class Normal(Distribution):
    def __init__(self, loc, scale):
      self.loc = _t(loc).to(torch.float32) if isinstance(loc, torch.Tensor) else torch.tensor(float(loc))
      self.scale = _t(scale).to(torch.float32) if isinstance(scale, torch.Tensor) else torch.tensor(float(scale))
      self.batch_shape = Shape(torch.broadcast_shapes(self.loc.shape, self.scale.shape))
      self.event_shape = Shape([])

    def sample(self, sample_shape=(), seed=None):
      assert tuple(sample_shape) == ()
      eps = NOISE.normal(tuple(self.batch_shape))
      return _wrap(self.loc + self.scale * eps)

    def mean(self):
      return _wrap(self.loc + torch.zeros_like(self.scale))

    mode = mean

    def log_prob(self, value):
      v = _t(value)
      return _wrap(-0.5 * ((v - self.loc) / self.scale) ** 2 - torch.log(self.scale) - 0.5 * math.log(2 * math.pi))

    def entropy(self):
      return _wrap(0.5 + 0.5 * math.log(2 * math.pi) + torch.log(self.scale) + torch.zeros_like(self.loc))

  class Bernoulli(Distribution):
    def __init__(self, logits=None, probs=None, dtype=None):
      assert logits is not None
      self.logits = _t(logits)
      self.batch_shape = Shape(self.logits.shape)
      self.event_shape = Shape([])

    def mean(self):
      return _wrap(torch.sigmoid(self.logits))

    def mode(self):
      return _wrap((self.logits > 0).to(torch.float32))

    def log_prob(self, value):
      v = _t(value).to(self.logits.dtype)
      return _wrap(-(F.softplus(-self.logits) * v + F.softplus(self.logits) * (1 - v)))

    def entropy(self):
      p = torch.sigmoid(self.logits)
      return _wrap(F.softplus(-self.logits) * p + F.softplus(self.logits) * (1 - p))

  class OneHotCategorical(Distribution):
    def __init__(self, logits=None, probs=None, dtype=float32):
      assert (logits is None) != (probs is None)
      self._logits = None if logits is None else _t(logits)
      self._probs = None if probs is None else _t(probs)
      self.dtype = dtype
      ref = self._logits if self._logits is not None else self._probs
      self.batch_shape = Shape(ref.shape[:-1])
      self.event_shape = Shape(ref.shape[-1:])

    def logits_parameter(self):
      return _wrap(self._logits if self._logits is not None else torch.log(self._probs))

    def probs_parameter(self):
      return _wrap(self._probs if self._probs is not None else torch.softmax(self._logits, -1))

    def _logp(self):
      return torch.log_softmax(_plain(self.logits_parameter()), -1)

    def mode(self):
      lg = _plain(self.logits_parameter())
      return _wrap(F.one_hot(torch.argmax(lg, -1), lg.shape[-1]).to(self.dtype))

    def mean(self):
      return self.probs_parameter()

    def log_prob(self, value):
      return _wrap((_t(value).to(torch.float32) * self._logp()).sum(-1))

    def entropy(self):
      lp = self._logp()
      return _wrap(-(torch.exp(lp) * lp).sum(-1))

    def sample(self, sample_shape=(), seed=None):
      lg = _plain(self.logits_parameter()).detach()
      g = NOISE.gumbel(tuple(lg.shape))
      return _wrap(F.one_hot(torch.argmax(lg + g, -1), lg.shape[-1]).to(self.dtype))

  class Independent(Distribution):
    def __init__(self, distribution, reinterpreted_batch_ndims=None):
      self.distribution = distribution
      n = int(reinterpreted_batch_ndims)
      self._n = n
      bs = list(distribution.batch_shape)
      self.batch_shape = Shape(bs[:len(bs) - n])
      self.event_shape = Shape(bs[len(bs) - n:] + list(distribution.event_shape))

    def _sum(self, x):
      t = _plain(x)
      return _wrap(t.sum(dim=tuple(range(-self._n, 0))) if self._n else t)

    def log_prob(self, value):
      return self._sum(self.distribution.log_prob(value))

    def entropy(self):
      return self._sum(self.distribution.entropy())

    def sample(self, sample_shape=(), seed=None):
      return self.distribution.sample(sample_shape, seed)

    def mode(self):
      return self.distribution.mode()

    def mean(self):
      return self.distribution.mean()

  def kl_divergence(a, b):
    assert isinstance(a, Independent) and isinstance(b, Independent) and a._n == b._n
    pa, pb = a.distribution, b.distribution
    assert isinstance(pa, OneHotCategorical) and isinstance(pb, OneHotCategorical)
    la, lb = pa._logp(), pb._logp()
    kl = (torch.exp(la) * (la - lb)).sum(-1)
    return a._sum(_wrap(kl))

  def _missing(name):
    class Missing(Distribution):
      def __init__(self, *a, **k):
        raise NotImplementedError(f'tfd.{name} is not part of the shim')
    Missing.__name__ = name
    return Missing
  tfd.Distribution, tfd.Normal, tfd.Bernoulli = Distribution, Normal, Bernoulli
  tfd.OneHotCategorical, tfd.Independent, tfd.kl_divergence = OneHotCategorical, Independent, kl_divergence
  for name in ('Uniform', 'TruncatedNormal', 'MultivariateNormalDiag', 'Deterministic'):
    setattr(tfd, name, _missing(name))
  return tfd


# ---------------------------------------------------------------------------------------------
# sonnet
# ---------------------------------------------------------------------------------------------


def _build_snt():
  snt = _Ns('sonnet')

  def walk(obj, seen, out, trainable_only):
    if id(obj) in seen:
      return
    if isinstance(obj, Variable):
      seen.add(id(obj))
      if obj.trainable or not trainable_only:
        out.append(obj)
    elif isinstance(obj, Module):
      seen.add(id(obj))
      for k in sorted(vars(obj)):
        if k in ('_variables_cache',):
          continue
        walk(vars(obj)[k], seen, out, trainable_only)
    elif isinstance(obj, dict):
      for k in sorted(obj, key=str):
        walk(obj[k], seen, out, trainable_only)
    elif isinstance(obj, (list, tuple)):
      for v in obj:
        walk(v, seen, out, trainable_only)

  class Module:
    def __init__(self, name=None):
      self._snt_name = name or type(self).__name__

    @property
    def name(self):
      return self._snt_name

    @property
    def variables(self):
      out = []
      walk(self, set(), out, False)
      return tuple(out)

    @property
    def trainable_variables(self):
      out = []
      walk(self, set(), out, True)
      return tuple(out)
  snt.Module = Module
  return snt


# ---------------------------------------------------------------------------------------------
# installation
# ---------------------------------------------------------------------------------------------


def install():
  """Put the stand-ins into sys.modules (idempotent).  Returns the fake `tensorflow` module."""
  if 'tensorflow' in sys.modules and getattr(sys.modules['tensorflow'], '_dd_shim', False):
    return sys.modules['tensorflow']
  assert 'tensorflow' not in sys.modules, 'a real tensorflow is already imported'
  tf = _build_tf()
  tf._dd_shim = True
  tfd = _build_tfd(tf)
  tfp = _Ns('tensorflow_probability')
  tfp.distributions = tfd
  snt = _build_snt()
  py = _Ns('tensorflow.python')
  dist = _Ns('tensorflow.python.distribute')
  values = _Ns('tensorflow.python.distribute.values')

  class PerReplica:
    pass
  values.PerReplica = PerReplica
  dist.values = values
  py.distribute = dist
  tf.python = py
  import yaml as pyyaml
  ruamel = _Ns('ruamel')
  ryaml = _Ns('ruamel.yaml')

  import re

  class Loader12(pyyaml.SafeLoader):
    """YAML 1.2 scalars as ruamel reads them: `off` / `no` stay strings, `1e-4` is a float."""
  Loader12.yaml_implicit_resolvers = {
      ch: [(tag, rx) for tag, rx in lst if tag not in ('tag:yaml.org,2002:bool', 'tag:yaml.org,2002:float')]
      for ch, lst in pyyaml.SafeLoader.yaml_implicit_resolvers.items()}
  Loader12.add_implicit_resolver(
      'tag:yaml.org,2002:bool', re.compile(r'^(?:true|True|TRUE|false|False|FALSE)$'), list('tTfF'))
  Loader12.add_implicit_resolver(
      'tag:yaml.org,2002:float',
      re.compile(r'^[-+]?(?:\.[0-9]+|[0-9]+(?:\.[0-9]*)?)(?:[eE][-+]?[0-9]+)?$|^[-+]?\.(?:inf|Inf|INF)$|^\.(?:nan|NaN|NAN)$'),
      list('-+0123456789.'))

  class YAML:
    def __init__(self, typ='safe'):
      pass

    def load(self, text):
      return pyyaml.load(text, Loader=Loader12)
  ryaml.YAML = YAML
  ruamel.yaml = ryaml
  mods = {
      'tensorflow': tf, 'tensorflow.math': tf.math, 'tensorflow.nn': tf.nn, 'tensorflow.nest': tf.nest,
      'tensorflow.keras': tf.keras, 'tensorflow.keras.mixed_precision': tf.keras.mixed_precision,
      'tensorflow.python': py, 'tensorflow.python.distribute': dist,
      'tensorflow.python.distribute.values': values,
      'tensorflow_probability': tfp, 'tensorflow_probability.distributions': tfd,
      'sonnet': snt, 'sonnet.v2': snt, 'ruamel': ruamel, 'ruamel.yaml': ryaml,
  }
  sys.modules.update(mods)
  if 'gym' not in sys.modules:                  # the reference's embodied package imports gym at module level
    gym = _Ns('gym')
    gym.spaces = _Ns('gym.spaces')
    sys.modules['gym'] = gym
    sys.modules['gym.spaces'] = gym.spaces
  return tf


@contextlib.contextmanager
def argv0(path):
  """agent.py finds configs.yaml next to sys.argv[0]."""
  old = sys.argv[0]
  sys.argv[0] = str(path)
  try:
    yield
  finally:
    sys.argv[0] = old
