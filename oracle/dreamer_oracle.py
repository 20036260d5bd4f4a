"""CPU oracle for the DreamerV2+ learner hot path  --  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (danijar/daydreamer, TF2 + TFP + sonnet) cannot
be imported in this image (tensorflow / tensorflow_probability / sonnet are not
installed, no network) and its own tests hold no golden vectors for this path
(SURVEY.md section 4).  This file is therefore a line-by-line restatement of the
reference arithmetic in PyTorch-CPU fp32 with explicit weights and explicit
noise tensors; it is anchored by closed-form known answers (tests/) and by
torch.autograd for every gradient.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this module.  The product path (daydreamer_b200/) never
does.

Path aliases:  A/ = /root/reference/embodied/agents/dreamerv2plus/

Conventions
  * all tensors fp32, row-major; LayerNorm eps = 1e-3 over the last axis
    (A/nets.py:594-600); Linear has no bias when a norm follows (A/nets.py:563).
  * categorical draws: idx = argmax(logit + g) with g injected Gumbel noise
    (tf.random.categorical is Gumbel-max; its stream is irreproducible so the
    noise is an explicit input here -- SURVEY.md section 7 "Sampling parity").
  * continuous actions: a = mean + std * eps with eps injected N(0,1).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-3


# ----------------------------------------------------------------------------
# Spec: sizes + hyper-parameters (resolved from A/configs.yaml)
# ----------------------------------------------------------------------------

@dataclasses.dataclass
class Spec:
  obs_dims: Dict[str, int]            # encoder/decoder vector keys -> dim
  act_dim: int = 12
  discrete: bool = False
  deter: int = 512
  units: int = 512
  stoch: int = 32
  classes: int = 32
  prior_layers: int = 3               # A/configs.yaml:53
  rssm_unimix: float = 0.01
  enc_layers: int = 4
  enc_units: int = 512
  head_layers: int = 4                # reward / cont / actor / critic (and decoder MLP unless set below)
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
  actor_unimix: float = 0.1           # discrete actor only
  lr: float = 1e-4
  adam_eps: float = 1e-6
  clip: float = 100.0
  wd: float = 1e-2
  beta1: float = 0.9
  beta2: float = 0.999
  slow_every: int = 100
  advnorm_decay: float = 0.99
  advnorm_max: float = 1e8
  # image keys -> (H, W, C); channel-concatenated for the CNN (A/nets.py:218-222)
  cnn_keys: Dict[str, Tuple[int, int, int]] = dataclasses.field(default_factory=dict)
  cnn_depth: int = 64                         # A/configs.yaml:54-55
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
    """(H, W, sum of channels) of the merged image, or None without cnn keys."""
    if not self.cnn_keys:
      return None
    shapes = [tuple(v) for v in self.cnn_keys.values()]
    assert all(x[:2] == shapes[0][:2] for x in shapes), shapes
    return shapes[0][:2] + (sum(x[2] for x in shapes),)

  @property
  def enc_conv_shapes(self):
    """[(Hin, Win, Cin, k, Hout, Wout, Cout)] of ImageEncoderSimple: stride 2,
    VALID, depth doubling (A/nets.py:298-305)."""
    out = []
    if not self.cnn_keys:
      return out
    h, w, c = self.img_shape
    depth = self.cnn_depth
    for k in self.enc_kernels:
      ho, wo = (h - k) // 2 + 1, (w - k) // 2 + 1
      out.append((h, w, c, k, ho, wo, depth))
      h, w, c, depth = ho, wo, depth, depth * 2
    return out

  @property
  def dec_conv_shapes(self):
    """Same tuples for ImageDecoderSimple: transposed, stride 2, VALID, out =
    in*2 + k - 2, starting from a 1x1 map of the features (A/nets.py:316-327,529-533)."""
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
    assert (h, w, c) == tuple(self.img_shape), ((h, w, c), self.img_shape)
    return out

  @property
  def cnn_out(self):
    sh = self.enc_conv_shapes
    return sh[-1][4] * sh[-1][5] * sh[-1][6] if sh else 0

  @property
  def embed(self):
    """Encoder output width: [cnn features, mlp features] (A/nets.py:220-231)."""
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


def a1_spec(**kw) -> Spec:
  """BASELINE.json configs[1]: a1 proprio, deter=units=512, stoch 32x32, A=12."""
  return Spec(obs_dims={'vector': 16}, **kw)


def tiny_spec(**kw) -> Spec:
  base = dict(
      obs_dims={'vector': 5}, act_dim=3, deter=24, units=20, stoch=4,
      classes=6, enc_layers=2, enc_units=16, head_layers=2, head_units=16,
      horizon=3, prior_layers=3)
  base.update(kw)
  return Spec(**base)


# ----------------------------------------------------------------------------
# Parameters (names mirror the reference's self.get() names; shapes SURVEY A.1)
# ----------------------------------------------------------------------------

def _uniform(rng, fan_in, fan_out, outscale=1.0):
  # A/nets.py:568-569  limit = sqrt(3 * outscale / mean(fan_in, fan_out))
  limit = math.sqrt(3.0 * outscale / ((fan_in + fan_out) / 2.0))
  return rng.uniform(-limit, limit, (fan_in, fan_out)).astype(np.float32)


def param_shapes(spec: Spec) -> Dict[str, Tuple[int, ...]]:
  """Ordered dict name -> shape of every trainable tensor + slow critic."""
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
    s[f'enc/conv{i}/kernel'] = (k, k, ci, co)                 # HWIO, A/nets.py:541
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
    s[f'{name}/kernel'] = (k, k, co, ci)                      # [k,k,out,in], A/nets.py:523
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


GROUPS = {
    'model': ('enc/', 'rssm/', 'dec/', 'reward/', 'cont/'),
    'actor': ('actor/',),
    'critic': ('critic/',),
    'slow_critic': ('slow_critic/',),
}


def group_names(spec: Spec, group: str):
  pre = GROUPS[group]
  return [k for k in param_shapes(spec) if k.startswith(pre)]


def init_params(spec: Spec, seed=1) -> Dict[str, np.ndarray]:
  """Reference init rule, drawn in name order from default_rng(seed)."""
  rng = np.random.default_rng(seed)
  out = {}
  for name, shape in param_shapes(spec).items():
    if name.endswith('/kernel') and len(shape) == 4:
      # Conv2D: U(+-sqrt(3/mean(in,out))) (A/nets.py:542-543); transposed:
      # U(+-sqrt(3/(k*k*mean(out,in)))) (A/nets.py:524-525)
      spatial = shape[0] * shape[1] if name.startswith('dec/') else 1
      limit = math.sqrt(3.0 / (spatial * (shape[2] + shape[3]) / 2.0))
      out[name] = rng.uniform(-limit, limit, shape).astype(np.float32)
    elif name.endswith('/kernel'):
      # DistLayer 'out' has outscale 0.1, 'std' 1.0 (A/nets.py:448-454);
      # rssm stats layers use plain Linear (outscale 1.0, A/nets.py:164).
      outscale = 0.1 if name.endswith('dist_out/out/kernel') else 1.0
      out[name] = _uniform(rng, shape[0], shape[1], outscale)
    elif name.endswith('norm/scale'):
      out[name] = np.ones(shape, np.float32)
    else:
      out[name] = np.zeros(shape, np.float32)
  return out


def is_decayed(name: str) -> bool:
  # A/tfutils.py:292  re.search('kernel', optimizer_name + '/' + var.name)
  return 'kernel' in name


# ----------------------------------------------------------------------------
# Primitive math
# ----------------------------------------------------------------------------

def layer_norm(x, scale, bias):
  mean = x.mean(-1, keepdim=True)
  var = ((x - mean) ** 2).mean(-1, keepdim=True)      # population variance
  return (x - mean) * torch.rsqrt(var + LN_EPS) * scale + bias


def dense(x, p, name, act=True):
  """Linear(no bias) + LayerNorm + elu  (A/nets.py:566-582)."""
  z = x @ p[f'{name}/kernel']
  z = layer_norm(z, p[f'{name}/norm/scale'], p[f'{name}/norm/bias'])
  return F.elu(z) if act else z


def linear_bias(x, p, name):
  return x @ p[f'{name}/kernel'] + p[f'{name}/bias']


def mlp_trunk(x, p, prefix, layers):
  for i in range(layers):
    x = dense(x, p, f'{prefix}/dense{i}')
  return x


def symlog(x):
  return torch.sign(x) * torch.log1p(torch.abs(x))


def symexp(x):
  return torch.sign(x) * torch.expm1(torch.abs(x))


def unimix_logits(z, classes, mix):
  """A/nets.py:165-170: log((1-m) softmax(z) + m/C) over the class axis."""
  z = z.reshape(z.shape[:-1] + (-1, classes))
  probs = torch.softmax(z, -1)
  probs = (1.0 - mix) * probs + mix / classes
  return torch.log(probs)


def sample_onehot(logit, gumbel):
  """A/tfutils.py:368-382: Gumbel-max index, one-hot, straight-through."""
  idx = torch.argmax(logit.detach() + gumbel, -1)
  onehot = F.one_hot(idx, logit.shape[-1]).to(logit.dtype)
  probs = torch.softmax(logit, -1)
  return onehot + (probs - probs.detach()), idx


def mode_onehot(logit):
  idx = torch.argmax(logit, -1)
  return F.one_hot(idx, logit.shape[-1]).to(logit.dtype), idx


def categorical_kl(lp, lq):
  """tfd.kl_divergence(OneHotCategorical(lp), OneHotCategorical(lq)) summed
  over the stoch axis (tfd.Independent(.., 1), A/nets.py:91)."""
  a = torch.log_softmax(lp, -1)
  b = torch.log_softmax(lq, -1)
  return (torch.exp(a) * (a - b)).sum(-1).sum(-1)


def categorical_entropy(logit):
  a = torch.log_softmax(logit, -1)
  return -(torch.exp(a) * a).sum(-1)


# ----------------------------------------------------------------------------
# RSSM (A/nets.py:11-183)
# ----------------------------------------------------------------------------

def gru(x, deter, p):
  """A/nets.py:149-160 (joint LN over 3D gates, update bias -1)."""
  y = dense(torch.cat([deter, x], -1), p, 'rssm/gru_out', act=False)
  r, c, u = torch.chunk(y, 3, -1)
  r = torch.sigmoid(r)
  c = torch.tanh(r * c)
  u = torch.sigmoid(u - 1.0)
  return u * c + (1.0 - u) * deter


def prior_logits(deter, p, spec: Spec):
  x = deter
  for i in range(spec.prior_layers):
    x = dense(x, p, f'rssm/img_out_{i}')
  z = linear_bias(x, p, 'rssm/img_stats')
  return unimix_logits(z, spec.classes, spec.rssm_unimix)


def img_step(state, action, p, spec: Spec, gumbel):
  """A/nets.py:119-138.  gumbel: [M,S,C] noise for the prior draw, or None to
  skip the (unused) draw and return the mode-free dict without 'stoch'."""
  stoch = state['stoch'].reshape(state['stoch'].shape[:-2] + (spec.sc,))
  x = dense(torch.cat([stoch, action], -1), p, 'rssm/img_in')
  deter = gru(x, state['deter'], p)
  logit = prior_logits(deter, p, spec)
  out = {'deter': deter, 'logit': logit}
  if gumbel is not None:
    out['stoch'], out['idx'] = sample_onehot(logit, gumbel)
  return out


def initial_state(m, p, spec: Spec):
  """'learned2' (A/nets.py:54-62,140-147): deter = tanh(v); stoch = mode of the
  prior at that deter (one-hot argmax, no gradient); logit = 0."""
  deter = torch.tanh(p['rssm/initial_deter'])[None].repeat(m, 1)
  stoch, _ = mode_onehot(prior_logits(deter, p, spec))
  logit = torch.zeros(m, spec.stoch, spec.classes)
  return {'deter': deter, 'stoch': stoch.detach(), 'logit': logit}


def obs_step(prev, action, embed, first, p, spec: Spec, g_post, g_prior=None):
  """A/nets.py:99-117."""
  m = len(first)
  keep = (1.0 - first)
  init = initial_state(m, p, spec)
  prev = {k: prev[k] * keep.reshape((m,) + (1,) * (prev[k].dim() - 1)) +
          init[k] * first.reshape((m,) + (1,) * (prev[k].dim() - 1))
          for k in ('deter', 'stoch', 'logit')}
  action = action * keep[:, None]
  prior = img_step(prev, action, p, spec, g_prior)
  x = dense(torch.cat([prior['deter'], embed], -1), p, 'rssm/obs_out')
  logit = unimix_logits(
      linear_bias(x, p, 'rssm/obs_stats'), spec.classes, spec.rssm_unimix)
  stoch, idx = sample_onehot(logit, g_post)
  post = {'stoch': stoch, 'deter': prior['deter'], 'logit': logit, 'idx': idx}
  return post, prior


def observe(embed, action, first, state, p, spec: Spec, g_post):
  """A/nets.py:66-76.  embed [B,L,E], action [B,L,A], first [B,L];
  g_post [L,B,S,C].  Returns post, prior stacked batch-major [B,L,...]."""
  B, L = first.shape
  if state is None:
    state = initial_state(B, p, spec)
  posts, priors = [], []
  for t in range(L):
    post, prior = obs_step(
        state, action[:, t], embed[:, t], first[:, t], p, spec, g_post[t])
    posts.append(post)
    priors.append(prior)
    state = {k: post[k] for k in ('deter', 'stoch', 'logit')}
  stack = lambda xs, k: torch.stack([x[k] for x in xs], 1)
  post = {k: stack(posts, k) for k in ('deter', 'stoch', 'logit', 'idx')}
  prior = {k: stack(priors, k) for k in ('deter', 'logit')}
  return post, prior


def kl_loss(post, prior, balance):
  """A/nets.py:178-183."""
  lhs = categorical_kl(post['logit'].detach(), prior['logit'])
  rhs = categorical_kl(post['logit'], prior['logit'].detach())
  return balance * lhs + (1.0 - balance) * rhs


# ----------------------------------------------------------------------------
# Heads
# ----------------------------------------------------------------------------

def feat_of(state, spec: Spec):
  st = state['stoch']
  return torch.cat(
      [state['deter'], st.reshape(st.shape[:-2] + (spec.sc,))], -1)


def head(feat, p, prefix, spec: Spec, out='out'):
  x = mlp_trunk(feat, p, prefix, spec.trunk('critic' if prefix == 'slow_critic' else prefix)[0])
  return linear_bias(x, p, f'{prefix}/dist_out/{out}')


def conv_layer(x, p, name, transp=False, norm=True):
  """Conv2D(stride 2, VALID) + bias [+ LayerNorm over channels + elu] on NHWC
  (A/nets.py:495-554).  Forward kernel HWIO; transposed kernel [k,k,out,in] with
  tf.nn.conv2d_transpose semantics: out[ho,wo,co] = sum x[hi,wi,ci] K[ho-2hi, wo-2wi, co, ci]."""
  w = p[f'{name}/kernel']
  x = x.permute(0, 3, 1, 2)
  if transp:
    x = F.conv_transpose2d(x, w.permute(3, 2, 0, 1), stride=2)
  else:
    x = F.conv2d(x, w.permute(3, 2, 0, 1), stride=2)
  x = x.permute(0, 2, 3, 1) + p[f'{name}/bias']
  if norm:
    x = F.elu(layer_norm(x, p[f'{name}/norm/scale'], p[f'{name}/norm/bias']))
  return x


def image_encoder(x, p, spec: Spec):
  """ImageEncoderSimple (A/nets.py:291-305): x [M,H,W,C] -> [M, cnn_out] (h,w,c order)."""
  for i in range(len(spec.enc_kernels)):
    x = conv_layer(x, p, f'enc/conv{i}')
  return x.reshape(x.shape[0], -1)


def image_decoder(feat, p, spec: Spec):
  """ImageDecoderSimple (A/nets.py:308-327): feat [M,F] -> sigmoid image [M,H,W,C]."""
  x = feat.reshape(feat.shape[0], 1, 1, feat.shape[-1])
  n = len(spec.dec_kernels)
  for i in range(n - 1):
    x = conv_layer(x, p, f'dec/conv{i}', transp=True)
  x = conv_layer(x, p, 'dec/conv_out', transp=True, norm=False)
  return torch.sigmoid(x)


def encoder(data, p, spec: Spec):
  """MultiEncoder (A/nets.py:212-232): [cnn(channel-concat images), mlp(concat vectors)]."""
  outs = []
  lead = data['is_first'].shape if 'is_first' in data else None
  if spec.cnn_keys:
    img = torch.cat([data[k] for k in spec.cnn_keys], -1)
    lead = img.shape[:-3]
    outs.append(image_encoder(img.reshape((-1,) + img.shape[-3:]), p, spec).reshape(lead + (-1,)))
  if spec.obs_dims:
    x = torch.cat([data[k].reshape(data[k].shape[:2] + (-1,))
                   for k in spec.obs_dims], -1)
    outs.append(mlp_trunk(x, p, 'enc', spec.enc_layers))
  return torch.cat(outs, -1)


def actor_dist(feat, p, spec: Spec):
  """A/nets.py:461-468 ('normal') / :480-491 ('onehot')."""
  x = mlp_trunk(feat, p, 'actor', spec.head_layers)
  out = linear_bias(x, p, 'actor/dist_out/out')
  if spec.discrete:
    return {'logit': unimix_logits(out, spec.act_dim, spec.actor_unimix)[..., 0, :]}
  std = linear_bias(x, p, 'actor/dist_out/std')
  lo, hi = spec.actor_minstd, spec.actor_maxstd
  return {'mean': torch.tanh(out), 'std': (hi - lo) * torch.sigmoid(std) + lo}


# ----------------------------------------------------------------------------
# Controllers (A/tfutils.py:414-527)
# ----------------------------------------------------------------------------

def init_ctrl(spec: Spec):
  z64 = lambda: torch.zeros((), dtype=torch.float64)
  ctrl = {
      'wmkl_scale': torch.ones(()),
      'actent_scale': torch.ones(() if spec.discrete else (spec.act_dim,)),
      'slow_updates': torch.tensor(-1, dtype=torch.int64),
  }
  for n in ('advnorm', 'retnorm', 'scorenorm'):
    ctrl[f'{n}_mean'] = z64()
    ctrl[f'{n}_sqrs'] = z64()
    ctrl[f'{n}_step'] = torch.tensor(0, dtype=torch.int64)
  return ctrl


def autoadapt_update(scale, avg, target, lo, hi, inverse, spec: Spec):
  """'mult' controller, A/tfutils.py:464-474."""
  below = avg < target / (1.0 + spec.ctrl_thres)
  above = avg > target * (1.0 + spec.ctrl_thres)
  if inverse:
    below, above = above, below
  inside = ~below & ~above
  adj = (above.float() * scale * (1.0 + spec.ctrl_vel) +
         below.float() * scale / (1.0 + spec.ctrl_vel) +
         inside.float() * scale)
  return torch.clamp(adj, lo, hi)


def normalize_update(ctrl, name, x, decay):
  x = x.detach().double()
  ctrl[f'{name}_step'] = ctrl[f'{name}_step'] + 1
  ctrl[f'{name}_mean'] = decay * ctrl[f'{name}_mean'] + (1 - decay) * x.mean()
  ctrl[f'{name}_sqrs'] = decay * ctrl[f'{name}_sqrs'] + (1 - decay) * (x ** 2).mean()


def normalize_stats(ctrl, name, decay, maxv):
  corr = 1.0 - decay ** ctrl[f'{name}_step'].double()
  mean = ctrl[f'{name}_mean'] / corr
  var = ctrl[f'{name}_sqrs'] / corr - mean ** 2
  scale = torch.rsqrt(torch.clamp(var, min=1.0 / maxv ** 2))
  return mean.float(), scale.float()


# ----------------------------------------------------------------------------
# Optimizer (A/tfutils.py:180-302)
# ----------------------------------------------------------------------------

def init_opt(params, names):
  return {'step': 0,
          'm': {k: torch.zeros_like(params[k]) for k in names},
          'v': {k: torch.zeros_like(params[k]) for k in names}}


def adam_update(params, grads, opt, names, spec: Spec):
  """Global-norm clip -> weight decay on kernels -> bias-corrected Adam.
  Mutates params/opt in place; returns the pre-clip global norm."""
  norm = torch.sqrt(sum((grads[k].double() ** 2).sum() for k in names)).float()
  scale = torch.clamp(spec.clip / norm, max=1.0) if spec.clip else 1.0
  opt['step'] += 1
  t = float(opt['step'])
  with torch.no_grad():
    for k in names:
      g = grads[k] * scale
      if spec.wd and is_decayed(k):
        params[k].mul_(1.0 - spec.wd * spec.lr)
      opt['m'][k].mul_(spec.beta1).add_(g, alpha=1.0 - spec.beta1)
      opt['v'][k].mul_(spec.beta2).addcmul_(g, g, value=1.0 - spec.beta2)
      mhat = opt['m'][k] / (1.0 - spec.beta1 ** t)
      vhat = opt['v'][k] / (1.0 - spec.beta2 ** t)
      params[k].sub_(spec.lr * mhat / (torch.sqrt(vhat) + spec.adam_eps))
  return norm


# ----------------------------------------------------------------------------
# lambda-return (A/agent.py:422-442, 'gve')
# ----------------------------------------------------------------------------

def lambda_return(reward, value, disc, lam):
  """reward [H,N] (r_1..r_H), value [H+1,N], disc [H,N] (= gamma*cont_1..H).
  R_H = V_H;  R_t = r_{t+1} + d_{t+1} ((1-lam) V_{t+1} + lam R_{t+1})."""
  vals = [value[-1]]
  interm = reward + disc * value[1:] * (1.0 - lam)
  for t in reversed(range(disc.shape[0])):
    vals.append(interm[t] + disc[t] * lam * vals[-1])
  return torch.stack(list(reversed(vals))[:-1], 0)


def balance_stats(mean, loss, target, thres):
  """A/tfutils.py:395-411."""
  pos = (target > thres).float()
  neg = (target <= thres).float()
  pred = (mean > thres).float()
  return dict(
      pos_loss=(loss * pos).sum() / pos.sum(),
      neg_loss=(loss * neg).sum() / neg.sum(),
      pos_acc=(pred * pos).sum() / pos.sum(),
      neg_acc=((1 - pred) * neg).sum() / neg.sum(),
      rate=pos.mean(), avg=target.mean(), pred=mean.mean())


def _std(x):
  return x.std(unbiased=False)


# ----------------------------------------------------------------------------
# The learner state and the full train step (A/agent.py:67-93)
# ----------------------------------------------------------------------------

class Learner:
  """Holds params / optimizer / controller state; train_step() is one
  Agent.train call of the reference restated with explicit noise."""

  def __init__(self, spec: Spec, params: Dict[str, np.ndarray] = None, seed=1):
    self.spec = spec
    params = params if params is not None else init_params(spec, seed)
    self.params = {k: torch.tensor(np.asarray(v), dtype=torch.float32)
                   for k, v in params.items()}
    self.names = {g: group_names(spec, g) for g in GROUPS}
    self.opts = {g: init_opt(self.params, self.names[g])
                 for g in ('model', 'actor', 'critic')}
    self.ctrl = init_ctrl(spec)

  # -- helpers --------------------------------------------------------------
  def _leaf(self, names):
    p = dict(self.params)
    for k in names:
      p[k] = self.params[k].clone().requires_grad_(True)
    return p

  @staticmethod
  def preprocess(batch):
    """A/agent.py:123-139 (reward transform 'off')."""
    data = {}
    for k, v in batch.items():
      if k.startswith('log_') or k == 'key':
        continue
      v = np.asarray(v)
      t = torch.as_tensor(v).float()
      data[k] = t / 255.0 if (v.ndim > 3 and v.dtype == np.uint8) else t
    data['cont'] = 1.0 - data['is_terminal']
    return data

  # -- acting ------------------------------------------------------------------
  def policy(self, obs, state, g_post, eps_act, mode='train'):
    """Agent.policy (A/agent.py:42-65) with expl_noise = eval_noise = 0: preprocess -> encoder
    -> ONE obs_step from the carried (latent, previous action) -> actor; `train` samples the
    action, `eval` takes the mode.  Explicit noise: g_post [P, S, C] Gumbel for the posterior draw,
    eps_act [P, A] standard normal (continuous actor) or Gumbel (one-hot actor).
    state: None (initial_policy_state: rssm.initial + zero action, A/agent.py:33-38) or the dict
    returned by the previous call.  Returns ({'action': [P, A]}, state)."""
    with torch.no_grad():
      p, spec = self.params, self.spec
      data = self.preprocess({k: np.asarray(v)[:, None] for k, v in obs.items()})   # [P, 1, ...]
      P = data['is_first'].shape[0]
      embed = encoder(data, p, spec)[:, 0]
      if state is None:
        latent = initial_state(P, p, spec)
        action = torch.zeros(P, spec.act_dim)
      else:
        latent = {k: torch.as_tensor(np.asarray(state[k])).float() for k in ('deter', 'stoch', 'logit')}
        latent['stoch'] = latent['stoch'].reshape(P, spec.stoch, spec.classes)
        latent['logit'] = latent['logit'].reshape(P, spec.stoch, spec.classes)
        action = torch.as_tensor(np.asarray(state['action'])).float()
      post, _ = obs_step(latent, action, embed, data['is_first'][:, 0], p, spec,
                         torch.as_tensor(np.asarray(g_post)).float())
      dist = actor_dist(feat_of(post, spec), p, spec)
      if spec.discrete:
        logit = dist['logit'][:, None, :]
        eps = torch.as_tensor(np.asarray(eps_act)).float()[:, None, :]
        act = (mode_onehot(logit)[0] if mode == 'eval' else sample_onehot(logit, eps)[0])[:, 0]
      else:
        eps = torch.as_tensor(np.asarray(eps_act)).float()
        act = dist['mean'] if mode == 'eval' else dist['mean'] + dist['std'] * eps
      state = {'deter': post['deter'].numpy(), 'stoch': post['stoch'].numpy().reshape(P, -1),
               'logit': post['logit'].numpy().reshape(P, -1), 'action': act.numpy()}
      return {'action': act.numpy()}, state

  # -- world model ------------------------------------------------------------
  def wm_loss(self, p, data, state, g_post, training=True):
    """A/agent.py:165-212.  Returns loss, aux dict, metrics."""
    spec = self.spec
    mets = {}
    embed = encoder(data, p, spec)
    post, prior = observe(
        embed, data['action'], data['is_first'], state, p, spec, g_post)
    feat = feat_of(post, spec)
    dec = recon = None
    if spec.dec_vec:
      target = torch.cat([data[k].reshape(data[k].shape[:2] + (-1,))
                          for k in spec.dec_vec], -1)
      dec = head(feat, p, 'dec', spec)
    if spec.cnn_keys:
      recon = image_decoder(feat.reshape(-1, spec.feat), p, spec)
      recon = recon.reshape(feat.shape[:2] + recon.shape[1:])
    rew = head(feat, p, 'reward', spec)[..., 0]
    con = head(feat, p, 'cont', spec)[..., 0]
    losses = {}
    kl = kl_loss(post, prior, spec.kl_balance)
    if training:
      self.ctrl['wmkl_scale'] = autoadapt_update(
          self.ctrl['wmkl_scale'], kl.detach().mean(), spec.wmkl_target,
          spec.wmkl_min, spec.wmkl_max, False, spec)
    scale = self.ctrl['wmkl_scale']
    losses['kl'] = scale * kl
    mets.update(wmkl_mean=kl.mean(), wmkl_std=_std(kl),
                wmkl_scale_mean=scale.mean(), wmkl_scale_std=scale.std(unbiased=False) if scale.dim() else torch.zeros(()))
    off = 0
    for k, shp in spec.cnn_keys.items():       # MSEDist(mean, 3, 'sum'), A/nets.py:283-288
      c = shp[2]
      losses[k] = ((recon[..., off:off + c] - data[k]) ** 2).sum((-1, -2, -3))
      off += c
    off = 0
    for k, d in spec.dec_vec.items():
      losses[k] = ((dec[..., off:off + d] - target[..., off:off + d]) ** 2).sum(-1)
      off += d
    losses['reward'] = (rew - symlog(data['reward'])) ** 2
    losses['cont'] = F.softplus(con) - data['cont'] * con
    for k, v in losses.items():
      mets[f'{k}_loss_mean'] = v.mean()
      mets[f'{k}_loss_std'] = _std(v)
    model_loss = sum(losses.values())
    mets['prior_ent_mean'] = categorical_entropy(prior['logit']).sum(-1).mean()
    mets['post_ent_mean'] = categorical_entropy(post['logit']).sum(-1).mean()
    mets['prior_ent_min'] = categorical_entropy(prior['logit']).sum(-1).min()
    mets['post_ent_min'] = categorical_entropy(post['logit']).sum(-1).min()
    mets['model_loss_mean'] = model_loss.mean()
    mets['model_loss_std'] = _std(model_loss)
    st = balance_stats(symexp(rew), losses['reward'], data['reward'], 0.1)
    mets.update({f'reward_{k}': v for k, v in st.items()})
    st = balance_stats(torch.sigmoid(con), losses['cont'], data['cont'], 0.5)
    mets.update({f'cont_{k}': v for k, v in st.items()})
    aux = {'embed': embed, 'post': post, 'prior': prior, 'dec': dec, 'recon': recon,
           'reward': rew, 'cont': con, 'kl': kl, 'losses': losses}
    return model_loss.mean(), aux, mets

  # -- imagination -----------------------------------------------------------
  def imagine(self, p, start, cont0, g_img, eps_act):
    """A/agent.py:234-261.  start: dict of [N,...] (constants).  g_img
    [H,N,S,C]; eps_act [H+1,N,A] (continuous) or Gumbel [H+1,N,A] (discrete).
    Returns traj dict, time-major [H+1,N,...]."""
    spec = self.spec
    H = spec.horizon

    def policy(state, t):
      dist = actor_dist(feat_of({k: v.detach() for k, v in state.items()
                                 if k in ('deter', 'stoch')}, spec), p, spec)
      if spec.discrete:
        a, _ = sample_onehot(dist['logit'], eps_act[t])
      else:
        a = dist['mean'] + dist['std'] * eps_act[t]
      return a

    state = {k: start[k] for k in ('deter', 'stoch', 'logit')}
    states, actions = [state], [policy(state, 0)]
    for t in range(H):
      nxt = img_step(states[-1], actions[-1], p, spec, g_img[t])
      state = {k: nxt[k] for k in ('deter', 'stoch', 'logit')}
      states.append(state)
      actions.append(policy(state, t + 1))
    traj = {k: torch.stack([s[k] for s in states], 0)
            for k in ('deter', 'stoch', 'logit')}
    traj['action'] = torch.stack(actions, 0)
    feat = feat_of(traj, spec)
    cont = torch.sigmoid(head(feat, p, 'cont', spec)[..., 0])
    traj['cont'] = torch.cat([cont0[None], cont[1:]], 0)
    traj['weight'] = torch.cumprod(spec.discount * traj['cont'], 0) / spec.discount
    traj['feat'] = feat
    return traj

  def _target(self, p, traj, slow_prefix):
    """VFunction.target 'gve' with rewfn = symexp(reward_head)[1:]."""
    spec = self.spec
    reward = symexp(head(traj['feat'], p, 'reward', spec)[..., 0])[1:]
    value = symexp(head(traj['feat'], p, slow_prefix, spec)[..., 0])
    disc = traj['cont'][1:] * spec.discount
    ret = lambda_return(reward, value, disc, spec.lam)
    return ret, value[:-1], reward

  def update_slow(self):
    """A/agent.py:444-454 (slow_target_fraction 1.0 -> hard copy)."""
    u = int(self.ctrl['slow_updates'])
    if u == -1 or u >= self.spec.slow_every:
      u = 0
      for k in self.names['critic']:
        self.params[k.replace('critic/', 'slow_critic/', 1)] = self.params[k].clone()
    self.ctrl['slow_updates'] = torch.tensor(u + 1, dtype=torch.int64)

  # -- the step ---------------------------------------------------------------
  def train_step(self, batch, noise, state=None, keep=False):
    """One Agent.train(batch, state).  noise: {'obs_post' [L,B,S,C],
    'img' [H,N,S,C], 'act' [H+1,N,A]} as numpy/torch fp32.  Returns
    (outs, state, metrics[, aux if keep])."""
    spec = self.spec
    noise = {k: torch.as_tensor(np.asarray(v)).float() for k, v in noise.items()}
    data = self.preprocess(batch)
    mets = {}
    aux_all = {}

    # 1. world model (A/agent.py:157-163)
    p = self._leaf(self.names['model'])
    loss, aux, m = self.wm_loss(p, data, state, noise['obs_post'])
    mets.update(m)
    grads = dict(zip(self.names['model'], torch.autograd.grad(
        loss, [p[k] for k in self.names['model']], allow_unused=True)))
    grads = {k: (g if g is not None else torch.zeros_like(self.params[k]))
             for k, g in grads.items()}
    if not torch.isfinite(loss):
      raise FloatingPointError('model_loss is not finite')
    mets['model_loss'] = loss.detach()
    if keep:
      aux_all['model_grads'] = {k: g.clone() for k, g in grads.items()}
    mets['model_grad_norm'] = adam_update(
        self.params, grads, self.opts['model'], self.names['model'], spec)
    mets['model_grad_steps'] = self.opts['model']['step']
    post = {k: v.detach() for k, v in aux['post'].items()}
    new_state = {k: post[k][:, -1] for k in ('deter', 'stoch', 'logit')}

    # 2. imagination with the UPDATED world model (A/agent.py:82-84,317-324)
    flat = lambda x: x.reshape((-1,) + tuple(x.shape[2:]))
    start = {k: flat(post[k]) for k in ('deter', 'stoch', 'logit')}
    cont0 = flat(data['cont'])
    p = self._leaf(self.names['actor'])
    traj = self.imagine(p, start, cont0, noise['img'], noise['act'])

    # 3. critic (A/agent.py:398-417)
    with torch.no_grad():
      target, _, reward = self._target(self.params, traj, 'slow_critic')
    pc = self._leaf(self.names['critic'])
    weight = traj['weight'].detach()
    feat_sg = traj['feat'].detach()
    vpred = head(feat_sg[:-1], pc, 'critic', spec)[..., 0]
    closs = ((vpred - symlog(target)) ** 2 * weight[:-1]).mean()
    cgrads = dict(zip(self.names['critic'], torch.autograd.grad(
        closs, [pc[k] for k in self.names['critic']])))
    if keep:
      aux_all['critic_grads'] = {k: g.clone() for k, g in cgrads.items()}
    mets['extr_critic_loss'] = closs.detach()
    mets['extr_critic_grad_norm'] = adam_update(
        self.params, cgrads, self.opts['critic'], self.names['critic'], spec)
    mets['extr_critic_grad_steps'] = self.opts['critic']['step']
    mets.update(
        extr_imag_reward_mean=reward.mean(), extr_imag_reward_std=_std(reward),
        extr_imag_critic_mean=symexp(vpred.detach()).mean(),
        extr_imag_critic_std=_std(symexp(vpred.detach())),
        extr_imag_return_mean=target.mean(), extr_imag_return_std=_std(target))
    self.update_slow()

    # 4. actor (A/agent.py:326-381)
    p = {**self.params, **{k: p[k] for k in self.names['actor']}}
    ret, base, _ = self._target(p, traj, 'slow_critic')
    normalize_update(self.ctrl, 'retnorm', ret, 0.999)       # impl 'off': stats only
    score = ret - base
    normalize_update(self.ctrl, 'scorenorm', score, 0.99)    # impl 'off'
    mets.update(extr_score_mean=score.mean(), extr_score_std=_std(score),
                extr_score_mag=score.abs().mean(), extr_score_max=score.abs().max())
    normalize_update(self.ctrl, 'advnorm', score, spec.advnorm_decay)
    mean, scale = normalize_stats(self.ctrl, 'advnorm', spec.advnorm_decay, spec.advnorm_max)
    score = (score - mean) * scale
    dist = actor_dist(feat_sg, p, spec)
    if spec.discrete:
      logp = (torch.log_softmax(dist['logit'], -1) * traj['action'].detach()).sum(-1)
      aloss = -logp[:-1] * score.detach()
      ent = categorical_entropy(dist['logit'])[:-1] / math.log(spec.act_dim)
      avg = ent.detach().mean()
    else:
      aloss = -score
      ent = 0.5 * math.log(2 * math.pi * math.e) + torch.log(dist['std'][:-1])
      lo = 0.5 * math.log(2 * math.pi * math.e * spec.actor_minstd ** 2)
      hi = 0.5 * math.log(2 * math.pi * math.e * spec.actor_maxstd ** 2)
      ent = (ent - lo) / (hi - lo)
      avg = ent.detach().mean((0, 1))
    self.ctrl['actent_scale'] = autoadapt_update(
        self.ctrl['actent_scale'], avg, spec.actent_target, spec.actent_min,
        spec.actent_max, True, spec)
    s = self.ctrl['actent_scale']
    ent_loss = -(s * ent)
    if not spec.discrete:
      ent_loss = ent_loss.sum(-1)
    mets.update(actent_mean=ent.mean(), actent_std=_std(ent),
                actent_scale_mean=s.mean(),
                actent_scale_std=_std(s) if s.dim() else torch.zeros(()))
    aloss = ((aloss + ent_loss) * weight[:-1]).mean()
    agrads = dict(zip(self.names['actor'], torch.autograd.grad(
        aloss, [p[k] for k in self.names['actor']])))
    if keep:
      aux_all['actor_grads'] = {k: g.clone() for k, g in agrads.items()}
    mets['actor_loss'] = aloss.detach()
    mets['actor_grad_norm'] = adam_update(
        self.params, agrads, self.opts['actor'], self.names['actor'], spec)
    mets['actor_grad_steps'] = self.opts['actor']['step']

    mets = {k: (float(v.detach()) if torch.is_tensor(v) else float(v))
            for k, v in mets.items()}
    if keep:
      aux_all.update(
          post=post, prior={k: v.detach() for k, v in aux['prior'].items()},
          embed=aux['embed'].detach(), kl=aux['kl'].detach(),
          traj={k: v.detach() for k, v in traj.items()},
          target=target, ret=ret.detach(), score=score.detach())
      return {}, new_state, mets, aux_all
    return {}, new_state, mets
