"""Shared helpers for the parity tests (seeded synthetic replay + noise)."""
import dataclasses
import numpy as np
import torch

from oracle import dreamer_oracle as O
from daydreamer_b200 import engine as E


def to_engine_spec(ospec):
  return E.Spec(**dataclasses.asdict(ospec))


def make_batch(spec, B, L, seed=0, term=0.05, mid_first=True):
  """SURVEY.md 8(d) synthetic replay: N(0,1) vectors, U(-1,1) actions,
  0.1*N(0,1) rewards, is_first[:,0]=True, Bernoulli is_terminal."""
  rng = np.random.default_rng(seed)
  vec = dict(spec.obs_dims)
  vec.update(getattr(spec, 'dec_dims', None) or {})
  b = {k: rng.standard_normal((B, L, d)).astype(np.float32) for k, d in vec.items()}
  for k, shp in getattr(spec, 'cnn_keys', {}).items():
    b[k] = rng.integers(0, 256, (B, L) + tuple(shp)).astype(np.uint8)
  b['action'] = rng.uniform(-1, 1, (B, L, spec.act_dim)).astype(np.float32)
  if getattr(spec, 'discrete', False):            # one-hot actions (embodied OneHotAction wrapper)
    b['action'] = np.eye(spec.act_dim, dtype=np.float32)[rng.integers(0, spec.act_dim, (B, L))]
  b['reward'] = (0.1 * rng.standard_normal((B, L))).astype(np.float32)
  f = np.zeros((B, L), bool)
  f[:, 0] = True
  if mid_first and L > 3:
    f[0, L // 2] = True          # an episode boundary inside a chunk
  b['is_first'] = f
  b['is_terminal'] = rng.uniform(size=(B, L)) < term
  b['is_last'] = b['is_terminal'].copy()
  return b


def make_noise(spec, B, L, seed=2):
  rng = np.random.default_rng(seed)
  N = B * L
  gum = lambda *s: (-np.log(-np.log(rng.uniform(1e-6, 1 - 1e-6, s)))).astype(np.float32)
  return {'obs_post': gum(L, B, spec.stoch, spec.classes),
          'img': gum(spec.horizon, N, spec.stoch, spec.classes),
          'act': (gum(spec.horizon + 1, N, spec.act_dim) if spec.discrete else
                  rng.standard_normal((spec.horizon + 1, N, spec.act_dim)).astype(np.float32))}


def tm(x):
  """[B, L, ...] -> time-major flat [L*B, ...]."""
  x = np.asarray(x)
  return np.ascontiguousarray(np.swapaxes(x, 0, 1)).reshape((-1,) + x.shape[2:])


def ref_rows_to_tm(x, B, L):
  """[..., N_ref=(b*L+l), F] -> [..., N_tm=(l*B+b), F] on axis -2."""
  x = np.asarray(x)
  lead = x.shape[:-2]
  y = x.reshape(lead + (B, L, x.shape[-1]))
  y = np.swapaxes(y, -3, -2)
  return np.ascontiguousarray(y).reshape(lead + (B * L, x.shape[-1]))


def load_batch(eng, spec, batch, noise=None, state=None):
  """Upload a [B, L, ...] numpy batch (+ injected noise) into engine buffers."""
  B, L = eng.B, eng.L
  dev = eng.obs.device
  put = lambda dst, arr: dst.copy_(torch.as_tensor(np.ascontiguousarray(arr, np.float32)).to(dev).reshape(dst.shape))
  if spec.obs_dims:
    obs = np.concatenate([np.asarray(batch[k], np.float32).reshape(B, L, -1)
                          for k in spec.obs_dims], -1)
    put(eng.obs, tm(obs))
  if getattr(spec, 'dec_dims', None) is not None and spec.dec_total:
    tgt = np.concatenate([np.asarray(batch[k], np.float32).reshape(B, L, -1)
                          for k in spec.dec_dims], -1)
    put(eng.dec_tgt, tm(tgt))
  if spec.cnn_keys:
    img = np.concatenate([np.asarray(batch[k], np.float32) / np.float32(255.0)
                          for k in spec.cnn_keys], -1)
    put(eng.img, tm(img))
  put(eng.action, tm(batch['action']))
  put(eng.first, tm(batch['is_first'].astype(np.float32)))
  put(eng.reward, tm(batch['reward']))
  put(eng.cont, tm(1.0 - batch['is_terminal'].astype(np.float32)))
  if state is None:
    eng.first0.fill_(1.0)
  else:
    put(eng.first0, batch['is_first'][:, 0].astype(np.float32))
    put(eng.prev_deter, state['deter'])
    put(eng.prev_stoch, np.asarray(state['stoch']).reshape(B, -1))
  if noise is not None:
    SC = spec.stoch * spec.classes
    put(eng.g_obs, noise['obs_post'].reshape(L, B, SC))
    put(eng.g_img, ref_rows_to_tm(noise['img'].reshape(spec.horizon, B * L, SC), B, L))
    put(eng.eps_act, ref_rows_to_tm(noise['act'], B, L))
