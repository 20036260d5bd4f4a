"""CPU: known-answer tests that pin the oracle (the reference holds no golden
vectors for this path -- SURVEY.md section 4 -- so closed forms anchor it), plus
the Philox known-answer vectors published with Random123 (tests/golden/)."""
import json
import math
import pathlib

import numpy as np
import pytest
import torch

from oracle import dreamer_oracle as O
from oracle import philox

GOLDEN = pathlib.Path(__file__).parent / 'golden'


def test_philox_known_answers():
  for case in json.loads((GOLDEN / 'philox4x32_10_kat.json').read_text()):
    c = [np.array([int(x, 16)], np.uint32) for x in case['counter']]
    k = [int(x, 16) for x in case['key']]
    out = philox.philox4x32(*c, k[0], k[1])
    assert [f'{int(o[0]):08x}' for o in out] == case['output'], case


def test_philox_streams_are_disjoint_and_uniform():
  u = philox.uniform(1 << 16, seed=7, stream=1, step=0)
  v = philox.uniform(1 << 16, seed=7, stream=2, step=0)
  w = philox.uniform(1 << 16, seed=7, stream=1, step=1)
  assert 0.0 < u.min() and u.max() < 1.0
  assert abs(u.mean() - 0.5) < 5e-3 and abs(np.corrcoef(u, v)[0, 1]) < 2e-2
  assert not np.array_equal(u, w)
  z = philox.normal(1 << 16, 7, 3, 0)
  assert abs(z.mean()) < 2e-2 and abs(z.std() - 1.0) < 2e-2


def test_unimix_logits_are_normalised():
  z = torch.randn(5, 4 * 6)
  lg = O.unimix_logits(z, 6, 0.01)
  assert torch.allclose(torch.logsumexp(lg, -1), torch.zeros(5, 4), atol=1e-6)
  assert (torch.exp(lg) >= 0.01 / 6 - 1e-7).all()


def test_kl_is_zero_for_equal_logits_and_positive_otherwise():
  a = torch.log_softmax(torch.randn(3, 4, 6), -1)
  b = torch.log_softmax(torch.randn(3, 4, 6), -1)
  assert torch.allclose(O.categorical_kl(a, a), torch.zeros(3), atol=1e-6)
  assert (O.categorical_kl(a, b) > 0).all()
  # KL balance keeps the value and splits the gradient 0.8 / 0.2
  post = {'logit': a.clone().requires_grad_(True)}
  prior = {'logit': b.clone().requires_grad_(True)}
  O.kl_loss(post, prior, 0.8).sum().backward()
  p2 = {'logit': a.clone().requires_grad_(True)}
  q2 = {'logit': b.clone().requires_grad_(True)}
  O.categorical_kl(p2['logit'], q2['logit']).sum().backward()
  assert torch.allclose(post['logit'].grad, 0.2 * p2['logit'].grad, atol=1e-6)
  assert torch.allclose(prior['logit'].grad, 0.8 * q2['logit'].grad, atol=1e-6)


def test_gru_gate_known_answer():
  spec = O.tiny_spec()
  p = {k: torch.tensor(v) for k, v in O.init_params(spec).items()}
  p['rssm/gru_out/kernel'] = torch.zeros_like(p['rssm/gru_out/kernel'])
  deter = torch.randn(2, spec.deter)
  # zero pre-activation, LN bias 0: r = .5, cand = tanh(0) = 0, u = sigmoid(-1)
  out = O.gru(torch.randn(2, spec.units), deter, p)
  u = 1.0 / (1.0 + math.e)
  assert torch.allclose(out, (1.0 - u) * deter, atol=1e-6)


def test_lambda_return_closed_form():
  H, N, g, lam = 7, 3, 0.9, 0.95
  r = torch.full((H, N), 2.0)
  v = torch.full((H + 1, N), 20.0)      # fixed point of V = r + g V
  ret = O.lambda_return(r, v, torch.full((H, N), g), lam)
  assert torch.allclose(ret, torch.full((H, N), 20.0), atol=1e-5)
  # lam = 0 is the one-step target, lam = 1 the Monte-Carlo return + bootstrap
  v = torch.randn(H + 1, N)
  assert torch.allclose(O.lambda_return(r, v, torch.full((H, N), g), 0.0), r + g * v[1:])
  mc = O.lambda_return(r, v, torch.full((H, N), g), 1.0)
  want = sum(2.0 * g ** k for k in range(H)) + g ** H * v[-1]
  assert torch.allclose(mc[0], want, atol=1e-5)


def test_symlog_symexp_inverse():
  x = torch.tensor([-30.0, -1.0, 0.0, 0.5, 100.0])
  assert torch.allclose(O.symexp(O.symlog(x)), x, rtol=1e-5)


def test_adam_first_step_is_lr_sign():
  spec = O.tiny_spec(wd=0.0)
  params = {'enc/dense0/kernel': torch.ones(3, 2), 'b': torch.zeros(2)}
  grads = {'enc/dense0/kernel': torch.tensor([[1.0, -2.0], [3.0, 4.0], [-5.0, 6.0]]),
           'b': torch.tensor([0.5, -0.5])}
  names = list(params)
  opt = O.init_opt(params, names)
  norm = O.adam_update(params, grads, opt, names, spec)
  assert abs(float(norm) - math.sqrt(sum(float((g ** 2).sum()) for g in grads.values()))) < 1e-5
  want = 1.0 - spec.lr * torch.sign(grads['enc/dense0/kernel'])
  assert torch.allclose(params['enc/dense0/kernel'], want, atol=1e-7)


def test_weight_decay_only_on_kernels_and_clip():
  spec = O.tiny_spec(clip=1.0)
  params = {'x/kernel': torch.ones(2, 2), 'x/norm/scale': torch.ones(2)}
  grads = {'x/kernel': torch.zeros(2, 2), 'x/norm/scale': torch.zeros(2)}
  grads['x/kernel'][0, 0] = 1e-30     # norm > 0 without moving anything visibly
  opt = O.init_opt(params, list(params))
  O.adam_update(params, grads, opt, list(params), spec)
  assert torch.allclose(params['x/kernel'][1], torch.full((2,), 1.0 - spec.wd * spec.lr))
  assert torch.equal(params['x/norm/scale'], torch.ones(2))


def test_straight_through_sample_is_one_hot_with_softmax_gradient():
  logit = torch.log_softmax(torch.randn(4, 3, 5), -1).requires_grad_(True)
  g = torch.zeros(4, 3, 5)
  g[..., 2] = 100.0
  s, idx = O.sample_onehot(logit, g)
  assert (idx == 2).all() and torch.equal(s.detach(), torch.nn.functional.one_hot(idx, 5).float())
  w = torch.randn(4, 3, 5)
  (s * w).sum().backward()
  ref = torch.log_softmax(logit.detach(), -1).requires_grad_(True)
  (torch.softmax(ref, -1) * w).sum().backward()
  assert torch.allclose(logit.grad, ref.grad, atol=1e-6)


def test_initial_state_is_learned2():
  spec = O.tiny_spec()
  p = {k: torch.tensor(v) for k, v in O.init_params(spec).items()}
  p['rssm/initial_deter'] = torch.randn(spec.deter)
  st = O.initial_state(3, p, spec)
  assert torch.allclose(st['deter'][1], torch.tanh(p['rssm/initial_deter']))
  assert torch.equal(st['stoch'].sum(-1), torch.ones(3, spec.stoch))
  assert torch.equal(st['logit'], torch.zeros(3, spec.stoch, spec.classes))


def test_train_step_is_deterministic_and_learns():
  from tests import util
  spec = O.tiny_spec()
  batch, noise = util.make_batch(spec, 3, 5), util.make_noise(spec, 3, 5)
  a, b = O.Learner(spec), O.Learner(spec)
  for _ in range(2):
    _, _, ma = a.train_step(batch, noise)
    _, _, mb = b.train_step(batch, noise)
  assert ma == {k: mb[k] for k in ma} or all(
      (np.isnan(ma[k]) and np.isnan(mb[k])) or ma[k] == mb[k] for k in ma)
  first = ma['model_loss']
  for _ in range(30):
    _, _, ma = a.train_step(batch, noise)
  assert ma['model_loss'] < first
  assert ma['model_grad_steps'] == 32 and ma['actor_grad_steps'] == 32


def test_conv_layers_match_the_tf_definition():
  """conv_layer vs naive loops over the TensorFlow definitions (tf.nn.conv2d VALID stride 2
  on HWIO kernels; tf.nn.conv2d_transpose = scatter of x[hi,wi,ci]*K[kh,kw,co,ci] at
  (2hi+kh, 2wi+kw), output size 2*in + k - 2; A/nets.py:523-547)."""
  rng = np.random.default_rng(0)
  x = rng.standard_normal((2, 9, 9, 3)).astype(np.float32)
  K = rng.standard_normal((4, 4, 3, 5)).astype(np.float32)
  b = rng.standard_normal(5).astype(np.float32)
  p = {'c/kernel': torch.tensor(K), 'c/bias': torch.tensor(b)}
  got = O.conv_layer(torch.tensor(x), p, 'c', norm=False).numpy()
  ho = (9 - 4) // 2 + 1
  want = np.zeros((2, ho, ho, 5), np.float64)
  for i in range(ho):
    for j in range(ho):
      patch = x[:, 2 * i:2 * i + 4, 2 * j:2 * j + 4, :].astype(np.float64)
      want[:, i, j, :] = np.einsum('nhwc,hwco->no', patch, K.astype(np.float64)) + b
  assert got.shape == want.shape and np.abs(got - want).max() < 1e-4
  # transposed
  x = rng.standard_normal((2, 3, 3, 4)).astype(np.float32)
  K = rng.standard_normal((5, 5, 6, 4)).astype(np.float32)       # [k, k, out, in]
  b = rng.standard_normal(6).astype(np.float32)
  p = {'t/kernel': torch.tensor(K), 't/bias': torch.tensor(b)}
  got = O.conv_layer(torch.tensor(x), p, 't', transp=True, norm=False).numpy()
  ho = 3 * 2 + 5 - 2
  want = np.zeros((2, ho, ho, 6), np.float64) + b
  for hi in range(3):
    for wi in range(3):
      contrib = np.einsum('nc,hwoc->nhwo', x[:, hi, wi, :].astype(np.float64), K.astype(np.float64))
      want[:, 2 * hi:2 * hi + 5, 2 * wi:2 * wi + 5, :] += contrib
  assert got.shape == want.shape and np.abs(got - want).max() < 1e-4


def test_image_geometry_and_loss():
  """Reference geometry (A/configs.yaml:54-55): 64 -> 31 -> 14 -> 6 -> 2 and 1 -> 5 -> 13 ->
  30 -> 64; embed = cnn(2*2*8*depth) + mlp; image loss = sum over H,W,C of (sigmoid - x/255)^2."""
  spec = O.Spec(obs_dims={'v': 19}, cnn_keys={'image': (64, 64, 3)}, discrete=True, act_dim=6)
  assert [s[4] for s in spec.enc_conv_shapes] == [31, 14, 6, 2]
  assert [s[6] for s in spec.enc_conv_shapes] == [64, 128, 256, 512]
  assert [s[4] for s in spec.dec_conv_shapes] == [5, 13, 30, 64]
  assert [s[6] for s in spec.dec_conv_shapes] == [256, 128, 64, 3]
  assert spec.cnn_out == 2048 and spec.embed == 2560
  shapes = O.param_shapes(spec)
  assert shapes['rssm/obs_out/kernel'] == (512 + 2560, 512)
  assert shapes['dec/conv0/kernel'] == (5, 5, 256, 1536)
  small = O.tiny_spec(obs_dims={}, cnn_keys={'image': (10, 10, 3)}, cnn_depth=4,
                      enc_kernels=(4, 4), dec_kernels=(4, 4))
  p = {k: torch.tensor(v) for k, v in O.init_params(small, 3).items()}
  for k in p:
    if k.startswith('dec/conv'):
      p[k] = torch.zeros_like(p[k])          # decoder outputs sigmoid(0) = 0.5 everywhere
  learner = O.Learner(small, {k: v.numpy() for k, v in p.items()})
  B, L = 2, 3
  from tests import util
  batch = util.make_batch(small, B, L, seed=0)
  batch['image'][:] = 255
  data = learner.preprocess(batch)
  g = torch.zeros(L, B, small.stoch, small.classes)
  _, aux, mets = learner.wm_loss(learner.params, data, None, g, training=False)
  assert abs(float(mets['image_loss_mean']) - 10 * 10 * 3 * 0.25) < 1e-3


def test_oracle_matches_committed_golden_steps():
  """tests/golden/oracle_steps.npz (made by tests/golden/make_oracle_golden.py): two seeded
  train steps of three small configurations.  Guards the oracle against drift between rounds;
  categorical indices must be identical, sums agree to fp32 reduction noise."""
  import importlib.util
  import pathlib
  here = pathlib.Path(__file__).parent / 'golden'
  spec = importlib.util.spec_from_file_location('make_oracle_golden', here / 'make_oracle_golden.py')
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  want = np.load(here / 'oracle_steps.npz')
  got = mod.build()
  assert set(got) == set(want.files)
  for k in want.files:
    if k.endswith('_idx'):
      assert np.array_equal(got[k], want[k]), k
    else:
      assert np.allclose(got[k], want[k], rtol=2e-4, atol=1e-5), (k, np.abs(got[k] - want[k]).max())
