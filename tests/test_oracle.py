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
