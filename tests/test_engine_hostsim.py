"""Host logic check (no GPU): the product's orchestration (engine.py, hand-written
backward) running on the torch-CPU primitive simulator must reproduce the
oracle's autograd train step."""
import os

import numpy as np
import pytest
import torch

os.environ.setdefault('DD_MACHINE_BWD', '1')      # the CPU tests cover the reverse programs too
os.environ.setdefault('DD_REVERSE', '1')

from oracle import dreamer_oracle as O
from daydreamer_b200 import engine as E
from tests import util
from tests.hostsim import HostSim


def rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)


@pytest.mark.parametrize('B,L,steps,kw', [
    (3, 5, 1, {}), (2, 6, 3, {}),
    (3, 5, 2, dict(discrete=True, act_dim=5, actent_target=0.4)),   # xarm/ur5-style one-hot actor
    # image + depth CNN encoder/decoder next to the proprio MLP (xarm-style), one-hot actor
    (2, 4, 2, dict(discrete=True, act_dim=5, cnn_keys={'image': (18, 18, 3), 'depth': (18, 18, 1)},
                   cnn_depth=4, enc_kernels=(4, 4), dec_kernels=(3, 4, 4))),
    # decoder MLP sized differently from the other heads (the reference's `debug` block)
    (2, 4, 1, dict(dec_layers=3, dec_units=12)),
    # the encoder sees more vector keys than the decoder reconstructs (reference default key regexes)
    (2, 4, 2, dict(obs_dims={'vector': 5, 'aux': 1, 'count': 1}, dec_dims={'vector': 5})),
    # image only, continuous actions
    (2, 3, 1, dict(obs_dims={}, cnn_keys={'image': (10, 10, 3)}, cnn_depth=4, enc_kernels=(4, 4),
                   dec_kernels=(4, 4))),
    # widths that tile across the layer machine's clusters: every trunk, the prior stack and the
    # rollout cell run as fused programs (operands reconstructed from their (x, x_lo) pairs)
    (2, 3, 2, dict(obs_dims={'vector': 16}, act_dim=12, deter=256, units=256, stoch=8, classes=32,
                   enc_units=256, head_units=256, horizon=2)),
    (2, 3, 1, dict(obs_dims={'vector': 8}, discrete=True, act_dim=8, deter=256, units=256, stoch=8,
                   classes=32, enc_units=256, head_units=256, horizon=2)),
])
def test_train_step_matches_oracle(B, L, steps, kw):
  torch.manual_seed(0)
  ospec = O.tiny_spec(**kw)
  spec = util.to_engine_spec(ospec)
  params = O.init_params(ospec, seed=1)
  # make LN params / biases non-trivial so every gradient path is exercised
  rng = np.random.default_rng(5)
  for k, v in params.items():
    if not k.endswith('/kernel'):
      params[k] = (v + 0.1 * rng.standard_normal(v.shape)).astype(np.float32)
  ol = O.Learner(ospec, params)
  eng = E.Engine(spec, HostSim(), B, L)
  eng.load_params(params)
  ostate = None
  estate = None
  for it in range(steps):
    batch = util.make_batch(ospec, B, L, seed=10 + it)
    noise = util.make_noise(ospec, B, L, seed=20 + it)
    _, ostate, omets, aux = ol.train_step(batch, noise, ostate, keep=True)
    util.load_batch(eng, spec, batch, noise, estate)
    # steps after the first do not refresh the slow critic -> exercise the skip path
    eng.launch_step(draw=False, slow_refreshed=eng.slow_refresh_due(int(eng.counters[E.C_SLOW])))
    estate = {'deter': eng.prev_deter.numpy().copy(), 'stoch': eng.prev_stoch.numpy().copy()}
    # posterior indices bit-exact, features close
    post_stoch = util.tm(aux['post']['stoch'].numpy().reshape(B, L, -1))
    assert np.array_equal(eng.feat[:, spec.deter:].numpy(), post_stoch)
    assert rel(eng.feat[:, :spec.deter].numpy(), util.tm(aux['post']['deter'].numpy())) < 1e-5
    # gradients of all three groups
    for group in ('model', 'critic', 'actor'):
      eg = eng.get_grads(group)
      og = aux[f'{group}_grads']
      for k in og:
        assert rel(eg[k], og[k].numpy()) < 2e-4, (it, group, k, rel(eg[k], og[k].numpy()))
    # trajectory
    traj = aux['traj']
    tstoch = util.ref_rows_to_tm(traj['stoch'].numpy().reshape(spec.horizon + 1, B * L, -1), B, L)
    assert np.array_equal(eng.tfeat[:, :, spec.deter:].numpy(), tstoch)
    assert rel(eng.r_ret.numpy(), util.ref_rows_to_tm(aux['ret'].numpy()[..., None], B, L)[..., 0]) < 1e-4
    # updated parameters
    ep = eng.get_params()
    for k, v in ol.params.items():
      # Adam moves every element by ~lr whatever its gradient: an element whose gradient is at
      # rounding noise may land a small fraction of lr away
      assert (rel(ep[k], v.numpy()) < 1e-5 or
              np.abs(ep[k] - v.numpy()).max() <= 0.05 * ospec.lr), (it, k)
    # metrics
    em = eng.read_metrics()
    for k, v in omets.items():
      assert k in em, k
      if np.isnan(v):
        assert np.isnan(em[k]), k
      else:
        assert abs(em[k] - v) <= 2e-4 * max(1.0, abs(v)), (it, k, em[k], v)
    assert set(em) == set(omets)


@pytest.mark.parametrize('case', ['all_first', 'L1', 'B1', 'all_terminal', 'no_terminal_no_first'])
def test_edge_case_batches_match_oracle(case):
  """Edge cases of the replay batch: every step an episode start (the carry is reset at each
  step), chunks of a single step, a single sequence, every step terminal (cont = 0 everywhere:
  imagination weights vanish after the start), and a continuation chunk without any boundary."""
  B, L = {'L1': (3, 1), 'B1': (1, 5)}.get(case, (2, 4))
  ospec = O.tiny_spec()
  spec = util.to_engine_spec(ospec)
  params = O.init_params(ospec, seed=2)
  ol = O.Learner(ospec, params)
  eng = E.Engine(spec, HostSim(), B, L)
  eng.load_params(params)
  ostate = estate = None
  for it in range(2):
    batch = util.make_batch(ospec, B, L, seed=30 + it, mid_first=False)
    if case == 'all_first':
      batch['is_first'][:] = True
    if case == 'all_terminal':
      batch['is_terminal'][:] = True
      batch['is_last'][:] = True
    if case == 'no_terminal_no_first':
      batch['is_terminal'][:] = False
      if it > 0:
        batch['is_first'][:] = False
    noise = util.make_noise(ospec, B, L, seed=40 + it)
    _, ostate, omets, aux = ol.train_step(batch, noise, ostate, keep=True)
    util.load_batch(eng, spec, batch, noise, estate)
    eng.launch_step(draw=False, slow_refreshed=eng.slow_refresh_due(int(eng.counters[E.C_SLOW])))
    estate = {'deter': eng.prev_deter.numpy().copy(), 'stoch': eng.prev_stoch.numpy().copy()}
    post_stoch = util.tm(aux['post']['stoch'].numpy().reshape(B, L, -1))
    assert np.array_equal(eng.feat[:, spec.deter:].numpy(), post_stoch)
    for group in ('model', 'critic', 'actor'):
      eg = eng.get_grads(group)
      for k, v in aux[f'{group}_grads'].items():
        scale = max(np.abs(v.numpy()).max(), 1e-6)
        assert np.abs(eg[k] - v.numpy()).max() <= 2e-4 * scale + 1e-9, (case, it, group, k)
    em = eng.read_metrics()
    for k, v in omets.items():
      if np.isnan(v):
        assert np.isnan(em[k]), (case, k)
      else:
        assert abs(em[k] - v) <= 2e-4 * max(1.0, abs(v)), (case, it, k, em[k], v)


@pytest.mark.parametrize('kw', [
    dict(prior_layers=1, enc_layers=1, head_layers=1, horizon=1),
    dict(stoch=1, classes=2, act_dim=1, deter=8, units=12),
    dict(obs_dims={'a': 1, 'b': 3, 'c': 2}, stoch=5, classes=3, prior_layers=2, horizon=2),
    dict(discrete=True, act_dim=2, stoch=2, classes=7, head_units=8, enc_units=8, horizon=4),
    dict(kl_balance=0.5, discount=0.9, lam=0.5, rssm_unimix=0.1, actor_minstd=0.3, clip=0.5, wd=0.0),
])
def test_odd_sizes_and_hyperparameters_match_oracle(kw):
  """Degenerate layer counts, single-group / two-class latents, one-dimensional actions, several
  vector keys, a gradient clip that is active, no weight decay."""
  B, L = 2, 3
  ospec = O.tiny_spec(**kw)
  spec = util.to_engine_spec(ospec)
  params = O.init_params(ospec, seed=4)
  rng = np.random.default_rng(6)
  for k, v in params.items():
    if not k.endswith('/kernel'):
      params[k] = (v + 0.1 * rng.standard_normal(v.shape)).astype(np.float32)
  ol = O.Learner(ospec, params)
  eng = E.Engine(spec, HostSim(), B, L)
  eng.load_params(params)
  ostate = estate = None
  for it in range(2):
    batch = util.make_batch(ospec, B, L, seed=50 + it)
    noise = util.make_noise(ospec, B, L, seed=60 + it)
    _, ostate, omets, aux = ol.train_step(batch, noise, ostate, keep=True)
    util.load_batch(eng, spec, batch, noise, estate)
    eng.launch_step(draw=False, slow_refreshed=eng.slow_refresh_due(int(eng.counters[E.C_SLOW])))
    estate = {'deter': eng.prev_deter.numpy().copy(), 'stoch': eng.prev_stoch.numpy().copy()}
    for group in ('model', 'critic', 'actor'):
      eg = eng.get_grads(group)
      for k, v in aux[f'{group}_grads'].items():
        assert rel(eg[k], v.numpy()) < 3e-4 or np.abs(eg[k] - v.numpy()).max() < 1e-8, (it, group, k)
    ep = eng.get_params()
    for k, v in ol.params.items():
      assert rel(ep[k], v.numpy()) < 1e-5, (it, k)
    em = eng.read_metrics()
    for k, v in omets.items():
      assert (np.isnan(v) and np.isnan(em[k])) or abs(em[k] - v) <= 2e-4 * max(1.0, abs(v)), (it, k, em[k], v)
