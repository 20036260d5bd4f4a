"""Agent.policy (A/agent.py:42-65): the acting path of the engine (PolicyEngine: encoder -> one
obs_step -> actor) against the oracle's restatement with injected noise, over several steps with
episode resets.  CPU: on the primitive simulator; -m gpu: on the CUDA kernels."""
import numpy as np
import pytest
import torch

from oracle import dreamer_oracle as O
from daydreamer_b200 import engine as E
from tests import util
from tests.hostsim import HostSim


def run_policy(be, kw, P=3, steps=4, tol=2e-5):
  ospec = O.tiny_spec(**kw)
  spec = util.to_engine_spec(ospec)
  params = O.init_params(ospec, seed=3)
  rng = np.random.default_rng(7)
  for k, v in params.items():
    if not k.endswith('/kernel'):
      params[k] = (v + 0.1 * rng.standard_normal(v.shape)).astype(np.float32)
  ol = O.Learner(ospec, params)
  eng = E.Engine(spec, be, 2, 3)
  eng.load_params(params)
  pol = E.PolicyEngine(eng, P)
  dev = pol.obs.device
  put = lambda dst, a: dst.copy_(torch.as_tensor(np.ascontiguousarray(a, np.float32)).to(dev).reshape(dst.shape))
  ostate = None
  put(pol.prev_action, np.zeros((P, spec.act_dim)))
  for t in range(steps):
    obs = {k: rng.standard_normal((P, d)).astype(np.float32) for k, d in ospec.obs_dims.items()}
    first = np.zeros(P, bool)
    if t == 0:
      first[:] = True
    if t == 2:
      first[1] = True                       # an episode boundary in the middle
    obs.update(is_first=first, is_terminal=np.zeros(P, bool), reward=np.zeros(P, np.float32))
    g = rng.gumbel(size=(P, ospec.stoch, ospec.classes)).astype(np.float32)
    eps = (rng.gumbel(size=(P, ospec.act_dim)) if ospec.discrete else
           rng.standard_normal((P, ospec.act_dim))).astype(np.float32)
    oout, ostate = ol.policy(obs, ostate, g, eps, 'train')
    put(pol.obs, np.concatenate([obs[k] for k in ospec.obs_dims], -1))
    put(pol.first, first.astype(np.float32))
    put(pol.g, g.reshape(P, -1))
    put(pol.eps, eps)
    pol.launch(draw=False)
    be.sync()
    act = pol.action.cpu().numpy()
    np.testing.assert_array_equal(pol.feat[:, spec.deter:].cpu().numpy(), ostate['stoch'])   # indices exact
    np.testing.assert_allclose(pol.feat[:, :spec.deter].cpu().numpy(), ostate['deter'], rtol=tol, atol=tol)
    if ospec.discrete:
      np.testing.assert_array_equal(act, oout['action'])
    else:
      np.testing.assert_allclose(act, oout['action'], rtol=tol, atol=tol)
    put(pol.prev_action, act)               # Agent.policy carries the action it returned
  return True


CASES = [dict(), dict(discrete=True, act_dim=5, actent_target=0.4),
         dict(obs_dims={'vector': 16}, act_dim=12, deter=256, units=256, stoch=8, classes=32,
              enc_units=256, head_units=256)]


@pytest.mark.parametrize('kw', CASES)
def test_policy_matches_oracle_hostsim(kw):
  assert run_policy(HostSim(), kw)


@pytest.mark.gpu
@pytest.mark.parametrize('kw', CASES)
def test_policy_matches_oracle_gpu(kw):
  from daydreamer_b200.backend import CudaBackend
  assert run_policy(CudaBackend(), kw, tol=1e-4)
