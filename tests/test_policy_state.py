"""Agent.policy keeps no recurrent state of its own: two drivers with the same number of environments
(the reference's train_eval loop alternates driver_train and driver_eval, embodied/run/train_eval.py)
share one PolicyEngine but each carries the state object policy() returned to it."""
import numpy as np
import torch

import bench
from daydreamer_b200 import engine as E
from daydreamer_b200.agent import Agent
from daydreamer_b200.config import make_config
from tests.hostsim import HostSim


def _agent(monkeypatch):
  cfg = make_config('a1', **{'rssm.deter': 32, 'rssm.units': 32, 'rssm.stoch': 4, 'rssm.classes': 4,
                            '.*\\.units': 16, 'batch_size': 2, 'replay_chunk': 4, 'imag_horizon': 2})
  obs, act = bench.spaces()
  agent = Agent(obs, act, None, cfg, seed=3, _backend_for_tests=HostSim())
  # fixed sampling noise: the caller's trajectory is then a function of its observations and state only
  launch = E.PolicyEngine.launch

  def fixed(self, draw=True):
    self.g.fill_(0.25)
    self.eps.fill_(0.5)
    return launch(self, draw=False)
  monkeypatch.setattr(E.PolicyEngine, 'launch', fixed)
  return agent


def _obs(rng, first):
  return {'vector': rng.standard_normal((1, 16)).astype(np.float32), 'is_first': np.array([first]),
          'is_last': np.array([False]), 'is_terminal': np.array([False]), 'reward': np.zeros(1, np.float32)}


def test_interleaved_policy_callers_keep_their_own_state(monkeypatch):
  solo = _agent(monkeypatch)
  rng = np.random.default_rng(0)
  obs_a = [_obs(rng, t == 0) for t in range(4)]
  obs_b = [_obs(rng, t == 0) for t in range(4)]
  state, want = None, []
  for o in obs_a:
    out, state = solo.policy(o, state, 'eval')
    want.append((out['action'].copy(), state['deter'].cpu().numpy().copy()))

  both = _agent(monkeypatch)
  sa = sb = None
  for t in range(4):
    out_a, sa = both.policy(obs_a[t], sa, 'eval')
    held = {k: (v.clone() if torch.is_tensor(v) else v.copy()) for k, v in sa.items()}
    out_b, sb = both.policy(obs_b[t], sb, 'eval')                 # a second driver with the same P
    for k in ('deter', 'stoch'):
      assert sa[k].data_ptr() != sb[k].data_ptr()                 # not views of one buffer
      assert torch.equal(sa[k], held[k])                          # ... and B's step did not touch A's state
    np.testing.assert_allclose(out_a['action'], want[t][0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(sa['deter'].cpu().numpy(), want[t][1], rtol=0, atol=1e-6)
