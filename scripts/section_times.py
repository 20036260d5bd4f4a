"""GPU: CUDA-event time of each phase of the train step, each phase captured as
its own CUDA graph (so launch overhead does not pollute the numbers)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from daydreamer_b200.agent import Agent
from daydreamer_b200.config import make_config

cfg = make_config('a1', **{'rssm.deter': 512, 'rssm.units': 512, 'batch_size': 16,
                          'replay_chunk': 64, 'imag_horizon': 15})
obs, act = bench.spaces()
agent = Agent(obs, act, None, cfg)
agent.use_graph = False
state = None
for i in range(2):
  _, state, m = agent.train(bench.synthetic_batch(16, i), state)
eng = agent._main
s = eng.spec
def tail():
  B, L = eng.B, eng.L
  r = slice((L - 1) * B, L * B)
  eng.be.copy_(eng.prev_deter, eng.feat[r, :s.deter]); eng.be.copy_(eng.prev_stoch, eng.feat[r, s.deter:])
  eng.be.copy_(eng.cont0.view(-1, 1), eng.cont)
phases = [('noise+init', lambda: (eng.draw_noise(), eng.be.stats_init(eng.mstat))),
          ('wm_forward', eng.wm_forward), ('wm_losses', eng.wm_losses), ('wm_backward', eng.wm_backward),
          ('model_opt', lambda: eng.opt_step('model', 0)), ('carry', tail), ('imagine', eng.imagine),
          ('traj_values', eng.traj_values), ('critic_update', eng.critic_update), ('actor_update', eng.actor_update)]
graphs = []
for name, fn in phases:
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  n0 = eng.be.launch_count()
  with torch.cuda.graph(g):
    fn()
  graphs.append((name, g, eng.be.launch_count() - n0))
tot = collections.OrderedDict()
for rep in range(6):
  for name, g, nl in graphs:
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    if rep >= 2:
      tot[name] = tot.get(name, 0.0) + e0.elapsed_time(e1) / 4
for name, g, nl in graphs:
  print(f'{name:14s} {tot[name]:7.3f} ms  {nl:5d} launches  {1000*tot[name]/max(nl,1):6.1f} us/launch')
print('sum', round(sum(tot.values()), 2), 'ms')
