"""Runs eager a1 train steps and brackets ONLY the two rollout programs (forward, reverse) with
cudaProfilerStart / Stop; wrap with `ncu --profile-from-start off --set full ...`."""
import sys, os
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
_, state, mets = agent.train(bench.synthetic_batch(16, 0), state)      # builds every program
eng, be = agent._main, agent.be
orig = be.program_run
targets = [h for h in (getattr(eng, '_rollout', None), getattr(eng, '_reverse', None)) if h is not None]


def run(h):
  hit = any(h is t for t in targets)
  if hit:
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
  orig(h)
  if hit:
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


be.program_run = run
_, state, mets = agent.train(bench.synthetic_batch(16, 1), state)
print('profiled', len(targets), 'programs; model_loss', mets['model_loss'])
