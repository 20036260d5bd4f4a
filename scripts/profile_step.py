"""Runs eager (no CUDA graph) train steps of the a1 bench workload; meant to be
wrapped by ncu: step 0 warms up, later steps are what gets profiled (DD_PROFILE_LAST=1 brackets the
last step with cudaProfilerStart / Stop for `ncu --profile-from-start off`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from daydreamer_b200.agent import Agent
from daydreamer_b200.config import make_config

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = make_config('a1', **{'rssm.deter': 512, 'rssm.units': 512, 'batch_size': 16,
                          'replay_chunk': 64, 'imag_horizon': 15})
obs, act = bench.spaces()
agent = Agent(obs, act, None, cfg)
agent.use_graph = False
state = None
bracket = os.environ.get('DD_PROFILE_LAST', '') != ''     # ncu --profile-from-start off: only the last step
for i in range(steps):
  n0 = agent.be.launch_count()
  if bracket and i == steps - 1:
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
  _, state, mets = agent.train(bench.synthetic_batch(16, i), state)
  if bracket and i == steps - 1:
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
  print('step', i, 'launches', agent.be.launch_count() - n0, 'model_loss', mets['model_loss'], flush=True)
