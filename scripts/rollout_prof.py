"""GPU: wait profile of the fused imagination-rollout program inside a real a1 train step:
CUDA-event time, per-CTA kernel cycles (-> effective SM clock) and the average wait cycles per
role.  DD_MACHINE_KS=1|2 selects the geometry."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
eng, be = agent._main, agent.be
NAMES = ['kernel', 'b:empty', 'a:ready', 'm:acce', 'm:full', 'e:stats', 'e:accf', '-', 'e:rdy8', 'a:empty', 'e:red']


def timed(fn, reps=5):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


for name, fn in (('imagine', eng.imagine),):
  ms = timed(fn)
  be.lib.dd_machine_prof(1, None, 0)
  fn()
  torch.cuda.synchronize()
  buf = (ctypes.c_ulonglong * (148 * 16))()
  be.lib.dd_machine_prof(0, buf, 148 * 16)
  arr = np.array(buf[:], dtype=np.float64).reshape(148, 16)
  used = arr[arr[:, 0] > 0]
  avg = used.mean(0)
  print(f'{name}: {ms:.3f} ms, {len(used)} CTAs, kernel {avg[0]:.0f} cycles/CTA -> effective clock '
        f'{avg[0] / (ms * 1e-3) / 1e9:.2f} GHz (one program launch per call; other launches of the phase included in ms)')
  print('   ' + '  '.join(f'{n}={avg[i]:.0f}' for i, n in enumerate(NAMES) if n != '-'))
  for kh in (0, 1):
    sel = used[[((i % 16) // 8) == kh for i in range(len(used))]] if len(used) > 64 else used
    a2 = sel.mean(0)
    print(f'   kh={kh}: ' + '  '.join(f'{n}={a2[i]:.0f}' for i, n in enumerate(NAMES) if n != '-'))
os.system('nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader')
