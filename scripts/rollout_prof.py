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
  TN = 2048
  buf = (ctypes.c_ulonglong * (148 * 16 + 4 * TN))()
  be.lib.dd_machine_prof(0, buf, 148 * 16 + 4 * TN)
  arr = np.array(buf[:148 * 16], dtype=np.float64).reshape(148, 16)
  tr = np.array(buf[148 * 16:], dtype=np.float64).reshape(4, TN)
  # CTA 0's timeline of the first layers: epilogue events per (tile, layer) pair, relative cycles
  EPI = ['PLAIN', 'LN', 'LNBWD', 'GRU', 'STATS', 'ACTOR', 'STATSBWD', 'GRUBWD']
  layers = getattr(eng, '_rollout_layers', [])
  t0 = tr[3, 0]
  rel = lambda x: int(x - t0) if x > 0 else -1
  print('  pair epi    N  kblk | k-loop  epilogue  publish  to-next-start   (cycles; k-loop = start -> accumulator full)')
  for pr in range(min(len(layers), 40)):
    e = [rel(tr[2, 16 * pr + k]) for k in range(12)]
    nxt = rel(tr[2, 16 * (pr + 1)]) if pr + 1 < len(layers) else -1
    L = layers[pr]
    kb = sum((sg.a.shape[1] + 31) // 32 for sg in L.segs) * max(1, L.N // 512)
    x = [rel(tr[2, 16 * pr + k]) for k in range(12, 16)]
    extra = '' if x[0] < 0 else '   marks (from start): ' + ' '.join(str(v - e[0]) for v in x) + f'  accf {e[1] - e[0]} done {e[9] - e[0]}'
    print(f'  {pr:4d} {EPI[L.epi]:6s} {L.N:4d} {kb:5d} | {e[1] - e[0]:6d} {e[9] - e[1]:9d} {e[11] - e[9]:8d} {nxt - e[11]:8d}{extra}')
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
