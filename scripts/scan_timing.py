import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from daydreamer_b200.agent import Agent
from daydreamer_b200.config import make_config
cfg = make_config('a1', **{'rssm.deter': 512, 'rssm.units': 512, 'batch_size': 16, 'replay_chunk': 64})
obs, act = bench.spaces()
agent = Agent(obs, act, None, cfg); agent.use_graph = False
_, st, _ = agent.train(bench.synthetic_batch(16, 0), None)
eng = agent._main
eng.scan_timing = torch.zeros(16, dtype=torch.int64, device='cuda')
for i in range(2):
  _, st, _ = agent.train(bench.synthetic_batch(16, i + 1), st)
torch.cuda.synchronize()
t = eng.scan_timing.cpu().numpy()[:16].astype(float)
names = ['P1 mask+img_in', 'B1', 'P2 LN+gru gemm', 'B2', 'P3 gates', 'B3', 'P4 obs gemm', 'B4', 'P5 LN+logits+sample', 'B5', 'P3a loads landed', '-', 'P2a staged', 'P2b LN done', 'P2c gemm done', '-']
tot = t.sum()
for n, v in zip(names, t):
  print(f'{n:22s} {v/64:9.0f} cycles/step  {100*v/tot:5.1f}%')
print('total cycles/step', tot / 64, ' (~%.1f us at 1.9 GHz)' % (tot / 64 / 1900))
# ---- backward (BPTT) kernel phase split
eng.bptt_timing = torch.zeros(12, dtype=torch.int64, device='cuda')
for i in range(2):
  _, st, _ = agent.train(bench.synthetic_batch(16, i + 3), st)
torch.cuda.synchronize()
t = eng.bptt_timing.cpu().numpy().astype(float)
names = ['P1a stats_bwd', 'B', 'P1b dzs@Wst^T', 'B', 'P2 LN+gemm+gru', 'B', 'P3 joint LN', 'B', 'P4 gemm Wgru^T', 'B',
         'P5 LN+gemm Win^T', 'B']
tot = t.sum()
for n, v in zip(names, t):
  print(f'bwd {n:20s} {v/64:9.0f} cycles/step  {100*v/tot:5.1f}%')
print('bwd total cycles/step', tot / 64, ' (~%.1f us at 1.9 GHz)' % (tot / 64 / 1900))
