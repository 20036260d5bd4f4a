"""CPU: histogram of dd_gemm calls (M, N, K, ta, tb, beta) in one a1 train step."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daydreamer_b200 import engine as E

class Recorder:
  device = torch.device('cpu')
  def __init__(self): self.calls = collections.Counter(); self.other = collections.Counter()
  def zeros(self, *shape, dtype=torch.float32): return torch.empty(*shape, dtype=dtype)
  def gemm(self, A, B, C, ta=False, tb=False, beta=0.0, bias=None):
    K = A.shape[0] if ta else A.shape[1]
    self.calls[(C.shape[0], C.shape[1], K, int(ta), int(tb), int(beta), bias is not None)] += 1
  def __getattr__(self, name):
    def f(*a, **k): self.other[name] += 1
    return f

spec = E.Spec(obs_dims={'vector': 16})
rec = Recorder()
eng = E.Engine(spec, rec, 16, 64)
eng.fused_scan = eng.fused_bptt = False
rec.calls.clear(); rec.other.clear()
eng.launch_step()
tot = sum(2.0 * m * n * k * c for (m, n, k, *_), c in rec.calls.items())
print(f'{sum(rec.calls.values())} gemm calls, {tot/1e9:.1f} GFLOP; other launches: {sum(rec.other.values())}')
rows = sorted(rec.calls.items(), key=lambda kv: -2.0 * kv[0][0] * kv[0][1] * kv[0][2] * kv[1])
for (m, n, k, ta, tb, beta, bias), c in rows:
  print(f'M={m:6d} N={n:5d} K={k:6d} ta={ta} tb={tb} beta={beta} bias={int(bias)} x{c:4d}  {2.0*m*n*k*c/1e9:8.2f} GF  {100*2.0*m*n*k*c/tot:5.1f}%')
print(sorted(rec.other.items(), key=lambda kv: -kv[1]))
