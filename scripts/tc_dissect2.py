import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daydreamer_b200.backend import CudaBackend
cu = CudaBackend(); cu.set_gemm_tc(1); cu.lib.dd_set_gemm_tc_persistent(0)
names = {2: 'load-only', 6: 'load-only no-fence', 10: 'loads consumed, no STS', 18: 'STS+fence, no LDG', 26: 'no LDG no STS (fence only)', 30: 'nothing'}
for (M, N, K, ta, tb) in ((15360, 512, 512, 0, 0), (15360, 512, 512, 0, 1), (1024, 512, 512, 0, 0)):
  A = torch.randn((K, M) if ta else (M, K), device='cuda'); B = torch.randn((N, K) if tb else (K, N), device='cuda')
  C = torch.zeros(M, N, device='cuda')
  out = []
  for dbg in names:
    cu.lib.dd_set_gemm_tc_debug(dbg)
    for _ in range(2): cu.gemm(A, B, C, ta=bool(ta), tb=bool(tb))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): cu.gemm(A, B, C, ta=bool(ta), tb=bool(tb))
    e1.record(); torch.cuda.synchronize()
    out.append(f'{names[dbg]} {e0.elapsed_time(e1) * 50:.0f}us')
  cu.lib.dd_set_gemm_tc_debug(0)
  print(f'M={M} N={N} K={K} tb={tb}: ' + ' | '.join(out), flush=True)
