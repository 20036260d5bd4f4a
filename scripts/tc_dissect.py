import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daydreamer_b200.backend import CudaBackend
cu = CudaBackend(); cu.set_gemm_tc(1)
M, N, K = 16384, 512, 1024
for ta, tb in ((0, 0), (0, 1), (1, 0), (1, 1)):
  A = torch.randn((K, M) if ta else (M, K), device='cuda'); B = torch.randn((N, K) if tb else (K, N), device='cuda')
  C = torch.zeros(M, N, device='cuda')
  out = []
  for dbg in (0, 1, 2, 3):
    cu.lib.dd_set_gemm_tc_debug(dbg)
    for _ in range(2): cu.gemm(A, B, C, ta=bool(ta), tb=bool(tb))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): cu.gemm(A, B, C, ta=bool(ta), tb=bool(tb))
    e1.record(); torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) * 200)
  cu.lib.dd_set_gemm_tc_debug(0)
  print(f'ta={ta} tb={tb} (A {"MN" if ta else "K"}-major, B {"K" if tb else "MN"}-major): full {out[0]:.0f}us | mma-only {out[1]:.0f}us | load-only {out[2]:.0f}us | neither {out[3]:.0f}us', flush=True)
