"""GPU: accuracy + speed of the tcgen05 3xTF32 GEMM vs the SIMT fp32 GEMM."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daydreamer_b200.backend import CudaBackend
cu = CudaBackend()
torch.manual_seed(0)
shapes = [(1024, 512, 512, 0, 0), (1024, 1536, 1024, 0, 0), (16384, 512, 1536, 0, 0), (16384, 512, 512, 0, 0),
          (16384, 1536, 512, 0, 1), (512, 512, 16384, 1, 0), (1536, 512, 16384, 1, 0), (1024, 1024, 512, 0, 1)]
for M, N, K, ta, tb in shapes:
  A = torch.randn((K, M) if ta else (M, K), device='cuda')
  B = torch.randn((N, K) if tb else (K, N), device='cuda')
  ref = ((A.t() if ta else A).double() @ (B.t() if tb else B).double())
  row = []
  for on in (1, 0):
    cu.set_gemm_tc(on)
    C = torch.zeros(M, N, device='cuda')
    beta = 1.0 if ta else 0.0
    cu.gemm(A, B, C, ta=bool(ta), tb=bool(tb), beta=beta)
    torch.cuda.synchronize()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
      cu.gemm(A, B, C, ta=bool(ta), tb=bool(tb), beta=beta)
    e0.record()
    for _ in range(10):
      cu.gemm(A, B, C, ta=bool(ta), tb=bool(tb), beta=beta)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    row.append((err, us, 2.0 * M * N * K / us / 1e6))
  print(f'M={M} N={N} K={K} ta={ta} tb={tb}: TC err {row[0][0]:.2e} {row[0][1]:.1f}us {row[0][2]:.1f} TF/s | '
        f'SIMT err {row[1][0]:.2e} {row[1][1]:.1f}us {row[1][2]:.1f} TF/s', flush=True)
cu.set_gemm_tc(1)
