"""GPU diagnostic: which operand element does tcgen05 read for each (row, k)?
Uses an identity on the (known-good) K-major side so C reproduces the other
operand; entries encode their own coordinates."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from daydreamer_b200.backend import CudaBackend
cu = CudaBackend()
cu.set_gemm_tc(1)
np.set_printoptions(linewidth=250)

def probe_b_mn(N):
  M, K = 128, 32
  A = torch.zeros(M, K); A[:K, :K] = torch.eye(K)               # K-major A (works)
  kk, nn = torch.meshgrid(torch.arange(K), torch.arange(N), indexing='ij')
  B = (kk * 256 + nn).float()                                     # [K, N] -> MN-major B
  C = torch.zeros(M, N, device='cuda')
  cu.gemm(A.cuda(), B.cuda(), C, ta=False, tb=False)
  torch.cuda.synchronize()
  c = C.cpu().numpy()[:K].astype(np.int64)
  print(f'--- B MN-major, N={N}: C[k, n] should be k*256+n; got (k\',n\') ---')
  ok = (c == B.numpy().astype(np.int64))
  print('exact fraction', ok.mean(), ' nonzero rows beyond K:', np.abs(C.cpu().numpy()[K:]).max())
  for k in (0, 1, 2, 7, 8, 9):
    print('k', k, [(int(v) // 256, int(v) % 256) for v in c[k, :12]], '...', [(int(v) // 256, int(v) % 256) for v in c[k, 60:64]])

def probe_a_mn():
  M, K, N = 128, 32, 128
  kk, mm = torch.meshgrid(torch.arange(K), torch.arange(M), indexing='ij')
  A = (kk * 256 + mm).float()                                     # [K, M] -> MN-major A
  B = torch.zeros(N, K); B[:K, :K] = torch.eye(K)                 # [N, K] K-major B (works)
  C = torch.zeros(M, N, device='cuda')
  cu.gemm(A.cuda(), B.cuda(), C, ta=True, tb=True)
  torch.cuda.synchronize()
  c = C.cpu().numpy()[:, :K].astype(np.int64)                     # C[m, k] should be k*256+m
  want = (kk * 256 + mm).numpy().T.astype(np.int64)
  print('--- A MN-major: C[m, k] should be k*256+m ---')
  print('exact fraction', (c == want).mean())
  for m in (0, 1, 2, 3, 4, 5, 8, 33):
    print('m', m, [(int(v) // 256, int(v) % 256) for v in c[m, :10]])

probe_b_mn(128)
probe_a_mn()
