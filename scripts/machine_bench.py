"""GPU: per-layer time of layer-machine programs (csrc/dd_machine.cu) next to the same layers as
dd_gemm + dd_ln_fwd launches.  Usage: python scripts/machine_bench.py > gpurun_out/machine_bench.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daydreamer_b200 import program as P
from daydreamer_b200.backend import CudaBackend

be = CudaBackend()
dev = be.device


def rnd(*s, scale=1.0):
  return torch.randn(*s, device=dev) * scale


def timeit(fn, iters=20, warm=3):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters * 1e3     # us


def chain(M, K0, U, nl, act=1):
  X = rnd(M, K0)
  Xlo = P.lo_of(X)
  Ws = [rnd(K0 if i == 0 else U, U, scale=(K0 if i == 0 else U) ** -0.5) for i in range(nl)]
  pairs = [torch.stack([w.t().contiguous(), P.lo_of(w.t().contiguous())]) for w in Ws]
  Wt = [p[0] for p in pairs]
  Wtlo = [p[1] for p in pairs]
  g, b = torch.ones(U, device=dev), torch.zeros(U, device=dev)
  Z = [torch.zeros(M, U, device=dev) for _ in range(nl)]
  H = [torch.zeros(M, U, device=dev) for _ in range(nl)]
  Hlo = [torch.zeros(M, U, device=dev) for _ in range(nl)]
  st = [torch.zeros(M, 2, device=dev) for _ in range(nl)]
  layers = []
  for i in range(nl):
    src = P.Seg(X, Xlo) if i == 0 else P.Seg(H[i - 1], Hlo[i - 1])
    layers.append(P.Layer([src], Wt[i], Wtlo[i], U, P.EPI_LN, act=act, gamma=g, beta=b, z=Z[i],
                          st=st[i], out=H[i], out_lo=Hlo[i]))
  h = be.program_build(layers, M)
  t_prog = timeit(lambda: be.program_run(h))
  H2 = [torch.zeros(M, U, device=dev) for _ in range(nl)]

  def legacy():
    for i in range(nl):
      be.gemm(X if i == 0 else H2[i - 1], Ws[i], Z[i])
      be.ln_fwd(Z[i], g, b, H2[i], st[i], act)
  hold = [t.clone() for t in H]
  t_leg = timeit(legacy)
  err = max((a - c).abs().max().item() for a, c in zip(hold, H2))
  flops = 2.0 * M * U * (K0 + (nl - 1) * U)
  print(f'chain M={M} K0={K0} U={U} layers={nl}: machine {t_prog:8.1f} us ({t_prog / nl:6.1f}/layer, '
        f'{flops / t_prog * 1e-6:6.1f} TFLOP/s)   gemm+ln launches {t_leg:8.1f} us ({t_leg / nl:6.1f}/layer)   '
        f'max|diff| {err:.2e}', flush=True)


def single(M, N, K, epi, act=0):
  X = rnd(M, K); Xlo = P.lo_of(X)
  W = rnd(K, N, scale=K ** -0.5); pw = torch.stack([W.t().contiguous(), P.lo_of(W.t().contiguous())]); Wt, Wtlo = pw[0], pw[1]
  g, b = torch.ones(N, device=dev), torch.zeros(N, device=dev)
  z, out, lo, st = (torch.zeros(M, N, device=dev), torch.zeros(M, N, device=dev),
                    torch.zeros(M, N, device=dev), torch.zeros(M, 2, device=dev))
  L = P.Layer([P.Seg(X, Xlo)], Wt, Wtlo, N, epi, act=act, gamma=g, beta=b, z=z, st=st, out=out, out_lo=lo)
  h = be.program_build([L], M)
  t = timeit(lambda: be.program_run(h))
  t2 = timeit(lambda: be.gemm(X, W, z))
  print(f'single M={M} N={N} K={K} epi={epi}: machine {t:7.1f} us ({2.0 * M * N * K / t * 1e-6:6.1f} TFLOP/s)   '
        f'dd_gemm alone {t2:7.1f} us', flush=True)


if __name__ == '__main__':
  single(1024, 512, 512, P.EPI_PLAIN)
  single(1024, 512, 512, P.EPI_LN, 1)
  single(1024, 1536, 1024, P.EPI_LN, 0)
  single(1024, 1024, 512, P.EPI_PLAIN)
  single(16384, 512, 512, P.EPI_LN, 1)
  single(16384, 512, 1536, P.EPI_LN, 1)
  chain(1024, 512, 512, 4)
  chain(1024, 512, 512, 16)
  chain(1024, 1536, 512, 4)
  chain(16384, 1536, 512, 4)
  chain(2048, 1536, 512, 4)
  chain(4096, 1536, 512, 4)
