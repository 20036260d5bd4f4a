"""GPU: where the layer machine waits.  Runs one program under each debug switch and prints the
per-CTA average wait cycles by wait code (1 producer:empty 2 producer:ready 3 mma:acc_empty
4 mma:full 5 epi:stats 6 epi:acc_full 7 epi:ready-order 8 epi:ready-input; 0 = kernel cycles)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from daydreamer_b200 import program as P
from daydreamer_b200.backend import CudaBackend

be = CudaBackend()
dev = be.device
NAMES = ['kernel', 'p:empty', 'p:ready', 'm:acce', 'm:full', 'e:stats', 'e:accf', 'e:rdy7', 'e:rdy8']


def build(M, K0, U, nl, epi):
  X = torch.randn(M, K0, device=dev); Xlo = P.lo_of(X)
  layers, keep = [], []
  prev = (X, Xlo)
  for i in range(nl):
    K = K0 if i == 0 else U
    Wt0 = (torch.randn(K, U, device=dev) * K ** -0.5).t().contiguous()
    pairw = torch.stack([Wt0, P.lo_of(Wt0)])
    Wt, Wtlo = pairw[0], pairw[1]
    g, b = torch.ones(U, device=dev), torch.zeros(U, device=dev)
    z, out, lo, st = (torch.zeros(M, U, device=dev), torch.zeros(M, U, device=dev),
                      torch.zeros(M, U, device=dev), torch.zeros(M, 2, device=dev))
    layers.append(P.Layer([P.Seg(*prev)], Wt, Wtlo, U, epi, act=1, gamma=g, beta=b, z=z, st=st, out=out, out_lo=lo))
    prev = (out, lo)
  return be.program_build(layers, M)


def run(tag, h, flags, iters=10, trace=False):
  be.lib.dd_set_machine_debug(flags)
  be.lib.dd_machine_prof(0, None, 0)
  for _ in range(3):
    be.program_run(h)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    be.program_run(h)
  b.record()
  torch.cuda.synchronize()
  us = a.elapsed_time(b) / iters * 1e3
  be.lib.dd_machine_prof(1, None, 0)
  be.program_run(h)
  torch.cuda.synchronize()
  TN = 2048
  buf = (ctypes.c_ulonglong * (148 * 16 + 4 * TN))()
  be.lib.dd_machine_prof(0, buf, 148 * 16 + 4 * TN)
  arr = np.array(buf[:148 * 16], dtype=np.float64).reshape(148, 16)
  tr = np.array(buf[148 * 16:], dtype=np.float64).reshape(4, TN)
  if trace:
    t0 = tr[3, 0]
    rel = lambda x: int(x - t0) if x > 0 else -1
    print('  kernel events: roles end', [rel(x) for x in tr[3, 1:7]], 'after syncthreads', rel(tr[3, 8]),
          'after cluster sync', rel(tr[3, 9]))
    n_it = int((tr[1, ::4] > 0).sum())
    print('  it: B.empty_ok B.issued | A.empty_ok A.issued | M.full_ok M.mmas_issued M.commit_issued M.syncwarp')
    for it in list(range(min(n_it, 10))):
      print(f'  {it:3d}: {rel(tr[0, 4 * it]):7d} {rel(tr[0, 4 * it + 1]):7d} | {rel(tr[0, 4 * it + 2]):7d} {rel(tr[0, 4 * it + 3]):7d} | '
            f'{rel(tr[1, 4 * it]):7d} {rel(tr[1, 4 * it + 2]):7d} {rel(tr[1, 4 * it + 3]):7d} {rel(tr[1, 4 * it + 1]):7d}')
    n_p = int((tr[2, ::16] > 0).sum())
    print('  pair: start accf_ok acc0_loaded z0_stored exch_in exch_out norm0_done h0_stored | epi_done arrived')
    for pr in range(min(n_p, 12)):
      e = [rel(tr[2, 16 * pr + k]) for k in range(12)]
      print(f'  {pr:3d}: ' + ' '.join(f'{x:7d}' for x in e[:8]) + f' | {e[9]:7d} {e[11]:7d}')
  used = arr[arr[:, 0] > 0]
  avg = used.mean(0) if len(used) else np.zeros(16)
  print(f'{tag:34s} flags={flags:2d} {us:8.1f} us  ctas={len(used):3d}  ' +
        ' '.join(f'{n}={avg[i]:.0f}' for i, n in enumerate(NAMES)), flush=True)
  be.lib.dd_set_machine_debug(0)


if __name__ == '__main__':
  for (M, K0, U, nl, epi, tag) in [(1024, 512, 512, 1, P.EPI_PLAIN, 'plain 1024x512x512'),
                                   (1024, 512, 512, 4, P.EPI_LN, 'LN chain x4 M=1024'),
                                   (16384, 512, 512, 1, P.EPI_LN, 'LN 16384x512x512')]:
    h = build(M, K0, U, nl, epi)
    for flags in (0,):
      run(tag, h, flags, trace=(M == 1024))
