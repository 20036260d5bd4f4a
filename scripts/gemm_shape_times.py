"""CPU: measured time per GEMM shape of one a1 train step.

Replays the engine's launch sequence on a recording backend (no GPU) to get the ORDERED list of
dd_gemm calls, and lines it up with the ordered GEMM-kernel launches of the ncu launch list
(`ncu --metrics gpu__time_duration.sum`, scripts/gpu_prof.sh).  Under ncu every launch is serialised
and cold-cache, so these are upper bounds per launch; the in-graph step is ~20 % faster overall."""
import collections, csv, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daydreamer_b200 import engine as E


class Recorder:
  device = torch.device('cpu')
  def __init__(self): self.gemms, self.seq, self.rows = [], [], []
  def zeros(self, *shape, dtype=torch.float32): return torch.empty(*shape, dtype=dtype)
  def gemm(self, A, B, C, ta=False, tb=False, beta=0.0, bias=None):
    K = A.shape[0] if ta else A.shape[1]
    self.seq.append(len(self.rows))           # row-op calls issued before this GEMM
    self.gemms.append((C.shape[0], C.shape[1], K, int(ta), int(tb), int(beta), int(bias is not None)))
  def ln_fwd(self, Z, scale, bias, H, stats, act):
    self.rows.append(('ln_fwd', Z.shape[0], Z.shape[1], 2))       # read Z, write H
  def ln_bwd(self, dH, Z, stats, scale, bias, dZ, dscale, dbias, act):
    self.rows.append(('ln_bwd', Z.shape[0], Z.shape[1], 3))       # read dH, Z, write dZ
  def observe_fwd_grid(self, *a): return 128
  def observe_bwd_grid(self, *a): return 128
  def __getattr__(self, name):
    return lambda *a, **k: None


def expected_kernel(M, N, K, ta, tb, beta, bias):
  """Kernel dd_gemm dispatches this call to (dd_gemm.cu / dd_gemm_tc.cu / dd_gemm_skinny.cu)."""
  if not ta and M <= 32 and N >= 16 and K >= 64:
    return f'skinny_{"nt" if tb else "nn"}_kernel<{16 if M <= 16 else 32}>'
  tc = M > 64 and N >= 32 and K >= 32 and (M if ta else K) % 4 == 0 and (K if tb else N) % 4 == 0
  if tc:
    t128 = -(-M // 128) * -(-N // 128)
    split_ok = beta != 1 or bias or K <= 32768
    if t128 > 222 and split_ok:
      return f'gemm_tc_persistent_kernel<128, 3, {ta}, {tb}>'
    if t128 >= 74:
      return f'gemm_tc_kernel<128, 3, {ta}, {tb}>'
    return f'gemm_tc_kernel<64, 4, {ta}, {tb}>'
  if M <= 16:
    return f'gemm_kernel<16, 64, 16, 1, 4, {ta}, {tb}>'
  if -(-M // 128) * -(-N // 128) >= 296:
    return f'gemm_kernel<128, 128, 16, 8, 8, {ta}, {tb}>'
  return f'gemm_kernel<64, 64, 16, 4, 4, {ta}, {tb}>'


def main(path):
  rec = Recorder()
  eng = E.Engine(E.Spec(obs_dims={'vector': 16}), rec, 16, 64)
  rec.gemms.clear()
  eng.launch_step(slow_refreshed=False)
  lines = [l for l in open(path) if not l.startswith('==')]
  launches = []
  for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
      continue
    v = float(r['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r.get('Metric Unit', 'ns'), 1e-3)
    launches.append((re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('dd::', ''), v))
  gk = [(n, v) for n, v in launches if re.search(r'gemm_tc|gemm_kernel|skinny_n|thin_', n)]
  print(f'# {len(rec.gemms)} dd_gemm calls in the replayed step, {len(gk)} GEMM-kernel launches in {path}')
  if len(gk) != len(rec.gemms):
    print('# counts differ: the launch list is not one whole step of this engine version; aligned prefix only')
  # the capture window is one step long but may start mid-step: rotate the replayed sequence to
  # the offset where the kernel each call dispatches to (csrc/dd_gemm.cu rules) matches best
  want = [expected_kernel(*g) for g in rec.gemms]
  names = [re.sub(r'^tc::', '', n) for n, _ in gk]
  n = len(want)
  score = lambda r: sum(want[(i + r) % n] == names[i] for i in range(min(n, len(names))))
  best = max(range(n), key=score)
  print(f'# window starts at call {best} of the step; {score(best)}/{len(names)} launches match the expected kernel')
  rec.gemms = rec.gemms[best:] + rec.gemms[:best]
  row0 = rec.seq[best]
  by = collections.OrderedDict()
  for shape, (name, us) in zip(rec.gemms, gk):
    key = shape + (re.sub(r'^tc::', '', name),)
    by.setdefault(key, []).append(us)
  tot = sum(sum(v) for v in by.values())
  print('# M      N     K      ta tb beta bias  x    mean us   total ms  share   TFLOP/s  kernel')
  for (M, N, K, ta, tb, beta, bias, name), v in sorted(by.items(), key=lambda kv: -sum(kv[1])):
    mean = sum(v) / len(v)
    print(f'{M:6d} {N:5d} {K:6d}  {ta}  {tb}  {beta}    {bias}  {len(v):4d} {mean:9.1f} {sum(v)/1e3:9.3f} {100*sum(v)/tot:6.1f}% '
          f'{2.0*M*N*K/mean/1e6:8.1f}  {name}')
  print(f'# total {tot/1e3:.2f} ms over {sum(len(v) for v in by.values())} launches')
  # LayerNorm(+elu) kernels: HBM/L2 streaming, algorithmic bytes = passes * M * N * 4
  rows = rec.rows[row0:] + rec.rows[:row0]
  lk = [(n, v) for n, v in launches if re.match(r'ln_(fwd|bwd)', n)]
  print(f'#\n# {len(rows)} LayerNorm calls in the replayed step, {len(lk)} ln_* launches in the list')
  agg = collections.OrderedDict()
  ok = 0
  for (kind, M, N, passes), (name, us) in zip(rows, lk):
    ok += name.startswith(kind)
    agg.setdefault((kind, M, N, passes), []).append(us)
  print(f'# {ok}/{len(lk)} launches match the expected kernel family')
  print('# kernel   M      N   x   mean us  total ms   GB/s (algorithmic bytes / time; measured copy peak 6573)')
  for (kind, M, N, passes), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    mean = sum(v) / len(v)
    print(f'{kind}  {M:6d} {N:5d} {len(v):4d} {mean:8.1f} {sum(v)/1e3:8.3f} {passes*M*N*4/mean/1e3:8.0f}')


if __name__ == '__main__':
  main(sys.argv[1] if len(sys.argv) > 1 else 'profiles/r01_launches_one_step.csv')
