"""CPU: compact text summary of an .ncu-rep (ncu -i ... --page raw --csv on stdin)."""
import csv, sys
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
  d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
  print('kernel:', d['Kernel Name'][:110])
  for w in WANT:
    if w in d:
      print(f'  {w} = {d[w]} {u[w]}')
  stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                   if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio') and v not in ('', 'n/a')), reverse=True)[:6]
  for v, k in stalls:
    print(f'  stall {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")} = {v:.3f}')
  print()
