"""Reads `ncu --page raw --csv` on stdin, prints the handful of metrics the roofline argument needs."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
  print('no data')
  sys.exit(0)
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_tensor_op_hmma.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
for r in rows[2:]:
  d = dict(zip(hdr, r))
  u = dict(zip(hdr, units))
  for k in want:
    if k in d:
      print(f'{k:90s} {d[k]:>20s} {u.get(k, "")}')
  extra = [k for k in hdr if ('tensor' in k or 'hmma' in k.lower()) and k not in want]
  for k in extra[:12]:
    print(f'{k:90s} {d[k]:>20s} {u.get(k, "")}')
  print()
