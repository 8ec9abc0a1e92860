"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py <report> [kernel-regex]."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
# a report, or the raw page already exported to CSV on the GPU box (tools/ncu_round2.sh)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__cycles_elapsed.avg']
for r in rows[2:]:
    print("=" * 100)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:72s} {r[i]} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st.append((float(r[i]), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print("stall cycles per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
