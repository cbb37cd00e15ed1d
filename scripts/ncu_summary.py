"""Dump the metrics we track from an ncu report (run here, no GPU needed)."""
import csv, subprocess, sys, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [
 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg.per_second',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active',
 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
 'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
 'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait',
 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_not_selected',
 'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_membar',
 'smsp__pcsamp_warps_issue_stalled_sleeping', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
 'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
 'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
 'smsp__pcsamp_warps_issue_stalled_misc', 'smsp__pcsamp_warps_issue_stalled_imc_miss',
 'smsp__pcsamp_warps_issue_stalled_tex_throttle', 'smsp__pcsamp_warps_issue_stalled_drain',
 'smsp__pcsamp_warps_issue_stalled_gmma', 'smsp__pcsamp_sample_buffers',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
 'dram__bytes_read.sum.per_second', 'l1tex__m_xbar2l1tex_read_bytes.sum',
 'launch__cluster_dim_x', 'launch__cluster_size',
]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index('Kernel Name')]}
    for w in want:
        if w in hdr:
            d[w] = f"{r[hdr.index(w)]} {units[hdr.index(w)]}".strip()
    out.append(d)
print(json.dumps(out, indent=1))
