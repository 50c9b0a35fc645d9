"""Summarise one .ncu-rep (first kernel) into the metrics DESIGN.md / profiles/ quote.  Usage: ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    lines.append(f"kernel: {d.get('Kernel Name')}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
            'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_config_size',
            'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
            'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
            'lts__t_sectors_srcunit_tex_op_write.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
    for k in keys:
        if k in d:
            lines.append(f"  {k:70s} {d[k]:>16s} {u[k]}")
    st = sorted(((float(d[h]), h) for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h and d[h] not in ('', 'n/a')), reverse=True)
    lines.append("  top stall reasons (warps stalled per issue-active cycle):")
    for v, h in st[:8]:
        lines.append(f"    {v:8.3f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
