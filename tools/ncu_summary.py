"""Summarise an .ncu-rep (raw page) into the handful of counters we track: python tools/ncu_summary.py rep [ids]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ids = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(len(data)))
def col(name):
    return hdr.index(name) if name in hdr else None
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__cycles_elapsed.max"]
for w in want:
    c = col(w)
    if c is None:
        continue
    print(f"{w:82s} {units[c]:10s}", [data[i][c][:26] for i in ids])
st = []
for c, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            st.append((max(float(data[i][c]) for i in ids), h, [data[i][c][:6] for i in ids]))
        except ValueError:
            pass
for v, h, vals in sorted(st, reverse=True)[:9]:
    print(f"{h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''):40s}", vals)
