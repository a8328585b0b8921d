"""Summarise an ncu --set full report (one or more launches) as text for profiles/.

    python tools/ncu_summary.py <report.ncu-rep> <units per launch> <unit name> [algorithmic bytes per unit]
"""
import csv, subprocess, sys

rep, units, uname = sys.argv[1], float(sys.argv[2]), sys.argv[3]
alg = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, un = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active"]
S = "smsp__average_warps_issue_stalled_"
print(f"# ncu --set full --clock-control none: {rep.split('/')[-1]}, {units:g} {uname} per launch")
for k, r in enumerate(rows[2:]):
    v, u = dict(zip(hdr, r)), dict(zip(hdr, un))
    print(f"\n## launch {k}")
    for key in KEYS:
        if key in v:
            print(f"{key:74s} {v[key]} {u.get(key, '')}")
    st = sorted(((float(v[h]), h[len(S):].replace("_per_issue_active.ratio", "")) for h in hdr
                 if h.startswith(S) and h.endswith("per_issue_active.ratio") and v[h]), reverse=True)
    print("warps stalled per issue-active cycle: " + ", ".join(f"{n}={x:.2f}" for x, n in st[:9]))

    def num(key):
        return float(v[key]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u.get(key, ""), 1)
    tr = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    line = f"derived: DRAM bytes/{uname} = {tr / units:.0f}"
    if alg:
        line += f" (algorithmic {alg:.0f}: x{tr / units / alg:.2f})"
    line += f"; warp-instr/{uname} = {float(v['smsp__inst_executed.sum']) / units:.1f}"
    line += f"; smem wavefronts/{uname} = {float(v['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']) / units:.1f}"
    line += f"; SM-cycles/{uname} = {float(v['sm__cycles_elapsed.max']) * 148 / units:.1f}"
    print(line)
