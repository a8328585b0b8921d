"""Turn a gpurun_out ncu capture into the committed summaries under profiles/.

    python profiles/summarize.py <tag> <round-label> [clips_in_capture]

Reads gpurun_out/prof_<tag>.ncu-rep (ncu --set full of the headline kernel), gpurun_out/
launches_<tag>.csv (gpu__time_duration per launch) and gpurun_out/bench_<tag>.json; writes
profiles/<round>_ncu_summary.txt, profiles/<round>_launches.csv, profiles/<round>_bench.json and
profiles/traffic.json (DRAM bytes per clip of the dominant kernel, read by bench.py)."""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag, rnd = sys.argv[1], sys.argv[2]
clips = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
G = ROOT / "gpurun_out"
rep = G / f"prof_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"
lines = [f"# ncu --set full --clock-control none, kernel captured with {clips} clips per launch (tag {tag})", ""]
traffic = None
for k, r in enumerate(rows[2:]):
    v = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    lines.append(f"## launch {k}")
    for key in KEYS:
        if key in v:
            lines.append(f"{key:78s} {v[key]} {u.get(key, '')}")
    st = sorted(((float(v[h]), h[len(STALL):].replace('_per_issue_active.ratio', '')) for h in hdr
                 if h.startswith(STALL) and h.endswith("per_issue_active.ratio") and v[h]), reverse=True)
    lines.append("warps stalled per issue-active cycle: " + ", ".join(f"{n}={x:.2f}" for x, n in st[:9]))
    def num(key):
        x = float(v[key]); un = u.get(key, "")
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(un, 1)
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    frames = clips * 512
    lines.append(f"derived: DRAM bytes/clip = {(rd + wr) / clips:.0f} (algorithmic 240160); "
                 f"warp-instr/frame = {float(v['smsp__inst_executed.sum']) / frames:.1f}; "
                 f"smem wavefronts/frame = {float(v['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']) / frames:.1f}")
    lines.append("")
    traffic = {"dram_bytes_per_clip": (rd + wr) / clips, "clips_in_capture": clips, "kernel": v["Kernel Name"],
               "note": f"ncu --set full, profiles/{rnd}_ncu_summary.txt"}
(ROOT / "profiles" / f"{rnd}_ncu_summary.txt").write_text("\n".join(lines))
if traffic:
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(traffic, indent=1))
for src, dst in ((G / f"launches_{tag}.csv", f"{rnd}_launches.csv"), (G / f"bench_{tag}.json", f"{rnd}_bench.json")):
    if src.exists():
        shutil.copy(src, ROOT / "profiles" / dst)
# SASS evidence: TMA bulk copy + mbarrier in the headline kernel
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "audio_edge_ml_pipeline_b200" / "libb2a.so")],
                      capture_output=True, text=True).stdout
cnt = collections.Counter()
cur = None
MNEMONICS = ("UBLKCP", "SYNCS", "USETMAXREG", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "LDS", "STS", "SHFL", "BAR")
for ln in sass.splitlines():
    if "Function :" in ln:
        cur = ln.split("Function :")[1].strip()
        continue
    if not (cur and "logmel512_kernelILb1ELi0ELb1ELb0" in cur):
        continue
    tok = ln.split()
    for t in tok[1:3]:                      # opcode is the 2nd token, or the 3rd after a predicate
        op = t.split(".")[0].rstrip(";")
        if op in MNEMONICS:
            cnt[op] += 1
            break
(ROOT / "profiles" / f"{rnd}_sass_evidence.txt").write_text(
    "logmel512_kernel<int16, mel, generated-mel> static SASS mnemonic counts (cuobjdump -sass libb2a.so):\n" +
    "\n".join(f"  {k}: {v}" for k, v in sorted(cnt.items())) +
    "\nUBLKCP = cp.async.bulk (TMA bulk copy); SYNCS = mbarrier init/arrive/try_wait; USETMAXREG = setmaxnreg\n"
    "(FFT warps 104 registers, mel/producer warps 64); FFMA2/FADD2/FMUL2 = packed FP32 (two lanes of a\n"
    "complex point per instruction); BAR = the mel warps' named barrier + the one-off prologue barrier.\n")
print("\n".join(lines[:40]))
