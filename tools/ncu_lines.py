"""Per-source-line totals from `ncu -i rep --page source --print-source cuda,sass --csv`:
samples, warp-instructions, shared-memory wavefronts (ideal / excessive).
usage: python tools/ncu_lines.py src.csv [top_n] [sort key: wf|ex|smp|inst]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "wf"
out, fname, hdr = [], "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ix = {h: k for k, h in enumerate(hdr) if h not in ("Source",)}
    elif hdr and r[0] != "-" and len(r) > 10 and r[2] == "-":      # a source line's aggregate row
        g = lambda k: float(r[ix[k]] or 0) if r[ix[k]] not in ("-", "") else 0.0
        out.append(dict(f=fname, ln=r[0], src=r[1].strip(), smp=g("# Samples"), inst=g("Instructions Executed"),
                        wf=g("L1 Wavefronts Shared"), ex=g("L1 Wavefronts Shared Excessive")))
T = {k: sum(o[k] for o in out) or 1 for k in ("smp", "inst", "wf", "ex")}
print(f"totals: samples {T['smp']:.0f}  warp-inst {T['inst']:.3g}  smem wavefronts {T['wf']:.3g}  excessive {T['ex']:.3g}")
for o in sorted(out, key=lambda o: -o[key])[:topn]:
    print(f"{o['f'][:14]:14s}:{o['ln']:>4s} wf {100 * o['wf'] / T['wf']:5.1f}% ex {100 * o['ex'] / T['wf']:5.1f}% "
          f"smp {100 * o['smp'] / T['smp']:5.1f}% inst {100 * o['inst'] / T['inst']:5.1f}%  {o['src'][:90]}")
