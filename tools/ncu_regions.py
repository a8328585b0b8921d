"""Bucket an ncu source-page CSV (ncu -i rep --page source --csv) by code region and list hot instructions.
usage: python tools/ncu_regions.py src.csv [bucket_bytes] [top_n]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x400
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hdr, data = rows[1], rows[2:]
ia, isrc, isamp, iex, iw = (hdr.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed', 'L1 Wavefronts Shared'))
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(data[0][ia], 16)
tot = sum(int(r[isamp] or 0) for r in data); totex = sum(int(r[iex] or 0) for r in data)
print('total samples', tot, 'warp-inst', totex)
bs = defaultdict(lambda: [0, 0, 0])
for r in data:
    b = (int(r[ia], 16) - base) // B
    bs[b][0] += int(r[isamp] or 0); bs[b][1] += int(r[iex] or 0); bs[b][2] += int(r[iw] or 0)
for b in sorted(bs):
    s, e, w = bs[b]
    if s * 200 > tot or e * 200 > totex:
        print(f"{b * B:#7x} samples {100 * s / tot:5.1f}%  inst {100 * e / totex:5.1f}%  smem-wf {w / 1e6:8.1f}M")
print()
for r in sorted(sorted(data, key=lambda r: -int(r[isamp] or 0))[:topn], key=lambda r: int(r[ia], 16)):
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print(f"{int(r[ia], 16) - base:#7x} {100 * int(r[isamp]) / tot:5.2f}% ex={int(r[iex]) / 1e6:7.1f}M  {r[isrc][:70]:70s} {st}")
