"""Aggregate the ncu source page (csv) of one kernel by address region and stall reason.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [lo_hex hi_hex]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][ia], 16)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 40
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sel = [r for r in data if lo <= int(r[ia], 16) - base <= hi]
tot = sum(int(r[isamp]) for r in sel)
print(f"region {lo:#x}..{hi:#x}: {len(sel)} instrs, {tot} samples, {sum(int(r[iex]) for r in sel)} warp-instr executed")
agg = collections.Counter()
for r in sel:
    for i in stall_cols:
        agg[hdr[i]] += int(r[i] or 0)
print("  by reason:", ", ".join(f"{k[6:]}:{v}" for k, v in agg.most_common(10)))
byop = collections.Counter()
for r in sel:
    s = r[isrc].split()
    op = s[1] if s[0].startswith("@") else s[0]
    byop[op.split(".")[0]] += int(r[isamp])
print("  by opcode:", ", ".join(f"{k}:{v}" for k, v in byop.most_common(12)))
for r in sorted(sel, key=lambda r: -int(r[isamp]))[:12]:
    print(f"   {int(r[ia], 16) - base:#7x} {r[isamp]:>5} {r[iex]:>7}  {r[isrc].strip()[:80]}")
