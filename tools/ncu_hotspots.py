"""Reduce the SASS page of an ncu capture (`ncu -i X.ncu-rep --page source --csv --print-source sass > page.csv`) to what
profiles/*_source_hotspots.txt hold: stall reasons over all samples, the instructions with the most samples, and samples /
executed instructions / shared-memory wavefronts per 1 KB of code.  usage: python tools/ncu_hotspots.py page.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernel = rows[0][1]
h = rows[1]
recs = [dict(zip(h, r)) for r in rows[2:] if len(r) == len(h)]
base = int(recs[0]["Address"], 16)
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"]) for r in recs)
inst = sum(int(r["Instructions Executed"]) for r in recs)
print("# %s\n# total warp-level samples %d, warp instructions executed %d" % (kernel, tot, inst))
print("# --- stall reasons, share of all samples")
for s in stalls:
    v = sum(int(r[s]) for r in recs)
    if v * 1000 >= tot:
        print("%-24s %5.1f%%" % (s, 100.0 * v / tot))
print("# --- %d instructions with the most samples: offset, share of samples, executions, SASS, top two stall reasons" % top)
for r in sorted(recs, key=lambda r: -int(r["# Samples"]))[:top]:
    st = sorted(((int(r[s]), s[6:]) for s in stalls), reverse=True)[:2]
    print("%05x %5.2f%% ex=%8.1fM %-60s %s" % (int(r["Address"], 16) - base, 100.0 * int(r["# Samples"]) / tot, int(r["Instructions Executed"]) / 1e6,
                                            r["Source"].strip()[:60], " ".join("%s=%d" % (n, v) for v, n in st)))
print("# --- samples, executed instructions and shared-memory wavefronts per 1 KB of code (blocks with > 0.4 % of the samples)")
blocks = {}
for r in recs:
    b = (int(r["Address"], 16) - base) // 1024
    x = blocks.setdefault(b, [0, 0, 0, 0])
    x[0] += int(r["# Samples"]); x[1] += int(r["Instructions Executed"])
    x[2] += int(r["L1 Wavefronts Shared"] or 0); x[3] += int(r["L1 Wavefronts Shared Ideal"] or 0)
for b in sorted(blocks):
    x = blocks[b]
    if x[0] * 250 >= tot:
        print("%05x samples %5.2f%%  instructions %6.2f G (%4.1f%%)  shared wavefronts %5.2f G (ideal %5.2f G)" %
              (b * 1024, 100.0 * x[0] / tot, x[1] / 1e9, 100.0 * x[1] / inst, x[2] / 1e9, x[3] / 1e9))
