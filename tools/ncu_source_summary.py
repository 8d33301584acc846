"""Summarise `ncu --page source --csv` of a conv3x3_tc2 capture: stall samples by kernel region and reason, and the
hottest SASS lines.  Usage: python tools/ncu_source_summary.py src.csv [top]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        samples = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    data.append((r[col["Address"]], r[col["Source"]], samples, {s: int(r[col[s]] or 0) for s in stalls},
                 int(r[col["Instructions Executed"]] or 0)))
total = sum(d[2] for d in data)
print("total samples", total, "instructions", len(data))
# regions: split at the role markers we can recognise in SASS
def region(idx, src):
    return None
agg = {}
for a, src, n, st, ex in data:
    for s, v in st.items():
        agg[s] = agg.get(s, 0) + v
print("stall reasons (all warps):")
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
    print("  %-26s %7d  %5.1f%%" % (s, v, 100.0 * v / max(total, 1)))
print("hottest lines:")
for i in sorted(range(len(data)), key=lambda i: -data[i][2])[:top]:
    a, src, n, st, ex = data[i]
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("  #%4d %6d %5.1f%%  ex=%8d  %-70s %s" % (i, n, 100.0 * n / max(total, 1), ex, src[:70], ", ".join("%s=%d" % b for b in best if b[1])))
# cumulative by index blocks of 100 instructions: shows which part of the kernel is hot
print("samples per 100-instruction block:")
for b in range(0, len(data), 100):
    n = sum(d[2] for d in data[b:b + 100])
    if n * 50 > total:
        print("  [%4d, %4d) %6d %5.1f%%" % (b, b + 100, n, 100.0 * n / max(total, 1)))
