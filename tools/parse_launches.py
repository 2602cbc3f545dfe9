"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one train step's kernels in order,
and the per-kernel-name totals.  usage: parse_launches.py file.csv [anchor-kernel-prefix]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
anchor = sys.argv[2] if len(sys.argv) > 2 else "embed_gather"
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[start:]:
    if len(r) > vi:
        try:
            seq.append((r[ki], float(r[vi].replace(",", ""))))
        except ValueError:
            pass
names = [s[0] for s in seq]
idxs = [i for i, n in enumerate(names) if anchor in n.split("(")[0]]
a, b = idxs[-2], idxs[-1]
tot = 0.0
agg = collections.OrderedDict()
for n, t in seq[a:b]:
    key = n.split("(")[0][:70]
    c = agg.setdefault(key, [0, 0.0]); c[0] += 1; c[1] += t
    tot += t
if "-v" in sys.argv:
    for n, t in seq[a:b]:
        print(f"{t/1000:9.1f} us  {n[:100]}")
print(f"step: {b-a} launches, {tot/1e6:.3f} ms (ncu, cold-cache, serialised)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1000:9.1f} us {100*t/tot:5.1f}%  x{c:<3d} {k}")
