"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` output.
usage: ncu_hot.py file.csv [kernel-substring] [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sub = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
kern, hdr, body = None, None, []
kernels = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if kern is not None:
            kernels.append((kern, hdr, body))
        kern, hdr, body = r[1], None, []
    elif r and r[0] == "Address":
        hdr = r
    elif kern is not None and hdr is not None and r:
        body.append(r)
if kern is not None:
    kernels.append((kern, hdr, body))
for kern, hdr, body in kernels:
    if sub not in kern:
        continue
    si = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(b[si]) for b in body)
    print(f"== {kern[:100]}  total samples {tot}, {len(body)} instructions")
    agg = {}
    for b in body:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(b[i])
    print("   stall mix:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:topn]
    for i in sorted(order):
        b = body[i]
        st = sorted(((int(b[j]), hdr[j][6:]) for j in stall_cols), reverse=True)[:2]
        print(f"  [{i:5d}] {100*int(b[si])/max(tot,1):5.1f}%  exec {b[ex]:>8s}  {b[src].strip()[:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
