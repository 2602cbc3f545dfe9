"""Condenses an `ncu --page raw --csv` export into one line per launch with the metrics the rooflines use.
usage: ncu_summary.py raw.csv > summary.txt ; also prints a JSON traffic map with --json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
C = {k: col(k) for k in ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                         "dram__bytes_write.sum", "launch__registers_per_thread",
                         "sm__warps_active.avg.pct_of_peak_sustained_active",
                         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                         "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                         "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                         "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                         "launch__shared_mem_per_block_dynamic"]}
def num(r, k):
    i = C.get(k)
    if i is None: return None
    try: return float(r[i].replace(",", ""))
    except ValueError: return None
def scale(r, k):   # bytes columns come in K/M/G byte units
    i = C.get(k)
    if i is None: return None
    u = units[i].lower()
    v = num(r, k)
    if v is None: return None
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
out = []
for r in rows[2:]:
    name = r[C["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("mtam::", "").replace("<unnamed>::", "")[:44]
    dur_i = C["gpu__time_duration.sum"]; du = units[dur_i]
    dur = num(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(du, 1)
    rd, wr = scale(r, "dram__bytes_read.sum") or 0, scale(r, "dram__bytes_write.sum") or 0
    out.append(dict(kernel=short, grid=r[C["Grid Size"]], block=r[C["Block Size"]], us=dur, dram_read_MB=rd / 1e6, dram_write_MB=wr / 1e6,
                    dram_GBs=(rd + wr) / dur / 1e3 if dur else 0, regs=num(r, "launch__registers_per_thread"),
                    dram_pct=num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    sm_pct=num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                    issue_pct=num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    smem_dyn=num(r, "launch__shared_mem_per_block_dynamic")))
if "--json" in sys.argv:
    print(json.dumps(out, indent=1))
else:
    print(f"{'kernel':44s} {'grid':>14s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'dram%':>6s} {'sm%':>5s} {'issue%':>6s} regs")
    for o in out:
        print(f"{o['kernel']:44s} {o['grid']:>14s} {o['us']:8.1f} {o['dram_read_MB']:8.1f} {o['dram_write_MB']:8.1f} {o['dram_GBs']:7.0f} "
              f"{(o['dram_pct'] or 0):6.1f} {(o['sm_pct'] or 0):5.1f} {(o['issue_pct'] or 0):6.1f} {int(o['regs'] or 0)}")
