"""Print the last N rows (kernel, us) of an ncu gpu__time_duration launch list CSV."""
import csv, sys
path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(x["Kernel Name"][:70], float(x["Metric Value"].replace(",", "")) / (1000 if x["Metric Unit"] == "ns" else 1))
        for x in csv.DictReader(lines)]
for k, v in rows[-n:]:
    print(f"{v:10.2f} us  {k}")
