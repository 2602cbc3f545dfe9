#!/bin/bash
# The ml-1m-shaped configurations (BASELINE configs[0]/[1]) of every model kind the library builds: one bench line each.
# usage (GPU box, repo root): bash tools/bench_models.sh <tag>
tag=${1:-r02}
for w in cfg1 cfg1_via_t_gru cfg2_pistrec cfg2_ta_sasrec cfg2_tisasrec cfg2_sasrec; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err || tail -3 gpurun_out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${tag}_$w.json"))
    print("$w", d["config"]["model"], "ms/step", round(d["ms_per_step"], 4), "seq/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e:
    print("$w", "FAILED", e)
PY
done
