#!/bin/bash
# One GPU call that produces what profiles/ holds for a round: the GPU test run, a default bench run, the reference arm,
# the ncu launch list of the same bench command, and one `--set full` capture of the step's kernels.
# usage (on the GPU box, from the repo root): bash tools/round_profile.sh <tag>
tag=${1:-r01}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; tail -2 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || tail -5 $out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err || tail -5 $out/bench_ref_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $out/ncu_list_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|ce_tc_kernel|gru_|hop_|tc_gemm_ws|seg_reduce|embed_gather" \
    -s 110 -c 40 -o $out/step_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $out/ncu_full_$tag.log 2>&1
ls -la $out | grep $tag
