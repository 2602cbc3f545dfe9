#!/bin/bash
# One GPU call that produces what profiles/ holds for a round: the GPU test run, smoke, a default bench run, the reference
# arm, the ncu launch list of the same bench command (one eager step), `--set full` captures of the step's kernels and of
# the two graded bandwidth kernels (gather, scatter-add = sort + segmented reduce) at cfg-4 shapes, and a SASS listing.
# usage (on the GPU box, from the repo root): bash tools/round_profile.sh <tag>
tag=${1:-r02}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; tail -2 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || tail -5 $out/bench_$tag.err
python bench.py --impl reference > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err || tail -5 $out/bench_ref_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-cfg4 > $out/ncu_list_$tag.log 2>&1
# (the .ncu-rep files stay on the box: gpurun_out/ may carry 64 MiB back, so only the raw-page CSV exports travel)
ncu --set full --clock-control none -k regex:"adam_|ce_tc_kernel|gru_|hop_|tc_gemm_ws|embed_gather" \
    -s 60 -c 40 -o /tmp/step_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-cfg4 > $out/ncu_full_$tag.log 2>&1
ncu -i /tmp/step_$tag.ncu-rep --page raw --csv > $out/step_${tag}_ncu_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"gather_rows|os_hist|os_pass|seg_reduce" -c 8 \
    -o /tmp/bw_$tag -f python tools/prof_scatter.py > $out/ncu_bw_$tag.log 2>&1
ncu -i /tmp/bw_$tag.ncu-rep --page raw --csv > $out/bw_${tag}_ncu_full_raw.csv 2>/dev/null
cuobjdump -sass mtamrecommender_b200/libmtam_b200.so | grep -E "Function :|UTCHMMA|UTCBAR|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|LDGSTS|SYNCS" \
    | awk '/Function :/{f=$3; next} {split($2,a,"."); c[f" "a[1]]++} END{for(k in c) print c[k], k}' | sort -k2,2 -k3,3 > $out/sass_$tag.txt
ls -la $out | grep $tag
