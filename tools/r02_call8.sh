#!/bin/bash
# Eighth GPU call of round 2 (one B200): three steps per pass as the default for fp64 BGK / TRT from 2048^2 nodes (one row per
# thread, 128-thread blocks): the whole -m gpu suite, smoke, the headline bench line, A/B, ncu of the bench shape.
R=${1:-r02h}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 700 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.txt 2>&1; step pytest-gpu $?
timeout 60 python __graft_entry__.py smoke > $O/${R}_smoke.txt 2>&1; step smoke $?
timeout 300 python bench.py > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2> $O/${R}_bench.err; step bench-c5 $?
timeout 300 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1_c5_bgk_f64_slab_k20.json 2>> $O/${R}_bench.err; step bench-c5-k20 $?
timeout 200 python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f64:trt,4096x32768:f64:bgk,4096x4096:f64:bgk,4096x4096:f64:trt,2048x2048:f64:bgk,2048x2048:f64:trt --variants 0,7 --steps 61 > $O/${R}_pair_ab_default.jsonl 2>&1; step ab-default $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_bgk_f64_c5 \
    python tools/pair_ab.py --cases 4096x32768:f64:bgk --variants 0 --once > /dev/null 2>&1; step ncu-lbmn3-c5 $?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_bench_n1_c5_bgk_f64.csv \
    python bench.py --steps 7 --warmup 3 --no-cpu > /dev/null 2>&1; step ncu-launch-list $?
cat $S
