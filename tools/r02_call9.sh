#!/bin/bash
# Ninth GPU call of round 2 (one B200): three-step kernel with HALO as a template parameter and the copies of a column issued by
# lane 0 of every warp (population q by warp q mod 4) instead of by one thread.
R=${1:-r02i}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_step_kernel_experimental or three_steps_per_pass or two_step" > $O/${R}_pytest_triples.txt 2>&1; step pytest-triples $?
timeout 200 python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:rr,4096x32768:f64:bgk,4096x4096:f64:bgk,2048x2048:f64:bgk,2048x2048:f64:trt,1024x1024:f64:trt --variants 10 --steps 61 > $O/${R}_pair_ab_triples.jsonl 2>&1; step ab-triples $?
timeout 100 env PLBM_MULTI_NT=256 python tools/pair_ab.py --cases 8192x8192:f64:bgk,4096x32768:f64:bgk --variants 10 --steps 61 >> $O/${R}_pair_ab_triples.jsonl 2>&1; step ab-triples-256 $?
timeout 300 python bench.py > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2> $O/${R}_bench.err; step bench-c5 $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_bgk_f64_c5 \
    python tools/pair_ab.py --cases 4096x32768:f64:bgk --variants 0 --once > /dev/null 2>&1; step ncu-lbmn3-c5 $?
cat $S
