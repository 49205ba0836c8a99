#!/bin/bash
# First GPU call of the next round (one B200):  gpurun --timeout 900 -- 'bash tools/round2_first_call.sh r02a'
# 1. what round 1 could not run after k_lbm2_bulk became the default: the whole -m gpu suite, serial, and the 32768^2 case;
# 2. parity gate + A/B of the experimental depth-generic kernel (plbm_lbmn.cu, variants 9 = two, 10 = three steps per pass).
R=${1:-r02a}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 200 env PLBM_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_parity.py -k "multi_step_kernel_experimental" -x -q -m gpu > $O/${R}_pytest_experimental.txt 2>&1; step experimental-parity $?
timeout 200 python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:rr,4096x32768:f64:bgk --variants 7,9,10 > $O/${R}_pair_ab.jsonl 2>&1; step ab $?
timeout 100 env PLBM_MULTI_NT=256 python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f32:bgk,4096x32768:f64:bgk --variants 10 > $O/${R}_pair_ab_nt256.jsonl 2>&1; step ab-nt256 $?
timeout 100 python tools/pair_ab.py --cases 8192x8192:f64:rr,8192x8192:f32:rr,8192x8192:f64:bgk --variants 0,11 > $O/${R}_pair_ab_fma.jsonl 2>&1; step ab-fma $?
# crossover between k_lbm2 and k_lbm2_bulk on mid-size grids (default: bulk from two waves of its blocks upwards)
for c in 1024x1024:f64:trt 2048x2048:f64:bgk 4096x4096:f64:bgk 4096x4096:f32:bgk; do
    timeout 60 python tools/pair_ab.py --cases $c --variants 6 --steps 201 >> $O/${R}_pair_ab_crossover.jsonl 2>&1
    timeout 60 env PLBM_PAIR_BULK=2 python tools/pair_ab.py --cases $c --variants 0 --steps 201 >> $O/${R}_pair_ab_crossover.jsonl 2>&1
    timeout 60 env PLBM_PAIR_BULK=2 PLBM_PAIR_BULK_FILL=1 python tools/pair_ab.py --cases $c --variants 0 --steps 201 >> $O/${R}_pair_ab_crossover.jsonl 2>&1
done; step ab-crossover $?
for sl in 32 128 256; do
    timeout 60 env PLBM_MULTI_SEGLEN=$sl python tools/pair_ab.py --cases 8192x8192:f64:bgk --variants 10 >> $O/${R}_pair_ab_seglen.jsonl 2>&1; step ab-seglen-$sl $?
done
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_bgk_f64_8192 \
    python tools/pair_ab.py --cases 8192x8192:f64:bgk --variants 10 --once > /dev/null 2>&1; step ncu-full-lbmn3 $?
timeout 120 env PLBM_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zz_round1_late.py -q -m gpu > $O/${R}_pytest_late.txt 2>&1; step late-tests $?
for c in dugks,f64,bgk,0 dugks,f64,bgk,3 dugks,f32,bgk,0 dugks,f32,bgk,3 fvm,f64,bgk,0 fvm,f64,bgk,3; do
    timeout 90 python tools/kbench.py --n 2048 --steps 50 --case $c >> $O/${R}_kbench_fma.jsonl 2>&1; step kbench-$c $?
done
timeout 400 python -m pytest tests -m gpu -q -x > $O/${R}_pytest_gpu.txt 2>&1; step pytest-gpu $?
timeout 60 python __graft_entry__.py smoke > $O/${R}_smoke.txt 2>&1; step smoke $?
cat $S
