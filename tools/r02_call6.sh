#!/bin/bash
# Sixth GPU call of round 2 (one B200): three steps per pass at twice the warps (k_lbmn_bulk wide shape, PLBM_MULTI_WIDE=1): ncu of the
# 128-thread shape shows 8 warps per SM, fp64 pipe 47 %, DRAM 62 % -- bound by latency, not by a pipe.  Plus the new default paths:
# TRT fp64 triples from 2048^2, sum-form marching kernel in fp32.
R=${1:-r02f}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 200 env PLBM_MULTI_WIDE=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_step_kernel_experimental" > $O/${R}_pytest_multi_wide.txt 2>&1; step pytest-multi-wide $?
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "three_steps_per_pass or two_step" > $O/${R}_pytest_trt_triples.txt 2>&1; step pytest-trt-triples $?
timeout 200 python -m pytest tests/test_gpu_fast_variants.py -m gpu -q > $O/${R}_pytest_fast.txt 2>&1; step pytest-fast $?
for wide in 0 1; do
    timeout 240 env PLBM_MULTI_WIDE=$wide python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:trt,8192x8192:f32:rr,4096x32768:f64:bgk,4096x4096:f64:bgk,2048x2048:f64:bgk --variants 10 --steps 61 >> $O/${R}_pair_ab_triples_wide.jsonl 2>&1; step ab-triples-wide-$wide $?
done
timeout 100 python tools/pair_ab.py --cases 8192x8192:f64:bgk,4096x32768:f64:bgk,4096x4096:f64:bgk,2048x2048:f64:bgk --variants 0 --steps 61 >> $O/${R}_pair_ab_triples_wide.jsonl 2>&1; step ab-pairs $?
timeout 120 env PLBM_MULTI_WIDE=1 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_wide_bgk_f64_8192 \
    python tools/pair_ab.py --cases 8192x8192:f64:bgk --variants 10 --once > /dev/null 2>&1; step ncu-lbmn3-wide $?
cat $S
