#!/bin/bash
# Seventh GPU call of round 2 (one B200): block shapes of the three-step kernel (barrier stall is its top stall: 2.1 per issue at
# 256 threads), then compute-sanitizer memcheck over the kernels added this round.
R=${1:-r02g}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
for shape in 1,32 1,64 1,128 0,32 0,64; do
    IFS=, read wide nt <<< "$shape"
    timeout 100 env PLBM_MULTI_WIDE=$wide PLBM_MULTI_NT=$nt python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_step_kernel_experimental" > $O/${R}_pytest_multi_w${wide}_nt$nt.txt 2>&1; step pytest-multi-w$wide-nt$nt $?
    timeout 200 env PLBM_MULTI_WIDE=$wide PLBM_MULTI_NT=$nt python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:rr,4096x32768:f64:bgk,2048x2048:f64:bgk --variants 10 --steps 61 >> $O/${R}_pair_ab_triples_shapes.jsonl 2>&1; step ab-w$wide-nt$nt $?
done
timeout 120 env PLBM_MULTI_WIDE=1 PLBM_MULTI_NT=32 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_w1_nt32_bgk_f64_8192 \
    python tools/pair_ab.py --cases 8192x8192:f64:bgk --variants 10 --once > /dev/null 2>&1; step ncu-lbmn3-w1-nt32 $?
timeout 900 env PLBM_SANITIZE_VARIANTS=0,3,4,9,10,11,12 compute-sanitizer --tool memcheck python tools/sanitize.py > $O/${R}_sanitizer_memcheck.txt 2>&1; step memcheck $?
cat $S
