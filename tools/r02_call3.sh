#!/bin/bash
# Third GPU call of round 2 (one B200): the "wide" shape of k_lbm2_bulk (8 bytes per thread, 256 threads: twice the warps on the
# same shared memory; PLBM_BULK_WIDE=1) -- bit parity, then A/B against the default shape.
R=${1:-r02c}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 400 env PLBM_BULK_WIDE=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q > $O/${R}_pytest_wide.txt 2>&1; step pytest-wide $?
for wide in 0 1; do
    timeout 200 env PLBM_BULK_WIDE=$wide python tools/pair_ab.py --cases 8192x8192:f32:bgk,8192x8192:f32:trt,8192x8192:f32:rr,8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,4096x32768:f64:bgk,2048x2048:f64:bgk,2048x2048:f32:rr,1024x1024:f64:trt --variants 0 >> $O/${R}_pair_ab_wide.jsonl 2>&1; step ab-wide-$wide $?
done
timeout 100 env PLBM_BULK_WIDE=1 python tools/pair_ab.py --cases 8192x8192:f32:bgk,8192x8192:f32:rr --variants 12 >> $O/${R}_pair_ab_wide.jsonl 2>&1; step ab-wide-scalar $?
timeout 120 env PLBM_BULK_WIDE=1 ncu --set full --clock-control none --import-source on -k regex:k_lbm2_bulk -c 1 -f -o $O/${R}_k_lbm2_bulk_rr_f32_wide_8192 \
    python tools/pair_ab.py --cases 8192x8192:f32:rr --variants 0 --once > /dev/null 2>&1; step ncu-rr-f32-wide $?
cat $S
