#!/bin/bash
# Fourth GPU call of round 2 (one B200): small-block shapes.  The two barriers per column keep all warps of a block in the same
# phase (loads / arithmetic / stores), so the pipes idle in turn (ncu r02b/r02c: fma pipe 60 %, math_pipe_throttle the top stall);
# smaller blocks, more of them per SM, decouple the phases.  k_lbm2_bulk: PLBM_BULK_NT=64/32; k_fv_march: PLBM_MARCH_NT=64/32.
R=${1:-r02d}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
for nt in 64 32; do
    timeout 300 env PLBM_BULK_NT=$nt python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q > $O/${R}_pytest_bulk_nt$nt.txt 2>&1; step pytest-bulk-nt$nt $?
done
for nt in 128 64 32; do
    timeout 200 env PLBM_BULK_NT=$nt python tools/pair_ab.py --cases 8192x8192:f32:bgk,8192x8192:f32:rr,8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,4096x32768:f64:bgk,2048x2048:f64:bgk,1024x1024:f64:trt --variants 0 >> $O/${R}_pair_ab_nt.jsonl 2>&1; step ab-nt-$nt $?
done
for shape in 128,3 128,4 64,6 64,8 32,12 32,16; do
    nt=${shape%,*}; mb=${shape#*,}
    for c in dugks,f64,bgk,4 dugks,f32,bgk,4 fvm,f64,bgk,4; do
        timeout 60 env PLBM_MARCH_NT=$nt PLBM_MARCH_MINB=$mb python tools/kbench.py --n 2048 --steps 50 --case $c >> $O/${R}_kbench_march_shapes.jsonl 2>&1
    done
done; step march-shapes $?
timeout 300 env PLBM_MARCH_NT=64 PLBM_MARCH_MINB=6 python -m pytest tests/test_gpu_fast_variants.py -m gpu -q > $O/${R}_pytest_march_nt64.txt 2>&1; step pytest-march-nt64 $?
timeout 120 env PLBM_BULK_NT=64 ncu --set full --clock-control none --import-source on -k regex:k_lbm2_bulk -c 1 -f -o $O/${R}_k_lbm2_bulk_rr_f32_nt64_8192 \
    python tools/pair_ab.py --cases 8192x8192:f32:rr --variants 0 --once > /dev/null 2>&1; step ncu-rr-f32-nt64 $?
timeout 120 env PLBM_MARCH_NT=64 PLBM_MARCH_MINB=6 ncu --set full --clock-control none --import-source on -k regex:k_fv_march -c 1 -f -o $O/${R}_k_fv_march_dugks_f64_nt64_2048 \
    python tools/kbench.py --n 2048 --steps 2 --case dugks,f64,bgk,4 > /dev/null 2>&1; step ncu-march-nt64 $?
cat $S
