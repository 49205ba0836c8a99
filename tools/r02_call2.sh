#!/bin/bash
# Second GPU call of round 2 (one B200):  gpurun --timeout 900 -- 'bash tools/r02_call2.sh r02b'
# Packed fp32 collisions (FFMA2 pairs, plbm_f32x2.cuh) are now the fp32 default of the two-step kernels: bit parity of the whole
# suite, A/B against the scalar form (variant 12), ncu of the RR fp32 kernel; where three steps per pass (variant 10) pays.
R=${1:-r02b}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 600 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.txt 2>&1; step pytest-gpu $?
timeout 60 python __graft_entry__.py smoke > $O/${R}_smoke.txt 2>&1; step smoke $?
timeout 200 python tools/pair_ab.py --cases 8192x8192:f32:bgk,8192x8192:f32:trt,8192x8192:f32:rr,4096x4096:f32:bgk,4096x4096:f32:rr,2048x2048:f32:rr --variants 0,12 > $O/${R}_pair_ab_packed.jsonl 2>&1; step ab-packed $?
timeout 100 env PLBM_PAIR_BULK=0 python tools/pair_ab.py --cases 8192x8192:f32:bgk,8192x8192:f32:rr --variants 0,12 > $O/${R}_pair_ab_packed_k_lbm2.jsonl 2>&1; step ab-packed-k_lbm2 $?
timeout 200 python tools/pair_ab.py --cases 1024x1024:f64:trt,2048x2048:f64:trt,4096x4096:f64:trt,8192x8192:f64:trt,4096x4096:f32:trt,8192x8192:f32:trt --variants 0,10 --steps 61 > $O/${R}_pair_ab_triples_trt.jsonl 2>&1; step ab-triples $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbm2_bulk -c 1 -f -o $O/${R}_k_lbm2_bulk_rr_f32_packed_8192 \
    python tools/pair_ab.py --cases 8192x8192:f32:rr --variants 0 --once > /dev/null 2>&1; step ncu-rr-f32-packed $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbm2_bulk -c 1 -f -o $O/${R}_k_lbm2_bulk_rr_f32_scalar_8192 \
    python tools/pair_ab.py --cases 8192x8192:f32:rr --variants 12 --once > /dev/null 2>&1; step ncu-rr-f32-scalar $?
for w in c3_rr_f32_8192; do
    timeout 240 python bench.py --workload $w --steps 200 --warmup 5 > $O/${R}_bench_n1_$w.json 2>> $O/${R}_bench.err; step bench-$w $?
done
cat $S
