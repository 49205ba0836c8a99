#!/bin/bash
# GPU call (one B200): where do three steps per pass beat two?  Every collision, both precisions, 1024^2 ... 8192^2, pairs
# (variant 7, k_lbm2_bulk) against triples (variant 10, k_lbmn_bulk: one row / two fp32 rows per thread, 128-thread blocks, copies
# issued by every warp from running line addresses, whole rounds of blocks, packed fp32 collisions).
R=${1:-r02k}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_step_kernel_experimental or three_steps_per_pass" > $O/${R}_pytest_triples.txt 2>&1; step pytest-triples $?
for n in 1024 2048 4096 8192; do
    timeout 300 python tools/pair_ab.py --cases ${n}x${n}:f64:bgk,${n}x${n}:f64:trt,${n}x${n}:f64:rr,${n}x${n}:f32:bgk,${n}x${n}:f32:trt,${n}x${n}:f32:rr --variants 7,10 --steps 61 >> $O/${R}_pair_ab_pairs_vs_triples.jsonl 2>&1; step ab-$n $?
done
timeout 100 python tools/pair_ab.py --cases 1536x1536:f64:bgk,1536x1536:f64:trt,1536x1536:f32:bgk,512x512:f64:trt --variants 6,7,10 --steps 61 >> $O/${R}_pair_ab_pairs_vs_triples.jsonl 2>&1; step ab-1536 $?
timeout 300 python bench.py > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2> $O/${R}_bench.err; step bench-c5 $?
cat $S
