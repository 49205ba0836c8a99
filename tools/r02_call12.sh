#!/bin/bash
# Short GPU call (one B200): segment length of the three-step kernel (four warm-up columns per segment: 6 % redundant at 64 columns).
R=${1:-r02p}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
for sl in 64 96 128 192 256 512; do
    timeout 100 env PLBM_MULTI_SEGLEN=$sl python tools/pair_ab.py --cases 4096x32768:f64:bgk,8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f32:bgk,4096x4096:f64:bgk,2048x2048:f64:bgk --variants 0 --steps 61 >> $O/${R}_pair_ab_seglen.jsonl 2>&1; step seglen-$sl $?
done
cat $S
