#!/bin/bash
# GPU call (one B200): the warp-specialised three-step kernel (k_lbm3_ws) and the closing dual triple -- parity gate, A/B against
# k_lbmn_bulk with and without the third lattice buffer, ncu of the new kernel, bench lines.
R=${1:-r02q}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${R}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ws_dual.py -m gpu -x -q > $O/${R}_pytest_ws_dual.txt 2>&1; step pytest-ws-dual $?
CASES=4096x32768:f64:bgk,8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:rr,2048x2048:f64:bgk,1024x1024:f64:trt
ab() {  # name, env...
    local name=$1; shift
    timeout 150 env "$@" python tools/pair_ab.py --cases $CASES --variants 0 --steps 20 --reps 4 >> $O/${R}_pair_ab.jsonl 2>&1; step ab-$name $?
}
ab bulk-nospare PLBM_TRIPLE_WS=0 PLBM_SPARE_LATTICE=0
ab ws-nospare PLBM_TRIPLE_WS=1 PLBM_SPARE_LATTICE=0
ab ws-spare PLBM_TRIPLE_WS=1 PLBM_SPARE_LATTICE=1
ab bulk-spare PLBM_TRIPLE_WS=0 PLBM_SPARE_LATTICE=1
ab ws-spare-seg128 PLBM_TRIPLE_WS=1 PLBM_SPARE_LATTICE=1 PLBM_WS_SEGLEN=128
ab ws-spare-seg32 PLBM_TRIPLE_WS=1 PLBM_SPARE_LATTICE=1 PLBM_WS_SEGLEN=32
timeout 150 env PLBM_TRIPLE_WS=1 ncu --set full --clock-control none --import-source on -k regex:k_lbm3_ws -c 1 -o $O/${R}_k_lbm3_ws_bgk_f64_c5 -f \
    python tools/pair_ab.py --cases 4096x32768:f64:bgk --variants 0 --once > $O/${R}_ncu.log 2>&1; step ncu-ws $?
timeout 200 env PLBM_TRIPLE_WS=1 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_ws_spare_k20.json 2> $O/${R}_bench.err; step bench-ws-spare $?
timeout 200 env PLBM_TRIPLE_WS=0 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_bulk_spare_k20.json 2>> $O/${R}_bench.err; step bench-bulk-spare $?
timeout 200 env PLBM_TRIPLE_WS=1 PLBM_SPARE_LATTICE=0 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_ws_nospare_k20.json 2>> $O/${R}_bench.err; step bench-ws-nospare $?
cat $S
