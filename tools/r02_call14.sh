#!/bin/bash
# GPU call (one B200): shape of k_lbm3_ws (segment length, consumer threads per block), e2e with / without the third lattice buffer
# (historical: PLBM_WS_NTC selected 96 / 192 consumer threads per block; both were slower and the instantiations were removed afterwards)
# (alternating processes, median of three cycles), compute-sanitizer over the new kernel and the closing dual triple.
R=${1:-r02r}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
CASES=4096x32768:f64:bgk,8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:rr,2048x2048:f64:bgk,1024x1024:f64:trt
ab() {  # name, env...
    local name=$1; shift
    timeout 150 env PLBM_TRIPLE_WS=1 PLBM_SPARE_LATTICE=1 "$@" python tools/pair_ab.py --cases $CASES --variants 0 --steps 20 --reps 3 >> $O/${R}_pair_ab.jsonl 2>&1; step ab-$name $?
}
ab seg96 PLBM_WS_SEGLEN=96
ab seg128 PLBM_WS_SEGLEN=128
ab seg192 PLBM_WS_SEGLEN=192
ab seg256 PLBM_WS_SEGLEN=256
ab ntc96-seg128 PLBM_WS_NTC=96 PLBM_WS_SEGLEN=128
ab ntc192-seg128 PLBM_WS_NTC=192 PLBM_WS_SEGLEN=128
ab ntc96-seg64 PLBM_WS_NTC=96 PLBM_WS_SEGLEN=64
SAN="PLBM_TRIPLES=2 PLBM_SPARE_LATTICE=2 PLBM_TRIPLE_WS=1 PLBM_SANITIZE_VARIANTS=0,10"
timeout 300 env $SAN compute-sanitizer --tool memcheck python tools/sanitize.py > $O/${R}_sanitizer_memcheck.txt 2>&1; step memcheck $?
timeout 400 env $SAN compute-sanitizer --tool racecheck python tools/sanitize.py > $O/${R}_sanitizer_racecheck.txt 2>&1; step racecheck $?
timeout 300 env $SAN compute-sanitizer --tool initcheck python tools/sanitize.py > $O/${R}_sanitizer_initcheck.txt 2>&1; step initcheck $?
for i in 1 2; do
    timeout 200 env PLBM_TRIPLE_WS=1 PLBM_WS_SEGLEN=128 PLBM_SPARE_LATTICE=1 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_spare_$i.json 2>> $O/${R}_bench.err; step bench-spare-$i $?
    timeout 200 env PLBM_TRIPLE_WS=1 PLBM_WS_SEGLEN=128 PLBM_SPARE_LATTICE=0 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_nospare_$i.json 2>> $O/${R}_bench.err; step bench-nospare-$i $?
done
cat $S
