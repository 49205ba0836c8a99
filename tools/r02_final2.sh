#!/bin/bash
# Second part of the final evidence pass (one B200): what the first part got wrong -- bench warm-up now loads every kernel (the
# driver's --steps 20 --warmup 5), C4 at a stable time step (CFL 0.5), compute-sanitizer after the variant-10 fall-through fix.
R=${1:-r02y}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_step_kernel_experimental or three_steps_per_pass" > $O/${R}_pytest_triples.txt 2>&1; step pytest-triples $?
timeout 300 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1_c5_bgk_f64_slab_k20.json 2> $O/${R}_bench.err; step bench-c5-k20 $?
timeout 300 python bench.py > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2>> $O/${R}_bench.err; step bench-c5 $?
for w in c4_dugks_f64_2048 c4_dugks_f32_2048 c4_fvm_bardow_f64_2048 c2_trt_f64_1024; do
    timeout 300 python bench.py --workload $w > $O/${R}_bench_n1_$w.json 2>> $O/${R}_bench.err; step bench-$w $?
done
for c in dugks,f64,bgk,0 dugks,f64,bgk,3 dugks,f64,bgk,4 dugks,f32,bgk,0 dugks,f32,bgk,4 fvm,f64,bgk,0 fvm,f64,bgk,3; do
    timeout 90 python tools/kbench.py --n 2048 --steps 50 --case $c >> $O/${R}_kbench_c4_stable_dt.jsonl 2>&1
done; step kbench-c4 $?
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize.py > $O/${R}_sanitizer_memcheck.txt 2>&1; step memcheck $?
timeout 900 env PLBM_SANITIZE_VARIANTS=0,4,7,10 compute-sanitizer --tool racecheck python tools/sanitize.py > $O/${R}_sanitizer_racecheck.txt 2>&1; step racecheck $?
timeout 600 env PLBM_SANITIZE_VARIANTS=0,4,10 compute-sanitizer --tool initcheck python tools/sanitize.py > $O/${R}_sanitizer_initcheck.txt 2>&1; step initcheck $?
cat $S
