#!/bin/bash
# Last GPU call of round 2 (one B200, ~4 GPU-minutes left): the final build -- smoke, the gates of the two three-step kernels and of the
# closing dual triple, the tests of test_gpu_parity.py that touch them, the bench line the driver asks for, ncu of k_lbm3_ws, racecheck.
R=${1:-r02t}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.txt 2>&1; step smoke $?
timeout 80 python -m pytest tests/test_gpu_ws_dual.py -m gpu -x -q > $O/${R}_pytest_ws_dual.txt 2>&1; step pytest-ws-dual $?
timeout 60 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1_c5_bgk_f64_slab_k20.json 2> $O/${R}_bench.err; step bench-c5-k20 $?
timeout 45 ncu --set full --clock-control none --import-source on -k regex:k_lbm3_ws -c 1 -o $O/${R}_k_lbm3_ws_bgk_f64_c5 -f \
    python tools/pair_ab.py --cases 4096x32768:f64:bgk --variants 0 --once > $O/${R}_ncu.log 2>&1; step ncu-ws $?
timeout 70 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "three_steps_per_pass or multi_step_kernel_experimental or fused_lbm_steps" > $O/${R}_pytest_parity_subset.txt 2>&1; step pytest-parity-subset $?
timeout 60 env PLBM_TRIPLES=2 PLBM_SPARE_LATTICE=2 PLBM_SANITIZE_VARIANTS=0 compute-sanitizer --tool racecheck python tools/sanitize.py > $O/${R}_sanitizer_racecheck.txt 2>&1; step racecheck $?
cat $S
