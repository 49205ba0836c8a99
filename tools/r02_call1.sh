#!/bin/bash
# First GPU call of round 2 (one B200):  gpurun --timeout 1800 -- 'bash tools/r02_call1.sh r02a'
# Everything this round's commits added has not yet seen a GPU: whole -m gpu suite first (no -x: list every failure), smoke,
# the headline bench line, then the opt-in kernels (parity gates, A/B), the C4 kernels, one bench line per BASELINE config.
R=${1:-r02a}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${R}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.txt 2>&1; step pytest-gpu $?
timeout 60 python __graft_entry__.py smoke > $O/${R}_smoke.txt 2>&1; step smoke $?
timeout 240 python bench.py > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2> $O/${R}_bench.err; step bench-c5 $?
timeout 120 env PLBM_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "experimental" > $O/${R}_pytest_experimental.txt 2>&1; step experimental-lbmn $?
timeout 300 env PLBM_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zz_round1_late.py tests/test_gpu_fast_variants.py -q -m gpu >> $O/${R}_pytest_experimental.txt 2>&1; step experimental-fma-march $?
# C4 kernels: bit-identical tile kernel (0), its FMA twin (3), marching kernel (4)
for c in dugks,f64,bgk,0 dugks,f64,bgk,3 dugks,f64,bgk,4 dugks,f32,bgk,0 dugks,f32,bgk,3 dugks,f32,bgk,4 fvm,f64,bgk,0 fvm,f64,bgk,3 fvm,f64,bgk,4 fvm,f32,bgk,4; do
    timeout 90 python tools/kbench.py --n 2048 --steps 50 --case $c >> $O/${R}_kbench_c4.jsonl 2>&1; step kbench-$c $?
done
for nt in 128 256; do for mb in 1 2 3 4; do
    timeout 60 env PLBM_MARCH_NT=$nt PLBM_MARCH_MINB=$mb python tools/kbench.py --n 2048 --steps 50 --case dugks,f64,bgk,4 >> $O/${R}_kbench_shapes.jsonl 2>&1
    timeout 60 env PLBM_MARCH_NT=$nt PLBM_MARCH_MINB=$mb python tools/kbench.py --n 2048 --steps 50 --case dugks,f32,bgk,4 >> $O/${R}_kbench_shapes.jsonl 2>&1
done; done; step shapes $?
# LBM: default (0), k_lbm2 (6), k_lbm2_bulk (7), depth-generic two / three steps per pass (9 / 10), FMA twin (11)
timeout 240 python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:rr,4096x32768:f64:bgk --variants 0,7,9,10,11 > $O/${R}_pair_ab.jsonl 2>&1; step ab $?
for c in 256x256:f64:bgk 512x512:f64:bgk 1024x1024:f64:trt 1024x1024:f32:bgk 2048x2048:f64:bgk 4096x4096:f64:bgk; do
    timeout 60 python tools/pair_ab.py --cases $c --variants 0,6 --steps 201 >> $O/${R}_pair_ab_crossover.jsonl 2>&1
    timeout 60 env PLBM_PAIR_BULK=2 python tools/pair_ab.py --cases $c --variants 0 --steps 201 >> $O/${R}_pair_ab_crossover.jsonl 2>&1
done; step ab-crossover $?
for w in c4_dugks_f64_2048 c4_dugks_f32_2048 c4_fvm_bardow_f64_2048 c3_rr_f64_8192 c3_rr_f32_8192 c2_trt_f64_1024 c1_bgk_f64_64; do
    timeout 240 python bench.py --workload $w --steps 20 --warmup 5 > $O/${R}_bench_n1_$w.json 2>> $O/${R}_bench.err; step bench-$w $?
done
timeout 240 python bench.py --workload c4_dugks_f64_2048 --variant 4 --steps 20 --warmup 5 > $O/${R}_bench_n1_c4_dugks_f64_2048_v4.json 2>> $O/${R}_bench.err; step bench-c4-v4 $?
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > $O/${R}_bench_ref_c5.json 2>> $O/${R}_bench.err; step bench-ref $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_fv_march -c 1 -f -o $O/${R}_k_fv_march_dugks_f64_2048 \
    python tools/kbench.py --n 2048 --steps 2 --case dugks,f64,bgk,4 > /dev/null 2>&1; step ncu-march-f64 $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_bgk_f64_8192 \
    python tools/pair_ab.py --cases 8192x8192:f64:bgk --variants 10 --once > /dev/null 2>&1; step ncu-full-lbmn3 $?
cat $S
