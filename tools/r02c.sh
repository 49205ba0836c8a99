#!/bin/bash
# Third GPU call of round 2 (one B200): the marching DUGKS kernel (variant 4) -- tolerance gate, block-shape sweep, ncu.
R=${1:-r02c}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 600 python -m pytest tests/test_gpu_fast_variants.py -q -m gpu > $O/${R}_pytest_fast.txt 2>&1; step fast-variants $?
timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nan" > $O/${R}_pytest_nan.txt 2>&1; step nan $?
for c in dugks,f64,bgk,0 dugks,f64,bgk,4 dugks,f32,bgk,0 dugks,f32,bgk,4 fvm,f64,bgk,0 fvm,f64,bgk,4 fvm,f32,bgk,4; do
    timeout 90 python tools/kbench.py --n 2048 --steps 50 --case $c >> $O/${R}_kbench.jsonl 2>&1; step kbench-$c $?
done
for nt in 128 256; do for mb in 1 2 3 4; do
    timeout 60 env PLBM_MARCH_NT=$nt PLBM_MARCH_MINB=$mb python tools/kbench.py --n 2048 --steps 50 --case dugks,f64,bgk,4 >> $O/${R}_kbench_shapes.jsonl 2>&1
    timeout 60 env PLBM_MARCH_NT=$nt PLBM_MARCH_MINB=$mb python tools/kbench.py --n 2048 --steps 50 --case dugks,f32,bgk,4 >> $O/${R}_kbench_shapes.jsonl 2>&1
done; done; step shapes $?
for n in 4096 8192; do
    timeout 60 python tools/kbench.py --n $n --steps 20 --case dugks,f64,bgk,4 >> $O/${R}_kbench_big.jsonl 2>&1
    timeout 60 python tools/kbench.py --n $n --steps 20 --case dugks,f64,bgk,0 >> $O/${R}_kbench_big.jsonl 2>&1
done; step big $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_fv_march -c 1 -f -o $O/${R}_k_fv_march_dugks_f64_2048 \
    python tools/kbench.py --n 2048 --steps 2 --case dugks,f64,bgk,4 > /dev/null 2>&1; step ncu-march-f64 $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_fv_march -c 1 -f -o $O/${R}_k_fv_march_dugks_f32_2048 \
    python tools/kbench.py --n 2048 --steps 2 --case dugks,f32,bgk,4 > /dev/null 2>&1; step ncu-march-f32 $?
for c in 1024x1024:f64:trt 1024x1024:f64:bgk 1536x1536:f64:bgk 2048x2048:f64:bgk 2048x2048:f32:bgk; do
    timeout 60 python tools/pair_ab.py --cases $c --variants 0 --steps 201 >> $O/${R}_pair_ab_mid.jsonl 2>&1
done; step ab-mid $?
timeout 240 python bench.py --workload c2_trt_f64_1024 --steps 20 --warmup 5 > $O/${R}_bench_n1_c2_trt_f64_1024.json 2>> $O/${R}_bench.err; step bench-c2 $?
timeout 240 python bench.py --workload c4_dugks_f64_2048 --variant 4 --steps 20 --warmup 5 > $O/${R}_bench_n1_c4_dugks_f64_2048_v4.json 2>> $O/${R}_bench.err; step bench-c4-v4 $?
cat $S
