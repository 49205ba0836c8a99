#!/bin/bash
# Fifth GPU call of round 2 (one B200): the sum-form marching kernel k_fv_march_s (PLBM_MARCH_FORM=1): tolerance gate, shapes, ncu.
R=${1:-r02e}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 300 env PLBM_MARCH_FORM=1 python -m pytest tests/test_gpu_fast_variants.py -m gpu -q > $O/${R}_pytest_march_s.txt 2>&1; step pytest-march-s $?
timeout 300 env PLBM_MARCH_FORM=1 PLBM_MARCH_NT=64 python -m pytest tests/test_gpu_fast_variants.py -m gpu -q > $O/${R}_pytest_march_s_nt64.txt 2>&1; step pytest-march-s-nt64 $?
for shape in 0,128,3 1,128,3 1,128,2 1,64,5; do
    IFS=, read form nt mb <<< "$shape"
    for c in dugks,f64,bgk,4 dugks,f32,bgk,4 fvm,f64,bgk,4 fvm,f32,bgk,4; do
        timeout 60 env PLBM_MARCH_FORM=$form PLBM_MARCH_NT=$nt PLBM_MARCH_MINB=$mb python tools/kbench.py --n 2048 --steps 50 --case $c >> $O/${R}_kbench_march_s.jsonl 2>&1
    done
done; step march-s-shapes $?
for n in 4096 8192; do
    for form in 0 1; do
        timeout 60 env PLBM_MARCH_FORM=$form python tools/kbench.py --n $n --steps 20 --case dugks,f64,bgk,4 >> $O/${R}_kbench_march_s_big.jsonl 2>&1
    done
done; step march-s-big $?
timeout 120 env PLBM_MARCH_FORM=1 ncu --set full --clock-control none --import-source on -k regex:k_fv_march -c 1 -f -o $O/${R}_k_fv_march_s_dugks_f64_2048 \
    python tools/kbench.py --n 2048 --steps 2 --case dugks,f64,bgk,4 > /dev/null 2>&1; step ncu-march-s $?
cat $S
