#!/bin/bash
# Multi-GPU call of round 2:  gpurun --gpus G --timeout 1500 -- 'bash tools/r02_multi.sh r02m G [tests-k-expr]'
# 1. tests/test_gpu_multi.py (slabs bitwise == single GPU, incl. the k_lbm2_bulk interior + k_lbm2<HALO> boundary mix, DUGKS/FVM halos,
#    ring-global diagnostics and vorticity); 2. bench.py at N = G (and N = 1 on the same box) with the selfcheck hash.
R=${1:-r02m}
G=${2:-2}
K=${3:-}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
nvidia-smi --query-gpu=index,name --format=csv > $O/${R}_gpus.txt 2>&1
nvidia-smi topo -m >> $O/${R}_gpus.txt 2>&1
if [ -n "$K" ]; then
    timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -k "$K" > $O/${R}_pytest_multi.txt 2>&1; step pytest-multi $?
else
    timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > $O/${R}_pytest_multi.txt 2>&1; step pytest-multi $?
fi
timeout 200 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2> $O/${R}_bench_n1.err; step bench-n1 $?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 100 --warmup 5 \
    > $O/${R}_bench_n${G}_c5_bgk_f64_slab.json 2> $O/${R}_bench_n$G.err; step bench-n$G $?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $G --steps 60 --warmup 5 --workload c5_bgk_f64_strong \
    > $O/${R}_bench_n${G}_c5_bgk_f64_strong.json 2> $O/${R}_bench_n${G}_strong.err; step bench-strong-n$G $?
cat $S
