#!/bin/bash
# Two-GPU call:  gpurun --gpus 2 --timeout 600 -- 'bash tools/r02_call15.sh r02s'
# The slab schedule with the closing dual triple (third lattice buffer on every rank), k_lbm3_ws interiors + k_lbmn_bulk<HALO, DUAL>
# boundaries: bitwise tests of what changed, then the bench at N = 1 and N = 2 with the driver's --steps 20 --warmup 5.
R=${1:-r02s}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
nvidia-smi --query-gpu=index,name --format=csv > $O/${R}_gpus.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -k "closing_dual or (three_steps_per_pass and p2p) or (bulk_interior and f64)" > $O/${R}_pytest_multi.txt 2>&1; step pytest-multi $?
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2> $O/${R}_bench_n1.err; step bench-n1 $?
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu \
    > $O/${R}_bench_n2_c5_bgk_f64_slab.json 2> $O/${R}_bench_n2.err; step bench-n2 $?
cat $S
