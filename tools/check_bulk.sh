#!/bin/bash
# Parity + A/B + ncu evidence of the bulk-copy flavour of the two-step kernel, run on the GPU box:
#   gpurun --timeout 340 -- 'bash tools/check_bulk.sh r01b'
# Ordered by importance (the box time left may cut the tail); everything goes to gpurun_out/, torch-free until the last step.
R=${1:-rXX}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
date +%s > $S
# 1. parity of both flavours (variants 5 default, 6 per-thread loads, 7 bulk, 8 bulk over the slab schedule's x ranges)
timeout 150 python -m pytest tests/test_gpu_parity.py -k two_step -x -q -m gpu > $O/${R}_pytest_two_step.txt 2>&1
echo "two_step rc=$? t=$(date +%s)" >> $S
# 2. A/B at 8192^2 (host clock around one 41-step call, best of 3)
timeout 120 python tools/pair_ab.py > $O/${R}_pair_ab_8192.jsonl 2>&1
echo "ab8192 rc=$? t=$(date +%s)" >> $S
# 3. full-size tiling parity with the bulk flavour as the default (C2 1024^2 TRT, C3 8192^2 RR fp64/fp32)
PLBM_PAIR_BULK=1 timeout 150 python -m pytest tests/test_gpu_fullsize.py -k "C2 or C3" -x -q -m gpu > $O/${R}_pytest_fullsize_bulk.txt 2>&1
echo "fullsize rc=$? t=$(date +%s)" >> $S
# 4. the bench shape (32768 rows x 4096 lines per GPU)
timeout 100 python tools/pair_ab.py --cases 4096x32768:f64:bgk --steps 21 --reps 2 > $O/${R}_pair_ab_c5.jsonl 2>&1
echo "abc5 rc=$? t=$(date +%s)" >> $S
# 5. ncu: launch list and one full capture of k_lbm2_bulk (8192^2 BGK fp64, one 3-step call)
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_pair_ab_8192_bgk_f64.csv \
    python tools/pair_ab.py --cases 8192x8192:f64:bgk --variants 7 --once > /dev/null 2>&1
echo "ncu-list rc=$? t=$(date +%s)" >> $S
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbm2_bulk -c 1 -f -o $O/${R}_k_lbm2_bulk_bgk_f64_8192 \
    python tools/pair_ab.py --cases 8192x8192:f64:bgk --variants 7 --once > /dev/null 2>&1
echo "ncu-full rc=$? t=$(date +%s)" >> $S
# 6. the bench line with the bulk flavour as the default
PLBM_PAIR_BULK=1 timeout 240 python bench.py --no-cpu > $O/${R}_bench_n1_c5_bgk_f64_bulk.json 2> $O/${R}_bench.err
echo "bench rc=$? t=$(date +%s)" >> $S
cat $S
