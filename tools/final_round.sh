#!/bin/bash
# Final evidence pass of a round on ONE B200 within a small time budget:
#   gpurun --timeout 280 -- 'bash tools/final_round.sh r01c 250'
# full GPU suite (pytest-xdist, perf guard and the 32768^2 case serial), smoke, bench line, ncu launch list + full capture.
R=${1:-rXX}
BUDGET=${2:-250}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
left() { echo $(( BUDGET - ( $(date +%s) - T0 ) )); }
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
python -c "from oracle import oracle; oracle.build()" ; step oracle-build $?
timeout 130 python -m pytest tests -m gpu -q -n 6 --deselect tests/test_gpu_perf.py -k "not C5" > $O/${R}_pytest_gpu_xdist.txt 2>&1; step pytest-xdist $?
timeout 40 python __graft_entry__.py smoke > $O/${R}_smoke.txt 2>&1; step smoke $?
[ $(left) -gt 60 ] && { timeout 90 python -m pytest tests/test_gpu_perf.py -q -m gpu > $O/${R}_pytest_gpu_perf.txt 2>&1; step pytest-perf $?; }
[ $(left) -gt 50 ] && { timeout 120 python bench.py > $O/${R}_bench_n1_c5_bgk_f64.json 2> $O/${R}_bench.err; step bench $?; }
[ $(left) -gt 40 ] && { timeout $(left) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_bench_n1_c5_bgk_f64.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1; step ncu-launch-list $?; }
[ $(left) -gt 30 ] && { timeout $(left) ncu --set full --clock-control none --import-source on -k regex:k_lbm2_bulk -c 1 -f -o $O/${R}_k_lbm2_bulk_bgk_f64_c5 \
    python tools/pair_ab.py --cases 4096x32768:f64:bgk --variants 0 --once > /dev/null 2>&1; step ncu-full-c5 $?; }
[ $(left) -gt 45 ] && { timeout $(left) python -m pytest tests/test_gpu_fullsize.py -k C5 -q -m gpu > $O/${R}_pytest_gpu_c5.txt 2>&1; step pytest-c5 $?; }
cat $S
