#!/bin/bash
# Final evidence pass of round 2 on ONE B200:  gpurun --timeout 2400 -- 'bash tools/r02_final.sh r02z'
# whole -m gpu suite (serial), smoke, one bench line per BASELINE config + the reference arm, ncu launch list and full capture of
# the headline kernel, compute-sanitizer over every kernel family.
R=${1:-r02z}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 800 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.txt 2>&1; step pytest-gpu $?
timeout 300 env PLBM_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zz_round1_late.py tests/test_gpu_fast_variants.py -m gpu -q > $O/${R}_pytest_opt_in.txt 2>&1; step pytest-opt-in $?
timeout 60 python __graft_entry__.py smoke > $O/${R}_smoke.txt 2>&1; step smoke $?
timeout 300 python bench.py > $O/${R}_bench_n1_c5_bgk_f64_slab.json 2> $O/${R}_bench.err; step bench-c5 $?
timeout 300 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1_c5_bgk_f64_slab_k20.json 2>> $O/${R}_bench.err; step bench-c5-k20 $?
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $O/${R}_bench_ref_c5.json 2>> $O/${R}_bench.err; step bench-ref $?
for w in c1_bgk_f64_64 c2_trt_f64_1024 c3_rr_f64_8192 c3_rr_f32_8192 c4_dugks_f64_2048 c4_dugks_f32_2048 c4_fvm_bardow_f64_2048; do
    timeout 300 python bench.py --workload $w > $O/${R}_bench_n1_$w.json 2>> $O/${R}_bench.err; step bench-$w $?
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_bench_n1_c5_bgk_f64.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > /dev/null 2>&1; step ncu-launch-list $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_bgk_f64_c5 \
    python tools/pair_ab.py --cases 4096x32768:f64:bgk --variants 0 --once > /dev/null 2>&1; step ncu-full-c5 $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_rr_f32_8192 \
    python tools/pair_ab.py --cases 8192x8192:f32:rr --variants 0 --once > /dev/null 2>&1; step ncu-full-rr-f32 $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_lbmn_bulk -c 1 -f -o $O/${R}_k_lbmn_bulk3_trt_f64_1024 \
    python tools/pair_ab.py --cases 1024x1024:f64:trt --variants 0 --once > /dev/null 2>&1; step ncu-full-c2 $?
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize.py > $O/${R}_sanitizer_memcheck.txt 2>&1; step memcheck $?
timeout 900 env PLBM_SANITIZE_VARIANTS=0,4,7,10 compute-sanitizer --tool racecheck python tools/sanitize.py > $O/${R}_sanitizer_racecheck.txt 2>&1; step racecheck $?
cat $S
