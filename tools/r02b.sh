#!/bin/bash
# Second GPU call of round 2 (one B200): new tests, launcher A/B on small and mid grids, one bench line per BASELINE config.
R=${1:-r02b}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/${R}_status.txt
T0=$(date +%s)
step() { echo "$1 rc=$2 elapsed=$(( $(date +%s) - T0 ))s" >> $S; }
: > $S
timeout 300 python -m pytest tests/test_gpu_ring1.py -x -q -m gpu > $O/${R}_pytest_ring1.txt 2>&1; step ring1 $?
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "checkpoint or set_indices or nan or unfused_dugks or lattice_hash or two_step or slab" > $O/${R}_pytest_new.txt 2>&1; step new-tests $?
timeout 200 python -m pytest tests/test_gpu_zz_round1_late.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/${R}_pytest_late_fullsize.txt 2>&1; step late-fullsize $?
for c in 256x256:f64:bgk 384x384:f64:bgk 512x512:f64:bgk 768x768:f64:bgk 1024x1024:f64:trt 1024x1024:f32:bgk 2048x2048:f64:bgk 4096x4096:f64:bgk 4096x4096:f32:bgk; do
    timeout 60 python tools/pair_ab.py --cases $c --variants 6 --steps 201 >> $O/${R}_pair_ab_crossover.jsonl 2>&1
    timeout 60 env PLBM_PAIR_BULK=2 python tools/pair_ab.py --cases $c --variants 0 --steps 201 >> $O/${R}_pair_ab_crossover.jsonl 2>&1
done; step ab-crossover $?
timeout 100 python tools/pair_ab.py --cases 8192x8192:f64:bgk,8192x8192:f64:rr,8192x8192:f32:bgk,4096x32768:f64:bgk --variants 0 > $O/${R}_pair_ab_big.jsonl 2>&1; step ab-big $?
for w in c5_bgk_f64_slab c4_dugks_f64_2048 c4_dugks_f32_2048 c4_fvm_bardow_f64_2048 c3_rr_f64_8192 c3_rr_f32_8192 c2_trt_f64_1024 c1_bgk_f64_64; do
    timeout 240 python bench.py --workload $w --steps 20 --warmup 5 > $O/${R}_bench_n1_$w.json 2>> $O/${R}_bench.err; step bench-$w $?
done
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > $O/${R}_bench_ref_c5.json 2>> $O/${R}_bench.err; step bench-ref $?
cat $S
