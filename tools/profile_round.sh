#!/bin/bash
# Evidence pass of a round, run on the GPU box:  gpurun --timeout 1500 -- 'bash tools/profile_round.sh r01'
# Writes bench lines, the ncu launch list and the ncu --set full captures of the dominant kernel to gpurun_out/.
R=${1:-rXX}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py > $O/${R}_bench_n1_c5_bgk_f64.json 2> $O/${R}_bench.err
for w in c3_rr_f64_8192 c3_rr_f32_8192 c2_trt_f64_1024 c1_bgk_f64_64; do
    timeout 200 python bench.py --workload $w --no-cpu > $O/${R}_bench_n1_$w.json 2>> $O/${R}_bench.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_bench_n1_c5_bgk_f64.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lbm2 -s 1 -c 1 -o $O/${R}_k_lbm2_c5_bgk_f64_slab \
    python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
for w in c3_rr_f64_8192 c3_rr_f32_8192 c2_trt_f64_1024; do
    timeout 300 ncu --set full --clock-control none -k regex:k_lbm2 -s 1 -c 1 -o $O/${R}_k_lbm2_$w \
        python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
done
cut -c1-300 $O/${R}_bench_n1_*.json
