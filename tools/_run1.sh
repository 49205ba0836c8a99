set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python bench.py > $O/r01b_bench_n1_c5_bgk_f64.json 2> $O/r01b_bench_n1_c5.err
for w in c3_rr_f64_8192 c3_rr_f32_8192 c2_trt_f64_1024 c1_bgk_f64_64; do timeout 200 python bench.py --workload $w --no-cpu > $O/r01b_bench_n1_$w.json 2>> $O/r01b_bench.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01b_launches_bench_n1_c5_bgk_f64.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lbm2 -s 1 -c 1 -o $O/r01b_k_lbm2_bgk_f64_c5 python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
for w in c3_rr_f64_8192 c3_rr_f32_8192 c2_trt_f64_1024; do timeout 300 ncu --set full --clock-control none -k regex:k_lbm2 -s 1 -c 1 -o $O/r01b_k_lbm2_$w python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1; done
cat $O/r01b_bench_n1_*.json | cut -c1-400
tail -3 $O/r01b_bench.err $O/r01b_bench_n1_c5.err
