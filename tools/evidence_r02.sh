set -x
P=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $P/r02_pytest_gpu.log 2>&1; tail -3 $P/r02_pytest_gpu.log
python bench.py > $P/r02_bench_1gpu_default.json 2> $P/r02_bench_1gpu_default.err; tail -c 300 $P/r02_bench_1gpu_default.json
python bench.py --impl reference --steps 5 --warmup 1 > $P/r02_bench_reference_arm.json 2>/dev/null
python tools/profile_kernels.py --workload c4 --states 256 --reps 2 > $P/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -o $P/r02_final_c4_256 -f python tools/profile_kernels.py --workload c4 --states 256 --reps 2 > $P/ncu_c4.log 2>&1
python tools/profile_kernels.py --workload c1 --reps 2 > $P/plain_c1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -o $P/r02_final_c1 -f python tools/profile_kernels.py --workload c1 --reps 2 > $P/ncu_c1.log 2>&1
python tools/profile_kernels.py --workload c3 --reps 2 > $P/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rollout_tc|sample_actions|score_reduce|select_elites|refit' -s 5 -c 5 -o $P/r02_final_c3 -f python tools/profile_kernels.py --workload c3 --reps 2 > $P/ncu_c3.log 2>&1
python bench.py --steps 2 --warmup 3 --extras none --no-cpu-baseline > $P/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r02_launches_c1_bench.csv python bench.py --steps 2 --warmup 3 --extras none --no-cpu-baseline > $P/ncu_bench.log 2>&1
for f in ncu_c4 ncu_c1 ncu_c3 ncu_bench; do tail -n 2 $P/$f.log; done
