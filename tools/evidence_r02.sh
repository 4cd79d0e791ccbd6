set -x
P=gpurun_out
python tools/profile_kernels.py --workload c4 --states 256 --reps 2 > $P/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -o $P/r02_final_c4_256 -f python tools/profile_kernels.py --workload c4 --states 256 --reps 2 > $P/ncu_c4.log 2>&1
python tools/profile_kernels.py --workload c1 --reps 2 > $P/plain_c1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -o $P/r02_final_c1 -f python tools/profile_kernels.py --workload c1 --reps 2 > $P/ncu_c1.log 2>&1
python tools/profile_kernels.py --workload c5 --states 20 --reps 2 > $P/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_tc_wide -s 1 -c 1 -o $P/r02_final_c5_20 -f python tools/profile_kernels.py --workload c5 --states 20 --reps 2 > $P/ncu_c5.log 2>&1
python tools/profile_kernels.py --workload c3 --reps 2 > $P/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rollout_tc|sample_actions|score_reduce|select_elites|refit' -s 5 -c 5 -o $P/r02_final_c3 -f python tools/profile_kernels.py --workload c3 --reps 2 > $P/ncu_c3.log 2>&1
python bench.py --steps 2 --warmup 3 --extras none --no-cpu-baseline > $P/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r02_launches_c1_bench.csv python bench.py --steps 2 --warmup 3 --extras none --no-cpu-baseline > $P/ncu_bench.log 2>&1
tail -2 $P/ncu_c4.log $P/ncu_c1.log $P/ncu_c5.log $P/ncu_c3.log $P/ncu_bench.log
ls -la $P/*.ncu-rep | tail
