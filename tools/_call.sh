set -x
tools/gpu_ab.sh r2a head r10 nopit:check default:check strict:check head default
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_r2a.log; tail -5 gpurun_out/gputests_r2a.log
python tools/ab_bench.py --sims 2000000 --reps 1 > gpurun_out/ncu_plain_r2a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:native_race_kernel -s 2 -c 1 -o gpurun_out/native_r2a -f python tools/ab_bench.py --sims 2000000 --reps 1 > gpurun_out/ncu_r2a.log 2>&1; tail -3 gpurun_out/ncu_r2a.log
