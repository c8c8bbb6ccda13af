set -x
timeout 900 python -m pytest tests/test_gpu_replay.py -x -q > gpurun_out/gputests_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_r2c.log; tail -5 gpurun_out/gputests_r2c.log
python tests/checkers/replay_bench.py > gpurun_out/replay_r2c.json 2> gpurun_out/replay_r2c.err; cat gpurun_out/replay_r2c.json
python tests/checkers/replay_bench.py --sims 40000 --reps 1 > gpurun_out/ncu_plain_replay_r2c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:replay_race_kernel -s 2 -c 1 -o gpurun_out/replay_r2c -f python tests/checkers/replay_bench.py --sims 40000 --reps 1 > gpurun_out/ncu_replay_r2c.log 2>&1; tail -3 gpurun_out/ncu_replay_r2c.log
