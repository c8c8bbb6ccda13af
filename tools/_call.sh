set -x
tools/gpu_ab.sh r2b head nopit rec64:check default:check rec64 default
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_r2b.log; tail -5 gpurun_out/gputests_r2b.log
python tests/checkers/replay_bench.py > gpurun_out/replay_r2b.json 2> gpurun_out/replay_r2b.err; cat gpurun_out/replay_r2b.json
python tests/checkers/ab_bench.py --sims 2000000 --reps 1 > gpurun_out/ncu_plain_r2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:native_race_kernel -s 2 -c 1 -o gpurun_out/native_r2b -f python tests/checkers/ab_bench.py --sims 2000000 --reps 1 > gpurun_out/ncu_r2b.log 2>&1; tail -3 gpurun_out/ncu_r2b.log
python tests/checkers/replay_bench.py --sims 40000 --reps 1 > gpurun_out/ncu_plain_replay_r2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:replay_race_kernel -s 2 -c 1 -o gpurun_out/replay_r2b -f python tests/checkers/replay_bench.py --sims 40000 --reps 1 > gpurun_out/ncu_replay_r2b.log 2>&1; tail -3 gpurun_out/ncu_replay_r2b.log
