set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_r2d.log; tail -30 gpurun_out/gputests_r2d.log
python tests/checkers/replay_bench.py > gpurun_out/replay_r2d.json 2> gpurun_out/replay_r2d.err; cat gpurun_out/replay_r2d.json
timeout 900 python bench.py --steps 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 3000 gpurun_out/bench_r2d.json; tail -5 gpurun_out/bench_r2d.err
