set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_r2j.log; tail -4 gpurun_out/gputests_r2j.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2j.log 2>&1; tail -2 gpurun_out/smoke_r2j.log
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1 -f"
python tests/checkers/ab_bench.py --sims 2000000 --reps 1 > gpurun_out/ncu_plain_r2j.log 2>&1 && $NCU -k regex:native_race_kernel -o gpurun_out/native_r2j python tests/checkers/ab_bench.py --sims 2000000 --reps 1 > gpurun_out/ncu_r2j.log 2>&1; tail -2 gpurun_out/ncu_r2j.log
python tests/checkers/ab_bench.py --sims 1000000 --reps 1 --mode trace >> gpurun_out/ncu_plain_r2j.log 2>&1 && $NCU -k regex:native_race_kernel -o gpurun_out/native_trace_r2j python tests/checkers/ab_bench.py --sims 1000000 --reps 1 --mode trace > gpurun_out/ncu_trace_r2j.log 2>&1; tail -2 gpurun_out/ncu_trace_r2j.log
python tests/checkers/ab_bench.py --sims 2000000 --reps 1 --mode laphist >> gpurun_out/ncu_plain_r2j.log 2>&1 && $NCU -k regex:native_race_kernel -o gpurun_out/native_laphist_r2j python tests/checkers/ab_bench.py --sims 2000000 --reps 1 --mode laphist > gpurun_out/ncu_laphist_r2j.log 2>&1; tail -2 gpurun_out/ncu_laphist_r2j.log
python tests/checkers/replay_bench.py --sims 40000 --reps 1 >> gpurun_out/ncu_plain_r2j.log 2>&1 && $NCU -k regex:replay_race_kernel -o gpurun_out/replay_r2j python tests/checkers/replay_bench.py --sims 40000 --reps 1 > gpurun_out/ncu_replay_r2j.log 2>&1; tail -2 gpurun_out/ncu_replay_r2j.log
