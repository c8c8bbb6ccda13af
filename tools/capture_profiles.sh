#!/bin/bash
# Round-end evidence pass on ONE B200 (run under gpurun from the repo root): every profiled command first exits 0
# without ncu; ncu captures the THIRD launch of each kernel (parameters resident, no L2 fill before it).
#   tools/capture_profiles.sh <tag>      -> gpurun_out/<tag>_*.ncu-rep, <tag>_bench_line.json, <tag>_bench_launch_list.csv, ...
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1 -f"
run() { "$@" > $out/_plain.log 2>&1 || { echo "FAILED without ncu: $*"; tail -3 $out/_plain.log; return 1; }; }
run python tests/checkers/ab_bench.py --sims 2000000 --reps 1 && $NCU -k regex:native_race_kernel -o $out/${tag}_native python tests/checkers/ab_bench.py --sims 2000000 --reps 1 > $out/${tag}_native.log 2>&1
run python tests/checkers/ab_bench.py --sims 1000000 --reps 1 --mode trace && $NCU -k regex:native_race_kernel -o $out/${tag}_native_trace python tests/checkers/ab_bench.py --sims 1000000 --reps 1 --mode trace > $out/${tag}_native_trace.log 2>&1
run python tests/checkers/ab_bench.py --sims 2000000 --reps 1 --mode laphist && $NCU -k regex:native_race_kernel -o $out/${tag}_native_laphist python tests/checkers/ab_bench.py --sims 2000000 --reps 1 --mode laphist > $out/${tag}_native_laphist.log 2>&1
run python tests/checkers/replay_bench.py --sims 40000 --reps 1 && $NCU -k regex:replay_race_kernel -o $out/${tag}_replay python tests/checkers/replay_bench.py --sims 40000 --reps 1 > $out/${tag}_replay.log 2>&1
python bench.py > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err && tail -c 400 $out/${tag}_bench_line.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
ls -la $out | tail -20
