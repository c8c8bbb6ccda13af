#!/bin/bash
# Turn the captures of tools/capture_profiles.sh (gpurun_out/<tag>_*.ncu-rep, bench lines) into the files under profiles/.
#   tools/process_captures.sh <tag>      (run on the build box after the gpurun call returned)
set -e
tag=${1:-r2}
g=gpurun_out
python tools/ncu_summary.py $g/${tag}_native.ncu-rep ${tag}_native 2000000 --kernel native --note "headline kernel native_race_kernel<5,0,0,32>: python tests/checkers/ab_bench.py --sims 2000000 --reps 1 (third launch captured; no L2 fill before it)" > /dev/null
python tools/ncu_summary.py $g/${tag}_native_trace.ncu-rep ${tag}_native_trace 1000000 --kernel native --note "trace variant <5,0,2,32>: ab_bench.py --sims 1000000 --reps 1 --mode trace" > /dev/null
python tools/ncu_summary.py $g/${tag}_native_laphist.ncu-rep ${tag}_native_laphist 2000000 --kernel native --note "lap-histogram variant <5,0,3,32>: ab_bench.py --sims 2000000 --reps 1 --mode laphist" > /dev/null
python tools/ncu_summary.py $g/${tag}_replay.ncu-rep ${tag}_replay 40000 --kernel replay --note "replay_race_kernel<10>: tests/checkers/replay_bench.py --sims 40000 --reps 1 (synthetic worst-case-sized tapes)" > /dev/null
tmp=$(mktemp -d)
(cd $tmp && cuobjdump -xelf all $OLDPWD/monte-carlo-gp_b200/csrc/native_kernel.o > /dev/null && cuobjdump -xelf all $OLDPWD/monte-carlo-gp_b200/csrc/replay_kernel.o > /dev/null)
python tools/sass_by_line.py $g/${tag}_native.ncu-rep $tmp/native_kernel.sm_100a.cubin native_race_kernelILi5ELb0ELi0ELi32 2000000 > profiles/${tag}_native_kernel_by_source_line.txt
python tools/sass_by_line.py $g/${tag}_replay.ncu-rep $tmp/replay_kernel.sm_100a.cubin replay_race_kernelILi10 40000 > profiles/${tag}_replay_kernel_by_source_line.txt
rm -rf $tmp
cp $g/${tag}_bench_reference_arm.json $g/${tag}_bench_launch_list.csv profiles/
echo "(the bench line is copied separately: it must be taken AFTER the summaries are committed, so that capture_matches_build is evaluated against them)"
for t in native native_trace native_laphist replay; do python - "$tag" "$t" <<'PY'
import json, sys
d = json.load(open(f"profiles/{sys.argv[1]}_{sys.argv[2]}_ncu_summary.json"))
print(sys.argv[2], round(d["executed_warp_instr_per_unit"], 1), "instr/unit", d["duration_ms"], "ms issue", round(d["issue_active_pct"], 1),
      "alu", round(d["pipe_alu_pct"], 1), "fma", round(d["pipe_fma_pct"], 1), "lsu-data", round(d["lsu_data_pipe_wavefronts_pct"], 1), "regs", d["registers_per_thread"],
      "dram r/w", d["dram_bytes_read"], d["dram_bytes_written"], d["stall_cycles_per_issue"])
PY
done
