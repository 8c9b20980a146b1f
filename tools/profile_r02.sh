#!/usr/bin/env bash
# ncu evidence for round 2 (run on the GPU box, one GPU):  bash tools/profile_r02.sh gpurun_out/r2/prof
#   launches.csv           every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
#   prof_step_f32_js       ncu --set full of head_step2_kernel, fp32 JS: two launches of the bare form (dsnt_head_step) and two
#                          of the single-launch form (dsnt_head_step_fused)
#   prof_step_bf16_js      the same, bf16 JS (20 warps, two window slots)
#   prof_pair_var          head_step_pair_kernel, 256x256 fp32 variance (cfg 5)
# Summaries are made off-line with tools/ncu_summary.py.  Numbers printed by a run under ncu are never bench values.
set -uo pipefail
OUT="${1:-gpurun_out/r2/prof}"
mkdir -p "$OUT"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > "$OUT/launches_bench.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_step2 -s 4 -c 4 -o "$OUT/prof_step_f32_js" -f \
    python tools/kbench.py --configs cfg4 --regs js --dtypes f32 --step-only --iters 3 > "$OUT/prof_step_f32_js.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_step2 -s 4 -c 4 -o "$OUT/prof_step_bf16_js" -f \
    python tools/kbench.py --configs cfg4 --regs js --dtypes bf16 --step-only --iters 3 > "$OUT/prof_step_bf16_js.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_step_pair -s 3 -c 1 -o "$OUT/prof_pair_var" -f \
    python tools/kbench.py --configs cfg5 --regs var --dtypes f32 --step-only --iters 3 > "$OUT/prof_pair_var.log" 2>&1
ls -la "$OUT"
