#!/bin/bash
# Tuning sweeps of the one-pass step kernel (developer tool; the measurements behind profiles/r01_v7_step2_sweeps.txt).
#   gpurun -- 'bash tools/sweep_step.sh > gpurun_out/sweep.txt 2>&1'
# Knobs (read once per process by csrc/step.cu):
#   DSNT_TUNE_STEP_V2=0          generic kernel instead of the 64x64 one
#   DSNT_TUNE_STEP_WARPS=n       warps per CTA            DSNT_TUNE_STEP_NWMAX=12|16|20|24   register budget variant
#   DSNT_TUNE_STEP_PACE_GBS=g    pacing target (0 = unpaced; default 6800)   DSNT_TUNE_STEP_PACE=c   pace in SM clocks
#   DSNT_TUNE_STEP_PACED=1       paced build for bf16 JS/MSE too             DSNT_TUNE_STEP_DEBUG=1  copy z -> dz, no arithmetic
#   DSNT_TUNE_STEP_NBUF2=n       cap on the ring buffers  DSNT_TUNE_STEP_STAGGER=ns  staggered start of the warps
K="timeout 200 python tools/kbench.py --configs cfg4 --step-only"
echo "== defaults"; $K --regs none,var,js,mse,kl --dtypes f32,bf16 2>&1 | grep one-pass
echo "== generic kernel"; DSNT_TUNE_STEP_V2=0 $K --regs none,var,js,mse --dtypes f32,bf16 2>&1 | grep one-pass
for g in 0 6400 6600 6800 7000 7200; do echo "== pace_gbs=$g"; DSNT_TUNE_STEP_PACE_GBS=$g $K --regs none,js --dtypes f32,bf16 2>&1 | grep one-pass; done
for w in 4 8 12 16; do echo "== copy through the ring, unpaced, bf16 W=$w"; DSNT_TUNE_STEP_PACE_GBS=0 DSNT_TUNE_STEP_DEBUG=1 DSNT_TUNE_STEP_WARPS=$w $K --regs none --dtypes bf16 2>&1 | grep one-pass; done
echo "== copy through the ring, paced"; DSNT_TUNE_STEP_DEBUG=1 $K --regs none --dtypes f32,bf16 2>&1 | grep one-pass
