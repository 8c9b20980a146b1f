mkdir -p gpurun_out/r2b
./tools/probe/cluster_probe > gpurun_out/r2b/cluster_probe.txt 2>&1; cat gpurun_out/r2b/cluster_probe.txt
for fl in 0 1 2 4 7; do
echo "== CS=2 flags $fl"
DSNT_TUNE_STEP_PAIR_VERBOSE=1 DSNT_TUNE_STEP_PAIR_CS=2 DSNT_TUNE_STEP_PAIR_FLAGS=$fl timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var --dtypes f32 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
done
echo "== CS=4 flags 0"
DSNT_TUNE_STEP_PAIR_VERBOSE=1 DSNT_TUNE_STEP_PAIR_CS=4 timeout 300 python tools/kbench.py --configs cfg5 --regs none,js --dtypes f32 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
