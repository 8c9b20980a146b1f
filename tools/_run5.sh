mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_round2.py -m gpu -q -x -k "too_large or edge_counts or pair" > gpurun_out/r2b/t3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b/t3.log
for cs in 2; do
echo "== NEW CS=$cs"
DSNT_TUNE_STEP_PAIR_CS=$cs timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32,bf16 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
echo "== NEW CS=$cs trace"
DSNT_B200_LIB=$PWD/tools/probe/libdsnt_trace.so DSNT_TUNE_STEP_PAIR_FLAGS=8 DSNT_TUNE_STEP_PAIR_CS=$cs timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only --iters 3 2>&1 | grep "pair trace" | awk '++n % 6 == 0'
done
