mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_round2.py -m gpu -q -x -k "too_large or edge_counts or pair" > gpurun_out/r2b/t5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b/t5.log
echo "== default (STAGE0 = 3, closed-form normalisation) CS=2 f32, bf16"
timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32,bf16 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
echo "== CS=4 bf16"
DSNT_TUNE_STEP_PAIR_CS=4 timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes bf16 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
for v in S1 S2 S4; do
echo "== variant $v CS=2 f32"
DSNT_B200_LIB=$PWD/tools/probe/lib_$v.so timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
done
echo "== two-kernel bf16 cfg5 for comparison"
timeout 300 python tools/kbench.py --configs cfg5 --regs js,var --dtypes bf16 2>&1 | grep -v "^HBM\|^cfg  "
