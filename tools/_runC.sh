O=gpurun_out/r2e; mkdir -p $O
for rep in 1 2; do
echo "== before (rep $rep)"; DSNT_B200_LIB=$PWD/tools/probe/lib_before.so timeout 120 python tools/kbench.py --configs cfg4 --regs js,mse --dtypes f32,bf16 --step-only 2>&1 | grep -v "^cfg  \|^HBM"
echo "== after: closed-form normalisation (rep $rep)"; timeout 120 python tools/kbench.py --configs cfg4 --regs js,mse --dtypes f32,bf16 --step-only 2>&1 | grep -v "^cfg  \|^HBM"
done
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sub.log
