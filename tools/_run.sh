mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py tests/test_gpu_level1.py -q -x > gpurun_out/v12_pytest_gpu.log 2>&1; tail -5 gpurun_out/v12_pytest_gpu.log
for k in 1 2 3 4; do echo "== L2 ctas/SM $k"; DSNT_TUNE_STEP_L2_CTAS=$k timeout 200 python tools/kbench.py --configs cfg5 --regs var,none --dtypes f32 2>&1 | grep cfg5; done
echo "== bf16"; timeout 200 python tools/kbench.py --configs cfg5 --regs var,js --dtypes bf16 2>&1 | grep cfg5
echo "== cfg4 through L2 staging instead of smem ring"; DSNT_TUNE_STEP_L2=1 timeout 200 python tools/kbench.py --configs cfg4 --regs var --dtypes f32 --step-only 2>&1 | grep cfg4
