mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_model.py tests/test_gpu_parity.py -q > gpurun_out/v8_pytest.log 2>&1; tail -3 gpurun_out/v8_pytest.log
K="timeout 200 python tools/kbench.py --configs cfg4 --step-only"
echo "== default"; $K --regs none,var,js,mse --dtypes f32,bf16 2>&1 | grep one-pass
echo "== bf16 js/mse PACED=1"; DSNT_TUNE_STEP_PACED=1 $K --regs js,mse --dtypes bf16 2>&1 | grep one-pass
timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/v8_bench.json 2> gpurun_out/v8_bench.err; cat gpurun_out/v8_bench.json
