mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_model.py tests/test_gpu_parity.py -q > gpurun_out/v11_pytest_gpu.log 2>&1; tail -3 gpurun_out/v11_pytest_gpu.log
timeout 300 python tools/stepbench.py > gpurun_out/v11_stepbench.txt 2>&1; cat gpurun_out/v11_stepbench.txt
