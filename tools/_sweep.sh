mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_model.py -q > gpurun_out/v10_pytest_gpu.log 2>&1; tail -3 gpurun_out/v10_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/v10_bench.json 2> gpurun_out/v10_bench.err; cut -c1-1900 gpurun_out/v10_bench.json; tail -3 gpurun_out/v10_bench.err
