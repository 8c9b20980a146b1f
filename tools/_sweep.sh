mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q > gpurun_out/v7_pytest_gpu.log 2>&1; tail -4 gpurun_out/v7_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/v7_bench.json 2> gpurun_out/v7_bench.err; cat gpurun_out/v7_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/v7_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v7_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_step2 -s 3 -c 1 -f -o gpurun_out/v7_step2_f32_js python tools/kbench.py --configs cfg4 --regs js --dtypes f32 --iters 3 --step-only > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_step2 -s 3 -c 1 -f -o gpurun_out/v7_step2_bf16_js python tools/kbench.py --configs cfg4 --regs js --dtypes bf16 --iters 3 --step-only > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_step2 -s 3 -c 1 -f -o gpurun_out/v7_step2_bf16_var python tools/kbench.py --configs cfg4 --regs var --dtypes bf16 --iters 3 --step-only > gpurun_out/ncu3.log 2>&1
