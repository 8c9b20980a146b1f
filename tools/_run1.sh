set -x
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_round2.py -m gpu -q -x -k "too_large or edge_counts or pair" > gpurun_out/r2b/t1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2b/t1.log
for cs in 2 4; do
DSNT_TUNE_STEP_PAIR_CS=$cs timeout 300 python tools/kbench.py --configs cfg5 --regs js,var,none,mse --dtypes f32,bf16 --step-only > gpurun_out/r2b/kb_cs$cs.txt 2>&1; cat gpurun_out/r2b/kb_cs$cs.txt
done
