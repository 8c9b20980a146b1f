mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_round2.py -m gpu -q -x -k "too_large or edge_counts or pair" > gpurun_out/r2b/t4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b/t4.log
for v in A B C D; do
for cs in 2 4; do
echo "== variant $v (A = geometry ahead + staged stores, B = no staging, C = no geometry, D = neither) CS=$cs"
lib=$PWD/tools/probe/lib_$v.so; [ $v = A ] && lib=$PWD/dsnt_pose2d_b200/libdsnt_b200.so
DSNT_B200_LIB=$lib DSNT_TUNE_STEP_PAIR_CS=$cs timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
done
done
echo "== bf16 A"
for cs in 2 4; do DSNT_TUNE_STEP_PAIR_CS=$cs timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes bf16 --step-only 2>&1 | grep -v "^HBM\|^cfg  "; done
echo "== trace CS=2"
DSNT_B200_LIB=$PWD/tools/probe/lib_T.so DSNT_TUNE_STEP_PAIR_FLAGS=8 DSNT_TUNE_STEP_PAIR_CS=2 timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only --iters 3 2>&1 | grep "pair trace" | awk '++n % 6 == 0'
