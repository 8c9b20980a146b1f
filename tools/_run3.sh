mkdir -p gpurun_out/r2b
echo "== OLD kernel (HEAD), order none,js,var,mse"
timeout 300 python tools/probe/old/tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
for cs in 2 4; do
echo "== NEW CS=$cs"
DSNT_TUNE_STEP_PAIR_CS=$cs timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
echo "== NEW CS=$cs trace"
DSNT_TUNE_STEP_PAIR_FLAGS=8 DSNT_TUNE_STEP_PAIR_CS=$cs timeout 300 python tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only --iters 3 2>&1 | grep -v "^HBM\|^cfg  " | awk '!/pair trace/ || ++n % 6 == 0'
done
echo "== OLD again"
timeout 300 python tools/probe/old/tools/kbench.py --configs cfg5 --regs none,js,var,mse --dtypes f32 --step-only 2>&1 | grep -v "^HBM\|^cfg  "
