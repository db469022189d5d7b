cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=";ilp=1;row_ahead=0;ilp=1,row_ahead=0;ilp=1,row_ahead=3;row_ahead=3;min_ctas=4;min_ctas=5;row_runs=0"
timeout 900 python tools/sweep_tile.py --size 4001 --quantum 3000,3200,3432 --modes 0,1,2,3 --variants "$V" > gpurun_out/r2_5_sweep_c4.log 2>&1
timeout 900 python tools/sweep_tile.py --cube 200 --quantum 2400,2800,3400 --modes 1,2,3 --variants "$V;fold_ahead=3;fold_ahead=5" > gpurun_out/r2_5_sweep_c5.log 2>&1
grep -o "q=.*frac=[0-9.]*\|'grid': [0-9]*\|'regs': [0-9]*" gpurun_out/r2_5_sweep_c4.log | paste - - - ; grep -o "q=.*frac=[0-9.]*\|'grid': [0-9]*\|'regs': [0-9]*" gpurun_out/r2_5_sweep_c5.log | paste - - -
