cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/sweep_tile.py --size 4001 --quantum 3400,4500,5100,6000,6800 --modes 1,0 --steps 50 --reps 3 --variants ";consumer_warps=6;consumer_warps=8;consumer_warps=8,min_ctas=2;consumer_warps=6,min_ctas=3" > gpurun_out/r2_11_sweep_c4.log 2>&1
grep "^q=\|FAILED" gpurun_out/r2_11_sweep_c4.log
