cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/sweep_tile.py --size 4001 --quantum 1800,2000,2200,2500,2700,3000,3400 --modes 1,0 --steps 50 --reps 3 --variants ";stages=3;stages=3,min_ctas=5;min_ctas=5" > gpurun_out/r2_10_sweep_c4.log 2>&1
timeout 900 python tools/sweep_tile.py --cube 200 --quantum 1800,2200,2800 --modes 1 --steps 30 --reps 3 --variants ";stages=3" > gpurun_out/r2_10_sweep_c5.log 2>&1
grep "^q=" gpurun_out/r2_10_sweep_c4.log; grep "^q=" gpurun_out/r2_10_sweep_c5.log
