cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_12_pytest.log
timeout 900 python tools/sweep_tile.py --size 4001 --quantum 3400,3800,4079 --modes 1,0,2,3 --steps 50 --reps 3 --variants ";row_ahead=0" > gpurun_out/r2_12_sweep_c4.log 2>&1
timeout 900 python tools/sweep_tile.py --cube 200 --quantum 2800,3200,3600,3900 --modes 1,2,3 --steps 30 --reps 3 --variants ";row_ahead=0" > gpurun_out/r2_12_sweep_c5.log 2>&1
tail -3 gpurun_out/r2_12_pytest.log; grep "^q=\|FAILED" gpurun_out/r2_12_sweep_c4.log; grep "^q=\|FAILED" gpurun_out/r2_12_sweep_c5.log
