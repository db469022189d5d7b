cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FVM_ROW_WARPS=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_unstructured.py tests/test_gpu_boundary_terms.py tests/test_gpu_solvers.py -m gpu -x -q > gpurun_out/r2_14_pytest_rw1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_14_pytest_rw1.log
V=";row_warps=1;row_warps=1,consumer_warps=5;row_warps=2,consumer_warps=6;row_warps=1,consumer_warps=5,ilp=2"
timeout 900 python tools/sweep_tile.py --size 4001 --quantum 3400 --modes 1,0,2 --steps 50 --reps 5 --variants "$V" > gpurun_out/r2_14_sweep_c4.log 2>&1
timeout 900 python tools/sweep_tile.py --cube 200 --quantum 2800 --modes 1,2 --steps 30 --reps 5 --variants "$V" > gpurun_out/r2_14_sweep_c5.log 2>&1
tail -3 gpurun_out/r2_14_pytest_rw1.log; grep "^q=\|FAILED" gpurun_out/r2_14_sweep_c4.log; grep "^q=\|FAILED" gpurun_out/r2_14_sweep_c5.log
