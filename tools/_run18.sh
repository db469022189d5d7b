cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2_18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_18_pytest.log
timeout 600 python tools/sweep_tile.py --cube 200 --modes 3 --steps 30 --reps 3 > gpurun_out/r2_18_spmv3d.log 2>&1
FVM_SPMV_VIEW=0 timeout 600 python tools/sweep_tile.py --cube 200 --modes 3 --steps 30 --reps 3 > gpurun_out/r2_18_spmv3d_noview.log 2>&1
( time timeout 900 python bench.py --workload c5 --cube 343 --steps 5 --warmup 3 ) > gpurun_out/r2_18_c5_343_n1.log 2> gpurun_out/r2_18_c5_343_n1.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/r2_18_c5_343_n1.err
tail -3 gpurun_out/r2_18_pytest.log; grep "^q=" gpurun_out/r2_18_spmv3d.log gpurun_out/r2_18_spmv3d_noview.log; tail -12 gpurun_out/r2_18_c5_343_n1.err; cut -c1-600 gpurun_out/r2_18_c5_343_n1.log
