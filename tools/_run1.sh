cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_1_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_1_pytest.log
timeout 600 python tools/sweep_tile.py --size 4001 --modes 0,1,2,3 --variants ";row_runs=0;row_runs=1;consumer_warps=6;consumer_warps=8" > gpurun_out/r2_1_sweep_c4.log 2>&1
timeout 600 python tools/sweep_tile.py --cube 200 --modes 1,2,3 --variants ";fold_ahead=0;fold_ahead=3;row_runs=0;row_runs=1;consumer_warps=6;consumer_warps=8" > gpurun_out/r2_1_sweep_c5.log 2>&1
(cd _r1 && timeout 600 python bench.py --steps 50 --no-cpu --no-cold > ../gpurun_out/r2_1_oldbench.log 2>&1)
(cd _r1 && timeout 600 python bench.py --workload c5 --cube 200 --steps 30 > ../gpurun_out/r2_1_oldbench_c5.log 2>&1)
tail -5 gpurun_out/r2_1_pytest.log; cat gpurun_out/r2_1_sweep_c4.log | tail -30
