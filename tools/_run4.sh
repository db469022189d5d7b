cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_unstructured.py -m gpu -x -q > gpurun_out/r2_4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_4_pytest.log
V=";l2_ahead=0;l2_ahead=1;l2_ahead=3;producers=1;producers=1,l2_ahead=0;ilp=1;producers=2"
timeout 900 python tools/sweep_tile.py --size 4001 --quantum 3200,3432 --modes 0,1,2,3 --variants "$V" > gpurun_out/r2_4_sweep_c4.log 2>&1
timeout 900 python tools/sweep_tile.py --cube 200 --modes 1,2,3 --variants "$V;fold_ahead=5" > gpurun_out/r2_4_sweep_c5.log 2>&1
tail -3 gpurun_out/r2_4_pytest.log; grep -o "q=.*frac=[0-9.]*\|'grid': [0-9]*" gpurun_out/r2_4_sweep_c4.log | paste - - ; grep -o "q=.*frac=[0-9.]*\|'grid': [0-9]*" gpurun_out/r2_4_sweep_c5.log | paste - -
