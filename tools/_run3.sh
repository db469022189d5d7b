cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_3_pytest.log
V=";producers=1;producers=2;ilp=1;row_ahead=0;producers=1,ilp=1,row_ahead=0;setmaxnreg=1;stages=3"
timeout 900 python tools/sweep_tile.py --size 4001 --modes 0,1,2,3 --variants "$V" > gpurun_out/r2_3_sweep_c4.log 2>&1
timeout 900 python tools/sweep_tile.py --cube 200 --modes 1,2,3 --variants "$V;fold_ahead=5" > gpurun_out/r2_3_sweep_c5.log 2>&1
tail -3 gpurun_out/r2_3_pytest.log; cut -c1-110 gpurun_out/r2_3_sweep_c4.log; cut -c1-110 gpurun_out/r2_3_sweep_c5.log
