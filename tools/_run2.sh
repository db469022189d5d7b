cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/src
export FVM_KERNEL_SRC_DIR=$GRAFT_REPO_ROOT/gpurun_out/src
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_2_pytest.log
python tools/sweep_tile.py --size 4001 --modes 1 --steps 3 > gpurun_out/r2_2_plain_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fvm_tile -s 5 -c 1 -o gpurun_out/r2_2_c4res python tools/sweep_tile.py --size 4001 --modes 1 --steps 3 > gpurun_out/r2_2_ncu_a.log 2>&1
python tools/sweep_tile.py --cube 160 --modes 1 --steps 3 --variants "fold_ahead=0" > gpurun_out/r2_2_plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fvm_tile -s 5 -c 1 -o gpurun_out/r2_2_c5res python tools/sweep_tile.py --cube 160 --modes 1 --steps 3 --variants "fold_ahead=0" > gpurun_out/r2_2_ncu_b.log 2>&1
tail -4 gpurun_out/r2_2_pytest.log
