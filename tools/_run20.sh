cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -s > gpurun_out/r2_20_dist_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_20_dist_pytest_n$N.log
tail -12 gpurun_out/r2_20_dist_pytest_n$N.log | cut -c1-250
