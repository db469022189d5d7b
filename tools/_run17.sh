cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_17_pytest.log
tail -25 gpurun_out/r2_17_pytest.log | cut -c1-200
