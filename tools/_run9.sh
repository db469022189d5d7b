cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_9_pytest.log
timeout 600 python tools/profile_cold.py > gpurun_out/r2_9_cold.log 2>&1
tail -15 gpurun_out/r2_9_pytest.log; head -60 gpurun_out/r2_9_cold.log
