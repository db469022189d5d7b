cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_7_pytest.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_7_bench.log 2> gpurun_out/r2_7_bench.err
( time timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_7_ref.log 2> gpurun_out/r2_7_ref.err
tail -3 gpurun_out/r2_7_pytest.log; tail -5 gpurun_out/r2_7_bench.err; tail -5 gpurun_out/r2_7_ref.err; free -g | head -2
