cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -s > gpurun_out/r2_8_dist_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_8_dist_pytest_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --verbose ) > gpurun_out/r2_8_bench_n$N.log 2> gpurun_out/r2_8_bench_n$N.err
tail -25 gpurun_out/r2_8_dist_pytest_n$N.log; tail -8 gpurun_out/r2_8_bench_n$N.err; grep -c "^{" gpurun_out/r2_8_bench_n$N.log
