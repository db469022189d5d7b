cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_dist_gpu.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2_13_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_13_pytest_n$N.log
for n in 8 4 2; do
  if [ $n -le $N ]; then
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --verbose ) > gpurun_out/r2_13_bench_n$n.log 2> gpurun_out/r2_13_bench_n$n.err
  fi
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-small > gpurun_out/r2_13_bench_n1.log 2> gpurun_out/r2_13_bench_n1.err
tail -4 gpurun_out/r2_13_pytest_n$N.log; for n in 1 2 4 8; do tail -3 gpurun_out/r2_13_bench_n$n.err; done
