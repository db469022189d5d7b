cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -s > gpurun_out/r2_21_dist_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_21_dist_pytest_n$N.log
for n in 8 4 2; do
  if [ $n -le $N ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_21_bench_n$n.log 2> gpurun_out/r2_21_bench_n$n.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --workload c5 --cube 343 --steps 10 --warmup 3 --solve > gpurun_out/r2_21_c5_343_n$n.log 2> gpurun_out/r2_21_c5_343_n$n.err
  fi
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_21_bench_n1.log 2> gpurun_out/r2_21_bench_n1.err
tail -5 gpurun_out/r2_21_dist_pytest_n$N.log | cut -c1-200; for n in 8 4 2; do tail -2 gpurun_out/r2_21_bench_n$n.err | cut -c1-200; tail -3 gpurun_out/r2_21_c5_343_n$n.err | cut -c1-200; done
