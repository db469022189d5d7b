cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -s > gpurun_out/r2_23_dist_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_23_dist_pytest_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --size 2830 --steps 20 --warmup 5 --c5-cube 128 --no-pipelined > gpurun_out/r2_23_b.log 2> gpurun_out/r2_23_b.err
tail -4 gpurun_out/r2_23_dist_pytest_n$N.log | cut -c1-200
python - <<'PY'
import json
for l in open('gpurun_out/r2_23_b.log'):
    if l.startswith('{'):
        d=json.loads(l); print('asm',d['ms_per_step'],'res',d['residual']['ms_per_eval'],'jac',d['jacobian']['ms_per_refresh'], 'parity', d.get('dist_parity'), 'c5', d['c5']['ms_per_step'])
PY
