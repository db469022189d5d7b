cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for a in 1 0; do
FVM_BENCH_ALIGN=$a timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2956$a bench.py --gpus 4 --size 2830 --steps 20 --warmup 5 --c5-cube 0 --no-parity --no-pipelined > gpurun_out/r2_22_align$a.log 2> gpurun_out/r2_22_align$a.err
done
python - <<'PY'
import json
for a in (1,0):
    for l in open(f'gpurun_out/r2_22_align{a}.log'):
        if l.startswith('{'):
            d=json.loads(l); print('align',a,'asm',d['ms_per_step'],'res',d['residual']['ms_per_eval'],'jac',d['jacobian']['ms_per_refresh'], d['config']['l2'][:20])
PY
