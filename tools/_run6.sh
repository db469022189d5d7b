cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r2_6_clocks.csv &
SMI=$!
V=";ilp=1;ilp=1,min_ctas=4;min_ctas=4;ilp=1,row_ahead=0;row_runs=0,ilp=1"
timeout 900 python tools/sweep_tile.py --size 4001 --quantum 3200,3432 --modes 0,1,2,3 --variants "$V" > gpurun_out/r2_6_sweep_c4.log 2>&1
timeout 900 python tools/sweep_tile.py --cube 200 --quantum 2600,2800,3000 --modes 1,2 --variants "$V" > gpurun_out/r2_6_sweep_c5.log 2>&1
kill $SMI
(cd _r1 && timeout 600 python bench.py --steps 200 --no-cpu --no-cold > ../gpurun_out/r2_6_oldbench.log 2>&1)
grep "^q=" gpurun_out/r2_6_sweep_c4.log; grep "^q=" gpurun_out/r2_6_sweep_c5.log; sort gpurun_out/r2_6_clocks.csv | uniq -c | sort -rn | head -8
python - <<'PY'
import json
for l in open('gpurun_out/r2_6_oldbench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('OLD', d['ms_per_step'], d['residual']['ms_per_eval'], d['spmv']['ms'])
PY
