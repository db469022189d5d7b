cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/src
export FVM_KERNEL_SRC_DIR=$GRAFT_REPO_ROOT/gpurun_out/src
# launch list of the bench command (cold-cache, serialised: shares only)
python bench.py --steps 20 --warmup 5 --no-cpu --no-small --c5-cube 0 > gpurun_out/r2_15_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_15_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-small --c5-cube 0 > gpurun_out/r2_15_ncu_l.log 2>&1
python tools/sweep_tile.py --size 4001 --modes 0,1 --steps 3 --reps 1 > gpurun_out/r2_15_plain_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fvm_tile -s 5 -c 1 -o gpurun_out/r2_15_c4asm python tools/sweep_tile.py --size 4001 --modes 0 --steps 3 --reps 1 > gpurun_out/r2_15_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fvm_tile -s 5 -c 1 -o gpurun_out/r2_15_c4res python tools/sweep_tile.py --size 4001 --modes 1 --steps 3 --reps 1 > gpurun_out/r2_15_ncu_b.log 2>&1
python tools/sweep_tile.py --cube 160 --modes 1 --steps 3 --reps 1 > gpurun_out/r2_15_plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fvm_tile -s 5 -c 1 -o gpurun_out/r2_15_c5res python tools/sweep_tile.py --cube 160 --modes 1 --steps 3 --reps 1 > gpurun_out/r2_15_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4; tail -2 gpurun_out/r2_15_ncu_l.log
