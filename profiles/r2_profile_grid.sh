# ncu --set full capture of the table scatter inside the training step (run under gpurun, one GPU)
set -x
COMMON="--steps 3 --warmup 3 --no-graph --no-prefetch --no-maintenance --no-cpu-baseline --no-ref-cuda --long-steps 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_grid_backward' -s 12 -c 2 \
    -o gpurun_out/r2b_prof_grid -f python bench.py $COMMON > gpurun_out/r2b_prof_grid.log 2>&1
tail -3 gpurun_out/r2b_prof_grid.log
