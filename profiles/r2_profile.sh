# Round-2 profiling recipe (run under gpurun, one GPU).  Outputs land in gpurun_out/ and are summarised into profiles/r2_*.md by
# profiles/ncu_extract.py on the CPU box.
set -x
mkdir -p gpurun_out
COMMON="--steps 3 --warmup 3 --no-graph --no-prefetch --no-maintenance --no-cpu-baseline --no-ref-cuda --long-steps 0"
# (1) one training step, every hot kernel once, full sections + source (the step's kernel order: march, sigma fwd, color fwd,
#     composite fwd, composite bwd, color bwd, sigma bwd, table scatter); launches 0..99 are the budget / warm-up steps
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_march_rays_train|k_ffmlp_backward_dual|k_grid_backward|k_composite_train|k_ffmlp_forward' -s 104 -c 8 \
    -o gpurun_out/r2_prof_step -f python bench.py $COMMON > gpurun_out/r2_prof_step.log 2>&1
# (2) launch list of the same command (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py $COMMON > gpurun_out/r2_launches.log 2>&1
