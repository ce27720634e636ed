# Final round-2 run on one B200 (under gpurun): full GPU test suite, the bench line with every arm, the measured experiments, the other
# BASELINE configs, and the ncu captures behind profiles/r2c_*.md.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/final_pytest.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.log; echo "bench rc=$?"; grep -E "timed region|e2e done|eager" gpurun_out/final_bench_n1.log
bash profiles/ab_env.sh NGP_SIGMA_PAIR_LOADS "1"
bash profiles/ab_env.sh NGP_SIGMA_TMA_LEVELS "1 2"
for c in c1 c3 c5 infer; do timeout 300 python bench.py --config $c > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.log; echo "$c rc=$?"; head -c 400 gpurun_out/final_bench_$c.json; echo; done
COMMON="--steps 3 --warmup 3 --no-graph --no-prefetch --no-maintenance --no-cpu-baseline --no-ref-cuda --long-steps 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2c_launches.csv python bench.py $COMMON > gpurun_out/r2c_launches.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'k_march_rays_train|k_ffmlp_backward_dual|k_grid_backward|k_composite_train|k_ffmlp_forward' -s 104 -c 8 \
    -o gpurun_out/r2c_prof_step -f python bench.py $COMMON > gpurun_out/r2c_prof_step.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
