for f in 2 6; do
  NGP_BWD_FLAGS=$f timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_f$f.json 2> gpurun_out/bench_f$f.err
done
