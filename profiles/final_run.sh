set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -4 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
tail -2 gpurun_out/bench_final_ref.err
