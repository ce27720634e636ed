# 2-GPU verification: smoke(), the compositor / DP tests, compositor group-size A/B, N=2 bench line
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python -m pytest tests/test_gpu_raymarching.py tests/test_gpu_golden.py tests/test_gpu_dp.py tests/test_gpu_optim.py tests/test_gpu_e2e.py tests/test_gpu_fused.py -m gpu -q --timeout 200 > gpurun_out/verify2_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|peer|Error" gpurun_out/verify2_pytest.log | tail -8 | cut -c1-250
CUDA_VISIBLE_DEVICES=0 bash profiles/ab_env.sh NGP_COMPOSITE_GROUP "8 16 32"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/verify2_n2.json 2> gpurun_out/verify2_n2.log; echo "n2 rc=$?"; grep -E "gradient exchange|timed region" gpurun_out/verify2_n2.log | sort | uniq | cut -c1-200 | tail -3
