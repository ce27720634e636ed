# 2-GPU probe: does the exchange run on torch symmetric memory + NVSwitch multicast here, and what does it cost?  (under `gpurun --gpus 2`)
run() { tag=$1; shift; timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --long-steps 0 --rays-per-step 160000 > gpurun_out/probe_$tag.json 2> gpurun_out/probe_$tag.log; echo "$tag rc=$?"; grep -E "gradient exchange|timed region|rror|Traceback" gpurun_out/probe_$tag.log | sort | uniq | cut -c1-300 | tail -5; }
run auto NGP_EXCHANGE_MEM=auto
run ipc NGP_EXCHANGE_MEM=ipc
timeout 400 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 180 2>&1 | tail -4 | cut -c1-300
