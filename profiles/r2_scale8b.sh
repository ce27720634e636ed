# 8-GPU strong scaling, final state: exchange over NVSwitch multicast (auto) vs plain peer loads / stores (ipc)
run() { tag=$1; shift; timeout 150 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 10 --long-steps 0 > gpurun_out/scale8b_$tag.json 2> gpurun_out/scale8b_$tag.log; echo "$tag rc=$?"; grep -E "gradient exchange|timed region|eager unpip|rror" gpurun_out/scale8b_$tag.log | sort | uniq | cut -c1-200 | tail -4; }
run nvls NGP_EXCHANGE_MEM=auto
run ipc NGP_EXCHANGE_MEM=ipc
