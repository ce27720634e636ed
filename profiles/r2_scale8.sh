# 8-GPU strong-scaling runs of the training step (under `gpurun --gpus 8`): exchange form x prefetch point
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 10 --long-steps 0 "$@" > gpurun_out/scale8_$tag.json 2> gpurun_out/scale8_$tag.log; echo "$tag rc=$?"; grep -E "exchange:|timed region|eager unpipelined|ERROR|Error" gpurun_out/scale8_$tag.log | sort | uniq | tail -5; }
run peer_exchange --exchange peer --prefetch-point exchange
run peer_start --exchange peer --prefetch-point start
run nccl_exchange --exchange nccl --prefetch-point exchange
run peer_noprefetch --exchange peer --no-prefetch
