"""Worker of tests/test_gpu_dp.py::test_peer_exchange_matches_nccl (one process per GPU under torch.distributed.run).

Three optimizer steps on identical per-rank gradients through FusedFieldOptimizer with exchange = "peer" (csrc/exchange.cu: reduce-scatter by
peer loads + sharded Adam + fp16 operand copies stored into every replica) or "nccl" (all-reduce + full optimizer pass on every rank).
Step 2 carries an inf on rank 1 only: both forms must skip it on every rank and back the loss scale off.  Rank 0 saves the resulting fp16
operand copies, the gathered fp32 masters / moments and the scaler state."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200"), os.path.join(ROOT, "tests")]


def run(mode, rank, world):
    from nerf_step import NeRFFieldFF
    from ngp_optim import FusedFieldOptimizer
    torch.manual_seed(1)
    model = NeRFFieldFF(bound=1, fused=True).cuda().train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.3, 0.3)
    opt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, lr=1e-2, init_scale=1024.0, exchange=mode)
    assert (opt.px is not None) == (mode == "peer")
    n = opt.sink.numel()
    shadows = []
    for step in range(3):
        g = torch.Generator(device="cuda").manual_seed(100 * step + rank)
        # sparse-ish gradient like the scatter produces: most entries zero
        grad = torch.randn(n, generator=g, device="cuda") * (torch.rand(n, generator=g, device="cuda") < 0.2)
        opt.sink.copy_(grad.half())
        if step == 1 and rank == world - 1:
            opt.sink[12345] = float("inf")
        opt.step()
        torch.cuda.synchronize()
        assert float(opt.sink.abs().max()) == 0.0, "gradient bucket must be cleared by the step"
        shadows.append(opt.shadow_flat.clone())
    if opt.px is not None:
        assert opt.px.error() == 0
    # every replica holds the same operand copies
    for s in shadows:
        ref = s.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, s), "replicas' fp16 operand copies differ"
    opt.gather_master()
    flat_p = torch.cat([p.detach().reshape(-1) for p in opt.params])
    out = dict(shadow=shadows[-1].cpu(), shadow1=shadows[1].cpu(), shadow0=shadows[0].cpu(), params=flat_p.cpu(), exp_avg=opt.exp_avg.cpu(), exp_avg_sq=opt.exp_avg_sq.cpu(),
               state=opt.state.cpu(), scale=float(opt.scale_tensor().item()),
               memory=(opt.px.memory if opt.px is not None else "nccl"), nvls=bool(opt.px is not None and opt.px.mc_sink))
    # shadow == half(master) everywhere after the gather
    assert torch.equal(flat_p.half(), shadows[-1])
    opt.detach()
    return out


if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run(sys.argv[2], rank, world)
    if rank == 0:
        torch.save(res, sys.argv[1])
    dist.barrier()
    dist.destroy_process_group()
