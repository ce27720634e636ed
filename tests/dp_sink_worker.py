"""Worker of tests/test_gpu_dp.py (launched under torch.distributed.run, one process per GPU): every rank runs the training
forward + backward on ITS shard of the same 8192 rays with the fp16 gradient sink installed, the sink is all-reduced (the step's only
exchange, ngp_optim.FusedFieldOptimizer.begin_exchange / finish_exchange), and rank 0 saves it."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200"), os.path.join(ROOT, "tests")]


def sink_after_backward(world, rank, N=8192, via="autograd", exchange="nccl"):
    import ngp_dp
    import ngp_synth as S
    from nerf_step import NeRFFieldFF, FusedTrainStep
    from ngp_optim import FusedFieldOptimizer
    from util import synth_rays, gen
    torch.manual_seed(1)
    model = NeRFFieldFF(bound=1, fused=True).cuda().train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.3, 0.3)
    grid, _ = S.box_union_density(128, seed=12)
    model.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).cuda())
    opt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, lr=1e-2, init_scale=1024.0, exchange=exchange)
    rays_o, rays_d, _, _ = synth_rays(N)
    target = torch.rand(N, 3, generator=gen(5))
    mine = ngp_dp.shard_indices(N, rank, world)
    ro, rd, tg = rays_o[mine].cuda(), rays_d[mine].cuda(), target[mine].cuda()
    with torch.autocast("cuda", dtype=torch.float16):
        out = model.render_train(ro, rd, perturb=False, force_all_rays=True)
        loss = ((out["image"] - tg) ** 2).sum() / (3.0 * N)
    (loss * opt.scale_tensor()).backward()
    if opt.px is not None:
        # peer path: the reduce-scatter half of the fused exchange (csrc/exchange.cu) — every rank ends up with the reduced gradient
        # of ITS shard in its own bucket; the shards are then collected for the comparison
        import _ngp_b200 as nb
        px = opt.px
        lo, hi = px.my_range
        px.barrier(0)
        nb.call("ngp_exchange_reduce", px.sinks, px.rank, px.world, lo, hi - lo, opt.state.data_ptr())
        px.barrier(1)
        torch.cuda.synchronize()
        full = opt.sink.detach().clone()
        for r in range(px.world):
            dist.broadcast(full[px.bounds[r]:px.bounds[r + 1]], src=r)
        assert px.error() == 0
        return full.float().cpu() / 1024.0, float(loss.detach())
    opt.begin_exchange()
    opt.finish_exchange()
    torch.cuda.synchronize()
    return opt.sink.detach().float().cpu() / 1024.0, float(loss.detach())


if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sink, loss = sink_after_backward(world, rank, exchange=(sys.argv[2] if len(sys.argv) > 2 else "nccl"))
    losses = torch.tensor([loss], device="cuda")
    dist.all_reduce(losses)
    if rank == 0:
        torch.save({"sink": sink, "loss_sum": float(losses.item())}, sys.argv[1])
    dist.barrier()
    dist.destroy_process_group()
