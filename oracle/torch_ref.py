"""torch_ref.py — the reference's "pure PyTorch" NeRF path on CPU (fp32, no --cuda_ray).  TEST INFRASTRUCTURE /
CPU BASELINE ONLY (bench.py `cpu_baseline` and `--impl reference`); never imported by torch-ngp_b200/.

The reference cannot run this path on a CPU as shipped: NeRFRenderer.run (nerf/renderer.py:125-253) calls the CUDA-only
raymarching.near_far_from_aabb (:141) and nerf/network.py uses the CUDA-only GridEncoder / SHEncoder.  Following
SURVEY §8d "CPU baseline plan", the exact control flow of `run` (num_steps=512 uniform samples, upsample_steps=0,
main_nerf.py:29-30) and the nn.Linear field of nerf/network.py:33-105 (sigma: 32->64->16, color: 31->64->64->3) are
restated here in vectorised torch, with the three CUDA-only leaves bound to vectorised torch restatements of
gridencoder.cu:87-245, shencoder.cu:49-68 (degree 4) and raymarching.cu:91-145.
"""
import math

import numpy as np
import torch
import torch.nn as nn

_PRIMES = (1, 2654435761, 805459861)


class TorchGridEncoder(nn.Module):
    """Differentiable torch restatement of the hash grid (gather + trilinear blend), fp32."""

    def __init__(self, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048):
        super().__init__()
        from . import oracle as O
        offsets, pls = O.grid_offsets(3, num_levels, level_dim, 2, base_resolution, log2_hashmap_size, desired_resolution)
        self.offsets = [int(o) for o in offsets]
        self.S = float(np.log2(pls)); self.H = base_resolution; self.L = num_levels; self.C = level_dim
        self.output_dim = num_levels * level_dim
        self.embeddings = nn.Parameter(torch.empty(self.offsets[-1], level_dim).uniform_(-1e-4, 1e-4))

    def forward(self, x, bound=1):
        x = (x + bound) / (2 * bound)
        outs = []
        for l in range(self.L):
            off, size = self.offsets[l], self.offsets[l + 1] - self.offsets[l]
            scale = float(np.float32(2.0) ** np.float32(l * self.S) * self.H - 1.0)
            res = int(math.ceil(scale)) + 1
            pos = x * scale + 0.5
            pg = torch.floor(pos)
            fr = pos - pg
            pg = pg.long()
            acc = 0
            for idx in range(8):
                w = 1.0
                c = []
                for d in range(3):
                    if idx & (1 << d):
                        w = w * fr[:, d]; c.append(pg[:, d] + 1)
                    else:
                        w = w * (1 - fr[:, d]); c.append(pg[:, d])
                if (res + 1) ** 3 <= size:                 # dense level
                    index = c[0] + c[1] * (res + 1) + c[2] * (res + 1) ** 2
                else:                                      # xor-prime hash, uint32 arithmetic
                    index = ((c[0] * _PRIMES[0]) & 0xffffffff) ^ ((c[1] * _PRIMES[1]) & 0xffffffff) ^ ((c[2] * _PRIMES[2]) & 0xffffffff)
                index = index % size + off
                acc = acc + w.unsqueeze(-1) * self.embeddings[index]
            outs.append(acc)
        return torch.cat(outs, dim=-1)


def sh4(d):
    """degree-4 real SH (16 channels), the closed forms of shencoder.cu:50-68 via their generating structure."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    c = [0.28209479177387814 * torch.ones_like(x),
         -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
         1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
         -1.0925484305920792 * xz, 0.54627421529603959 * (x2 - y2),
         0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
         0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
         1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(c, dim=-1)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.cu:91-145, vectorised."""
    rd = 1.0 / rays_d
    t0 = (aabb[:3] - rays_o) * rd
    t1 = (aabb[3:] - rays_o) * rd
    tmin = torch.minimum(t0, t1); tmax = torch.maximum(t0, t1)
    near = tmin.max(dim=-1).values; far = tmax.min(dim=-1).values
    miss = near > far
    near = near.clamp(min=min_near)
    big = torch.finfo(torch.float32).max
    return torch.where(miss, torch.full_like(near, big), near), torch.where(miss, torch.full_like(far, big), far)


class TorchNeRF(nn.Module):
    """nerf/network.py:11-105 (nn.Linear MLPs) + nerf/renderer.py:125-253 (`run`, num_steps=512, upsample_steps=0)."""

    def __init__(self, bound=1, hidden=64, geo_feat_dim=15, num_steps=512):
        super().__init__()
        self.bound = bound; self.num_steps = num_steps
        self.encoder = TorchGridEncoder(desired_resolution=2048 * bound)
        self.sigma_net = nn.Sequential(nn.Linear(32, hidden, bias=False), nn.ReLU(), nn.Linear(hidden, 1 + geo_feat_dim, bias=False))
        self.color_net = nn.Sequential(nn.Linear(16 + geo_feat_dim, hidden, bias=False), nn.ReLU(),
                                       nn.Linear(hidden, hidden, bias=False), nn.ReLU(), nn.Linear(hidden, 3, bias=False))
        self.register_buffer("aabb", torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32))

    def render(self, rays_o, rays_d, bg_color=1.0, perturb=True):
        N, T = rays_o.shape[0], self.num_steps
        nears, fars = near_far_from_aabb(rays_o, rays_d, self.aabb)
        nears = nears.unsqueeze(-1); fars = fars.unsqueeze(-1)
        z = torch.linspace(0.0, 1.0, T).unsqueeze(0).expand(N, T)
        z = nears + (fars - nears) * z
        sample_dist = (fars - nears) / T
        if perturb:
            z = z + (torch.rand(z.shape) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, self.aabb[:3]), self.aabb[3:]).reshape(-1, 3)
        h = self.sigma_net(self.encoder(xyzs, self.bound))
        sigma = torch.exp(h[:, 0].clamp(max=15)).view(N, T)
        geo = h[:, 1:]
        deltas = torch.cat([z[:, 1:] - z[:, :-1], sample_dist], dim=-1)
        alphas = 1 - torch.exp(-deltas * sigma)
        shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(shifted, dim=-1)[:, :-1]
        dirs = rays_d.unsqueeze(1).expand(N, T, 3).reshape(-1, 3)
        rgb = torch.sigmoid(self.color_net(torch.cat([sh4(dirs), geo], dim=-1))).view(N, T, 3)
        wsum = weights.sum(-1)
        image = (weights.unsqueeze(-1) * rgb).sum(-2) + (1 - wsum).unsqueeze(-1) * bg_color
        return image


def train_steps(n_rays, steps, warmup, rays_o, rays_d, target, threads=None, seed=0):
    """Time `steps` CPU training iterations (forward, MSE, backward, Adam) of n_rays rays each; returns seconds/step."""
    import time
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    model = TorchNeRF()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    times = []
    for it in range(warmup + steps):
        s = (it * n_rays) % max(1, rays_o.shape[0] - n_rays + 1)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        img = model.render(rays_o[s:s + n_rays], rays_d[s:s + n_rays])
        loss = torch.nn.functional.mse_loss(img, target[s:s + n_rays])
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times)), float(loss.detach())
