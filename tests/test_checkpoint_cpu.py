"""Checkpoint interchange with the reference trainer's format (nerf/utils.py:1015-1136): FusedFieldOptimizer <-> torch.optim.Adam
state for the reference's parameter-group layout (network_ff.py:137-149), GradScaler state, model keys."""
import os
import sys
import types

import pytest
import torch

import ngp_checkpoint
from nerf_step import NeRFFieldFF
from ngp_optim import FusedFieldOptimizer


def reference_adam(m, lr=1e-2):
    groups = [{"params": list(m.encoder.parameters()), "lr": lr}, {"params": list(m.sigma_net.parameters()), "lr": lr},
              {"params": list(m.encoder_dir.parameters()), "lr": lr}, {"params": list(m.color_net.parameters()), "lr": lr}]
    return torch.optim.Adam(groups, betas=(0.9, 0.99), eps=1e-15)


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return NeRFFieldFF()


def test_adam_state_roundtrip(model, tmp_path):
    opt_ref = reference_adam(model)
    for it in range(3):
        for p in (model.encoder.embeddings, model.sigma_net.weights, model.color_net.weights):
            p.grad = torch.randn_like(p) * 1e-3
        opt_ref.step()
    scaler_sd = {"scale": 4096.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 17}
    ref_ck = {"epoch": 3, "global_step": 300, "stats": {"loss": [0.1]}, "mean_count": 1234, "mean_density": 0.5,
              "model": {**model.state_dict(), "aabb_infer": model.aabb_train.clone()}, "optimizer": opt_ref.state_dict(), "scaler": scaler_sd}
    path = str(tmp_path / "ref.pth")
    torch.save(ref_ck, path)

    torch.manual_seed(1)
    m2 = NeRFFieldFF()
    fused = FusedFieldOptimizer(m2.encoder, m2.sigma_net, m2.color_net)
    ck, missing, unexpected = ngp_checkpoint.load(path, m2, fused)
    assert missing == [] and unexpected == ["aabb_infer"]
    assert m2.mean_count == 1234 and m2.mean_density == 0.5 and ck["epoch"] == 3
    for k, v in model.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k
    assert int(fused.state[3]) == 3 and float(fused.scale_tensor()) == 4096.0 and int(fused.state[1]) == 17
    sd = fused.state_dict()
    for i in range(3):
        assert torch.equal(sd["state"][i]["exp_avg"], opt_ref.state_dict()["state"][i]["exp_avg"])
        assert torch.equal(sd["state"][i]["exp_avg_sq"], opt_ref.state_dict()["state"][i]["exp_avg_sq"])
        assert float(sd["state"][i]["step"]) == 3.0
    assert [g["params"] for g in sd["param_groups"]] == [[0], [1], [], [2]]

    # and back: a checkpoint written here loads into the reference's optimizer / scaler objects
    path2 = str(tmp_path / "ours.pth")
    ngp_checkpoint.save(path2, m2, fused, epoch=4, global_step=400)
    ck2 = torch.load(path2, weights_only=False)
    assert set(ck2) >= {"epoch", "global_step", "stats", "mean_count", "mean_density", "model", "optimizer", "scaler"}
    assert "aabb_infer" in ck2["model"] and "density_grid" in ck2["model"]
    opt3 = reference_adam(model)
    opt3.load_state_dict(ck2["optimizer"])
    assert torch.equal(opt3.state_dict()["state"][2]["exp_avg"], opt_ref.state_dict()["state"][2]["exp_avg"])
    sc = torch.amp.GradScaler("cpu", enabled=True)
    sc.load_state_dict(ck2["scaler"])
    assert sc.state_dict()["scale"] == 4096.0 and sc.state_dict()["_growth_tracker"] == 17
    fused.detach()


@pytest.mark.skipif(not os.path.isdir("/root/reference/nerf"), reason="reference tree only exists in the build container")
def test_checkpoint_loads_into_reference_network(model, tmp_path):
    """a checkpoint written by ngp_checkpoint.save loads into the reference's own NeRFNetwork (strict) and vice versa."""
    for m in ("trimesh", "pysdf", "mcubes", "tensorboardX", "lpips", "torch_ema", "torchmetrics", "imageio", "matplotlib",
              "matplotlib.pyplot", "cv2"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.append("/root/reference")
    try:
        import importlib
        utils_stub = types.ModuleType("nerf.utils")
        utils_stub.custom_meshgrid = lambda *a: torch.meshgrid(*a, indexing="ij")
        sys.modules["nerf.utils"] = utils_stub
        net = importlib.import_module("nerf.network_ff").NeRFNetwork(bound=1, cuda_ray=True)
        path = str(tmp_path / "ours.pth")
        ngp_checkpoint.save(path, model, None)
        ck = torch.load(path, weights_only=False)
        net.load_state_dict(ck["model"], strict=True)
        assert torch.equal(net.encoder.embeddings, model.encoder.embeddings)
        path2 = str(tmp_path / "ref.pth")
        torch.save({"model": net.state_dict(), "epoch": 1, "global_step": 1, "stats": {}, "mean_count": 7, "mean_density": 0.1}, path2)
        m2 = NeRFFieldFF()
        _, missing, unexpected = ngp_checkpoint.load(path2, m2)
        assert missing == [] and unexpected == ["aabb_infer"] and m2.mean_count == 7
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k.startswith(("nerf.", "encoding", "activation")) or k == "nerf"]:
            sys.modules.pop(k, None)
