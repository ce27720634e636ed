"""ngp_checkpoint.py — the reference trainer's checkpoint format (nerf/utils.py:1015-1136), read and written for the hot path's
model and fused optimizer, so checkpoints move between torch-ngp and this package in both directions.

File = torch.save of {'epoch', 'global_step', 'stats', 'mean_count', 'mean_density', 'model': state_dict
[, 'optimizer': Adam state, 'lr_scheduler', 'scaler': GradScaler state, 'ema']}.  Model keys are those of
nerf/network_ff.py's NeRFNetwork (aabb_train, aabb_infer, density_grid, density_bitfield, step_counter, encoder.embeddings,
encoder.offsets, sigma_net.weights, color_net.weights); NeRFFieldFF has no aabb_infer, which load() tolerates like the
reference does (strict=False with a report of missing / unexpected keys).
"""
import torch


def save(path, model, optimizer=None, epoch=0, global_step=0, stats=None, full=True, lr_scheduler=None):
    state = {"epoch": epoch, "global_step": global_step,
             "stats": stats if stats is not None else {"loss": [], "valid_loss": [], "results": [], "checkpoints": [], "best_result": None}}
    if getattr(model, "cuda_ray", False):
        state["mean_count"] = model.mean_count
        state["mean_density"] = model.mean_density
    if full and optimizer is not None:
        state["optimizer"] = optimizer.state_dict()
        if hasattr(optimizer, "scaler_state_dict"):
            state["scaler"] = optimizer.scaler_state_dict()
        if lr_scheduler is not None:
            state["lr_scheduler"] = lr_scheduler.state_dict()
    sd = model.state_dict()
    if "aabb_infer" not in sd and "aabb_train" in sd:
        sd["aabb_infer"] = sd["aabb_train"].clone()         # the reference model registers both (renderer.py:80-83)
    state["model"] = sd
    torch.save(state, path)
    return state


def load(path, model, optimizer=None, model_only=False, map_location=None):
    """Returns (checkpoint dict, missing keys, unexpected keys)."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    sd = ck["model"] if "model" in ck else ck
    missing, unexpected = model.load_state_dict(sd, strict=False)
    emb = getattr(getattr(model, "encoder", None), "embeddings", None)
    if emb is not None:
        from ngp_autograd import sync_half_table
        sync_half_table(emb)                                  # an owner-maintained fp16 kernel operand follows the loaded master
    if "model" not in ck:
        return ck, list(missing), list(unexpected)
    if getattr(model, "cuda_ray", False):
        if "mean_count" in ck:
            model.mean_count = ck["mean_count"]
        if "mean_density" in ck:
            model.mean_density = ck["mean_density"]
    if not model_only and optimizer is not None:
        if "optimizer" in ck:
            optimizer.load_state_dict(ck["optimizer"])
        if "scaler" in ck and ck["scaler"] and hasattr(optimizer, "load_scaler_state_dict"):
            optimizer.load_scaler_state_dict(ck["scaler"])
    return ck, list(missing), list(unexpected)
