"""Checkpoint loading with the reference's legacy key remap (reference src/eval.py:15-52)."""
import torch


def remap_legacy_keys(state_dict):
    """`pose_net.conv.` -> `pose_net.` and `sample_net.` -> `sample_nets.0.` (eval.py:41-42)."""
    out = {}
    for k, v in state_dict.items():
        k = k.replace("pose_net.conv.", "pose_net.")
        if k.startswith("sample_net."):
            k = "sample_nets.0." + k[len("sample_net."):]
        out[k] = v
    return out


def load_checkpoint_with_legacy_fix(checkpoint_path, model, device="cpu"):
    """Same positional arguments as the reference's helper (src/eval.py:27: `(checkpoint_path, model, device)`)."""
    if isinstance(checkpoint_path, torch.nn.Module):          # (model, path) order of round 1: still accepted
        checkpoint_path, model = model, checkpoint_path
    ckpt = torch.load(checkpoint_path, map_location=device, weights_only=False)
    sd = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    try:
        model.load_state_dict(sd, strict=True)
    except RuntimeError:
        model.load_state_dict(remap_legacy_keys(sd), strict=True)
    return model
