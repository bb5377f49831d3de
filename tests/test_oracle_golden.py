"""Pins oracle/handmvnet_oracle.py against fixtures produced by the real reference
(oracle/gen_golden.py; reference src/models/handmvnet.py:158-266)."""
import glob
import os

import numpy as np
import pytest
import torch

import handmvnet_oracle as O

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("preprocess"))


def _run(g):
    backbone = str(g["meta_backbone"]) if "meta_backbone" in g.files else "resnet"
    cfg = O.release_config(int(g["meta_num_views"]), bool(g["meta_crop"]), backbone)
    sd = O.make_state_dict(cfg, seed=int(g["meta_seed_w"]), randomize_norm=bool(g["meta_randomize_norm"]))
    x, bbox, intr = O.make_inputs(int(g["meta_batch"]), int(g["meta_num_views"]), seed=int(g["meta_seed_x"]))
    if bool(g["meta_crop"]):
        return O.forward(sd, cfg, x, bbox, intr, return_taps=True)
    return O.forward(sd, cfg, x, return_taps=True)


def test_fixtures_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_fixture(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    out, taps = _run(g)
    # outputs (fp32 CPU, same torch build: differences are accumulation-order only)
    hm = out["heatmap"]
    np.testing.assert_array_equal(hm.flatten(-2).argmax(-1).numpy(), g["out_heatmap_argmax"])
    np.testing.assert_allclose(hm[..., ::4, ::4].numpy(), g["out_heatmap_sub"], rtol=1e-4, atol=1e-4 * max(1.0, float(np.abs(g["out_heatmap_sub"]).max())))
    np.testing.assert_allclose(out["joints_crop_img"].numpy(), g["out_joints_crop_img"], rtol=0, atol=1e-3)
    # Conditioning: the unit-gain HRNet fixture (the reference's own init) feeds tokens of magnitude ~3e6 straight into
    # to_q / to_k (no pre-norm, layers.py:213-215), so the attention logits are ~1e13 and every softmax is a hard max: the
    # fp32 GEMM summation order of the host's BLAS (AVX2 vs AVX-512 kernels) then moves the fusion stages by up to 3e-3
    # although everything up to and including the tokens reproduces bit for bit.  Such a fixture pins the transformer and
    # the graph head at 1e-2 / 5e-6 m; every well-conditioned fixture at 1e-4 / 1e-6 m.
    ill = "stage_tokens_absmean" in g.files and float(g["stage_tokens_absmean"]) > 1e3
    down_tol, kp_tol = (1e-2, 5e-6) if ill else (1e-4, 1e-6)
    # final keypoints: the north-star criterion is 0.1 mm (1e-4 m); the oracle must be far inside it
    assert np.abs(out["joints_cam"].numpy() - g["out_joints_cam"]).max() < kp_tol  # metres
    # every stage fingerprint
    keys = [k[len("stage_"):-len("_shape")] for k in g.files if k.startswith("stage_") and k.endswith("_shape")]
    assert len(keys) >= 12
    for key in keys:
        t = taps[key]
        assert list(t.shape) == list(g[f"stage_{key}_shape"]), key
        flat = t.reshape(-1)
        ref = g[f"stage_{key}_val"]
        got = flat[torch.from_numpy(g[f"stage_{key}_idx"])].numpy()
        scale = float(g[f"stage_{key}_absmean"]) + 1e-12
        tol = down_tol if key.startswith("fusion") or key == "joints_cam" else 1e-4
        assert np.abs(got - ref).max() / scale < tol, key
        assert abs(float(flat.double().mean()) - float(g[f"stage_{key}_mean"])) / scale < tol, key


def test_sampler_gather_identity():
    """gather-then-conv == dense conv then grid_sample."""
    cfg = O.release_config(5)
    sd = O.make_state_dict(cfg, seed=4)
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(2, 1024, 32, 32, generator=g)
    xy = torch.rand(2, 21, 2, generator=g) * 31
    xy[0, :5] = xy[0, :5].round()                 # integer coords (the hard-argmax case)
    xy[0, 5] = torch.tensor([31.0, 31.0]); xy[0, 6] = torch.tensor([0.0, 31.0])
    a = O.sample_net(sd, feat, xy)
    b = O.sample_net_gather(sd, feat, xy)
    assert (a - b).abs().max() < 2e-4 * a.abs().max()


def test_known_answers():
    # soft-argmax of a one-hot heatmap is that pixel (x = col, y = row); utils.py:35-62
    hm = torch.zeros(1, 2, 32, 32)
    hm[0, 0, 7, 19] = 1.0
    hm[0, 1, 31, 0] = 1.0
    xy = O.soft_argmax_2d(hm)
    assert torch.allclose(xy[0, 0], torch.tensor([19.0, 7.0]), atol=1e-5)
    assert torch.allclose(xy[0, 1], torch.tensor([0.0, 31.0]), atol=1e-5)
    # adjacency: rows sum to 1, wrist has 6 entries, tips 2, others 3; 61 non-zeros
    a = O.hand_adjacency()
    assert torch.allclose(a.sum(1), torch.ones(21))
    nnz = (a > 0).sum(1)
    assert int(nnz[0]) == 6 and all(int(nnz[t]) == 2 for t in (4, 8, 12, 16, 20)) and int((a > 0).sum()) == 61
    t = O.cheb_basis(3)
    lap = torch.eye(21) - a
    assert torch.allclose(t[1], lap, atol=1e-6) and torch.allclose(t[2], 2 * lap @ lap - torch.eye(21), atol=1e-6)
    # PE table layers.py:136-150
    pe = O.positional_table(524, 105)
    assert pe.shape == (105, 524) and pe[0, 0] == 0 and pe[0, 1] == 1
    # view-count mismatch must raise (the reference fails late / regroups silently)
    cfg = O.release_config(5)
    with pytest.raises(ValueError):
        O.forward({}, cfg, torch.zeros(1, 8, 3, 256, 256))


def test_preprocess_oracle_matches_reference_fixture(golden_dir):
    """crop_and_pad_image + ToTensor + Resize(antialias) + Normalize: the oracle restatement against the fixture made by
    oracle/gen_golden_preprocess.py with the reference's own crop function and torchvision's transforms."""
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    frames, bboxes = O.make_frames(8, seed=0)
    np.testing.assert_array_equal(bboxes, g["bboxes"])
    out = O.preprocess(frames, bboxes)
    assert tuple(out.shape) == tuple(g["shape"])
    np.testing.assert_allclose(out[:, :, ::8, ::8].numpy(), g["sub"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out.reshape(-1)[torch.from_numpy(g["idx"])].numpy(), g["val"], rtol=0, atol=1e-6)
    assert abs(float(out.double().mean()) - float(g["mean"])) < 1e-6
    # edge cases of the crop: a box entirely outside the frame is all zeros before normalisation
    z = O.crop_and_pad_image(frames[0], (700, 500, 760, 560))
    assert z.shape == (60, 60, 3) and not z.any()
