"""Seeded fuzz of the scene kernels (depth estimation, fog-density map) and of awx_fogloss against the oracle on
random frame sizes / label dtypes / loss variants."""

import numpy as np
import pytest
import torch

from oracle import prep as op, weather as ow, loss as ol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(12))
def test_scene_fuzz(seed):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import scipy_gaussian_weights
    rng = np.random.RandomState(50 + seed)
    b = int(rng.randint(1, 4))
    h = int(rng.choice([3, 5, 18, 33, 64, 97]))
    w = int(rng.choice([3, 7, 16, 40, 80, 131]))
    imgs = rng.randint(0, 256, (b, h, w, 3)).astype(np.uint8)
    if seed % 4 == 0:
        imgs[0] = 200   # flat frame: zero contrast / zero Laplacian everywhere
    got = ops_prep.estimate_depth(torch.from_numpy(imgs), scipy_gaussian_weights(2.0)).cpu().numpy()
    for i in range(b):
        assert np.array_equal(got[i], op.estimate_depth(imgs[i])), (h, w, i)
    depth = np.stack([ow.depth_from_noise(rng.normal(0, 10, (h, w))) for _ in range(b)])
    imgf = imgs.astype(np.float32) / 255.0
    fog = ops_prep.fog_density_map(torch.from_numpy(imgf), torch.from_numpy(depth)).cpu().numpy()
    con = ops_prep.local_contrast(torch.from_numpy(imgf)).cpu().numpy()
    for i in range(b):
        c_ref = op.local_contrast(imgf[i])
        f_ref = op.fog_density_map(imgf[i], depth[i])
        if w % 16 == 0:
            assert np.array_equal(con[i], c_ref), (h, w, i)
            assert np.array_equal(fog[i], f_ref), (h, w, i)
        else:
            # OpenCV's scalar tail (last W mod 8 columns) rounds the product separately: 1 ulp of the contrast
            assert np.abs(con[i] - c_ref).max() <= 2.4e-7 * max(float(c_ref.max()), 1e-3) + 1e-9, (h, w, i)
            assert np.abs(fog[i] - f_ref).max() <= 5e-6, (h, w, i)


@pytest.mark.parametrize("seed", range(12))
def test_loss_fuzz(seed):
    import adverse_weather_semantic_segmentation_robustness_benchmark_b200 as p
    rng = np.random.RandomState(70 + seed)
    b = int(rng.randint(1, 4))
    c = int(rng.choice([19, 19, 5, 33]))
    h = int(rng.choice([6, 17, 32, 50]))
    w = int(rng.choice([5, 16, 33, 64]))
    gen = torch.Generator().manual_seed(seed)
    scale = float(rng.choice([0.5, 2.0, 8.0]))
    base = str(rng.choice(["cross_entropy", "focal"]))
    use_fd, use_depth = bool(rng.rand() < 0.7), bool(rng.rand() < 0.6)
    logits = (torch.randn(b, c, h, w, generator=gen) * scale)
    label = torch.randint(0, c, (b, h, w), generator=gen)
    if rng.rand() < 0.5:
        label = label.to(torch.uint8)
    fd = torch.rand(b, h, w, generator=gen) if use_fd or not use_depth else None
    dpred = torch.rand(b, 1, h, w, generator=gen) * 30
    dtgt = torch.rand(b, h, w, generator=gen) * 30
    # reference arithmetic on the CPU (autograd)
    lg_r = logits.clone().requires_grad_(True)
    dp_r = dpred.clone().requires_grad_(True)
    pred_r = {"segmentation": lg_r}
    tgt_r = {"label": label}
    if use_depth:
        pred_r["depth"] = dp_r
        tgt_r["depth"] = dtgt
    want = ol.fog_loss(pred_r, tgt_r, fd, base_loss=base)
    want["total_loss"].backward()
    # product
    lg = logits.clone().cuda().requires_grad_(True)
    dp = dpred.clone().cuda().requires_grad_(True)
    pred = {"segmentation": lg}
    tgt = {"label": label.cuda()}
    if use_depth:
        pred["depth"] = dp
        tgt["depth"] = dtgt.cuda()
    got = p.FogDensityAwareLoss(base_loss=base)(pred, tgt, None if fd is None else fd.cuda())
    got["total_loss"].backward()
    for k in ("total_loss", "segmentation_loss"):
        np.testing.assert_allclose(float(got[k].detach()), float(want[k].detach()), rtol=2e-5)
    gd = lg.grad.cpu() - lg_r.grad
    assert float((gd.abs() - (2e-5 * lg_r.grad.abs() + 1e-9)).max()) <= 0
    if use_depth:
        np.testing.assert_allclose(float(got["depth_loss"].detach()), float(want["depth_loss"].detach()), rtol=2e-5)
        assert float(((dp.grad.cpu() - dp_r.grad).abs() - (2e-5 * dp_r.grad.abs() + 1e-9)).max()) <= 0
