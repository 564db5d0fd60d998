"""GPU parity: awx_fogloss / FogDensityAwareLoss / EnsembleModel against golden vectors of the reference.

Bars: loss values rel 1e-5; gradients |d| <= 1e-5*|ref| + 1e-9 (they carry a 1/N factor);
fused logits bit-exact; disagreement map |d| <= 1e-5*|ref| + 2e-6.
"""

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import adverse_weather_semantic_segmentation_robustness_benchmark_b200 as p
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    _lib.load()
    return p


VARIANTS = {
    "ce_fd_depth": ("cross_entropy", True, True, True),
    "ce_fd_nodepth": ("cross_entropy", True, False, False),
    "ce_nofd_nodepth": ("cross_entropy", False, False, False),
    "focal_fd_depth": ("focal", True, True, True),
    "ce_pathB": ("cross_entropy", False, True, True),
    "ce_fd_dpred_only": ("cross_entropy", True, True, False),
}


def _close(got, want, rtol, atol):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    ex = np.abs(got - want) - (atol + rtol * np.abs(want))
    assert ex.max() <= 0, f"max abs diff {np.abs(got - want).max():.3e}, excess {ex.max():.3e}"


@pytest.mark.parametrize("tag", sorted(VARIANTS))
@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_loss_forward_backward_golden(pkg, golden, tag, device):
    g = golden("loss")
    base, use_fd, dpred, dtgt = VARIANTS[tag]
    lg = torch.from_numpy(g["logits"]).to(device).requires_grad_(True)
    dp = torch.from_numpy(g["depth"]).to(device).requires_grad_(True)
    pred = {"segmentation": lg}
    tgt = {"label": torch.from_numpy(g["label"]).to(device)}
    if dpred:
        pred["depth"] = dp
    if dtgt:
        tgt["depth"] = torch.from_numpy(g["dtgt"]).to(device)
    fn = pkg.FogDensityAwareLoss(base_loss=base)
    r = fn(pred, tgt, torch.from_numpy(g["fd"]).to(device) if use_fd else None)
    assert set(r) == {"total_loss", "segmentation_loss", "depth_loss"}
    r["total_loss"].backward()
    _close(r["total_loss"].item(), g[f"{tag}_total"], 1e-5, 0)
    _close(r["segmentation_loss"].item(), g[f"{tag}_seg"], 1e-5, 0)
    dl = r["depth_loss"]
    _close(dl.item() if torch.is_tensor(dl) else dl, g[f"{tag}_depthloss"], 1e-5, 0)
    assert lg.grad.device.type == device
    _close(lg.grad.cpu().numpy(), g[f"{tag}_dlogits"], 1e-5, 1e-9)
    want = g[f"{tag}_ddepth"]
    if want.size:
        _close(dp.grad.cpu().numpy(), want, 2e-5, 1e-9)
    else:
        assert dp.grad is None


def test_loss_bad_label_raises(pkg):
    fn = pkg.FogDensityAwareLoss()
    lg = torch.zeros(1, 19, 4, 4)
    with pytest.raises(IndexError):
        fn({"segmentation": lg}, {"label": torch.full((1, 4, 4), 255)})


def test_loss_uint8_labels_and_odd_sizes(pkg):
    from oracle import loss as ol
    torch.manual_seed(0)
    for c, h, w in ((19, 33, 35), (5, 16, 24), (19, 64, 64)):
        lg = torch.randn(2, c, h, w, requires_grad=True)
        lab = torch.randint(0, c, (2, h, w)).to(torch.uint8)
        fd = torch.rand(2, h, w)
        want = ol.fog_loss({"segmentation": lg}, {"label": lab}, fd)
        want["total_loss"].backward()
        gref = lg.grad.clone()
        lg.grad = None
        got = pkg.FogDensityAwareLoss()({"segmentation": lg}, {"label": lab}, fd)
        got["total_loss"].backward()
        _close(got["total_loss"].item(), want["total_loss"].item(), 1e-5, 0)
        _close(lg.grad.numpy(), gref.numpy(), 1e-5, 1e-9)
        lg.grad = None


def test_loss_deterministic(pkg):
    torch.manual_seed(1)
    lg = torch.randn(4, 19, 128, 128, device="cuda")
    lab = torch.randint(0, 19, (4, 128, 128), device="cuda")
    fd = torch.rand(4, 128, 128, device="cuda")
    fn = pkg.FogDensityAwareLoss()
    a = fn({"segmentation": lg}, {"label": lab}, fd)["total_loss"]
    b = fn({"segmentation": lg}, {"label": lab}, fd)["total_loss"]
    assert a.item() == b.item()


class _Fixed(nn.Module):
    def __init__(self, seg, depth=None):
        super().__init__()
        self.seg, self.depth = seg, depth

    def forward(self, x):
        out = {"segmentation": self.seg}
        if self.depth is not None:
            out["depth"] = self.depth
        return out


@pytest.mark.parametrize("strategy", ["weighted_average", "max_confidence", "mean_anything"])
@pytest.mark.parametrize("ts", [True, False])
def test_ensemble_model_golden(pkg, golden, strategy, ts):
    g = golden("fusion")
    l1, l2 = torch.from_numpy(g["l1"]), torch.from_numpy(g["l2"])
    d1, d2 = torch.from_numpy(g["d1"]), torch.from_numpy(g["d2"])
    ens = pkg.EnsembleModel(19, True, strategy, ts, segformer=_Fixed(l1, d1), deeplabv3plus=_Fixed(l2, d2))
    with torch.no_grad():
        ens.ensemble_weights.copy_(torch.from_numpy(g["raw_w"]))
        if ts:
            ens.temperature.copy_(torch.from_numpy(g["temp"]))
        r = ens(torch.zeros(2, 3, 24, 40))
    tag = f"{strategy}_{'T' if ts else 'noT'}"
    keys = {"segmentation", "segformer_seg", "deeplabv3plus_seg", "depth", "segformer_depth", "deeplabv3plus_depth"}
    assert set(r) == keys
    assert torch.equal(r["segmentation"].cpu(), torch.from_numpy(g[f"{tag}_seg"])), "fused logits must be bit-exact"
    assert torch.equal(r["depth"].cpu(), torch.from_numpy(g[f"{tag}_depth"]))
    dis = ens.get_ensemble_disagreement(torch.zeros(2, 3, 24, 40))
    assert dis.shape == (2, 24, 40) and float(dis.min()) >= -1e-6
    _close(dis.cpu().numpy(), g[f"{tag}_dis"], 1e-5, 2e-6)
    assert isinstance(ens.temperature if ts else ens.ensemble_weights, nn.Parameter)


def test_ensemble_training_gradients(pkg):
    """Differentiable fusion + loss: gradients w.r.t. members, fusion weights and temperature
    against torch autograd of the reference's expressions."""
    from oracle import fusion as of_, loss as ol
    torch.manual_seed(3)
    b, c, h, w = 2, 19, 16, 24
    a0 = torch.randn(b, c, h, w)
    b0 = torch.randn(b, c, h, w)
    lab = torch.randint(0, c, (b, h, w))
    fd = torch.rand(b, h, w)

    def run(ours):
        a = a0.clone().requires_grad_(True)
        bb = b0.clone().requires_grad_(True)
        if ours:
            ens = pkg.EnsembleModel(c, False, "weighted_average", True, segformer=_Fixed(a), deeplabv3plus=_Fixed(bb))
            with torch.no_grad():
                ens.ensemble_weights.copy_(torch.tensor([0.3, 0.9]))
                ens.temperature.copy_(torch.tensor([1.7]))
            ens.train()
            out = ens(torch.zeros(b, 3, h, w))
            loss = pkg.FogDensityAwareLoss()(out, {"label": lab}, fd)["total_loss"]
            loss.backward()
            return loss.item(), a.grad, bb.grad, ens.ensemble_weights.grad, ens.temperature.grad
        rw = torch.tensor([0.3, 0.9], requires_grad=True)
        t = torch.tensor([1.7], requires_grad=True)
        fused = of_.fuse_logits(a, bb, "weighted_average", rw, t)
        loss = ol.fog_loss({"segmentation": fused}, {"label": lab}, fd)["total_loss"]
        loss.backward()
        return loss.item(), a.grad, bb.grad, rw.grad, t.grad

    got, want = run(True), run(False)
    _close(got[0], want[0], 1e-5, 0)
    for gg, ww in zip(got[1:], want[1:]):
        _close(gg.cpu().numpy(), ww.numpy(), 2e-4, 1e-8)


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_temperature_only_gradient_with_frozen_members(pkg, mode):
    """Post-hoc calibration: frozen (detached) members, frozen fusion weights, trainable temperature, in eval() mode
    as well as train().  The reference's eager expression always carries autograd to `temperature`
    (models/model.py:443-462); the drop-in must too, with the same value."""
    from oracle import fusion as of_
    torch.manual_seed(5)
    b, c, h, w = 2, 19, 12, 20
    a0, b0 = torch.randn(b, c, h, w), torch.randn(b, c, h, w)
    lab = torch.randint(0, c, (b, h, w))
    ens = pkg.EnsembleModel(c, False, "weighted_average", True, segformer=_Fixed(a0), deeplabv3plus=_Fixed(b0))
    with torch.no_grad():
        ens.ensemble_weights.copy_(torch.tensor([0.3, 0.9]))
        ens.temperature.copy_(torch.tensor([1.7]))
    ens.ensemble_weights.requires_grad_(False)
    getattr(ens, mode)()
    out = ens(torch.zeros(b, 3, h, w))["segmentation"]
    assert out.requires_grad, "the fused logits must carry a graph to the temperature"
    torch.nn.functional.cross_entropy(out, lab.to(out.device)).backward()
    assert ens.ensemble_weights.grad is None
    t = torch.tensor([1.7], requires_grad=True)
    ref = of_.fuse_logits(a0, b0, "weighted_average", torch.tensor([0.3, 0.9]), t)
    torch.nn.functional.cross_entropy(ref, lab).backward()
    _close(ens.temperature.grad.cpu().numpy(), t.grad.numpy(), 2e-4, 1e-8)
    # ConfidenceCalibration.temperature_scale is the reference's `logits / temperature`: differentiable as well
    cal = pkg.ConfidenceCalibration()
    t2 = torch.tensor([2.5], requires_grad=True)
    x = a0.clone().requires_grad_(True)
    y = cal.temperature_scale(x, t2)
    assert torch.equal(y.detach().cpu(), a0 / torch.tensor([2.5]))
    (y * y).sum().backward()
    t3 = torch.tensor([2.5], requires_grad=True)
    x3 = a0.clone().requires_grad_(True)
    ((x3 / t3) ** 2).sum().backward()
    _close(t2.grad.cpu().numpy(), t3.grad.numpy(), 2e-4, 1e-6)
    _close(x.grad.cpu().numpy(), x3.grad.numpy(), 1e-5, 1e-8)
    # without grad mode (or nothing trainable) the plain kernel path is taken
    with torch.no_grad():
        assert not ens(torch.zeros(b, 3, h, w))["segmentation"].requires_grad


def test_ops_follow_the_device_of_their_operands(pkg):
    """Tensors on cuda:1 while cuda:0 is current: the wrappers must switch device (outputs, bins and workspaces on
    cuda:1, kernels launched there) instead of launching on device 0 with device-1 pointers."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib
    torch.cuda.set_device(0)
    d1 = torch.device("cuda", 1)
    gen = torch.Generator().manual_seed(2)
    la, lb = torch.randn(1, 19, 16, 32, generator=gen), torch.randn(1, 19, 16, 32, generator=gen)
    tgt = torch.randint(0, 19, (1, 16, 32), generator=gen)
    kw = dict(strategy=_lib.FUSE_WEIGHTED, w0=0.4, w1=0.6, temperature=1.3, auroc_bins=1024)
    want = ops.score(la, lb, tgt, **kw)["bins"].cpu()                      # on cuda:0
    got = ops.score(la.to(d1), lb.to(d1), tgt, **kw)["bins"]              # members on cuda:1, labels on the host
    assert got.device == d1 and torch.equal(got.cpu(), want)
    assert torch.cuda.current_device() == 0
    with pytest.raises(ValueError, match="different CUDA devices"):
        ops.score(la.to(d1), lb.cuda(0), tgt, **kw)
    out = pkg.FogDensityAwareLoss()({"segmentation": la.to(d1).requires_grad_(True)}, {"label": tgt.to(d1)})
    assert out["total_loss"].device == d1


def test_depth_density_forward_backward(pkg, golden):
    """_estimate_fog_density_from_depth in libawx: forward against the reference's golden output, gradient
    against autograd of the oracle's restatement (the indicator edge mask carries no gradient; min / max do)."""
    from oracle import loss as ol
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_loss
    g = golden("loss")
    depth = torch.from_numpy(g["depth"]).squeeze(1)
    got = pkg.FogDensityAwareLoss()._estimate_fog_density_from_depth(depth)
    _close(got.cpu().numpy(), g["fd_from_depth"], 1e-6, 1e-7)
    gen = torch.Generator().manual_seed(3)
    for shape in ((2, 24, 40), (1, 2, 2), (3, 17, 33)):
        d = (torch.rand(*shape, generator=gen) * 40).requires_grad_(True)
        up = torch.randn(*shape, generator=gen)
        ref = ol.fog_density_from_depth(d)
        (ref * up).sum().backward()
        dd = d.detach().clone().cuda().requires_grad_(True)
        out = ops_loss.depth_density(dd)
        (out * up.cuda()).sum().backward()
        _close(out.detach().cpu().numpy(), ref.detach().numpy(), 1e-6, 1e-7)
        _close(dd.grad.cpu().numpy(), d.grad.numpy(), 2e-5, 1e-7)
    # ties at the extrema share the min / max gradient evenly, as torch's full-reduction backward does
    d = torch.tensor([[[0.0, 5.0, 5.0], [0.0, 2.0, 3.0], [1.0, 4.0, 5.0]]], requires_grad=True)
    up = torch.arange(9.0).reshape(1, 3, 3) + 1.0
    (ol.fog_density_from_depth(d) * up).sum().backward()
    dd = d.detach().clone().cuda().requires_grad_(True)
    (ops_loss.depth_density(dd) * up.cuda()).sum().backward()
    _close(dd.grad.cpu().numpy(), d.grad.numpy(), 2e-5, 1e-7)


@pytest.mark.parametrize("strategy", ["weighted_average", "mean", "max_confidence"])
@pytest.mark.parametrize("ts", [True, False])
def test_fusion_backward_kernel(pkg, strategy, ts):
    """awx_fuse_backward (with depth heads: C = 1 through the same kernel) against torch autograd of the
    reference's fusion expressions, for every strategy, with and without temperature scaling."""
    from oracle import fusion as of_
    torch.manual_seed(11)
    b, c, h, w = 2, 19, 12, 20
    a0, b0 = torch.randn(b, c, h, w) * 2, torch.randn(b, c, h, w) * 2
    d10, d20 = torch.rand(b, 1, h, w), torch.rand(b, 1, h, w)
    up, upd = torch.randn(b, c, h, w), torch.randn(b, 1, h, w)

    a, bb = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    d1, d2 = d10.clone().requires_grad_(True), d20.clone().requires_grad_(True)
    ens = pkg.EnsembleModel(c, True, strategy, ts, segformer=_Fixed(a, d1), deeplabv3plus=_Fixed(bb, d2))
    with torch.no_grad():
        ens.ensemble_weights.copy_(torch.tensor([0.3, 0.9]))
        if ts:
            ens.temperature.copy_(torch.tensor([1.7]))
    ens.train()
    out = ens(torch.zeros(b, 3, h, w))
    ((out["segmentation"] * up.cuda()).sum() + (out["depth"] * upd.cuda()).sum()).backward()

    ar, br = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    d1r, d2r = d10.clone().requires_grad_(True), d20.clone().requires_grad_(True)
    rw = torch.tensor([0.3, 0.9], requires_grad=True)
    t = torch.tensor([1.7], requires_grad=True)
    fused = of_.fuse_logits(ar, br, strategy, rw, t if ts else None)
    depth = of_.fuse_depth(d1r, d2r, strategy, rw)
    ((fused * up).sum() + (depth * upd).sum()).backward()

    assert torch.equal(out["segmentation"].detach().cpu(), fused.detach())
    for got, want in ((a.grad, ar.grad), (bb.grad, br.grad), (d1.grad, d1r.grad), (d2.grad, d2r.grad)):
        _close(got.cpu().numpy(), want.numpy(), 1e-5, 1e-7)
    if strategy == "weighted_average":
        _close(ens.ensemble_weights.grad.cpu().numpy(), rw.grad.numpy(), 1e-4, 1e-5)
    if ts:
        _close(ens.temperature.grad.cpu().numpy(), t.grad.numpy(), 1e-4, 1e-5)


def test_ring_kernel_matches_register_kernel(monkeypatch):
    """awx_fogloss's TMA-staged kernel (C = 19, aligned planes) against the register kernel (AWX_LOSS_KERNEL=v1) on
    frames with a tail tile per image (H*W = 1 200 is not a multiple of the 608-pixel tile), labels of both dtypes
    and an out-of-range label: same sums, same gradients, same bad-label count."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_loss
    gen = torch.Generator().manual_seed(3)
    b, c, h, w = 3, 19, 30, 40
    logits = (torch.randn(b, c, h, w, generator=gen) * 3).cuda()
    fog = torch.rand(b, h, w, generator=gen).cuda()
    dp, dt = (torch.rand(b, h, w, generator=gen) * 40).cuda(), (torch.rand(b, h, w, generator=gen) * 40).cuda()
    for dtype in (torch.int64, torch.uint8):
        labels = torch.randint(0, c, (b, h, w), generator=gen).to(dtype).cuda()
        labels[1, 2, 3] = 200
        for focal in (False, True):
            outs = {}
            for kern in ("ring", "v1"):
                if kern == "v1":
                    monkeypatch.setenv("AWX_LOSS_KERNEL", "v1")
                else:
                    monkeypatch.delenv("AWX_LOSS_KERNEL", raising=False)
                outs[kern] = ops_loss.fogloss_raw(logits, labels, fog, dp, dt, 2.0, focal, True, True)
            ring, v1 = outs["ring"], outs["v1"]
            assert len(ring) == len(v1)
            for a, r in zip(ring, v1):
                if isinstance(a, torch.Tensor):
                    if a.dtype.is_floating_point:
                        torch.testing.assert_close(a, r, rtol=1e-6, atol=1e-12)
                    else:
                        assert torch.equal(a, r)
                else:
                    assert a == r
