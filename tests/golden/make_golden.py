"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE in the build container.

    python tests/golden/make_golden.py

The reference holds no golden vectors for this path, so its own
outputs on seeded inputs are the pin.  Inputs are stored next to outputs so the files
are self-contained; library versions are stored in ``versions``.  Nothing here runs on
the GPU box.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refshim  # noqa: E402


def versions() -> str:
    import cv2, scipy, sklearn
    return json.dumps({"numpy": np.__version__, "cv2": cv2.__version__, "scipy": scipy.__version__,
                       "torch": torch.__version__, "sklearn": sklearn.__version__})


def weather_cases():
    pre = refshim.preprocessing()
    out = {"versions": versions()}
    shapes = {"s": (40, 56), "m": (96, 160)}
    for tag, (h, w) in shapes.items():
        np.random.seed(7)
        image = np.random.randint(0, 255, (h, w, 3), dtype=np.uint8)
        out[f"{tag}_image"] = image
        for kind in ("fog", "rain", "snow", "night"):
            for seed, intensity in ((42, None), (43, 0.5), (44, 0.9)):
                t = pre.WeatherDegradationTransforms(seed=seed)
                res = t.apply_weather_effect(image.copy(), kind, intensity)
                out[f"{tag}_{kind}_seed{seed}"] = res
        t = pre.WeatherDegradationTransforms(seed=11)
        out[f"{tag}_depth_seed11"] = t._generate_synthetic_depth(h, w)
    # many short streaks hugging the borders (thick-line clipping)
    np.random.seed(3)
    image = np.random.randint(0, 255, (24, 32, 3), dtype=np.uint8)
    out["b_image"] = image
    for seed in (1, 2, 3):
        t = pre.WeatherDegradationTransforms(seed=seed)
        out[f"b_rain_seed{seed}"] = t.apply_weather_effect(image.copy(), "rain", 0.8)
        t = pre.WeatherDegradationTransforms(seed=seed)
        out[f"b_snow_seed{seed}"] = t.apply_weather_effect(image.copy(), "snow", 0.7)
    np.savez_compressed(os.path.join(HERE, "weather.npz"), **out)


def metric_cases():
    met = refshim.metrics()
    out = {"versions": versions()}
    torch.manual_seed(42)
    # the reference's own (unused) conftest fixture shape: tests/conftest.py:192-209
    cases = {
        "c5": (2, 5, 64, 128, torch.int64, 0.0),
        "c19_i64_ign": (1, 19, 48, 64, torch.int64, 0.02),
        "c19_u8": (2, 19, 32, 48, torch.uint8, 0.0),
    }
    for tag, (b, c, h, w, ldt, ign) in cases.items():
        la = torch.randn(b, c, h, w)
        lb = torch.randn(b, c, h, w) * 1.5 + 0.3 * la
        tgt = torch.randint(0, c, (b, h, w)).to(ldt)
        if ign > 0:
            tgt[torch.rand(b, h, w) < ign] = 255
        out[f"{tag}_la"] = la.numpy()
        out[f"{tag}_lb"] = lb.numpy()
        out[f"{tag}_target"] = tgt.numpy()
        iou_m = met.IoUMetrics(c)
        r = iou_m.compute_iou(la, tgt)
        out[f"{tag}_mean_iou"] = np.float64(r["mean_iou"])
        out[f"{tag}_per_class_iou"] = r["per_class_iou"]
        out[f"{tag}_valid_classes"] = r["valid_classes"]
        out[f"{tag}_pixel_accuracy"] = np.float64(iou_m.compute_pixel_accuracy(la, tgt))
        cal = met.ConfidenceCalibration()
        d = cal.compute_ece(la, tgt, return_details=True)
        out[f"{tag}_ece"] = np.float64(d["ece"])
        out[f"{tag}_ece_scalar"] = np.float64(cal.compute_ece(la, tgt))
        for key in ("accuracy", "confidence", "proportion", "error", "bin_lower", "bin_upper"):
            out[f"{tag}_ece_{key}"] = np.array([x[key] for x in d["bin_details"]], dtype=np.float64)
        out[f"{tag}_ece_overall_accuracy"] = np.float64(d["overall_accuracy"])
        out[f"{tag}_ece_overall_confidence"] = np.float64(d["overall_confidence"])
        rel = cal.compute_reliability_diagram_data(la, tgt)
        out[f"{tag}_rel_centers"] = rel["bin_centers"]
        ens = met.EnsembleDisagreementMetrics()
        out[f"{tag}_mi"] = ens.compute_disagreement_map([la, lb]).numpy()
        out[f"{tag}_var"] = ens.compute_variance_map([la, lb]).numpy()
        out[f"{tag}_js"] = ens.compute_jensen_shannon_divergence(la, lb).numpy()
        out[f"{tag}_auroc"] = np.float64(ens.compute_disagreement_auroc([la, lb], tgt))
        rob = met.RobustnessMetrics(num_classes=c)
        comp = rob.compute_comprehensive_metrics(la, tgt, [la, lb], "fog")
        out[f"{tag}_comp_keys"] = json.dumps(sorted(comp.keys()))
        out[f"{tag}_comp_vals"] = np.array([comp[k] for k in sorted(comp.keys())], dtype=np.float64)
        out[f"{tag}_temp_scaled"] = cal.temperature_scale(la, 1.7).numpy()
    rob = met.RobustnessMetrics()
    out["degr"] = np.array([rob.compute_robustness_degradation_ratio(a, b)
                            for a, b in ((0.5, 0.4), (0.0, 0.3), (0.4, 0.5), (0.78, 0.65))])
    summ = rob.create_robustness_summary({
        "clean": {"mean_iou": 0.5, "expected_calibration_error": 0.02, "ensemble_disagreement_auroc": 0.7},
        "fog": {"mean_iou": 0.3, "expected_calibration_error": 0.05},
        "night": {"mean_iou": 0.45, "expected_calibration_error": 0.03, "ensemble_disagreement_auroc": 0.6},
    })
    out["summary_keys"] = json.dumps(sorted(summ.keys()))
    out["summary_vals"] = np.array([summ[k] for k in sorted(summ.keys())], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)


def fusion_cases():
    out = {"versions": versions()}
    torch.manual_seed(5)
    b, c, h, w = 2, 19, 24, 40
    l1 = torch.randn(b, c, h, w) * 2
    l2 = torch.randn(b, c, h, w) * 2 + 0.5 * l1
    d1 = torch.rand(b, 1, h, w)
    d2 = torch.rand(b, 1, h, w)
    out["l1"], out["l2"], out["d1"], out["d2"] = l1.numpy(), l2.numpy(), d1.numpy(), d2.numpy()
    raw_w = torch.tensor([0.3, 0.9])
    temp = torch.tensor([1.7])
    out["raw_w"], out["temp"] = raw_w.numpy(), temp.numpy()
    for strat in ("weighted_average", "max_confidence", "mean_anything"):
        for ts in (True, False):
            ens = refshim.ensemble_with_fixed_members(l1, l2, strat, ts, raw_w, temp, d1, d2)
            with torch.no_grad():
                r = ens(torch.zeros(b, 3, h, w))
            tag = f"{strat}_{'T' if ts else 'noT'}"
            out[f"{tag}_seg"] = r["segmentation"].numpy()
            out[f"{tag}_depth"] = r["depth"].numpy()
            out[f"{tag}_dis"] = ens.get_ensemble_disagreement(torch.zeros(b, 3, h, w)).numpy()
    np.savez_compressed(os.path.join(HERE, "fusion.npz"), **out)


def loss_cases():
    m = refshim.model()
    out = {"versions": versions()}
    torch.manual_seed(9)
    b, c, h, w = 2, 19, 24, 40
    logits = torch.randn(b, c, h, w)
    depth = torch.rand(b, 1, h, w) * 50
    label = torch.randint(0, c, (b, h, w))
    dtgt = torch.rand(b, h, w) * 50
    fd = torch.rand(b, h, w)
    out.update(logits=logits.numpy(), depth=depth.numpy(), label=label.numpy(),
               dtgt=dtgt.numpy(), fd=fd.numpy())
    variants = {
        "ce_fd_depth": dict(base="cross_entropy", fd=True, dpred=True, dtgt=True),
        "ce_fd_nodepth": dict(base="cross_entropy", fd=True, dpred=False, dtgt=False),
        "ce_nofd_nodepth": dict(base="cross_entropy", fd=False, dpred=False, dtgt=False),
        "focal_fd_depth": dict(base="focal", fd=True, dpred=True, dtgt=True),
        "ce_pathB": dict(base="cross_entropy", fd=False, dpred=True, dtgt=True),
        "ce_fd_dpred_only": dict(base="cross_entropy", fd=True, dpred=True, dtgt=False),
    }
    for tag, v in variants.items():
        lg = logits.clone().requires_grad_(True)
        dp = depth.clone().requires_grad_(True)
        pred = {"segmentation": lg}
        tgt = {"label": label}
        if v["dpred"]:
            pred["depth"] = dp
        if v["dtgt"]:
            tgt["depth"] = dtgt
        fn = m.FogDensityAwareLoss(base_loss=v["base"])
        r = fn(pred, tgt, fd if v["fd"] else None)
        r["total_loss"].backward()
        out[f"{tag}_total"] = np.float64(r["total_loss"].item())
        out[f"{tag}_seg"] = np.float64(r["segmentation_loss"].item())
        dl = r["depth_loss"]
        out[f"{tag}_depthloss"] = np.float64(dl.item() if torch.is_tensor(dl) else dl)
        out[f"{tag}_dlogits"] = lg.grad.numpy()
        out[f"{tag}_ddepth"] = dp.grad.numpy() if dp.grad is not None else np.zeros(0, np.float32)
    out["fd_from_depth"] = m.FogDensityAwareLoss()._estimate_fog_density_from_depth(depth.squeeze(1)).numpy()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


def prep_cases():
    """fog-density map, depth estimation, style transfer, temperature grid."""
    pre = refshim.preprocessing()
    met = refshim.metrics()
    ldr = refshim.loader()
    out = {"versions": versions()}
    # W multiples of 16: no scalar tail in OpenCV's filter2D vector body (see csrc/scene.cu header)
    for tag, (h, w) in {"a": (64, 96), "b": (50, 70), "c": (33, 48)}.items():
        np.random.seed(21)
        image = np.random.randint(0, 255, (h, w, 3), dtype=np.uint8)
        out[f"{tag}_image"] = image
        t = pre.WeatherDegradationTransforms(seed=13)
        depth = t._generate_synthetic_depth(h, w)
        out[f"{tag}_depth"] = depth
        imgf = image.astype(np.float32) / 255.0
        out[f"{tag}_fogmap"] = t.get_fog_density_map(imgf, depth)
        t2 = pre.WeatherDegradationTransforms(seed=14)
        out[f"{tag}_fogmap_seed14"] = t2.get_fog_density_map(imgf)  # depth drawn from the global RNG
        out[f"{tag}_est_depth"] = pre.DepthEstimationPreprocessor().estimate_depth(image)
    # style transfer: every byte value in every channel, plus a random frame
    ramp = np.repeat(np.arange(256, dtype=np.uint8).reshape(16, 16, 1), 3, axis=2)
    np.random.seed(5)
    rnd = np.random.randint(0, 256, (40, 56, 3), dtype=np.uint8)
    out["style_ramp"], out["style_rnd"] = ramp, rnd
    pipe = ldr.WeatherAugmentationPipeline()
    for kind in ("fog", "rain", "snow", "night", "clean"):
        out[f"style_ramp_{kind}"] = pipe._apply_style_transfer(ramp.copy(), kind)
        out[f"style_rnd_{kind}"] = pipe._apply_style_transfer(rnd.copy(), kind)
    # full pipeline with the reference's RNG order (weather choice, corruption draws, style coin)
    np.random.seed(77)
    frame = np.random.randint(0, 255, (48, 64, 3), dtype=np.uint8)
    out["aug_frame"] = frame
    for seed in (1, 2, 3, 4):
        np.random.seed(seed)
        p2 = ldr.WeatherAugmentationPipeline(style_transfer_prob=0.6)
        np.random.seed(seed)
        out[f"aug_seed{seed}"] = p2.apply_domain_adaptation_augmentation(frame.copy())
    # temperature grid search
    torch.manual_seed(31)
    cal = met.ConfidenceCalibration()
    for tag, (b, c, h, w, scale, ign) in {"t19": (2, 19, 16, 24, 3.0, 0.05), "t5": (1, 5, 12, 20, 0.5, 0.0)}.items():
        logits = torch.randn(b, c, h, w) * scale
        targets = torch.randint(0, c, (b, h, w))
        if ign > 0:
            targets[torch.rand(b, h, w) < ign] = 255
        out[f"{tag}_logits"], out[f"{tag}_targets"] = logits.numpy(), targets.numpy()
        out[f"{tag}_best_t"] = np.float64(cal.optimize_temperature(logits, targets))
        rows = logits.view(-1, c)[targets.view(-1) != 255]
        tg = targets.view(-1)[targets.view(-1) != 255]
        out[f"{tag}_nll"] = np.array([torch.nn.functional.cross_entropy(rows / t, tg).item()
                                      for t in torch.linspace(0.1, 10.0, 100)], dtype=np.float64)
    # EnsembleDisagreementMetrics on LISTS of three and four members (the list form of metrics.py:336-438)
    ens = met.EnsembleDisagreementMetrics()
    torch.manual_seed(77)
    for tag, (n, b, c, h, w, ign) in {"m3": (3, 2, 19, 24, 32, 0.03), "m4": (4, 1, 5, 16, 20, 0.0)}.items():
        base = torch.randn(b, c, h, w) * 2
        members = [base * 0.5 + torch.randn(b, c, h, w) * (1.0 + 0.5 * k) for k in range(n)]
        targets = torch.randint(0, c, (b, h, w))
        if ign > 0:
            targets[torch.rand(b, h, w) < ign] = 255
        for k, m in enumerate(members):
            out[f"{tag}_member{k}"] = m.numpy()
        out[f"{tag}_targets"] = targets.numpy()
        out[f"{tag}_mi"] = ens.compute_disagreement_map(members).numpy()
        out[f"{tag}_var"] = ens.compute_variance_map(members).numpy()
        out[f"{tag}_auroc"] = np.float64(ens.compute_disagreement_auroc(members, targets))
    np.savez_compressed(os.path.join(HERE, "prep.npz"), **out)


def trainer_batches(seed, n_batches, bsz, c, h, w, with_depth, with_other):
    """Synthetic validation loader + per-batch model outputs (shared with the tests: everything a test needs is
    stored in the fixture, this only builds it)."""
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(seed)
    names = ["clean", "fog", "rain", "snow", "night"] + (["hail"] if with_other else [])
    batches = []
    for _ in range(n_batches):
        b = {"image": torch.rand(bsz, 3, h, w, generator=gen),
             "label": torch.randint(0, c, (bsz, h, w), generator=gen),
             "weather_condition": [names[i] for i in rng.randint(0, len(names), bsz)],
             "_seg": torch.randn(bsz, c, h, w, generator=gen) * 2}
        if with_depth:
            b["depth"] = torch.rand(bsz, h, w, generator=gen) * 50
            b["_depth"] = torch.rand(bsz, 1, h, w, generator=gen) * 50
        batches.append(b)
    return batches


def trainer_cases():
    """SURVEY.md 8f row 1, trainer twin: AdverseWeatherTrainer.validate_epoch / _estimate_fog_density run
    UNBOUND on a namespace carrying the members they read (model, val_loader, device, loss_fn, metrics)."""
    from types import SimpleNamespace
    tr = refshim.trainer()
    met = refshim.metrics()
    T = tr.AdverseWeatherTrainer
    out = {"versions": versions()}
    # fog-density maps: the global torch CPU generator, one rand(h, w) per frame
    names = ["fog", "clean", "rain", "snow", "night", "hail", "fog"]
    torch.manual_seed(123)
    fd = T._estimate_fog_density(None, {"weather_condition": names, "image": torch.zeros(len(names), 3, 20, 28)})
    out["fd_names"] = np.array(names)
    out["fd_maps"] = fd.numpy()
    assert T._estimate_fog_density(None, {"image": torch.zeros(1, 3, 4, 4)}) is None

    class _Replay(torch.nn.Module):
        def __init__(self, batches):
            super().__init__()
            self.batches, self.i = batches, 0

        def forward(self, x):
            b = self.batches[self.i]
            self.i += 1
            o = {"segmentation": b["_seg"]}
            if "_depth" in b:
                o["depth"] = b["_depth"]
            return o

    cases = {"v_depth": (5, 3, 3, 19, 24, 32, True, True, "cross_entropy"),
             "v_nodepth": (6, 2, 4, 19, 16, 48, False, False, "cross_entropy"),
             "v_focal": (7, 2, 2, 7, 20, 20, True, True, "focal")}
    for tag, (seed, nb, bsz, c, h, w, with_depth, with_other, base) in cases.items():
        batches = trainer_batches(seed, nb, bsz, c, h, w, with_depth, with_other)
        ns = SimpleNamespace(model=_Replay(batches), val_loader=batches, device=torch.device("cpu"),
                             loss_fn=tr.FogDensityAwareLoss(base_loss=base), metrics=met.RobustnessMetrics(c))
        ns._estimate_fog_density = lambda b: T._estimate_fog_density(ns, b)
        torch.manual_seed(1000 + seed)
        res = T.validate_epoch(ns)
        out[f"{tag}_args"] = np.array([seed, nb, bsz, c, h, w, int(with_depth), int(with_other)])
        out[f"{tag}_base"] = np.array(base)
        out[f"{tag}_keys"] = np.array(sorted(res))
        out[f"{tag}_vals"] = np.array([float(res[k]) for k in sorted(res)], dtype=np.float64)
        for i, b in enumerate(batches):
            for k, v in b.items():
                out[f"{tag}_b{i}_{k}"] = np.array(v) if k == "weather_condition" else v.numpy()
    # plain criterion (the trainer's non-fog-aware branch, trainer.py:431-434)
    batches = trainer_batches(9, 2, 3, 19, 16, 24, False, False)
    ns = SimpleNamespace(model=_Replay(batches), val_loader=batches, device=torch.device("cpu"),
                         loss_fn=torch.nn.CrossEntropyLoss(), metrics=met.RobustnessMetrics(19))
    res = T.validate_epoch(ns)
    out["v_plain_args"] = np.array([9, 2, 3, 19, 16, 24, 0, 0])
    out["v_plain_keys"] = np.array(sorted(res))
    out["v_plain_vals"] = np.array([float(res[k]) for k in sorted(res)], dtype=np.float64)
    for i, b in enumerate(batches):
        for k, v in b.items():
            out[f"v_plain_b{i}_{k}"] = np.array(v) if k == "weather_condition" else v.numpy()
    np.savez_compressed(os.path.join(HERE, "trainer.npz"), **out)


def evaluate_batches(seed, n_batches, bsz, c, h, w, label_dtype, with_other):
    """Synthetic test loader with the member logits of every batch stashed next to it."""
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(seed)
    names = ["clean", "fog", "rain", "snow", "night"] + (["hail"] if with_other else [])
    batches = []
    for _ in range(n_batches):
        lab = torch.randint(0, c, (bsz, h, w), generator=gen)
        lab[torch.rand(bsz, h, w, generator=gen) < 0.03] = 255
        batches.append({"image": torch.zeros(bsz, 3, 4, 4),
                        "label": lab.to(label_dtype),
                        "weather_condition": [names[i] for i in rng.randint(0, len(names), bsz)],
                        "_la": torch.randn(bsz, c, h, w, generator=gen) * 2,
                        "_lb": torch.randn(bsz, c, h, w, generator=gen) * 2})
    batches[0]["weather_condition"][0] = "clean"   # the degradation ratios need a clean frame
    return batches


def evaluate_cases():
    """the reference's OWN evaluate_model (scripts/evaluate.py:134-274) run on its OWN
    EnsembleModel (trained-looking fusion weights, temperature != 1) with two replaying members injected, on
    loaders with int64 / uint8 labels, mixed weather and a condition the config does not list; plus the same
    model behind a DistributedDataParallel-style wrapper (the reference then finds no `segformer` attribute and
    reports no ensemble AUROC) and a single-member model."""
    ev = refshim.evaluate_script()
    met = refshim.metrics()
    out = {"versions": versions()}

    class _Member(torch.nn.Module):
        def __init__(self, batches, key):
            super().__init__()
            self.batches, self.key, self.i = batches, key, 0

        def forward(self, x):
            b = self.batches[self.i]
            self.i += 1
            return {"segmentation": b[self.key]}

    class _Wrapper(torch.nn.Module):      # what DistributedDataParallel looks like from outside
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, x):
            return self.module(x)

    class _Config:
        def __init__(self, d):
            self.d = d

        def get(self, k, default=None):
            return self.d.get(k, default)

    conds = ["clean", "fog", "rain", "snow", "night"]
    cases = {  # tag: (seed, batches, batch size, C, H, W, label dtype, other condition, strategy, T scaling, wrapped, single)
        "e_weighted": (11, 4, 3, 19, 12, 16, torch.int64, True, "weighted_average", True, False, False),
        "e_u8": (12, 3, 4, 19, 8, 20, torch.uint8, True, "weighted_average", True, False, False),
        "e_maxconf": (13, 3, 3, 19, 12, 16, torch.int64, False, "max_confidence", True, False, False),
        "e_mean_not": (14, 3, 3, 19, 12, 16, torch.int64, True, "mean", False, False, False),
        "e_wrapped": (15, 3, 3, 19, 12, 16, torch.int64, True, "weighted_average", True, True, False),
        "e_single": (16, 3, 3, 19, 12, 16, torch.int64, False, None, False, False, True),
    }
    for tag, (seed, nb, bsz, c, h, w, ldt, other, strategy, ts, wrapped, single) in cases.items():
        batches = evaluate_batches(seed, nb, bsz, c, h, w, ldt, other)
        if single:
            model = _Member(batches, "_la")
        else:
            model = refshim.ensemble_with_fixed_members(batches[0]["_la"], batches[0]["_lb"], strategy, ts,
                                                        raw_weights=torch.tensor([0.3, 0.9]),
                                                        temperature=torch.tensor([1.7]))
            model.segformer, model.deeplabv3plus = _Member(batches, "_la"), _Member(batches, "_lb")
            if wrapped:
                model = _Wrapper(model)
        res = ev.evaluate_model(model, batches, met.RobustnessMetrics(c), torch.device("cpu"),
                                _Config({"data.weather_conditions": conds}))
        out[f"{tag}_args"] = np.array([seed, nb, bsz, c, h, w, int(other), int(ts), int(wrapped), int(single)])
        out[f"{tag}_strategy"] = np.array(str(strategy))
        out[f"{tag}_keys"] = np.array(sorted(res))
        out[f"{tag}_vals"] = np.array([float(res[k]) for k in sorted(res)], dtype=np.float64)
        for i, b in enumerate(batches):
            for k, v in b.items():
                if k != "image":
                    out[f"{tag}_b{i}_{k}"] = np.array(v) if k == "weather_condition" else v.numpy()
    np.savez_compressed(os.path.join(HERE, "evaluate.npz"), **out)


if __name__ == "__main__":
    if not refshim.available():
        raise SystemExit("reference not present; golden files can only be regenerated in the build container")
    weather_cases()
    metric_cases()
    fusion_cases()
    loss_cases()
    prep_cases()
    trainer_cases()
    evaluate_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
