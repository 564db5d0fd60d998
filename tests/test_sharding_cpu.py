"""World-size-2 (gloo, CPU) test of the sharded sweep's host logic: ranks fill their per-condition
integer bins from disjoint frame shards, one all_reduce(SUM) merges them, every rank finalises
the reference's result dict; the merged bins and the results must equal the single-process ones
bit for bit (integer bins are order independent)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

CONDITIONS = ("clean", "fog", "rain")
C, NB = 19, 256


def _frames(n=6, h=24, w=32):
    gen = torch.Generator().manual_seed(123)
    out = []
    for i in range(n):
        la = torch.randn(1, C, h, w, generator=gen)
        lb = torch.randn(1, C, h, w, generator=gen)
        tgt = torch.randint(0, C, (1, h, w), generator=gen).to(torch.uint8)
        out.append((CONDITIONS[i % len(CONDITIONS)], la, lb, tgt))
    return out


def _bins_from_oracle(ev, frames):
    """Fill the evaluator's bins on the CPU from oracle maps (stands in for awx_score launches)."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    from oracle import metrics as om, fusion as of_
    lay = ev.layout
    for cond, la, lb, tgt in frames:
        row = ev.bins[ev.conditions.index(cond)].numpy()
        fused = of_.fuse_logits(la, lb, "weighted_average", torch.tensor([0.3, 0.9]), torch.tensor([1.7]))
        row[lay.confusion:lay.confusion + C * C] += om.confusion_matrix(fused, tgt, C).numpy().reshape(-1)
        e = om.ece(fused, tgt)
        row[lay.ece_count:lay.ece_count + 15] += e["count"]
        row[lay.ece_correct:lay.ece_correct + 15] += e["correct"]
        conf, pred = om.confidence_and_prediction(fused)
        idx = om.ece_bin_index(conf.reshape(-1).numpy(), om.ece_edges(15).numpy())
        fx = (conf.reshape(-1).double().numpy() * 2.0 ** 31).astype(np.int64)
        for b in range(15):
            s = int(fx[idx == b].sum())
            row[lay.ece_conf_hi + b] += s >> 32
            row[lay.ece_conf_lo + b] += s & 0xffffffff
        mi = om.mi_map([la, lb]).numpy().reshape(-1)
        wrong = (om.mean_prob_prediction([la, lb]) != tgt).numpy().reshape(-1)
        mb = om.mi_bin_index(mi, NB, float(np.float32(np.log(2.0))))
        row[lay.auroc_pos:lay.auroc_pos + NB] += np.bincount(mb[wrong], minlength=NB)
        row[lay.auroc_neg:lay.auroc_neg + NB] += np.bincount(mb[~wrong], minlength=NB)
        row[lay.counters + _lib.CNT_VALID] += tgt.numel()
        row[lay.counters + _lib.CNT_PIXELS] += tgt.numel()
        row[lay.counters + _lib.CNT_CORRECT] += int((pred == tgt).sum())
        row[lay.counters + _lib.CNT_ENS_WRONG] += int(wrong.sum())


def _make_ev():
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation.streaming import StreamingEvaluator
    return StreamingEvaluator(C, CONDITIONS, 15, NB, "weighted_average", (0.3, 0.9), 1.7, bins_device="cpu")


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = _make_ev()
    _bins_from_oracle(ev, _frames()[rank::world])      # frames sharded rank::world
    ev.all_reduce()
    q.put((rank, ev.bins.clone().numpy(), ev.finalize()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_merge_equals_single_process():
    ev = _make_ev()
    _bins_from_oracle(ev, _frames())
    want_bins, want = ev.bins.numpy().copy(), ev.finalize()
    assert {"overall_miou", "miou_clean", "miou_fog", "ece_rain", "expected_calibration_error",
            "ensemble_disagreement_auroc", "robustness_degradation_fog", "robustness_degradation_ratio"} <= set(want)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bins, res in got:
        assert np.array_equal(bins, want_bins), f"rank {rank}: merged bins differ from the single-process bins"
        assert res.keys() == want.keys()
        for k in want:
            assert res[k] == want[k], (k, res[k], want[k])


def test_finalize_matches_oracle_end_to_end():
    """Merged bins -> the reference's evaluate_model numbers (overall mIoU, per-condition ECE, AUROC)."""
    from oracle import metrics as om, fusion as of_
    ev = _make_ev()
    frames = _frames()
    _bins_from_oracle(ev, frames)
    res = ev.finalize()
    raw_w, temp = torch.tensor([0.3, 0.9]), torch.tensor([1.7])
    fused = torch.cat([of_.fuse_logits(la, lb, "weighted_average", raw_w, temp) for _, la, lb, _ in frames])
    tg = torch.cat([t for *_, t in frames])
    assert res["overall_miou"] == om.iou(fused, tg, C)["mean_iou"]
    np.testing.assert_allclose(res["expected_calibration_error"], om.ece(fused, tg)["ece"], rtol=1e-5, atol=1e-8)
    fog = [f for f in frames if f[0] == "fog"]
    f_fog = torch.cat([of_.fuse_logits(la, lb, "weighted_average", raw_w, temp) for _, la, lb, _ in fog])
    t_fog = torch.cat([t for *_, t in fog])
    assert res["miou_fog"] == om.iou(f_fog, t_fog, C)["mean_iou"]
    np.testing.assert_allclose(res["ece_fog"], om.ece(f_fog, t_fog)["ece"], rtol=1e-5, atol=1e-8)
    exact = om.disagreement_auroc([torch.cat([f[1] for f in frames]), torch.cat([f[2] for f in frames])], tg)
    pc = ev.per_condition()
    assert abs(res["ensemble_disagreement_auroc"] - exact) <= 0.01   # 256 bins: loose by construction
    clean = om.iou(torch.cat([of_.fuse_logits(la, lb, "weighted_average", raw_w, temp) for c, la, lb, _ in frames if c == "clean"]),
                   torch.cat([t for c, *_, t in frames if c == "clean"]), C)["mean_iou"]
    assert res["robustness_degradation_fog"] == om.degradation_ratio(clean, res["miou_fog"])
    assert set(pc) == set(CONDITIONS)
