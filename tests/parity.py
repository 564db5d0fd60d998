"""Integer-parity gates shared by the GPU tests (and mirrored by bench.py's `parity` object).

The reference derives three integer quantities from fp32 PROBABILITIES (torch's softmax), which no other
implementation can reproduce bit for bit: the ECE bin of a pixel, the arg-max of the fused probabilities (the
ECE's accuracy term) and the arg-max of the mean member probabilities (the AUROC's positive flag).  The CUDA
path re-evaluates every pixel that is close to such a decision boundary in fp64 and COUNTS those for which the
outcome is inside the reference's own rounding noise (AWX_CNT_ECE_AMBIG / _EPRED_AMBIG / _MARG_AMBIG).  The
gates below make those counters the only slack:

* the counters themselves are bounded (a kernel that flags half the frame cannot pass);
* the TRUE mismatch against the oracle -- pixels sitting in a different ECE bin, `correct` flags that differ,
  ensemble-wrong flags that differ -- must not exceed the corresponding counter.
"""

import numpy as np


def amb_bound(n_valid: int, per_pixel: float = 2e-6, floor: int = 0) -> int:
    """Upper bound on a self-reported ambiguity counter: `per_pixel` of the valid pixels (+ `floor` for the
    small frames of the unit tests, where one unlucky pixel is already more than 2e-6 of the frame)."""
    return int(np.floor(per_pixel * n_valid)) + floor


def ece_mismatch(bins, ref: dict) -> dict:
    """True mismatch of the ECE bins against the oracle's integer counts."""
    d_count = np.abs(np.asarray(bins.ece_count) - ref["count"])
    d_correct = np.abs(np.asarray(bins.ece_correct) - ref["correct"])
    # a pixel binned differently shows up in two bins; an odd total means the in-no-bin count differs too
    return {"moved_pixels": int((d_count.sum() + 1) // 2), "correct_l1": int(d_correct.sum())}


def genuine_ece_near_edges(logits, target, num_bins: int = 15, ulps: float = 4.0) -> int:
    """Valid pixels whose fp64 confidence lies within `ulps` fp32 ulp of an interior ECE bin edge: the pixels for
    which the reference's own fp32 softmax decides the bin (the kernel's criterion is 3 ulp).  This count -- a
    property of the data, ~1.5e-7 per edge and unit of confidence density -- is the cap on AWX_CNT_ECE_AMBIG."""
    import torch
    import torch.nn.functional as F
    conf = F.softmax(logits.double(), dim=1).max(dim=1).values
    c32 = conf.float()
    ulp = (torch.nextafter(c32, torch.full_like(c32, 2.0)) - c32).double()
    edges = torch.linspace(0, 1, num_bins + 1)[1:-1].double()
    near = torch.zeros_like(conf, dtype=torch.bool)
    for e in edges:
        near |= (conf - e).abs() <= ulps * ulp
    return int((near & (target != 255)).sum())


def assert_ece_parity(bins, ref: dict, n_valid: int, _lib, per_pixel: float = 2e-6, floor: int = 0,
                      extra_amb: int = 0, cap: int = None) -> dict:
    """ECE counts == the oracle's up to the pixels the kernel itself reports as ambiguous, those being few
    (`cap`: genuine_ece_near_edges of the same data, or `per_pixel` of the frame + `floor`)."""
    amb = bins.counter(_lib.CNT_ECE_AMBIG) + extra_amb
    eamb = bins.counter(_lib.CNT_EPRED_AMBIG)
    if cap is None:
        cap = amb_bound(n_valid, per_pixel, floor)
    assert amb <= cap + extra_amb, f"{amb} ECE-ambiguous pixels reported, more than {cap} ({per_pixel:g} of {n_valid})"
    assert eamb <= cap, f"{eamb} prediction-ambiguous pixels reported, more than {cap}"
    mm = ece_mismatch(bins, ref)
    assert mm["moved_pixels"] <= amb, f"{mm['moved_pixels']} pixels binned differently, only {amb} reported ambiguous"
    # a moved pixel that is correct changes two `correct` bins; a prediction-ambiguous pixel changes one
    assert mm["correct_l1"] <= 2 * amb + eamb, (mm, amb, eamb)
    mm.update(ece_ambiguous=amb, epred_ambiguous=eamb)
    return mm


def genuine_marg_ties(members, target, rel: float = 1e-6) -> int:
    """Valid pixels whose label is one of SEVERAL classes within `rel` of the maximum of the mean member
    probabilities, evaluated in fp64: the pixels for which the reference's own fp32 arithmetic decides the
    arg-max.  Saturated softmaxes (two members certain of different classes: mean 0.5 / 0.5) make these common,
    so for such data this count -- not a fixed fraction of the frame -- is the cap on AWX_CNT_MARG_AMBIG."""
    import torch
    import torch.nn.functional as F
    m = sum(F.softmax(x.double(), dim=1) for x in members) / len(members)
    cand = m >= m.max(dim=1, keepdim=True).values * (1.0 - rel)
    lab = target.long().clamp(0, m.shape[1] - 1).unsqueeze(1)
    hit = cand.gather(1, lab).squeeze(1) & (cand.sum(dim=1) > 1) & (target != 255) & (target.long() < m.shape[1])
    return int(hit.sum())


def assert_ens_wrong_parity(bins, wrong_ref: int, n_valid: int, _lib, per_pixel: float = 2e-6, floor: int = 0,
                            cap: int = None) -> dict:
    """AWX_CNT_ENS_WRONG (= sum of the AUROC positives) == the oracle's count up to AWX_CNT_MARG_AMBIG."""
    mamb = bins.counter(_lib.CNT_MARG_AMBIG)
    if cap is None:
        cap = amb_bound(n_valid, per_pixel, floor)
    assert mamb <= cap, f"{mamb} mean-probability ties reported, more than {cap}"
    diff = abs(bins.counter(_lib.CNT_ENS_WRONG) - int(wrong_ref))
    assert diff <= mamb, f"ensemble-wrong count differs by {diff}, only {mamb} pixels reported ambiguous"
    return {"ens_wrong_mismatch": diff, "marg_ambiguous": mamb}


def cv2_blur_follows_the_restated_order() -> bool:
    """Does THIS machine's cv2.GaussianBlur (CV_32FC3) follow the operation order csrc/blur_strip.cuh restates --
    row filter fma(c, t0, fl((l + r) t1)) / left-to-right chain, column filter fl(c t0) then fma per tap pair?
    OpenCV picks its filter kernels by CPU dispatch (and fuses multiply-adds only in FMA builds), so the 0-LSB bar of
    the rain / snow images is asserted where a self-check on a random image says the order holds, the 1-LSB bar
    elsewhere.  fma is emulated in fp64 (the product of two fp32 values is exact there)."""
    import cv2
    rng = np.random.RandomState(0)
    h, w = 24, 48
    img = rng.rand(h, w, 3).astype(np.float32)

    def fma(a, b, c):
        return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)

    def mul(a, b):
        return (a * np.float32(b)).astype(np.float32)

    def pad(x, r, axis):
        n = x.shape[axis]
        lo = np.flip(np.take(x, range(1, r + 1), axis=axis), axis)
        hi = np.flip(np.take(x, range(n - r - 1, n - 1), axis=axis), axis)
        return np.concatenate([lo, x, hi], axis=axis)

    for k, sigma in ((3, 0.5), (7, 1.0)):
        r = k // 2
        taps = cv2.getGaussianKernel(k, sigma, cv2.CV_32F).ravel()
        t = [np.float32(taps[r + j]) for j in range(r + 1)]
        p = pad(img, r, 1)
        at = lambda q, off, axis, n: np.take(q, range(r + off, r + off + n), axis=axis)
        if r == 1:
            hres = fma(at(p, 0, 1, w), t[0], mul((at(p, -1, 1, w) + at(p, 1, 1, w)).astype(np.float32), t[1]))
        else:
            hres = mul(at(p, -r, 1, w), t[r])
            for j in range(-r + 1, r + 1):
                hres = fma(at(p, j, 1, w), t[abs(j)], hres)
        q = pad(hres, r, 0)
        v = mul(at(q, 0, 0, h), t[0])
        for j in range(1, r + 1):
            v = fma((at(q, -j, 0, h) + at(q, j, 0, h)).astype(np.float32), t[j], v)
        if not np.array_equal(v, cv2.GaussianBlur(img, (k, k), sigma)):
            return False
    return True
