"""Oracle: ensemble logit fusion and the reverse-KL "JS" disagreement map.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``P/models/model.py:442-486`` (fusion) and ``:498-513`` (disagreement), where
``P/`` = /root/reference/src/adverse_weather_semantic_segmentation_robustness_benchmark/.
torch-CPU eager ops are used op-for-op (three separately rounded fp32 ops for
the weighted average, then a true division by the temperature).
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

STRATEGIES = ("weighted_average", "max_confidence", "mean")


def member_weights(raw_weights: torch.Tensor) -> torch.Tensor:
    """softmax over the two learnable scalars (model.py:444)."""
    return F.softmax(raw_weights, dim=0)


def fuse_logits(l1: torch.Tensor, l2: torch.Tensor, strategy: str,
                raw_weights: torch.Tensor | None = None,
                temperature: torch.Tensor | None = None) -> torch.Tensor:
    """Fused segmentation logits [B,C,H,W] (model.py:443-462).

    ``temperature`` None means ``temperature_scaling=False`` (no division).
    Any strategy string other than the first two takes the mean branch, as
    the reference's ``else`` does.
    """
    if strategy == "weighted_average":
        w = member_weights(raw_weights)
        out = w[0] * l1 + w[1] * l2
    elif strategy == "max_confidence":
        c1 = F.softmax(l1, dim=1).max(dim=1)[0]
        c2 = F.softmax(l2, dim=1).max(dim=1)[0]
        pick1 = (c1 > c2).float().unsqueeze(1)
        out = pick1 * l1 + (1 - pick1) * l2
    else:
        out = (l1 + l2) / 2
    if temperature is not None:
        out = out / temperature
    return out


def fuse_depth(d1: torch.Tensor, d2: torch.Tensor, strategy: str,
               raw_weights: torch.Tensor | None = None) -> torch.Tensor:
    """Depth fusion: weighted for weighted_average, else mean (model.py:471-478)."""
    if strategy == "weighted_average":
        w = member_weights(raw_weights)
        return w[0] * d1 + w[1] * d2
    return (d1 + d2) / 2


def reverse_kl_disagreement(l1: torch.Tensor, l2: torch.Tensor) -> torch.Tensor:
    """0.5*[KL(m||p)+KL(m||q)] as F.kl_div(log p, m) computes it (model.py:502-511,
    and identically evaluation/metrics.py:456-465).  Not a true JSD."""
    p = F.softmax(l1, dim=1)
    q = F.softmax(l2, dim=1)
    m = (p + q) / 2
    k1 = F.kl_div(p.log(), m, reduction="none").sum(dim=1)
    k2 = F.kl_div(q.log(), m, reduction="none").sum(dim=1)
    return (k1 + k2) / 2
