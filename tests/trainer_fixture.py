"""Helpers shared by the CPU and GPU tests of the trainer twin: the validation loader stored in
tests/golden/trainer.npz (made by tests/golden/make_golden.py::trainer_cases) and a model replaying its outputs."""

import torch


def trainer_fixture_batches(g, tag):
    """The validation loader stored in tests/golden/trainer.npz and a model replaying its stored outputs."""
    nb = int(g[f"{tag}_args"][1])
    batches = []
    for i in range(nb):
        b = {}
        for k in ("image", "label", "_seg", "depth", "_depth"):
            if f"{tag}_b{i}_{k}" in g.files:
                b[k] = torch.from_numpy(g[f"{tag}_b{i}_{k}"])
        b["weather_condition"] = [str(x) for x in g[f"{tag}_b{i}_weather_condition"]]
        batches.append(b)
    return batches


class ReplayModel(torch.nn.Module):
    def __init__(self, batches, device=None):
        super().__init__()
        self.batches, self.i, self.device = batches, 0, device

    def forward(self, x):
        b = self.batches[self.i]
        self.i += 1
        out = {"segmentation": b["_seg"]}
        if "_depth" in b:
            out["depth"] = b["_depth"]
        return {k: v.to(self.device) for k, v in out.items()} if self.device is not None else out


def evaluate_fixture_batches(g, tag):
    """The test loader stored in tests/golden/evaluate.npz (make_golden.py::evaluate_cases): labels, weather
    conditions and both members' logits per batch; the images are placeholders (the models replay stored logits)."""
    nb, bsz = int(g[f"{tag}_args"][1]), int(g[f"{tag}_args"][2])
    batches = []
    for i in range(nb):
        b = {"image": torch.zeros(bsz, 3, 4, 4)}
        for k in ("label", "_la", "_lb"):
            b[k] = torch.from_numpy(g[f"{tag}_b{i}_{k}"])
        b["weather_condition"] = [str(x) for x in g[f"{tag}_b{i}_weather_condition"]]
        batches.append(b)
    return batches


class ReplayMember(torch.nn.Module):
    """One ensemble member replaying the stored logits of successive batches."""

    def __init__(self, batches, key, device=None):
        super().__init__()
        self.batches, self.key, self.i, self.device = batches, key, 0, device

    def forward(self, x):
        out = self.batches[self.i][self.key]
        self.i += 1
        return {"segmentation": out.to(self.device) if self.device is not None else out}
