"""The built library is sm_100a machine code with the instruction classes the design rests on (no GPU needed:
cuobjdump disassembles the in-tree libawx.so).  A regression guard for "the fast path silently fell back to generic
code": bulk async copies + mbarriers in the score / loss kernels, packed fp32 pairs in the score and strip kernels,
cp.async staging in the strip kernels, MUFU.EX2 and fp64 in the screened fog kernel."""

import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adverse_weather_semantic_segmentation_robustness_benchmark_b200", "libawx.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs cuobjdump and a built libawx.so")


@pytest.fixture(scope="module")
def kernels():
    """{demangled-ish kernel name: Counter of opcodes}"""
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", txt))
    out, name = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            out[name][m.group(1)] += 1
    return archs, out


def _matching(kernels, needle):
    found = {k: v for k, v in kernels.items() if needle in k}
    assert found, f"no kernel named *{needle}* in libawx.so"
    return found


def test_only_sm_100a(kernels):
    archs, _ = kernels
    assert archs == {"sm_100a"}, archs


def test_score_kernels_use_bulk_copies_mbarriers_and_packed_math(kernels):
    _, k = kernels
    for name, ops in _matching(k, "score_v2_kernel").items():
        assert ops["UBLKCP"] > 0 and ops["SYNCS"] > 0, f"{name}: no bulk async copy / mbarrier"
        assert ops["FFMA2"] > 0 and ops["MUFU"] > 0, f"{name}: no packed fp32 / MUFU"


def test_loss_ring_kernel_uses_bulk_copies(kernels):
    _, k = kernels
    for name, ops in _matching(k, "fogloss_ring_kernel").items():
        assert ops["UBLKCP"] > 0 and ops["SYNCS"] > 0, name


def test_strip_kernels_are_packed_and_staged_with_cp_async(kernels):
    _, k = kernels
    found = _matching(k, "blur_strip_kernel")
    assert len(found) == 4          # {3, 7 taps} x {rain, snow}
    for name, ops in found.items():
        assert ops["LDGSTS"] >= 5, f"{name}: row bytes are not staged with cp.async"
        assert ops["FFMA2"] > 50 and ops["FADD2"] > 20 and ops["FMUL2"] > 20, f"{name}: filters are not packed"
        assert ops["F2IP"] > 0 and ops["BAR"] >= 1, name


def test_fog_and_night_kernels(kernels):
    _, k = kernels
    (name, ops), = _matching(k, "fog_kernel").items()
    assert ops["MUFU"] >= 4 and ops["FFMA2"] > 0 and ops["DFMA"] > 0, f"{name}: screen (MUFU, packed) + exact path (fp64)"
    assert ops["BAR"] == 0 and ops["STS"] == 0, f"{name}: must not stage through shared memory"
    (name, ops), = _matching(k, "night_kernel").items()
    assert ops["DADD"] > 0 and ops["BAR"] == 0 and ops["STS"] == 0, name
