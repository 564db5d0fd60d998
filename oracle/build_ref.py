"""Recipe for ``oracle/_ref/``: the REFERENCE's own hot-path modules, byte-compiled where they lie.

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python, so its "build" is a
byte-compilation: the three files of SURVEY.md section 8a

    src/<pkg>/data/preprocessing.py      WeatherDegradationTransforms            (8a rows 1-7)
    src/<pkg>/models/model.py            EnsembleModel, FogDensityAwareLoss      (8a rows 8-10)
    src/<pkg>/evaluation/metrics.py      IoU / ECE / disagreement / Robustness   (8a rows 11-14)

are compiled from ``/root/reference`` straight into ``oracle/_ref/*.refbin`` (marshalled code objects, the .pyc format) -- build products only: no reference
source is copied into the repo, ``oracle/_ref/`` is git-ignored, and (like the repo's own ``libawx.so``) it travels to
the GPU box with the snapshot, where ``/root/reference`` does not exist.  ``oracle/reference.py`` loads them.

    python -m oracle.build_ref            (also run by __graft_entry__.build() when the reference is present)
"""

from __future__ import annotations

import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("AWX_REFERENCE_ROOT", "/root/reference")
PKG = "adverse_weather_semantic_segmentation_robustness_benchmark"
EXT = ".refbin"   # .pyc content; the neutral extension keeps file-sync tools from dropping it as a cache file
FILES = {"preprocessing": "data/preprocessing.py", "model": "models/model.py", "metrics": "evaluation/metrics.py"}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", PKG))


def build() -> str:
    """Byte-compile the three modules into oracle/_ref/; returns the output directory."""
    if not available():
        raise RuntimeError(f"{REF_ROOT} is not present: oracle/_ref can only be built in the build container")
    os.makedirs(OUT, exist_ok=True)
    meta = {"python": sys.version.split()[0], "magic": __import__("importlib.util").util.MAGIC_NUMBER.hex(), "files": {}}
    for name, rel in FILES.items():
        src = os.path.join(REF_ROOT, "src", PKG, rel)
        py_compile.compile(src, cfile=os.path.join(OUT, name + EXT), dfile=f"<reference>/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        meta["files"][name] = rel
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    return OUT


if __name__ == "__main__":
    print(build())
