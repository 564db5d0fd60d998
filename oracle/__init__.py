"""CPU oracle for the per-pixel robustness-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline -- never as a fallback for the CUDA path.

What it is: a functional restatement (NumPy / OpenCV / SciPy / torch-CPU /
scikit-learn -- the reference's own dependencies) of the arithmetic the
reference performs on this path, with the stochastic draws separated from the
per-pixel arithmetic so that both sides can be fed identical parameters.

Pinning: the reference ships no golden vectors for this path (SURVEY.md
section 8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
executed in the build container by ``tests/golden/make_golden.py`` and
committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every oracle function against them (bit-exact), and
``tests/test_oracle_live.py`` re-checks against the live reference whenever
``/root/reference`` is present.  Third-party arithmetic (cv2.line / cv2.circle
/ cv2.GaussianBlur / scipy gaussian_filter / torch softmax / sklearn
roc_auc_score) is *called*, not restated, exactly as the reference calls it;
the reference pins only version floors for those libraries, so parity at the
dependency level is "unpinned" and every parity report prints the versions.
"""

from . import weather, fusion, metrics, loss  # noqa: F401
