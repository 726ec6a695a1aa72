"""Drive the UNMODIFIED reference FINCH from /root/reference (authoring container only).

TEST INFRASTRUCTURE ONLY, and never used at run time on the GPU box (/root/reference does not
exist there): tests/golden/make_golden.py calls this to produce the committed fixtures and
tests/test_oracle_golden.py re-checks the oracle against the live reference when it is present.

For n > 70 000 the reference calls pyflann (finch.py:30-38), which is not installed; an exact
cosine 2-NN stand-in module is injected instead - results obtained
through it are labelled "reference + exact-NN stand-in".
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("SLIC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "clustering", "finch.py"))


class _ExactFLANN:
    """pyflann.FLANN look-alike: nn(pts, qpts, num_neighbors=2, ...) -> exact cosine 2-NN."""

    def nn(self, pts, qpts, num_neighbors=2, **kwargs):
        from oracle.finch_oracle import _unit_rows
        unit = _unit_rows(np.ascontiguousarray(pts))
        n = unit.shape[0]
        res = np.empty((n, 2), dtype=np.int64)
        dst = np.empty((n, 2), dtype=unit.dtype)
        for s in range(0, n, 4096):
            e = min(n, s + 4096)
            d = unit[s:e] @ unit.T
            d *= -1
            d += 1
            np.clip(d, 0, 2, out=d)
            r = np.arange(e - s)
            d[r, np.arange(s, e)] = -1.0          # the query point itself comes back first
            j0 = np.argmin(d, axis=1)
            d[r, j0] = np.inf
            j1 = np.argmin(d, axis=1)
            res[s:e, 0], res[s:e, 1] = j0, j1
            dst[s:e, 0], dst[s:e, 1] = 0.0, d[r, j1]
        return res, dst


def load_reference_finch(with_exact_flann=False):
    """Import /root/reference/clustering/finch.py by path, optionally with the stand-in."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    saved = sys.modules.get("pyflann")
    if with_exact_flann:
        fake = types.ModuleType("pyflann")
        fake.FLANN = _ExactFLANN
        fake.__all__ = ["FLANN"]
        sys.modules["pyflann"] = fake
    try:
        spec = importlib.util.spec_from_file_location(
            "_reference_finch_%d" % int(with_exact_flann),
            os.path.join(REFERENCE_ROOT, "clustering", "finch.py"))
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    finally:
        if with_exact_flann:
            if saved is None:
                sys.modules.pop("pyflann", None)
            else:
                sys.modules["pyflann"] = saved
    return mod
