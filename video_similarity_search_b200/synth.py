"""Seeded synthetic embedding generators shared by the oracle, the tests and bench.py.

The shapes are the BASELINE.json configs: a Gaussian mixture with
unit-variance noise around K standard-normal centres, cast to float32 and NOT L2-normalised
(the reference clusters un-normalised projection-head outputs, /root/reference
models/resnet.py:294-299).  Only numpy is used so the same bytes come out on every box.
"""
import numpy as np

# name -> (N, D, K, seed)
CONFIGS = {
    "C1": (9537, 512, 101, 0),        # UCF101 split-1 train size, ResNet-18 3D width
    "C3": (240000, 512, 400, 0),      # Kinetics-400 train size
    "C5": (1000000, 1024, 1000, 0),   # scale sweep, S3D width
}


def gaussian_mixture(n, d, k, seed, dtype=np.float32, return_labels=False, chunk=65536):
    """centres ~ N(0,I) [k,d]; lab ~ U{0..k-1}; X = centres[lab] + N(0,I).

    Generated in row chunks so the 1M x 1024 case never holds a float64 [N,D] temporary;
    the stream of random numbers (and therefore the bytes) is identical to the one-shot form.
    """
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((k, d))
    lab = rng.integers(0, k, n)
    out = np.empty((n, d), dtype=dtype)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        out[s:e] = (centres[lab[s:e]] + rng.standard_normal((e - s, d))).astype(dtype)
    if return_labels:
        return out, lab.astype(np.int64), centres
    return out


def config(name, dtype=np.float32, return_labels=False):
    n, d, k, seed = CONFIGS[name]
    return gaussian_mixture(n, d, k, seed, dtype=dtype, return_labels=return_labels)


def retrieval_queries(q, centres, seed, dtype=np.float32):
    """C2 queries: drawn from the SAME centres as the database, different seed."""
    rng = np.random.default_rng(seed)
    k, d = centres.shape
    lab = rng.integers(0, k, q)
    x = (centres[lab] + rng.standard_normal((q, d))).astype(dtype)
    return x, lab.astype(np.int64)


def c2_retrieval(dtype=np.float32):
    """(train [9537,512], train_labels, test [3783,512], test_labels) - iic_retrieve_clips UCF101 shape."""
    n, d, k, seed = CONFIGS["C1"]
    train, ytr, centres = gaussian_mixture(n, d, k, seed, dtype=np.float32, return_labels=True)
    test, yte = retrieval_queries(3783, centres, 1, dtype=np.float32)
    return train.astype(dtype), ytr, test.astype(dtype), yte


def iid_normal(n, d, seed, dtype=np.float32):
    """No-structure stress input (small top-1/top-2 gaps)."""
    return np.random.default_rng(seed).standard_normal((n, d)).astype(dtype)
