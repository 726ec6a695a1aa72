"""Generate tests/golden/*.npz by running the UNMODIFIED reference FINCH
(/root/reference/clustering/finch.py, imported by path) on seeded synthetic inputs.

Run in the authoring container only:   python tests/golden/make_golden.py
Inputs are not stored - they are re-created from (generator, n, d, k, seed) by
video_similarity_search_b200.synth; outputs (c, num_clust, req_c, per-level first neighbours,
min_sim) are.  The per-level first neighbours and min_sim are captured by wrapping the
reference module's own clust_rank / get_clust (no source edits).
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_harness as rh                     # noqa: E402
from video_similarity_search_b200 import synth                 # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> dict(gen=..., n, d, k, seed, and FINCH kwargs / harness knobs)
CASES = {
    "gmm_1200x64":        dict(gen="gmm", n=1200, d=64, k=12, seed=3),
    "iid_600x32":         dict(gen="iid", n=600, d=32, seed=5),
    "gmm_3000x128":       dict(gen="gmm", n=3000, d=128, k=30, seed=7),
    "gmm_3000x128_rank":  dict(gen="gmm", n=3000, d=128, k=30, seed=7, use_initial_rank=True),
    "gmm_3000x128_req":   dict(gen="gmm", n=3000, d=128, k=30, seed=7, req_clust=20),
    "gmm_3000x128_noexit": dict(gen="gmm", n=3000, d=128, k=30, seed=7, ensure_early_exit=False),
    "gmm_2500x96_flann":  dict(gen="gmm", n=2500, d=96, k=25, seed=11, flann_threshold=1000),
    "gmm_777x200_odd":    dict(gen="gmm", n=777, d=200, k=9, seed=13),
    "c1_9537x512":        dict(gen="gmm", n=9537, d=512, k=101, seed=0),
    # degenerate shapes: the smallest inputs, and all-zero rows (unit row = 0, every distance exactly 1.0: the
    # first neighbour is the lowest other index, np.argmin's tie rule with no rounding involved)
    "tiny_2x8":           dict(gen="iid", n=2, d=8, seed=21),
    "tiny_3x8":           dict(gen="iid", n=3, d=8, seed=22),
    "tiny_5x16":          dict(gen="iid", n=5, d=16, seed=23),
    "zeros_500x32":       dict(gen="gmm_zero_rows", n=500, d=32, k=6, seed=24, zero_rows=[0, 7, 8, 250, 499]),
}


def make_input(case):
    if case["gen"] == "gmm":
        return synth.gaussian_mixture(case["n"], case["d"], case["k"], case["seed"])
    if case["gen"] == "gmm_zero_rows":
        x = synth.gaussian_mixture(case["n"], case["d"], case["k"], case["seed"])
        x[case["zero_rows"]] = 0
        return x
    return synth.iid_normal(case["n"], case["d"], case["seed"])


def run_reference(case):
    x = make_input(case)
    thr = case.get("flann_threshold")
    mod = rh.load_reference_finch(with_exact_flann=thr is not None)
    if thr is not None:
        mod.FLANN_THRESHOLD = thr            # module constant, finch.py:19
    nn_levels, min_sims = [], []
    orig_rank, orig_clust = mod.clust_rank, mod.get_clust

    def rank_spy(mat, initial_rank=None, distance="cosine"):
        adj, dist = orig_rank(mat, initial_rank, distance)
        if initial_rank is not None:
            nn_levels.append(np.asarray(initial_rank, dtype=np.int64))
        elif len(dist) != 0:
            nn_levels.append(np.argmin(dist, axis=1).astype(np.int64))
        else:
            nn_levels.append(np.zeros(0, dtype=np.int64))    # stand-in level: not observable
        return adj, dist

    def clust_spy(a, orig_dist, min_sim=None):
        min_sims.append(np.nan if min_sim is None else float(min_sim))
        return orig_clust(a, orig_dist, min_sim)

    mod.clust_rank, mod.get_clust = rank_spy, clust_spy
    kwargs = dict(distance="cosine", verbose=False,
                  ensure_early_exit=case.get("ensure_early_exit", True),
                  req_clust=case.get("req_clust"))
    if case.get("use_initial_rank"):
        plain = rh.load_reference_finch()
        _, dist = plain.clust_rank(x.astype(np.float32), None, "cosine")
        kwargs["initial_rank"] = np.argmin(dist, axis=1)
    with contextlib.redirect_stdout(io.StringIO()):
        c, num_clust, req_c = mod.FINCH(x, **kwargs)
    n_main = len(num_clust) + 1 if len(nn_levels) > len(num_clust) else len(nn_levels)
    out = dict(c=c.astype(np.int32), num_clust=np.asarray(num_clust, dtype=np.int64),
               req_c=np.zeros(0, np.int32) if req_c is None else np.asarray(req_c, np.int32),
               min_sim=np.float64(min_sims[1]) if len(min_sims) > 1 else np.float64(np.nan),
               n_levels_run=np.int64(n_main))
    for lvl, nn in enumerate(nn_levels[:n_main]):
        out["nn_level%d" % lvl] = nn
    if "initial_rank" in kwargs:
        out["initial_rank"] = np.asarray(kwargs["initial_rank"], dtype=np.int64)
    return out


def main():
    for name, case in CASES.items():
        out = run_reference(case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, out["num_clust"].tolist(), "min_sim", float(out["min_sim"]),
              "levels", int(out["n_levels_run"]))


if __name__ == "__main__":
    main()
