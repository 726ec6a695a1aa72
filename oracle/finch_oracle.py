"""CPU oracle for the FINCH hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy / scipy / scikit-learn restatement of the reference algorithm in
/root/reference/clustering/finch.py (vendored from ssarfraz/FINCH-Clustering).  It is the
checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  Nothing under video_similarity_search_b200/ does.

Parity pin: the oracle is validated in the authoring container against the UNMODIFIED
reference module (oracle/reference_harness.py imports it by file path) - see
tests/golden/make_golden.py, which commits the reference's outputs as fixtures, and
tests/test_oracle_golden.py, which checks this restatement against them on every run.
For N > 70 000 the reference needs pyflann (absent; approximate Euclidean kd-trees, see
SURVEY.md D2), so above that size parity is pinned only against the reference driven by an
exact-NN stand-in: "parity unpinned against real FLANN".

Third-party arithmetic the reference leans on (not under /root/reference):
  scikit-learn (requirements.txt:5 pins 0.22.0; 1.9.0 here) - cosine_distances:
      normalise rows (norm via einsum, zero norms -> 1), X @ Y.T, 1 - S, clip to [0, 2]
  scipy.sparse.csgraph.connected_components(directed=True, connection='weak')
  numpy argmin / unique / cumsum
Each function below cites the reference lines it follows.
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components
from sklearn.metrics import pairwise_distances

FLANN_THRESHOLD = 70000  # finch.py:19


# --------------------------------------------------------------------------------------
# a2: first neighbours (finch.py:22-29)
# --------------------------------------------------------------------------------------
def _unit_rows(mat):
    """sklearn.preprocessing.normalize as cosine_similarity applies it: in the dtype of mat."""
    nrm = np.sqrt(np.einsum("ij,ij->i", mat, mat))
    nrm[nrm == 0.0] = 1.0
    return mat / nrm[:, None]


def first_neighbors_dense(mat):
    """finch.py:27-29 - full pairwise cosine distance, diagonal forced to 1e12, row argmin.

    Returns (nn int64 [n], dist [n,n] in the dtype of mat)."""
    dist = pairwise_distances(mat, mat, metric="cosine")
    np.fill_diagonal(dist, 1e12)
    return np.argmin(dist, axis=1), dist


def first_neighbors_blocked(mat, block=4096, rows=None):
    """Exact cosine first neighbour without the n x n matrix (the stand-in used where the
    reference would call pyflann, finch.py:30-38).  Same arithmetic as cosine_distances,
    evaluated one row block at a time.  `rows` restricts the queries (for sampled checks).

    Returns (nn int64, d1 distance to nn, gap = second-smallest minus smallest distance)."""
    unit = _unit_rows(np.ascontiguousarray(mat))
    n = unit.shape[0]
    rows = np.arange(n) if rows is None else np.asarray(rows)
    nn = np.empty(len(rows), dtype=np.int64)
    d1 = np.empty(len(rows), dtype=unit.dtype)
    gap = np.empty(len(rows), dtype=unit.dtype)
    for s in range(0, len(rows), block):
        r = rows[s:s + block]
        d = unit[r] @ unit.T
        d *= -1
        d += 1
        np.clip(d, 0, 2, out=d)
        d[np.arange(len(r)), r] = 1e12
        j = np.argmin(d, axis=1)
        best = d[np.arange(len(r)), j]
        d[np.arange(len(r)), j] = np.inf
        nn[s:s + block] = j
        d1[s:s + block] = best
        gap[s:s + block] = d.min(axis=1) - best
    return nn, d1, gap


# --------------------------------------------------------------------------------------
# a4: link graph + components (finch.py:40-47, 50-55)
# --------------------------------------------------------------------------------------
def link_graph(nn):
    """finch.py:41-46: A = (P + I)(P + I)^T with the diagonal cleared, P[i, nn[i]] = 1.
    Entries are 2 for mutual first neighbours, 1 for one-way links and for rows that share
    a first neighbour."""
    n = len(nn)
    p = sp.csr_matrix((np.ones(n, dtype=np.float32), (np.arange(n), nn)), shape=(n, n))
    p = p + sp.eye(n, dtype=np.float32, format="csr")
    a = (p @ p.T).tolil()
    a.setdiag(0)
    return a


def components(adj, dist=None, min_sim=None):
    """finch.py:50-55.  With min_sim the links whose weighted distance exceeds it are cut
    first.  scipy numbers components by their smallest member index."""
    if min_sim is not None:
        adj[np.where((dist * adj.toarray()) > min_sim)] = 0
    num, lab = connected_components(csgraph=adj, directed=True, connection="weak", return_labels=True)
    return lab, num


# --------------------------------------------------------------------------------------
# a5/a6: centroids and label composition (finch.py:58-82)
# --------------------------------------------------------------------------------------
def cluster_means(data, lab):
    """finch.py:58-71 (cool_mean): float64 per-cluster mean of the ORIGINAL rows via a
    running sum over label-sorted rows, differenced at the cluster boundaries."""
    _, cnt = np.unique(lab, return_counts=True)
    order = np.argsort(lab)
    acc = np.vstack((np.zeros((1, data.shape[1])), data[order, :]))
    np.cumsum(acc, axis=0, out=acc)
    hi = np.cumsum(cnt)
    lo = np.insert(hi, 0, 0)[:-1]
    return (acc[hi, :] - acc[lo, :]) / cnt[:, None]


def compose(prev, lab, data):
    """finch.py:74-82 (get_merge)."""
    if len(prev) != 0:
        _, inv = np.unique(prev, return_inverse=True)
        cur = lab[inv]
    else:
        cur = lab
    return cur, cluster_means(data, cur)


# --------------------------------------------------------------------------------------
# a1/a7: driver (finch.py:108-178)
# --------------------------------------------------------------------------------------
def _rank(mat, initial_rank, exact_above_threshold):
    """finch.py:22-47 (clust_rank).  Returns (adjacency, dist) where dist == [] whenever the
    reference would not have a dense matrix (given initial_rank, or n > FLANN_THRESHOLD)."""
    n = mat.shape[0]
    if initial_rank is not None:
        dist = []
        nn = np.asarray(initial_rank)
    elif n <= FLANN_THRESHOLD:
        nn, dist = first_neighbors_dense(mat)
    else:
        if not exact_above_threshold:
            raise MemoryError("You should use pyflann for inputs larger than %d samples." % FLANN_THRESHOLD)
        nn, _, _ = first_neighbors_blocked(mat)
        dist = []
    return link_graph(nn), dist, nn


def finch(data, initial_rank=None, req_clust=None, ensure_early_exit=True, verbose=False,
          exact_above_threshold=True, return_trace=False, nn0_override=None):
    """finch.py:108-178 with distance='cosine'.  Returns (c int [N,P], num_clust list, req_c)
    and, with return_trace, a dict of per-level first neighbours / min_sim for diagnostics.

    nn0_override (tests only): replace the level-0 first neighbours by the given ones while keeping
    everything else of the dense mode (distances, min_sim).  Used to show that a partition difference
    is explained entirely by rows whose float32 top-1/top-2 distances tie."""
    data = data.astype(np.float32)                                    # :131
    trace = {"nn": [], "min_sim": None}
    min_sim = None
    adj, dist, nn = _rank(data, initial_rank, exact_above_threshold)  # :134
    if nn0_override is not None:
        nn = np.asarray(nn0_override)
        adj = link_graph(nn)
    trace["nn"].append(np.asarray(nn))
    group, n0 = components(adj, [], None)                             # :136
    c, mat = compose([], group, data)                                 # :137
    if verbose:
        print("Partition 0: {} clusters".format(n0))
    if ensure_early_exit and len(dist) != 0:                          # :142-144
        min_sim = np.max(dist * adj.toarray())
    trace["min_sim"] = min_sim

    num_clust = [n0]
    c_ = c
    drop = 2
    k = 1
    while drop > 1:                                                   # :151
        adj, dist, nn = _rank(mat, None, exact_above_threshold)
        trace["nn"].append(np.asarray(nn))
        u, cur = components(adj, dist, min_sim)
        c_, mat = compose(c_, u, data)
        num_clust.append(cur)
        c = np.column_stack((c, c_))
        drop = num_clust[-2] - cur
        if cur == 1 or drop < 1:                                      # :160-163
            num_clust = num_clust[:-1]
            c = c[:, :-1]
            break
        if verbose:
            print("Partition {}: {} clusters".format(k, num_clust[k]))
        k += 1

    req_c = None
    if req_clust is not None:                                         # :169-176
        if req_clust not in num_clust:
            ind = [i for i, v in enumerate(num_clust) if v >= req_clust]
            req_c = refine_to(c[:, ind[-1]], data, req_clust)
        else:
            req_c = c[:, num_clust.index(req_clust)]
    if return_trace:
        return c, num_clust, req_c, trace
    return c, num_clust, req_c


# --------------------------------------------------------------------------------------
# a8: req_clust refinement (finch.py:85-105)
# --------------------------------------------------------------------------------------
def _closest_link_only(adj, dist):
    """finch.py:85-94 (update_adj): keep the two smallest-distance non-zero entries of the
    link graph - on a symmetric matrix that is one pair and its mirror."""
    r, cidx = adj.nonzero()
    order = np.argsort(dist[r, cidx])[:2]
    out = sp.lil_matrix(adj.get_shape())
    out[[r[order[0]], r[order[1]]], [cidx[order[0]], cidx[order[1]]]] = 1
    return out


def refine_to(lab, data, req_clust):
    """finch.py:97-105 (req_numclust): one closest-linked-pair merge per iteration."""
    todo = len(np.unique(lab)) - req_clust
    cur, mat = compose([], lab, data)
    for _ in range(todo):
        nn, dist = first_neighbors_dense(mat)
        adj = _closest_link_only(link_graph(nn), dist)
        u, _ = components(adj, [], None)
        cur, mat = compose(cur, u, data)
    return cur


# --------------------------------------------------------------------------------------
# a9: fit_cluster finch branch (cluster_masks.py:79-86)
# --------------------------------------------------------------------------------------
def fit_cluster_finch(embeddings, finch_partition=0):
    c, num_clust, _ = finch(np.asarray(embeddings))
    return c[:, finch_partition]
