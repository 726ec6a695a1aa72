"""FINCH first-neighbour clustering on B200 - drop-in for /root/reference/clustering/finch.py.

Same call signature and return values as the reference's FINCH (finch.py:108-178):

    c, num_clust, req_c = FINCH(data, initial_rank=None, req_clust=None, distance='cosine',
                                ensure_early_exit=True, verbose=True)

What runs where
  host (this file): the level loop and its exit rules, exactly as finch.py:134-176 has them.
  device (libslic_b200.so through backend.CudaBackend):
    a2  first neighbours   slic_normalize_rows + slic_nn_top1 (tcgen05 f16 screen, exact re-rank in
                           float32 at level 0 / float64 at levels >= 1) or slic_nn_exact_top1 (small n)
    a4  components         slic_finch_components (union-find on nn[], optional min_sim cut)
    a5/a6 merge + means    slic_compose_labels, slic_segmented_mean (float64)
    a7  min_sim            slic_finch_min_sim
    a8  req_clust          slic_finch_closest_link + the calls above
There is no CPU path: without a CUDA device the backend constructor raises.

Differences from the reference, all deliberate:
  * N > 70 000 does not need pyflann: the exact cosine first neighbour is computed at every size.
    The reference's control flow for that case is kept - no dense distances => no min_sim filter
    (finch.py:30-38,142-144).
  * only distance='cosine' (the one value any caller passes) is implemented.
"""
import numpy as np
import torch

from .. import _lib
from .. import backend as _backend

FLANN_THRESHOLD = 70000  # finch.py:19 - above it the reference has no dense distance matrix


class _Level:
    """Result of the clust_rank step (finch.py:22-47) for one level, kept on the device."""

    def __init__(self, nn, dist, unit, dense):
        self.nn, self.dist, self.unit, self.dense = nn, dist, unit, dense


def _rank(be, mat, initial_rank, first_neighbors=None):
    """finch.py:22-38.  `dense` mirrors `len(orig_dist) != 0` in the reference."""
    n = mat.shape[0]
    if initial_rank is not None:
        nn = be.to_device(np.asarray(initial_rank).astype(np.int32, copy=False), torch.int32)
        if nn.shape[0] != n:
            raise ValueError("initial_rank must have one entry per row of data")
        return _Level(nn, None, None, False)
    if n == 1:
        return _Level(torch.zeros(1, dtype=torch.int32, device=mat.device), torch.zeros(1, dtype=mat.dtype,
                      device=mat.device), mat, n <= FLANN_THRESHOLD)
    if first_neighbors is not None:
        nn, dist, unit = first_neighbors(mat)
    else:
        nn, dist, unit = be.first_neighbors(mat)
    return _Level(nn, dist, unit, n <= FLANN_THRESHOLD)


def _clust(be, lvl, min_sim):
    """finch.py:50-55."""
    if min_sim is not None and lvl.dense:
        return be.components(lvl.nn, min_sim=float(min_sim), unit=lvl.unit, dist=lvl.dist)
    return be.components(lvl.nn)


class _Sums:
    """float64 row sums and row counts of the current partition's clusters (device)."""

    def __init__(self, sums, counts):
        self.sums, self.counts = sums, counts


def _merge(be, prev, u, data, num_clust, state=None):
    """finch.py:74-82: compose the labels and return the float64 centroids of the ORIGINAL rows.
    The reference re-reads all N rows at every level (cool_mean, :58-71).  The clusters of a level are unions
    of the previous level's clusters, so here only level 0 reads the rows; later levels add the previous level's
    per-cluster float64 sums (slic_merge_cluster_sums) - the same means to float64 rounding.
    Returns (labels [N], means [C, D], _Sums)."""
    cur = be.compose_labels(prev, u)
    if state is None:
        sums, counts, means = be.cluster_sums(data, cur, num_clust)
    else:
        sums, counts, means = be.merge_cluster_sums(state.sums, state.counts, u, num_clust)
    return cur, means, _Sums(sums, counts)


def _req_numclust(be, labels, n_labels, data, req_clust):
    """finch.py:97-105: merge the closest linked pair until req_clust clusters remain."""
    cur, mat, state = _merge(be, None, labels, data, n_labels)
    for _ in range(n_labels - req_clust):
        n = mat.shape[0]
        lvl = _rank(be, mat, None)
        i, j = be.closest_link(lvl.nn, lvl.unit, lvl.dist)          # update_adj, finch.py:85-94
        link = torch.arange(n, dtype=torch.int32, device=mat.device)
        link[j] = i                                                 # the single surviving link
        u, cnt = be.components(link)
        cur, mat, state = _merge(be, cur, u, data, cnt, state)
    return cur


def _finch_loop(be, data, initial_rank, ensure_early_exit, verbose, first_neighbors):
    """The level loop of finch.py:134-167 driven from Python, one backend call per step.  Used with stand-in
    backends (tests) and as the continuation when the native driver's label buffer is too small.
    -> (columns list of device [N] labels, num_clust list)."""
    min_sim = None
    lvl = _rank(be, data, initial_rank, first_neighbors)            # finch.py:134
    group, n0 = _clust(be, lvl, None)                               # finch.py:136
    c_, mat, state = _merge(be, None, group, data, n0)              # finch.py:137
    if verbose:
        print('Partition 0: {} clusters'.format(n0))
    if ensure_early_exit and lvl.dense and lvl.dist is not None and data.shape[0] > 1:
        min_sim = be.min_sim(lvl.nn, lvl.unit, lvl.dist)            # finch.py:142-144

    columns = [c_]
    num_clust = [n0]
    exit_clust = 2
    k = 1
    while exit_clust > 1:                                           # finch.py:151
        lvl = _rank(be, mat, None)
        u, cur = _clust(be, lvl, min_sim)
        c_, mat, state = _merge(be, c_, u, data, cur, state)
        num_clust.append(cur)
        columns.append(c_)
        exit_clust = num_clust[-2] - cur
        if cur == 1 or exit_clust < 1:                              # finch.py:160-163
            num_clust = num_clust[:-1]
            columns = columns[:-1]
            break
        if verbose:
            print('Partition {}: {} clusters'.format(k, num_clust[k]))
        k += 1
    return columns, num_clust


def _is_host_matrix(data):
    return isinstance(data, np.ndarray) or (isinstance(data, torch.Tensor) and not data.is_cuda)


def _finch_native(be, data, initial_rank, ensure_early_exit, first_neighbors):
    """The hierarchy through the native driver (csrc/finch_driver.cu): ONE C-ABI call.
    -> (c numpy [N, P], num_clust, data on the device or None)."""
    _lib.call("slic_set_flann_threshold", int(FLANN_THRESHOLD))    # finch.py:19 is a module constant; keep it one here
    if _is_host_matrix(data) and first_neighbors is None:
        # the reference-facing call: host matrix in, host labels out; the upload is pipelined behind the level-0
        # screen inside slic_finch_host
        x = data.numpy() if isinstance(data, torch.Tensor) else data
        x = np.ascontiguousarray(x, dtype=np.float32)               # finch.py:131
        if x.ndim != 2 or x.shape[0] < 1:
            raise ValueError("data must be a non-empty [N, D] matrix")
        c, num_clust, _ = be.finch_host(x, initial_rank, ensure_early_exit)
        return c, num_clust, None
    dev = be.to_device(data, torch.float32)                         # finch.py:131
    if dev.dim() != 2 or dev.shape[0] < 1:
        raise ValueError("data must be a non-empty [N, D] matrix")
    n = dev.shape[0]
    nn0 = dist0 = unit0 = None
    dense0 = False
    if initial_rank is not None:
        nn0 = be.to_device(np.asarray(initial_rank).astype(np.int32, copy=False), torch.int32)
        if nn0.shape[0] != n:
            raise ValueError("initial_rank must have one entry per row of data")
    elif first_neighbors is not None and n > 1:
        comm = first_neighbors.native_comm(dev) if hasattr(first_neighbors, "native_comm") else None
        if comm is not None:
            # multi-GPU, one process per GPU: normalise + shared level-0 search + the rest of the hierarchy behind ONE
            # call on every rank (csrc/comm.cu), no host synchronisation between the stages
            c, num_clust, _ = be.finch_native_comm(comm, dev, ensure_early_exit, host_labels=True)
            return c, num_clust, dev
        nn0, dist0, unit0 = first_neighbors(dev)
        dense0 = n <= FLANN_THRESHOLD
    # (host_labels: the device writes the label matrix straight into page-locked host memory - FINCH returns numpy)
    c, num_clust, _ = be.finch_native(dev, nn0, dist0, unit0, dense0, ensure_early_exit, host_labels=True)
    return c, num_clust, dev


def FINCH(data, initial_rank=None, req_clust=None, distance='cosine', ensure_early_exit=True, verbose=True,
          backend=None, first_neighbors=None):
    """FINCH clustering (reference: clustering/finch.py:108-178).

    :param data: [N, D] features in rows - numpy array or torch tensor (CPU or CUDA), any float dtype;
                 cast to float32 as the reference does (finch.py:131).
    :param initial_rank: optional [N] first-neighbour indices (skips the level-0 search and, as in the
                 reference, disables the min_sim filter).
    :param req_clust: optional exact number of clusters to refine to.
    :param distance: only 'cosine'.
    :param ensure_early_exit: apply the min_sim purity filter when level 0 had dense distances.
    :param verbose: print 'Partition k: n clusters' lines like the reference.
    :param backend: device backend (default: the process-wide CudaBackend).
    :param first_neighbors: optional callable mat -> (nn, dist, unit) overriding the level-0 search
                 (the multi-GPU row-sharded search in sharded.py plugs in here).
    :return: c int32 [N, P] (numpy), num_clust list[int], req_c int32 [N] or None.
    """
    if distance != 'cosine':
        raise NotImplementedError("only distance='cosine' is implemented (the value every reference caller passes)")
    be = backend if backend is not None else _backend.default_backend()
    if isinstance(data, torch.Tensor):
        data = data.detach()

    c = None
    dev = None
    if hasattr(be, "finch_native"):
        try:
            c, num_clust, dev = _finch_native(be, data, initial_rank, ensure_early_exit, first_neighbors)
            if verbose:
                for k, v in enumerate(num_clust):
                    print('Partition {}: {} clusters'.format(k, v))
        except _lib.SlicError as e:
            if "status -5" not in str(e):   # SLIC_ERR_OVERFLOW: more levels than the driver's label buffer holds
                raise
    if c is None:
        dev = be.to_device(data, torch.float32)                     # finch.py:131
        if dev.dim() != 2 or dev.shape[0] < 1:
            raise ValueError("data must be a non-empty [N, D] matrix")
        columns, num_clust = _finch_loop(be, dev, initial_rank, ensure_early_exit, verbose, first_neighbors)
        c = be.to_host(torch.stack(columns, dim=1))

    req_c = None
    if req_clust is not None:                                       # finch.py:169-176
        if req_clust not in num_clust:
            if dev is None:
                dev = be.to_device(data, torch.float32)
            ind = [i for i, v in enumerate(num_clust) if v >= req_clust]
            col = be.to_device(np.ascontiguousarray(c[:, ind[-1]]), torch.int32)
            req_c = be.to_host(_req_numclust(be, col, num_clust[ind[-1]], dev, req_clust))
        else:
            req_c = np.ascontiguousarray(c[:, num_clust.index(req_clust)])
    return c, num_clust, req_c
