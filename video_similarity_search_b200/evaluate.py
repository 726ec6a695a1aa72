"""Retrieval helpers on B200 - drop-in for the distance / top-k functions of
/root/reference/evaluate.py (get_distance_matrix :208-223, get_closest_data_mat :226-231,
get_closest_data :234-238, get_topk_acc :287-307).

The reference materialises the dense [Q, N] matrix with sklearn and then partitions it.  Here
get_distance_matrix returns a lazy CosineDistances / dense-on-demand object: the unit rows stay on
the GPU and get_closest_data_mat / get_topk_acc run the fused top-k (no [Q, N] matrix in HBM beyond
one row block).  Indexing or np.asarray() on the object materialises it (callers such as
evaluate.py:384 index rows for plotting), so existing code keeps working.
"""
import numpy as np
import torch

from . import backend as _backend


def _as_float_array(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def _pair_dtype(x, y):
    """sklearn's _return_float_dtype: float32 only if both inputs are float32, else float64."""
    if x.dtype == np.float32 and (y is None or y.dtype == np.float32):
        return np.float32
    return np.float64


class DistanceMatrix:
    """Lazy [Q, N] distance matrix.  Behaves like the ndarray the reference returns (shape, dtype,
    indexing, np.asarray) but keeps the embeddings on the device until a dense view is needed."""

    def __init__(self, be, q, x, metric, same):
        self._be, self._q, self._x, self.metric, self.same = be, q, x, metric, same
        self._dense = None
        self._unit_q = self._unit_x = self._bf_q = self._bf_x = None
        self.shape = (q.shape[0], x.shape[0])
        self.dtype = np.dtype(np.float32 if x.dtype == torch.float32 else np.float64)

    def _units(self):
        if self._unit_x is None:
            screen = self._x.shape[0] >= _backend.SCREEN_MIN_ROWS
            self._unit_x, self._bf_x = self._be.normalize_rows(self._x, want_f16=screen)
            if self.same:
                self._unit_q, self._bf_q = self._unit_x, self._bf_x
            else:
                self._unit_q, self._bf_q = self._be.normalize_rows(self._q, want_f16=screen)
        return self._unit_q, self._unit_x

    def device_matrix(self):
        """Dense matrix on the device, diagonal = +inf when built from one set (evaluate.py:221-222)."""
        if self.metric == 'cosine':
            uq, ux = self._units()
            m = self._be.distance_matrix(uq, ux, 'cosine', same=self.same)
        else:
            m = self._be.distance_matrix(self._q, self._x, 'euclidean', same=self.same)
        if self.same:
            m.fill_diagonal_(float('inf'))
        return m

    def numpy(self):
        if self._dense is None:
            self._dense = self._be.to_host(self.device_matrix())
        return self._dense

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, item):
        return self.numpy()[item]

    def __len__(self):
        return self.shape[0]

    def topk(self, k):
        """(idx int32 [Q,k], dist [Q,k]) on the device, ascending distance, ties -> lowest column."""
        if self.metric == 'cosine':
            uq, ux = self._units()
            return self._be.topk_cosine(uq, ux, k, self_offset=0 if self.same else -1, q_f16=self._bf_q,
                                        x_f16=self._bf_x)
        return self._be.rows_topk(self.device_matrix(), k)


def get_distance_matrix(x_embeddings, y_embeddings=None, dist_metric='cosine', backend=None, lazy=True):
    """evaluate.py:208-223.  Returns a DistanceMatrix (see module docstring): every caller in the reference
    (evaluate.py:367-384, validation.py:81-82, :130-131) hands the result to get_topk_acc / get_closest_data(_mat) or
    reads .shape / indexes rows, all of which the lazy object serves.  lazy=False returns the materialised
    np.ndarray itself (what the reference returns), for code that checks isinstance(..., np.ndarray)."""
    assert(dist_metric in ['cosine', 'euclidean'])                  # evaluate.py:211
    be = backend or _backend.default_backend()
    x = _as_float_array(x_embeddings)
    y = None if y_embeddings is None else _as_float_array(y_embeddings)
    dt = _pair_dtype(x, y)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    xd = be.to_device(x.astype(dt, copy=False), tdt)
    yd = xd if y is None else be.to_device(y.astype(dt, copy=False), tdt)
    dm = DistanceMatrix(be, xd, yd, dist_metric, same=y is None)
    return dm if lazy else dm.numpy()


def _topk_from_any(distance_matrix, top_k, backend=None):
    if isinstance(distance_matrix, DistanceMatrix):
        return distance_matrix.topk(top_k)
    be = backend or _backend.default_backend()
    m = _as_float_array(distance_matrix)
    if m.dtype not in (np.float32, np.float64):
        m = m.astype(np.float64)
    md = be.to_device(np.ascontiguousarray(m), torch.float32 if m.dtype == np.float32 else torch.float64)
    return be.rows_topk(md, top_k)


def get_closest_data_mat(distance_matrix, top_k, backend=None):
    """evaluate.py:226-231: indices of the top_k smallest distances of every row, ascending.
    Returns int64 [Q, top_k] (numpy) like np.argpartition / take_along_axis do."""
    idx, _ = _topk_from_any(distance_matrix, top_k, backend)
    return idx.cpu().numpy().astype(np.int64)


def get_closest_data(distance_matrix, exemplar_idx, top_k, backend=None):
    """evaluate.py:234-238: the same for one row."""
    if isinstance(distance_matrix, DistanceMatrix):
        row = distance_matrix.numpy()[exemplar_idx][None, :]
        backend = backend or distance_matrix._be
    else:
        row = _as_float_array(distance_matrix)[exemplar_idx][None, :]
    return get_closest_data_mat(np.ascontiguousarray(row), top_k, backend)[0]


def _int_labels(x_labels, y_labels):
    xa, ya = np.asarray(x_labels), np.asarray(y_labels)
    if xa.dtype.kind in 'iu' and ya.dtype.kind in 'iu':
        return xa.astype(np.int64), ya.astype(np.int64)
    uniq, inv = np.unique(np.concatenate([xa.ravel(), ya.ravel()]), return_inverse=True)
    return inv[:xa.size].astype(np.int64), inv[xa.size:].astype(np.int64)


def get_topk_acc(distance_matrix, x_labels, y_labels=None, top_ks=[1, 5, 10, 20], backend=None):
    """evaluate.py:287-307: hit@k for every k in top_ks (query label among the labels of its k nearest)."""
    be = backend or (distance_matrix._be if isinstance(distance_matrix, DistanceMatrix) else _backend.default_backend())
    top_k = top_ks[-1]
    idx, _ = _topk_from_any(distance_matrix, top_k, be)
    if y_labels is None:
        y_labels = x_labels
    xl, yl = _int_labels(x_labels, y_labels)
    ks = [int(k) for k in top_ks]
    if sorted(ks) != ks:
        raise ValueError("top_ks must be ascending (the reference takes top_ks[-1] as the largest)")
    hits = be.hit_at_k(idx, be.to_device(xl, torch.int64), be.to_device(yl, torch.int64), ks)
    return be.to_host(hits).astype(np.float64) / float(len(xl))
