"""normalized_mutual_info_score / adjusted_mutual_info_score on the device - the two scikit-learn calls
online_train.py:633-642 makes on the true labels and the FINCH labels after every clustering
(SURVEY.md section 8f, rank 3).  Same names, argument order and defaults as sklearn.metrics.

The contingency counts, the mutual information, the entropies and the expected mutual information come from
one C-ABI call (slic_cluster_metrics, csrc/metrics.cu); the ratios and sklearn's special cases below are host
arithmetic on six float64 numbers (sklearn/metrics/cluster/_supervised.py).
"""
import numpy as np
import torch

from . import backend as _backend

_EPS = float(np.finfo("float64").eps)


_I32 = (-(1 << 31), (1 << 31) - 1)


def _dense(be, labels):
    """np.unique(labels, return_inverse=True)[1] (sklearn check_clusterings + contingency_matrix): labels of any
    integer values -> dense ids, computed on the device (slic_dense_labels); returns (int32 [n], distinct values).
    Integer labels must fit int32 (class and cluster ids do); other label types are rejected."""
    if isinstance(labels, torch.Tensor):
        t = labels.detach().reshape(-1)
        if t.dtype.is_floating_point or t.dtype == torch.bool:
            raise TypeError("labels must be integers")
        if t.dtype == torch.int64 and t.numel() and (int(t.min()) < _I32[0] or int(t.max()) > _I32[1]):
            raise ValueError("integer labels must fit int32")
        dev = be.to_device(t, torch.int32)
    else:
        arr = np.asarray(labels).reshape(-1)
        if arr.dtype.kind not in "iu":
            raise TypeError("labels must be integers")
        if arr.size and (int(arr.min()) < _I32[0] or int(arr.max()) > _I32[1]):
            raise ValueError("integer labels must fit int32")
        dev = be.to_device(arr.astype(np.int32, copy=False), torch.int32)
    if dev.shape[0] == 0:
        return dev, 0
    dense, _, count = be.dense_labels(dev)
    return dense, count


def _generalized_average(u, v, average_method):
    """sklearn _supervised.py _generalized_average."""
    if average_method == "min":
        return min(u, v)
    if average_method == "geometric":
        return float(np.sqrt(u * v))
    if average_method == "arithmetic":
        return float(np.mean([u, v]))
    if average_method == "max":
        return max(u, v)
    raise ValueError("'average_method' must be 'min', 'geometric', 'arithmetic', or 'max'")


def cluster_scores(labels_true, labels_pred, want_emi=True, backend=None):
    """-> dict(mi, h_true, h_pred, emi, classes, clusters, n) for two labelings of the same rows."""
    be = backend or _backend.default_backend()
    lt, nt = _dense(be, labels_true)
    lp, npred = _dense(be, labels_pred)
    if lt.shape[0] != lp.shape[0]:
        raise ValueError("labels_true and labels_pred must have same size, got %d and %d" % (lt.shape[0], lp.shape[0]))
    if lt.shape[0] == 0:
        return dict(mi=0.0, h_true=0.0, h_pred=0.0, emi=0.0, classes=0, clusters=0, n=0)
    mi, ht, hp, emi, classes, clusters = be.cluster_metrics(lt, lp, nt, npred, want_emi)
    return dict(mi=mi, h_true=ht, h_pred=hp, emi=emi, classes=int(classes), clusters=int(clusters), n=int(lt.shape[0]))


def mutual_info_score(labels_true, labels_pred, backend=None):
    return cluster_scores(labels_true, labels_pred, want_emi=False, backend=backend)["mi"]


def normalized_mutual_info_score(labels_true, labels_pred, average_method="arithmetic", backend=None):
    """sklearn.metrics.normalized_mutual_info_score (online_train.py:634)."""
    s = cluster_scores(labels_true, labels_pred, want_emi=False, backend=backend)
    if (s["classes"] == s["clusters"] == 1) or (s["classes"] == s["clusters"] == 0):
        return 1.0                                   # a single cluster on both sides: perfect match by convention
    if abs(s["mi"]) < _EPS:                          # mi = 0 cannot be a perfect match here
        return 0.0
    normalizer = _generalized_average(s["h_true"], s["h_pred"], average_method)
    return float(s["mi"] / normalizer)


def adjusted_mutual_info_score(labels_true, labels_pred, average_method="arithmetic", backend=None):
    """sklearn.metrics.adjusted_mutual_info_score (online_train.py:640)."""
    s = cluster_scores(labels_true, labels_pred, want_emi=True, backend=backend)
    if (s["classes"] == s["clusters"] == 1) or (s["classes"] == s["clusters"] == 0):
        return 1.0
    normalizer = _generalized_average(s["h_true"], s["h_pred"], average_method)
    denominator = normalizer - s["emi"]
    # guard against a denominator that rounding pushed to (or across) zero, with sklearn's sign convention
    if denominator < 0:
        denominator = min(denominator, -_EPS)
    else:
        denominator = max(denominator, _EPS)
    return float((s["mi"] - s["emi"]) / denominator)
