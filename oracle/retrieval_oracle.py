"""CPU oracle for the retrieval half of the hot path.  TEST INFRASTRUCTURE ONLY.

/root/reference/evaluate.py and iic_retrieve_clips.py cannot be imported (matplotlib, fvcore,
skvideo are absent), so these are restatements of the few numpy/sklearn lines on the path;
every function cites the lines it follows.  Ties inside a row are NOT pinned by the reference
(np.argpartition / quicksort argsort are unstable) - tests compare on rows whose k / k+1
boundary gap exceeds the margin stated in the test, plus exact hit counts.
"""
import numpy as np
from sklearn.metrics.pairwise import cosine_distances, euclidean_distances


def distance_matrix(x, y=None, dist_metric="cosine"):
    """evaluate.py:208-223 (get_distance_matrix)."""
    assert dist_metric in ["cosine", "euclidean"]
    if dist_metric == "cosine":
        dm = cosine_distances(x, Y=y)
    else:
        dm = euclidean_distances(x, Y=y)
    if y is None:
        np.fill_diagonal(dm, float("inf"))
    return dm


def closest_data_mat(dm, top_k):
    """evaluate.py:226-231 (get_closest_data_mat): k smallest per row, ascending."""
    part = np.argpartition(dm, top_k, axis=-1)
    vals = np.take_along_axis(dm, part[:, :top_k], axis=-1)
    return np.take_along_axis(part, np.argsort(vals, axis=-1), axis=-1)


def closest_data(dm, exemplar_idx, top_k):
    """evaluate.py:234-238 (get_closest_data)."""
    row = dm[exemplar_idx]
    part = np.argpartition(row, top_k)
    return part[np.argsort(row[part[:top_k]])]


def topk_acc(dm, x_labels, y_labels=None, top_ks=(1, 5, 10, 20)):
    """evaluate.py:287-307 (get_topk_acc): hit@k = query label among labels of its k nearest."""
    idx = closest_data_mat(dm, top_k=top_ks[-1])
    if y_labels is None:
        y_labels = x_labels
    y_labels = np.asarray(y_labels)
    acc = []
    for i, xl in enumerate(x_labels):
        acc.append([int(xl in y_labels[idx[i, :k]]) for k in top_ks])
    return np.mean(np.array(acc), axis=0)


def topk_retrieval_counts(x_train, y_train, x_test, y_test, ks=(1, 5, 10, 20, 50)):
    """iic_retrieve_clips.py:275-306 (topk_retrieval) on the raw [videos, clips, D] arrays:
    mean over clips, cosine distances test x train, full argsort, hit counts per k."""
    x_train = np.mean(x_train, 1).reshape((-1, x_train.shape[-1]))
    y_train = y_train[:, 0].reshape(-1)
    x_test = np.mean(x_test, 1).reshape((-1, x_test.shape[-1]))
    y_test = y_test[:, 0].reshape(-1)
    d = cosine_distances(x_test, x_train)
    order = np.argsort(d)
    correct = {}
    for k in ks:
        lab = y_train[order[:, :k]]
        correct[k] = int(np.sum((lab == y_test[:, None]).any(axis=1)))
    return correct, order, d


def boundary_gaps(d_sorted_vals, ks):
    """gap between the k-th and (k+1)-th smallest distance per row, for the tie-margin filter."""
    return {k: d_sorted_vals[:, k] - d_sorted_vals[:, k - 1] for k in ks}


def coclr_nn_accuracy(test_feature, test_label, train_feature, train_label, ks=(1, 5, 10, 20, 50)):
    """coclr_classify.py:784-810 restated with the same torch (CPU) calls: centring, F.normalize, matmul, torch.topk per
    k, kNN accuracy.  Returns (accs, sim) - sim for tie-margin bookkeeping in the tests."""
    import torch
    import torch.nn.functional as F
    test_feature = torch.as_tensor(test_feature, dtype=torch.float32)
    train_feature = torch.as_tensor(train_feature, dtype=torch.float32)
    test_label, train_label = torch.as_tensor(test_label), torch.as_tensor(train_label)
    test_feature = test_feature - test_feature.mean(dim=0, keepdim=True)
    train_feature = train_feature - train_feature.mean(dim=0, keepdim=True)
    test_feature = F.normalize(test_feature, p=2, dim=1)
    train_feature = F.normalize(train_feature, p=2, dim=1)
    sim = test_feature.matmul(train_feature.t())
    accs = []
    for k in ks:
        _, topkidx = torch.topk(sim, k, dim=1)
        accs.append(torch.any(train_label[topkidx] == test_label.unsqueeze(1), dim=1).float().mean().item())
    return accs, sim.numpy()


def pdist_v2(vector1, vector2, eps, dist_metric):
    """loss/triplet_loss.py:438-447, same torch calls."""
    import torch
    import torch.nn.functional as F
    vector1, vector2 = torch.as_tensor(vector1), torch.as_tensor(vector2)
    rows = []
    for i in range(len(vector1)):
        if dist_metric == 'euclidean':
            rows.append(F.pairwise_distance(vector1[i], vector2, eps=eps).unsqueeze(0))
        else:
            rows.append(1 - F.cosine_similarity(vector1[i].unsqueeze(0), vector2, dim=1).unsqueeze(0))
    return torch.cat(rows, dim=0).numpy()
