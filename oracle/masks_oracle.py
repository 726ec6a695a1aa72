"""CPU oracle for the label-equality masks that consume FINCH labels.  TEST INFRASTRUCTURE ONLY.

The reference has no mask module; these restate the four call sites."""
import numpy as np


def queue_positive_mask(k_label, queue_label):
    """models/infoNCE.py:281-283 (UberNCE): [B, 1+K] bool, column 0 all ones."""
    k_label = np.asarray(k_label)
    queue_label = np.asarray(queue_label)
    m = k_label[:, None] == queue_label[None, :]
    return np.concatenate([np.ones((m.shape[0], 1), dtype=bool), m], axis=1)


def in_batch_masks(labels, other=None):
    """loss/triplet_loss.py:136-142, 291-297 (and :254-261 with a memory-bank `other`):
    for every unique label, positives = where(labels == label), negatives = where(~(other == label)).
    Returned as the dense [U, L] positive mask, its complement over `other`, and the unique labels."""
    labels = np.asarray(labels)
    other = labels if other is None else np.asarray(other)
    uniq = np.unique(labels)
    pos = uniq[:, None] == labels[None, :]
    neg = ~(uniq[:, None] == other[None, :])
    return uniq, pos, neg


def label_to_indices(data_labels, label_set=None):
    """datasets/triplets_dataset.py:99-104: {label: np.where(data_labels == label)[0]}."""
    data_labels = np.asarray(data_labels)
    if label_set is None:
        label_set = set(data_labels.tolist())
    return {lab: np.where(data_labels == lab)[0] for lab in label_set}


def group_by_label(labels, num_labels=None):
    """CSR form of label_to_indices for dense labels 0..C-1: (order, offsets), ascending index
    inside every group (what np.where yields)."""
    labels = np.asarray(labels)
    c = int(labels.max()) + 1 if num_labels is None else num_labels
    order = np.argsort(labels, kind="stable").astype(np.int32)
    offsets = np.zeros(c + 1, dtype=np.int32)
    np.cumsum(np.bincount(labels, minlength=c), out=offsets[1:])
    return order, offsets
