"""CPU oracle for the cluster-quality scores and the label hand-over.  TEST INFRASTRUCTURE ONLY (see
finch_oracle.py for the rules).

The reference computes NMI / AMI by calling scikit-learn directly (online_train.py:22 imports
normalized_mutual_info_score and adjusted_mutual_info_score from sklearn.metrics; call sites :634, :640), so the
oracle is the same third-party call (scikit-learn: requirements.txt:5 pins 0.22.0, 1.9.0 installed here - both use
average_method='arithmetic' by default).  The unshuffle loop and the text format restate online_train.py:648-658
and datasets/ucf101.py:124-134.
"""
import numpy as np
from sklearn import metrics as skm


def normalized_mutual_info_score(labels_true, labels_pred):
    return float(skm.normalized_mutual_info_score(labels_true, labels_pred))      # online_train.py:634


def adjusted_mutual_info_score(labels_true, labels_pred):
    return float(skm.adjusted_mutual_info_score(labels_true, labels_pred))        # online_train.py:640


def mutual_info_score(labels_true, labels_pred):
    return float(skm.mutual_info_score(labels_true, labels_pred))


def unshuffled_assignments(cluster_labels, idxs, dataset_len):
    """online_train.py:648-652, as written (sequential: a repeated index keeps the last label)."""
    out = [None] * dataset_len
    for i in range(len(cluster_labels)):
        out[idxs[i]] = cluster_labels[i]
    return out


def write_vid_clusters(path, assignments):
    """online_train.py:654-658."""
    with open(path, "w") as f:
        for label in assignments:
            f.write('{}\n'.format(label))


def read_cluster_labels(path):
    """datasets/ucf101.py:128-131."""
    with open(path, 'r') as f:
        cluster_labels = f.readlines()
    return [int(id.replace('\n', '')) for id in cluster_labels]
