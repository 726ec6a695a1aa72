"""Cluster-label hand-over to the data loaders.

  online_train.py:644-658   cluster labels -> unshuffled dataset order -> `vid_clusters.txt` (one label per line)
  datasets/ucf101.py:124-134, datasets/kinetics.py:88-96   read_cluster_labels
  datasets/triplets_dataset.py:99-104                      label -> row-indices table (clustering.cluster_masks.label_to_indices)

The scatter back to dataset order runs on the device (slic_scatter_last_wins: where the distributed sampler
repeated a dataset index, the last occurrence wins - as the reference's sequential loop); the text file format is
kept byte for byte so the reference's own readers can consume it, and a `.npy` side path avoids the text round trip.
"""
import os

import numpy as np
import torch

from . import backend as _backend

UNASSIGNED = -1


def unshuffled_assignments(cluster_labels, idxs, dataset_len, backend=None):
    """online_train.py:648-652.  cluster_labels [n] (row i of the embedding matrix), idxs [n] (its dataset index).
    -> int32 numpy [dataset_len]; slots no row maps to hold UNASSIGNED (the reference leaves None there)."""
    be = backend or _backend.default_backend()
    lab = be.to_device(torch.as_tensor(np.asarray(cluster_labels) if not isinstance(cluster_labels, torch.Tensor)
                                       else cluster_labels.detach()), torch.int32)
    pos = be.to_device(torch.as_tensor(np.asarray(idxs) if not isinstance(idxs, torch.Tensor) else idxs.detach()),
                       torch.int64)
    if lab.shape[0] != pos.shape[0]:
        raise ValueError("cluster_labels and idxs must have one entry per embedding row")
    out, bad = be.scatter_last_wins(lab, pos, int(dataset_len), fill=UNASSIGNED)
    if bad:
        raise IndexError("%d dataset indices outside [0, %d)" % (bad, dataset_len))   # list assignment index out of range
    return be.to_host(out)


def write_vid_clusters(path, assignments):
    """online_train.py:654-658: one label per line, '{}\\n'.format(label); an unassigned slot is written as the
    reference writes it ('None')."""
    arr = np.asarray(assignments)
    lines = np.char.add(arr.astype(np.int64).astype(str), "\n")
    if (arr == UNASSIGNED).any():
        lines[arr == UNASSIGNED] = "None\n"
    with open(path, "w") as f:
        f.write("".join(lines.tolist()))
    print('Saved cluster labels to', path)


def read_cluster_labels(path, is_master_proc=False):
    """datasets/ucf101.py:124-134 / datasets/kinetics.py:88-96: list of ints, or None without a path.
    (A 'None' line raises ValueError, as int('None') does in the reference.)"""
    if not path:
        if is_master_proc:
            print('cluster_path not defined....')
        return None
    with open(path, 'r') as f:
        cluster_labels = [int(tok) for tok in f.read().split()]
    if is_master_proc:
        print('retrieved {} cluster id from file: {}'.format(len(cluster_labels), path))
    return cluster_labels


def save_cluster_labels_npy(path, assignments):
    """Binary side path: int32 .npy next to (or instead of) vid_clusters.txt."""
    np.save(path, np.asarray(assignments, dtype=np.int32))


def load_cluster_labels(path):
    """.npy or the reference's text format, by extension."""
    if os.path.splitext(path)[1] == ".npy":
        return np.load(path).astype(np.int32)
    return np.asarray(read_cluster_labels(path), dtype=np.int32)
