"""Embedding hand-off without the host round trip.

The reference collects the per-batch embeddings of every rank with an all-gather, moves each batch to the
CPU, concatenates there (evaluate.py:189-201) and later converts the [N, D] matrix to numpy for FINCH
(clustering/cluster_masks.py:80) - 2 x N x D x 4 bytes over PCIe and a Python list of N / B host tensors.
`EmbeddingCollector` keeps the same call shape and the SAME ROW ORDER (batch by batch, ranks interleaved
inside a batch exactly as du_helper.all_gather concatenates them), but the rows stay in one preallocated
device matrix that FINCH / fit_cluster / the sharded search take as they are.

    col = EmbeddingCollector(len(loader.dataset), dim, group=None)      # capacity is an upper bound
    for inputs, targets, indexes in loader:                             # evaluate.py:170-191
        col.append(encoder(inputs).flatten(1), targets, indexes)
    embeddings, labels, idxs = col.finish()                             # device [N, D], host lists (evaluate.py:198-200)

torch provides device memory and the collective (NCCL on GPUs, gloo in the CPU tests); there is no numerical
work here.
"""
import torch
import torch.distributed as dist


class EmbeddingCollector:
    def __init__(self, capacity, dim, device=None, dtype=torch.float32, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        # the distributed sampler pads the last batch: leave room for one padded batch per rank
        self.capacity = int(capacity)
        self.emb = torch.empty((self.capacity, dim), dtype=dtype, device=self.device)
        self.labels = torch.empty(self.capacity, dtype=torch.int64, device=self.device)
        self.idxs = torch.empty(self.capacity, dtype=torch.int64, device=self.device)
        self.rows = 0

    def _grow(self, need):
        cap = max(need, self.capacity * 2)
        for name in ("emb", "labels", "idxs"):
            old = getattr(self, name)
            new = torch.empty((cap,) + tuple(old.shape[1:]), dtype=old.dtype, device=old.device)
            new[: self.rows] = old[: self.rows]
            setattr(self, name, new)
        self.capacity = cap

    def append(self, embedd, targets, indexes):
        """One batch of this rank (evaluate.py:180-191): embeddings [b, D], targets [b], dataset indexes [b].
        With world > 1 the batch is all-gathered first - rank-major inside the batch, as torch.cat of
        du_helper.all_gather's list - and stays on the device."""
        embedd = embedd.detach().to(self.device, self.emb.dtype).contiguous()
        targets = torch.as_tensor(targets).detach().to(self.device, torch.int64).contiguous()
        indexes = torch.as_tensor(indexes).detach().to(self.device, torch.int64).contiguous()
        b = embedd.shape[0]
        total = b * self.world
        if self.rows + total > self.capacity:
            self._grow(self.rows + total)
        r0 = self.rows
        if self.world == 1:
            self.emb[r0:r0 + b] = embedd
            self.labels[r0:r0 + b] = targets
            self.idxs[r0:r0 + b] = indexes
        else:
            # gathered straight into the destination rows: no intermediate list, no concatenation
            dist.all_gather_into_tensor(self.emb[r0:r0 + total], embedd, group=self.group)
            dist.all_gather_into_tensor(self.labels[r0:r0 + total], targets, group=self.group)
            dist.all_gather_into_tensor(self.idxs[r0:r0 + total], indexes, group=self.group)
        self.rows += total

    def finish(self, labels_on_host=True):
        """-> (embeddings [N, D] on the device, labels, idxs).  labels / idxs as Python lists like
        evaluate.py:199-200 (they index host-side dataset tables), or device tensors with labels_on_host=False."""
        emb = self.emb[: self.rows]
        if labels_on_host:
            both = torch.stack((self.labels[: self.rows], self.idxs[: self.rows])).cpu()
            return emb, both[0].tolist(), both[1].tolist()
        return emb, self.labels[: self.rows], self.idxs[: self.rows]
