"""numpy stand-in for backend.CudaBackend - TEST INFRASTRUCTURE ONLY.

Lets the host-side logic (the FINCH level loop and exit rules, fit_cluster, the evaluate / iic
wrappers, the row-sharded search under gloo) run on a box without a GPU.  It is never imported by the
package; the product path constructs CudaBackend, which raises without a CUDA device.
Each method follows the documented contract of the C-ABI call it stands in for.
"""
import numpy as np
import torch
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class FakeBackend:
    name = "fake-numpy"
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def to_device(self, array, dtype=None):
        t = torch.as_tensor(array)
        return (t.to(dtype) if dtype is not None else t).contiguous()

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def to_host(self, t):
        return t.cpu().numpy()

    # K1
    def normalize_rows(self, x, want_f16=True):
        a = _np(x)
        nrm = np.sqrt(np.einsum("ij,ij->i", a.astype(np.float64), a.astype(np.float64))).astype(a.dtype)
        nrm[nrm == 0] = 1
        unit = torch.from_numpy(a / nrm[:, None])
        return unit, (unit.to(torch.float16) if want_f16 else None)

    def center_columns(self, x):
        a = _np(x)
        return torch.from_numpy(a - a.astype(np.float64).mean(0).astype(np.float32)[None, :])

    def _sims(self, q, x):
        return _np(q).astype(np.float64) @ _np(x).astype(np.float64).T

    def nn_exact_top1(self, q_unit, x_unit, self_offset=-1, q_rows=None):
        q = _np(q_unit) if q_rows is None else _np(q_unit)[_np(q_rows)]
        s = self._sims(q, x_unit)
        if self_offset >= 0:
            rows = np.arange(len(q)) if q_rows is None else _np(q_rows)
            s[np.arange(len(q)), rows + self_offset] = -np.inf
        dt = _np(x_unit).dtype
        dmat = np.clip(dt.type(1) - s.astype(dt), 0, 2).astype(dt)     # distances in the reference dtype
        dmat[np.isneginf(s)] = np.inf
        j = np.argmin(dmat, axis=1)                                      # ties -> lowest index (np.argmin)
        return torch.from_numpy(j.astype(np.int32)), torch.from_numpy(dmat[np.arange(len(q)), j])

    def nn_top1(self, q_unit, q_f16, x_unit, x_f16, self_offset=-1, eps=0.0):
        return self.nn_exact_top1(q_unit, x_unit, self_offset)

    def first_neighbors(self, x, row_range=None):
        self.calls.append(("first_neighbors", tuple(x.shape), str(x.dtype), row_range))
        unit, _ = self.normalize_rows(x, want_f16=False)
        r0, r1 = (0, x.shape[0]) if row_range is None else row_range
        nn, d = self.nn_exact_top1(unit[r0:r1], unit, self_offset=r0)
        return nn, d, unit

    # multi-process share of the symmetric self-search (slic_nn_top1_sym_part / slic_unpack_neighbor_keys)
    PART_BLOCK = 64
    SYM_MIN_ROWS = 256          # (the CUDA backend: 16384)
    KEY_NONE = 0x7fffffff7fffffff

    def supports_triangle_parts(self, x):
        return x.dtype == torch.float32 and x.shape[0] >= self.SYM_MIN_ROWS

    def first_neighbors_part(self, x, part, parts, reduce_max=None):
        """Block pairs (bi <= bj) of the distance matrix are dealt round-robin to the parts; a part evaluates its pairs
        in both directions and keeps, per row, the smallest (distance bits, neighbour) key.  reduce_max: the row-best
        exchange of the two-phase search - exercised with this part's rows' best similarity to a column sample."""
        self.calls.append(("first_neighbors_part", tuple(x.shape), part, parts))
        unit, _ = self.normalize_rows(x, want_f16=False)
        u = _np(unit)
        n, b = len(u), self.PART_BLOCK
        if reduce_max is not None and parts > 1:
            lo = np.iinfo(np.int32).min
            bests = np.full(n, lo, dtype=np.int32)
            r0, r1 = n * part // parts, n * (part + 1) // parts
            s = (u[r0:r1].astype(np.float64) @ u[::7].astype(np.float64).T).astype(np.float32)
            s[s > 0.99999] = -1.0                                       # (the row itself, if sampled)
            enc = s.max(axis=1).view(np.uint32).astype(np.int64)
            enc = np.where(enc & 0x80000000, ~enc & 0xffffffff, enc | 0x80000000) ^ 0x80000000
            bests[r0:r1] = enc.astype(np.uint32).view(np.int32)
            t = torch.from_numpy(bests)
            reduce_max(t)
            self.exchanged_bests = t
            assert int((t == lo).sum()) == 0, "row bests of some rows were not delivered by the exchange"
        keys = np.full(n + 1, self.KEY_NONE, dtype=np.int64)
        keys[n] = 1
        nb = (n + b - 1) // b
        t = 0
        for bi in range(nb):
            for bj in range(bi, nb):
                t += 1
                if (t - 1) % parts != part:
                    continue
                ri, rj = np.arange(bi * b, min(n, bi * b + b)), np.arange(bj * b, min(n, bj * b + b))
                s = u[ri].astype(np.float64) @ u[rj].astype(np.float64).T
                dm = np.clip(np.float32(1) - s.astype(np.float32), 0, 2).astype(np.float32)
                for rows, cols, m in ((ri, rj, dm), (rj, ri, dm.T)):
                    k = (m.view(np.uint32).astype(np.int64) << 32) | cols[None, :].astype(np.int64)
                    k[rows[:, None] == cols[None, :]] = self.KEY_NONE
                    keys[rows] = np.minimum(keys[rows], k.min(axis=1))
        return torch.from_numpy(keys), unit

    def unpack_neighbor_keys(self, keys):
        k = _np(keys)
        n = len(k) - 1
        nn = (k[:n] & 0xffffffff).astype(np.int32)
        d = (k[:n] >> 32).astype(np.uint32).view(np.float32)
        complete = bool(k[n] == 1 and not (k[:n] == self.KEY_NONE).any())
        return torch.from_numpy(nn), torch.from_numpy(d.copy()), complete

    # label hand-over / cluster-quality scores (slic_scatter_last_wins, slic_cluster_metrics)
    def dense_labels(self, labels):
        uniq, inv = np.unique(_np(labels), return_inverse=True)
        return torch.from_numpy(inv.astype(np.int32)), torch.from_numpy(uniq.astype(np.int32)), len(uniq)

    def scatter_last_wins(self, values, positions, n_out, fill=-1):
        v, p = _np(values), _np(positions)
        out = np.full(n_out, fill, dtype=np.int32)
        ok = (p >= 0) & (p < n_out)
        for i in np.flatnonzero(ok):
            out[p[i]] = v[i]
        return torch.from_numpy(out), int((~ok).sum())

    def cluster_metrics(self, labels_true, labels_pred, num_true, num_pred, want_emi=True):
        from sklearn.metrics.cluster._expected_mutual_info_fast import expected_mutual_information
        import scipy.sparse as sp
        lt, lp = _np(labels_true).astype(np.int64), _np(labels_pred).astype(np.int64)
        n = len(lt)
        cont = np.zeros((num_true, num_pred), dtype=np.int64)
        np.add.at(cont, (lt, lp), 1)
        a, b = cont.sum(1), cont.sum(0)

        def ent(c):
            c = c[c > 0].astype(np.float64)
            return 0.0 if c.size <= 1 else float(-np.sum((c / n) * (np.log(c) - np.log(n))))
        i, j = np.nonzero(cont)
        nij = cont[i, j].astype(np.float64)
        t = (nij / n) * (np.log(nij) - np.log(n)) + (nij / n) * (-np.log((a[i] * b[j]).astype(np.float64)) + 2 * np.log(n))
        t[np.abs(t) < np.finfo(np.float64).eps] = 0.0
        mi = max(float(t.sum()), 0.0)
        emi = float(expected_mutual_information(sp.csr_matrix(cont), n)) if want_emi else 0.0
        return [mi, ent(a), ent(b), emi, float((a > 0).sum()), float((b > 0).sum())]

    def distance_matrix(self, q, x, metric="cosine", same=False):
        dt = _np(x).dtype
        if metric == "cosine":
            m = np.clip(dt.type(1) - self._sims(q, x).astype(dt), 0, 2).astype(dt)
        else:
            qa, xa = _np(q).astype(np.float64), _np(x).astype(np.float64)
            d2 = (qa * qa).sum(1)[:, None] + (xa * xa).sum(1)[None, :] - 2 * qa @ xa.T
            m = np.sqrt(np.maximum(d2, 0)).astype(dt)
        if same:
            np.fill_diagonal(m, 0)
        return torch.from_numpy(m)

    def rows_topk(self, mat, k):
        m = _np(mat)
        order = np.lexsort((np.broadcast_to(np.arange(m.shape[1]), m.shape), m), axis=1)[:, :k]
        return torch.from_numpy(order.astype(np.int32)), torch.from_numpy(np.take_along_axis(m, order, 1))

    def topk_cosine(self, q_unit, x_unit, k, self_offset=-1, q_f16=None, x_f16=None, eps=0.0):
        m = _np(self.distance_matrix(q_unit, x_unit))
        if self_offset >= 0:
            m[np.arange(m.shape[0]), np.arange(m.shape[0]) + self_offset] = np.inf
        return self.rows_topk(torch.from_numpy(m), k)

    def topk_neighbors(self, q, x, k, same=False):
        ux, _ = self.normalize_rows(x, want_f16=False)
        uq = ux if same else self.normalize_rows(q, want_f16=False)[0]
        return self.topk_cosine(uq, ux, k, self_offset=0 if same else -1)

    def hit_at_k(self, topk_idx, q_labels, x_labels, ks):
        idx, ql, xl = _np(topk_idx), _np(q_labels), _np(x_labels)
        return torch.tensor([int(((xl[idx[:, :k]] == ql[:, None]).any(1)).sum()) for k in ks], dtype=torch.int32)

    # K2
    def _links(self, nn, unit=None):
        nn = _np(nn).astype(np.int64)
        n = len(nn)
        i = np.arange(n)
        rows, cols, kind = [i], [nn], [np.zeros(n, int)]
        order = np.argsort(nn, kind="stable")
        s = nn[order]
        for h in np.unique(s):
            mem = order[s == h]
            if len(mem) > 1:
                a, b = np.triu_indices(len(mem), 1)
                rows.append(mem[a]); cols.append(mem[b]); kind.append(np.ones(len(a), int))
        return np.concatenate(rows), np.concatenate(cols), np.concatenate(kind)

    def _pair_dist(self, unit, r, c):
        u = _np(unit)
        s = np.einsum("ij,ij->i", u[r].astype(np.float64), u[c].astype(np.float64))
        return np.clip(u.dtype.type(1) - s.astype(u.dtype), 0, 2).astype(u.dtype)

    def components(self, nn, min_sim=None, unit=None, dist=None):
        nnv = _np(nn).astype(np.int64)
        n = len(nnv)
        if min_sim is None:
            r, c = np.arange(n), nnv
        else:
            r, c, kind = self._links(nn)
            d = np.where(kind == 0, _np(dist)[r].astype(np.float64), self._pair_dist(unit, r, c).astype(np.float64))
            w = np.where((kind == 0) & (nnv[c] == r), 2.0, 1.0)
            keep = ~(d * w > float(min_sim)) & (r != c)
            r, c = r[keep], c[keep]
        g = csr_matrix((np.ones(len(r)), (r, c)), shape=(n, n))
        num, lab = connected_components(g, directed=True, connection="weak")
        return torch.from_numpy(lab.astype(np.int32)), int(num)

    def min_sim(self, nn, unit, dist):
        nnv = _np(nn).astype(np.int64)
        r, c, kind = self._links(nn)
        d = np.where(kind == 0, _np(dist)[r], self._pair_dist(unit, r, c)).astype(np.float32)
        w = np.where((kind == 0) & (nnv[c] == r), 2.0, 1.0).astype(np.float32)
        keep = r != c
        return np.float32(np.max(d[keep] * w[keep]))

    def closest_link(self, nn, unit, dist):
        r, c, kind = self._links(nn)
        d = np.where(kind == 0, _np(dist)[r].astype(np.float64), self._pair_dist(unit, r, c).astype(np.float64))
        d[r == c] = np.inf
        m = d.min()
        cand = [(min(a, b), max(a, b)) for a, b in zip(r[d == m], c[d == m])]
        i, j = min(cand)
        return int(i), int(j)

    # K3
    def compose_labels(self, prev, u):
        return u.clone() if prev is None else torch.from_numpy(_np(u)[_np(prev).astype(np.int64)])

    def segmented_mean(self, data, labels, num_clust):
        a, lab = _np(data).astype(np.float64), _np(labels).astype(np.int64)
        out = np.zeros((num_clust, a.shape[1]))
        np.add.at(out, lab, a)
        return torch.from_numpy(out / np.bincount(lab, minlength=num_clust)[:, None])

    # K4
    def cluster_sums(self, data, labels, num_clust):
        x, lab = _np(data).astype(np.float64), _np(labels).astype(np.int64)
        sums = np.zeros((num_clust, x.shape[1]))
        np.add.at(sums, lab, x)
        counts = np.bincount(lab, minlength=num_clust).astype(np.int32)
        return torch.from_numpy(sums), torch.from_numpy(counts), torch.from_numpy(sums / counts[:, None])

    def merge_cluster_sums(self, sums_prev, counts_prev, u, num_clust):
        sp, cp, lab = _np(sums_prev), _np(counts_prev).astype(np.int64), _np(u).astype(np.int64)
        sums = np.zeros((num_clust, sp.shape[1]))
        np.add.at(sums, lab, sp)
        counts = np.bincount(lab, weights=cp, minlength=num_clust).astype(np.int32)
        return torch.from_numpy(sums), torch.from_numpy(counts), torch.from_numpy(sums / counts[:, None])

    def label_mask(self, a, b, prepend_ones=False, negate=False):
        m = (_np(a)[:, None] == _np(b)[None, :]) != negate
        if prepend_ones:
            m = np.concatenate([np.ones((m.shape[0], 1), bool), m], 1)
        return torch.from_numpy(m)

    def label_mask_bits(self, a, b, negate=False):
        m = (_np(a)[:, None] == _np(b)[None, :]) != negate
        pad = (-m.shape[1]) % 32
        m = np.pad(m, ((0, 0), (0, pad)))
        words = np.packbits(m.reshape(m.shape[0], -1, 32), axis=-1, bitorder="little").view(np.uint32)[..., 0]
        return torch.from_numpy(words.astype(np.int64).astype(np.uint32).view(np.int32))

    def group_by_label(self, labels, num_labels):
        lab = _np(labels).astype(np.int64)
        order = np.argsort(lab, kind="stable").astype(np.int32)
        off = np.zeros(num_labels + 1, np.int32)
        np.cumsum(np.bincount(lab, minlength=num_labels), out=off[1:])
        return torch.from_numpy(order), torch.from_numpy(off)
