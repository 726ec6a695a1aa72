"""Multi-GPU level-0 first-neighbour search across the ranks of one box, every rank holding the full
embedding matrix.  Two ways to share the O(N^2 D) stage:
  * triangle parts (large float32 self-searches, the default): the symmetric screen computes only the tiles on
    or right of the diagonal; rank r takes a contiguous 1/world share of that triangle and the per-row (distance,
    neighbour) keys are merged by ONE all-reduce (MIN) over NCCL / NVLink - 8 (N + 1) bytes;
  * row shards (small inputs, retrieval, fallback): rank r searches rows [r * ceil(N / G), ...) against all
    columns; ids and distances (which the min_sim filter needs) are all-gathered.
Levels >= 1 (n_1 << N), the components and the means are serial work of a few milliseconds and run
redundantly on every rank, so all ranks return the same partition without a broadcast.

One process per GPU (torch.distributed, backend nccl on GPUs; gloo works for the CPU tests with the
stand-in backend).  The reference runs FINCH on rank 0 only while the other ranks wait at a barrier
(online_train.py:619-627, 660-662); here those ranks do a 1/world share of the O(N^2 D) stage.
"""
import torch
import torch.distributed as dist

from . import backend as _backend
from .clustering.finch import FINCH


class PeerWindowsUnavailable(RuntimeError):
    """Raised on EVERY rank when some rank could not set up its peer window; callers fall back to the NCCL scheme."""


class PeerGroup:
    """NVLink peer windows of the ranks of a process group (one process per GPU; csrc/comm.cu).  The 64-byte CUDA IPC
    handles travel through the process group once, at construction; after that the level-0 search of the group needs
    no collective call at all - the ranks publish into and read from each other's windows from inside their kernels."""

    def __init__(self, be, group=None, max_rows=1 << 20):
        self.be, self.group, self.max_rows = be, group, int(max_rows)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        # Every step is collective-safe: a rank that cannot create, export or map a window (no peer access between the
        # devices, CUDA IPC not permitted in this container, ...) still takes part in the exchanges below, and ALL ranks
        # learn of the failure from one all-reduce - nobody is left waiting in a barrier.
        self.comm, error = None, None
        try:
            self.comm, handle = be.comm_window_create(self.max_rows)
        except Exception as e:             # noqa: BLE001 - reported below, on every rank
            handle, error = bytes(64), e
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=be.device)
        everyone = torch.empty(64 * self.world, dtype=torch.uint8, device=be.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        if error is None:
            try:
                be.comm_connect(self.comm, self.rank, self.world, bytes(everyone.cpu().tolist()))
            except Exception as e:         # noqa: BLE001
                error = e
        ok = torch.tensor([0 if error is not None else 1], dtype=torch.int32, device=be.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # (also: every rank has mapped every window before anyone publishes)
        if int(ok.item()) == 0:
            if self.comm is not None:
                be.comm_destroy(self.comm)
                self.comm = None
            raise PeerWindowsUnavailable("NVLink peer windows could not be set up on every rank%s"
                                         % ("" if error is None else ": %s" % error))

    def first_neighbors(self, mat):
        """-> (nn, dist, unit, status): status is a device int32[2], non-zero = repeat the search another way."""
        return self.be.comm_first_neighbors(self.comm, mat)

    def close(self):
        if self.comm is not None:
            torch.cuda.synchronize(self.be.device)
            dist.barrier(group=self.group)  # nobody unmaps a window a peer's kernel may still be touching
            self.be.comm_destroy(self.comm)
            self.comm = None


def shard_range(n, rank, world):
    per = (n + world - 1) // world
    r0 = min(rank * per, n)
    return r0, min(r0 + per, n), per


def sharded_first_neighbors(be, group=None, timings=None, triangle=True, peer=True):
    """Returns a callable mat -> (nn, dist, unit) for FINCH(first_neighbors=...).
    triangle=True (default): large float32 self-searches are shared as parts of the symmetric screen's triangle;
      peer=True (default, up to 8 ranks of one box): through NVLink peer windows - one fused kernel per rank, no collective
        call on the data path (PeerGroup); peer=False: two screen launches with an NCCL all-reduce MAX of the row bests
        between them and an all-reduce MIN of the keys after (the round-1 scheme, also used by the CPU stand-in tests);
    otherwise (and for small inputs) query rows are sharded and the ids all-gathered."""
    def peer_group(n):
        # one set of windows per (backend, process group), shared by every search built on them and grown on demand;
        # None when the windows cannot be had on this box (decided once, identically on every rank)
        key = (id(be), id(group))
        if key in _PEER_UNAVAILABLE:
            return None
        pg = _PEER_GROUPS.get(key)
        if pg is not None and pg.max_rows < n:
            pg.close()
            _PEER_GROUPS.pop(key, None)
            pg = None
        if pg is None:
            try:
                pg = _PEER_GROUPS[key] = PeerGroup(be, group, max_rows=max(n, 1 << 18))
            except PeerWindowsUnavailable as e:
                import warnings
                warnings.warn("%s - falling back to the NCCL all-reduce scheme" % e)
                _PEER_UNAVAILABLE.add(key)
                return None
        return pg

    def search(mat):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        n = mat.shape[0]
        use_triangle = world > 1 and triangle and hasattr(be, "first_neighbors_part") and be.supports_triangle_parts(mat)
        pg = peer_group(n) if (use_triangle and peer and world <= 8 and hasattr(be, "comm_first_neighbors")) else None
        if pg is not None:
            nn, d, unit, status = pg.first_neighbors(mat)
            if not status.any().item():      # (the one host read-back of the stage; identical on every rank)
                return nn, d, unit
            use_triangle = False             # degenerate input (candidate log overflow): row-sharded full square below
        if use_triangle:
            # The score matrix of a self-search is symmetric: the ranks share the tiles on or right of its diagonal
            # (half the flops of the row-sharded full square) and every rank ends up with, for EVERY row, the best
            # neighbour among the pairs it saw, as (distance, neighbour) keys; every rank sees the same merged array,
            # hence takes the same branch below.
            # Two exchange steps: (1) all-reduce MAX of 4 N bytes of row bests - each rank screens 1 / G of the ROWS
            # against a sample of the columns first, so that every rank's candidate filter starts tight for all rows;
            # (2) all-reduce MIN of the 8 (N + 1) bytes of keys.
            keys, unit = be.first_neighbors_part(
                mat, rank, world, reduce_max=lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group))
            dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
            nn, d, complete = be.unpack_neighbor_keys(keys)
            if complete:
                return nn, d, unit
            # (a candidate log overflowed on some rank - degenerate input: repeat on the row-sharded full square)
        r0, r1, per = shard_range(n, rank, world)
        nn_loc, d_loc, unit = be.first_neighbors(mat, row_range=(r0, r1))
        if world == 1:
            return nn_loc, d_loc, unit
        nn_pad = torch.full((per,), -1, dtype=torch.int32, device=mat.device)
        d_pad = torch.zeros((per,), dtype=mat.dtype, device=mat.device)
        nn_pad[: r1 - r0] = nn_loc
        d_pad[: r1 - r0] = d_loc
        nn_all = torch.empty(per * world, dtype=torch.int32, device=mat.device)
        d_all = torch.empty(per * world, dtype=mat.dtype, device=mat.device)
        # the one exchange step of the path: 4 N bytes of ids + 4 N bytes of distances in total
        dist.all_gather_into_tensor(nn_all, nn_pad, group=group)
        dist.all_gather_into_tensor(d_all, d_pad, group=group)
        return nn_all[:n].contiguous(), d_all[:n].contiguous(), unit

    def native_comm(mat):
        """The peer-window handle for a whole-hierarchy call (slic_comm_finch) on this matrix, or None when the shared
        search does not apply (clustering/finch.py then falls back to search() + the single-GPU driver)."""
        world = dist.get_world_size(group)
        if (world > 1 and world <= 8 and triangle and peer and hasattr(be, "finch_native_comm")
                and be.supports_triangle_parts(mat)):
            pg = peer_group(mat.shape[0])
            return None if pg is None else pg.comm
        return None

    search.native_comm = native_comm
    return search


_PEER_GROUPS = {}
_PEER_UNAVAILABLE = set()


def close_peer_groups():
    """Unmap and free the NVLink peer windows (collective: every rank calls it; before destroy_process_group)."""
    for key in list(_PEER_GROUPS):
        _PEER_GROUPS.pop(key).close()


UPLOAD_CHUNKS = 4   # host -> device copy of a rank's shard in this many pieces, each all-gathered while the next one is copied


def upload_replicated(data, group=None, backend=None):
    """Host matrix held by EVERY rank -> full [N, D] float32 device matrix on every rank, with 1 / G of the PCIe
    traffic per rank: rank r copies rows [r * ceil(N / G), ...) host -> device, and all-gathers over NVLink assemble the
    matrix (the G ranks of a box share the host's memory and PCIe root, so G full uploads cost G times the bytes over
    the same links).  The shard travels in UPLOAD_CHUNKS pieces: the all-gather of piece c runs while piece c + 1 is
    still crossing PCIe (copy stream + current stream).  Device-resident input is returned as it is."""
    be = backend or _backend.default_backend()
    if isinstance(data, torch.Tensor) and data.device.type != "cpu":
        return be.to_device(data.detach(), torch.float32)
    host = data.detach() if isinstance(data, torch.Tensor) else torch.as_tensor(data)
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or host.dim() != 2:
        return be.to_device(host, torch.float32)
    if host.dtype != torch.float32:
        host = host.to(torch.float32)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, d = host.shape
    r0, r1, per = shard_range(n, rank, world)
    full = torch.empty((world, per, d), dtype=torch.float32, device=be.device)
    on_gpu = full.device.type == "cuda"
    chunks = UPLOAD_CHUNKS if (on_gpu and per >= 4096 * UPLOAD_CHUNKS) else 1
    step = (per + chunks - 1) // chunks
    main = torch.cuda.current_stream(be.device) if on_gpu else None
    copier = _copy_stream(be.device) if on_gpu and chunks > 1 else None
    if copier is not None:
        copier.wait_stream(main)
    for c in range(chunks):
        c0, c1 = c * step, min(per, (c + 1) * step)
        if c1 <= c0:
            break
        mine = full[rank, c0:c1]                          # this rank's piece lands in place
        rows = max(0, min(r1 - r0, c1) - c0)              # real rows of the piece (the last shards may be ragged)
        ctx = torch.cuda.stream(copier) if copier is not None else _nullcontext()
        with ctx:
            if rows > 0:
                piece = host[r0 + c0:r0 + c0 + rows]
                if on_gpu and hasattr(be, "copy_to_device") and not piece.is_pinned() and piece.is_contiguous():
                    be.copy_to_device(mine[:rows], piece)      # pageable source: staged by host threads (csrc/host_entry.cu)
                else:
                    mine[:rows].copy_(piece, non_blocking=True)
            if rows < c1 - c0:
                mine[rows:].zero_()                       # padding rows of the ragged last shards
        if copier is not None:
            main.wait_stream(copier)                      # (only the copies enqueued so far)
        if chunks == 1:
            dist.all_gather_into_tensor(full.view(world * per, d), mine.reshape(per, d), group=group)
        else:
            dist.all_gather([full[g, c0:c1] for g in range(world)], mine, group=group)
    return full.view(world * per, d)[:n]


_COPY_STREAMS = {}


def _copy_stream(device):
    key = str(device)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def FINCH_sharded(data, group=None, backend=None, **kwargs):
    """FINCH with the level-0 nearest-neighbour stage shared by the ranks of the process group.  Every rank
    must pass the same `data` (host or device); every rank returns the same (c, num_clust, req_c)."""
    be = backend or _backend.default_backend()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return FINCH(data, backend=be, **kwargs)
    dev = upload_replicated(data, group=group, backend=be)
    return FINCH(dev, backend=be, first_neighbors=sharded_first_neighbors(be, group), **kwargs)


def topk_neighbors_sharded(q, x, k, same=False, group=None, backend=None):
    """Retrieval top-k (iic_retrieve_clips.py:295-296, evaluate.py:226-231) with the QUERY rows sharded over
    the process group and the database replicated: rank r searches rows [r * ceil(Q / G), ...) against all of
    x, then the [Q, k] ids and distances are all-gathered (8 * Q * k bytes in total).  same=True: q is x and
    every row excludes itself.  Every rank must pass the same q and x; every rank returns the same result."""
    be = backend or _backend.default_backend()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return be.topk_neighbors(q, x, k, same=same)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nq = q.shape[0]
    r0, r1, per = shard_range(nq, rank, world)
    idx_pad = torch.full((per, k), -1, dtype=torch.int32, device=x.device)
    d_pad = torch.zeros((per, k), dtype=x.dtype, device=x.device)
    if r1 > r0:
        use_screen = x.shape[0] >= _backend.SCREEN_MIN_ROWS and k <= _backend.TOPK_SCREEN_MAX_K
        ux, xb = be.normalize_rows(x, want_f16=use_screen)
        if same:
            uq, qb = ux[r0:r1], (xb[r0:r1] if xb is not None else None)
        else:
            uq, qb = be.normalize_rows(q[r0:r1].contiguous(), want_f16=use_screen)
        idx, d = be.topk_cosine(uq, ux, k, self_offset=r0 if same else -1, q_f16=qb, x_f16=xb)
        idx_pad[: r1 - r0] = idx
        d_pad[: r1 - r0] = d
    idx_all = torch.empty((per * world, k), dtype=torch.int32, device=x.device)
    d_all = torch.empty((per * world, k), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(idx_all, idx_pad, group=group)
    dist.all_gather_into_tensor(d_all, d_pad, group=group)
    return idx_all[:nq].contiguous(), d_all[:nq].contiguous()


def replicate(data, src=0, group=None, backend=None):
    """Broadcast the [N, D] float32 matrix held by rank `src` to every rank's GPU (NCCL over NVLink)."""
    be = backend or _backend.default_backend()
    shape = torch.zeros(2, dtype=torch.int64, device=be.device)
    if dist.get_rank(group) == src:
        t = be.to_device(data, torch.float32)
        shape[0], shape[1] = t.shape
    dist.broadcast(shape, src, group=group)
    if dist.get_rank(group) != src:
        t = torch.empty((int(shape[0]), int(shape[1])), dtype=torch.float32, device=be.device)
    dist.broadcast(t, src, group=group)
    return t
