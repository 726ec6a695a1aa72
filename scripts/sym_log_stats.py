"""Analyse the candidate log of one symmetric screen (SLIC_SYM_DEBUG=2 dump): who logs what (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SLIC_SYM_DEBUG"] = "2"
import numpy as np, torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
be = CudaBackend()
n, d, k, seed = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (20011, 96, 0, 2))]
x = synth.gaussian_mixture(n, d, k, seed) if k else np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)
unit, ub = be.normalize_rows(be.to_device(x))
be.nn_top1(unit, ub, unit, ub, self_offset=0)
raw = open("/tmp/slic_sym_log.bin", "rb").read()
hdr = np.frombuffer(raw[:24], dtype=np.int64); n, regions, region = [int(v) for v in hdr]
off = 24
cnt = np.frombuffer(raw[off:off + 4 * regions], dtype=np.int32); off += 4 * regions
q = np.frombuffer(raw[off:off + 4 * regions * region], dtype=np.int32).reshape(regions, region); off += 4 * regions * region
nb = np.frombuffer(raw[off:off + 4 * regions * region], dtype=np.int32).reshape(regions, region); off += 4 * regions * region
s = np.frombuffer(raw[off:off + 4 * regions * region], dtype=np.float32).reshape(regions, region)
print("n", n, "regions", regions, "region", region, "records", cnt.sum(), "per row %.1f" % (cnt.sum() / n))
order = np.argsort(-cnt)[:8]
for r in order:
    c = cnt[r]; cta, w = divmod(r, 8)
    qq, nn, ss = q[r, :c], nb[r, :c], s[r, :c]
    # row role: query is one of this warp's rows => q // 128 pattern; approximate: role by q < nb? (triangle: row role has nb >= q's tile)
    print("region %d (cta %d warp %d): %d records; distinct queries %d; q range [%d, %d]; nb range [%d, %d]; score min %.3f med %.3f max %.3f"
          % (r, cta, w, c, len(np.unique(qq)), qq.min(), qq.max(), nn.min(), nn.max(), ss.min(), np.median(ss), ss.max()))
    u, uc = np.unique(qq, return_counts=True)
    top = np.argsort(-uc)[:5]
    print("    busiest queries:", [(int(u[i]), int(uc[i])) for i in top])
allq = np.concatenate([q[r, :cnt[r]] for r in range(regions)])
u, uc = np.unique(allq, return_counts=True)
print("records per query row: mean %.1f median %.0f p99 %.0f max %d (row %d)" % (uc.mean(), np.median(uc), np.percentile(uc, 99), uc.max(), u[np.argmax(uc)]))
