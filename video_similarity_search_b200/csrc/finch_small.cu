// The small levels of the FINCH hierarchy in ONE cooperative launch (clustering/finch.py:151-167 for n <= 2048).
//
// After the first one or two levels a FINCH hierarchy works on a few hundred centroids (BASELINE config 3:
// 240 000 -> 21 436 -> 404 -> 105 -> 21 -> 5).  Driven from the host every such level is ~25 launches of kernels that
// run for 2 microseconds each plus a read-back of the cluster count - launch latency, not work.  Here the level loop
// itself runs on the device: one persistent grid (one CTA per SM) walks the levels, separated by grid-wide barriers,
// and evaluates the reference's exit rules (finch.py:151, 158-163) itself.  Per level:
//   A  (entering level) normalise the float64 centroids (sklearn normalize: norm = sqrt(sum x^2), 0 -> 1)   finch.py:27
//   B  Gram matrix of the unit rows, float64, 64 x 64 tiles on or right of the diagonal, K split over CTAs when
//      there are few tiles; partial tiles are summed in a fixed order (deterministic across runs and ranks)
//   B2 first neighbour of every row: argmin of clip(1 - s, 0, 2), ties -> lowest index                 finch.py:28-29
//   C  lock-free union-find over the links i - nn[i]: without the min_sim filter every CTA builds the forest in its
//      own shared memory (no barrier); with it the grid shares the sibling-pair distances on a forest   finch.py:40-55
//      in global memory (a link survives iff weight * distance <= min_sim, weight 2 for mutual pairs)
//   D  label = rank of the root (smallest member) among roots - scipy's numbering; exit rules
//   E  compose the labels of all N rows (get_merge, finch.py:74-79); float64 sums / counts / means of the new
//      clusters from the previous level's sums, members added in ascending order, and their unit rows   finch.py:58-71
// Three grid barriers per level (four with the filter).  Bound: latency; the Gram tiles are FP64-pipe work of at
// most 2 * 2048^2 * d flop.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace slic {

constexpr int SL_THREADS = 256;
constexpr int SL_TILE = 64, SL_BK = 16, SL_LD = 66;
constexpr int SL_MAX_KSPLIT = 16;

__device__ __forceinline__ int sl_find_ro(const int* parent, int x) {   // read-only: every CTA walks the same forest
    int p = parent[x];
    while (p != x) {
        x = p;
        p = parent[x];
    }
    return x;
}
__device__ __forceinline__ int sl_find(int* parent, int x) {
    volatile int* p = parent;
    int px = p[x];
    while (px != x) {
        const int ppx = p[px];
        if (ppx != px) p[x] = ppx;
        x = px;
        px = ppx;
    }
    return x;
}
__device__ __forceinline__ void sl_union(int* parent, int a, int b) {
    while (true) {
        a = sl_find(parent, a);
        b = sl_find(parent, b);
        if (a == b) return;
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;
    }
}

// exclusive scan of one int per thread over the CTA (256 threads); *total = sum
__device__ __forceinline__ int sl_block_scan(int v, int* total, int* s_warp /*[9]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();   // s_warp may still be read from the previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < SL_THREADS / 32 ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < SL_THREADS / 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SL_THREADS / 32) s_warp[lane] = wi - w;
        if (lane == SL_THREADS / 32 - 1) s_warp[8] = wi;
    }
    __syncthreads();
    *total = s_warp[8];
    return s_warp[warp] + incl - v;
}

__device__ __forceinline__ int sl_ksplit(int tiles, int kslabs) {
    int ks = 296 / tiles;   // a function of the level's shape only: the summation order never depends on the grid
    if (ks < 1) ks = 1;
    if (ks > SL_MAX_KSPLIT) ks = SL_MAX_KSPLIT;
    if (ks > kslabs) ks = kslabs;
    return ks;
}

// sum of the K-split partial products of one Gram entry, partials added in split order; the loads are independent
__device__ __forceinline__ double sl_gram_at(const double* __restrict__ gram, int ksplit, int64_t plane, int64_t off) {
    double v[SL_MAX_KSPLIT];
#pragma unroll
    for (int ks = 0; ks < SL_MAX_KSPLIT; ++ks) v[ks] = ks < ksplit ? gram[(int64_t)ks * plane + off] : 0.0;
    double s = v[0];
#pragma unroll
    for (int ks = 1; ks < SL_MAX_KSPLIT; ++ks) s += v[ks];
    return s;
}

// unit row of one centroid (sklearn normalize: norm = sqrt(sum x^2), 0 -> 1) by the whole CTA; `row` holds d values
__device__ __forceinline__ void sl_unit_row_cta(const double* __restrict__ row, int d, double* __restrict__ out,
                                                double* s_red /*[9]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc = 0.0;
    for (int k = tid; k < d; k += SL_THREADS) acc = fma(row[k], row[k], acc);
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < SL_THREADS / 32; ++w) t += s_red[w];
        double nrm = sqrt(t);
        s_red[8] = nrm == 0.0 ? 1.0 : nrm;
    }
    __syncthreads();
    const double nrm = s_red[8];
    for (int k = tid; k < d; k += SL_THREADS) out[k] = row[k] / nrm;
}

__global__ void __launch_bounds__(SL_THREADS) finch_small_levels_kernel(const SmallLevelsArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_lab[SMALL_LEVEL_MAX_ROWS];   // root, then label, of every node
    __shared__ int s_aux[SMALL_LEVEL_MAX_ROWS];   // forest / rank of the roots / member lists
    __shared__ __align__(16) double As[SL_BK][SL_LD];
    __shared__ __align__(16) double Bs[SL_BK][SL_LD];
    __shared__ int s_warp[9];
    __shared__ double s_red[9];
    __shared__ int s_cnt;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t gthreads = (int64_t)gridDim.x * SL_THREADS, gtid = (int64_t)blockIdx.x * SL_THREADS + tid;
    const int gwarps = gridDim.x * (SL_THREADS / 32), gwarp = blockIdx.x * (SL_THREADS / 32) + warp;
    const int d = a.d;
    const int kslabs = (d + SL_BK - 1) / SL_BK;
    const double min_sim = (a.use_filter && a.min_sim_dev) ? (double)*a.min_sim_dev : 0.0;

    int levels = a.summary[0];
    int m = a.summary[2 + levels - 1];
    int buf = 0, status = 0, exit_clust = 2;
    int stamp = 0;
#define SL_STAMP()                                                                          \
    do {                                                                                    \
        if (a.trace && blockIdx.x == 0 && tid == 0 && stamp < SMALL_TRACE_STAMPS) {         \
            unsigned long long t_;                                                          \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
            a.trace[stamp] = t_;                                                            \
        }                                                                                   \
        ++stamp;                                                                            \
    } while (0)
    SL_STAMP();
    if (m > SMALL_LEVEL_MAX_ROWS) {
        status = 2;   // the host continues with the large-level path
        exit_clust = 0;
    } else if (m > 1) {
        // ---- A (entering level only; later levels are normalised where their means are formed): unit rows -------
        const double* means = a.means[0];
        for (int row = gwarp; row < m; row += gwarps) {
            const double* xr = means + (int64_t)row * d;
            double acc = 0.0;
            for (int k = lane; k < d; k += 32) acc = fma(xr[k], xr[k], acc);
            acc = warp_sum(acc);
            double nrm = sqrt(acc);
            if (nrm == 0.0) nrm = 1.0;
            for (int k = lane; k < d; k += 32) a.unit[(int64_t)row * d + k] = xr[k] / nrm;
        }
        grid.sync();
    }
    SL_STAMP();
    while (exit_clust > 1) {                                              // finch.py:151
        if (m == 1) break;
        const int T = (m + SL_TILE - 1) / SL_TILE, mp = T * SL_TILE;
        const int64_t plane = (int64_t)mp * mp;
        const int tiles = T * (T + 1) / 2;
        const int ksplit = sl_ksplit(tiles, kslabs);
        const int slabs_per = (kslabs + ksplit - 1) / ksplit;

        // ---- B: partial Gram tiles (the next K slab is fetched into registers while the current one is multiplied) ---
        {
            const int ty = tid >> 4, tx = tid & 15;
            const int lr = tid >> 2, kq = (tid & 3) * 4;   // 64 rows x 16 k per operand: thread t loads row t / 4, 4 k's
            for (int item = blockIdx.x; item < tiles * ksplit; item += gridDim.x) {
                const int tile = item / ksplit, ks = item % ksplit;
                int ti = 0, rem = tile;
                while (rem >= T - ti) {
                    rem -= T - ti;
                    ++ti;
                }
                const int tj = ti + rem;
                const int row0 = ti * SL_TILE, col0 = tj * SL_TILE;
                const int slab0 = ks * slabs_per, slab1 = min(slab0 + slabs_per, kslabs);
                const int ra = row0 + lr, rb = col0 + lr;
                const double* pa = a.unit + (int64_t)ra * d;
                const double* pb = a.unit + (int64_t)rb * d;
                double fa[4], fb[4];
                auto fetch = [&](int sl) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = sl * SL_BK + kq + i;
                        fa[i] = (ra < m && k < d) ? pa[k] : 0.0;
                        fb[i] = (rb < m && k < d) ? pb[k] : 0.0;
                    }
                };
                double acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
                if (slab0 < slab1) fetch(slab0);
                for (int sl = slab0; sl < slab1; ++sl) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        As[kq + i][lr] = fa[i];
                        Bs[kq + i][lr] = fb[i];
                    }
                    __syncthreads();
                    if (sl + 1 < slab1) fetch(sl + 1);
#pragma unroll
                    for (int k = 0; k < SL_BK; ++k) {
                        const double2 a01 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
                        const double2 a23 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
                        const double2 b01 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
                        const double2 b23 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
                        const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                        const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
                    }
                    __syncthreads();
                }
                double* g = a.gram + (int64_t)ks * plane;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = row0 + ty * 4 + i, c = col0 + tx * 4 + j;
                        g[(int64_t)r * mp + c] = acc[i][j];
                        if (ti != tj) g[(int64_t)c * mp + r] = acc[i][j];
                    }
            }
        }
        grid.sync();
        SL_STAMP();

        // ---- B2: first neighbours ------------------------------------------------------------------------------
        for (int row = gwarp; row < m; row += gwarps) {
            double best = CUDART_INF;
            int bj = 0x7fffffff;
            for (int c = lane; c < m; c += 32) {
                if (c == row) continue;
                const double dist = cosine_distance_from_sim<double>(sl_gram_at(a.gram, ksplit, plane, (int64_t)row * mp + c));
                if (closer(dist, c, best, bj)) {
                    best = dist;
                    bj = c;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double so = __shfl_xor_sync(0xffffffffu, best, o);
                const int jo = __shfl_xor_sync(0xffffffffu, bj, o);
                if (closer(so, jo, best, bj)) {
                    best = so;
                    bj = jo;
                }
            }
            if (lane == 0) {
                a.nn[row] = bj;
                a.dist[row] = best;
                a.parent[row] = row;
            }
        }
        grid.sync();
        SL_STAMP();

        // ---- C: links ----------------------------------------------------------------------------------------
        if (a.use_filter) {
            // with the min_sim cut the sibling pairs need distances: the grid shares the work on a forest in global memory
            for (int64_t i = gtid; i < m; i += gthreads) {
                const int j = a.nn[i];
                if (j < 0 || j >= m || j == (int)i) continue;
                const double w = (a.nn[j] == (int)i) ? 2.0 : 1.0;
                if (a.dist[i] * w > min_sim) continue;
                sl_union(a.parent, (int)i, j);
            }
            // rows sharing a first neighbour are linked iff their own distance <= min_sim (weight 1)
            for (int i = gwarp; i < m; i += gwarps) {
                const int hub = a.nn[i];
                for (int j0 = i + 1; j0 < m; j0 += 32) {
                    const int j = j0 + lane;
                    if (j < m && a.nn[j] == hub) {
                        const double s = sl_gram_at(a.gram, ksplit, plane, (int64_t)i * mp + j);
                        if (cosine_distance_from_sim<double>(s) <= min_sim) sl_union(a.parent, i, j);
                    }
                }
            }
            grid.sync();
            for (int i = tid; i < m; i += SL_THREADS) s_lab[i] = sl_find_ro(a.parent, i);
        } else {
            // plain first-neighbour graph: every CTA builds the same forest in its own shared memory (roots are the
            // smallest members whatever the order of the unions) - no barrier, no global traffic
            for (int i = tid; i < m; i += SL_THREADS) s_aux[i] = i;
            __syncthreads();
            for (int i = tid; i < m; i += SL_THREADS) {
                const int j = a.nn[i];
                if (j >= 0 && j < m && j != i) sl_union(s_aux, i, j);
            }
            __syncthreads();
            for (int i = tid; i < m; i += SL_THREADS) s_lab[i] = sl_find_ro(s_aux, i);
        }
        __syncthreads();
        SL_STAMP();

        // ---- D: labels (every CTA, redundantly) and the exit rules ---------------------------------------------
        int cur;
        {
            constexpr int PER = SMALL_LEVEL_MAX_ROWS / SL_THREADS;
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = tid * PER + i;
                cnt += (idx < m && s_lab[idx] == idx) ? 1 : 0;
            }
            int run = sl_block_scan(cnt, &cur, s_warp);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = tid * PER + i;
                if (idx < m && s_lab[idx] == idx) s_aux[idx] = run++;
            }
        }
        __syncthreads();
        {
            constexpr int PER = SMALL_LEVEL_MAX_ROWS / SL_THREADS;
            int lab[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = i * SL_THREADS + tid;
                lab[i] = idx < m ? s_aux[s_lab[idx]] : 0;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = i * SL_THREADS + tid;
                if (idx < m) s_lab[idx] = lab[i];
            }
        }
        __syncthreads();
        exit_clust = m - cur;
        if (cur == 1 || exit_clust < 1) break;                              // finch.py:160-163: level dropped
        if (levels >= a.capacity) {
            status = 1;
            break;
        }

        // ---- E: compose the labels, merge the sums, unit rows of the new centroids -----------------------------
        {
            const int* prev = a.cols + (int64_t)(levels - 1) * a.n_rows;
            int* out = a.cols + (int64_t)levels * a.n_rows;
            for (int64_t i = gtid; i < a.n_rows; i += gthreads) out[i] = s_lab[prev[i]];
        }
        {
            const double* sums_prev = a.sums[buf];
            const int* counts_prev = a.counts[buf];
            double* sums_new = a.sums[buf ^ 1];
            int* counts_new = a.counts[buf ^ 1];
            double* means_new = a.means[buf ^ 1];
            for (int c = blockIdx.x; c < cur; c += gridDim.x) {
                // ordered member list of cluster c
                constexpr int PER = SMALL_LEVEL_MAX_ROWS / SL_THREADS;
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int idx = tid * PER + i;
                    cnt += (idx < m && s_lab[idx] == c) ? 1 : 0;
                }
                int members;
                int run = sl_block_scan(cnt, &members, s_warp);
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int idx = tid * PER + i;
                    if (idx < m && s_lab[idx] == c) s_aux[run++] = idx;
                }
                if (tid == 0) s_cnt = 0;
                __syncthreads();
                int rows = 0;
                for (int e = tid; e < members; e += SL_THREADS) rows += counts_prev[s_aux[e]];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
                if (lane == 0 && rows) atomicAdd(&s_cnt, rows);
                __syncthreads();
                const double total_rows = (double)s_cnt;
                for (int k = tid; k < d; k += SL_THREADS) {
                    double acc = 0.0;
                    for (int e = 0; e < members; ++e) acc += sums_prev[(int64_t)s_aux[e] * d + k];
                    sums_new[(int64_t)c * d + k] = acc;
                    means_new[(int64_t)c * d + k] = acc / total_rows;
                }
                if (tid == 0) counts_new[c] = s_cnt;
                __syncthreads();   // means_new row complete (this CTA wrote all of it); s_aux / s_cnt free again
                sl_unit_row_cta(means_new + (int64_t)c * d, d, a.unit + (int64_t)c * d, s_red);
                __syncthreads();
            }
        }
        if (blockIdx.x == 0 && tid == 0) a.summary[2 + levels] = cur;
        ++levels;
        m = cur;
        buf ^= 1;
        grid.sync();
        SL_STAMP();
    }
    SL_STAMP();
#undef SL_STAMP
    if (blockIdx.x == 0 && tid == 0) {
        a.summary[0] = levels;
        a.summary[1] = status;
    }
}

size_t small_levels_gram_elems(int64_t m) {
    // K-split levels hold at most 256 partial tiles (sl_ksplit); unsplit levels one padded m x m matrix
    const int64_t T = (m + SL_TILE - 1) / SL_TILE, mp = T * SL_TILE;
    const int64_t split_bound = (int64_t)9 * 512 * 512;   // max over T of sl_ksplit(T (T + 1) / 2) * (64 T)^2 is 2.1 M (T = 8, 16)
    return (size_t)(mp * mp > split_bound ? mp * mp : split_bound);
}

int launch_small_levels(const SmallLevelsArgs& args, cudaStream_t st) {
    static int coop = -1;
    int dev = 0;
    SLIC_CUDA_OK(cudaGetDevice(&dev));
    if (coop < 0) SLIC_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) {
        set_error("finch: the device does not support cooperative launches");
        return SLIC_ERR_UNSUPPORTED;
    }
    SmallLevelsArgs a = args;
    void* params[] = {&a};
    SLIC_CUDA_OK(cudaLaunchCooperativeKernel((void*)finch_small_levels_kernel, dim3((unsigned)num_sms()), dim3(SL_THREADS),
                                             params, 0, st));
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

}  // namespace slic
