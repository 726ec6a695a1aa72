// The small levels of the FINCH hierarchy in ONE cooperative launch (clustering/finch.py:151-167 for n <= 2048).
//
// After the first one or two levels a FINCH hierarchy works on a few hundred centroids (BASELINE config 3:
// 240 000 -> 21 436 -> 404 -> 105 -> 21 -> 5).  Driven from the host every such level is ~25 launches of kernels that
// run for 2 microseconds each plus a read-back of the cluster count - launch latency, not work.  Here the level loop
// itself runs on the device: one persistent grid (one CTA per SM) walks the levels, separated by grid-wide barriers,
// and evaluates the reference's exit rules (finch.py:151, 158-163) itself.  Per level:
//   A  normalise the float64 centroids (sklearn normalize: norm = sqrt(sum x^2), 0 -> 1)            finch.py:27
//   B  Gram matrix of the unit rows, float64, 64 x 64 tiles on or right of the diagonal, K split over CTAs when
//      there are few tiles; partial tiles are summed in a fixed order (deterministic across runs and ranks)
//   B2 first neighbour of every row: argmin of clip(1 - s, 0, 2), ties -> lowest index                 finch.py:28-29
//   C  lock-free union-find over the links i - nn[i] (and, with the min_sim filter, the sibling pairs;   finch.py:40-55
//      a link survives iff weight * distance <= min_sim, weight 2 for mutual pairs)
//   D  every CTA flattens the forest into shared memory: label = rank of the root (smallest member) among roots -
//      scipy's numbering; exit rules
//   E  compose the labels of all N rows (get_merge, finch.py:74-79); float64 sums / counts / means of the new
//      clusters from the previous level's sums, members added in ascending order                        finch.py:58-71
// Bound: latency (grid barriers); the Gram tiles are FP64-pipe work of at most 2 * 2048^2 * d flop.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace slic {

constexpr int SL_THREADS = 256;
constexpr int SL_TILE = 64, SL_BK = 16, SL_LD = 66;
constexpr int SL_MAX_KSPLIT = 8;

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
    int ks = 256 / tiles;   // a function of the level's shape only: the summation order never depends on the grid
    if (ks < 1) ks = 1;
    if (ks > SL_MAX_KSPLIT) ks = SL_MAX_KSPLIT;
    if (ks > kslabs) ks = kslabs;
    return ks;
}

__global__ void __launch_bounds__(SL_THREADS) finch_small_levels_kernel(const SmallLevelsArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_lab[SMALL_LEVEL_MAX_ROWS];   // root, then label, of every node
    __shared__ int s_aux[SMALL_LEVEL_MAX_ROWS];   // rank of the roots / member lists
    __shared__ __align__(16) double As[SL_BK][SL_LD];
    __shared__ __align__(16) double Bs[SL_BK][SL_LD];
    __shared__ int s_warp[9];
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
    while (exit_clust > 1) {                                              // finch.py:151
        if (m == 1) break;
        if (m > SMALL_LEVEL_MAX_ROWS) {
            status = 2;   // the host continues with the large-level path
            break;
        }
        const double* means = a.means[buf];
        const int T = (m + SL_TILE - 1) / SL_TILE, mp = T * SL_TILE;
        const int tiles = T * (T + 1) / 2;
        const int ksplit = sl_ksplit(tiles, kslabs);
        const int slabs_per = (kslabs + ksplit - 1) / ksplit;

        // ---- A: unit rows ------------------------------------------------------------------------------------
        for (int row = gwarp; row < m; row += gwarps) {
            const double* xr = means + (int64_t)row * d;
            double acc = 0.0;
            for (int k = lane; k < d; k += 32) acc = fma(xr[k], xr[k], acc);
            acc = warp_sum(acc);
            double nrm = sqrt(acc);
            if (nrm == 0.0) nrm = 1.0;
            for (int k = lane; k < d; k += 32) a.unit[(int64_t)row * d + k] = xr[k] / nrm;
            if (lane == 0) a.parent[row] = row;
        }
        grid.sync();

        // ---- B: partial Gram tiles -----------------------------------------------------------------------------
        {
            const int ty = tid >> 4, tx = tid & 15;
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
                double acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
                for (int sl = slab0; sl < slab1; ++sl) {
                    const int k0 = sl * SL_BK;
                    {   // 64 rows x 16 k per operand: thread t loads row t / 4, k = (t % 4) * 4 .. + 3
                        const int r = tid >> 2, kq = (tid & 3) * 4;
                        const int ra = row0 + r, rb = col0 + r;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int k = k0 + kq + i;
                            As[kq + i][r] = (ra < m && k < d) ? a.unit[(int64_t)ra * d + k] : 0.0;
                            Bs[kq + i][r] = (rb < m && k < d) ? a.unit[(int64_t)rb * d + k] : 0.0;
                        }
                    }
                    __syncthreads();
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
                double* g = a.gram + (int64_t)ks * mp * mp;
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

        // ---- B2: first neighbours ------------------------------------------------------------------------------
        for (int row = gwarp; row < m; row += gwarps) {
            double best = CUDART_INF;
            int bj = 0x7fffffff;
            for (int c = lane; c < m; c += 32) {
                if (c == row) continue;
                double s = 0.0;
                for (int ks = 0; ks < ksplit; ++ks) s += a.gram[((int64_t)ks * mp + row) * mp + c];
                const double dist = cosine_distance_from_sim<double>(s);
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
            }
        }
        grid.sync();

        // ---- C: links ----------------------------------------------------------------------------------------
        for (int64_t i = gtid; i < m; i += gthreads) {
            const int j = a.nn[i];
            if (j < 0 || j >= m || j == (int)i) continue;
            if (a.use_filter) {
                const double w = (a.nn[j] == (int)i) ? 2.0 : 1.0;
                if (a.dist[i] * w > min_sim) continue;
            }
            sl_union(a.parent, (int)i, j);
        }
        if (a.use_filter) {
            // rows sharing a first neighbour are linked iff their own distance <= min_sim (weight 1)
            for (int i = gwarp; i < m; i += gwarps) {
                const int hub = a.nn[i];
                for (int j0 = i + 1; j0 < m; j0 += 32) {
                    const int j = j0 + lane;
                    if (j < m && a.nn[j] == hub) {
                        double s = 0.0;
                        for (int ks = 0; ks < ksplit; ++ks) s += a.gram[((int64_t)ks * mp + i) * mp + j];
                        if (cosine_distance_from_sim<double>(s) <= min_sim) sl_union(a.parent, i, j);
                    }
                }
            }
        }
        grid.sync();

        // ---- D: labels (every CTA, redundantly) and the exit rules ---------------------------------------------
        for (int i = tid; i < m; i += SL_THREADS) s_lab[i] = sl_find_ro(a.parent, i);
        __syncthreads();
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

        // ---- E: compose the labels, merge the sums -------------------------------------------------------------
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
                __syncthreads();   // s_aux / s_cnt are rewritten by the next trip
            }
        }
        if (blockIdx.x == 0 && tid == 0) a.summary[2 + levels] = cur;
        ++levels;
        m = cur;
        buf ^= 1;
        grid.sync();
    }
    if (blockIdx.x == 0 && tid == 0) {
        a.summary[0] = levels;
        a.summary[1] = status;
    }
}

size_t small_levels_gram_elems(int64_t m) {
    // K-split levels hold at most 256 partial tiles (sl_ksplit); unsplit levels one padded m x m matrix
    const int64_t T = (m + SL_TILE - 1) / SL_TILE, mp = T * SL_TILE;
    const int64_t split_bound = (int64_t)SL_MAX_KSPLIT * 64 * SL_TILE * SL_TILE;   // ksplit * mp^2 <= 8 * (256 / 8 tiles -> T <= 8)^2 ...
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
