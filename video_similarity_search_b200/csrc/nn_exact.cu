// K1-exact: brute-force similarity in the reference dtype with float64 accumulation on the CUDA
// cores.  Three users: (i) the parity anchor / cross-check of the tensor-core screen, (ii) the
// finisher for rows whose candidate list overflowed in the screen, (iii) the dense-matrix and
// top-k retrieval entry points (evaluate.py:208-231, iic_retrieve_clips.py:295-296).
//
// Tiling: 64 x 64 outputs per CTA, 16-wide k slabs staged in shared memory as float64, 4 x 4
// register micro-tiles.  This kernel is FP64-pipe bound (2 * nq * n * d flop), it is not the
// throughput path - the tcgen05 screen in nn_screen_tc.cu is.
#include <math_constants.h>

#include "common.cuh"

namespace slic {

constexpr int EX_BM = 64, EX_BN = 64, EX_BK = 16, EX_LD = 66, EX_THREADS = 256;

template <typename T>
__device__ __forceinline__ void load_slab(double (*dst)[EX_LD], const T* __restrict__ base, const int* __restrict__ rows,
                                          int64_t row0, int64_t nrows, int d, int k0) {
    // 64 rows x 16 k: thread t loads row t/4, k = (t%4)*4 .. +3
    const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
    const int64_t row = row0 + r;
    const T* src = nullptr;
    if (row < nrows) src = base + (int64_t)(rows ? rows[row] : row) * d;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + kq + i;
        dst[kq + i][r] = (src && k < d) ? (double)__ldg(src + k) : 0.0;
    }
}

__device__ __forceinline__ void mma_slab(const double (*As)[EX_LD], const double (*Bs)[EX_LD], double acc[4][4], int ty,
                                         int tx) {
#pragma unroll
    for (int k = 0; k < EX_BK; ++k) {
        const double2 a01 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
        const double2 a23 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
        const double2 b01 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
        const double2 b23 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
}

// ---- top-1 ---------------------------------------------------------------------------------
// grid = (row tiles, column splits).  part_* are [splits][nq].
template <typename T>
__global__ void __launch_bounds__(EX_THREADS) exact_top1_kernel(const T* __restrict__ q, const int* __restrict__ q_rows,
                                                                int64_t nq, const T* __restrict__ x, int64_t n, int d,
                                                                int64_t self_offset, int tiles_per_split,
                                                                double* __restrict__ part_score,
                                                                int* __restrict__ part_idx) {
    __shared__ __align__(16) double As[EX_BK][EX_LD];
    __shared__ __align__(16) double Bs[EX_BK][EX_LD];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int64_t row0 = (int64_t)blockIdx.x * EX_BM;
    const int64_t n_col_tiles = ceil_div(n, EX_BN);
    const int64_t ct0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t ct1 = ct0 + tiles_per_split < n_col_tiles ? ct0 + tiles_per_split : n_col_tiles;

    double best_s[4];
    int best_j[4];
    int64_t self_col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        best_s[i] = CUDART_INF;  // holds the best DISTANCE (rounded to T) seen so far
        best_j[i] = 0x7fffffff;
        const int64_t r = row0 + ty * 4 + i;
        self_col[i] = -1;
        if (self_offset >= 0 && r < nq) self_col[i] = (int64_t)(q_rows ? q_rows[r] : r) + self_offset;
    }

    for (int64_t ct = ct0; ct < ct1; ++ct) {
        const int64_t col0 = ct * EX_BN;
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        for (int k0 = 0; k0 < d; k0 += EX_BK) {
            load_slab<T>(As, q, q_rows, row0, nq, d, k0);
            load_slab<T>(Bs, x, nullptr, col0, n, d, k0);
            __syncthreads();
            mma_slab(As, Bs, acc, ty, tx);
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = CUDART_INF;
            int jb = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t col = col0 + tx * 4 + j;
                const double dist = (double)cosine_distance_from_sim<T>(acc[i][j]);
                if (col < n && col != self_col[i] && closer(dist, (int)col, s, jb)) {
                    s = dist;
                    jb = (int)col;
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const double so = __shfl_xor_sync(0xffffffffu, s, o);
                const int jo = __shfl_xor_sync(0xffffffffu, jb, o);
                if (closer(so, jo, s, jb)) {
                    s = so;
                    jb = jo;
                }
            }
            if (closer(s, jb, best_s[i], best_j[i])) {
                best_s[i] = s;
                best_j[i] = jb;
            }
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t r = row0 + ty * 4 + i;
            if (r < nq) {
                part_score[(int64_t)blockIdx.y * nq + r] = best_s[i];
                part_idx[(int64_t)blockIdx.y * nq + r] = best_j[i];
            }
        }
    }
}

template <typename T>
__global__ void exact_top1_merge_kernel(const double* __restrict__ part_score, const int* __restrict__ part_idx,
                                        int64_t nq, int splits, int* __restrict__ idx_out, T* __restrict__ dist_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nq) return;
    double s = CUDART_INF;
    int jb = 0x7fffffff;
    for (int p = 0; p < splits; ++p) {
        const double so = part_score[(int64_t)p * nq + r];
        const int jo = part_idx[(int64_t)p * nq + r];
        if (closer(so, jo, s, jb)) {
            s = so;
            jb = jo;
        }
    }
    idx_out[r] = jb == 0x7fffffff ? -1 : jb;
    if (dist_out) dist_out[r] = (T)s;  // already a distance in T
}

// ---- dense distance matrix ---------------------------------------------------------------
// metric 0: inputs are unit rows, out = clip(1 - s, 0, 2).  metric 1: raw rows + float64 squared norms,
// out = sqrt(max(qq + xx - 2 s, 0)).  same != 0 zeroes the diagonal (sklearn, X is Y);
// inf_col_offset >= 0 writes +inf at column row + inf_col_offset (self exclusion before a top-k).
template <typename T>
__global__ void __launch_bounds__(EX_THREADS) exact_matrix_kernel(const T* __restrict__ q,
                                                                  const int* __restrict__ q_rows, int64_t nq,
                                                                  const T* __restrict__ x, int64_t n, int d, int metric,
                                                                  const double* __restrict__ q_sq,
                                                                  const double* __restrict__ x_sq, int same,
                                                                  int64_t inf_col_offset, T* __restrict__ out,
                                                                  int64_t ld) {
    __shared__ __align__(16) double As[EX_BK][EX_LD];
    __shared__ __align__(16) double Bs[EX_BK][EX_LD];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int64_t row0 = (int64_t)blockIdx.y * EX_BM, col0 = (int64_t)blockIdx.x * EX_BN;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < d; k0 += EX_BK) {
        load_slab<T>(As, q, q_rows, row0, nq, d, k0);
        load_slab<T>(Bs, x, nullptr, col0, n, d, k0);
        __syncthreads();
        mma_slab(As, Bs, acc, ty, tx);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + ty * 4 + i;
        if (r >= nq) continue;
        const int64_t inf_col = inf_col_offset >= 0 ? (int64_t)(q_rows ? q_rows[r] : r) + inf_col_offset : -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = col0 + tx * 4 + j;
            if (c >= n) continue;
            T v;
            if (metric == SLIC_METRIC_COSINE) {
                v = cosine_distance_from_sim<T>(acc[i][j]);
            } else {
                double d2 = q_sq[r] + x_sq[c] - 2.0 * acc[i][j];
                v = (T)sqrt(d2 > 0.0 ? d2 : 0.0);
            }
            if (same && r == c) v = (T)0;
            if (c == inf_col) v = (T)CUDART_INF;
            out[r * ld + c] = v;
        }
    }
}

template <typename T>
__global__ void sq_norms_kernel(const T* __restrict__ x, int64_t n, int d, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    double a = 0;
    for (int k = lane; k < d; k += 32) {
        double v = (double)x[row * d + k];
        a = fma(v, v, a);
    }
    a = warp_sum(a);
    if (lane == 0) out[row] = a;
}

// ---- row-wise top-k of a dense matrix (radix select + bitonic sort of the k winners) --------
template <typename T> struct SortKey;
template <> struct SortKey<float> {
    typedef uint32_t type;
    static constexpr int BYTES = 4;
    __device__ static type make(float v) {
        uint32_t b = __float_as_uint(v);
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
    static constexpr type MAXKEY = 0xffffffffu;
};
template <> struct SortKey<double> {
    typedef uint64_t type;
    static constexpr int BYTES = 8;
    __device__ static type make(double v) {
        uint64_t b = (uint64_t)__double_as_longlong(v);
        return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    }
    static constexpr type MAXKEY = 0xffffffffffffffffull;
};

constexpr int SEL_THREADS = 256;
constexpr int SEL_MAXK = 2048;

template <typename T>
__global__ void __launch_bounds__(SEL_THREADS) rows_topk_kernel(const T* __restrict__ mat, int64_t nq, int64_t n,
                                                                int64_t ld, int k, int kpow2, int* __restrict__ idx_out,
                                                                T* __restrict__ val_out) {
    typedef typename SortKey<T>::type K;
    __shared__ int hist[256];
    __shared__ K s_prefix;
    __shared__ int s_remaining, s_equal_total, s_cnt_less, s_cnt_eq;
    __shared__ K s_keys[SEL_MAXK];
    __shared__ int s_idx[SEL_MAXK];
    const T* row = mat + (int64_t)blockIdx.x * ld;
    const int tid = threadIdx.x;

    if (tid == 0) {
        s_prefix = 0;
        s_remaining = k;
    }
    K known = 0;
    for (int pass = SortKey<T>::BYTES - 1; pass >= 0; --pass) {
        hist[tid] = 0;
        __syncthreads();
        const K prefix = s_prefix;
        for (int64_t j = tid; j < n; j += SEL_THREADS) {
            const K key = SortKey<T>::make(row[j]);
            if (((key ^ prefix) & known) == 0) atomicAdd(&hist[(int)((key >> (8 * pass)) & 255)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, b = 0, rem = s_remaining;
            for (; b < 256; ++b) {
                if (cum + hist[b] >= rem) break;
                cum += hist[b];
            }
            s_remaining = rem - cum;
            s_prefix = prefix | ((K)b << (8 * pass));
            s_equal_total = hist[b];
        }
        known |= (K)255 << (8 * pass);
        __syncthreads();
    }
    const K kth = s_prefix;             // key of the k-th smallest entry
    const int need_eq = s_remaining;    // how many entries equal to kth belong to the top-k
    const int n_less = k - need_eq;
    const bool eq_all = (s_equal_total == need_eq);
    if (tid == 0) {
        s_cnt_less = 0;
        s_cnt_eq = 0;
    }
    for (int j = tid; j < kpow2; j += SEL_THREADS) {
        s_keys[j] = SortKey<T>::MAXKEY;
        s_idx[j] = 0x7fffffff;
    }
    __syncthreads();
    for (int64_t j = tid; j < n; j += SEL_THREADS) {
        const K key = SortKey<T>::make(row[j]);
        if (key < kth) {
            const int p = atomicAdd(&s_cnt_less, 1);
            s_keys[p] = key;
            s_idx[p] = (int)j;
        } else if (eq_all && key == kth) {
            const int p = n_less + atomicAdd(&s_cnt_eq, 1);
            s_keys[p] = key;
            s_idx[p] = (int)j;
        }
    }
    if (!eq_all && tid < 32) {
        // ties at the boundary: take the lowest column indices, in order
        int got = 0;
        for (int64_t base = 0; base < n && got < need_eq; base += 32) {
            const int64_t j = base + tid;
            const bool hit = j < n && SortKey<T>::make(row[j]) == kth;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            const int pos = got + __popc(m & ((1u << tid) - 1u));
            if (hit && pos < need_eq) {
                s_keys[n_less + pos] = kth;
                s_idx[n_less + pos] = (int)j;
            }
            got += __popc(m);
        }
    }
    __syncthreads();
    // bitonic sort of kpow2 (key, idx) pairs, ascending
    for (int size = 2; size <= kpow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (kpow2 >> 1); t += SEL_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const K ka = s_keys[lo], kb = s_keys[hi];
                const int ia = s_idx[lo], ib = s_idx[hi];
                const bool a_after_b = (ka > kb) || (ka == kb && ia > ib);
                if (a_after_b == up) {
                    s_keys[lo] = kb; s_keys[hi] = ka;
                    s_idx[lo] = ib; s_idx[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < k; j += SEL_THREADS) {
        const int c = s_idx[j];
        idx_out[(int64_t)blockIdx.x * k + j] = c;
        if (val_out) val_out[(int64_t)blockIdx.x * k + j] = row[c];
    }
}

static int pick_splits(int64_t row_tiles, int64_t col_tiles) {
    const int64_t target = (int64_t)num_sms() * 2;
    int64_t s = row_tiles >= target ? 1 : ceil_div(target, row_tiles);
    if (s > col_tiles) s = col_tiles;
    if (s < 1) s = 1;
    return (int)s;
}

template <typename T>
static int exact_top1_impl(const T* q, const int* q_rows, int64_t nq, const T* x, int64_t n, int d, int64_t self_offset,
                           int* idx_out, T* dist_out, cudaStream_t st) {
    const int64_t row_tiles = ceil_div(nq, EX_BM), col_tiles = ceil_div(n, EX_BN);
    const int splits = pick_splits(row_tiles, col_tiles);
    const int tiles_per_split = (int)ceil_div(col_tiles, splits);
    const int real_splits = (int)ceil_div(col_tiles, tiles_per_split);
    Scratch ps, pi;
    SLIC_CUDA_OK(ps.alloc((int64_t)real_splits * nq * sizeof(double), st));
    SLIC_CUDA_OK(pi.alloc((int64_t)real_splits * nq * sizeof(int), st));
    dim3 grid((unsigned)row_tiles, (unsigned)real_splits);
    exact_top1_kernel<T><<<grid, EX_THREADS, 0, st>>>(q, q_rows, nq, x, n, d, self_offset, tiles_per_split,
                                                      ps.as<double>(), pi.as<int>());
    SLIC_LAUNCH_OK();
    exact_top1_merge_kernel<T><<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(ps.as<double>(), pi.as<int>(), nq,
                                                                            real_splits, idx_out, dist_out);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

template <typename T>
static int matrix_impl(const T* q, const int* q_rows, int64_t nq, const T* x, int64_t n, int d, int metric, int same,
                       int64_t inf_col_offset, T* out, int64_t ld, cudaStream_t st) {
    Scratch qs, xs;
    const double* qsp = nullptr;
    const double* xsp = nullptr;
    if (metric == SLIC_METRIC_EUCLIDEAN) {
        SLIC_CUDA_OK(qs.alloc(nq * sizeof(double), st));
        SLIC_CUDA_OK(xs.alloc(n * sizeof(double), st));
        sq_norms_kernel<T><<<(unsigned)ceil_div(nq, 8), 256, 0, st>>>(q, nq, d, qs.as<double>());
        SLIC_LAUNCH_OK();
        sq_norms_kernel<T><<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(x, n, d, xs.as<double>());
        SLIC_LAUNCH_OK();
        qsp = qs.as<double>();
        xsp = xs.as<double>();
    }
    // rows on grid.y (<= 65535 tiles = 4.19M rows)
    dim3 grid((unsigned)ceil_div(n, EX_BN), (unsigned)ceil_div(nq, EX_BM));
    SLIC_REQUIRE(grid.y <= 65535, "distance_matrix: too many query rows for one launch");
    exact_matrix_kernel<T><<<grid, EX_THREADS, 0, st>>>(q, q_rows, nq, x, n, d, metric, qsp, xsp, same, inf_col_offset, out,
                                                        ld);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

static int next_pow2(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

template <typename T>
static int rows_topk_impl(const T* mat, int64_t nq, int64_t n, int64_t ld, int k, int* idx_out, T* val_out,
                          cudaStream_t st) {
    rows_topk_kernel<T><<<(unsigned)nq, SEL_THREADS, 0, st>>>(mat, nq, n, ld, k, next_pow2(k), idx_out, val_out);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

// q_rows (optional): the query rows to process (indices into q); outputs are indexed by position in the list and
// the excluded column of list entry i is q_rows[i] + self_offset.
template <typename T>
static int topk_cosine_impl(const T* q, const int* q_rows, int64_t nq, const T* x, int64_t n, int d, int k,
                            int64_t self_offset, int* idx_out, T* dist_out, cudaStream_t st) {
    // one row block of the distance matrix at a time (<= 1 GiB), selected immediately
    int64_t rows = ((int64_t)1 << 30) / (n * (int64_t)sizeof(T));
    rows = rows < 64 ? 64 : (rows / 64) * 64;
    if (rows > nq) rows = nq;
    Scratch block;
    SLIC_CUDA_OK(block.alloc(rows * n * sizeof(T), st));
    for (int64_t r0 = 0; r0 < nq; r0 += rows) {
        const int64_t nr = nq - r0 < rows ? nq - r0 : rows;
        if (q_rows) {
            SLIC_PROPAGATE(matrix_impl<T>(q, q_rows + r0, nr, x, n, d, SLIC_METRIC_COSINE, 0, self_offset >= 0 ? self_offset : -1,
                                          block.as<T>(), n, st));
        } else {
            const int64_t inf_off = self_offset >= 0 ? self_offset + r0 : -1;
            SLIC_PROPAGATE(matrix_impl<T>(q + r0 * d, nullptr, nr, x, n, d, SLIC_METRIC_COSINE, 0, inf_off, block.as<T>(), n, st));
        }
        SLIC_PROPAGATE(rows_topk_impl<T>(block.as<T>(), nr, n, n, k, idx_out + r0 * k, dist_out ? dist_out + r0 * k : nullptr, st));
    }
    return SLIC_OK;
}

int exact_topk_cosine_rows(const void* q, const int* q_rows, int64_t nq, const void* x, int64_t n, int d, int dtype, int k,
                           int64_t self_offset, int* idx_out, void* dist_out, cudaStream_t st) {
    if (nq == 0) return SLIC_OK;
    if (dtype == SLIC_F32)
        return topk_cosine_impl<float>((const float*)q, q_rows, nq, (const float*)x, n, d, k, self_offset, idx_out,
                                       (float*)dist_out, st);
    return topk_cosine_impl<double>((const double*)q, q_rows, nq, (const double*)x, n, d, k, self_offset, idx_out,
                                    (double*)dist_out, st);
}

}  // namespace slic

extern "C" {

int slic_nn_exact_top1(const void* q_unit_dev, const int32_t* q_rows_dev, int64_t nq, const void* x_unit_dev, int64_t n,
                       int32_t d, int32_t dtype, int64_t self_offset, int32_t* idx_out_dev, void* dist_out_dev,
                       slic_stream_t stream) {
    SLIC_REQUIRE(nq >= 0 && n > 0 && n < ((int64_t)1 << 31) && d > 0, "nn_exact_top1: bad shape");
    SLIC_REQUIRE(q_unit_dev && x_unit_dev && idx_out_dev, "nn_exact_top1: null pointer");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "nn_exact_top1: bad dtype");
    if (nq == 0) return SLIC_OK;
    cudaStream_t st = slic::as_stream(stream);
    if (dtype == SLIC_F32)
        return slic::exact_top1_impl<float>((const float*)q_unit_dev, q_rows_dev, nq, (const float*)x_unit_dev, n, d,
                                            self_offset, idx_out_dev, (float*)dist_out_dev, st);
    return slic::exact_top1_impl<double>((const double*)q_unit_dev, q_rows_dev, nq, (const double*)x_unit_dev, n, d,
                                         self_offset, idx_out_dev, (double*)dist_out_dev, st);
}

int slic_distance_matrix(const void* q_dev, int64_t nq, const void* x_dev, int64_t n, int32_t d, int32_t dtype,
                         int32_t metric, int32_t same_matrix, void* out_dev, int64_t ld_out, slic_stream_t stream) {
    SLIC_REQUIRE(nq >= 0 && n >= 0 && d > 0 && ld_out >= n, "distance_matrix: bad shape");
    SLIC_REQUIRE(q_dev && x_dev && out_dev, "distance_matrix: null pointer");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "distance_matrix: bad dtype");
    SLIC_REQUIRE(metric == SLIC_METRIC_COSINE || metric == SLIC_METRIC_EUCLIDEAN, "distance_matrix: bad metric");
    if (nq == 0 || n == 0) return SLIC_OK;
    cudaStream_t st = slic::as_stream(stream);
    if (dtype == SLIC_F32)
        return slic::matrix_impl<float>((const float*)q_dev, nullptr, nq, (const float*)x_dev, n, d, metric, same_matrix,
                                        -1, (float*)out_dev, ld_out, st);
    return slic::matrix_impl<double>((const double*)q_dev, nullptr, nq, (const double*)x_dev, n, d, metric, same_matrix,
                                     -1, (double*)out_dev, ld_out, st);
}

int slic_rows_topk(const void* dist_dev, int64_t nq, int64_t n, int64_t ld, int32_t dtype, int32_t k,
                   int32_t* idx_out_dev, void* val_out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(nq >= 0 && n > 0 && n < ((int64_t)1 << 31) && ld >= n, "rows_topk: bad shape");
    SLIC_REQUIRE(k > 0 && k <= n && k <= slic::SEL_MAXK, "rows_topk: k must satisfy 1 <= k <= min(n, 2048)");
    SLIC_REQUIRE(dist_dev && idx_out_dev, "rows_topk: null pointer");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "rows_topk: bad dtype");
    if (nq == 0) return SLIC_OK;
    cudaStream_t st = slic::as_stream(stream);
    if (dtype == SLIC_F32)
        return slic::rows_topk_impl<float>((const float*)dist_dev, nq, n, ld, k, idx_out_dev, (float*)val_out_dev, st);
    return slic::rows_topk_impl<double>((const double*)dist_dev, nq, n, ld, k, idx_out_dev, (double*)val_out_dev, st);
}

int slic_topk_cosine(const void* q_unit_dev, int64_t nq, const void* x_unit_dev, int64_t n, int32_t d, int32_t dtype,
                     int32_t k, int64_t self_offset, int32_t* idx_out_dev, void* dist_out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(nq >= 0 && n > 0 && n < ((int64_t)1 << 31) && d > 0, "topk_cosine: bad shape");
    SLIC_REQUIRE(k > 0 && k <= n && k <= slic::SEL_MAXK, "topk_cosine: k must satisfy 1 <= k <= min(n, 2048)");
    SLIC_REQUIRE(q_unit_dev && x_unit_dev && idx_out_dev, "topk_cosine: null pointer");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "topk_cosine: bad dtype");
    if (nq == 0) return SLIC_OK;
    cudaStream_t st = slic::as_stream(stream);
    return slic::exact_topk_cosine_rows(q_unit_dev, nullptr, nq, x_unit_dev, n, d, dtype, k, self_offset, idx_out_dev,
                                        dist_out_dev, st);
}

}  // extern "C"
