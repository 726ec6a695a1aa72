// Native FINCH driver: the whole hierarchy of clustering/finch.py:108-178 behind ONE C-ABI call.
//
//   slic_finch       device-resident [N, D] float32 matrix in, [N, P] int32 label matrix out (device)
//   slic_finch_host  host buffers in and out; the host -> device copy of the embeddings is cut into row chunks
//                    that are normalised as they land, and the level-0 tensor-core screen is launched BEFORE the
//                    first chunk has arrived: its persistent CTAs consume the triangle of chunk pairs as the gates
//                    open (nn_screen_tc.cu, GateSpec), so all but the first chunk of the PCIe transfer is hidden
//                    behind the O(N^2 D) stage.
//
// The level loop follows the reference line by line (exit rules finch.py:151-163, min_sim mode :142-144, the
// "no dense distances above 70 000 rows" control flow :30-38); every numeric step is one of the kernels behind
// include/slic_b200.h.  Host work per level: one 4-byte read-back of the cluster count.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <vector>

#include "common.cuh"

namespace slic {

// finch.py:19 - above it the reference has no dense distance matrix.  A module-level constant there, library-level
// state here (slic_set_flann_threshold), so that tests can exercise the "no dense distances" control flow at small n.
static int64_t FLANN_THRESHOLD = 70000;
constexpr int64_t SCREEN_MIN_ROWS = 2048;    // below: the tensor-core screen is launch overhead, use the exact kernel
constexpr int FINCH_MAX_LEVELS = 64;
constexpr int64_t GATED_MIN_ROWS = 32768;    // host entry: below this the upload is too short to be worth pipelining
constexpr int GATED_CHUNKS = 8;             // measured at C3 (492 MB): 4 / 8 / 12 chunks -> 31.6 / 30.6 / 30.4 ms end to end

// experiments: SLIC_GATED_CHUNKS=<c> overrides the chunk count of the pipelined upload (<= 1: no pipelining)
static int gated_chunks() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("SLIC_GATED_CHUNKS");
        cached = e ? atoi(e) : GATED_CHUNKS;
        if (cached < 1) cached = 1;
        if (cached > 64) cached = 64;
    }
    return cached;
}

// Pipelined upload policy.  The gated launch relies on the upload stream's kernels becoming resident NEXT TO the
// persistent screen kernel, which CUDA does not promise: serialised launches (CUDA_LAUNCH_BLOCKING, compute-sanitizer,
// Nsight's kernel replay), MPS / time slicing or an oversubscribed device keep the gates shut.  So:
//   * slic_set_upload_overlap(0) / SLIC_UPLOAD_OVERLAP=0 switch it off explicitly;
//   * it is off by itself when the environment shows a serialising tool;
//   * a gate that stays shut for ~3 s makes the kernel give up WITHOUT trapping (nn_screen_tc.cu, gate_wait): the driver
//     repeats the search after the upload and keeps the overlap off for the rest of the process.
static int g_overlap_user = -1;          // -1: not set by the caller, 0 / 1: slic_set_upload_overlap
static bool g_overlap_failed = false;    // a gated launch timed out in this process
static bool env_set(const char* name) {
    const char* e = getenv(name);
    return e && *e && strcmp(e, "0") != 0;
}
static bool upload_overlap_allowed() {
    if (g_overlap_failed) return false;
    if (g_overlap_user >= 0) return g_overlap_user != 0;
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("SLIC_UPLOAD_OVERLAP");
        if (e && *e) cached = atoi(e) != 0 ? 1 : 0;
        else
            cached = !(env_set("CUDA_LAUNCH_BLOCKING") || getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") ||
                       getenv("NV_SANITIZER_INJECTION_PORT_BASE") || getenv("NV_SANITIZER_INJECTION_PORT_RANGE_BEGIN") ||
                       getenv("SANITIZER_INJECTION_PATH") || getenv("NSYS_PROFILING_SESSION_ID") || getenv("CUDA_MPS_PIPE_DIRECTORY"))
                         ? 1 : 0;
    }
    return cached != 0;
}

// diagnostic timeline of the last slic_finch_host call (slic_host_trace): CUDA events, read after the call
static bool g_host_trace = false;
static cudaEvent_t g_up0 = nullptr, g_up1 = nullptr, g_t0 = nullptr, g_t_search = nullptr, g_t_end = nullptr;

static int d_pad_of(int d) { return (d + 63) / 64 * 64; }

// Pinned mailbox for the read-backs of a call (a pageable D2H costs an extra staging hop): ints [0, 16) receive the
// per-level block {search counters [8], cluster count}, ints [16, 16 + 2 + 64] the final summary, then min_sim.
constexpr int MB_LEVEL = 0, MB_SUMMARY = 16, MB_MIN_SIM = MB_SUMMARY + 2 + FINCH_MAX_LEVELS, MB_INTS = MB_MIN_SIM + 2;
struct Mailbox {
    int* in = nullptr;    // device -> host
    int* out = nullptr;   // host -> device (the summary seed)
    cudaEvent_t ev = nullptr;
};
static int mailbox(Mailbox** out) {
    static thread_local Mailbox mb;
    if (!mb.in) {
        SLIC_CUDA_OK(cudaHostAlloc((void**)&mb.in, MB_INTS * sizeof(int), cudaHostAllocDefault));
        SLIC_CUDA_OK(cudaHostAlloc((void**)&mb.out, MB_INTS * sizeof(int), cudaHostAllocDefault));
        SLIC_CUDA_OK(cudaEventCreateWithFlags(&mb.ev, cudaEventDisableTiming));
    }
    *out = &mb;
    return SLIC_OK;
}

// out[i, l] = cols[l][i] for l < *levels: the [N, P] C-contiguous matrix np.column_stack builds (finch.py:157).
// P is read from device memory (the level loop may have ended on the device, finch_small.cu).
__global__ void stack_columns_kernel(const int* __restrict__ cols, const int* __restrict__ levels_dev, int64_t n,
                                     int* __restrict__ out) {
    const int levels = *levels_dev;
    const int64_t total = n * levels, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t i = t / levels;
        const int l = (int)(t % levels);
        out[t] = cols[(int64_t)l * n + i];
    }
}

// opens upload gate `g` (own kernel, one warp: a cudaMemsetAsync may be a driver kernel of unknown shape, and whatever
// runs on the copy stream must fit next to the persistent screen kernel)
__global__ void open_gate_kernel(int* gate) {
    if (threadIdx.x == 0) {
        __threadfence();
        atomicExch(gate, 1);
    }
}

// initial_rank (int64, finch.py:22-23) -> int32; entries outside [0, n) are counted (the reference would raise)
__global__ void convert_rank_kernel(const int64_t* __restrict__ in, int64_t n, int* __restrict__ out, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = in[i];
    out[i] = (int)v;
    if (v < 0 || v >= n) atomicAdd(bad, 1);
}

typedef std::unique_ptr<Scratch> ScratchPtr;
static ScratchPtr new_scratch() { return ScratchPtr(new Scratch()); }

struct Level0 {
    const int* nn;        // [n]
    const float* dist;    // [n] or nullptr
    const float* unit;    // [n, d] or nullptr
    bool dense;           // the reference would hold a dense distance matrix (len(orig_dist) != 0, finch.py:142)
    // The level-0 search was launched asynchronously: its counters land in async_stats (device, 16 ints: [1] rows left
    // unsettled, [4] pipeline error, [5] log overflow; [8] is the driver's cluster count slot).  When any is set the
    // driver repeats the search synchronously through slic_nn_top1 on (retry_unit, retry_ub).
    int* async_stats = nullptr;
    const uint16_t* retry_ub = nullptr;
    int* retry_nn = nullptr;
    float* retry_dist = nullptr;
    // every row has a first neighbour other than itself (our own searches): each component then has >= 2 rows
    bool no_self_links = false;
};

static int d_pad_of_(int d) { return (d + 63) / 64 * 64; }

// Everything after the level-0 search.  labels_out: device [n, capacity] ints, filled as [n, P] row-major.
//
// Host synchronisations: one per level driven from here (levels of more than 2048 clusters: the cluster count sizes
// the next level's launches) + one at the end; every one of them is placed AFTER the next batch of device work has
// been enqueued, so the device never waits for the host.  Levels of <= 2048 clusters run inside ONE cooperative
// launch with the loop on the device (finch_small.cu).
static int finch_levels(const float* data, int64_t n, int d, const Level0& l0, bool ensure_early_exit, int capacity,
                        int* labels_out, int* num_clust_host, int* num_levels_host, float* min_sim_host,
                        int* has_min_sim_host, cudaStream_t st) {
    Mailbox* mb;
    SLIC_PROPAGATE(mailbox(&mb));
    const int dp = d_pad_of_(d);
    std::vector<int> num_clust;
    Scratch cols, blk0, ms, csr_off, csr_cnt, summary;
    SLIC_CUDA_OK(cols.alloc((size_t)n * capacity * sizeof(int), st));
    SLIC_CUDA_OK(csr_off.alloc((size_t)(n + 1) * sizeof(int), st));
    SLIC_CUDA_OK(csr_cnt.alloc((size_t)n * sizeof(int), st));
    SLIC_CUDA_OK(ms.alloc(2 * sizeof(float), st));
    SLIC_CUDA_OK(summary.alloc((2 + FINCH_MAX_LEVELS) * sizeof(int), st));
    int* blk = l0.async_stats;
    if (!blk) {
        SLIC_CUDA_OK(blk0.alloc(16 * sizeof(int), st));
        SLIC_CUDA_OK(cudaMemsetAsync(blk0.ptr, 0, 16 * sizeof(int), st));
        blk = blk0.as<int>();
    }
    const bool have_min_sim = ensure_early_exit && l0.dense && l0.dist && l0.unit && n > 1;    // finch.py:142-144

    // ---- level 0: components of the first-neighbour graph (finch.py:136), centroids (:137), min_sim (:142-144) ----
    ScratchPtr sums = new_scratch(), means = new_scratch(), counts_own = new_scratch();
    int* counts = csr_cnt.as<int>();   // level 0: the row counts come out of the components pass
    int cur = 0;
    int64_t bound = (l0.no_self_links && n > 1) ? n / 2 : n;
    for (int attempt = 0;; ++attempt) {
        SLIC_PROPAGATE(finch_components_csr(l0.nn, n, 0, nullptr, nullptr, 0, SLIC_F32, nullptr, cols.as<int>(), blk + 8,
                                            csr_off.as<int>(), csr_cnt.as<int>(), st, blk + 9));
        SLIC_CUDA_OK(cudaMemcpyAsync(mb->in + MB_LEVEL, blk, 10 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SLIC_CUDA_OK(cudaEventRecord(mb->ev, st));
        // the means are enqueued before the host knows the cluster count (it lives on the device; `bound` sizes buffers)
        sums = new_scratch();
        means = new_scratch();
        SLIC_CUDA_OK(sums->alloc((size_t)bound * d * sizeof(double), st));
        SLIC_CUDA_OK(means->alloc((size_t)bound * d * sizeof(double), st));
        {
            Scratch order;
            SLIC_CUDA_OK(order.alloc((size_t)n * sizeof(int), st));
            SLIC_PROPAGATE(order_rows_by_label(cols.as<int>(), n, bound, order.as<int>(), st));
            SLIC_PROPAGATE(cluster_sums_csr<float>(data, nullptr, order.as<int>(), csr_off.as<int>(), n, d, (int)bound,
                                                   blk + 8, sums->as<double>(), nullptr, means->as<double>(), st));
        }
        if (have_min_sim)
            SLIC_PROPAGATE(slic_finch_min_sim(l0.nn, n, l0.unit, d, SLIC_F32, l0.dist, ms.as<float>(), st));
        SLIC_CUDA_OK(cudaEventSynchronize(mb->ev));
        const int* s = mb->in + MB_LEVEL;
        cur = s[8];
        if (s[9] != 0 && !(s[1] | s[4] | s[5])) {   // finch.py:41-43 would raise on such an index (scipy: out of bounds)
            set_error("finch: %d first-neighbour indices lie outside [0, n)", s[9]);
            return SLIC_ERR_INVALID_ARG;
        }
        if (l0.retry_nn && attempt == 0 && (s[1] | s[4] | s[5])) {
            if (s[4] == 3) {
                g_overlap_failed = true;   // the upload did not run next to the gated kernel: search again, now that it has landed
            } else if (s[4] != 0 && s[4] != 2) {
                set_error("nn_screen_kernel: pipeline barrier timed out");
                return SLIC_ERR_CUDA;
            }
            // rows left to the exact finisher, or a full candidate log (degenerate input): the synchronous entry has the
            // fallbacks; everything enqueued above is repeated on its result
            SLIC_PROPAGATE(slic_nn_top1(l0.unit, l0.retry_ub, n, l0.unit, l0.retry_ub, n, d, dp, SLIC_F32, 0, 0.f,
                                        l0.retry_nn, l0.retry_dist, nullptr, st));
            SLIC_CUDA_OK(cudaMemsetAsync(blk, 0, 16 * sizeof(int), st));
            continue;
        } else if (s[1] | s[4] | s[5]) {
            set_error("finch: the level-0 neighbour search did not complete (counters %d %d %d)", s[1], s[4], s[5]);
            return SLIC_ERR_CUDA;
        }
        if (cur > bound) {   // (a caller-supplied neighbour array with self links) - repeat with room for n clusters
            bound = n;
            continue;
        }
        break;
    }
    num_clust.push_back(cur);

    // ---- levels driven from the host: more clusters than the device-side loop takes ------------------------------
    int exit_clust = 2;
    bool loop_open = true;   // finch.py:151 still running
    while (loop_open && num_clust.back() > SMALL_LEVEL_MAX_ROWS) {
        const int64_t m = num_clust.back();
        const int levels = (int)num_clust.size();
        const bool filter = have_min_sim && m <= FLANN_THRESHOLD;          // finch.py:51-52 (needs dense distances)
        Scratch unit, ub, nn, dist, u, blkl, off2;
        SLIC_CUDA_OK(unit.alloc((size_t)m * d * sizeof(double), st));
        SLIC_CUDA_OK(ub.alloc((size_t)m * dp * 2, st));
        SLIC_CUDA_OK(nn.alloc((size_t)m * sizeof(int), st));
        SLIC_CUDA_OK(dist.alloc((size_t)m * sizeof(double), st));
        SLIC_CUDA_OK(u.alloc((size_t)m * sizeof(int), st));
        SLIC_CUDA_OK(blkl.alloc(16 * sizeof(int), st));
        SLIC_CUDA_OK(off2.alloc((size_t)(m + 1) * sizeof(int), st));
        SLIC_PROPAGATE(slic_normalize_rows(means->ptr, m, d, SLIC_F64, unit.ptr, nullptr, ub.as<uint16_t>(), dp, st));
        ScratchPtr s2, c2, m2;
        int64_t bound_l = filter ? m : m / 2;   // without the cut every component has >= 2 members
        for (int attempt = 0;; ++attempt) {
            if (attempt == 0) {
                SLIC_CUDA_OK(cudaMemsetAsync(blkl.ptr, 0, 16 * sizeof(int), st));
                SLIC_PROPAGATE(nn_top1_self_async(unit.ptr, ub.as<uint16_t>(), m, d, dp, SLIC_F64, nn.as<int>(), dist.ptr,
                                                  blkl.as<int>(), st));
            }
            SLIC_PROPAGATE(finch_components_csr(nn.as<int>(), m, filter ? 1 : 0, ms.as<float>(), unit.ptr, d, SLIC_F64,
                                                dist.ptr, u.as<int>(), blkl.as<int>() + 8, off2.as<int>(), nullptr, st));
            SLIC_CUDA_OK(cudaMemcpyAsync(mb->in + MB_LEVEL, blkl.ptr, 9 * sizeof(int), cudaMemcpyDeviceToHost, st));
            SLIC_CUDA_OK(cudaEventRecord(mb->ev, st));
            // optimistic: compose the labels and merge the sums while the host waits for the count
            if (levels < capacity)
                SLIC_PROPAGATE(slic_compose_labels(cols.as<int>() + (size_t)(levels - 1) * n, u.as<int>(), n,
                                                   cols.as<int>() + (size_t)levels * n, st));   // get_merge, finch.py:74-79
            s2 = new_scratch();
            c2 = new_scratch();
            m2 = new_scratch();
            SLIC_CUDA_OK(s2->alloc((size_t)bound_l * d * sizeof(double), st));
            SLIC_CUDA_OK(c2->alloc((size_t)bound_l * sizeof(int), st));
            SLIC_CUDA_OK(m2->alloc((size_t)bound_l * d * sizeof(double), st));
            {
                Scratch order;
                SLIC_CUDA_OK(order.alloc((size_t)m * sizeof(int), st));
                SLIC_PROPAGATE(order_rows_by_label(u.as<int>(), m, bound_l, order.as<int>(), st));
                SLIC_PROPAGATE(cluster_sums_csr<double>(sums->as<double>(), counts, order.as<int>(), off2.as<int>(), m, d,
                                                        (int)bound_l, blkl.as<int>() + 8, s2->as<double>(), c2->as<int>(),
                                                        m2->as<double>(), st));
            }
            SLIC_CUDA_OK(cudaEventSynchronize(mb->ev));
            const int* s = mb->in + MB_LEVEL;
            cur = s[8];
            if (attempt == 0 && (s[1] | s[4] | s[5])) {
                if (s[4] != 0 && s[4] != 2) {
                    set_error("nn_screen_kernel: pipeline barrier timed out");
                    return SLIC_ERR_CUDA;
                }
                SLIC_PROPAGATE(slic_nn_top1(unit.ptr, ub.as<uint16_t>(), m, unit.ptr, ub.as<uint16_t>(), m, d, dp, SLIC_F64, 0,
                                            0.f, nn.as<int32_t>(), dist.ptr, nullptr, st));
                continue;
            }
            if (cur > bound_l) {
                bound_l = m;
                continue;
            }
            break;
        }
        exit_clust = (int)m - cur;
        if (cur == 1 || exit_clust < 1) {                                  // finch.py:160-163: level dropped
            loop_open = false;
            break;
        }
        if (levels >= capacity || levels >= FINCH_MAX_LEVELS) {
            set_error("finch: more than %d partitions; enlarge the label buffer", levels);
            return SLIC_ERR_OVERFLOW;
        }
        sums.swap(s2);
        means.swap(m2);
        counts_own.swap(c2);
        counts = counts_own->as<int>();
        num_clust.push_back(cur);
        if (exit_clust <= 1) loop_open = false;                            // finch.py:151
    }

    // ---- the remaining levels: one cooperative launch, loop and exit rules on the device ------------------------
    int* seed = mb->out;
    seed[0] = (int)num_clust.size();
    seed[1] = 0;
    for (size_t l = 0; l < num_clust.size(); ++l) seed[2 + l] = num_clust[l];
    SLIC_CUDA_OK(cudaMemcpyAsync(summary.ptr, seed, (2 + num_clust.size()) * sizeof(int), cudaMemcpyHostToDevice, st));
    Scratch sl_sums, sl_counts, sl_means, sl_unit, sl_gram, sl_nn, sl_dist, sl_parent, sl_trace;
    const int64_t m_in = num_clust.back();
    if (loop_open && m_in > 1) {
        SLIC_CUDA_OK(sl_sums.alloc((size_t)m_in * d * sizeof(double), st));
        SLIC_CUDA_OK(sl_counts.alloc((size_t)m_in * sizeof(int), st));
        SLIC_CUDA_OK(sl_means.alloc((size_t)m_in * d * sizeof(double), st));
        SLIC_CUDA_OK(sl_unit.alloc((size_t)m_in * d * sizeof(double), st));
        SLIC_CUDA_OK(sl_gram.alloc(small_levels_gram_elems(m_in) * sizeof(double), st));
        SLIC_CUDA_OK(sl_nn.alloc((size_t)m_in * sizeof(int), st));
        SLIC_CUDA_OK(sl_dist.alloc((size_t)m_in * sizeof(double), st));
        SLIC_CUDA_OK(sl_parent.alloc((size_t)m_in * sizeof(int), st));
        SmallLevelsArgs a;
        a.n_rows = n;
        a.d = d;
        a.capacity = capacity < FINCH_MAX_LEVELS ? capacity : FINCH_MAX_LEVELS;
        a.cols = cols.as<int>();
        a.summary = summary.as<int>();
        a.sums[0] = sums->as<double>();
        a.sums[1] = sl_sums.as<double>();
        a.counts[0] = counts;
        a.counts[1] = sl_counts.as<int>();
        a.means[0] = means->as<double>();
        a.means[1] = sl_means.as<double>();
        a.unit = sl_unit.as<double>();
        a.gram = sl_gram.as<double>();
        a.nn = sl_nn.as<int>();
        a.dist = sl_dist.as<double>();
        a.parent = sl_parent.as<int>();
        a.use_filter = (have_min_sim && m_in <= FLANN_THRESHOLD) ? 1 : 0;
        a.min_sim_dev = ms.as<float>();
        a.trace = nullptr;
        static int small_trace = -1;   // SLIC_SMALL_TRACE=1: print the phase timeline of the device-side level loop
        if (small_trace < 0) {
            const char* e = getenv("SLIC_SMALL_TRACE");
            small_trace = e && atoi(e) == 1 ? 1 : 0;
        }
        if (small_trace) {
            SLIC_CUDA_OK(sl_trace.alloc(SMALL_TRACE_STAMPS * sizeof(unsigned long long), st));
            SLIC_CUDA_OK(cudaMemsetAsync(sl_trace.ptr, 0, SMALL_TRACE_STAMPS * sizeof(unsigned long long), st));
            a.trace = sl_trace.as<unsigned long long>();
        }
        SLIC_PROPAGATE(launch_small_levels(a, st));
        if (small_trace) {
            unsigned long long h[SMALL_TRACE_STAMPS];
            SLIC_CUDA_OK(cudaMemcpyAsync(h, sl_trace.ptr, sizeof(h), cudaMemcpyDeviceToHost, st));
            SLIC_CUDA_OK(cudaStreamSynchronize(st));
            fprintf(stderr, "[slic] small levels from m=%lld: phase stamps (us since start):", (long long)m_in);
            for (int i = 1; i < SMALL_TRACE_STAMPS && h[i]; ++i) fprintf(stderr, " %.1f", (double)(h[i] - h[0]) * 1e-3);
            fprintf(stderr, "\n");
        }
    }
    {
        const int64_t want = ceil_div(n * 8, 256);
        const int64_t cap_blocks = (int64_t)num_sms() * 16;
        stack_columns_kernel<<<(unsigned)(want < cap_blocks ? want : cap_blocks), 256, 0, st>>>(cols.as<int>(), summary.as<int>(),
                                                                                          n, labels_out);
        SLIC_LAUNCH_OK();
    }
    SLIC_CUDA_OK(cudaMemcpyAsync(mb->in + MB_SUMMARY, summary.ptr, (2 + FINCH_MAX_LEVELS) * sizeof(int), cudaMemcpyDeviceToHost,
                                 st));
    if (have_min_sim)
        SLIC_CUDA_OK(cudaMemcpyAsync(mb->in + MB_MIN_SIM, ms.ptr, sizeof(float), cudaMemcpyDeviceToHost, st));
    SLIC_CUDA_OK(cudaStreamSynchronize(st));
    const int* fin = mb->in + MB_SUMMARY;
    if (fin[1] == 1) {
        set_error("finch: more than %d partitions; enlarge the label buffer", capacity);
        return SLIC_ERR_OVERFLOW;
    }
    if (fin[1] != 0) {
        set_error("finch: device-side level loop reported status %d", fin[1]);
        return SLIC_ERR_CUDA;
    }
    const int levels = fin[0];
    for (int l = 0; l < levels; ++l) num_clust_host[l] = fin[2 + l];
    *num_levels_host = levels;
    float min_sim = 0.f;
    if (have_min_sim) memcpy(&min_sim, mb->in + MB_MIN_SIM, sizeof(float));
    if (min_sim_host) *min_sim_host = min_sim;
    if (has_min_sim_host) *has_min_sim_host = have_min_sim ? 1 : 0;
    return SLIC_OK;
}

// ---- pipelined upload ---------------------------------------------------------------------------
struct Upload {
    const float* x_host;
    int64_t n;
    int d, d_pad;
    float* data;
    float* unit;
    uint16_t* ub;
    int* gates;
    int num_chunks;
    int64_t chunk_rows;
    cudaStream_t copy_stream, main_stream;
    cudaEvent_t done;
    cudaStream_t prep_stream;       // normalise + gate of chunk c run here while chunk c + 1 is being copied
    cudaEvent_t* chunk_landed;      // [num_chunks] (pipelined upload only)
};

// Runs on the host right after the gated screen kernel has been launched on main_stream: enqueue, chunk by chunk,
// copy (copy stream) -> normalise -> open the gate (preparation stream, behind an event per chunk), then make the main
// stream wait for the last chunk (the exact re-rank reads float32 unit rows of every chunk).  The copies run back to
// back: with copy and normalise on ONE stream the copy engine idled while each chunk was normalised by guest CTAs
// squeezed in next to the persistent screen kernel - the upload of C3 lasted 12 ms against 9.3 ms for the bare copy,
// and the screen, whose available work grows with the square of the rows that have arrived, starved that much longer.
static int run_upload(void* ctx) {
    Upload* up = static_cast<Upload*>(ctx);
    if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_up0, up->copy_stream));
    StagePool& pool = stage_pool();
    const bool staged = up->num_chunks > 1 && (size_t)up->chunk_rows * up->d * sizeof(float) <= ((size_t)256 << 20) &&
                        host_is_pageable(up->x_host);   // (3 pinned staging buffers of one chunk each: bounded)
    if (staged) SLIC_PROPAGATE(pool.ensure((size_t)up->chunk_rows * up->d * sizeof(float)));
    for (int c = 0; c < up->num_chunks; ++c) {
        const int64_t r0 = (int64_t)c * up->chunk_rows;
        const int64_t rows = (r0 + up->chunk_rows <= up->n ? up->chunk_rows : up->n - r0);
        const float* src = up->x_host + r0 * up->d;
        if (staged) {
            const int slot = c % STAGE_SLOTS;
            if (pool.busy[slot]) SLIC_CUDA_OK(cudaEventSynchronize(pool.done[slot]));   // its previous DMA has drained
            parallel_host_copy(pool.buf[slot], src, (size_t)rows * up->d * sizeof(float));
            src = static_cast<const float*>(pool.buf[slot]);
        }
        SLIC_CUDA_OK(cudaMemcpyAsync(up->data + r0 * up->d, src, (size_t)rows * up->d * sizeof(float),
                                     cudaMemcpyHostToDevice, up->copy_stream));
        if (staged) {
            const int slot = c % STAGE_SLOTS;
            SLIC_CUDA_OK(cudaEventRecord(pool.done[slot], up->copy_stream));
            pool.busy[slot] = true;
        }
        cudaStream_t prep = up->copy_stream;
        if (up->chunk_landed) {
            prep = up->prep_stream;
            SLIC_CUDA_OK(cudaEventRecord(up->chunk_landed[c], up->copy_stream));
            SLIC_CUDA_OK(cudaStreamWaitEvent(prep, up->chunk_landed[c], 0));
        }
        SLIC_PROPAGATE(slic_normalize_rows(up->data + r0 * up->d, rows, up->d, SLIC_F32, up->unit + r0 * up->d, nullptr,
                                           up->ub + r0 * up->d_pad, up->d_pad, prep));
        if (up->gates) {
            open_gate_kernel<<<1, 32, 0, prep>>>(up->gates + c);
            SLIC_LAUNCH_OK();
        }
    }
    cudaStream_t last = up->chunk_landed ? up->prep_stream : up->copy_stream;
    if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_up1, last));
    SLIC_CUDA_OK(cudaEventRecord(up->done, last));
    SLIC_CUDA_OK(cudaStreamWaitEvent(up->main_stream, up->done, 0));
    return SLIC_OK;
}

struct HostStreams {
    cudaStream_t main = nullptr, copy = nullptr, prep = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    std::vector<cudaEvent_t> landed;   // one per upload chunk
    int init() {
        SLIC_CUDA_OK(cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking));
        SLIC_CUDA_OK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        SLIC_CUDA_OK(cudaStreamCreateWithFlags(&prep, cudaStreamNonBlocking));
        SLIC_CUDA_OK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        SLIC_CUDA_OK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        return SLIC_OK;
    }
    int chunk_events(int count) {
        while ((int)landed.size() < count) {
            cudaEvent_t e = nullptr;
            SLIC_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            landed.push_back(e);
        }
        return SLIC_OK;
    }
    ~HostStreams() {
        if (main) cudaStreamSynchronize(main);
        if (copy) cudaStreamSynchronize(copy);
        if (prep) cudaStreamSynchronize(prep);
        for (cudaEvent_t e : landed) cudaEventDestroy(e);
        if (ready) cudaEventDestroy(ready);
        if (done) cudaEventDestroy(done);
        if (prep) cudaStreamDestroy(prep);
        if (copy) cudaStreamDestroy(copy);
        if (main) cudaStreamDestroy(main);
    }
};

static int finch_host_impl(const float* x_host, int64_t n, int d, const int64_t* initial_rank_host, bool ensure_early_exit,
                           int capacity, int* labels_out_host, int* num_clust_host, int* num_levels_host,
                           float* min_sim_host, int* has_min_sim_host, HostStreams& hs) {
    const int dp = d_pad_of(d);
    cudaStream_t st = hs.main;
    Scratch data, unit, ub, nn, dist, gates, labels, rank64, blk;
    // Declared AFTER the buffers, hence destroyed BEFORE them: on every return path (errors included) both streams are
    // drained before a buffer goes back to the pool - the copy stream may still be writing chunks that `st` never
    // waited for when an error cut the call short.
    struct Drain {
        HostStreams& h;
        ~Drain() {
            cudaStreamSynchronize(h.copy);
            cudaStreamSynchronize(h.prep);
            cudaStreamSynchronize(h.main);
        }
    } drain{hs};
    SLIC_CUDA_OK(data.alloc((size_t)n * d * sizeof(float), st));
    SLIC_CUDA_OK(nn.alloc((size_t)n * sizeof(int), st));
    // page-locked destination: the stacking kernel writes the [N, P] matrix into it directly (one synchronisation, no
    // copy after the level count is known); pageable destination: device buffer + copy
    int* labels_direct = static_cast<int*>(pinned_device_view(labels_out_host));
    if (!labels_direct) SLIC_CUDA_OK(labels.alloc((size_t)n * capacity * sizeof(int), st));
    SLIC_CUDA_OK(blk.alloc(16 * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(blk.ptr, 0, 16 * sizeof(int), st));
    Level0 l0;
    l0.nn = nn.as<int>();
    l0.dist = nullptr;
    l0.unit = nullptr;
    l0.dense = false;
    l0.async_stats = blk.as<int>();

    if (initial_rank_host) {
        // finch.py:22-23: the caller's neighbours are used as they are; no distances => no min_sim (:142)
        SLIC_CUDA_OK(rank64.alloc((size_t)n * sizeof(int64_t), st));
        SLIC_CUDA_OK(cudaMemcpyAsync(rank64.ptr, initial_rank_host, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        convert_rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(rank64.as<int64_t>(), n, nn.as<int>(),
                                                                      blk.as<int>() + 9);
        SLIC_LAUNCH_OK();
        SLIC_CUDA_OK(cudaMemcpyAsync(data.ptr, x_host, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, st));
    } else if (n == 1) {
        SLIC_CUDA_OK(cudaMemsetAsync(nn.ptr, 0, sizeof(int), st));
        SLIC_CUDA_OK(cudaMemcpyAsync(data.ptr, x_host, (size_t)d * sizeof(float), cudaMemcpyHostToDevice, st));
        l0.dense = true;   // distances would exist, but min_sim needs n > 1
    } else {
        SLIC_CUDA_OK(unit.alloc((size_t)n * d * sizeof(float), st));
        SLIC_CUDA_OK(dist.alloc((size_t)n * sizeof(float), st));
        const bool screen = n >= SCREEN_MIN_ROWS;
        if (screen) SLIC_CUDA_OK(ub.alloc((size_t)n * dp * 2, st));
        Upload up = {x_host, n, d, dp, data.as<float>(), unit.as<float>(), ub.as<uint16_t>(), nullptr, 1, n, hs.copy, st,
                     hs.done, hs.prep, nullptr};
        if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_t0, st));
        int guest_threads = 0, guest_regs = 0;
        SLIC_PROPAGATE(normalize_kernel_shape(&guest_threads, &guest_regs));
        if (screen && n >= GATED_MIN_ROWS && gated_chunks() > 1 && upload_overlap_allowed() &&
            screen_can_overlap_upload(n, dp, guest_threads, guest_regs)) {
            // pipelined: screen first, upload behind it
            up.num_chunks = gated_chunks();
            up.chunk_rows = ceil_div(ceil_div(n, up.num_chunks), 256) * 256;
            up.num_chunks = (int)ceil_div(n, up.chunk_rows);
            // A pageable source is staged through pinned buffers: size them NOW.  cudaFreeHost / cudaHostAlloc synchronise
            // the device - issued while the gated kernel is already waiting for its first chunk they would wait for a
            // kernel that waits for them (measured: the 2 s gate time-out, then the whole process falls back to
            // upload-then-search, 70 ms per call instead of 33).
            if ((size_t)up.chunk_rows * d * sizeof(float) <= ((size_t)256 << 20) && host_is_pageable(x_host))
                SLIC_PROPAGATE(stage_pool().ensure((size_t)up.chunk_rows * d * sizeof(float)));
            SLIC_CUDA_OK(gates.alloc((up.num_chunks + 1) * sizeof(int), st));
            SLIC_CUDA_OK(cudaMemsetAsync(gates.ptr, 0, (up.num_chunks + 1) * sizeof(int), st));
            up.gates = gates.as<int>();
            SLIC_PROPAGATE(hs.chunk_events(up.num_chunks));
            up.chunk_landed = hs.landed.data();
            // Every kernel the copy stream will run must be LOADED before the screen kernel starts to wait for it: with
            // lazy module loading the first launch of a function may synchronise the context - behind the very kernel
            // that is waiting.  normalize_kernel_shape() above loaded the normalise kernel; open a spare gate here.
            open_gate_kernel<<<1, 32, 0, st>>>(up.gates + up.num_chunks);
            SLIC_LAUNCH_OK();
            // the copy stream may touch the buffers (allocated in stream order on `st`) and the zeroed gates only
            // after this point
            SLIC_CUDA_OK(cudaEventRecord(hs.ready, st));
            SLIC_CUDA_OK(cudaStreamWaitEvent(hs.copy, hs.ready, 0));
            GateSpec gs = {gates.as<int>(), up.num_chunks, up.chunk_rows};
            SLIC_PROPAGATE(nn_top1_f32_gated(unit.as<float>(), ub.as<uint16_t>(), n, unit.as<float>(), ub.as<uint16_t>(), n,
                                             d, dp, 0, 0.f, nn.as<int>(), dist.as<float>(), nullptr, &gs, run_upload, &up,
                                             st, blk.as<int>()));
        } else {
            SLIC_CUDA_OK(cudaEventRecord(hs.ready, st));
            SLIC_CUDA_OK(cudaStreamWaitEvent(hs.copy, hs.ready, 0));
            if (!screen) up.ub = nullptr;
            // one chunk, no gates: copy + normalise on the copy stream, then search
            if (screen) {
                SLIC_PROPAGATE(run_upload(&up));
                SLIC_PROPAGATE(nn_top1_f32_gated(unit.as<float>(), ub.as<uint16_t>(), n, unit.as<float>(), ub.as<uint16_t>(), n,
                                                 d, dp, 0, 0.f, nn.as<int>(), dist.as<float>(), nullptr, nullptr, nullptr,
                                                 nullptr, st, blk.as<int>()));
            } else {
                SLIC_CUDA_OK(cudaMemcpyAsync(data.ptr, x_host, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, st));
                SLIC_PROPAGATE(slic_normalize_rows(data.ptr, n, d, SLIC_F32, unit.ptr, nullptr, nullptr, dp, st));
                SLIC_PROPAGATE(slic_nn_exact_top1(unit.ptr, nullptr, n, unit.ptr, n, d, SLIC_F32, 0, nn.as<int32_t>(),
                                                  dist.ptr, st));
            }
        }
        l0.dist = dist.as<float>();
        l0.unit = unit.as<float>();
        l0.dense = n <= FLANN_THRESHOLD;
        l0.retry_ub = ub.as<uint16_t>();
        l0.retry_nn = nn.as<int>();
        l0.retry_dist = dist.as<float>();
        l0.no_self_links = true;
        if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_t_search, st));
    }
    int levels = 0;
    SLIC_PROPAGATE(finch_levels(data.as<float>(), n, d, l0, ensure_early_exit, capacity,
                                labels_direct ? labels_direct : labels.as<int>(), num_clust_host, &levels, min_sim_host,
                                has_min_sim_host, st));
    *num_levels_host = levels;
    if (!labels_direct)
        SLIC_CUDA_OK(cudaMemcpyAsync(labels_out_host, labels.ptr, (size_t)n * levels * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_t_end, st));
    SLIC_CUDA_OK(cudaStreamSynchronize(st));
    return SLIC_OK;
}

// comm.cu (single-process multi-GPU FINCH): everything after a level-0 search that was carried out elsewhere, on the
// device whose stream this is; labels go straight to the host.  blk16: the search's asynchronous status block
// ([1] rows without a neighbour, [4] pipeline error, [5] incomplete) - when set, the search is repeated here on this
// one device through the synchronous path before the hierarchy is built (nn / dist are then overwritten).
int finch_tail_device(const float* data, int64_t n, int d, int* nn, float* dist, const float* unit, const uint16_t* ub,
                      int* blk16, bool ensure_early_exit, int capacity, int* labels_out_dev, int* num_clust_host,
                      int* num_levels_host, float* min_sim_host, int* has_min_sim_host, cudaStream_t st) {
    Level0 l0;
    l0.nn = nn;
    l0.dist = dist;
    l0.unit = unit;
    l0.dense = n <= FLANN_THRESHOLD;
    l0.async_stats = blk16;
    l0.retry_ub = ub;
    l0.retry_nn = nn;
    l0.retry_dist = dist;
    l0.no_self_links = true;
    return finch_levels(data, n, d, l0, ensure_early_exit, capacity, labels_out_dev, num_clust_host, num_levels_host,
                        min_sim_host, has_min_sim_host, st);
}

int finch_tail_to_host(const float* data, int64_t n, int d, int* nn, float* dist, const float* unit, const uint16_t* ub,
                       int* blk16, bool ensure_early_exit, int capacity, int* labels_out_host, int* num_clust_host,
                       int* num_levels_host, float* min_sim_host, int* has_min_sim_host, cudaStream_t st) {
    Scratch labels;
    int* labels_direct = static_cast<int*>(pinned_device_view(labels_out_host));   // (see finch_host_impl)
    if (!labels_direct) SLIC_CUDA_OK(labels.alloc((size_t)n * capacity * sizeof(int), st));
    Level0 l0;
    l0.nn = nn;
    l0.dist = dist;
    l0.unit = unit;
    l0.dense = n <= FLANN_THRESHOLD;
    l0.async_stats = blk16;
    l0.retry_ub = ub;
    l0.retry_nn = nn;
    l0.retry_dist = dist;
    l0.no_self_links = true;
    int levels = 0;
    SLIC_PROPAGATE(finch_levels(data, n, d, l0, ensure_early_exit, capacity, labels_direct ? labels_direct : labels.as<int>(),
                                num_clust_host, &levels, min_sim_host, has_min_sim_host, st));
    *num_levels_host = levels;
    if (!labels_direct) {
        SLIC_CUDA_OK(cudaMemcpyAsync(labels_out_host, labels.ptr, (size_t)n * levels * sizeof(int), cudaMemcpyDeviceToHost, st));
        SLIC_CUDA_OK(cudaStreamSynchronize(st));
    }
    return SLIC_OK;
}

// the host-matrix entry on the CURRENT device with its own streams (comm.cu falls back to it for inputs the multi-GPU
// search does not take: caller-supplied neighbours, fewer rows than the symmetric screen needs)
int finch_host_single(const float* x_host, int64_t n, int d, const int64_t* initial_rank_host, bool ensure_early_exit,
                      int capacity, int* labels_out_host, int* num_clust_host, int* num_levels_host, float* min_sim_host,
                      int* has_min_sim_host) {
    HostStreams hs;
    SLIC_PROPAGATE(hs.init());
    return finch_host_impl(x_host, n, d, initial_rank_host, ensure_early_exit, capacity, labels_out_host, num_clust_host,
                           num_levels_host, min_sim_host, has_min_sim_host, hs);
}

}  // namespace slic

extern "C" {

int slic_host_trace(int32_t enable, float* ms_out_host) {
    using namespace slic;
    if (enable && !g_up0) {
        SLIC_CUDA_OK(cudaEventCreate(&g_up0));
        SLIC_CUDA_OK(cudaEventCreate(&g_up1));
        SLIC_CUDA_OK(cudaEventCreate(&g_t0));
        SLIC_CUDA_OK(cudaEventCreate(&g_t_search));
        SLIC_CUDA_OK(cudaEventCreate(&g_t_end));
    }
    if (ms_out_host && g_host_trace) {
        SLIC_CUDA_OK(cudaEventElapsedTime(ms_out_host + 0, g_t0, g_up0));        // start -> first copy begins
        SLIC_CUDA_OK(cudaEventElapsedTime(ms_out_host + 1, g_up0, g_up1));       // upload (copies + normalise + gates)
        SLIC_CUDA_OK(cudaEventElapsedTime(ms_out_host + 2, g_t0, g_t_search));   // start -> level-0 search complete
        SLIC_CUDA_OK(cudaEventElapsedTime(ms_out_host + 3, g_t0, g_t_end));      // start -> labels copied back
    }
    g_host_trace = enable != 0;
    return SLIC_OK;
}

int slic_set_upload_overlap(int32_t enable) {
    slic::g_overlap_user = enable < 0 ? -1 : (enable != 0 ? 1 : 0);
    if (enable > 0) slic::g_overlap_failed = false;
    return SLIC_OK;
}

int slic_set_flann_threshold(int64_t rows) {
    SLIC_REQUIRE(rows >= 0, "set_flann_threshold: negative");
    slic::FLANN_THRESHOLD = rows;
    return SLIC_OK;
}

int slic_finch(const float* data_dev, int64_t n, int32_t d, const int32_t* nn0_dev, const float* dist0_dev,
               const float* unit0_dev, int32_t level0_dense, int32_t ensure_early_exit, int32_t capacity,
               int32_t* labels_out_dev, int32_t* num_clust_out_host, int32_t* num_levels_out_host,
               float* min_sim_out_host, int32_t* has_min_sim_out_host, slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n >= 1 && n < ((int64_t)1 << 31) && d > 0, "finch: bad shape");
    SLIC_REQUIRE(data_dev && labels_out_dev && num_clust_out_host && num_levels_out_host, "finch: null pointer");
    SLIC_REQUIRE(capacity >= 1, "finch: the label buffer needs at least one column");
    SLIC_PROPAGATE(slic_require_device());
    cudaStream_t st = as_stream(stream);
    const int dp = d_pad_of(d);
    Scratch unit, ub, nn, dist, blk;
    SLIC_CUDA_OK(blk.alloc(16 * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(blk.ptr, 0, 16 * sizeof(int), st));
    Level0 l0;
    l0.async_stats = blk.as<int>();
    if (nn0_dev) {
        l0.nn = nn0_dev;
        l0.dist = dist0_dev;
        l0.unit = unit0_dev;
        l0.dense = level0_dense != 0;
    } else if (n == 1) {
        SLIC_CUDA_OK(nn.alloc(sizeof(int), st));
        SLIC_CUDA_OK(cudaMemsetAsync(nn.ptr, 0, sizeof(int), st));
        l0.nn = nn.as<int>();
        l0.dist = nullptr;
        l0.unit = nullptr;
        l0.dense = true;
    } else {
        const bool screen = n >= SCREEN_MIN_ROWS;
        SLIC_CUDA_OK(unit.alloc((size_t)n * d * sizeof(float), st));
        SLIC_CUDA_OK(nn.alloc((size_t)n * sizeof(int), st));
        SLIC_CUDA_OK(dist.alloc((size_t)n * sizeof(float), st));
        if (screen) SLIC_CUDA_OK(ub.alloc((size_t)n * dp * 2, st));
        SLIC_PROPAGATE(slic_normalize_rows(data_dev, n, d, SLIC_F32, unit.ptr, nullptr, screen ? ub.as<uint16_t>() : nullptr,
                                           dp, st));
        if (screen)
            SLIC_PROPAGATE(nn_top1_f32_gated(unit.as<float>(), ub.as<uint16_t>(), n, unit.as<float>(), ub.as<uint16_t>(), n, d,
                                             dp, 0, 0.f, nn.as<int>(), dist.as<float>(), nullptr, nullptr, nullptr, nullptr,
                                             st, blk.as<int>()));
        else
            SLIC_PROPAGATE(slic_nn_exact_top1(unit.ptr, nullptr, n, unit.ptr, n, d, SLIC_F32, 0, nn.as<int32_t>(), dist.ptr,
                                              st));
        l0.nn = nn.as<int>();
        l0.dist = dist.as<float>();
        l0.unit = unit.as<float>();
        l0.dense = n <= FLANN_THRESHOLD;
        l0.retry_ub = ub.as<uint16_t>();
        l0.retry_nn = nn.as<int>();
        l0.retry_dist = dist.as<float>();
        l0.no_self_links = true;
    }
    return finch_levels(data_dev, n, d, l0, ensure_early_exit != 0, capacity, labels_out_dev, num_clust_out_host,
                        num_levels_out_host, min_sim_out_host, has_min_sim_out_host, st);
}

int slic_finch_host(const float* x_host, int64_t n, int32_t d, const int64_t* initial_rank_host,
                    int32_t ensure_early_exit, int32_t capacity, int32_t* labels_out_host,
                    int32_t* num_clust_out_host, int32_t* num_levels_out_host, float* min_sim_out_host,
                    int32_t* has_min_sim_out_host) {
    using namespace slic;
    SLIC_REQUIRE(n >= 1 && n < ((int64_t)1 << 31) && d > 0, "finch_host: bad shape");
    SLIC_REQUIRE(x_host && labels_out_host && num_clust_out_host && num_levels_out_host, "finch_host: null pointer");
    SLIC_REQUIRE(capacity >= 1 && capacity <= FINCH_MAX_LEVELS, "finch_host: capacity must be in [1, 64]");
    SLIC_PROPAGATE(slic_require_device());
    HostStreams hs;
    SLIC_PROPAGATE(hs.init());
    return finch_host_impl(x_host, n, d, initial_rank_host, ensure_early_exit != 0, capacity, labels_out_host,
                           num_clust_out_host, num_levels_out_host, min_sim_out_host, has_min_sim_out_host, hs);
}

}  // extern "C"
