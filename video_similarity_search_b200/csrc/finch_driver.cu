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
#include <cuda_bf16.h>
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

// diagnostic timeline of the last slic_finch_host call (slic_host_trace): CUDA events, read after the call
static bool g_host_trace = false;
static cudaEvent_t g_up0 = nullptr, g_up1 = nullptr, g_t0 = nullptr, g_t_search = nullptr, g_t_end = nullptr;

static int d_pad_of(int d) { return (d + 63) / 64 * 64; }

// 64-byte pinned mailbox for the per-level scalar read-backs (a pageable D2H costs an extra staging hop)
static int mailbox(void** out) {
    static thread_local void* box = nullptr;
    if (!box) SLIC_CUDA_OK(cudaHostAlloc(&box, 64, cudaHostAllocDefault));
    *out = box;
    return SLIC_OK;
}

static int read_back(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st) {
    void* box;
    SLIC_PROPAGATE(mailbox(&box));
    SLIC_CUDA_OK(cudaMemcpyAsync(box, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    SLIC_CUDA_OK(cudaStreamSynchronize(st));
    memcpy(dst_host, box, bytes);
    return SLIC_OK;
}

struct ColumnPtrs {
    const int* col[FINCH_MAX_LEVELS];
};

// out[i, l] = col[l][i]: the [N, P] C-contiguous matrix np.column_stack builds (finch.py:157)
__global__ void stack_columns_kernel(ColumnPtrs cols, int levels, int64_t n, int* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * levels) return;
    const int64_t i = t / levels;
    const int l = (int)(t % levels);
    out[t] = cols.col[l][i];
}

// opens upload gate `g` (own kernel, one warp: a cudaMemsetAsync may be a driver kernel of unknown shape, and whatever
// runs on the copy stream must fit next to the persistent screen kernel)
__global__ void open_gate_kernel(int* gate) {
    if (threadIdx.x == 0) {
        __threadfence();
        atomicExch(gate, 1);
    }
}

__global__ void convert_rank_kernel(const int64_t* __restrict__ in, int64_t n, int* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

typedef std::unique_ptr<Scratch> ScratchPtr;
static ScratchPtr new_scratch() { return ScratchPtr(new Scratch()); }

struct Level0 {
    const int* nn;        // [n]
    const float* dist;    // [n] or nullptr
    const float* unit;    // [n, d] or nullptr
    bool dense;           // the reference would hold a dense distance matrix (len(orig_dist) != 0, finch.py:142)
};

// clust_rank (finch.py:22-38) for a float64 level: unit rows, first neighbours, distances
static int rank_f64(const double* mat, int64_t n, int d, Scratch& unit, Scratch& nn, Scratch& dist, cudaStream_t st) {
    const int dp = d_pad_of(d);
    const bool screen = n >= SCREEN_MIN_ROWS;
    Scratch ub;
    SLIC_CUDA_OK(unit.alloc((size_t)n * d * sizeof(double), st));
    SLIC_CUDA_OK(nn.alloc((size_t)n * sizeof(int), st));
    SLIC_CUDA_OK(dist.alloc((size_t)n * sizeof(double), st));
    if (screen) SLIC_CUDA_OK(ub.alloc((size_t)n * dp * 2, st));
    SLIC_PROPAGATE(slic_normalize_rows(mat, n, d, SLIC_F64, unit.ptr, nullptr, screen ? ub.as<uint16_t>() : nullptr, dp, st));
    if (screen)
        return slic_nn_top1(unit.ptr, ub.as<uint16_t>(), n, unit.ptr, ub.as<uint16_t>(), n, d, dp, SLIC_F64, 0, 0.f,
                            nn.as<int32_t>(), dist.ptr, nullptr, st);
    return slic_nn_exact_top1(unit.ptr, nullptr, n, unit.ptr, n, d, SLIC_F64, 0, nn.as<int32_t>(), dist.ptr, st);
}

// Everything after the level-0 search.  labels_out: device [n, capacity] ints, filled as [n, P] row-major.
static int finch_levels(const float* data, int64_t n, int d, const Level0& l0, bool ensure_early_exit, int capacity,
                        int* labels_out, int* num_clust_host, int* num_levels_host, float* min_sim_host,
                        int* has_min_sim_host, cudaStream_t st) {
    std::vector<ScratchPtr> cols;     // composed labels of every kept level, [n] each
    std::vector<int> num_clust;
    Scratch count_dev;
    SLIC_CUDA_OK(count_dev.alloc(sizeof(int), st));

    // level 0: components of the first-neighbour graph (finch.py:136), centroids (:137)
    cols.push_back(new_scratch());
    SLIC_CUDA_OK(cols[0]->alloc((size_t)n * sizeof(int), st));
    SLIC_PROPAGATE(slic_finch_components(l0.nn, n, 0, 0.0, nullptr, 0, SLIC_F32, nullptr, cols[0]->as<int32_t>(),
                                         count_dev.as<int32_t>(), st));
    int cur = 0;
    SLIC_PROPAGATE(read_back(&cur, count_dev.ptr, sizeof(int), st));
    num_clust.push_back(cur);
    ScratchPtr sums = new_scratch(), counts = new_scratch(), means = new_scratch();
    SLIC_CUDA_OK(sums->alloc((size_t)cur * d * sizeof(double), st));
    SLIC_CUDA_OK(counts->alloc((size_t)cur * sizeof(int), st));
    SLIC_CUDA_OK(means->alloc((size_t)cur * d * sizeof(double), st));
    SLIC_PROPAGATE(slic_cluster_sums(data, cols[0]->as<int32_t>(), n, d, cur, sums->as<double>(), counts->as<int32_t>(),
                                     means->as<double>(), st));

    bool have_min_sim = false;
    float min_sim = 0.f;
    if (ensure_early_exit && l0.dense && l0.dist && l0.unit && n > 1) {    // finch.py:142-144
        Scratch ms;
        SLIC_CUDA_OK(ms.alloc(sizeof(float), st));
        SLIC_PROPAGATE(slic_finch_min_sim(l0.nn, n, l0.unit, d, SLIC_F32, l0.dist, ms.as<float>(), st));
        SLIC_PROPAGATE(read_back(&min_sim, ms.ptr, sizeof(float), st));
        have_min_sim = true;
    }

    int exit_clust = 2;
    while (exit_clust > 1) {                                              // finch.py:151
        const int64_t m = num_clust.back();
        if (m == 1) break;   // a single centroid links to itself: one cluster, the level is dropped (:160-163)
        Scratch unit, nn, dist, u;
        SLIC_PROPAGATE(rank_f64(means->as<double>(), m, d, unit, nn, dist, st));
        SLIC_CUDA_OK(u.alloc((size_t)m * sizeof(int), st));
        const bool filter = have_min_sim && m <= FLANN_THRESHOLD;          // finch.py:51-52 (needs dense distances)
        SLIC_PROPAGATE(slic_finch_components(nn.as<int32_t>(), m, filter ? 1 : 0, (double)min_sim, unit.ptr, d, SLIC_F64,
                                             dist.ptr, u.as<int32_t>(), count_dev.as<int32_t>(), st));
        SLIC_PROPAGATE(read_back(&cur, count_dev.ptr, sizeof(int), st));
        exit_clust = num_clust.back() - cur;
        if (cur == 1 || exit_clust < 1) break;                             // finch.py:160-163: level dropped
        if ((int)cols.size() >= capacity || (int)cols.size() >= FINCH_MAX_LEVELS) {
            set_error("finch: more than %d partitions; enlarge the label buffer", (int)cols.size());
            return SLIC_ERR_OVERFLOW;
        }
        cols.push_back(new_scratch());
        SLIC_CUDA_OK(cols.back()->alloc((size_t)n * sizeof(int), st));
        SLIC_PROPAGATE(slic_compose_labels(cols[cols.size() - 2]->as<int32_t>(), u.as<int32_t>(), n,
                                           cols.back()->as<int32_t>(), st));             // get_merge, finch.py:74-79
        ScratchPtr s2 = new_scratch(), c2 = new_scratch(), m2 = new_scratch();
        SLIC_CUDA_OK(s2->alloc((size_t)cur * d * sizeof(double), st));
        SLIC_CUDA_OK(c2->alloc((size_t)cur * sizeof(int), st));
        SLIC_CUDA_OK(m2->alloc((size_t)cur * d * sizeof(double), st));
        SLIC_PROPAGATE(slic_merge_cluster_sums(sums->as<double>(), counts->as<int32_t>(), u.as<int32_t>(), m, d, cur,
                                               s2->as<double>(), c2->as<int32_t>(), m2->as<double>(), st));
        sums.swap(s2);
        counts.swap(c2);
        means.swap(m2);
        num_clust.push_back(cur);
    }

    const int levels = (int)cols.size();
    ColumnPtrs cp;
    for (int l = 0; l < levels; ++l) cp.col[l] = cols[l]->as<int>();
    stack_columns_kernel<<<(unsigned)ceil_div(n * levels, 256), 256, 0, st>>>(cp, levels, n, labels_out);
    SLIC_LAUNCH_OK();
    for (int l = 0; l < levels; ++l) num_clust_host[l] = num_clust[l];
    *num_levels_host = levels;
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
};

// Runs on the host right after the gated screen kernel has been launched on main_stream: enqueue, chunk by chunk,
// copy -> normalise -> open the gate on the copy stream, then make the main stream wait for the last chunk (the
// exact re-rank reads float32 unit rows of every chunk).
static int run_upload(void* ctx) {
    Upload* up = static_cast<Upload*>(ctx);
    if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_up0, up->copy_stream));
    for (int c = 0; c < up->num_chunks; ++c) {
        const int64_t r0 = (int64_t)c * up->chunk_rows;
        const int64_t rows = (r0 + up->chunk_rows <= up->n ? up->chunk_rows : up->n - r0);
        SLIC_CUDA_OK(cudaMemcpyAsync(up->data + r0 * up->d, up->x_host + r0 * up->d, (size_t)rows * up->d * sizeof(float),
                                     cudaMemcpyHostToDevice, up->copy_stream));
        SLIC_PROPAGATE(slic_normalize_rows(up->data + r0 * up->d, rows, up->d, SLIC_F32, up->unit + r0 * up->d, nullptr,
                                           up->ub + r0 * up->d_pad, up->d_pad, up->copy_stream));
        if (up->gates) {
            open_gate_kernel<<<1, 32, 0, up->copy_stream>>>(up->gates + c);
            SLIC_LAUNCH_OK();
        }
    }
    if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_up1, up->copy_stream));
    SLIC_CUDA_OK(cudaEventRecord(up->done, up->copy_stream));
    SLIC_CUDA_OK(cudaStreamWaitEvent(up->main_stream, up->done, 0));
    return SLIC_OK;
}

struct HostStreams {
    cudaStream_t main = nullptr, copy = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    int init() {
        SLIC_CUDA_OK(cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking));
        SLIC_CUDA_OK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        SLIC_CUDA_OK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        SLIC_CUDA_OK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        return SLIC_OK;
    }
    ~HostStreams() {
        if (main) cudaStreamSynchronize(main);
        if (copy) cudaStreamSynchronize(copy);
        if (ready) cudaEventDestroy(ready);
        if (done) cudaEventDestroy(done);
        if (copy) cudaStreamDestroy(copy);
        if (main) cudaStreamDestroy(main);
    }
};

static int finch_host_impl(const float* x_host, int64_t n, int d, const int64_t* initial_rank_host, bool ensure_early_exit,
                           int capacity, int* labels_out_host, int* num_clust_host, int* num_levels_host,
                           float* min_sim_host, int* has_min_sim_host, HostStreams& hs) {
    const int dp = d_pad_of(d);
    cudaStream_t st = hs.main;
    Scratch data, unit, ub, nn, dist, gates, labels, rank64;
    SLIC_CUDA_OK(data.alloc((size_t)n * d * sizeof(float), st));
    SLIC_CUDA_OK(nn.alloc((size_t)n * sizeof(int), st));
    SLIC_CUDA_OK(labels.alloc((size_t)n * capacity * sizeof(int), st));
    Level0 l0 = {nn.as<int>(), nullptr, nullptr, false};

    if (initial_rank_host) {
        // finch.py:22-23: the caller's neighbours are used as they are; no distances => no min_sim (:142)
        SLIC_CUDA_OK(rank64.alloc((size_t)n * sizeof(int64_t), st));
        SLIC_CUDA_OK(cudaMemcpyAsync(rank64.ptr, initial_rank_host, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        convert_rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(rank64.as<int64_t>(), n, nn.as<int>());
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
                     hs.done};
        if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_t0, st));
        int guest_threads = 0, guest_regs = 0;
        SLIC_PROPAGATE(normalize_kernel_shape(&guest_threads, &guest_regs));
        if (screen && n >= GATED_MIN_ROWS && gated_chunks() > 1 &&
            screen_can_overlap_upload(n, dp, guest_threads, guest_regs)) {
            // pipelined: screen first, upload behind it
            up.num_chunks = gated_chunks();
            up.chunk_rows = ceil_div(ceil_div(n, up.num_chunks), 256) * 256;
            up.num_chunks = (int)ceil_div(n, up.chunk_rows);
            SLIC_CUDA_OK(gates.alloc((up.num_chunks + 1) * sizeof(int), st));
            SLIC_CUDA_OK(cudaMemsetAsync(gates.ptr, 0, (up.num_chunks + 1) * sizeof(int), st));
            up.gates = gates.as<int>();
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
                                             st));
        } else {
            SLIC_CUDA_OK(cudaEventRecord(hs.ready, st));
            SLIC_CUDA_OK(cudaStreamWaitEvent(hs.copy, hs.ready, 0));
            if (!screen) up.ub = nullptr;
            // one chunk, no gates: copy + normalise on the copy stream, then search
            if (screen) {
                SLIC_PROPAGATE(run_upload(&up));
                SLIC_PROPAGATE(slic_nn_top1(unit.ptr, ub.as<uint16_t>(), n, unit.ptr, ub.as<uint16_t>(), n, d, dp, SLIC_F32, 0,
                                            0.f, nn.as<int32_t>(), dist.ptr, nullptr, st));
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
        if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_t_search, st));
    }
    int levels = 0;
    SLIC_PROPAGATE(finch_levels(data.as<float>(), n, d, l0, ensure_early_exit, capacity, labels.as<int>(), num_clust_host,
                                &levels, min_sim_host, has_min_sim_host, st));
    *num_levels_host = levels;
    SLIC_CUDA_OK(cudaMemcpyAsync(labels_out_host, labels.ptr, (size_t)n * levels * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (g_host_trace) SLIC_CUDA_OK(cudaEventRecord(g_t_end, st));
    SLIC_CUDA_OK(cudaStreamSynchronize(st));
    return SLIC_OK;
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
    Scratch unit, ub, nn, dist;
    Level0 l0;
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
            SLIC_PROPAGATE(slic_nn_top1(unit.ptr, ub.as<uint16_t>(), n, unit.ptr, ub.as<uint16_t>(), n, d, dp, SLIC_F32, 0, 0.f,
                                        nn.as<int32_t>(), dist.ptr, nullptr, st));
        else
            SLIC_PROPAGATE(slic_nn_exact_top1(unit.ptr, nullptr, n, unit.ptr, n, d, SLIC_F32, 0, nn.as<int32_t>(), dist.ptr,
                                              st));
        l0.nn = nn.as<int>();
        l0.dist = dist.as<float>();
        l0.unit = unit.as<float>();
        l0.dense = n <= FLANN_THRESHOLD;
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
