// Host side of the C ABI (include/fastsk_b200.h): argument checks, the combination queue,
// virtual-stream bookkeeping of approx mode, and the launch sequence of fsk_kernels.cuh.
// Mirrors FastSK::compute_kernel (fastsk.cpp:30-118) and KernelFunction::compute_kernel /
// kernel_build_parallel (fastsk_kernel.cpp:24-106, 145-322) without any CPU compute path.
#include "../../include/fastsk_b200.h"
#include "fsk_kernels.cuh"
#include "fsk_dense.cuh"
#include "fsk_bucket.cuh"
#include "fsk_segment.cuh"

#include <algorithm>
#if defined(__linux__)
#include <sys/syscall.h>
#include <unistd.h>
#endif
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <functional>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace fsk;

namespace {

std::string g_create_error;

enum RecMode { MODE_R32 = 0, MODE_R64 = 1, MODE_KV = 2 };
constexpr int SEG_TICKET = 16;   // d_ticket[0 .. MAX_PASS) belong to the sort passes
constexpr int HEAVY_COUNT = 32;  // d_ticket[32]: heavy runs listed by segment_kernel in this batch (cleared with the tickets)
enum ProfClass { PC_PACK = 0, PC_SORT, PC_SEGMENT, PC_ACCUMULATE, PC_WELFORD, PC_NORMALISE, PC_COUNT };

struct ProfSpan {
    cudaEvent_t a, b;
    int cls;
};

}  // namespace

struct fsk_handle {
    // parameters (FastSK::FastSK, fastsk.cpp:19-28)
    int g = 0, m = 0, k = 0, t = -1;
    bool approx = false, skip_variance = false;
    double delta = 0.025;
    int max_iters = -1;
    // configuration
    int device = 0, rank = 0, world = 1;
    bool have_seed = false;
    uint64_t seed = 0;
    bool run_seed_set = false;       // in-process team: the leader draws the wall-clock seed once per upload for every member
    uint64_t run_seed = 0;
    // in-process multi-GPU (fsk_set_devices): team[0] is this handle, team[i] drives devices[i] from its own host thread
    std::vector<int> devices;
    std::vector<fsk_handle*> team;
    fsk_handle* leader = nullptr;    // set on team members other than the leader
    // sharded finalisation: the partial kernels of all ranks in rank order (own buffer included), and the output rows held here
    std::vector<const void*> peer_parts;
    std::vector<void*> ipc_opened;   // peer mappings this handle opened with cudaIpcOpenMemHandle
    bool weights_auto = false;       // out_weights came from the team's own probe (dropped when the team changes)
    std::vector<double> out_weights; // share of the output rows per rank (empty: equal shares); 0 = this rank hands no rows back
    bool sharded = false;            // d_train / d_test hold only the rows [tr_r0, +tr_nr) / [te_r0, +te_nr)
    int64_t tr_r0 = 0, tr_nr = 0, te_r0 = 0, te_nr = 0;
    size_t train_cap = 0, test_cap = 0;
    std::vector<int32_t> user_queue;
    int opt_batch = 0;
    int opt_acc_path = 0;            // 0 auto, 1 global RED, 2 row-stationary shared memory, 3 dense tensor-core contraction
    bool safe_rank = false;          // onesweep ranking: false = one atomic per key, verified afterwards; true = match masks
    int opt_rows_threads = 0;        // threads per row CTA of the accumulate (0 = by N)
    int opt_overlap = 0;             // 1 = pre-pass of the next batch on its own low-priority stream (measured: no gain, the
                                     // row CTAs own the whole SM's shared memory; profiles/r01_overlap_experiment.txt)
    int seg_rows = SEG_ROWS_DEFAULT;
    int opt_seg_fused = 0;           // 0 auto (= off), 1 off, 2 on: fused last sort pass + segmentation (fsk_bucket.cuh) for two-digit keys
    bool fused_seg = false;
    uint32_t image_cap = 0;          // ids in the shared-memory image of a bucket
    int opt_acc_prefetch = 1;        // 1 = L2 prefetch of the next chunk's id ranges (0: off, for the A/B measurement)
    int opt_acc_unroll = 2;          // id units in flight per lane of the accumulate (2 or 4)
    int opt_wave = 32;               // accumulate launch = opt_wave x (CTAs resident on the chip) rows (4: 171.4, 16: 169.7, 32: 168.9, 64: 168.6, one launch: 169.5 ms per 384 combinations)
    bool profile = false;
    std::string err;

    // state
    cudaStream_t stream = nullptr;       // accumulate, Welford, normalisation, copies
    cudaStream_t pre_stream = nullptr;   // pack / sort / segment of the NEXT batch, overlapping the accumulate of the current one
    cudaStream_t ls = nullptr;           // stream the launch helpers and profiling spans currently target
    cudaEvent_t ev_pre[2] = {nullptr, nullptr}, ev_acc[2] = {nullptr, nullptr}, ev_sync = nullptr;
    int64_t batch_index = 0;
    int buf = 0;
    bool uploaded = false, built = false, finalized = false;
    int64_t n_train = 0, n_test = 0, N = 0, nfeat = 0, n_pairs = 0, n_train_pairs = 0, ncomb = 0;
    int A = 0, b = 0, cpw = 0, NW = 1;
    bool gw32 = false;
    int keybits = 0, idbits = 0, mode = MODE_R32, rec_bytes = 4;
    SortPlan plan{};
    int B = 1;                       // slots per batch
    uint32_t sort_tiles = 0, seg_tiles = 0;
    int sort_items = 16;
    bool variance_mode = false;
    int T_eff = 1;

    // device buffers
    void* d_gw0 = nullptr;
    uint64_t* d_gw1 = nullptr;
    uint32_t* d_wseq = nullptr;
    void *d_recA = nullptr, *d_recB = nullptr;
    uint32_t *d_valA = nullptr, *d_valB = nullptr;
    unsigned char* d_zero = nullptr;   // [ghist | tickets | status] cleared once per batch
    size_t zero_bytes = 0;
    uint32_t *d_ghist = nullptr, *d_ticket = nullptr, *d_status = nullptr, *d_seg_status = nullptr;
    uint32_t *d_woff32 = nullptr, *d_fill = nullptr;   // window offsets per sequence; tasks filed so far per (slot, sequence)
    void* d_ids[2] = {nullptr, nullptr};               // sequence id of every sorted record (u16 when N <= 65536, else u32); double-buffered
    size_t ids_stride = 0;                             // per-slot stride of d_ids, a multiple of 64 elements
    uint32_t pad_mask = 7;                             // runs start on multiples of pad_mask + 1 ids in d_ids (a 16-byte unit or a 128-byte line)
    int opt_pad = 0;                                   // 0 auto, 1 unit, 2 line
    bool ids16 = false;
    uint2* d_task[2] = {nullptr, nullptr};
    // directory form of the segmentation (small key spaces): per-window keys + one task per (run, block of sequences)
    int opt_seg_dir = 0;                               // 0 auto, 1 off, 2 on
    int opt_dir_blocks = 32;                           // target number of row blocks (power-of-two block size)
    bool dir_mode = false;
    int dir_bshift = 0;
    uint32_t dir_nb = 0;
    uint16_t* d_wkey = nullptr;
    uint2* d_tdir[2] = {nullptr, nullptr};
    int opt_dense_u8 = 1;                              // dense regime: byte operands / int32 accumulators when no sequence has more than 255 windows (0: always fp16 / fp32)
    int opt_wf_regs = 1;                               // tensor-core variance mode: running means in registers across a round's slots (0: streamed through L2)
    int opt_fit_smem = 1;                              // accumulate launches ask for the shared memory of their longest row only
    int opt_pf_stride = 128;                           // L2 prefetch granularity of the accumulate's id ranges (64 or 128 bytes)
    int opt_seg_lean = 0;                              // 0 auto (on for records that carry the id), 1 off, 2 on: register-blocked segmentation
    bool lean_seg = false;
    uint32_t lean_tiles = 0;
    int opt_count_updates = 1;                         // profile: also count entries / runs / pair updates (0: spans only)
    bool rows_path = false;
    int rows_threads = 256;
    size_t rows_smem = 0;
    int64_t col_width = 0;                             // columns of K per shared-memory row window (>= N: one window)
    int col_windows = 1;
    int opt_acc_cols = 0;                              // 0 auto, else forced window width (tests)
    int wave_rows = 148;                               // rows per accumulate launch
    int64_t maxwin = 0;
    // dense regime (fsk_dense.cuh): per-sequence k-mer counts of a batch, contracted on the tensor cores
    bool dense_path = false;
    uint32_t nks = 0;                                  // k-mer columns per slot, padded to a multiple of DG_BK
    size_t dense_ld = 0;                               // fp16 elements per row of d_C (= B * nks)
    int dense_chunk = 1;                               // slots per GEMM: keeps every fp32 accumulator below 2^24
    __half* d_C = nullptr;
    // heavy runs of the sparse regime: runs longer than heavy_tau leave the row path for a tensor-core contraction
    int opt_heavy_tau = 0;                             // 0 auto, -1 off, > 0 forced threshold
    int opt_ids32 = 0;                                 // test hook: 32-bit id stream although N <= 65000
    int opt_gemm_shape = 0;                            // 0 auto, 1 one tile per CTA, 2 two tiles per CTA sharing B
    int opt_heavy_cap = 0;                             // 0 auto, else columns of d_H (tests: a small list overflows)
    uint32_t heavy_tau = 0;                            // 0 = feature off
    uint32_t heavy_tau_min = 0;                        // the break-even threshold the adaptive one never goes below
    uint32_t heavy_now = 0;                            // threshold in force for the batch being launched (0 while the feature sleeps)
    bool heavy_u8 = false;                             // byte columns in d_H (no sequence has more than 255 windows)
    uint32_t heavy_cap = 0;                            // columns of d_H = upper bound on the heavy runs of a batch
    __half* d_H = nullptr;
    uint2* d_heavy_list = nullptr;
    uint32_t* d_heavy_bits = nullptr;                  // [slot][heavy_bits_stride] in the per-batch zero region
    size_t heavy_bits_stride = 0;
    // adaptive: when a batch lists no heavy run the feature sleeps (no marking, no launches) and is probed again every 32nd
    // batch -- the sequences are the same for every combination, so run-length statistics hardly change between batches
    bool heavy_live = true, heavy_probe_pending = false;
    int heavy_idle = 0;
    uint32_t* h_heavy_count = nullptr;                 // pinned
    cudaEvent_t ev_heavy = nullptr;
    CUtensorMap tmap_H;
    uint32_t* d_pair_order = nullptr;                  // the same for pairs of tile rows (two-tile shape of the contraction)
    uint32_t pair_tiles = 0;
    uint32_t* d_tile_order = nullptr;                  // lower-triangle tiles (I << 16 | J) in L2-friendly launch order
    CUtensorMap tmap_C;
    bool wf_regs = false, dense_u8 = false;            // fixed at upload: form of the Welford contraction, byte operands (ld of d_C then in bytes)
    unsigned long long* d_Kint = nullptr;   // integer partial (exact / skip_variance), or per-slot Ks in variance mode
    int ks_slots = 1;
    std::vector<double*> d_Khat;            // one per local virtual stream
    double* d_Kf = nullptr;                 // fp64 partial (variance mode): sum of local K_hat
    double *d_diag = nullptr, *d_train = nullptr, *d_test = nullptr;
    double *d_block_sums = nullptr, *d_var = nullptr;
    WelfordSpec* d_wf = nullptr;            // variance mode: per-slot running means for the fused Welford flush
    bool wf_active = false;                 // the batch being launched carries d_wf
    int wf_groups = 0;                      // ... for this many virtual streams (groups of consecutive slots)
    int wf_depth = 1;                       // variance mode: iterations of a stream speculated per round (1 = none, no second mean buffer)
    int opt_spec_depth = 0;                 // 0 auto, else forced upper bound
    uint32_t sums_stride = WELFORD_BLOCKS;  // partial sums per slot
    unsigned long long* d_counters = nullptr;   // entries, runs, pair updates
    uint32_t* d_flag = nullptr;                 // set by segment_kernel when a sorted batch is not non-decreasing

    std::vector<int32_t> queue;
    std::vector<double> stdevs;

    // statistics
    int64_t combos_done = 0, launches = 0;
    int rank_fallbacks = 0;
    double ms[PC_COUNT] = {0, 0, 0, 0, 0, 0};
    std::vector<ProfSpan> spans;
    std::vector<cudaEvent_t> event_pool;
};

namespace {

int fail(fsk_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(h, e_ == cudaErrorMemoryAllocation ? FSK_ENOMEM : FSK_ECUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
    } while (0)

// Device allocations go through a process-wide cache of exact-size blocks: a handle that computes again on inputs of the same
// shape (or a new handle doing so: every FastSK object of a benchmark loop) re-uses the ~40 GB of buffers of the previous one
// instead of ~30 cudaFree + cudaMalloc calls, each of which synchronises the device.  Blocks are plain cudaMalloc memory
// (CUDA IPC needs that for the peer-mapped partial kernels).  A miss that cannot be served drops the whole cache of that
// device and retries; fsk_trim_cache() returns everything to the driver.
struct DevCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_blocks;       // (device, bytes) -> block
    std::unordered_map<void*, std::pair<int, size_t>> live;
    size_t cached_bytes = 0;
    std::unordered_map<void*, int> exported;                         // blocks other processes may have mapped (CUDA IPC)
    void trim(int device, bool exported_too = false) {                // device < 0: all devices (caller holds mu)
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto it = free_blocks.begin(); it != free_blocks.end();) {
            // a block whose IPC handle went out stays allocated until fsk_trim_cache: a peer may still have it mapped
            if ((device < 0 || it->first.first == device) && (exported_too || !exported.count(it->second))) {
                cudaSetDevice(it->first.first);
                cudaFree(it->second);
                cached_bytes -= it->first.second;
                it = free_blocks.erase(it);
            } else ++it;
        }
        cudaSetDevice(cur);
    }
};
DevCache g_cache;

cudaError_t cached_malloc(void** p, size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache.mu);
    auto it = g_cache.free_blocks.find({dev, bytes});
    if (it != g_cache.free_blocks.end()) {
        *p = it->second;
        g_cache.free_blocks.erase(it);
        g_cache.cached_bytes -= bytes;
        g_cache.live[*p] = {dev, bytes};
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        g_cache.trim(dev);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) g_cache.live[*p] = {dev, bytes};
    return e;
}
void cached_free(void* p) {
    std::lock_guard<std::mutex> lk(g_cache.mu);
    auto it = g_cache.live.find(p);
    if (it == g_cache.live.end()) { cudaFree(p); return; }
    g_cache.free_blocks.insert({it->second, p});
    g_cache.cached_bytes += it->second.second;
    g_cache.live.erase(it);
}

// free device memory as the sizing decisions should see it: what the driver reports plus what this library holds in its
// cache (a miss that does not fit drops the cache)
cudaError_t mem_info_with_cache(size_t* free_b, size_t* total_b) {
    cudaError_t e = cudaMemGetInfo(free_b, total_b);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache.mu);
    for (auto& kv : g_cache.free_blocks)
        if (kv.first.first == dev) *free_b += kv.first.second;
    return cudaSuccess;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel) and only upwards: ~25 of these per upload were
// a millisecond of a small build
template <typename F>
cudaError_t smem_opt_in(F func, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> have;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    int& cur = have[{dev, (const void*)func}];
    if (bytes <= cur) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

template <typename T>
int dev_alloc(fsk_handle* h, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    CU(cached_malloc((void**)p, count * sizeof(T)));
    return FSK_OK;
}
#define ALLOC(ptr, count)                                   \
    do {                                                    \
        int rc_ = dev_alloc(h, &(ptr), (size_t)(count));    \
        if (rc_) return rc_;                                \
    } while (0)

template <typename T>
void dev_free(T*& p) {
    if (p) cached_free((void*)p);
    p = nullptr;
}

void release_device(fsk_handle* h) {
    if (h->pre_stream) cudaStreamSynchronize(h->pre_stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    dev_free(h->d_gw0); dev_free(h->d_gw1); dev_free(h->d_wseq);
    dev_free(h->d_recA); dev_free(h->d_recB); dev_free(h->d_valA); dev_free(h->d_valB);
    dev_free(h->d_zero);
    h->d_ghist = h->d_ticket = h->d_status = h->d_seg_status = nullptr;
    dev_free(h->d_woff32); dev_free(h->d_fill); dev_free(h->d_C); dev_free(h->d_tile_order); dev_free(h->d_pair_order); dev_free(h->d_H); dev_free(h->d_heavy_list);
    for (int i = 0; i < 2; ++i) { dev_free(h->d_ids[i]); dev_free(h->d_task[i]); dev_free(h->d_tdir[i]); }
    dev_free(h->d_wkey);
    dev_free(h->d_Kint); dev_free(h->d_Kf);
    for (auto& p : h->d_Khat) dev_free(p);
    h->d_Khat.clear();
    h->ipc_opened.clear();
    h->peer_parts.clear();
    h->sharded = false;
    h->train_cap = h->test_cap = 0;
    dev_free(h->d_diag); dev_free(h->d_train); dev_free(h->d_test);
    dev_free(h->d_block_sums); dev_free(h->d_var); dev_free(h->d_wf); dev_free(h->d_counters); dev_free(h->d_flag);
    for (auto& s : h->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    h->spans.clear();
    for (auto e : h->event_pool) cudaEventDestroy(e);
    h->event_pool.clear();
    h->uploaded = h->built = h->finalized = false;
}

// FSK_TRACE=1 in the environment: wall-clock milliseconds of the host-side phases on stderr
struct Trace {
    bool on;
    const char* what;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(const char* w) : on(getenv("FSK_TRACE") != nullptr), what(w), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* phase) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[fsk] %s: %s %.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

int64_t nchoosek64(int n, int k) {   // C(n,k); the reference's int version is exact for g <= 20 (shared.cpp:335-345)
    if (k < 0 || k > n) return 0;
    if (2 * k > n) k = n - k;
    int64_t r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

// kept positions of combination number `idx` in the lexicographic order getCombinations emits
// (shared.cpp:347-360), by combinatorial unranking instead of re-enumerating all C(g,k) subsets
// on every iteration (fastsk_kernel.cpp:216-221).
void unrank_combination(int g, int k, int64_t idx, int* pos) {
    int p = 0;
    for (int d = 0; d < k; ++d) {
        for (;; ++p) {
            const int64_t with_p = nchoosek64(g - 1 - p, k - 1 - d);
            if (idx < with_p) break;
            idx -= with_p;
        }
        pos[d] = p++;
    }
}

int ceil_log2(int64_t v) {   // bits needed to store values 0 .. v-1 (at least 1)
    int b = 1;
    while (((int64_t)1 << b) < v) ++b;
    return b;
}

// --- profiling spans -------------------------------------------------------------------------
cudaEvent_t get_event(fsk_handle* h) {
    if (!h->event_pool.empty()) {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void resolve_spans(fsk_handle* h) {
    if (h->spans.empty()) return;
    cudaStreamSynchronize(h->pre_stream);
    cudaStreamSynchronize(h->stream);
    for (auto& s : h->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) h->ms[s.cls] += ms;
        h->event_pool.push_back(s.a);
        h->event_pool.push_back(s.b);
    }
    h->spans.clear();
}
struct Span {
    fsk_handle* h;
    ProfSpan s{};
    bool on;
    Span(fsk_handle* h_, int cls) : h(h_), on(h_->profile) {
        if (!on) return;
        s.cls = cls;
        s.a = get_event(h);
        s.b = get_event(h);
        cudaEventRecord(s.a, h->ls);
    }
    ~Span() {
        if (!on) return;
        cudaEventRecord(s.b, h->ls);
        h->spans.push_back(s);
    }
};

// --- typed launch helpers ----------------------------------------------------------------------
template <typename RecT, bool KV>
int launch_pack(fsk_handle* h, int nb, const BatchSpec& spec) {
    dim3 grid(nb, (unsigned)((h->nfeat + 256 * PACK_ITEMS - 1) / (256 * PACK_ITEMS)));
    RecT* rec = (RecT*)h->d_recA;
    const uint32_t n = (uint32_t)h->nfeat;
    if (h->NW == 2)
        pack_hist_kernel<RecT, KV, uint64_t, 2><<<grid, 256, 0, h->ls>>>((const uint64_t*)h->d_gw0, h->d_gw1, h->d_wseq, n, rec,
                                                                           h->d_valA, h->d_ghist, spec, h->plan, h->idbits, h->dir_mode ? h->d_wkey : nullptr);
    else if (h->gw32)
        pack_hist_kernel<RecT, KV, uint32_t, 1><<<grid, 256, 0, h->ls>>>((const uint32_t*)h->d_gw0, nullptr, h->d_wseq, n, rec,
                                                                           h->d_valA, h->d_ghist, spec, h->plan, h->idbits, h->dir_mode ? h->d_wkey : nullptr);
    else
        pack_hist_kernel<RecT, KV, uint64_t, 1><<<grid, 256, 0, h->ls>>>((const uint64_t*)h->d_gw0, nullptr, h->d_wseq, n, rec,
                                                                           h->d_valA, h->d_ghist, spec, h->plan, h->idbits, h->dir_mode ? h->d_wkey : nullptr);
    h->launches++;
    CU(cudaGetLastError());
    return FSK_OK;
}

template <typename RecT, bool KV, int ITEMS>
size_t sort_smem() {
    return sizeof(RecT) * SORT_THREADS * ITEMS + (KV ? 4 * SORT_THREADS * ITEMS : 0) + 4 * (8 * RADIX + RADIX + 8 + 8 * RADIX);
}

template <typename RecT, bool KV, int ITEMS>
int launch_sort(fsk_handle* h, int nb) {
    const uint32_t n = (uint32_t)h->nfeat;
    const size_t smem = sort_smem<RecT, KV, ITEMS>();
    // opt in to more than 48 KB of dynamic shared memory (per device, so not cached across handles)
    auto kernel = h->safe_rank ? onesweep_kernel<RecT, KV, ITEMS, false> : onesweep_kernel<RecT, KV, ITEMS, true>;
    CU(smem_opt_in(kernel, (int)smem));
    const bool fused = h->fused_seg && !h->safe_rank;     // fused: partition by the high digit only (bucket_segment_kernel does the rest)
    for (int p = fused ? h->plan.npass - 1 : 0; p < h->plan.npass; ++p) {
        const int shift = (KV ? 0 : h->idbits) + h->plan.shift[p];
        uint32_t* status = h->d_status + (size_t)p * h->B * h->sort_tiles * RADIX;
        kernel<<<h->sort_tiles * nb, SORT_THREADS, smem, h->ls>>>(
            (const RecT*)h->d_recA, (RecT*)h->d_recB, h->d_valA, h->d_valB, n, h->sort_tiles, (uint32_t)nb, shift, h->plan.bits[p],
            h->d_ghist + (size_t)p * RADIX, status, h->d_ticket + p);
        h->launches++;
        CU(cudaGetLastError());
        std::swap(h->d_recA, h->d_recB);
        std::swap(h->d_valA, h->d_valB);
    }
    return FSK_OK;
}

template <typename RecT, bool KV>
int launch_segment(fsk_handle* h, int nb) {
    const uint32_t n = (uint32_t)h->nfeat;
    unsigned long long* stat = (h->profile && h->opt_count_updates) ? h->d_counters : nullptr;
    if (!h->dir_mode) {
        init_fill_kernel<<<dim3((unsigned)((h->N + 255) / 256), nb), 256, 0, h->ls>>>(h->d_fill, h->d_woff32, (uint32_t)h->N);
        h->launches++;
    }
    if (h->fused_seg && !h->safe_rank) {
        if (!std::is_same<RecT, uint32_t>::value || KV) return fail(h, FSK_ESTATE, "fused segmentation needs 32-bit records");
        const int lo_bits = h->plan.bits[0], hi_bits = h->plan.bits[1];
        dim3 grid(1u << hi_bits, (unsigned)nb);
        const size_t smem = bucket_smem_bytes(h->image_cap);
        const uint32_t* gh = h->d_ghist + (size_t)(h->plan.npass - 1) * RADIX;
#define BK_ARGS (const uint32_t*)h->d_recA, n, gh, h->idbits, h->idbits + h->plan.shift[0], lo_bits, (uint32_t)h->N, h->ids_stride, h->pad_mask, \
                h->image_cap, h->d_fill, (uint16_t*)h->d_ids[h->buf], h->d_task[h->buf], h->d_flag, stat
        if (stat) bucket_segment_kernel<true><<<grid, BK_THREADS, smem, h->ls>>>(BK_ARGS);
        else bucket_segment_kernel<false><<<grid, BK_THREADS, smem, h->ls>>>(BK_ARGS);
#undef BK_ARGS
        h->launches++;
        CU(cudaGetLastError());
        return FSK_OK;
    }
    if (h->lean_seg) {
        if (KV) return fail(h, FSK_ESTATE, "the register-blocked segmentation needs records that carry the sequence id");
        // (writes the fill between the runs itself: no memset of the id stream)
        const unsigned lgrid = h->lean_tiles * (unsigned)nb;
        const int lush = h->ids16 ? 3 : 2;
        const size_t lsmem = (size_t)lean_cap(sizeof(RecT)) * (h->ids16 ? 2 : 4);
#define LEAN_ARGS(IDT) (const RecT*)h->d_recA, n, h->lean_tiles, h->ids_stride, h->idbits, (uint32_t)h->N, lush, h->pad_mask, h->d_fill, \
                  (IDT*)h->d_ids[h->buf], h->d_task[h->buf], h->d_seg_status, h->d_ticket + SEG_TICKET, h->d_flag, stat, h->heavy_now, \
                  h->d_ticket + HEAVY_COUNT, h->d_heavy_list, h->heavy_cap, h->d_heavy_bits, h->heavy_bits_stride, h->d_tdir[h->buf], \
                  h->dir_bshift, h->dir_nb, h->keybits
#define LEAN_GO(IDT, DR, HV, ST) segment_lean_kernel<RecT, IDT, DR, HV, ST><<<lgrid, LEAN_THREADS, lsmem, h->ls>>>(LEAN_ARGS(IDT))
#define LEAN_HS(IDT, DR) do { if (hv) { if (stt) LEAN_GO(IDT, DR, true, true); else LEAN_GO(IDT, DR, true, false); } \
                              else { if (stt) LEAN_GO(IDT, DR, false, true); else LEAN_GO(IDT, DR, false, false); } } while (0)
        const bool hv = h->heavy_now != 0, stt = stat != nullptr;
        if (h->ids16) { if (h->dir_mode) LEAN_HS(uint16_t, true); else LEAN_HS(uint16_t, false); }
        else { if (h->dir_mode) LEAN_HS(uint32_t, true); else LEAN_HS(uint32_t, false); }
#undef LEAN_HS
#undef LEAN_GO
#undef LEAN_ARGS
        h->launches++;
        CU(cudaGetLastError());
        return FSK_OK;
    }
    // the gaps between the aligned runs must read as "no sequence": 0xFF.. clamps to the dump word in the accumulate
    CU(cudaMemsetAsync(h->d_ids[h->buf], 0xff, (size_t)nb * h->ids_stride * (h->ids16 ? 2 : 4), h->ls));
    const unsigned grid = h->seg_tiles * (unsigned)nb;
    const int ush = h->ids16 ? 3 : 2;
#define SEG_ARGS (const RecT*)h->d_recA, h->d_valA, n, h->seg_tiles, h->ids_stride, h->idbits, (uint32_t)h->N, ush, h->pad_mask, h->d_fill
#define SEG_ARGS2 h->d_task[h->buf], h->d_seg_status, h->d_ticket + SEG_TICKET, h->d_flag, stat, h->heavy_now, \
                  h->d_ticket + HEAVY_COUNT, h->d_heavy_list, h->heavy_cap, h->d_heavy_bits, h->heavy_bits_stride
    if (h->dir_mode) {
        constexpr int MB = sizeof(RecT) == 4 ? 3 : 2;
#define DIR_ARGS , h->d_tdir[h->buf], h->dir_bshift, h->dir_nb, h->keybits
#define SEG_DIR(IDT, HV, ST) segment_kernel<RecT, false, IDT, HV, SEG_ROWS_DEFAULT, MB, true, ST><<<grid, SEG_THREADS, 0, h->ls>>>(SEG_ARGS, (IDT*)h->d_ids[h->buf], SEG_ARGS2 DIR_ARGS)
        if (KV) return fail(h, FSK_ESTATE, "the directory form needs records that carry the sequence id");
        const bool hv = h->heavy_now != 0, st = stat != nullptr;
        if (h->ids16) { if (hv) { if (st) SEG_DIR(uint16_t, true, true); else SEG_DIR(uint16_t, true, false); } else { if (st) SEG_DIR(uint16_t, false, true); else SEG_DIR(uint16_t, false, false); } }
        else { if (hv) { if (st) SEG_DIR(uint32_t, true, true); else SEG_DIR(uint32_t, true, false); } else { if (st) SEG_DIR(uint32_t, false, true); else SEG_DIR(uint32_t, false, false); } }
#undef SEG_DIR
#undef DIR_ARGS
    } else if (h->ids16 && h->heavy_now)
        segment_kernel<RecT, KV, uint16_t, true><<<grid, SEG_THREADS, 0, h->ls>>>(SEG_ARGS, (uint16_t*)h->d_ids[h->buf], SEG_ARGS2);
    else if (h->ids16)
        segment_kernel<RecT, KV, uint16_t, false><<<grid, SEG_THREADS, 0, h->ls>>>(SEG_ARGS, (uint16_t*)h->d_ids[h->buf], SEG_ARGS2);
    else if (h->heavy_now)
        segment_kernel<RecT, KV, uint32_t, true><<<grid, SEG_THREADS, 0, h->ls>>>(SEG_ARGS, (uint32_t*)h->d_ids[h->buf], SEG_ARGS2);
    else
        segment_kernel<RecT, KV, uint32_t, false><<<grid, SEG_THREADS, 0, h->ls>>>(SEG_ARGS, (uint32_t*)h->d_ids[h->buf], SEG_ARGS2);
#undef SEG_ARGS
#undef SEG_ARGS2
    h->launches++;
    CU(cudaGetLastError());
    return FSK_OK;
}

template <typename IdT>
int launch_accumulate(fsk_handle* h, int nb, unsigned long long* K, size_t slot_stride) {
    const uint32_t n = (uint32_t)h->nfeat;
    const IdT* ids = (const IdT*)h->d_ids[h->buf];
    if (h->rows_path) {
        // slot_stride != 0 (variance mode): every slot adds into its own K; else all slots add into one K
        // variance mode (wf_active): one group per virtual stream, its slots applied in order; else all slots add into one K
        const int groups = h->wf_active ? h->wf_groups : (slot_stride ? nb : 1), per_group = slot_stride ? 1 : nb;
        const int wave = std::max(1, h->wave_rows / groups);
        // one pass over the rows per column window of K (a single window unless N columns exceed shared memory)
        for (int64_t col0 = 0, win = 0; col0 < h->N; col0 += h->col_width, ++win)
        for (int64_t hi = h->N - 1; hi >= col0; hi -= wave) {
            dim3 grid((unsigned)std::min<int64_t>(wave, hi - col0 + 1), groups);
            auto kern = h->dir_mode ? (h->opt_acc_prefetch == 0 ? accumulate_rows_kernel<unsigned long long, IdT, 2, false, true>
                                       : h->opt_acc_unroll == 4 ? accumulate_rows_kernel<unsigned long long, IdT, 4, true, true>
                                                                : accumulate_rows_kernel<unsigned long long, IdT, 2, true, true>)
                        : h->opt_acc_prefetch == 0 ? accumulate_rows_kernel<unsigned long long, IdT, 2, false>
                        : h->opt_acc_unroll == 4 ? accumulate_rows_kernel<unsigned long long, IdT, 4>
                                                 : accumulate_rows_kernel<unsigned long long, IdT, 2>;
            DirSpec dir;
            dir.wkey = h->d_wkey; dir.tdir = h->d_tdir[h->buf]; dir.bshift = h->dir_bshift; dir.keybits = h->keybits; dir.nb = h->dir_nb; dir.pf_stride = (uint32_t)h->opt_pf_stride;
            // shared memory of the launch = its longest row (the launches walk the rows from the longest down): the shorter half of
            // the rows then fits two CTAs per SM
            const size_t smem_launch = h->opt_fit_smem ? std::min(h->rows_smem, (size_t)(std::min<int64_t>(hi - col0 + 1, h->col_width) + 32) * 4)
                                                       : h->rows_smem;
            kern<<<grid, h->rows_threads, smem_launch, h->ls>>>(
                ids, h->ids_stride, h->d_task[h->buf], h->d_woff32, n, (uint32_t)hi, per_group, K, slot_stride,
                h->wf_active ? h->d_wf : nullptr, (uint32_t)col0, (uint32_t)h->col_width, (uint32_t)(win * h->N), h->d_heavy_bits,
                h->heavy_bits_stride, dir);
            h->launches++;
        }
    } else {
        dim3 grid((unsigned)((h->nfeat + 255) / 256), nb);
        accumulate_global_kernel<unsigned long long, IdT><<<grid, 256, 0, h->ls>>>(ids, h->ids_stride, h->d_task[h->buf], h->d_wseq, n, K, slot_stride);
        h->launches++;
    }
    CU(cudaGetLastError());
    return FSK_OK;
}

// Dense regime: per-sequence k-mer counts of every slot, then K += C C^T on the tensor cores (fsk_dense.cuh).
int run_batch_dense(fsk_handle* h, int nb, const BatchSpec& spec, unsigned long long* K, size_t slot_stride) {
    h->ls = h->stream;
    {
        Span sp(h, PC_PACK);
        dim3 grid((unsigned)h->N);
        const size_t smem = (size_t)DENSE_COUNT_WARPS * h->nks * 2;
        constexpr int CT = DENSE_COUNT_WARPS * 32;
        if (h->dense_u8) {
            if (h->NW == 2)
                dense_count_kernel<uint64_t, 2, true><<<grid, CT, smem, h->ls>>>((const uint64_t*)h->d_gw0, h->d_gw1, h->d_woff32, h->nks, h->dense_ld, nb, h->d_C, spec);
            else if (h->gw32)
                dense_count_kernel<uint32_t, 1, true><<<grid, CT, smem, h->ls>>>((const uint32_t*)h->d_gw0, nullptr, h->d_woff32, h->nks, h->dense_ld, nb, h->d_C, spec);
            else
                dense_count_kernel<uint64_t, 1, true><<<grid, CT, smem, h->ls>>>((const uint64_t*)h->d_gw0, nullptr, h->d_woff32, h->nks, h->dense_ld, nb, h->d_C, spec);
        } else if (h->NW == 2)
            dense_count_kernel<uint64_t, 2><<<grid, CT, smem, h->ls>>>((const uint64_t*)h->d_gw0, h->d_gw1, h->d_woff32, h->nks, h->dense_ld, nb, h->d_C, spec);
        else if (h->gw32)
            dense_count_kernel<uint32_t, 1><<<grid, CT, smem, h->ls>>>((const uint32_t*)h->d_gw0, nullptr, h->d_woff32, h->nks, h->dense_ld, nb, h->d_C, spec);
        else
            dense_count_kernel<uint64_t, 1><<<grid, CT, smem, h->ls>>>((const uint64_t*)h->d_gw0, nullptr, h->d_woff32, h->nks, h->dense_ld, nb, h->d_C, spec);
        h->launches++;
        CU(cudaGetLastError());
    }
    {
        Span sp(h, PC_ACCUMULATE);
        const unsigned T = (unsigned)((h->N + DG_TILE - 1) / DG_TILE);
        const unsigned tiles = T * (T + 1) / 2;
        if (h->wf_active) {   // variance mode: every stream's tiles walk the stream's slots in order (Welford in the epilogue)
            if (h->dense_u8) syrk_tc_welford_kernel<true, true><<<dim3(tiles, (unsigned)h->wf_groups), DW_THREADS_REGS, dw_smem(true), h->ls>>>(h->tmap_C, h->d_tile_order, h->N, h->nks, h->d_wf);
            else if (h->wf_regs) syrk_tc_welford_kernel<true><<<dim3(tiles, (unsigned)h->wf_groups), DW_THREADS_REGS, dw_smem(true), h->ls>>>(h->tmap_C, h->d_tile_order, h->N, h->nks, h->d_wf);
            else syrk_tc_welford_kernel<false><<<dim3(tiles, (unsigned)h->wf_groups), DW_THREADS, dw_smem(false), h->ls>>>(h->tmap_C, h->d_tile_order, h->N, h->nks, h->d_wf);
            h->launches++;
        } else {
            for (int c0 = 0; c0 < nb; c0 += h->dense_chunk) {
                const int cs = std::min(h->dense_chunk, nb - c0);
                // one tile per CTA (two CTAs per SM: one's epilogue under the other's MMAs) measured faster than the pair shape at
                // every N from 1 000 to 32 000, byte and fp16 operands alike (profiles/r02_gemm_shape_by_n.txt): the pair is opt-in
                const bool two = h->opt_gemm_shape == 2;
                if (two && h->dense_u8)
                    syrk_tc_kernel<2, true><<<dim3(h->pair_tiles, 1), DG_THREADS, dg_smem(2), h->ls>>>(h->tmap_C, h->d_pair_order, h->N, (uint32_t)c0 * h->nks, 0u, (uint32_t)cs * h->nks, K, 0, nullptr);
                else if (two)
                    syrk_tc_kernel<2><<<dim3(h->pair_tiles, 1), DG_THREADS, dg_smem(2), h->ls>>>(h->tmap_C, h->d_pair_order, h->N, (uint32_t)c0 * h->nks, 0u, (uint32_t)cs * h->nks, K, 0, nullptr);
                else if (h->dense_u8)
                    syrk_tc_kernel<1, true><<<dim3(tiles, 1), DG_THREADS, dg_smem(1), h->ls>>>(h->tmap_C, h->d_tile_order, h->N, (uint32_t)c0 * h->nks, 0u, (uint32_t)cs * h->nks, K, 0, nullptr);
                else
                    syrk_tc_kernel<1><<<dim3(tiles, 1), DG_THREADS, dg_smem(1), h->ls>>>(h->tmap_C, h->d_tile_order, h->N, (uint32_t)c0 * h->nks, 0u, (uint32_t)cs * h->nks, K, 0, nullptr);
                h->launches++;
            }
        }
        CU(cudaGetLastError());
    }
    h->combos_done += nb;
    if (h->spans.size() > 2048) resolve_spans(h);
    return FSK_OK;
}

// One batch of combinations: partial kernels added into K (+ slot * slot_stride).
int run_batch(fsk_handle* h, const int32_t* combos, int nb, unsigned long long* K, size_t slot_stride) {
    BatchSpec spec;
    memset(&spec, 0, sizeof spec);
    int pos[MAX_K];
    for (int s = 0; s < nb; ++s) {
        if (combos[s] < 0 || combos[s] >= h->ncomb) return fail(h, FSK_EINVAL, "combination index %d out of range [0, %lld)", combos[s], (long long)h->ncomb);
        unrank_combination(h->g, h->k, combos[s], pos);
        int nseg = 0;
        for (int j = 0; j < h->k;) {   // maximal stretch of consecutive kept positions inside one g-mer word
            int e = j + 1;
            while (e < h->k && pos[e] == pos[e - 1] + 1 && pos[e] / h->cpw == pos[j] / h->cpw) ++e;
            const int src = (pos[j] / h->cpw) * 64 + (pos[j] % h->cpw) * h->b, width = (e - j) * h->b;
            spec.seg[s][nseg++] = (uint16_t)(src | (width - 1) << 7);
            j = e;
        }
        spec.nseg[s] = (uint8_t)nseg;
    }
    if (h->dense_path) return run_batch_dense(h, nb, spec, K, slot_stride);
    h->heavy_now = 0;
    if (h->heavy_tau) {
        if (h->heavy_probe_pending && cudaEventQuery(h->ev_heavy) == cudaSuccess) {
            h->heavy_probe_pending = false;
            const uint32_t cnt = *h->h_heavy_count;           // candidates of an earlier batch, incl. those the list had no room for
            if (h->opt_heavy_tau == 0) {
                if (cnt == 0) { h->heavy_live = false; h->heavy_idle = 0; }
                // The list is first come, first served: when it overflows, short runs crowd out the long ones that matter most.
                // Run-length statistics barely change between batches (same sequences), so steer the threshold instead.
                else if (cnt > h->heavy_cap) h->heavy_tau = (uint32_t)std::min<int64_t>(h->nfeat, (int64_t)h->heavy_tau * 3 / 2 + 1);
                else if (cnt < h->heavy_cap / 8 && h->heavy_tau > h->heavy_tau_min)
                    h->heavy_tau = std::max(h->heavy_tau_min, h->heavy_tau * 3 / 4);
            }
        }
        if (!h->heavy_live && ++h->heavy_idle >= 32) h->heavy_live = true;
        if (h->heavy_live) h->heavy_now = h->heavy_tau;
    }
    // The pack / sort / segment of this batch go to pre_stream and may overlap the accumulate of the previous batch on
    // the main stream (they are HBM / L2 bound, the accumulate is shared-memory bound); ids/task are double-buffered.
    h->buf = (int)(h->batch_index++ & 1);
    h->ls = h->pre_stream;
    CU(cudaStreamWaitEvent(h->pre_stream, h->ev_acc[h->buf], 0));
    CU(cudaMemsetAsync(h->d_zero, 0, h->zero_bytes, h->ls));
    int rc;
    {
        Span sp(h, PC_PACK);
        if (h->mode == MODE_R32) rc = launch_pack<uint32_t, false>(h, nb, spec);
        else if (h->mode == MODE_R64) rc = launch_pack<uint64_t, false>(h, nb, spec);
        else rc = launch_pack<uint64_t, true>(h, nb, spec);
        if (rc) return rc;
    }
    {
        Span sp(h, PC_SORT);
        if (h->mode == MODE_R32) rc = launch_sort<uint32_t, false, 16>(h, nb);
        else if (h->mode == MODE_R64) rc = launch_sort<uint64_t, false, 16>(h, nb);
        else rc = launch_sort<uint64_t, true, 12>(h, nb);
        if (rc) return rc;
    }
    {
        Span sp(h, PC_SEGMENT);
        if (h->mode == MODE_R32) rc = launch_segment<uint32_t, false>(h, nb);
        else if (h->mode == MODE_R64) rc = launch_segment<uint64_t, false>(h, nb);
        else rc = launch_segment<uint64_t, true>(h, nb);
        if (rc) return rc;
    }
    if (h->heavy_now) {
        // the runs segment_kernel listed become fp16 columns and one tensor-core contraction adds their H H^T to K; the column
        // count lives on the device, so the three launches are unconditional (they return at once when the list is empty)
        Span sp(h, PC_ACCUMULATE);
        const uint32_t* cnt = h->d_ticket + HEAVY_COUNT;
        const unsigned T = (unsigned)((h->N + DG_TILE - 1) / DG_TILE);
        const bool two = h->opt_gemm_shape != 1 && (h->opt_gemm_shape == 2 || T >= 4);
        if (h->heavy_u8) {
            uint8_t* H8 = reinterpret_cast<uint8_t*>(h->d_H);
            heavy_zero_kernel<uint8_t><<<148 * 8, 256, 0, h->ls>>>(H8, (size_t)h->heavy_cap, h->N, cnt);
            if (h->mode == MODE_R32)
                heavy_fill_kernel<uint32_t, uint8_t><<<148 * 4, 256, 0, h->ls>>>((const uint32_t*)h->d_recA, (uint32_t)h->nfeat, h->idbits, h->d_heavy_list, cnt,
                                                                                  H8, (size_t)h->heavy_cap, h->d_counters + 3);
            else
                heavy_fill_kernel<uint64_t, uint8_t><<<148 * 4, 256, 0, h->ls>>>((const uint64_t*)h->d_recA, (uint32_t)h->nfeat, h->idbits, h->d_heavy_list, cnt,
                                                                                  H8, (size_t)h->heavy_cap, h->d_counters + 3);
            if (two)
                syrk_tc_kernel<2, true><<<dim3(h->pair_tiles, 1), DG_THREADS, dg_smem(2), h->ls>>>(h->tmap_H, h->d_pair_order, h->N, 0u, 0u, h->heavy_cap, K, 0, cnt);
            else
                syrk_tc_kernel<1, true><<<dim3(T * (T + 1) / 2, 1), DG_THREADS, dg_smem(1), h->ls>>>(h->tmap_H, h->d_tile_order, h->N, 0u, 0u, h->heavy_cap, K, 0, cnt);
        } else {
            heavy_zero_kernel<__half><<<148 * 8, 256, 0, h->ls>>>(h->d_H, (size_t)h->heavy_cap, h->N, cnt);
            if (h->mode == MODE_R32)
                heavy_fill_kernel<uint32_t, __half><<<148 * 4, 256, 0, h->ls>>>((const uint32_t*)h->d_recA, (uint32_t)h->nfeat, h->idbits, h->d_heavy_list, cnt,
                                                                                 h->d_H, (size_t)h->heavy_cap, h->d_counters + 3);
            else
                heavy_fill_kernel<uint64_t, __half><<<148 * 4, 256, 0, h->ls>>>((const uint64_t*)h->d_recA, (uint32_t)h->nfeat, h->idbits, h->d_heavy_list, cnt,
                                                                                 h->d_H, (size_t)h->heavy_cap, h->d_counters + 3);
            if (two)
                syrk_tc_kernel<2><<<dim3(h->pair_tiles, 1), DG_THREADS, dg_smem(2), h->ls>>>(h->tmap_H, h->d_pair_order, h->N, 0u, 0u, h->heavy_cap, K, 0, cnt);
            else
                syrk_tc_kernel<1><<<dim3(T * (T + 1) / 2, 1), DG_THREADS, dg_smem(1), h->ls>>>(h->tmap_H, h->d_tile_order, h->N, 0u, 0u, h->heavy_cap, K, 0, cnt);
        }
        h->launches += 3;
        CU(cudaGetLastError());
        if (!h->heavy_probe_pending) {   // did this batch have any heavy run?  read back without waiting
            CU(cudaMemcpyAsync(h->h_heavy_count, cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->ls));
            CU(cudaEventRecord(h->ev_heavy, h->ls));
            h->heavy_probe_pending = true;
        }
    }
    CU(cudaEventRecord(h->ev_pre[h->buf], h->pre_stream));
    h->ls = h->stream;
    CU(cudaStreamWaitEvent(h->stream, h->ev_pre[h->buf], 0));
    {
        Span sp(h, PC_ACCUMULATE);
        rc = h->ids16 ? launch_accumulate<uint16_t>(h, nb, K, slot_stride) : launch_accumulate<uint32_t>(h, nb, K, slot_stride);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev_acc[h->buf], h->stream));
    h->combos_done += nb;
    if (h->spans.size() > 2048) resolve_spans(h);
    return FSK_OK;
}

// TMA descriptor of a K-major fp16 operand matrix [N rows][ld columns]: boxes of 128 rows x 64 columns landing with the 128-byte
// swizzle the UMMA descriptors of syrk_tc_kernel expect; rows past N read as zero.  cuTensorMapEncodeTiled is resolved through
// the runtime (no link-time dependency on libcuda).
int encode_operand_map(fsk_handle* h, CUtensorMap* map, __half* ptr, size_t ld, bool u8 = false) {   // u8: byte elements, boxes of 128 x 128
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(h, FSK_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)h->N};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * (u8 ? 1 : 2)};
    const cuuint32_t box[2] = {(cuuint32_t)(u8 ? 2 * DG_BK : DG_BK), (cuuint32_t)DG_TILE};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = ((EncodeFn)fn)(map, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstride, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(h, FSK_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    CU(smem_opt_in(syrk_tc_kernel<1>, (int)dg_smem(1)));
    CU(smem_opt_in(syrk_tc_kernel<2>, (int)dg_smem(2)));
    CU(smem_opt_in(syrk_tc_kernel<1, true>, (int)dg_smem(1)));
    CU(smem_opt_in(syrk_tc_kernel<2, true>, (int)dg_smem(2)));
    CU(smem_opt_in(syrk_tc_welford_kernel<true>, (int)dw_smem(true)));
    CU(smem_opt_in(syrk_tc_welford_kernel<true, true>, (int)dw_smem(true)));
    CU(smem_opt_in(syrk_tc_welford_kernel<false>, (int)dw_smem(false)));
    return FSK_OK;
}

// launch order of the output tiles of syrk_tc_kernel: bands of 16 tile rows, column by column inside a band
int make_tile_order(fsk_handle* h) {
    const int64_t T = (h->N + DG_TILE - 1) / DG_TILE;
    if (T > 65535) return fail(h, FSK_EINVAL, "too many sequences for the tensor-core contraction");
    std::vector<uint32_t> order;
    order.reserve((size_t)(T * (T + 1) / 2));
    for (int64_t b0 = 0; b0 < T; b0 += 16) {
        const int64_t b1 = std::min<int64_t>(T, b0 + 16);
        for (int64_t J = 0; J < b1; ++J)
            for (int64_t I = std::max(b0, J); I < b1; ++I) order.push_back((uint32_t)(I << 16 | J));
    }
    ALLOC(h->d_tile_order, order.size());
    CU(cudaMemcpy(h->d_tile_order, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    // the same for the two-tile shape: entries (P << 16 | J) name the tile rows 2 P and 2 P + 1, J <= 2 P + 1; bands of 8 pairs
    const int64_t TP = (T + 1) / 2;
    order.clear();
    for (int64_t b0 = 0; b0 < TP; b0 += 8) {
        const int64_t b1 = std::min<int64_t>(TP, b0 + 8);
        for (int64_t J = 0; J < std::min<int64_t>(T, 2 * b1); ++J)
            for (int64_t P = std::max(b0, J / 2); P < b1; ++P) order.push_back((uint32_t)(P << 16 | J));
    }
    h->pair_tiles = (uint32_t)order.size();
    ALLOC(h->d_pair_order, order.size());
    CU(cudaMemcpy(h->d_pair_order, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return FSK_OK;
}

int build_queue(fsk_handle* h) {
    if (!h->user_queue.empty()) {
        for (int32_t c : h->user_queue)
            if (c < 0 || c >= h->ncomb) return fail(h, FSK_EINVAL, "combination %d out of range [0, %lld)", c, (long long)h->ncomb);
        h->queue = h->user_queue;
        return FSK_OK;
    }
    // fastsk_kernel.cpp:31-47: identity, then std::shuffle with std::default_random_engine
    h->queue.resize((size_t)h->ncomb);
    for (int64_t i = 0; i < h->ncomb; ++i) h->queue[(size_t)i] = (int32_t)i;
    auto rng = std::default_random_engine{};
    const uint64_t seed = h->have_seed ? h->seed : (h->run_seed_set ? h->run_seed : (uint64_t)std::time(0));
    rng.seed((std::default_random_engine::result_type)seed);
    std::shuffle(h->queue.begin(), h->queue.end(), rng);
    return FSK_OK;
}

int effective_streams(const fsk_handle* h, int64_t nq) {   // fastsk_kernel.cpp:54-61
    int T = h->t == -1 ? 20 : h->t;
    if (T < 1) T = 1;
    if (T > nq) T = (int)nq;
    return T;
}

// What this shard processes.  Integer modes (exact, approx && skip_variance): the union of what every
// reference thread would process -- stream tid runs queue[tid + T*r] until max_iters or the end of the
// queue (fastsk_kernel.cpp:257-262, 275-278) -- dealt round-robin to the ranks; K is an integer sum, so
// any partition gives the same result.  Variance mode: the ids of the virtual streams this rank owns.
void shard_work(const fsk_handle* h, std::vector<int32_t>& out) {
    out.clear();
    const int64_t nq = (int64_t)h->queue.size();
    const int T = effective_streams(h, nq);
    if (h->approx && !h->skip_variance) {
        for (int tid = h->rank; tid < T; tid += h->world) out.push_back(tid);
        return;
    }
    std::vector<int32_t> work;
    if (!h->approx) {
        work = h->queue;
    } else {
        for (int tid = 0; tid < T; ++tid) {
            int64_t avail = (nq - tid + T - 1) / T;
            if (h->max_iters != -1) avail = std::min<int64_t>(avail, std::max(1, h->max_iters));
            for (int64_t r = 0; r < avail; ++r) work.push_back(h->queue[(size_t)(tid + T * r)]);
        }
    }
    for (size_t i = (size_t)h->rank; i < work.size(); i += (size_t)h->world) out.push_back(work[i]);
}

// the partial kernels the normalisation sums: this handle's own buffer, or those of every rank (sharded finalisation)
template <typename T>
PeerParts<T> make_parts(const fsk_handle* h, const T* own) {
    PeerParts<T> parts;
    memset(&parts, 0, sizeof parts);
    if (h->peer_parts.empty()) {
        parts.p[0] = own;
        parts.n = 1;
    } else {
        parts.n = (int)h->peer_parts.size();
        for (int r = 0; r < parts.n; ++r) parts.p[r] = (const T*)h->peer_parts[(size_t)r];
    }
    return parts;
}

// normalised train rows [tr_r0, +tr_nr) and test rows [te_r0, +te_nr) of the summed partial kernels (fastsk_kernel.cpp:96-103
// fused with the merge of the partials, :285-315); one rank alone holds all rows
template <typename T>
int finalize_typed(fsk_handle* h, const T* K) {
    h->ls = h->stream;
    Span sp(h, PC_NORMALISE);
    const PeerParts<T> parts = make_parts<T>(h, K);
    diag_kernel<T><<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(parts, h->N, h->d_diag);
    h->launches++;
    const unsigned tcols = (unsigned)((h->n_train + 31) / 32);
    if (h->tr_nr > 0) {
        const unsigned trows = (unsigned)((h->tr_nr + 31) / 32);
        normalise_block_kernel<T><<<dim3(tcols, trows), 256, 0, h->stream>>>(parts, h->d_diag, h->tr_r0, h->tr_nr, h->n_train, h->d_train, 0);
        normalise_block_kernel<T><<<dim3(trows, tcols), 256, 0, h->stream>>>(parts, h->d_diag, h->tr_r0, h->tr_nr, h->n_train, h->d_train, 1);
        h->launches += 2;
    }
    if (h->te_nr > 0) {       // (test rows lie below every train column: no mirrored tiles)
        const unsigned trows = (unsigned)((h->te_nr + 31) / 32);
        normalise_block_kernel<T><<<dim3(tcols, trows), 256, 0, h->stream>>>(parts, h->d_diag, h->n_train + h->te_r0, h->te_nr, h->n_train, h->d_test, 0);
        h->launches++;
    }
    CU(cudaGetLastError());
    return FSK_OK;
}

// Did segment_kernel see a batch that the sort left out of order?  (Only possible if the optimistic ranking of
// onesweep_kernel was not served in lane order.)  Synchronises the stream.
int sort_was_unstable(fsk_handle* h, bool* bad) {
    uint32_t f = 0;
    CU(cudaMemcpyAsync(&f, h->d_flag, sizeof f, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *bad = f != 0;
    if (h->dense_path || h->heavy_tau) {     // the tensor-core kernels' barrier waits are bounded: did one run out?
        unsigned int fault = 0;
        CU(cudaMemcpyFromSymbol(&fault, fsk_dev_fault, sizeof fault));
        if (fault) {
            const unsigned int zero = 0;
            cudaMemcpyToSymbol(fsk_dev_fault, &zero, sizeof zero);
            return fail(h, FSK_ECUDA, "a tensor-core kernel gave up waiting on a barrier (TMA / MMA pipeline protocol error); the result is invalid");
        }
    }
    return FSK_OK;
}

int build_partial_once(fsk_handle* h);
int upload_one(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test,
               const int32_t* codes_test = nullptr, const int64_t* offsets_test = nullptr);
int reset_one(fsk_handle* h);
int accumulate_one(fsk_handle* h, const int32_t* combos, int64_t n, int sync);
int build_one(fsk_handle* h);
int finalize_one(fsk_handle* h);
int sync_one(fsk_handle* h);
void* own_part(const fsk_handle* h) { return h->variance_mode ? (void*)h->d_Kf : (void*)h->d_Kint; }
// Peer mappings opened with cudaIpcOpenMemHandle are kept for the life of the process (until fsk_trim_cache), keyed by the
// 64 handle bytes: mapping a 10 GB partial kernel costs ~0.1 s, the ranks' partial buffers are the same cached blocks from
// one compute to the next, and seven such opens per compute were most of an 8-rank job's fixed cost.
struct IpcCache {
    std::mutex mu;
    std::map<std::pair<int, std::string>, void*> open;     // (device, handle bytes) -> mapping
};
IpcCache g_ipc;
void release_peers(fsk_handle* h) {
    h->ipc_opened.clear();
    h->peer_parts.clear();
}

// run fn(member) for every member of the team, each from its own host thread (one thread per GPU, like the reference's one
// std::thread per stream, fastsk_kernel.cpp:86-93); the first failure's message is copied to the leader
int team_run(fsk_handle* h, const std::function<int(fsk_handle*)>& fn) {
    const size_t n = h->team.size();
    std::vector<int> rcs(n, FSK_OK);
    std::vector<std::thread> th;
    for (size_t i = 1; i < n; ++i) th.emplace_back([&, i] { rcs[i] = fn(h->team[i]); });
    rcs[0] = fn(h);
    for (auto& t : th) t.join();
    cudaSetDevice(h->device);
    for (size_t i = 0; i < n; ++i)
        if (rcs[i]) {
            if (i) h->err = "device " + std::to_string(h->team[i]->device) + ": " + h->team[i]->err;
            return rcs[i];
        }
    return FSK_OK;
}
bool is_team(const fsk_handle* h) { return h->team.size() > 1; }

void destroy_one(fsk_handle* h) {
    cudaSetDevice(h->device);
    release_device(h);
    if (h->stream) {
        if (h->pre_stream != h->stream) cudaStreamDestroy(h->pre_stream);
        cudaStreamDestroy(h->stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(h->ev_pre[i]); cudaEventDestroy(h->ev_acc[i]); }
        cudaEventDestroy(h->ev_sync);
    }
    if (h->ev_heavy) cudaEventDestroy(h->ev_heavy);
    if (h->h_heavy_count) cudaFreeHost(h->h_heavy_count);
    delete h;
}

// (re)create the team members 1 .. n-1 as copies of the leader's configuration
void sync_team(fsk_handle* h) {
    if (h->weights_auto) { h->out_weights.clear(); h->weights_auto = false; }
    if (h->devices.size() < 2) {
        if (h->team.size() > 1) { h->rank = 0; h->world = 1; }      // (the leader was member 0 of n)
        for (size_t i = 1; i < h->team.size(); ++i) destroy_one(h->team[i]);
        h->team.clear();
        return;
    }
    if (h->team.empty()) h->team.push_back(h);
    while (h->team.size() > h->devices.size()) { destroy_one(h->team.back()); h->team.pop_back(); }
    while (h->team.size() < h->devices.size()) { h->team.push_back(new fsk_handle()); h->team.back()->leader = h; }
    const int n = (int)h->devices.size();
    for (int i = 0; i < n; ++i) {
        fsk_handle* w = h->team[(size_t)i];
        if (i > 0 && w->stream && w->device != h->devices[(size_t)i]) {   // the member moves to another GPU: start it afresh there
            destroy_one(w);
            h->team[(size_t)i] = w = new fsk_handle();
            w->leader = h;
        }
        w->device = h->devices[(size_t)i];
        w->rank = i; w->world = n;
        if (i == 0) continue;
        w->g = h->g; w->m = h->m; w->k = h->k; w->t = h->t; w->approx = h->approx; w->skip_variance = h->skip_variance;
        w->delta = h->delta; w->max_iters = h->max_iters; w->ncomb = h->ncomb;
        w->have_seed = h->have_seed; w->seed = h->seed; w->user_queue = h->user_queue;
        w->opt_batch = h->opt_batch; w->opt_acc_path = h->opt_acc_path; w->safe_rank = h->safe_rank;
        w->opt_rows_threads = h->opt_rows_threads; w->opt_overlap = h->opt_overlap; w->opt_seg_fused = h->opt_seg_fused;
        w->opt_acc_prefetch = h->opt_acc_prefetch; w->opt_acc_unroll = h->opt_acc_unroll; w->opt_wave = h->opt_wave;
        w->profile = h->profile; w->opt_pad = h->opt_pad; w->opt_acc_cols = h->opt_acc_cols; w->opt_heavy_tau = h->opt_heavy_tau;
        w->opt_ids32 = h->opt_ids32; w->opt_gemm_shape = h->opt_gemm_shape; w->opt_heavy_cap = h->opt_heavy_cap;
        w->out_weights = h->out_weights;
        w->opt_seg_lean = h->opt_seg_lean; w->opt_spec_depth = h->opt_spec_depth; w->opt_pf_stride = h->opt_pf_stride; w->opt_fit_smem = h->opt_fit_smem; w->opt_wf_regs = h->opt_wf_regs; w->opt_dense_u8 = h->opt_dense_u8;
        w->opt_seg_dir = h->opt_seg_dir; w->opt_dir_blocks = h->opt_dir_blocks; w->opt_count_updates = h->opt_count_updates;
    }
}

}  // namespace

// ================================================================================================
extern "C" {

const char* fsk_version(void) { return "fastsk_b200 0.1 (sm_100a)"; }

const char* fsk_last_error(const fsk_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int fsk_create(fsk_handle** out, int g, int m, int t, int approx, double delta, int max_iters, int skip_variance) {
    if (!out) return fail(nullptr, FSK_EINVAL, "out is NULL");
    *out = nullptr;
    // the reference's validate_args (shared.cpp:380-391; never called there) asks g > m and g <= 20
    if (g < 1 || m < 0 || g <= m) return fail(nullptr, FSK_EINVAL, "g must be greater than m (g = %d, m = %d)", g, m);
    if (g - m > MAX_K || g > 64) return fail(nullptr, FSK_EINVAL, "g - m must be at most %d and g at most 64 (g = %d, m = %d)", MAX_K, g, m);
    if (t == 0 || t < -1) return fail(nullptr, FSK_EINVAL, "t must be -1 or positive (t = %d)", t);
    if (nchoosek64(g, m) > 0x7fffffffLL) return fail(nullptr, FSK_EINVAL, "C(g, m) does not fit a 32-bit combination index");
    fsk_handle* h = new fsk_handle();
    h->g = g; h->m = m; h->k = g - m; h->t = t;
    h->approx = approx != 0; h->delta = delta; h->max_iters = max_iters; h->skip_variance = skip_variance != 0;
    h->ncomb = nchoosek64(g, m);
    *out = h;
    return FSK_OK;
}

void fsk_destroy(fsk_handle* h) {
    if (!h) return;
    for (size_t i = 1; i < h->team.size(); ++i) destroy_one(h->team[i]);
    h->team.clear();
    destroy_one(h);
}

int fsk_set_device(fsk_handle* h, int device) {
    if (device == h->device && h->devices.size() < 2) return FSK_OK;
    // moving to another GPU drops what the handle holds on the old one (a second compute on the same object is allowed)
    if (h->stream) { cudaSetDevice(h->device); release_device(h); }
    if (h->stream) {
        if (h->pre_stream != h->stream) cudaStreamDestroy(h->pre_stream);
        cudaStreamDestroy(h->stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(h->ev_pre[i]); cudaEventDestroy(h->ev_acc[i]); }
        cudaEventDestroy(h->ev_sync);
        h->stream = h->pre_stream = h->ls = nullptr;
    }
    if (h->ev_heavy) { cudaEventDestroy(h->ev_heavy); h->ev_heavy = nullptr; }
    if (h->h_heavy_count) { cudaFreeHost(h->h_heavy_count); h->h_heavy_count = nullptr; }
    h->device = device;
    h->devices.clear();
    sync_team(h);
    return FSK_OK;
}

// In-process multi-GPU: the reference fans one compute_kernel call out over T threads (fastsk_kernel.cpp:54-94); here one
// call fans out over the listed GPUs, one host thread each.  devices == NULL && n == -1 means every visible GPU.
int fsk_set_devices(fsk_handle* h, const int* devices, int n) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) { cudaGetLastError(); visible = 0; }
    std::vector<int> dev;
    if (n == -1 && !devices) for (int i = 0; i < visible; ++i) dev.push_back(i);
    else {
        if (n < 1 || !devices) return fail(h, FSK_EINVAL, "fsk_set_devices needs at least one device");
        dev.assign(devices, devices + n);
    }
    if ((int)dev.size() > MAX_PEERS) return fail(h, FSK_EINVAL, "at most %d devices", MAX_PEERS);
    for (size_t i = 0; i < dev.size(); ++i) {
        if (dev[i] < 0 || (visible && dev[i] >= visible)) return fail(h, FSK_EINVAL, "device %d is not visible (%d devices)", dev[i], visible);
        for (size_t j = 0; j < i; ++j) if (dev[j] == dev[i]) return fail(h, FSK_EINVAL, "device %d listed twice", dev[i]);
    }
    if (dev.empty()) return fail(h, FSK_ECUDA, "no CUDA device is visible");
    if (h->world > 1 && !is_team(h)) return fail(h, FSK_ESTATE, "fsk_set_devices and fsk_set_shard are alternatives (one process per GPU, or one process for all)");
    int rc = fsk_set_device(h, dev[0]);
    if (rc) return rc;
    if (dev.size() > 1) {
        h->devices = dev;
        sync_team(h);
    } else {
        h->rank = 0; h->world = 1;
    }
    return FSK_OK;
}
int fsk_set_seed(fsk_handle* h, uint64_t seed) { h->have_seed = true; h->seed = seed; return FSK_OK; }
int fsk_set_combo_sequence(fsk_handle* h, const int32_t* combos, int64_t n) {
    if (n < 0 || (n > 0 && !combos)) return fail(h, FSK_EINVAL, "bad combination sequence");
    h->user_queue.assign(combos, combos + n);
    return FSK_OK;
}
int fsk_set_shard(fsk_handle* h, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) return fail(h, FSK_EINVAL, "bad shard %d of %d", rank, world);
    if (world > MAX_PEERS) return fail(h, FSK_EINVAL, "at most %d ranks", MAX_PEERS);
    if (is_team(h) && world > 1) return fail(h, FSK_ESTATE, "fsk_set_devices and fsk_set_shard are alternatives");
    h->rank = rank; h->world = world;
    return FSK_OK;
}
int fsk_set_option(fsk_handle* h, const char* key, int64_t value) {
    if (!key) return fail(h, FSK_EINVAL, "key is NULL");
    if (!strcmp(key, "batch")) {
        if (value < 0 || value > MAX_BATCH) return fail(h, FSK_EINVAL, "batch must be in [0, %d]", MAX_BATCH);
        h->opt_batch = (int)value;
    } else if (!strcmp(key, "acc_path")) {
        if (value < 0 || value > 3) return fail(h, FSK_EINVAL, "acc_path must be 0 (auto), 1 (global RED), 2 (shared-memory rows) or 3 (dense tensor-core contraction)");
        h->opt_acc_path = (int)value;
    } else if (!strcmp(key, "safe_rank")) {
        h->safe_rank = value != 0;
    } else if (!strcmp(key, "rows_threads")) {
        if (value != 0 && (value < 32 || value > 1024 || value % 32)) return fail(h, FSK_EINVAL, "rows_threads must be 0 or a multiple of 32 up to 1024");
        h->opt_rows_threads = (int)value;
    } else if (!strcmp(key, "overlap")) {
        if (h->stream) return fail(h, FSK_ESTATE, "overlap must be set before the first upload");
        h->opt_overlap = value != 0;
    } else if (!strcmp(key, "wave")) {
        if (value < 1 || value > 1024) return fail(h, FSK_EINVAL, "wave must be in [1, 1024]");
        h->opt_wave = (int)value;
    } else if (!strcmp(key, "pad")) {
        if (value < 0 || value > 2) return fail(h, FSK_EINVAL, "pad must be 0 (auto), 1 (16-byte units) or 2 (128-byte lines)");
        h->opt_pad = (int)value;
    } else if (!strcmp(key, "heavy_tau")) {
        if (value < -1) return fail(h, FSK_EINVAL, "heavy_tau must be -1 (off), 0 (auto) or a positive run length");
        h->opt_heavy_tau = (int)value;
    } else if (!strcmp(key, "ids32")) {
        h->opt_ids32 = value != 0;
    } else if (!strcmp(key, "gemm_shape")) {
        if (value < 0 || value > 2) return fail(h, FSK_EINVAL, "gemm_shape must be 0 (auto), 1 or 2 tiles per CTA");
        h->opt_gemm_shape = (int)value;
    } else if (!strcmp(key, "heavy_cap")) {
        if (value != 0 && (value < 64 || value % 64 || value > 65536)) return fail(h, FSK_EINVAL, "heavy_cap must be 0 (auto) or a multiple of 64 up to 65536");
        h->opt_heavy_cap = (int)value;
    } else if (!strcmp(key, "acc_cols")) {
        if (value != 0 && (value < 32 || value % 32)) return fail(h, FSK_EINVAL, "acc_cols must be 0 (auto) or a positive multiple of 32");
        h->opt_acc_cols = (int)value;
    } else if (!strcmp(key, "seg_fused")) {
        if (value < 0 || value > 2) return fail(h, FSK_EINVAL, "seg_fused must be 0 (auto), 1 (off) or 2 (on)");
        h->opt_seg_fused = (int)value;
    } else if (!strcmp(key, "seg_dir")) {
        if (value < 0 || value > 2) return fail(h, FSK_EINVAL, "seg_dir must be 0 (auto), 1 (off) or 2 (on)");
        h->opt_seg_dir = (int)value;
    } else if (!strcmp(key, "fit_smem")) {
        h->opt_fit_smem = value != 0;
    } else if (!strcmp(key, "wf_regs")) {
        h->opt_wf_regs = value != 0;
    } else if (!strcmp(key, "dense_u8")) {
        h->opt_dense_u8 = value != 0;
    } else if (!strcmp(key, "pf_stride")) {
        if (value != 64 && value != 128) return fail(h, FSK_EINVAL, "pf_stride must be 64 or 128");
        h->opt_pf_stride = (int)value;
    } else if (!strcmp(key, "spec_depth")) {
        if (value < 0 || value > MAX_BATCH) return fail(h, FSK_EINVAL, "spec_depth must be in [0, %d]", MAX_BATCH);
        h->opt_spec_depth = (int)value;
    } else if (!strcmp(key, "seg_lean")) {
        if (value < 0 || value > 2) return fail(h, FSK_EINVAL, "seg_lean must be 0 (auto), 1 (off) or 2 (on)");
        h->opt_seg_lean = (int)value;
    } else if (!strcmp(key, "dir_blocks")) {
        if (value < 1 || value > 4096) return fail(h, FSK_EINVAL, "dir_blocks must be in [1, 4096]");
        h->opt_dir_blocks = (int)value;
    } else if (!strcmp(key, "count_updates")) {
        h->opt_count_updates = value != 0;
    } else if (!strcmp(key, "acc_prefetch")) {
        h->opt_acc_prefetch = value != 0;
    } else if (!strcmp(key, "acc_unroll")) {
        if (value != 2 && value != 4) return fail(h, FSK_EINVAL, "acc_unroll must be 2 or 4");
        h->opt_acc_unroll = (int)value;
    } else if (!strcmp(key, "profile")) {
        h->profile = value != 0;
    } else {
        return fail(h, FSK_EINVAL, "unknown option '%s'", key);
    }
    return FSK_OK;
}

}  // extern "C"

namespace {
// (codes, offsets) hold all N sequences back to back -- or, when codes_test is given, the train sequences only, and the test
// sequences come from (codes_test, offsets_test): the two halves of compute_kernel(Xtrain, Xtest) need no concatenation
int upload_one(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test,
               const int32_t* codes_test, const int64_t* offsets_test) {
    if (!codes || !offsets) return fail(h, FSK_EINVAL, "codes/offsets is NULL");
    if ((codes_test == nullptr) != (offsets_test == nullptr)) return fail(h, FSK_EINVAL, "codes_test and offsets_test go together");
    const bool split = codes_test != nullptr;
    Trace tr("upload");
    // the reference dereferences Xtrain[0] / Xtest[0] unconditionally (fastsk.cpp:33,41); compute_train passes no test set
    if (n_train < 1 || n_test < 0) return fail(h, FSK_EINVAL, "need at least one train sequence (n_train = %lld, n_test = %lld)", (long long)n_train, (long long)n_test);
    const int64_t N = n_train + n_test;
    if (N >= (1LL << 31)) return fail(h, FSK_EINVAL, "too many sequences");
    // fastsk.cpp:53-58: g longer than the shortest sequence is fatal (exit(1) there, an error code here)
    int64_t shortest_train = INT64_MAX, shortest_test = INT64_MAX, nfeat = 0, maxwin = 0;
    std::vector<int64_t> off0((size_t)N + 1);      // offsets of the sequences in the (virtual) concatenation, from 0
    off0[0] = 0;
    for (int64_t i = 0; i < N; ++i) {
        const int64_t len = (split && i >= n_train) ? offsets_test[i - n_train + 1] - offsets_test[i - n_train] : offsets[i + 1] - offsets[i];
        if (len < 0) return fail(h, FSK_EINVAL, "offsets must be non-decreasing");
        off0[(size_t)i + 1] = off0[(size_t)i] + len;
        if (i < n_train) shortest_train = std::min(shortest_train, len);
        else shortest_test = std::min(shortest_test, len);
        nfeat += len - h->g + 1;
        maxwin = std::max(maxwin, len - h->g + 1);
    }
    if (h->g > shortest_train)
        return fail(h, FSK_EINVAL, "g cannot be longer than the shortest sequence in a dataset: g = %d, but shortest train sequence has length %lld", h->g, (long long)shortest_train);
    if (n_test > 0 && h->g > shortest_test)
        return fail(h, FSK_EINVAL, "g cannot be longer than the shortest sequence in a dataset: g = %d, but shortest test sequence has length %lld", h->g, (long long)shortest_test);
    if (nfeat >= (1LL << 30)) return fail(h, FSK_EINVAL, "too many g-mers (%lld); at most 2^30 - 1", (long long)nfeat);

    // dense re-coding: only equality of characters matters.  Host threads over slices of the characters (the
    // 10 M characters of configs[3] took 20 ms on one core: 2 % of an 8-GPU build).
    const int64_t total = off0[(size_t)N];
    const int64_t total_a = split ? off0[(size_t)n_train] : total;          // characters of the first source array
    const int32_t* __restrict__ src_a = codes + offsets[0];
    const int32_t* __restrict__ src_b = split ? codes_test + offsets_test[0] : nullptr;
    const int nthr = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)8, (int64_t)std::thread::hardware_concurrency(), total >> 18}));
    // fn(thread, source pointer, first position in the concatenation, count) over slices of the characters
    auto parallel = [&](const std::function<void(int, const int32_t*, int64_t, int64_t)>& fn) {
        auto piece = [&](int t, int64_t a, int64_t e) {
            if (a < total_a) fn(t, src_a + a, a, std::min(e, total_a) - a);
            if (e > total_a) { const int64_t a2 = std::max(a, total_a); fn(t, src_b + (a2 - total_a), a2, e - a2); }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; ++t) th.emplace_back(piece, t, total * t / nthr, total * (t + 1) / nthr);
        piece(0, 0, total / nthr);
        for (auto& x : th) x.join();
    };
    std::vector<int32_t> tmax((size_t)nthr, 0), tmin((size_t)nthr, 0);
    parallel([&](int t, const int32_t* c, int64_t, int64_t cnt) {
        int32_t mx = tmax[(size_t)t], mn = tmin[(size_t)t];
        for (int64_t i = 0; i < cnt; ++i) { mx = std::max(mx, c[i]); mn = std::min(mn, c[i]); }
        tmax[(size_t)t] = mx; tmin[(size_t)t] = mn;
    });
    const int32_t maxv = *std::max_element(tmax.begin(), tmax.end());
    if (*std::min_element(tmin.begin(), tmin.end()) < 0)
        return fail(h, FSK_EINVAL, "negative character code in the input");
    std::vector<int32_t> remap;
    std::vector<uint8_t> dense((size_t)total);
    int A = 0;
    if (maxv < (1 << 22)) {
        std::vector<std::vector<uint8_t>> seen((size_t)nthr, std::vector<uint8_t>((size_t)maxv + 1, 0));
        parallel([&](int t, const int32_t* c, int64_t, int64_t cnt) {
            uint8_t* sn = seen[(size_t)t].data();
            for (int64_t i = 0; i < cnt; ++i) sn[c[i]] = 1;
        });
        remap.assign((size_t)maxv + 1, -1);
        for (int32_t v = 0; v <= maxv; ++v) {
            bool any = false;
            for (int t = 0; t < nthr; ++t) any |= seen[(size_t)t][(size_t)v] != 0;
            if (any) remap[(size_t)v] = A++;
        }
        if (A > 256) return fail(h, FSK_EINVAL, "alphabet of %d distinct characters; at most 256 are supported", A);
        parallel([&](int, const int32_t* c, int64_t at, int64_t cnt) {
            for (int64_t i = 0; i < cnt; ++i) dense[(size_t)(at + i)] = (uint8_t)remap[(size_t)c[i]];
        });
    } else {
        std::vector<int32_t> uniq(src_a, src_a + total_a);
        if (split) uniq.insert(uniq.end(), src_b, src_b + (total - total_a));
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        A = (int)uniq.size();
        if (A > 256) return fail(h, FSK_EINVAL, "alphabet of %d distinct characters; at most 256 are supported", A);
        parallel([&](int, const int32_t* c, int64_t at, int64_t cnt) {
            for (int64_t i = 0; i < cnt; ++i) dense[(size_t)(at + i)] = (uint8_t)(std::lower_bound(uniq.begin(), uniq.end(), c[i]) - uniq.begin());
        });
    }
    tr.lap("length scan + dense re-coding");
    const int b = ceil_log2(A);
    const int cpw = 64 / b;
    if (h->k * b > 64) return fail(h, FSK_EINVAL, "(g - m) * bits-per-character = %d * %d exceeds the 64-bit key", h->k, b);
    if (h->g > 2 * cpw) return fail(h, FSK_EINVAL, "g * bits-per-character = %d * %d exceeds the 128-bit g-mer word", h->g, b);

    CU(cudaSetDevice(h->device));
    int cc_major = 0, cc_minor = 0;      // (cudaGetDeviceProperties takes ~7 ms: more than a whole small build)
    CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, h->device));
    CU(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, h->device));
    if (cc_major < 10) return fail(h, FSK_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", h->device, cc_major, cc_minor);
    if (!h->stream) {
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));        // accumulate CTAs are placed first
        if (h->opt_overlap) CU(cudaStreamCreateWithPriority(&h->pre_stream, cudaStreamNonBlocking, prio_lo));
        else h->pre_stream = h->stream;
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&h->ev_pre[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_acc[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming));
        h->ls = h->stream;
    }
    tr.lap("device properties, streams");
    release_device(h);
    tr.lap("release of the previous buffers");

    h->n_train = n_train; h->n_test = n_test; h->N = N; h->nfeat = nfeat;
    h->n_pairs = N * (N + 1) / 2;
    h->n_train_pairs = n_train * (n_train + 1) / 2;
    h->A = A; h->b = b; h->cpw = cpw;
    h->NW = h->g > cpw ? 2 : 1;
    h->gw32 = h->NW == 1 && h->g * b <= 32;
    h->keybits = h->k * b;
    h->idbits = ceil_log2(N);
    if (h->keybits + h->idbits <= 32) { h->mode = MODE_R32; h->rec_bytes = 4; h->sort_items = 16; }
    else if (h->keybits + h->idbits <= 64) { h->mode = MODE_R64; h->rec_bytes = 8; h->sort_items = 16; }
    else { h->mode = MODE_KV; h->rec_bytes = 12; h->sort_items = 12; }
    h->plan.npass = (h->keybits + 7) / 8;
    {
        const int base = h->keybits / h->plan.npass, rem = h->keybits % h->plan.npass;
        int sh = 0;
        for (int p = 0; p < h->plan.npass; ++p) {
            h->plan.bits[p] = (uint8_t)(base + (p < rem ? 1 : 0));
            h->plan.shift[p] = (uint8_t)sh;
            sh += h->plan.bits[p];
        }
    }
    h->variance_mode = h->approx && !h->skip_variance;

    h->maxwin = maxwin;
    // accumulate path: rows of K in shared memory (4 B per column, one CTA per row) whenever a row fits;
    // otherwise global RED on the packed triangle
    int max_smem = 0, n_sm = 148;
    CU(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
    // a row of K that does not fit is split into column windows, each a pass of its own over the row's tasks (up to 16)
    {
        const int64_t fit = (((int64_t)max_smem - 1024 - 128) / 4) & ~31LL;           // columns one CTA can hold
        const int64_t windows = h->opt_acc_cols ? (N + h->opt_acc_cols - 1) / h->opt_acc_cols : (N + fit - 1) / fit;
        h->col_windows = (int)std::max<int64_t>(1, windows);
        h->col_width = h->opt_acc_cols ? h->opt_acc_cols : std::min<int64_t>(fit, ((N + windows - 1) / windows + 31) & ~31LL);
        if (h->opt_acc_cols && h->opt_acc_cols > fit) return fail(h, FSK_EINVAL, "acc_cols = %d exceeds the %lld columns a CTA can hold", h->opt_acc_cols, (long long)fit);
    }
    const bool rows_ok = h->col_windows <= 16 && (double)maxwin * (double)maxwin < 4294967296.0;
    if (h->opt_acc_path == 2 && !rows_ok) return fail(h, FSK_EINVAL, "acc_path = 2 needs at most 16 column windows of K (N = %lld) and fewer than 65536 windows per sequence", (long long)N);
    h->rows_path = h->opt_acc_path == 2 || (h->opt_acc_path == 0 && rows_ok);
    {
        // dense regime: few distinct k-mers per combination, so K += C C^T on the tensor cores beats the sort + sparse
        // update.  Cost model per combination: N^2 / 2 x nks MACs at an effective 5e14 MAC/s against
        // nfeat^2 / (2 keys) shared-memory updates at 1.2e12 /s plus 2.7e-11 s per window of sort + segmentation.
        const bool dense_ok = h->keybits <= 12 && maxwin <= 2048;
        const uint32_t nks = dense_ok ? (uint32_t)(((1u << h->keybits) + DG_BK - 1) / DG_BK * DG_BK) : 0;
        double keys = 1;
        for (int i = 0; i < h->k; ++i) keys *= A;
        // (byte operands: 2.5e15 int8 op/s measured on the contraction = 1.2e15 MAC/s; fp16: 1.25e15 flop/s = 6e14 MAC/s)
        const bool u8_ok = h->opt_dense_u8 && maxwin <= 255 && nks % (2 * DG_BK) == 0;
        const double t_dense = 0.5 * (double)N * (double)N * nks / (u8_ok ? 1.0e15 : 5e14) + (double)N * nks * 2 / 3e12 + 5e-6;
        const double t_sparse = (double)nfeat * (double)nfeat / (2.0 * keys) / 1.2e12 + (double)nfeat * 2.7e-11 + 5e-6;
        if (h->opt_acc_path == 3 && !dense_ok)
            return fail(h, FSK_EINVAL, "acc_path = 3 needs at most 12 key bits and 2048 windows per sequence (key bits = %d, windows = %lld)", h->keybits, (long long)maxwin);
        h->dense_path = h->opt_acc_path == 3 || (h->opt_acc_path == 0 && dense_ok && t_dense < t_sparse);
        h->nks = nks;
        if (h->dense_path) h->rows_path = false;
    }
    h->rows_smem = (size_t)std::min<int64_t>(N, h->col_width) * 4 + 128;   // + one dump word per lane for masked-off ids
    h->rows_threads = N >= 16384 ? 1024 : (N >= 4096 ? 512 : 256);
    if (h->opt_rows_threads) h->rows_threads = h->opt_rows_threads;
    h->ids16 = N <= 65000 && !h->opt_ids32;   // u16 ids leave room for the 32 dump words b + 1 + lane (packed 16-bit min in the accumulate)
    {
        // runs of the id stream are aligned (segment_kernel): to one 16-byte unit always, to a 128-byte line when the
        // padding costs at most half of the stream again (few, long runs -- the HBM fetches whole lines either way)
        const int64_t unit = h->ids16 ? 8 : 4, line = unit * 8;
        const int64_t max_runs = h->keybits >= 31 ? nfeat : std::min<int64_t>(nfeat, (int64_t)1 << h->keybits);
        const bool line_ok = (line - 1) * max_runs <= nfeat / 2;
        const int64_t align = h->opt_pad == 1 ? unit : (h->opt_pad == 2 ? line : (line_ok ? line : unit));
        h->pad_mask = (uint32_t)(align - 1);
        const int64_t cap = nfeat + (align - 1) * max_runs + 64;
        if (cap >= (1LL << 30)) return fail(h, FSK_EINVAL, "too many g-mers (%lld) for the aligned id stream", (long long)nfeat);
        h->ids_stride = (size_t)((cap + 63) / 64 * 64);
        // fused last pass + segmentation: two radix digits, 32-bit records, 16-bit ids, row-stationary accumulate
        const bool fused_ok = h->plan.npass == 2 && h->mode == MODE_R32 && h->ids16 && h->rows_path;
        if (h->opt_seg_fused == 2 && !fused_ok)
            return fail(h, FSK_EINVAL, "seg_fused = 2 needs keys of two radix digits (9..16 bits), 32-bit records and 16-bit ids");
        // measured on B200 (profiles/r01_fused_bucket_experiment.txt): the fused kernel halves the sort time but its task
        // filing (global atomics with return + scattered 8-byte stores) is no faster than segment_kernel's, and the 64-byte run
        // alignment it needs costs the accumulate 4 %: 1280 against 1346 combinations/s on configs[3].  Opt-in only.
        h->fused_seg = fused_ok && h->opt_seg_fused == 2;
        if (h->fused_seg) {
            const int lo_bits = h->plan.bits[0], hi_bits = h->plan.bits[1];
            const int64_t cap2 = ((nfeat + 63) / 64 * 64) + ((int64_t)1 << hi_bits) * (int64_t)bucket_id_stride(lo_bits) + 64;
            if (cap2 >= (1LL << 30)) return fail(h, FSK_EINVAL, "too many g-mers (%lld) for the aligned id stream", (long long)nfeat);
            h->ids_stride = std::max(h->ids_stride, (size_t)((cap2 + 63) / 64 * 64));
            // two CTAs per SM: each may take half of the SM's shared memory less the 1 KB the system reserves per CTA
            h->pad_mask = 31;   // runs start on 64-byte boundaries (the DRAM access granularity); the kernel maps image positions
                                // to runs by 32-id blocks.  128-byte lines would cost 51 pad ids per 141-id run of configs[3].
            const int64_t per_cta = ((int64_t)max_smem + 1024) / 2 - 1024 - 512;
            const int64_t cap_max = ((per_cta - (int64_t)bucket_smem_bytes(0)) * 32 / 66) & ~63LL;
            h->image_cap = (uint32_t)std::max<int64_t>(64, std::min<int64_t>(cap_max, (nfeat + 64 * ((int64_t)1 << lo_bits) + 127) & ~63LL));
            CU(smem_opt_in(bucket_segment_kernel<false>, (int)bucket_smem_bytes(h->image_cap)));
            CU(smem_opt_in(bucket_segment_kernel<true>, (int)bucket_smem_bytes(h->image_cap)));
        }
    }
    {
        // directory form of the segmentation: key spaces of at most 2^16 k-mers on the row path, when the directory (one
        // entry per key and block of rows) is not much larger than the per-record task list it replaces
        int bs = 0;
        while (((N + ((int64_t)1 << bs) - 1) >> bs) > h->opt_dir_blocks) ++bs;
        const int64_t nb = (N + ((int64_t)1 << bs) - 1) >> bs;
        const bool dir_ok = h->rows_path && !h->dense_path && h->mode != MODE_KV && h->keybits <= 16 && !h->fused_seg;
        if (h->opt_seg_dir == 2 && !dir_ok)
            return fail(h, FSK_EINVAL, "seg_dir = 2 needs the row path, at most 16 key bits and records that carry the sequence id");
        // opt-in: measured on configs[3] (profiles/r02_segment_forms.txt) the directory form halves the segmentation (no atomic,
        // no per-record scatter) but its 9.25 M random 8-byte look-ups per combination cost the accumulate more than that
        h->dir_mode = dir_ok && h->opt_seg_dir == 2;
        h->dir_bshift = bs;
        h->dir_nb = (uint32_t)nb;
        const bool lean_ok = h->mode != MODE_KV && !h->fused_seg && !h->dense_path;
        if (h->opt_seg_lean == 2 && !lean_ok) return fail(h, FSK_EINVAL, "seg_lean = 2 needs records that carry the sequence id");
        h->lean_seg = lean_ok && h->opt_seg_lean != 1;
    }
    {
        // rows per accumulate launch: the CTAs resident at once, times opt_wave (default 1)
        const int per_sm = std::max(1, std::min(2048 / h->rows_threads, (int)((size_t)(max_smem + 1024) / (h->rows_smem + 1024))));
        h->wave_rows = n_sm * per_sm * std::max(1, h->opt_wave);
    }
    if (h->rows_path) {
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint16_t, 2, false>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint32_t, 2, false>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint16_t, 2>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint16_t, 4>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint32_t, 2>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint32_t, 4>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint16_t, 2, false, true>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint32_t, 2, false, true>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint16_t, 2, true, true>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint16_t, 4, true, true>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint32_t, 2, true, true>, (int)h->rows_smem));
        CU(smem_opt_in(accumulate_rows_kernel<unsigned long long, uint32_t, 4, true, true>, (int)h->rows_smem));
    }

    // batch: combinations per launch group.  The row path flushes every row of K once per batch, so it
    // wants the batch as large as memory allows; the u32 shared-memory accumulators bound it by 2^32 / maxwin^2.
    const int64_t task_bytes = h->dir_mode ? 2 * (((int64_t)h->dir_nb << h->keybits) * 8) + nfeat * 2 : nfeat * 2 * 8;
    const int64_t per_slot_bytes = nfeat * (2 * (h->mode == MODE_R32 ? 4 : 8) + (h->mode == MODE_KV ? 8 : 0)) + task_bytes +
                                   2 * (int64_t)h->ids_stride * (h->ids16 ? 2 : 4) +
                                   (int64_t)h->plan.npass * ((nfeat + 3071) / 3072) * RADIX * 4 + N * 4 + 4096;
    size_t free_b = 0, total_b = 0;
    CU(mem_info_with_cache(&free_b, &total_b));
    int64_t khat_streams = 0;      // variance mode: one fp64 running mean per local virtual stream (allocated by the build)
    if (h->variance_mode) {
        const int64_t nq = h->user_queue.empty() ? h->ncomb : (int64_t)h->user_queue.size();
        khat_streams = (effective_streams(h, nq) + h->world - 1) / h->world;
    }
    const int64_t k_bytes = h->n_pairs * 8 * (h->variance_mode ? 2 + khat_streams : 1) + (int64_t)n_train * N * 8;   // (+ a second set of means when rounds speculate: checked below)
    int64_t Bsel = h->opt_batch > 0 ? h->opt_batch : MAX_BATCH;
    if (h->opt_batch == 0) {
        if (!h->rows_path) Bsel = std::max<int64_t>(1, (6LL << 20) / std::max<int64_t>(1, nfeat));
        const int64_t budget = ((int64_t)free_b - k_bytes) * 4 / 10;
        // (variance mode on the global-RED path keeps this iteration's integer partial kernel of every slot)
        const bool slot_ks = h->variance_mode && !(h->rows_path || h->dense_path);
        Bsel = std::min(Bsel, std::max<int64_t>(1, budget / std::max<int64_t>(1, per_slot_bytes + (slot_ks ? h->n_pairs * 8 : 0))));
    }
    Bsel = std::min<int64_t>(Bsel, MAX_BATCH);
    Bsel = std::min<int64_t>(Bsel, std::max<int64_t>(1, (int64_t)(4294967295.0 / ((double)maxwin * (double)maxwin))));
    Bsel = std::min<int64_t>(Bsel, std::max<int64_t>(1, h->user_queue.empty() ? h->ncomb : (int64_t)h->user_queue.size()));
    h->wf_depth = 1;
    if (h->variance_mode) {
        // a round gives every live virtual stream of this rank `depth` consecutive slots (iterations speculated past a possible
        // stop); that needs a second buffer of running means to roll a stream back from.  Without room for it -- or on the
        // global-RED path, whose Welford pass is not fused -- one slot per stream, as the reference iterates.
        const int64_t T_local = std::max<int64_t>(1, khat_streams);
        int64_t depth = (h->rows_path || h->dense_path) ? std::max<int64_t>(1, Bsel / T_local) : 1;
        if (h->opt_spec_depth) depth = std::min<int64_t>(depth, h->opt_spec_depth);
        if (h->max_iters > 0) depth = std::min<int64_t>(depth, h->max_iters);
        if (depth > 1 && (double)free_b * 0.9 < (double)k_bytes + (double)T_local * h->n_pairs * 8 + (double)per_slot_bytes * T_local * 2) depth = 1;
        h->wf_depth = (int)depth;
        Bsel = std::min<int64_t>(Bsel, T_local * depth);
    }
    if (h->dense_path) {
        if (h->opt_batch == 0) Bsel = std::min<int64_t>(MAX_BATCH, std::max<int64_t>(1, h->user_queue.empty() ? h->ncomb : (int64_t)h->user_queue.size()));
        Bsel = std::min<int64_t>(Bsel, std::max<int64_t>(1, (4LL << 30) / (N * (int64_t)h->nks * 2)));   // C stays below 4 GB
        if (h->variance_mode) {
            const int64_t T_local = std::max<int64_t>(1, khat_streams);
            int64_t depth = std::max<int64_t>(1, Bsel / T_local);
            if (h->opt_spec_depth) depth = std::min<int64_t>(depth, h->opt_spec_depth);
            if (h->max_iters > 0) depth = std::min<int64_t>(depth, h->max_iters);
            if (depth > 1 && (double)free_b * 0.9 < (double)k_bytes + (double)T_local * h->n_pairs * 8) depth = 1;
            h->wf_depth = (int)depth;
            Bsel = std::min<int64_t>(Bsel, T_local * depth);
        }
        h->dense_chunk = (int)std::max<int64_t>(1, std::min<int64_t>(Bsel, 16777215 / std::max<int64_t>(1, maxwin * maxwin)));
    }
    if (h->rows_path && !h->variance_mode && h->mode != MODE_KV && maxwin <= 2048 && h->opt_heavy_tau >= 0 && h->opt_batch == 0) {
        // the heavy-run stage needs every fp32 accumulator of its contraction below 2^24 (slots x maxwin^2): a somewhat smaller
        // batch keeps the stage available (it is worth 6x on skewed inputs; the last 100 slots of a batch are worth < 1 %)
        const int64_t b_heavy = (int64_t)(16777215.0 / ((double)maxwin * (double)maxwin));
        if (b_heavy >= 64 && b_heavy < Bsel) Bsel = b_heavy;
    }
    h->B = (int)Bsel;
    {
        // heavy runs -> tensor cores: integer modes of the row path, records that carry the sequence id, counts exact in fp16
        // (<= 2048 windows per sequence) and every fp32 accumulator below 2^24 (slots x maxwin^2).  The threshold is the larger
        // of the measured break-even (~0.05 N; 0.06 N used) and what keeps the batch's heavy runs within 65536 columns / 8 GB.
        h->heavy_tau = 0;
        h->heavy_u8 = false;
        const bool ok = h->rows_path && !h->variance_mode && h->mode != MODE_KV && maxwin <= 2048 &&
                        (double)Bsel * (double)maxwin * (double)maxwin < 16777216.0 && !h->fused_seg && h->opt_heavy_tau >= 0 &&
                        !h->opt_overlap;   // (the list and its bitmap are single-buffered: not with the two-stream overlap)
        if (h->opt_heavy_tau > 0 && !ok)
            return fail(h, FSK_EINVAL, "heavy_tau needs the row path in an integer mode, at most 2048 windows per sequence and batch x windows^2 < 2^24");
        if (ok) {
            // a run of d records costs d^2 / 2 updates at 1.3e12 /s on the row path, one column = N^2 / 2 MACs at 6e14 /s here
            const int64_t tau = h->opt_heavy_tau > 0 ? h->opt_heavy_tau : std::max<int64_t>(1024, (N * 5 + 99) / 100);
            if (tau < nfeat) {                                            // a run that long must be possible at all
                h->heavy_tau = h->heavy_tau_min = (uint32_t)tau;
                h->heavy_cap = h->opt_heavy_cap ? (uint32_t)h->opt_heavy_cap
                                                : (uint32_t)std::max<int64_t>(128, std::min<int64_t>(65536, ((8LL << 30) / (N * 2)) & ~127LL));
                // byte columns (kind::i8 contraction) when every count fits a byte and the list is whole 128-column k-blocks
                h->heavy_u8 = h->opt_dense_u8 && maxwin <= 255 && h->heavy_cap % 128 == 0;
            }
        }
        h->heavy_live = true;
        h->heavy_probe_pending = false;
        h->heavy_idle = 0;
    }
    h->sort_tiles = (uint32_t)((nfeat + SORT_THREADS * h->sort_items - 1) / (SORT_THREADS * h->sort_items));
    h->seg_rows = SEG_ROWS_DEFAULT;
    h->seg_tiles = (uint32_t)((nfeat + seg_tile_records(h->seg_rows) - 1) / seg_tile_records(h->seg_rows));
    h->lean_tiles = (uint32_t)((nfeat + 3 + lean_tile(h->rec_bytes == 4 ? 4 : 8) - 1) / lean_tile(h->rec_bytes == 4 ? 4 : 8));

    tr.lap("path selection, kernel attributes");
    // device inputs
    uint8_t* d_codes = nullptr;
    int64_t *d_off = nullptr, *d_woff = nullptr;
    std::vector<int64_t> woff((size_t)N + 1);
    woff[0] = 0;
    for (int64_t i = 0; i < N; ++i) woff[(size_t)i + 1] = woff[(size_t)i] + (off0[(size_t)i + 1] - off0[(size_t)i] - h->g + 1);
    ALLOC(d_codes, total);
    ALLOC(d_off, N + 1);
    ALLOC(d_woff, N + 1);
    CU(cudaMemcpyAsync(d_codes, dense.data(), (size_t)total, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_off, off0.data(), sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_woff, woff.data(), sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, h->stream));
    if (h->gw32) { uint32_t* p; ALLOC(p, nfeat); h->d_gw0 = p; }
    else { uint64_t* p; ALLOC(p, nfeat); h->d_gw0 = p; }
    if (h->NW == 2) ALLOC(h->d_gw1, nfeat);
    ALLOC(h->d_wseq, nfeat);
    {
        const int blocks = (int)std::min<int64_t>((N + 7) / 8, 148 * 16);
        if (h->NW == 2)
            build_gwords_kernel<uint64_t, 2><<<blocks, 256, 0, h->stream>>>(d_codes, d_off, d_woff, N, h->g, b, cpw, (uint64_t*)h->d_gw0, h->d_gw1, h->d_wseq);
        else if (h->gw32)
            build_gwords_kernel<uint32_t, 1><<<blocks, 256, 0, h->stream>>>(d_codes, d_off, d_woff, N, h->g, b, cpw, (uint32_t*)h->d_gw0, nullptr, h->d_wseq);
        else
            build_gwords_kernel<uint64_t, 1><<<blocks, 256, 0, h->stream>>>(d_codes, d_off, d_woff, N, h->g, b, cpw, (uint64_t*)h->d_gw0, nullptr, h->d_wseq);
        h->launches = 1;
        CU(cudaGetLastError());
    }

    // scratch for B slots (the dense path sorts nothing: one record keeps the pointers valid)
    const size_t bn = h->dense_path ? 1 : (size_t)h->B * (size_t)nfeat;
    if (h->dense_path) { h->sort_tiles = h->seg_tiles = h->lean_tiles = 1; h->ids_stride = 64; }
    // (+ 64 bytes: the register-blocked segmentation reads whole aligned 16-byte vectors around the slots' ends)
    { unsigned char* p; ALLOC(p, bn * (h->mode == MODE_R32 ? 4 : 8) + 64); h->d_recA = p; }
    { unsigned char* p; ALLOC(p, bn * (h->mode == MODE_R32 ? 4 : 8) + 64); h->d_recB = p; }
    if (h->mode == MODE_KV) { ALLOC(h->d_valA, bn); ALLOC(h->d_valB, bn); }
    const int B = h->B;
    const size_t ghist_words = (size_t)B * MAX_PASS * RADIX, ticket_words = 64;
    const size_t rowcount_words = 0;
    const size_t status_words = (size_t)h->plan.npass * B * h->sort_tiles * RADIX;
    const size_t seg_status_words = (size_t)B * std::max(h->seg_tiles, h->lean_tiles);
    h->heavy_bits_stride = h->heavy_tau ? ((h->ids_stride >> (h->ids16 ? 3 : 2)) + 31) / 32 + 1 : 0;
    const size_t heavy_words = (size_t)B * h->heavy_bits_stride;
    h->zero_bytes = 4 * (ghist_words + ticket_words + status_words + seg_status_words + rowcount_words + heavy_words);
    ALLOC(h->d_zero, h->zero_bytes);
    h->d_ghist = (uint32_t*)h->d_zero;
    h->d_ticket = h->d_ghist + ghist_words;
    h->d_status = h->d_ticket + ticket_words;
    h->d_seg_status = h->d_status + status_words;
    h->d_heavy_bits = h->d_seg_status + seg_status_words;
    for (int i = 0; i < 2; ++i) {
        unsigned char* p;
        ALLOC(p, (size_t)B * h->ids_stride * (h->ids16 ? 2 : 4));
        h->d_ids[i] = p;
        ALLOC(h->d_task[i], h->dir_mode ? 1 : bn);
        if (h->dir_mode) ALLOC(h->d_tdir[i], (size_t)B * ((size_t)h->dir_nb << h->keybits));
    }
    if (h->dir_mode) ALLOC(h->d_wkey, bn);
    ALLOC(h->d_fill, (h->dense_path || h->dir_mode) ? 1 : (size_t)B * (size_t)N);
    {
        std::vector<uint32_t> w32((size_t)N + 1);
        for (int64_t i = 0; i <= N; ++i) w32[(size_t)i] = (uint32_t)woff[(size_t)i];
        ALLOC(h->d_woff32, N + 1);
        CU(cudaMemcpy(h->d_woff32, w32.data(), sizeof(uint32_t) * (size_t)(N + 1), cudaMemcpyHostToDevice));
    }
    if (h->dense_path) {
        // variance mode: the form of the Welford contraction is fixed here (its means' layout and its operands depend on it)
        h->wf_regs = h->variance_mode && h->opt_wf_regs;
        h->dense_u8 = h->opt_dense_u8 && maxwin <= 255 && h->nks % (2 * DG_BK) == 0 && (h->wf_regs || !h->variance_mode);   // (the streamed Welford form is fp16 only)
        h->dense_ld = (size_t)B * h->nks;                        // elements per row: bytes when dense_u8
        ALLOC(h->d_C, h->dense_u8 ? ((size_t)N * h->dense_ld + 1) / 2 : (size_t)N * h->dense_ld);
        int rc_ = encode_operand_map(h, &h->tmap_C, h->d_C, h->dense_ld, h->dense_u8);
        if (rc_) return rc_;
        const int count_smem = DENSE_COUNT_WARPS * (int)h->nks * 2;
        CU(smem_opt_in(dense_count_kernel<uint64_t, 2>, count_smem));
        CU(smem_opt_in(dense_count_kernel<uint64_t, 1>, count_smem));
        CU(smem_opt_in(dense_count_kernel<uint32_t, 1>, count_smem));
        CU(smem_opt_in(dense_count_kernel<uint64_t, 2, true>, count_smem));
        CU(smem_opt_in(dense_count_kernel<uint64_t, 1, true>, count_smem));
        CU(smem_opt_in(dense_count_kernel<uint32_t, 1, true>, count_smem));
    }
    if (h->heavy_tau) {
        // optional stage: it must not be what makes a large upload run out of memory (K and the outputs are still to come)
        size_t free_now = 0, total_now = 0;
        CU(mem_info_with_cache(&free_now, &total_now));
        if ((double)free_now < (double)N * h->heavy_cap * 2.0 + (double)k_bytes + (double)(2LL << 30)) h->heavy_tau = h->heavy_tau_min = 0;
    }
    if (h->heavy_tau) {
        ALLOC(h->d_H, h->heavy_u8 ? ((size_t)N * h->heavy_cap + 1) / 2 : (size_t)N * h->heavy_cap);
        ALLOC(h->d_heavy_list, h->heavy_cap);
        if (!h->h_heavy_count) CU(cudaMallocHost((void**)&h->h_heavy_count, sizeof(uint32_t)));
        if (!h->ev_heavy) CU(cudaEventCreateWithFlags(&h->ev_heavy, cudaEventDisableTiming));
        int rc_ = encode_operand_map(h, &h->tmap_H, h->d_H, h->heavy_cap, h->heavy_u8);
        if (rc_) return rc_;
    }
    if (h->dense_path || h->heavy_tau) {
        int rc_ = make_tile_order(h);
        if (rc_) return rc_;
    }
    ALLOC(h->d_counters, 4);
    CU(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(unsigned long long), h->stream));
    ALLOC(h->d_flag, 1);
    CU(cudaMemsetAsync(h->d_flag, 0, sizeof(uint32_t), h->stream));

    // accumulators
    // variance mode keeps a per-slot Ks only on the global-RED path; the row and dense paths fold the Welford step into the
    // accumulate and never materialise Ks
    h->ks_slots = (h->variance_mode && !(h->rows_path || h->dense_path)) ? B : 1;
    ALLOC(h->d_Kint, (size_t)h->ks_slots * h->n_pairs);
    CU(cudaMemsetAsync(h->d_Kint, 0, sizeof(unsigned long long) * (size_t)h->ks_slots * h->n_pairs, h->stream));
    if (h->variance_mode) {
        ALLOC(h->d_Kf, h->n_pairs);
        CU(cudaMemsetAsync(h->d_Kf, 0, sizeof(double) * (size_t)h->n_pairs, h->stream));
        // partial sums of the variance per slot: one per Welford block, per row (fused flush of the row path) or per epilogue
        // warp of every tile (fused epilogue of the dense path)
        const int64_t Tt = (N + DG_TILE - 1) / DG_TILE;
        h->sums_stride = (uint32_t)std::max<int64_t>(WELFORD_BLOCKS, h->dense_path ? 4 * Tt * (Tt + 1) : N * h->col_windows);
        ALLOC(h->d_block_sums, (size_t)B * h->sums_stride);
        ALLOC(h->d_var, B);
        ALLOC(h->d_wf, 1);
    }
    tr.lap("allocations, H2D, g-mer words (launched)");
    CU(cudaStreamSynchronize(h->stream));
    tr.lap("stream synchronise (memsets of K)");
    cached_free(d_codes); cached_free(d_off); cached_free(d_woff);

    h->combos_done = 0;
    for (double& v : h->ms) v = 0;
    h->stdevs.clear();
    int rc = build_queue(h);
    if (rc) return rc;
    h->uploaded = true;
    return FSK_OK;
}
}  // namespace

extern "C" {

int fsk_upload(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test) {
    return fsk_upload_split(h, codes, offsets, n_train, nullptr, nullptr, n_test);
}

int fsk_upload_split(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, const int32_t* codes_test,
                     const int64_t* offsets_test, int64_t n_test) {
    if (!is_team(h)) return upload_one(h, codes, offsets, n_train, n_test, codes_test, offsets_test);
    // every member must work through the same combination order: the leader draws the wall-clock seed of an unseeded
    // shuffle (fastsk_kernel.cpp:36-38) once for the whole team
    sync_team(h);
    const uint64_t seed = (uint64_t)std::time(0);
    for (fsk_handle* w : h->team) { w->run_seed_set = true; w->run_seed = seed; }
    return team_run(h, [&](fsk_handle* w) { return upload_one(w, codes, offsets, n_train, n_test, codes_test, offsets_test); });
}

int fsk_reset_partial(fsk_handle* h) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (is_team(h)) return team_run(h, reset_one);
    return reset_one(h);
}

int fsk_accumulate_combos(fsk_handle* h, const int32_t* combos, int64_t n, int sync) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (h->variance_mode) return fail(h, FSK_ESTATE, "fsk_accumulate_combos needs an integer mode (exact or skip_variance)");
    if (n < 0 || (n > 0 && !combos)) return fail(h, FSK_EINVAL, "bad combination list");
    if (!is_team(h)) return accumulate_one(h, combos, n, sync);
    const int64_t G = (int64_t)h->team.size();      // dealt round-robin, like shard_work
    return team_run(h, [&](fsk_handle* w) {
        std::vector<int32_t> mine;
        for (int64_t i = w->rank; i < n; i += G) mine.push_back(combos[i]);
        return accumulate_one(w, mine.data(), (int64_t)mine.size(), sync);
    });
}

int fsk_build_partial(fsk_handle* h) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (is_team(h)) return team_run(h, build_one);
    return build_one(h);
}

}  // extern "C"

namespace {
int reset_one(fsk_handle* h) {
    CU(cudaSetDevice(h->device));
    CU(cudaMemsetAsync(h->d_Kint, 0, sizeof(unsigned long long) * (size_t)h->ks_slots * h->n_pairs, h->stream));
    if (h->d_Kf) CU(cudaMemsetAsync(h->d_Kf, 0, sizeof(double) * (size_t)h->n_pairs, h->stream));
    h->built = h->finalized = false;
    return FSK_OK;
}

int accumulate_one(fsk_handle* h, const int32_t* combos, int64_t n, int sync) {
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev_sync, h->stream));
    CU(cudaStreamWaitEvent(h->pre_stream, h->ev_sync, 0));
    for (int64_t i = 0; i < n; i += h->B) {
        const int nb = (int)std::min<int64_t>(h->B, n - i);
        int rc = run_batch(h, combos + i, nb, h->d_Kint, 0);
        if (rc) return rc;
    }
    if (sync) {
        bool bad = false;
        int rc = sort_was_unstable(h, &bad);
        if (rc) return rc;
        if (bad) return fail(h, FSK_ECUDA, "sort verification failed: records out of order after the optimistic ranking; set option safe_rank = 1");
    }
    return FSK_OK;
}

int build_one(fsk_handle* h) {
    CU(cudaSetDevice(h->device));
    for (int attempt = 0;; ++attempt) {
        int rc = build_partial_once(h);
        if (rc) return rc;
        bool bad = false;
        rc = sort_was_unstable(h, &bad);
        if (rc) return rc;
        if (!bad) break;
        if (h->safe_rank || attempt > 0) return fail(h, FSK_ECUDA, "sort verification failed with the safe ranking");
        h->safe_rank = true;       // never observed; repeat the whole build with the match-mask ranking
        h->rank_fallbacks++;
        CU(cudaMemsetAsync(h->d_flag, 0, sizeof(uint32_t), h->stream));
    }
    h->built = true;
    return FSK_OK;
}

int build_partial_once(fsk_handle* h) {
    int rc = reset_one(h);
    if (rc) return rc;
    h->stdevs.clear();
    const int64_t nq = (int64_t)h->queue.size();
    if (nq < 1) return fail(h, FSK_EINVAL, "empty combination queue");
    const int T = effective_streams(h, nq);
    h->T_eff = T;

    if (!h->variance_mode) {
        std::vector<int32_t> mine;
        shard_work(h, mine);
        rc = accumulate_one(h, mine.data(), (int64_t)mine.size(), 0);
        if (rc) return rc;
    } else {
        // variance mode: T independent virtual streams (fastsk_kernel.cpp:188-281), stream tid owned by rank tid % world.
        // A round gives every live stream up to wf_depth CONSECUTIVE iterations, applied in order by one launch group; the
        // host then walks each stream's variance statistics in order and applies the stop rule (fastsk_kernel.cpp:243-262).
        // A stream whose rule fires before the round's last iteration is rolled back: its running mean is rebuilt from the
        // buffer the round started from (`cur`), with only the iterations it really ran.  One host round trip per round.
        struct Stream { int tid; int64_t item; int iter; bool working; double *cur, *alt; int depth, used; };
        std::vector<Stream> streams;
        std::vector<int32_t> my_streams;
        shard_work(h, my_streams);
        const bool fused_wf = h->rows_path || h->dense_path;
        const bool pingpong = fused_wf && h->wf_depth > 1;
        // (the tensor-core path keeps its running means tile-major: 128 x 128 cells per lower-triangle tile, fsk_dense.cuh)
        const int64_t Ttiles = (h->N + DG_TILE - 1) / DG_TILE;
        const size_t khat_elems = h->dense_path ? (size_t)(Ttiles * (Ttiles + 1) / 2) * (size_t)(DG_TILE * DG_TILE) : (size_t)h->n_pairs;
        if (!my_streams.empty()) {   // one allocation for the running means of all local streams (cudaMalloc/cudaFree are slow)
            double* all;
            const size_t per = khat_elems * (pingpong ? 2 : 1);
            ALLOC(all, per * my_streams.size());
            h->d_Khat.push_back(all);
            CU(cudaMemsetAsync(all, 0, sizeof(double) * per * my_streams.size(), h->stream));
            for (size_t i = 0; i < my_streams.size(); ++i) {
                double* c = all + i * per;
                streams.push_back({my_streams[i], my_streams[i], 1, true, c, pingpong ? c + khat_elems : c, 0, 0});
            }
        }
        std::vector<double> var_host((size_t)h->B);
        const int64_t Tt = (h->N + DG_TILE - 1) / DG_TILE;
        const int n_sums = !fused_wf ? WELFORD_BLOCKS : (h->dense_path ? (int)(4 * Tt * (Tt + 1)) : (int)(h->N * h->col_windows));
        // one launch group: the given streams, stream i running its next depth[i] iterations
        auto run_round = [&](const std::vector<Stream*>& grp, int nslots) -> int {
            int32_t combos[MAX_BATCH];
            WelfordSpec wf;
            memset(&wf, 0, sizeof wf);
            int slot = 0;
            for (size_t gi = 0; gi < grp.size(); ++gi) {
                Stream* st = grp[gi];
                wf.khat_in[gi] = st->cur; wf.khat_out[gi] = st->alt; wf.iter0[gi] = st->iter;
                wf.slot0[gi] = (uint16_t)slot; wf.depth[gi] = (uint16_t)st->depth;
                for (int d = 0; d < st->depth; ++d) combos[slot++] = h->queue[(size_t)(st->item + (int64_t)T * d)];
            }
            wf.sums = h->d_block_sums; wf.sums_stride = h->sums_stride; wf.n_train = h->n_train;
            if (fused_wf) CU(cudaMemcpyAsync(h->d_wf, &wf, sizeof wf, cudaMemcpyHostToDevice, h->stream));
            if (fused_wf && h->rows_path && h->col_windows > 1)   // rows below a column window write no partial sum for it
                CU(cudaMemsetAsync(h->d_block_sums, 0, sizeof(double) * (size_t)nslots * h->sums_stride, h->stream));
            h->wf_active = fused_wf;
            h->wf_groups = (int)grp.size();
            int rc2 = run_batch(h, combos, nslots, h->d_Kint, (size_t)h->n_pairs);
            h->wf_active = false;
            if (rc2) return rc2;
            Span sp(h, PC_WELFORD);
            if (!fused_wf) {   // global-RED path: Ks of each slot (zero on entry), a separate Welford pass per stream (depth is 1)
                for (size_t gi = 0; gi < grp.size(); ++gi) {
                    welford_kernel<unsigned long long><<<WELFORD_BLOCKS, 256, 0, h->stream>>>(
                        h->d_Kint + gi * (size_t)h->n_pairs, grp[gi]->cur, h->n_pairs, h->n_train_pairs, grp[gi]->iter,
                        h->d_block_sums + gi * (size_t)h->sums_stride);
                    h->launches++;
                }
            }
            welford_final_kernel<<<nslots, 256, 0, h->stream>>>(h->d_block_sums, (size_t)h->sums_stride, n_sums, h->d_var);
            h->launches++;
            CU(cudaGetLastError());
            return FSK_OK;
        };
        while (true) {
            std::vector<Stream*> active;
            for (auto& s : streams) if (s.working) active.push_back(&s);
            if (active.empty()) break;
            // groups of streams that fit the batch; every stream of a group gets the same share of the slots
            const size_t per_group = std::min<size_t>(active.size(), (size_t)h->B);
            std::vector<Stream*> redo;
            for (size_t a0 = 0; a0 < active.size(); a0 += per_group) {
                std::vector<Stream*> grp(active.begin() + (long)a0, active.begin() + (long)std::min(active.size(), a0 + per_group));
                const int share = fused_wf ? std::max(1, std::min(h->wf_depth, h->B / (int)grp.size())) : 1;
                int nslots = 0;
                for (Stream* st : grp) {
                    int64_t left = (nq - st->item + T - 1) / T;                                  // items left in its strided queue
                    if (h->max_iters != -1) left = std::min<int64_t>(left, std::max(1, h->max_iters - st->iter + 1));
                    st->depth = (int)std::max<int64_t>(1, std::min<int64_t>(share, left));
                    nslots += st->depth;
                }
                rc = run_round(grp, nslots);
                if (rc) return rc;
                CU(cudaMemcpyAsync(var_host.data(), h->d_var, sizeof(double) * (size_t)nslots, cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                int slot = 0;
                for (Stream* st : grp) {
                    st->used = 0;
                    for (int d = 0; d < st->depth && st->working; ++d) {
                        // fastsk_kernel.cpp:130-142 then 244-256
                        double v = var_host[(size_t)(slot + d)] / (double)h->n_train_pairs;
                        if (st->iter == 1) v = 9999999;
                        else v /= st->iter - 1;
                        const double sd = std::sqrt(v / st->iter);
                        if (st->tid == 0) h->stdevs.push_back(sd);
                        if (h->delta / sd > 1.96) st->working = false;
                        if (h->max_iters != -1 && st->iter >= h->max_iters) st->working = false;
                        st->item += T;
                        if (st->item >= nq) st->working = false;
                        st->iter++;
                        st->used++;
                    }
                    slot += st->depth;
                    if (st->used < st->depth) redo.push_back(st);          // stopped inside the round: `alt` ran too far
                    else std::swap(st->cur, st->alt);
                }
            }
            // roll back: rebuild the mean of every stream that stopped early from where its round started
            for (size_t a0 = 0; a0 < redo.size();) {
                std::vector<Stream*> grp;
                int nslots = 0;
                while (a0 < redo.size() && nslots + redo[a0]->used <= h->B) {
                    Stream* st = redo[a0++];
                    st->item -= (int64_t)T * st->used; st->iter -= st->used; st->depth = st->used;   // as the round found it
                    nslots += st->depth;
                    grp.push_back(st);
                }
                rc = run_round(grp, nslots);
                if (rc) return rc;
                for (Stream* st : grp) { st->item += (int64_t)T * st->used; st->iter += st->used; std::swap(st->cur, st->alt); }
            }
        }
        // merge (fastsk_kernel.cpp:296-313): sum of the streams' running means, in stream order
        for (auto& s : streams) {
            if (h->dense_path) welford_untile_kernel<<<(unsigned)(Ttiles * (Ttiles + 1) / 2), 256, 0, h->stream>>>(h->d_Kf, s.cur, h->d_tile_order, h->N, h->wf_regs);
            else add_f64_kernel<<<592, 256, 0, h->stream>>>(h->d_Kf, s.cur, h->n_pairs);
            h->launches++;
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
        for (auto& p : h->d_Khat) dev_free(p);
        h->d_Khat.clear();
    }
    CU(cudaStreamSynchronize(h->stream));
    return FSK_OK;
}

// normalisation of the rows this handle is responsible for: all of them, or -- when the partial kernels of all ranks are
// reachable (peer_parts) -- rank r's share of the train rows and of the test rows
int finalize_one(fsk_handle* h) {
    CU(cudaSetDevice(h->device));
    const int W = (int)h->peer_parts.size();
    h->sharded = W > 1;
    const int64_t r = h->sharded ? h->rank : 0, w = h->sharded ? W : 1;
    // rows [n * cum(r) / total, n * cum(r + 1) / total): equal shares, or by the ranks' weights (a box whose GPUs do not reach
    // the host equally fast hands the results back through the fast ones: fsk_set_output_weights)
    double before = (double)r, total = (double)w, mine = 1.0;
    if (h->sharded && (int)h->out_weights.size() == W) {
        before = 0; total = 0;
        for (int q = 0; q < W; ++q) { if (q < r) before += h->out_weights[(size_t)q]; total += h->out_weights[(size_t)q]; }
        mine = h->out_weights[(size_t)r];
    }
    auto cut = [&](int64_t n, double c) { return (int64_t)std::min<double>((double)n, std::floor((double)n * (c / total) + 1e-9)); };
    const bool last = before + mine >= total - 1e-12;
    h->tr_r0 = cut(h->n_train, before); h->tr_nr = (last ? h->n_train : cut(h->n_train, before + mine)) - h->tr_r0;
    h->te_r0 = cut(h->n_test, before); h->te_nr = (last ? h->n_test : cut(h->n_test, before + mine)) - h->te_r0;
    if (!h->d_diag) ALLOC(h->d_diag, h->N);
    const size_t need_tr = (size_t)h->tr_nr * h->n_train, need_te = (size_t)h->te_nr * h->n_train;
    if (!h->d_train || need_tr > h->train_cap) { dev_free(h->d_train); ALLOC(h->d_train, need_tr); h->train_cap = need_tr; }
    if (need_te && (!h->d_test || need_te > h->test_cap)) { dev_free(h->d_test); ALLOC(h->d_test, need_te); h->test_cap = need_te; }
    int rc = h->variance_mode ? finalize_typed<double>(h, h->d_Kf) : finalize_typed<unsigned long long>(h, h->d_Kint);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    h->finalized = true;
    return FSK_OK;
}

int sync_one(fsk_handle* h) {
    if (!h->stream) return FSK_OK;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (h->uploaded) {
        bool bad = false;
        int rc = sort_was_unstable(h, &bad);
        if (rc) return rc;
        if (bad) return fail(h, FSK_ECUDA, "sort verification failed: records out of order after the optimistic ranking; set option safe_rank = 1");
    }
    return FSK_OK;
}

// Which members should hand the results back to the host?  All of them, unless the box's GPUs reach host memory unequally:
// probe a device -> host copy on all members at once, then on the faster half alone; if the half alone moves clearly more
// bytes per second than all together, only those members take output rows.  Once per process and device list.
int team_output_weights(fsk_handle* h) {
    if (!h->out_weights.empty() || h->team.size() < 4 || getenv("FSK_EQUAL_OUTPUT_SHARES")) return FSK_OK;
    static std::mutex mu;
    static std::map<std::string, std::vector<double>> known;
    std::string key;
    for (fsk_handle* w : h->team) key += std::to_string(w->device) + ",";
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = known.find(key);
        if (it != known.end()) { h->out_weights = it->second; h->weights_auto = true; for (fsk_handle* w : h->team) w->out_weights = it->second; return FSK_OK; }
    }
    const size_t n = h->team.size(), bytes = (size_t)256 << 20;
    std::vector<double> t_all(n, 0.0), t_sub(n, 0.0);
    std::vector<char> in_sub(n, 0);
    auto probe = [&](std::vector<double>& t, bool subset) {
        return team_run(h, [&](fsk_handle* w) {
            if (subset && !in_sub[(size_t)w->rank]) return (int)FSK_OK;
            return fsk_probe_d2h(w->device, bytes, &t[(size_t)w->rank]);
        });
    };
    int rc = probe(t_all, false);            // (first touch of the scratch buffers)
    if (rc == FSK_OK) rc = probe(t_all, false);
    if (rc) return FSK_OK;                   // no probe, equal shares
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return t_all[a] < t_all[b]; });
    for (size_t i = 0; i < n / 2; ++i) in_sub[order[i]] = 1;
    if (probe(t_sub, true)) return FSK_OK;
    const double agg_all = (double)n * bytes / *std::max_element(t_all.begin(), t_all.end());
    double worst_sub = 0;
    for (size_t i = 0; i < n; ++i) if (in_sub[i]) worst_sub = std::max(worst_sub, t_sub[i]);
    const double agg_sub = (double)(n / 2) * bytes / worst_sub;
    std::vector<double> wts(n, 1.0);
    if (agg_sub > 1.25 * agg_all) for (size_t i = 0; i < n; ++i) wts[i] = in_sub[i] ? 1.0 : 0.0;
    if (getenv("FSK_TRACE")) {
        fprintf(stderr, "[fsk] device -> host, all %zu GPUs at once: %.1f GB/s; the faster half alone: %.1f GB/s; output rows from:", n, agg_all / 1e9, agg_sub / 1e9);
        for (size_t i = 0; i < n; ++i) if (wts[i] > 0) fprintf(stderr, " %d", h->team[i]->device);
        fprintf(stderr, "\n");
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        known[key] = wts;
    }
    h->out_weights = wts;
    h->weights_auto = true;
    for (fsk_handle* w : h->team) w->out_weights = wts;
    return FSK_OK;
}

// peer access between all devices of the team, and every member's list of the partial kernels of all members
int team_link_peers(fsk_handle* h) {
    for (fsk_handle* w : h->team) {
        cudaSetDevice(w->device);
        for (fsk_handle* o : h->team) {
            if (o == w) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, w->device, o->device);
            if (!can) { cudaSetDevice(h->device); return fail(h, FSK_ECUDA, "device %d cannot access device %d's memory", w->device, o->device); }
            const cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaSetDevice(h->device); return fail(h, FSK_ECUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    }
    cudaSetDevice(h->device);
    for (fsk_handle* w : h->team) {
        w->peer_parts.clear();
        for (fsk_handle* o : h->team) w->peer_parts.push_back(own_part(o));
    }
    return FSK_OK;
}
}  // namespace

extern "C" {

int fsk_partial_buffer(fsk_handle* h, void** dev_ptr, int64_t* n_elems, int* dtype) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (dev_ptr) *dev_ptr = own_part(h);
    if (n_elems) *n_elems = h->n_pairs;
    if (dtype) *dtype = h->variance_mode ? FSK_DT_F64 : FSK_DT_I64;
    return FSK_OK;
}

/* ---- sharded finalisation between the processes of a torchrun launch: CUDA IPC over NVLink --------------------------- */

int fsk_ipc_export_partial(fsk_handle* h, void* handle_out) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (!handle_out) return fail(h, FSK_EINVAL, "handle_out is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == FSK_IPC_HANDLE_BYTES, "FSK_IPC_HANDLE_BYTES");
    CU(cudaSetDevice(h->device));
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, own_part(h)));
    memcpy(handle_out, &mh, sizeof mh);
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        g_cache.exported[own_part(h)] = h->device;
    }
    return FSK_OK;
}

int fsk_set_peer_partials(fsk_handle* h, const void* handles, int world) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (is_team(h)) return fail(h, FSK_ESTATE, "a team of in-process devices exchanges its partial kernels by itself");
    if (world != h->world || !handles) return fail(h, FSK_EINVAL, "need the handles of all %d ranks", h->world);
    CU(cudaSetDevice(h->device));
    release_peers(h);
    h->peer_parts.assign((size_t)world, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == h->rank) { h->peer_parts[(size_t)r] = own_part(h); continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, (const char*)handles + (size_t)r * FSK_IPC_HANDLE_BYTES, sizeof mh);
        const std::pair<int, std::string> key(h->device, std::string((const char*)&mh, sizeof mh));
        void* q = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_ipc.mu);
            auto it = g_ipc.open.find(key);
            if (it != g_ipc.open.end()) q = it->second;
            else {
                const cudaError_t e = cudaIpcOpenMemHandle(&q, mh, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    release_peers(h);
                    return fail(h, FSK_ECUDA, "cudaIpcOpenMemHandle of rank %d's partial kernel failed: %s", r, cudaGetErrorString(e));
                }
                g_ipc.open[key] = q;
            }
        }
        h->ipc_opened.push_back(q);
        h->peer_parts[(size_t)r] = q;
    }
    return FSK_OK;
}

int fsk_set_peer_pointers(fsk_handle* h, void* const* parts, int world) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (is_team(h)) return fail(h, FSK_ESTATE, "a team of in-process devices exchanges its partial kernels by itself");
    if (world != h->world || !parts) return fail(h, FSK_EINVAL, "need the partial buffers of all %d ranks", h->world);
    CU(cudaSetDevice(h->device));
    release_peers(h);
    for (int r = 0; r < world; ++r) {
        if (!parts[r]) { h->peer_parts.clear(); return fail(h, FSK_EINVAL, "partial buffer of rank %d is NULL", r); }
        h->peer_parts.push_back(r == h->rank ? own_part(h) : parts[r]);
    }
    return FSK_OK;
}

int fsk_release_peers(fsk_handle* h) {
    if (h->stream) CU(cudaSetDevice(h->device));
    if (!is_team(h)) release_peers(h);
    return FSK_OK;
}

int fsk_set_output_weights(fsk_handle* h, const double* weights, int world) {
    if (!weights || world < 1) { h->out_weights.clear(); for (fsk_handle* w : h->team) w->out_weights.clear(); return FSK_OK; }
    if (world > MAX_PEERS) return fail(h, FSK_EINVAL, "at most %d ranks", MAX_PEERS);
    double tot = 0;
    for (int r = 0; r < world; ++r) {
        if (!(weights[r] >= 0)) return fail(h, FSK_EINVAL, "output weights must be non-negative");
        tot += weights[r];
    }
    if (!(tot > 0)) return fail(h, FSK_EINVAL, "at least one rank must hand rows back");
    h->out_weights.assign(weights, weights + world);
    for (fsk_handle* w : h->team) w->out_weights = h->out_weights;
    return FSK_OK;
}

int fsk_probe_d2h(int device, size_t bytes, double* seconds) {
    if (!seconds || bytes == 0) return FSK_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); g_create_error = "no usable CUDA device"; return FSK_ECUDA; }
    // scratch per device, kept for the life of the process (a probe is a few hundred MB, once per launch)
    static std::mutex mu;
    static std::map<int, std::pair<void*, void*>> scratch;
    static std::map<int, size_t> cap;
    void *d = nullptr, *hst = nullptr;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cap[device] < bytes) {
            if (scratch[device].first) { cudaFree(scratch[device].first); cudaFreeHost(scratch[device].second); }
            if (cudaMalloc(&d, bytes) != cudaSuccess || cudaHostAlloc(&hst, bytes, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                if (d) cudaFree(d);
                scratch[device] = {nullptr, nullptr}; cap[device] = 0;
                return FSK_ENOMEM;
            }
            scratch[device] = {d, hst}; cap[device] = bytes;
        }
        d = scratch[device].first; hst = scratch[device].second;
    }
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = cudaMemcpy(hst, d, bytes, cudaMemcpyDeviceToHost);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (e != cudaSuccess) { cudaGetLastError(); return FSK_ECUDA; }
    return FSK_OK;
}

int fsk_output_rows(fsk_handle* h, int64_t* train_r0, int64_t* train_nr, int64_t* test_r0, int64_t* test_nr) {
    if (!h->finalized) return fail(h, FSK_ESTATE, "no kernel computed yet");
    if (train_r0) *train_r0 = h->tr_r0;
    if (train_nr) *train_nr = is_team(h) ? h->n_train : h->tr_nr;
    if (test_r0) *test_r0 = h->te_r0;
    if (test_nr) *test_nr = is_team(h) ? h->n_test : h->te_nr;
    return FSK_OK;
}

int fsk_finalize(fsk_handle* h) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (!is_team(h)) return finalize_one(h);
    // every member normalises its rows of the outputs, reading the partial kernels of all members over NVLink
    int rc = team_link_peers(h);
    if (rc) return rc;
    team_output_weights(h);
    return team_run(h, finalize_one);
}

int fsk_compute(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test) {
    return fsk_compute_split(h, codes, offsets, n_train, nullptr, nullptr, n_test);
}

int fsk_compute_split(fsk_handle* h, const int32_t* codes, const int64_t* offsets, int64_t n_train, const int32_t* codes_test,
                      const int64_t* offsets_test, int64_t n_test) {
    int rc = fsk_upload_split(h, codes, offsets, n_train, codes_test, offsets_test, n_test);
    if (rc) return rc;
    rc = fsk_build_partial(h);
    if (rc) return rc;
    return fsk_finalize(h);
}

int fsk_stream(fsk_handle* h, void** cuda_stream) {
    if (!h->stream) return fail(h, FSK_ESTATE, "no stream before upload");
    *cuda_stream = (void*)h->stream;
    return FSK_OK;
}
int fsk_synchronize(fsk_handle* h) {
    if (!h->stream) return FSK_OK;
    if (is_team(h)) return team_run(h, sync_one);
    return sync_one(h);
}

int fsk_shape(fsk_handle* h, int64_t* n_train, int64_t* n_test, int64_t* nfeat, int64_t* n_combos) {
    if (n_train) *n_train = h->n_train;
    if (n_test) *n_test = h->n_test;
    if (nfeat) *nfeat = h->nfeat;
    if (n_combos) *n_combos = h->ncomb;
    return FSK_OK;
}

#define NEED_FINAL()                                                                       \
    do {                                                                                   \
        if (!h->finalized) return fail(h, FSK_ESTATE, "no kernel computed yet");           \
        CU(cudaSetDevice(h->device));                                                      \
    } while (0)

// `out` is always the FULL row-major matrix; a handle that holds only some rows (sharded finalisation) fills those rows
int fsk_get_train_kernel(fsk_handle* h, double* out) {
    if (!h->finalized) return fail(h, FSK_ESTATE, "no kernel computed yet");
    auto one = [out](fsk_handle* w) -> int {
        fsk_handle* h = w;
        CU(cudaSetDevice(h->device));
        if (h->tr_nr == 0) return FSK_OK;
        CU(cudaMemcpyAsync(out + (size_t)h->tr_r0 * h->n_train, h->d_train, sizeof(double) * (size_t)h->tr_nr * h->n_train, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return FSK_OK;
    };
    return is_team(h) ? team_run(h, one) : one(h);
}
int fsk_get_test_kernel(fsk_handle* h, double* out) {
    if (!h->finalized) return fail(h, FSK_ESTATE, "no kernel computed yet");
    auto one = [out](fsk_handle* w) -> int {
        fsk_handle* h = w;
        CU(cudaSetDevice(h->device));
        if (h->te_nr == 0) return FSK_OK;
        CU(cudaMemcpyAsync(out + (size_t)h->te_r0 * h->n_train, h->d_test, sizeof(double) * (size_t)h->te_nr * h->n_train, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return FSK_OK;
    };
    return is_team(h) ? team_run(h, one) : one(h);
}
// device-resident outputs of THIS handle: all rows, or the rows fsk_output_rows names after a sharded finalisation
int fsk_train_kernel_device(fsk_handle* h, void** dev_ptr) { NEED_FINAL(); *dev_ptr = h->d_train; return FSK_OK; }
int fsk_test_kernel_device(fsk_handle* h, void** dev_ptr) { NEED_FINAL(); *dev_ptr = h->d_test; return FSK_OK; }

// The three getters below see the SUM of the partial kernels this handle can reach: its own buffer, or every rank's while
// the peers are mapped (a team always; the processes of a torchrun launch between fsk_set_peer_partials and
// fsk_release_peers).  They stage through a bounded device buffer, so nothing of the triangle's size is allocated.
int fsk_get_kernel_packed(fsk_handle* h, double* out) {
    NEED_FINAL();
    const int64_t rows_per = std::max<int64_t>(1, std::min<int64_t>(h->N, (64LL << 20) / std::max<int64_t>(1, h->N)));
    double* d_stage;
    ALLOC(d_stage, (size_t)rows_per * h->N);
    int rc = FSK_OK;
    for (int64_t i0 = 0; i0 < h->N && rc == FSK_OK; i0 += rows_per) {
        const int64_t nr = std::min(rows_per, h->N - i0);
        const int64_t c0 = i0 * (i0 + 1) / 2, c1 = (i0 + nr) * (i0 + nr + 1) / 2;
        dim3 grid((unsigned)std::min<int64_t>((h->N + 255) / 256, 64), (unsigned)nr);
        if (h->variance_mode) normalise_packed_kernel<double><<<grid, 256, 0, h->stream>>>(make_parts<double>(h, h->d_Kf), h->d_diag, i0, d_stage);
        else normalise_packed_kernel<unsigned long long><<<grid, 256, 0, h->stream>>>(make_parts<unsigned long long>(h, h->d_Kint), h->d_diag, i0, d_stage);
        h->launches++;
        cudaError_t e = cudaMemcpyAsync(out + c0, d_stage, sizeof(double) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, FSK_ECUDA, "packed kernel copy failed: %s", cudaGetErrorString(e));
    }
    cached_free(d_stage);
    return rc;
}

}  // extern "C"
namespace {
template <typename T, typename O>
int get_summed(fsk_handle* h, const T* own, O* out) {
    CU(cudaSetDevice(h->device));
    if (h->peer_parts.empty() && std::is_same<T, O>::value) {
        CU(cudaMemcpyAsync(out, own, sizeof(O) * (size_t)h->n_pairs, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return FSK_OK;
    }
    const int64_t chunk = std::min<int64_t>(h->n_pairs, 32LL << 20);
    O* d_stage;
    ALLOC(d_stage, (size_t)chunk);
    int rc = FSK_OK;
    for (int64_t c0 = 0; c0 < h->n_pairs && rc == FSK_OK; c0 += chunk) {
        const int64_t n = std::min(chunk, h->n_pairs - c0);
        sum_parts_kernel<T, O><<<592, 256, 0, h->stream>>>(make_parts<T>(h, own), c0, n, d_stage);
        h->launches++;
        cudaError_t e = cudaMemcpyAsync(out + c0, d_stage, sizeof(O) * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, FSK_ECUDA, "unnormalised kernel copy failed: %s", cudaGetErrorString(e));
    }
    cached_free(d_stage);
    return rc;
}
}  // namespace
extern "C" {

int fsk_get_unnormalised_i64(fsk_handle* h, int64_t* out) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (h->variance_mode) return fail(h, FSK_ESTATE, "the integer kernel exists only in the exact and skip_variance modes");
    if (is_team(h) && h->peer_parts.empty()) { int rc = team_link_peers(h); if (rc) return rc; }
    return get_summed<unsigned long long, long long>(h, h->d_Kint, (long long*)out);
}
int fsk_get_unnormalised_f64(fsk_handle* h, double* out) {
    if (!h->uploaded) return fail(h, FSK_ESTATE, "nothing uploaded");
    if (is_team(h) && h->peer_parts.empty()) { int rc = team_link_peers(h); if (rc) return rc; }
    if (h->variance_mode) return get_summed<double, double>(h, h->d_Kf, out);
    return get_summed<unsigned long long, double>(h, h->d_Kint, out);
}

/* pinned host memory for inputs and outputs (the getters and fsk_compute then run at PCIe speed, and -- in a team -- every
 * GPU copies its rows over its own link at the same time); no torch needed */
}  // extern "C"
namespace {
// Pages of the big host buffers are interleaved over the NUMA nodes while they are allocated / pinned: eight GPUs copying
// their rows at once then write to the memory of both sockets instead of one (set_mempolicy(MPOL_INTERLEAVE); a no-op where
// the call is not permitted or there is one node).
struct Interleave {
    bool on = false;
    Interleave() {
#if defined(__linux__) && defined(SYS_set_mempolicy)
        unsigned long mask[16];
        memset(mask, 0, sizeof mask);
        int nodes = 0;
        for (int n = 0; n < 1024; ++n) {
            char path[64];
            snprintf(path, sizeof path, "/sys/devices/system/node/node%d", n);
            if (access(path, F_OK) != 0) break;
            mask[n / (8 * sizeof(unsigned long))] |= 1ul << (n % (8 * sizeof(unsigned long)));
            ++nodes;
        }
        if (nodes > 1) on = syscall(SYS_set_mempolicy, 3 /* MPOL_INTERLEAVE */, mask, (unsigned long)(nodes + 1)) == 0;
        if (getenv("FSK_TRACE")) fprintf(stderr, "[fsk] host buffer: %d NUMA node(s), interleave %s\n", nodes, on ? "on" : "off");
#endif
    }
    ~Interleave() {
#if defined(__linux__) && defined(SYS_set_mempolicy)
        if (on) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
    }
};
}  // namespace
extern "C" {

int fsk_host_alloc(void** out, size_t bytes) {
    if (!out) return FSK_EINVAL;
    *out = nullptr;
    Interleave il;
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("cudaHostAlloc failed: ") + cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? FSK_ENOMEM : FSK_ECUDA; }
    return FSK_OK;
}
int fsk_host_free(void* p) {
    if (p) cudaFreeHost(p);
    return FSK_OK;
}
/* make an existing host range (e.g. a shared-memory mapping all ranks write their rows into) DMA-able */
int fsk_host_register(void* p, size_t bytes) {
    Interleave il;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); g_create_error = std::string("cudaHostRegister failed: ") + cudaGetErrorString(e); return FSK_ECUDA; }
    cudaGetLastError();
    return FSK_OK;
}
int fsk_host_unregister(void* p) {
    cudaHostUnregister(p);
    cudaGetLastError();
    return FSK_OK;
}
int fsk_selftest_division(int device, uint64_t seed, uint64_t n, uint64_t* mismatches) {
    if (!mismatches) return FSK_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); g_create_error = "no usable CUDA device"; return FSK_ECUDA; }
    unsigned long long* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof *d) != cudaSuccess) { cudaGetLastError(); return FSK_ENOMEM; }
    cudaMemset(d, 0, sizeof *d);
    division_selftest_kernel<<<148 * 8, 256>>>(seed, n, d);
    unsigned long long hm = 0;
    const cudaError_t e = cudaMemcpy(&hm, d, sizeof hm, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("division self-test failed: ") + cudaGetErrorString(e); return FSK_ECUDA; }
    *mismatches = hm;
    return FSK_OK;
}
int fsk_trim_cache(void) {
    {
        std::lock_guard<std::mutex> lk(g_ipc.mu);
        for (auto& kv : g_ipc.open) { cudaSetDevice(kv.first.first); cudaIpcCloseMemHandle(kv.second); }
        g_ipc.open.clear();
    }
    std::lock_guard<std::mutex> lk(g_cache.mu);
    g_cache.trim(-1, true);
    g_cache.exported.clear();
    return FSK_OK;
}

int fsk_get_stdevs(fsk_handle* h, double* out, int64_t cap, int64_t* n) {
    if (n) *n = (int64_t)h->stdevs.size();
    for (int64_t i = 0; out && i < cap && i < (int64_t)h->stdevs.size(); ++i) out[i] = h->stdevs[(size_t)i];
    return FSK_OK;
}

int fsk_get_queue(fsk_handle* h, int32_t* out, int64_t cap, int64_t* n) {
    if (!h->uploaded) {   // queue of a handle that has seen no data yet: build it now (depends on g, m, seed only)
        int rc = build_queue(h);
        if (rc) return rc;
    }
    if (n) *n = (int64_t)h->queue.size();
    for (int64_t i = 0; out && i < cap && i < (int64_t)h->queue.size(); ++i) out[i] = h->queue[(size_t)i];
    return FSK_OK;
}

int fsk_get_shard_work(fsk_handle* h, int32_t* out, int64_t cap, int64_t* n) {
    if (!h->uploaded) {
        int rc = build_queue(h);
        if (rc) return rc;
    }
    std::vector<int32_t> mine;
    shard_work(h, mine);
    if (n) *n = (int64_t)mine.size();
    for (int64_t i = 0; out && i < cap && i < (int64_t)mine.size(); ++i) out[i] = mine[(size_t)i];
    return FSK_OK;
}

int fsk_save_kernel(fsk_handle* h, const char* path) {
    NEED_FINAL();
    if (!path || !*path) return FSK_OK;   // fastsk.cpp:226: empty file name is a no-op
    FILE* f = fopen(path, "w");
    if (!f) return fail(h, FSK_EINVAL, "cannot open %s for writing", path);
    // blocks of full rows of the square matrix, normalised on the device: nothing of the triangle's size on the host
    const int64_t rows_per = std::max<int64_t>(1, std::min<int64_t>(h->N, (32LL << 20) / std::max<int64_t>(1, h->N)));
    std::vector<double> rows((size_t)rows_per * h->N);
    double* d_stage = nullptr;
    int rc = dev_alloc(h, &d_stage, (size_t)rows_per * h->N);
    for (int64_t i0 = 0; i0 < h->N && rc == FSK_OK; i0 += rows_per) {
        const int64_t nr = std::min(rows_per, h->N - i0);
        dim3 grid((unsigned)((h->N + 31) / 32), (unsigned)((nr + 31) / 32));
        const dim3 grid_t(grid.y, grid.x);
        if (h->variance_mode) {
            normalise_block_kernel<double><<<grid, 256, 0, h->stream>>>(make_parts<double>(h, h->d_Kf), h->d_diag, i0, nr, h->N, d_stage, 0);
            normalise_block_kernel<double><<<grid_t, 256, 0, h->stream>>>(make_parts<double>(h, h->d_Kf), h->d_diag, i0, nr, h->N, d_stage, 1);
        } else {
            normalise_block_kernel<unsigned long long><<<grid, 256, 0, h->stream>>>(make_parts<unsigned long long>(h, h->d_Kint), h->d_diag, i0, nr, h->N, d_stage, 0);
            normalise_block_kernel<unsigned long long><<<grid_t, 256, 0, h->stream>>>(make_parts<unsigned long long>(h, h->d_Kint), h->d_diag, i0, nr, h->N, d_stage, 1);
        }
        h->launches += 2;
        cudaError_t e = cudaMemcpyAsync(rows.data(), d_stage, sizeof(double) * (size_t)nr * h->N, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { rc = fail(h, FSK_ECUDA, "kernel rows copy failed: %s", cudaGetErrorString(e)); break; }
        for (int64_t i = 0; i < nr; ++i) {
            for (int64_t j = 0; j < h->N; ++j) fprintf(f, "%d:%e ", (int)(j + 1), rows[(size_t)(i * h->N + j)]);
            fprintf(f, "\n");
        }
    }
    if (d_stage) cached_free(d_stage);
    fclose(f);
    return rc;
}

int fsk_get_stats(fsk_handle* h, fsk_stats* out) {
    if (!out) return fail(h, FSK_EINVAL, "out is NULL");
    memset(out, 0, sizeof *out);
    out->n_seq = h->N; out->nfeat = h->nfeat; out->n_pairs = h->n_pairs; out->n_combos_total = h->ncomb;
    out->combos_done = h->combos_done;
    out->kernel_launches = h->launches;
    out->key_bits = h->keybits; out->id_bits = h->idbits; out->record_bytes = h->rec_bytes; out->sort_passes = h->plan.npass;
    out->alphabet = h->A; out->bits_per_char = h->b; out->batch = h->B; out->acc_bytes = 8;
    out->acc_path = h->dense_path ? 3 : (h->rows_path ? 2 : 1);
    out->heavy_tau = (int32_t)h->heavy_tau;
    if (h->uploaded) {
        CU(cudaSetDevice(h->device));
        resolve_spans(h);
        unsigned long long c[4] = {0, 0, 0, 0};
        CU(cudaMemcpy(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost));
        out->entries = (int64_t)c[0]; out->runs = (int64_t)c[1]; out->pair_updates = (int64_t)c[2]; out->heavy_runs = (int64_t)c[3];
    }
    out->ms_pack = h->ms[PC_PACK]; out->ms_sort = h->ms[PC_SORT]; out->ms_segment = h->ms[PC_SEGMENT];
    out->ms_accumulate = h->ms[PC_ACCUMULATE]; out->ms_welford = h->ms[PC_WELFORD]; out->ms_normalise = h->ms[PC_NORMALISE];
    for (int i = 0; i < PC_COUNT; ++i) out->ms_total += h->ms[i];
    out->n_devices = is_team(h) ? (int32_t)h->team.size() : 1;
    out->seg_mode = (h->dir_mode ? 1 : (h->fused_seg ? 2 : 0)) + (h->lean_seg ? 4 : 0);
    out->dense_mode = h->dense_path ? (h->dense_u8 ? 2 : 1) + (h->wf_regs ? 4 : 0) : 0;
    // a team: counts add up over the members, times are the slowest member's
    for (size_t i = 1; !h->leader && i < h->team.size(); ++i) {
        fsk_stats o;
        int rc = fsk_get_stats(h->team[i], &o);
        if (rc) { h->err = h->team[i]->err; return rc; }
        out->combos_done += o.combos_done; out->kernel_launches += o.kernel_launches; out->entries += o.entries;
        out->runs += o.runs; out->pair_updates += o.pair_updates; out->heavy_runs += o.heavy_runs;
        out->ms_pack = std::max(out->ms_pack, o.ms_pack); out->ms_sort = std::max(out->ms_sort, o.ms_sort);
        out->ms_segment = std::max(out->ms_segment, o.ms_segment); out->ms_accumulate = std::max(out->ms_accumulate, o.ms_accumulate);
        out->ms_welford = std::max(out->ms_welford, o.ms_welford); out->ms_normalise = std::max(out->ms_normalise, o.ms_normalise);
        out->ms_total = std::max(out->ms_total, o.ms_total);
    }
    if (h->stream) cudaSetDevice(h->device);
    return FSK_OK;
}

}  // extern "C"
