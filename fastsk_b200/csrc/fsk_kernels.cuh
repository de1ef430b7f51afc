// Device kernels of the gapped k-mer kernel-matrix build (sm_100a).
//
// Per batch of combinations ("slots") the pipeline is
//   pack_hist   g-mer word -> kept characters packed into a key, tagged with the sequence id;
//               digit histograms of every radix pass in the same sweep
//               (replaces the column gather of fastsk_kernel.cpp:224-228)
//   onesweep    stable LSD radix sort over the key bits only, one kernel per 8-bit digit:
//               warp match ranking + decoupled look-back (replaces cntsrtna, shared.cpp:156-191,
//               and the gather of sorted features, fastsk_kernel.cpp:233-238)
//   segment     run boundaries of the sorted records -> one task (run start, prefix length) per record,
//               filed under the record's sequence (shared.cpp:280-315)
//   accumulate  K[i][j] += c_i * c_j on the packed lower triangle (shared.cpp:316-327), one row of K
//               per CTA in shared memory
// plus normalise (fastsk_kernel.cpp:96-103) and the Welford / variance pass
// (fastsk_kernel.cpp:108-143).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace fsk {

constexpr int MAX_K = 32;        // kept positions per combination
constexpr int MAX_BATCH = 384;   // combinations per launch group (BatchSpec travels in kernel-parameter space: 25 KB of the 32 KB)
constexpr int MAX_PASS = 8;      // 64 key bits / 8
constexpr int RADIX = 256;
constexpr int SORT_THREADS = 256;
constexpr int PACK_ITEMS = 64;   // windows per thread of pack_hist_kernel

// The kept characters of a combination, as maximal stretches of consecutive kept positions inside one g-mer
// word: seg = source bit position (word * 64 + shift) | (width in bits - 1) << 7.  The stretches are laid into
// the key from bit 0 upwards; any injective packing gives the same runs (only equality of k-mers matters).
struct BatchSpec {               // by value in kernel parameter space
    uint16_t seg[MAX_BATCH][MAX_K];
    uint8_t nseg[MAX_BATCH];
};
struct SortPlan {
    int npass;
    uint8_t shift[MAX_PASS];     // digit position relative to key bit 0
    uint8_t bits[MAX_PASS];
};

// Variance mode with the Welford step fused into the accumulate's flush (fastsk_kernel.cpp:108-143): the running mean of
// the virtual stream that owns each slot, its iteration number, and where the per-CTA sums of delta * delta2 go.  Lives
// in device memory; the host rewrites it before every round.
struct WelfordSpec {
    // per group = one virtual stream of this round: its slots [slot0, slot0 + depth) are CONSECUTIVE iterations of the stream
    // (speculated past a possible stop; the host rolls a stream back from khat_in when its stop rule fires inside the round)
    const double* khat_in[MAX_BATCH];   // the stream's running mean when the round starts
    double* khat_out[MAX_BATCH];        // ... and where the round leaves it (the same buffer when depth is 1)
    int32_t iter0[MAX_BATCH];           // iteration number of the group's first slot
    uint16_t slot0[MAX_BATCH], depth[MAX_BATCH];
    double* sums;                       // [slot][sums_stride]
    uint32_t sums_stride;
    int64_t n_train;                    // rows below n_train are the train x train block (the first n_train_pairs cells)
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v));
}

// a / b, correctly rounded, for the Welford step's division by the iteration number (fastsk_kernel.cpp:124: K_hat += d / iter):
// with rb = RN(1 / b) computed once per slot, q0 = RN(a * rb) is within one ulp of the quotient, the remainder a - q0 * b is
// exact in one FMA, and q0 + rem * rb rounds to RN(a / b) (Markstein's theorem; b is a small positive integer, so its
// significand is never the all-ones exception).  ~4 instructions instead of the ~45 of the general division routine, same
// bits: fsk_selftest_division compares the two on the device.
__device__ __forceinline__ double div_by_iter(double a, double b, double rb) {
    const double q0 = __dmul_rn(a, rb);
    const double rem = __fma_rn(-q0, b, a);
    return __fma_rn(rem, rb, q0);
}
__global__ void division_selftest_kernel(uint64_t seed, uint64_t n, unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;                 // splitmix64
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
        uint64_t y = (x + 0x9E3779B97F4A7C15ull); y = (y ^ (y >> 30)) * 0xBF58476D1CE4E5B9ull; y = (y ^ (y >> 27)) * 0x94D049BB133111EBull; y ^= y >> 31;
        const double b = (double)(1 + (x % ((i & 7) == 0 ? 2000000000ull : 4096ull)));       // iteration numbers
        // numerators: integers (counts minus a mean), fractions of every magnitude, both signs
        double a = (double)(int64_t)(y >> ((i >> 3) & 31)) * (((i >> 8) & 1) ? 1.0 : 1.0 / 1048576.0);
        if ((i >> 9) & 1) a = -a;
        if ((i & 0xff) == 0) a = __longlong_as_double((long long)((y & 0x800fffffffffffffull) | ((uint64_t)(1023 + (int)(x % 200) - 100) << 52)));
        const double rb = __ddiv_rn(1.0, b);
        if (__double_as_longlong(div_by_iter(a, b, rb)) != __double_as_longlong(__ddiv_rn(a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// exclusive scan of one value per thread over a 256-thread block; `total` gets the block sum
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* warp_sums /* >= 8 */, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t s = warp_sums[w];
        if (w < warp) wbase += s;
        tot += s;
    }
    total = tot;
    __syncthreads();
    return wbase + inc - v;
}

// ------------------------------------------------------------------------------------------
// g-mer words: every length-g window of every sequence packed b bits per character, once per
// upload.  Windows are numbered in (sequence, position) order like extractFeatures
// (shared.cpp:55-91); nothing g-times-larger is ever materialised.
template <typename GwT, int NW>
__global__ void build_gwords_kernel(const uint8_t* __restrict__ codes, const int64_t* __restrict__ offsets,
                                    const int64_t* __restrict__ woffs, int64_t nseq, int g, int b, int cpw,
                                    GwT* __restrict__ gw0, uint64_t* __restrict__ gw1, uint32_t* __restrict__ wseq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t s = warp; s < nseq; s += nwarps) {
        const uint8_t* c = codes + offsets[s];
        const int64_t nw = offsets[s + 1] - offsets[s] - g + 1;
        const int64_t w0 = woffs[s];
        for (int64_t p = lane; p < nw; p += 32) {
            uint64_t lo = 0, hi = 0;
            for (int t = 0; t < g; ++t) {
                const uint64_t ch = c[p + t];
                if (NW == 1 || t < cpw) lo |= ch << (t * b);
                else hi |= ch << ((t - cpw) * b);
            }
            gw0[w0 + p] = (GwT)lo;
            if (NW == 2) gw1[w0 + p] = hi;
            wseq[w0 + p] = (uint32_t)s;
        }
    }
}

// ------------------------------------------------------------------------------------------
// pack + histogram.  grid = (slots, tiles): CTAs launched together read the same windows for different
// combinations, so the g-mer words come from HBM once per batch and from L2 for the other slots.
//
// A stretch of kept characters moves from bit `src` of the g-mer word to bit `dst` of the key: one rotate and
// one masked OR (key |= rotr(word, src - dst) & (mask << dst)).  A thread packs PACK_SUB windows at a time: the
// stretch and digit descriptors are read from parameter space once per sub-batch (uniform loads), the inner
// loops run on registers only, and PACK_SUB loads per array are in flight.
constexpr int PACK_SUB = 8;

template <typename KeyT>
__device__ __forceinline__ KeyT rotr_key(KeyT x, uint32_t r) {
    if (sizeof(KeyT) == 4) return (KeyT)__funnelshift_r((uint32_t)x, (uint32_t)x, r);
    r &= 63u;
    return r ? (KeyT)(((uint64_t)x >> r) | ((uint64_t)x << (64u - r))) : x;
}

template <typename RecT, bool KV, typename GwT, int NW>
__global__ void __launch_bounds__(256)
pack_hist_kernel(const GwT* __restrict__ gw0, const uint64_t* __restrict__ gw1, const uint32_t* __restrict__ wseq,
                 uint32_t nfeat, RecT* __restrict__ rec, uint32_t* __restrict__ val, uint32_t* __restrict__ ghist,
                 const __grid_constant__ BatchSpec spec, const __grid_constant__ SortPlan plan, int idbits,
                 uint16_t* __restrict__ wkey /* directory form of the segmentation: the key of every window, in window order */) {
    // a 32-bit g-mer word holds a key of at most 32 bits: keep the arithmetic in 32 bits then
    using KeyT = typename std::conditional<sizeof(GwT) == 4, uint32_t, uint64_t>::type;
    constexpr uint32_t KB = sizeof(KeyT) * 8;
    __shared__ uint32_t sh[MAX_PASS * RADIX];
    const int slot = blockIdx.x;
    const int nseg = spec.nseg[slot];
    const int npass = plan.npass;
    for (int i = threadIdx.x; i < npass * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const size_t sbase = (size_t)slot * nfeat;
    // a CTA packs PACK_ITEMS x 256 consecutive windows: its npass x 256 histogram counters go to the slot's global
    // histogram once, so few CTAs per slot keep the same-address atomics on those few lines off the critical path
    const uint32_t tile0 = blockIdx.y * (256 * PACK_ITEMS);
    for (int it = 0; it < PACK_ITEMS; it += PACK_SUB) {
        const uint32_t w0 = tile0 + it * 256 + threadIdx.x;
        if (tile0 + it * 256 >= nfeat) break;
        KeyT lo[PACK_SUB], hi[NW == 2 ? PACK_SUB : 1], key[PACK_SUB];
        uint32_t seq[PACK_SUB];
#pragma unroll
        for (int i = 0; i < PACK_SUB; ++i) {
            const uint32_t w = w0 + i * 256;
            lo[i] = 0;
            seq[i] = 0;
            if (NW == 2) hi[i] = 0;
            if (w < nfeat) {
                lo[i] = (KeyT)gw0[w];
                if (NW == 2) hi[i] = (KeyT)gw1[w];
                seq[i] = wseq[w];
            }
            key[i] = 0;
        }
        uint32_t dst = 0;
        for (int j = 0; j < nseg; ++j) {
            const uint32_t e = spec.seg[slot][j];
            const uint32_t src = e & 63u, width = (e >> 7) + 1u;
            const KeyT m = (KeyT)((KeyT) ~(KeyT)0 >> (KB - width)) << dst;
            const uint32_t r = (src - dst) & (KB - 1u);
            if (NW == 2 && (e & 64u)) {
#pragma unroll
                for (int i = 0; i < PACK_SUB; ++i) key[i] |= rotr_key<KeyT>(hi[i], r) & m;
            } else {
#pragma unroll
                for (int i = 0; i < PACK_SUB; ++i) key[i] |= rotr_key<KeyT>(lo[i], r) & m;
            }
            dst += width;
        }
#pragma unroll
        for (int i = 0; i < PACK_SUB; ++i) {
            const uint32_t w = w0 + i * 256;
            if (w < nfeat) {
                if (KV) {
                    rec[sbase + w] = (RecT)key[i];
                    val[sbase + w] = seq[i];
                } else {
                    rec[sbase + w] = ((RecT)key[i] << idbits) | (RecT)seq[i];
                }
                if (wkey) wkey[sbase + w] = (uint16_t)key[i];
            }
        }
        for (int p = 0; p < npass; ++p) {
            const uint32_t dsh = plan.shift[p], dmask = (1u << plan.bits[p]) - 1u;
            uint32_t* hp = sh + p * RADIX;
#pragma unroll
            for (int i = 0; i < PACK_SUB; ++i)
                if (w0 + i * 256 < nfeat) atomicAdd(&hp[(uint32_t)(key[i] >> dsh) & dmask], 1u);
        }
    }
    __syncthreads();
    uint32_t* gh = ghist + (size_t)slot * MAX_PASS * RADIX;
    for (int i = threadIdx.x; i < npass * RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&gh[i], sh[i]);
}

// ------------------------------------------------------------------------------------------
// One LSD pass (onesweep): a single read and a single write of the records.  Tiles take a
// ticket so that every predecessor of a running tile has already started (forward progress of
// the look-back).  Stable: keys keep their input order within a digit, so sequence ids stay
// ascending inside every run of equal k-mers (the property countAndUpdateTri relies on).
// status word = flag (2 bits: 1 = tile aggregate, 2 = inclusive prefix) | count (30 bits).
template <typename RecT, bool KV, int ITEMS, bool OPTIMISTIC>
__global__ void __launch_bounds__(SORT_THREADS, (sizeof(RecT) == 4 ? 4 : 2))
onesweep_kernel(const RecT* __restrict__ in, RecT* __restrict__ out, const uint32_t* __restrict__ vin,
                uint32_t* __restrict__ vout, uint32_t n, uint32_t tiles_per_slot, uint32_t nslots, int shift, int bits,
                const uint32_t* __restrict__ ghist /* [slot][MAX_PASS][RADIX], pre-offset to this pass */,
                uint32_t* __restrict__ status /* [slot][tile][RADIX] of this pass */, uint32_t* __restrict__ ticket) {
    constexpr int TILE = SORT_THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RecT* skeys = reinterpret_cast<RecT*>(smem_raw);
    uint32_t* svals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(RecT) * TILE);
    uint32_t* warp_hist = svals + (KV ? TILE : 0);   // [8][RADIX]
    uint32_t* scatter_base = warp_hist + 8 * RADIX;   // [RADIX]
    uint32_t* warp_sums = scatter_base + RADIX;       // [8]
    uint32_t* match_mask = warp_sums + 8;             // [8][RADIX]
    __shared__ uint32_t s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    for (int i = tid; i < 8 * RADIX; i += SORT_THREADS) { warp_hist[i] = 0; if (!OPTIMISTIC) match_mask[i] = 0; }
    __syncthreads();
    // tickets deal the slots round-robin: the tiles in flight at any time are spread over all slots of the batch, so a
    // tile's look-back crosses only the few running tiles of its own slot (not every resident CTA of the chip)
    const uint32_t tile = s_ticket / nslots;
    const uint32_t slot = s_ticket - tile * nslots;
    const size_t sbase = (size_t)slot * n;
    const uint32_t tile0 = tile * TILE;
    const uint32_t dmask = (1u << bits) - 1;

    RecT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rank2[(ITEMS + 1) / 2];   // two 16-bit ranks per register (a rank is < TILE <= 4096)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t idx = tile0 + warp * (32 * ITEMS) + j * 32 + lane;
        key[j] = 0;
        val[j] = 0;
        if (idx < n) {
            key[j] = in[sbase + idx];
            if (KV) val[j] = vin[sbase + idx];
        }
    }
    // In-warp stable ranking of the digits.
    //   OPTIMISTIC: one shared-memory atomicAdd with return per key.  Lanes of one instruction that hit the same counter
    //   are served in lane order on this hardware (tools/rank_bench.cu: 0 mismatches in 74 M keys, 1.75x faster than the
    //   next best) but nothing documents that, so segment_kernel verifies that the final records are non-decreasing -- which,
    //   the scatter being a permutation, holds iff every pass was stable -- and the host repeats the build with the safe
    //   ranking if the check ever fails.
    //   safe: lanes holding the same digit find each other through a per-warp match mask (atomicOr of the lane bit, read
    //   back, cleared by the lowest lane); 2.8x faster than match.any, whose MATCH instruction saturates the ADU pipe.
    const uint32_t lane_lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t idx = tile0 + warp * (32 * ITEMS) + j * 32 + lane;
        const bool valid = idx < n;
        const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
        uint32_t* wh = warp_hist + warp * RADIX + d;
        uint32_t rk;
        if (OPTIMISTIC) {
            rk = valid ? atomicAdd(wh, 1u) : 0u;
        } else {
            uint32_t* mm = match_mask + warp * RADIX + d;
            if (valid) atomicOr(mm, 1u << lane);
            __syncwarp();
            uint32_t peers = 0, prev = 0;
            if (valid) { peers = *mm; prev = *wh; }
            __syncwarp();
            const uint32_t lower = peers & lane_lt;
            if (valid && lower == 0) { *wh = prev + __popc(peers); *mm = 0; }
            __syncwarp();
            rk = prev + __popc(lower);
        }
        if (j & 1) rank2[j >> 1] |= rk << 16;
        else rank2[j >> 1] = rk;
    }
    __syncthreads();

    // thread d owns digit d: tile total, published at once as the tile's aggregate
    uint32_t cw[8], total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        cw[w] = warp_hist[w * RADIX + tid];
        total += cw[w];
    }
    uint32_t* my_status = status + ((size_t)slot * tiles_per_slot + tile) * RADIX + tid;
    st_volatile_u32(my_status, total | (tile == 0 ? 0x80000000u : 0x40000000u));

    uint32_t dummy;
    const uint32_t gexcl = block_excl_scan_256(ghist[(size_t)slot * MAX_PASS * RADIX + tid], warp_sums, dummy);
    const uint32_t dstart = block_excl_scan_256(total, warp_sums, dummy);
    // position of a key inside the tile = (start of its digit + keys of the digit in earlier warps) + rank: one lookup
    {
        uint32_t run = dstart;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            warp_hist[w * RADIX + tid] = run;
            run += cw[w];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t idx = tile0 + warp * (32 * ITEMS) + j * 32 + lane;
        if (idx < n) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
            const uint32_t pos = warp_hist[warp * RADIX + d] + ((rank2[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
            skeys[pos] = key[j];
            if (KV) svals[pos] = val[j];
        }
    }

    // decoupled look-back (after the local scatter was issued, so its latency overlaps), LOOK predecessors per step
    // with independent loads in flight
    uint32_t excl = 0;
    if (tile > 0) {
        constexpr int LOOK = 8;
        int64_t pt = (int64_t)tile - 1;
        bool done = false;
        while (!done) {
            uint32_t sv[LOOK];
#pragma unroll
            for (int i = 0; i < LOOK; ++i) sv[i] = (pt - i >= 0) ? ld_volatile_u32(my_status - (size_t)(tile - (pt - i)) * RADIX) : 0x80000000u;
#pragma unroll
            for (int i = 0; i < LOOK; ++i) {
                if (done) break;
                uint32_t v = sv[i];
                while ((v >> 30) == 0) v = ld_volatile_u32(my_status - (size_t)(tile - (pt - i)) * RADIX);
                excl += v & 0x3fffffffu;
                if (v >> 31) done = true;
            }
            pt -= LOOK;
        }
        st_volatile_u32(my_status, (total + excl) | 0x80000000u);
    }
    scatter_base[tid] = gexcl + excl - dstart;
    __syncthreads();
    const uint32_t count = min((uint32_t)TILE, n - tile0);
    for (uint32_t i = tid; i < count; i += SORT_THREADS) {
        const RecT kk = skeys[i];
        const uint32_t d = (uint32_t)(kk >> shift) & dmask;
        const uint32_t dst = scatter_base[d] + i;
        out[sbase + dst] = kk;
        if (KV) vout[sbase + dst] = svals[i];
    }
}

// ------------------------------------------------------------------------------------------
// segmentation of the sorted records (shared.cpp:280-315 without materialising the per-run count
// table).
//
// Every sorted record i is one TASK of its own sequence b = seq(i): "add 1 to K[b][seq(j)] for every
// record j of the same run from the run's first record rs(i) to the last record ge(i) of b's own group".
// Summed over the c_b records of b in the run this gives K[b][a] += c_a * c_b for every a <= b of the run
// and K[b][b] += c_b^2 -- exactly the update of shared.cpp:316-327 -- with no per-(k-mer, sequence)
// compaction pass: the sorted records themselves are the entry list (ids ascend inside a run because the
// sort is stable).
//
// Outputs:
//   ids   the sequence id of every sorted record (IdT = u16 when N <= 65000, else u32), with every RUN
//         ALIGNED: run q starts at X_q = sum over earlier runs of their length rounded up to the alignment
//         (pad_mask + 1 ids: one 16-byte unit, or one 128-byte line when that costs little memory).  The cells
//         between a run's end and the next run's start keep the 0xFF.. fill the host put there.  A task's
//         prefix then starts on a unit boundary, and whatever follows its last id inside the last unit is
//         either a larger id of the same run or fill: the accumulate needs no masks, only min(id, b + 1).
//   task  filed by row: task[woff[b] + fill[b]++] = (X_run >> unit_shift, ge - rs + 1); a sequence has exactly
//         as many tasks as windows.
// X(i) = exclusive prefix sum, over sorted positions e < i, of [e is the last record of its run] x
// roundup(run length): a warp scan, a CTA scan, and a decoupled look-back over the tiles of the slot (tiles
// take tickets, slot-major, so that the task array being filled stays L2-resident).
// A warp owns SEG_ROWS x 32 consecutive records; run heads / group tails are warp ballots, so the last
// head at or before a record and the first tail at or after it are bit scans.
constexpr int SEG_THREADS = 256;
constexpr int SEG_ROWS_DEFAULT = 12;   // measured on C4 (ms per 384 combinations): 8: 76.9, 10: 69.5, 12: 64.8, 16: 66.4
__host__ __device__ constexpr int seg_tile_records(int rows) { return (SEG_THREADS / 32) * rows * 32; }

// fill[slot][b] = woff[b]: where the next task of sequence b goes
__global__ void init_fill_kernel(uint32_t* __restrict__ fill, const uint32_t* __restrict__ woff, uint32_t nseq) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nseq) fill[(size_t)blockIdx.y * nseq + b] = woff[b];
}

template <typename RecT, bool KV>
struct RecOps {
    static __device__ __forceinline__ bool same_key(RecT a, RecT b, int idbits) { return KV ? a == b : ((a ^ b) >> idbits) == 0; }
    static __device__ __forceinline__ bool key_less(RecT a, RecT b, int idbits) { return KV ? a < b : (a >> idbits) < (b >> idbits); }
};

//
// DIRECTORY FORM (DIR = true; key spaces of at most 2^16 k-mers, where every k-mer has a slot of its own): nothing is filed
// per record -- no atomic, no scattered task store.  The rows of K are cut into blocks of 2^bshift sequences, and the LAST
// record of every (run, block) group writes ONE entry tdir[block][key] = (run start, records of the run up to the end of
// this block).  The row CTA of sequence b then finds the task of each of its windows by itself: key of the window
// (pack_hist_kernel's wkey) -> tdir[b >> bshift][key].  The prefix it reads may run past b's own group to the end of b's
// block; those ids are larger than b and fall into the dump words like every other id beyond the row.  Entries per slot:
// (run, block) pairs, ~0.22 per record on the headline workload, against one atomic + one 8-byte scatter per record.
template <typename RecT, bool KV, typename IdT, bool HEAVY = false, int ROWS = SEG_ROWS_DEFAULT, int MINB = (sizeof(RecT) == 4 ? 3 : 2),
          bool DIR = false, bool DIRSTATS = false>
__global__ void __launch_bounds__(SEG_THREADS, MINB)
segment_kernel(const RecT* __restrict__ rec, const uint32_t* __restrict__ val, uint32_t n, uint32_t tiles_per_slot,
               size_t ids_stride, int idbits, uint32_t nseq, int unit_shift, uint32_t pad_mask, uint32_t* __restrict__ fill,
               IdT* __restrict__ ids, uint2* __restrict__ task, uint32_t* __restrict__ scan_status /* [slot][tile] */,
               uint32_t* __restrict__ ticket, uint32_t* __restrict__ unsorted_flag,
               unsigned long long* __restrict__ stat_counters,
               uint32_t heavy_tau /* 0 = off: runs longer than this may leave the sparse path */, uint32_t* __restrict__ heavy_count,
               uint2* __restrict__ heavy_list /* (slot, first sorted record) of every heavy run of the batch */, uint32_t heavy_cap,
               uint32_t* __restrict__ heavy_bits /* [slot][units / 32]: bit set = the run starting at that id unit is in the list */,
               size_t heavy_bits_stride, uint2* __restrict__ tdir = nullptr /* [slot][block][key] */, int dir_bshift = 0,
               uint32_t dir_nb = 0, int dir_keybits = 0) {
    using Ops = RecOps<RecT, KV>;
    constexpr bool NEED_LEN = !DIR || DIRSTATS;
    constexpr int SEG_ROWS = ROWS;
    constexpr int SEG_WARP_RECS = SEG_ROWS * 32;
    constexpr int SEG_TILE = seg_tile_records(ROWS);
    __shared__ uint32_t s_ticket, s_tile_base, s_warp_tot[SEG_THREADS / 32];
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t slot = s_ticket / tiles_per_slot;
    const uint32_t tile = s_ticket - slot * tiles_per_slot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t sbase = (size_t)slot * n;
    const RecT* __restrict__ R = rec + sbase;
    const uint32_t* __restrict__ V = KV ? val + sbase : nullptr;
    const uint32_t seg0 = tile * SEG_TILE + warp * SEG_WARP_RECS;
    const bool live = seg0 < n;                              // the last tile's trailing warps own no records
    const RecT idmask = KV ? (RecT)0 : (((RecT)1 << idbits) - 1);

    RecT r[SEG_ROWS];
    uint32_t sq[SEG_ROWS];
#pragma unroll
    for (int k = 0; k < SEG_ROWS; ++k) {
        const uint32_t i = seg0 + k * 32 + lane;
        r[k] = 0;
        sq[k] = 0;
        if (i < n) {
            r[k] = R[i];
            sq[k] = KV ? V[i] : (uint32_t)(r[k] & idmask);
        }
    }
    // the records just outside the warp's segment
    RecT r_before = 0, r_after = 0;
    uint32_t s_before = 0, s_after = 0;
    const uint32_t seg_end = live ? min(seg0 + SEG_WARP_RECS, n) : seg0;   // exclusive
    if (live && seg0 > 0) { r_before = R[seg0 - 1]; s_before = KV ? V[seg0 - 1] : (uint32_t)(r_before & idmask); }
    if (live && seg_end < n) { r_after = R[seg_end]; s_after = KV ? V[seg_end] : (uint32_t)(r_after & idmask); }

    uint32_t hm[SEG_ROWS], tm[NEED_LEN ? SEG_ROWS : 1], rtm[SEG_ROWS];   // run-head / group-tail / run-tail ballots (warp-uniform)
    uint32_t btm[DIR ? SEG_ROWS : 1];                     // DIR: last record of its (run, block of sequences) group
    uint32_t n_groups = 0;
#pragma unroll
    for (int k = 0; k < SEG_ROWS; ++k) {
        const uint32_t i = seg0 + k * 32 + lane;
        // previous record of lane 0 is lane 31 of the previous row; next record of lane 31 is lane 0 of the next row
        RecT pr = __shfl_up_sync(0xffffffffu, r[k], 1);
        uint32_t ps = __shfl_up_sync(0xffffffffu, sq[k], 1);
        RecT nr = __shfl_down_sync(0xffffffffu, r[k], 1);
        uint32_t ns = __shfl_down_sync(0xffffffffu, sq[k], 1);
        const RecT pr_row = k > 0 ? __shfl_sync(0xffffffffu, r[k > 0 ? k - 1 : 0], 31) : r_before;
        const uint32_t ps_row = k > 0 ? __shfl_sync(0xffffffffu, sq[k > 0 ? k - 1 : 0], 31) : s_before;
        const RecT nr_row = k < SEG_ROWS - 1 ? __shfl_sync(0xffffffffu, r[k < SEG_ROWS - 1 ? k + 1 : k], 0) : r_after;
        const uint32_t ns_row = k < SEG_ROWS - 1 ? __shfl_sync(0xffffffffu, sq[k < SEG_ROWS - 1 ? k + 1 : k], 0) : s_after;
        if (lane == 0) { pr = pr_row; ps = ps_row; }
        if (lane == 31) { nr = nr_row; ns = ns_row; }
        const bool valid = i < n;
        const bool head = valid && (i == 0 || !Ops::same_key(pr, r[k], idbits));
        const bool ghead = valid && (i == 0 || pr != r[k] || (KV && ps != sq[k]));
        const bool tail = valid && (i + 1 >= n || nr != r[k] || (KV && ns != sq[k]));
        const bool rtail = valid && (i + 1 >= n || !Ops::same_key(nr, r[k], idbits));
        // the sort must have left the records non-decreasing (key, then sequence id): see onesweep_kernel
        if (valid && i > 0 && (KV ? (pr > r[k] || (pr == r[k] && ps > sq[k])) : pr > r[k])) *unsorted_flag = 1u;
        hm[k] = __ballot_sync(0xffffffffu, head);
        if (NEED_LEN) tm[k] = __ballot_sync(0xffffffffu, tail);
        rtm[k] = __ballot_sync(0xffffffffu, rtail);
        if (DIR) btm[k] = __ballot_sync(0xffffffffu, rtail || (valid && (ns >> dir_bshift) != (sq[k] >> dir_bshift)));
        if (NEED_LEN) n_groups += __popc(__ballot_sync(0xffffffffu, ghead));
    }

    // run start of the segment's first record when its run began before the segment: walk back in blocks of 32
    // (typical runs are short), then lower_bound over the sorted records for very long runs
    uint32_t carry_head = seg0;
    if (live && seg0 > 0 && !(hm[0] & 1u)) {
        const RecT r_first = __shfl_sync(0xffffffffu, r[0], 0);
        bool found = false;
        uint32_t p = seg0;                  // records [p, seg0) all have the key of r_first
        for (int step = 0; step < 8 && p > 0 && !found; ++step) {
            const uint32_t j = p - 1 - lane;               // may wrap below 0
            const bool inb = (uint32_t)lane < p;
            const bool differs = inb && !Ops::same_key(R[inb ? j : 0], r_first, idbits);
            const uint32_t dm = __ballot_sync(0xffffffffu, differs);
            if (dm) { p = p - (__ffs(dm) - 1); found = true; }      // first differing record going backwards is at p-1-l
            else p = p > 32 ? p - 32 : 0;
        }
        if (!found && p > 0) {              // lower_bound of the key in [0, p)
            uint32_t lo = 0, hi = p;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (Ops::key_less(R[mid], r_first, idbits)) lo = mid + 1;
                else hi = mid;
            }
            p = lo;
        }
        carry_head = p;
    }

    // first group tail at or after each record, scanning the rows backwards
    uint32_t ge[NEED_LEN ? SEG_ROWS : 1];
    if (NEED_LEN) {
        uint32_t next_tail = 0xffffffffu;   // first tail in the rows after row k (warp-uniform)
        const uint32_t lane_ge = 0xffffffffu << lane;
#pragma unroll
        for (int k = SEG_ROWS - 1; k >= 0; --k) {
            const uint32_t base = seg0 + k * 32;
            const uint32_t m = tm[k] & lane_ge;
            ge[k] = m ? base + (__ffs(m) - 1) : next_tail;
            if (tm[k]) next_tail = base + (__ffs(tm[k]) - 1);
        }
    }
    // run start and prefix length of every record; a group that runs past the warp's segment is followed in
    // global memory (rare)
    unsigned long long updates = 0;
    uint32_t rs[SEG_ROWS], len[NEED_LEN ? SEG_ROWS : 1];
    {
        uint32_t last_head = carry_head;
        const uint32_t lane_le = 0xffffffffu >> (31 - lane);
#pragma unroll
        for (int k = 0; k < SEG_ROWS; ++k) {
            const uint32_t base = seg0 + k * 32;
            const uint32_t m = hm[k] & lane_le;
            rs[k] = m ? base + (31 - __clz(m)) : last_head;
            if (hm[k]) last_head = base + (31 - __clz(hm[k]));
            if (NEED_LEN) len[k] = ge[k] - rs[k] + 1;           // garbage where ge is unknown or the record is out of range: fixed below
        }
        if (NEED_LEN && __any_sync(0xffffffffu, ge[SEG_ROWS - 1] == 0xffffffffu && seg0 + (SEG_ROWS - 1) * 32 + lane < n)) {
#pragma unroll
            for (int k = 0; k < SEG_ROWS; ++k) {
                const uint32_t i = seg0 + k * 32 + lane;
                if (i < n && ge[k] == 0xffffffffu) {
                    uint32_t g = seg_end - 1;
                    while (g + 1 < n && R[g + 1] == r[k] && (!KV || V[g + 1] == sq[k])) ++g;
                    len[k] = g - rs[k] + 1;
                }
            }
        }
    }

    // aligned position of every record's run: exclusive scan of the rounded-up lengths of the runs that END before it
    uint32_t xs[SEG_ROWS];
    uint32_t wtot = 0;                      // warp-uniform running total
#pragma unroll
    for (int k = 0; k < SEG_ROWS; ++k) {
        xs[k] = wtot;
        if (rtm[k]) {                       // warp-uniform: most rows of long runs hold no run end
            const uint32_t i = seg0 + k * 32 + lane;
            const uint32_t c = ((rtm[k] >> lane) & 1u) ? ((i - rs[k] + 1 + pad_mask) & ~pad_mask) : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            xs[k] += inc - c;
            wtot += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    if (lane == 0) s_warp_tot[warp] = wtot;
    __syncthreads();
    uint32_t wbase = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < SEG_THREADS / 32; ++w) {
        const uint32_t t = s_warp_tot[w];
        if (w < warp) wbase += t;
        tile_total += t;
    }
    uint32_t* __restrict__ st = scan_status + (size_t)slot * tiles_per_slot;
    if (threadIdx.x == 0 && tile > 0) st_volatile_u32(st + tile, tile_total | 0x40000000u);

    // file the tasks: all the atomics of half the rows in flight before the first dependent store; the look-back of
    // the tile's base runs under the first half's atomics
    constexpr int HALF = SEG_ROWS / 2;
    uint32_t base = 0;
#pragma unroll
    for (int h0 = 0; h0 < SEG_ROWS; h0 += HALF) {
        uint32_t pos[DIR ? 1 : HALF];            // fill[] starts at woff[seq]: the atomic returns the task's address
        if (!DIR) {
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                const uint32_t i = seg0 + (h0 + k) * 32 + lane;
                pos[k] = 0;
                if (i < n) {
                    pos[k] = atomicAdd(&fill[(size_t)slot * nseq + sq[h0 + k]], 1u);
                }
            }
        }
        if (h0 == 0) {
            if (warp == 0) {
                uint32_t excl = 0;
                if (tile > 0) {
                    int64_t pt = (int64_t)tile - 1;
                    while (true) {
                        const int64_t mine = pt - lane;
                        uint32_t v = 0x80000000u;            // before the slot's first tile: inclusive 0
                        if (mine >= 0) {
                            do { v = ld_volatile_u32(st + mine); } while ((v >> 30) == 0);
                        }
                        const uint32_t inclusive = __ballot_sync(0xffffffffu, (v >> 31) != 0);
                        const bool take = inclusive == 0 || lane <= __ffs(inclusive) - 1;
                        excl += __reduce_add_sync(0xffffffffu, take ? (v & 0x3fffffffu) : 0u);
                        if (inclusive) break;
                        pt -= 32;
                    }
                }
                if (lane == 0) {
                    s_tile_base = excl;
                    st_volatile_u32(st + tile, (excl + tile_total) | 0x80000000u);
                }
            }
            __syncthreads();
            base = s_tile_base + wbase;
        }
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
            const uint32_t i = seg0 + (h0 + k) * 32 + lane;
            if (i < n) {
                const uint32_t X = base + xs[h0 + k];
                ids[(size_t)slot * ids_stride + X + (i - rs[h0 + k])] = (IdT)sq[h0 + k];
                const bool writes = !DIR || ((btm[DIR ? h0 + k : 0] >> lane) & 1u);      // DIR: one entry per (run, block) group
                uint32_t ln = DIR ? i - rs[h0 + k] + 1 : len[NEED_LEN ? h0 + k : 0];
                if (HEAVY && heavy_tau) {                  // (compiled out of the variant launched while the stage sleeps)
                    // a run of more than heavy_tau records is a (nearly) dense column of the count matrix: its d^2/2 updates go to
                    // the tensor-core contraction (heavy_fill_kernel + syrk_tc_kernel) if the batch's list has room, and then the
                    // accumulate skips its tasks.  The records are sorted, so the run is that long iff the record heavy_tau places
                    // after its start has the same key.
                    const uint32_t far = rs[h0 + k] + heavy_tau;
                    if ((writes || i == rs[h0 + k]) && far < n && Ops::same_key(R[far], r[h0 + k], idbits)) {
                        ln |= 0x80000000u;                       // "ask heavy_bits": only the run's first record knows whether the list had room
                        if (i == rs[h0 + k]) {
                            const uint32_t at = atomicAdd(heavy_count, 1u);
                            if (at < heavy_cap) {
                                heavy_list[at] = make_uint2(slot, i);
                                const uint32_t xu = X >> unit_shift;
                                atomicOr(&heavy_bits[(size_t)slot * heavy_bits_stride + (xu >> 5)], 1u << (xu & 31u));
                            }
                        }
                    }
                }
                if (DIR) {
                    if (writes) {
                        const uint32_t key = (uint32_t)(r[h0 + k] >> idbits);
                        tdir[(((size_t)slot * dir_nb + (sq[h0 + k] >> dir_bshift)) << dir_keybits) + key] = make_uint2(X >> unit_shift, ln);
                    }
                    if (DIRSTATS) updates += len[NEED_LEN ? h0 + k : 0];
                } else {
                    task[sbase + pos[k]] = make_uint2(X >> unit_shift, ln);
                    updates += ln & 0x7fffffffu;
                }
            }
        }
    }
    if (NEED_LEN && stat_counters) {
        uint32_t n_runs = 0;
#pragma unroll
        for (int k = 0; k < SEG_ROWS; ++k) n_runs += __popc(hm[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) updates += __shfl_xor_sync(0xffffffffu, updates, o);
        if (lane == 0) {
            atomicAdd(&stat_counters[0], (unsigned long long)n_groups);
            atomicAdd(&stat_counters[1], (unsigned long long)n_runs);
            atomicAdd(&stat_counters[2], updates);
        }
    }
}

// ------------------------------------------------------------------------------------------
// accumulate (shared.cpp:316-327): K[b][ids[j]] += 1 for every task (first unit u, length len) of row b
// and every j in [u * PER, u * PER + len).
//
// Measured on B200 (profiles/r01_atomic_microbench.txt): scattered RED into the 10 GB packed triangle
// runs at 20 G updates/s (DRAM sector read-modify-write), L2-resident at 210 G/s, shared-memory atomics
// at > 800 G/s.  So the update is ordered by OUTPUT ROW: one CTA owns row b of K for a whole batch of
// combinations, keeps it in shared memory (4 B x (b+1) <= 227 KB), streams the id ranges of its tasks and
// adds the row to HBM once per batch.
//
// The id ranges are read in aligned 16-byte units (PER = 8 u16 ids).  A warp takes 32 tasks (one coalesced
// load), cuts their ranges into units, and deals the concatenated units over its lanes 32 at a time
// (load-balanced expansion: the task owning position i is the number of task starts at or before i), so
// every load instruction is 32 x 16 B in a handful of cache lines and every lane then issues PER shared-
// memory atomics.  Runs start on unit boundaries (segment_kernel), so only the LAST unit of a task can hold
// ids that are not part of it, and those are larger than b (later sequences of the run, or the 0xFF.. fill):
// min(id, b + 1 + lane) sends them to one of 32 dump words behind the row -- no masks, no branches.
// the id stream is read once per task and never again by this CTA: no L1 allocation, 64-byte L2 prefetch
// (128-byte prefetch and no prefetch measured within 1 %: profiles/r01_v6_l2fetch_ldhint_experiment.txt)
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void smem_inc(uint32_t* row, uint32_t byte_off) {
    atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(row) + byte_off), 1u);
}

// col0: first column of K held by this CTA's shared-memory row (0 unless the row is split over several CTAs because N
// columns do not fit).  id - col0 wraps for ids below col0 -- to at least 2^16 - col0 >= dump in 16-bit arithmetic because
// N + 32 <= 2^16 there -- so one unsigned min sends both the ids before the window and those after it to the dump words.
template <typename IdT>
__device__ __forceinline__ void apply_unit(uint32_t* row, const uint4 v, const uint32_t dump, const uint32_t col0) {
    if (sizeof(IdT) == 2) {
        const uint32_t d2 = dump | (dump << 16);           // dump <= 0xffff whenever ids are 16-bit
        const uint32_t c2 = col0 | (col0 << 16);
        const uint32_t a = __vminu2(__vsub2(v.x, c2), d2), b = __vminu2(__vsub2(v.y, c2), d2), c = __vminu2(__vsub2(v.z, c2), d2),
                       d = __vminu2(__vsub2(v.w, c2), d2);
        smem_inc(row, (a << 2) & 0x3fffcu); smem_inc(row, (a >> 14) & 0x3fffcu);
        smem_inc(row, (b << 2) & 0x3fffcu); smem_inc(row, (b >> 14) & 0x3fffcu);
        smem_inc(row, (c << 2) & 0x3fffcu); smem_inc(row, (c >> 14) & 0x3fffcu);
        smem_inc(row, (d << 2) & 0x3fffcu); smem_inc(row, (d >> 14) & 0x3fffcu);
    } else {
        atomicAdd(&row[min(v.x - col0, dump)], 1u); atomicAdd(&row[min(v.y - col0, dump)], 1u);
        atomicAdd(&row[min(v.z - col0, dump)], 1u); atomicAdd(&row[min(v.w - col0, dump)], 1u);
    }
}

// grid = (rows of this wave, groups): CTA x owns row b = row_hi - x; the slots [group * slots_per_group,
// +slots_per_group) add into the same K.  The host launches the rows in waves of a few CTAs per SM, longest
// rows first.
constexpr uint32_t PF_STRIDE = 128, PF_LINES = 3, PF_SKIP = 0, PF_REACH = 4096;     // one prefetch per 128-byte line, reach 384 bytes of ids (PF_SKIP = 48, not prefetching a line the range barely enters, measured 179.6 against 171.4 ms)
// DIR: the tasks come from the run directory of segment_kernel's directory form: task of window t of row b in slot s =
// tdir[s][b >> bshift][wkey[s][window]] (two dependent loads, the first coalesced, the second a hit in L2: the rows of
// one launch belong to one or two blocks, so they all read the same 2^keybits x 8-byte column of each slot).
struct DirSpec {
    const uint16_t* wkey;        // [slot][window]
    const uint2* tdir;           // [slot][block][key]
    int bshift, keybits;
    uint32_t nb;
    uint32_t pf_stride;          // experiment: 64 = one L2 prefetch per 64 bytes of a task's id range instead of per 128-byte line
};
template <typename AccT, typename IdT, int UNROLL, bool PREFETCH = true, bool DIR = false>
__global__ void __launch_bounds__(1024)
accumulate_rows_kernel(const IdT* __restrict__ ids, size_t ids_stride, const uint2* __restrict__ task,
                       const uint32_t* __restrict__ woff, uint32_t n, uint32_t row_hi, int slots_per_group,
                       AccT* __restrict__ K, size_t k_group_stride, const WelfordSpec* __restrict__ wf, uint32_t col0,
                       uint32_t col_width, uint32_t sums_off, const uint32_t* __restrict__ heavy_bits, size_t heavy_bits_stride,
                       const DirSpec dir) {
    constexpr int PER = 16 / sizeof(IdT);
    constexpr int SH = PER == 8 ? 3 : 2;
    extern __shared__ uint32_t row[];
    __shared__ uint32_t next_chunk;
    const uint32_t b = row_hi - blockIdx.x;
    const int group = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t wb = woff[b], nw = woff[b + 1] - wb;
    const uint32_t cps = (nw + 31) >> 5;                     // chunks of 32 tasks per slot
    // integer modes: the slots of the group add into one row, in any order (one phase over all their chunks).  Variance mode:
    // the group's slots are consecutive iterations of one virtual stream, applied IN ORDER: one phase per slot, each followed by
    // the Welford step on the stream's running mean.
    const uint32_t slot_first = wf ? (uint32_t)wf->slot0[group] : (uint32_t)group * (uint32_t)slots_per_group;
    const uint32_t nphases = wf ? (uint32_t)wf->depth[group] : 1u;
    const uint32_t phase_chunks = wf ? cps : cps * (uint32_t)slots_per_group;
    uint32_t nchunks = phase_chunks;                         // end of the current phase's chunk range
    // this CTA holds columns [col0, col0 + ncols) of row b (all b + 1 of them unless N columns exceed shared memory: then the
    // host launches the row once per column window, and every window streams the row's tasks)
    const uint32_t ncols = min(b + 1 - col0, col_width);
    if (threadIdx.x == 0) next_chunk = 0;
    for (uint32_t i = threadIdx.x; i < ncols + 32; i += blockDim.x) row[i] = 0;
    __syncthreads();
    const uint32_t lane_le = 0xffffffffu >> (31 - lane);
    const uint32_t dump = ncols + lane;                      // 32 words behind the row
    const uint2* __restrict__ task_g = DIR ? nullptr : task + (size_t)slot_first * n + wb;
    const uint16_t* __restrict__ wkey_g = DIR ? dir.wkey + (size_t)slot_first * n + wb : nullptr;
    const uint2* __restrict__ tdir_g = DIR ? dir.tdir + ((((size_t)slot_first) * dir.nb + (b >> dir.bshift)) << dir.keybits) : nullptr;
    const size_t tdir_slot = DIR ? ((size_t)dir.nb << dir.keybits) : 0;
    const IdT* __restrict__ ids_g = ids + (size_t)slot_first * ids_stride;
    // chunk c -> (slot, first task); the task of this lane, or an empty one
    auto grab = [&]() -> uint32_t {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(&next_chunk, 1u);
        return __shfl_sync(0xffffffffu, c, 0);
    };
    // DIR: the key of this lane's window in chunk c (0xffffffff: no window), fetched one chunk before the task itself
    auto load_key = [&](uint32_t c) -> uint32_t {
        uint32_t k = 0xffffffffu;
        if (DIR && c < nchunks) {
            const uint32_t s = c / cps;
            const uint32_t t = ((c - s * cps) << 5) + lane;
            if (t < nw) k = wkey_g[(size_t)s * n + t];
        }
        return k;
    };
    // the task as stored (a plain load, nothing depends on it yet) ...
    auto load_task = [&](uint32_t c, uint32_t key) -> uint2 {
        uint2 q = make_uint2(0, 0);
        if (c < nchunks) {
            const uint32_t s = c / cps;
            const uint32_t t = ((c - s * cps) << 5) + lane;
            if (DIR) { if (key != 0xffffffffu) q = tdir_g[(size_t)s * tdir_slot + key]; }
            else if (t < nw) q = task_g[(size_t)s * n + t];
        }
        return q;
    };
    // ... and as applied, one chunk later: a task of a long run (segment_kernel) is an ordinary task unless the tensor-core
    // contraction took the run
    auto settle_task = [&](uint32_t c, uint2 q) -> uint2 {
        if (q.y >> 31) {
            const uint32_t s = c / cps;
            const uint32_t w = heavy_bits[((size_t)slot_first + s) * heavy_bits_stride + (q.x >> 5)];
            q.y &= 0x7fffffffu;
            // taken: the task must still own one unit (the expansion below hands positions to consecutive lanes, so no
            // lane in the middle may be empty): the last unit of the slot's id stream, which is always 0xFF.. fill
            if ((w >> (q.x & 31u)) & 1u) q = make_uint2((uint32_t)(ids_stride / PER) - 1u, 1u);
        }
        return q;
    };
    // Software pipeline over the chunks of 32 tasks, three deep: chunk c is applied while the id ranges of c1 are pulled into
    // L2 (its tasks were fetched a whole chunk ago, so nothing waits for them), the tasks of c2 are in flight and -- DIR --
    // so are the keys of c3.  The kernel's stalls are memory latency (ncu: ~28 % of the samples on the first use of the id
    // loads, 7 % on the task loads with a two-deep pipeline), and a prefetch holds no register and no scoreboard entry.
    for (uint32_t phase = 0; phase < nphases; ++phase) {
    if (phase) {                                             // next slot of the stream: its chunks [phase * cps, + cps)
        nchunks = (phase + 1) * cps;
        if (threadIdx.x == 0) next_chunk = phase * cps;
        __syncthreads();
    }
    uint32_t c = grab(), c1 = grab(), c2 = grab();
    uint2 q = settle_task(c, load_task(c, load_key(c)));
    uint2 q1 = load_task(c1, load_key(c1));
    uint32_t k2 = load_key(c2);
    while (c < nchunks) {
        const uint32_t s = c / cps;
        const uint4* __restrict__ ip = reinterpret_cast<const uint4*>(ids_g + (size_t)s * ids_stride);   // ids_stride is a multiple of 64
        const uint2 q2 = load_task(c2, k2);
        const uint32_t c3 = grab();
        const uint32_t k3 = load_key(c3);
        q1 = settle_task(c1, q1);
        if (PREFETCH && q1.y) {                      // per-lane prefetches (the bulk form is warp-uniform: 32 serial issues)
            const char* p = reinterpret_cast<const char*>(ids_g + (size_t)(c1 / cps) * ids_stride) + (size_t)q1.x * 16;
            const uint32_t bytes = q1.y * (uint32_t)sizeof(IdT);
            if (dir.pf_stride == 64u) {
#pragma unroll
                for (uint32_t o = 0; o < PF_LINES * PF_STRIDE; o += 64u)
                    if (bytes > o) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
            } else {
#pragma unroll
                for (uint32_t o = 0; o < PF_LINES * PF_STRIDE; o += PF_STRIDE)
                    if (bytes > o + PF_SKIP) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
            }
            // long ranges (skewed inputs): up to PF_REACH bytes; the uniform workload never gets here
            for (uint32_t o = PF_LINES * PF_STRIDE; o < min(bytes, PF_REACH); o += PF_STRIDE)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
        }
        const uint32_t my_units = (q.y + PER - 1) >> SH;
        uint32_t incl = my_units;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t P = incl - my_units;                 // first position of this lane's task in the concatenation
        const uint32_t W = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t u0 = q.x - P;                         // unit index = u0[owner] + position
        uint32_t started = 0;
        // loads of one step: UNROLL x 32 consecutive positions of the concatenation
        auto issue = [&](uint32_t base, uint4 (&v)[UNROLL]) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t wbase = base + 32 * u;
                const uint32_t rel = P - wbase;
                const uint32_t m = __reduce_or_sync(0xffffffffu, (rel < 32u && my_units) ? (1u << rel) : 0u);
                const uint32_t j = (started + __popc(m & lane_le) - 1) & 31;
                started += __popc(m);
                const uint32_t unit = __shfl_sync(0xffffffffu, u0, j) + wbase + lane;
                if (wbase + lane < W) v[u] = ldg_stream_u4(ip + unit);
            }
        };
        auto apply = [&](uint32_t base, const uint4 (&v)[UNROLL]) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (base + 32 * u + lane < W) apply_unit<IdT>(row, v[u], dump, col0);
        };
        // (double-buffering these loads against the applies measured no gain: profiles/r01_v6_unroll_pipe_segexp_experiment.txt)
        for (uint32_t base = 0; base < W; base += 32 * UNROLL) {
            uint4 v[UNROLL];
            issue(base, v);
            apply(base, v);
        }
        c = c1; q = q1;
        c1 = c2; q1 = q2;
        c2 = c3; k2 = k3;
    }
    __syncthreads();
    if (wf) {
        // variance mode: the row holds this iteration's partial kernel Ks of the stream; apply the Welford step to the
        // stream's running mean right here (no Ks in HBM, no separate pass): fastsk_kernel.cpp:121-135.  The first slot of the
        // round reads the mean the round started from, the later ones what the previous slot left.
        __shared__ double ws[32];
        const size_t roff = ((size_t)b * (b + 1) >> 1) + col0;
        const double* __restrict__ kin = (phase == 0 ? wf->khat_in[group] : wf->khat_out[group]) + roff;
        double* __restrict__ kout = wf->khat_out[group] + roff;
        const double diter = (double)(wf->iter0[group] + (int32_t)phase);
        const double riter = __ddiv_rn(1.0, diter);
        const bool train = (int64_t)b < wf->n_train;
        double acc = 0.0;
        for (uint32_t i = threadIdx.x; i < ncols; i += blockDim.x) {
            const double ks = (double)row[i];
            const double k0 = kin[i];
            const double delta = __dsub_rn(ks, k0);
            const double nh = __dadd_rn(k0, div_by_iter(delta, diter, riter));
            kout[i] = nh;
            row[i] = 0;
            if (train) acc = __dadd_rn(acc, __dmul_rn(delta, __dsub_rn(ks, nh)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (lane == 0) ws[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) t = __dadd_rn(t, ws[w]);
            wf->sums[(size_t)(slot_first + phase) * wf->sums_stride + sums_off + b] = t;
        }
        continue;                                            // (the next phase's barrier orders the row's zeroes)
    }
    }   // phases
    if (wf) return;
    AccT* __restrict__ Krow = K + (size_t)group * k_group_stride + ((size_t)b * (b + 1) >> 1) + col0;
    for (uint32_t i = threadIdx.x; i < ncols; i += blockDim.x) {
        const uint32_t v = row[i];
        if (v) atomicAdd(&Krow[i], (AccT)v);                 // RED: no read round trip on the flush
    }
}

// Fallback when a row of K does not fit in shared memory (N > ~56 000): the same tasks, added with global
// RED on the packed triangle.  One warp per task; tasks live at their row's window range, so the window ->
// sequence table gives the row.
template <typename AccT, typename IdT>
__global__ void __launch_bounds__(256)
accumulate_global_kernel(const IdT* __restrict__ ids, size_t ids_stride, const uint2* __restrict__ task,
                         const uint32_t* __restrict__ wseq, uint32_t n, AccT* __restrict__ K, size_t k_slot_stride) {
    constexpr int PER = 16 / sizeof(IdT);
    const int slot = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t t0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
    const IdT* __restrict__ ip = ids + (size_t)slot * ids_stride;
    AccT* __restrict__ Ks = K + (size_t)slot * k_slot_stride;
    for (uint32_t t = t0; t < min(t0 + 32u, n); ++t) {
        const uint2 q = task[(size_t)slot * n + t];
        const size_t b = wseq[t];
        AccT* __restrict__ Krow = Ks + (b * (b + 1) >> 1);
        for (uint32_t a = lane; a < q.y; a += 32) atomicAdd(&Krow[ip[(size_t)q.x * PER + a]], (AccT)1);
    }
}

// ------------------------------------------------------------------------------------------
// normalisation (fastsk_kernel.cpp:96-103): out = K_ij / sqrt(K_ii * K_jj) with IEEE mul, sqrt,
// div (bit-identical to the reference's x86-64 doubles); the diagonal formula K_ii/sqrt(K_ii*K_ii)
// is the same expression.
// The unnormalised kernel as the normalisation sees it: the SUM of the partial kernels of all ranks.  parts.p[r] is rank r's
// packed triangle -- this GPU's own buffer, or a peer's HBM mapped over NVLink (cudaDeviceEnablePeerAccess inside one
// process, cudaIpcOpenMemHandle between the processes of a torchrun launch).  The reduction of the partial kernels
// (fastsk_kernel.cpp:285-315: the mutex-striped merge) is therefore FUSED into the normalisation: every cell crosses NVLink
// once, as an operand load of the kernel that consumes it, and no reduced copy of K is ever written.  Integer partials add
// exactly; fp64 partials (variance mode) are added in rank order on every rank.
constexpr int MAX_PEERS = 16;
template <typename T>
struct PeerParts {
    const T* p[MAX_PEERS];
    int n;
};
template <typename T>
__device__ __forceinline__ double peer_sum(const PeerParts<T>& parts, int64_t idx) {
    if (std::is_same<T, double>::value) {
        double s = (double)parts.p[0][idx];
        for (int r = 1; r < parts.n; ++r) s = __dadd_rn(s, (double)parts.p[r][idx]);
        return s;
    }
    unsigned long long s = (unsigned long long)parts.p[0][idx];
    for (int r = 1; r < parts.n; ++r) s += (unsigned long long)parts.p[r][idx];
    return (double)s;
}

template <typename T>
__global__ void diag_kernel(const __grid_constant__ PeerParts<T> parts, int64_t n, double* __restrict__ diag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) diag[i] = peer_sum(parts, (i * (i + 1) >> 1) + i);
}

__device__ __forceinline__ double norm_entry(double v, double di, double dj) {
    return __ddiv_rn(v, __dsqrt_rn(__dmul_rn(di, dj)));
}

// square block rows [r0, r0+nr) x cols [0, nc) of the symmetric matrix -> out (row-major, ld = nc).
// 32x32 tiles; tiles above the diagonal are read transposed through shared memory so that the
// packed triangle is always read along its contiguous direction.
// Two launches: pass 0 does the tiles on or below the diagonal with the tile COLUMN in blockIdx.x (neighbouring CTAs walk
// along a row of K), pass 1 the tiles above it with the tile ROW in blockIdx.x (neighbouring CTAs read neighbouring 256-byte
// pieces of the SAME rows j of K: 8 x 157 of them in a row at 8 ranks instead of 256 bytes here and there -- the mirrored
// half is most of rank 0's share, and it is read from seven peers over NVLink).
template <typename T>
__global__ void __launch_bounds__(256)
normalise_block_kernel(const __grid_constant__ PeerParts<T> parts, const double* __restrict__ diag, int64_t r0, int64_t nr, int64_t nc,
                       double* __restrict__ out, int pass) {
    __shared__ double tile[32][33];
    // tile origin (row offset within block, col)
    const int64_t ti = (int64_t)(pass ? blockIdx.x : blockIdx.y) * 32, tj = (int64_t)(pass ? blockIdx.y : blockIdx.x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool upper = (r0 + ti + 31) < tj;   // whole tile strictly above the diagonal: read mirrored
    if (upper != (pass != 0)) return;
    if (!upper) {
        for (int y = ty; y < 32; y += 8) {
            const int64_t i = r0 + ti + y, j = tj + tx;
            if (ti + y < nr && j < nc) {
                const int64_t a = i >= j ? i : j, b = i >= j ? j : i;
                out[(ti + y) * nc + j] = norm_entry(peer_sum(parts, (a * (a + 1) >> 1) + b), diag[i], diag[j]);
            }
        }
    } else {
        for (int y = ty; y < 32; y += 8) {      // read K[j][i] with threads along i (contiguous)
            const int64_t j = tj + y, i = r0 + ti + tx;
            if (ti + tx < nr && j < nc) tile[y][tx] = norm_entry(peer_sum(parts, (j * (j + 1) >> 1) + i), diag[i], diag[j]);
        }
        __syncthreads();
        for (int y = ty; y < 32; y += 8) {
            const int64_t j = tj + tx;
            if (ti + y < nr && j < nc) out[(ti + y) * nc + j] = tile[tx][y];
        }
    }
}

// rows [i0, i0 + gridDim.y) of the packed triangle, normalised (out is indexed like the triangle, from row i0's first cell)
template <typename T>
__global__ void normalise_packed_kernel(const __grid_constant__ PeerParts<T> parts, const double* __restrict__ diag, int64_t i0,
                                        double* __restrict__ out) {
    const int64_t i = i0 + blockIdx.y;
    const int64_t base0 = i0 * (i0 + 1) >> 1;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= i; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = (i * (i + 1) >> 1) + j;
        out[p - base0] = norm_entry(peer_sum(parts, p), diag[i], diag[j]);
    }
}

// cells [c0, c0 + n) of the summed triangle, as int64 or fp64 (getters of the unnormalised kernel when K is sharded)
template <typename T, typename O>
__global__ void sum_parts_kernel(const __grid_constant__ PeerParts<T> parts, int64_t c0, int64_t n, O* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (std::is_same<T, double>::value || std::is_same<O, double>::value) out[i] = (O)peer_sum(parts, c0 + i);
        else {
            unsigned long long s = 0;
            for (int r = 0; r < parts.n; ++r) s += (unsigned long long)parts.p[r][c0 + i];
            out[i] = (O)s;
        }
    }
}

template <typename T>
__global__ void to_f64_kernel(const T* __restrict__ in, double* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)in[i];
}

// ------------------------------------------------------------------------------------------
// Welford step of one virtual stream (fastsk_kernel.cpp:108-143), fused with re-zeroing the
// per-iteration integer partial (fastsk_kernel.cpp:192-194).  The sum of delta*delta2 over the
// train x train triangle is reduced per block; welford_final_kernel adds the block sums in a
// fixed order, so a run is reproducible.
constexpr int WELFORD_BLOCKS = 592;   // 4 x 148 SMs
template <typename AccT>
__global__ void __launch_bounds__(256, 4)     // 4 CTAs per SM: the 592 blocks are one full wave
welford_kernel(AccT* __restrict__ Ks, double* __restrict__ K_hat, int64_t n_pairs, int64_t n_train_pairs, int iter,
               double* __restrict__ block_sums) {
    __shared__ double ws[8];
    const double diter = (double)iter;
    const double riter = __ddiv_rn(1.0, diter);
    double acc = 0.0;
    // four independent element loads per array in flight per thread; a thread still visits its elements in increasing
    // order, so the per-thread partial sums (and the result) do not depend on the unrolling
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; p + 3 * stride < n_pairs; p += 4 * stride) {
        double ks[4], kh[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ks[u] = (double)Ks[p + u * stride];
            kh[u] = K_hat[p + u * stride];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t q = p + u * stride;
            Ks[q] = 0;
            const double delta = __dsub_rn(ks[u], kh[u]);
            const double nh = __dadd_rn(kh[u], div_by_iter(delta, diter, riter));
            K_hat[q] = nh;
            if (q < n_train_pairs) acc = __dadd_rn(acc, __dmul_rn(delta, __dsub_rn(ks[u], nh)));
        }
    }
    for (; p < n_pairs; p += stride) {
        const double ks = (double)Ks[p];
        Ks[p] = 0;
        double kh = K_hat[p];
        const double delta = __dsub_rn(ks, kh);
        kh = __dadd_rn(kh, div_by_iter(delta, diter, riter));
        K_hat[p] = kh;
        if (p < n_train_pairs) acc = __dadd_rn(acc, __dmul_rn(delta, __dsub_rn(ks, kh)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s = __dadd_rn(s, ws[w]);
        block_sums[blockIdx.x] = s;
    }
}

// grid = slots: block s adds the nblocks partial sums of slot s (block_sums + s * stride) in a fixed order
__global__ void welford_final_kernel(const double* __restrict__ block_sums_all, size_t stride, int nblocks, double* __restrict__ out_all) {
    __shared__ double ws[32];
    const double* __restrict__ block_sums = block_sums_all + (size_t)blockIdx.x * stride;
    double* __restrict__ out = out_all + blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc = __dadd_rn(acc, block_sums[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = __dadd_rn(s, ws[w]);
        *out = s;
    }
}

// Ksfinal += K_hat of one stream (fastsk_kernel.cpp:296-313)
__global__ void add_f64_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __dadd_rn(dst[i], src[i]);
}

}  // namespace fsk
