// Device kernels of the gapped k-mer kernel-matrix build (sm_100a).
//
// Per batch of combinations ("slots") the pipeline is
//   pack_hist   g-mer word -> kept characters packed into a key, tagged with the sequence id;
//               digit histograms of every radix pass in the same sweep
//               (replaces the column gather of fastsk_kernel.cpp:224-228)
//   onesweep    stable LSD radix sort over the key bits only, one kernel per 8-bit digit:
//               warp match ranking + decoupled look-back (replaces cntsrtna, shared.cpp:156-191,
//               and the gather of sorted features, fastsk_kernel.cpp:233-238)
//   segment     run boundaries and per-(k-mer, sequence) counts (shared.cpp:280-315)
//   accumulate  K[i][j] += c_i * c_j on the packed lower triangle (shared.cpp:316-327)
// plus normalise (fastsk_kernel.cpp:96-103) and the Welford / variance pass
// (fastsk_kernel.cpp:108-143).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fsk {

constexpr int MAX_K = 32;        // kept positions per combination
constexpr int MAX_BATCH = 48;    // combinations per launch group (kernel-parameter budget)
constexpr int MAX_PASS = 8;      // 64 key bits / 8
constexpr int RADIX = 256;
constexpr int SORT_THREADS = 256;
constexpr int SEG_THREADS = 256;
constexpr int SEG_ITEMS = 8;
constexpr int SEG_TILE = SEG_THREADS * SEG_ITEMS;
constexpr int ACC_ROWS = 64;     // entries (rows of updates) per accumulate CTA
constexpr int PACK_ITEMS = 8;

struct BatchSpec {               // by value in kernel parameter space
    uint8_t src[MAX_BATCH][MAX_K];   // bit position (word * 64 + shift) of each kept character
};
struct SortPlan {
    int npass;
    uint8_t shift[MAX_PASS];     // digit position relative to key bit 0
    uint8_t bits[MAX_PASS];
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v));
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
// pack + histogram.  grid = (tiles, slots).
template <typename RecT, bool KV, typename GwT, int NW>
__global__ void __launch_bounds__(256)
pack_hist_kernel(const GwT* __restrict__ gw0, const uint64_t* __restrict__ gw1, const uint32_t* __restrict__ wseq,
                 uint32_t nfeat, RecT* __restrict__ rec, uint32_t* __restrict__ val, uint32_t* __restrict__ ghist,
                 const __grid_constant__ BatchSpec spec, const __grid_constant__ SortPlan plan, int k, int b, int idbits) {
    __shared__ uint32_t sh[MAX_PASS * RADIX];
    const int slot = blockIdx.y;
    const int npass = plan.npass;
    for (int i = threadIdx.x; i < npass * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t cmask = (1ull << b) - 1;
    const size_t sbase = (size_t)slot * nfeat;
    const uint32_t tile0 = blockIdx.x * (256 * PACK_ITEMS);
#pragma unroll 2
    for (int it = 0; it < PACK_ITEMS; ++it) {
        const uint32_t w = tile0 + it * 256 + threadIdx.x;
        if (w < nfeat) {
            const uint64_t lo = gw0[w];
            uint64_t hi = 0;
            if (NW == 2) hi = gw1[w];
            uint64_t key = 0;
            for (int j = 0; j < k; ++j) {
                const uint32_t sp = spec.src[slot][j];
                const uint64_t word = (NW == 2 && (sp & 64)) ? hi : lo;
                key = (key << b) | ((word >> (sp & 63)) & cmask);
            }
            const uint32_t seq = wseq[w];
            if (KV) {
                rec[sbase + w] = (RecT)key;
                val[sbase + w] = seq;
            } else {
                rec[sbase + w] = (RecT)((key << idbits) | seq);
            }
            for (int p = 0; p < npass; ++p) {
                const uint32_t d = (uint32_t)(key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1);
                atomicAdd(&sh[p * RADIX + d], 1u);
            }
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
template <typename RecT, bool KV, int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS)
onesweep_kernel(const RecT* __restrict__ in, RecT* __restrict__ out, const uint32_t* __restrict__ vin,
                uint32_t* __restrict__ vout, uint32_t n, uint32_t tiles_per_slot, int shift, int bits,
                const uint32_t* __restrict__ ghist /* [slot][MAX_PASS][RADIX], pre-offset to this pass */,
                uint32_t* __restrict__ status /* [slot][tile][RADIX] of this pass */, uint32_t* __restrict__ ticket) {
    constexpr int TILE = SORT_THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RecT* skeys = reinterpret_cast<RecT*>(smem_raw);
    uint32_t* svals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(RecT) * TILE);
    uint32_t* warp_hist = svals + (KV ? TILE : 0);   // [8][RADIX]
    uint32_t* digit_start = warp_hist + 8 * RADIX;    // [RADIX]
    uint32_t* scatter_base = digit_start + RADIX;     // [RADIX]
    uint32_t* warp_sums = scatter_base + RADIX;       // [8]
    __shared__ uint32_t s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    for (int i = tid; i < 8 * RADIX; i += SORT_THREADS) warp_hist[i] = 0;
    __syncthreads();
    const uint32_t slot = s_ticket / tiles_per_slot;
    const uint32_t tile = s_ticket - slot * tiles_per_slot;
    const size_t sbase = (size_t)slot * n;
    const uint32_t tile0 = tile * TILE;
    const uint32_t dmask = (1u << bits) - 1;

    RecT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rank[ITEMS];
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
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t idx = tile0 + warp * (32 * ITEMS) + j * 32 + lane;
        const bool valid = idx < n;
        const uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & dmask) : (0x80000000u | lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t lower = peers & ((1u << lane) - 1);
        uint32_t prev = 0;
        if (valid) prev = warp_hist[warp * RADIX + d];
        __syncwarp();
        if (valid && lower == 0) warp_hist[warp * RADIX + d] = prev + __popc(peers);
        __syncwarp();
        rank[j] = prev + __popc(lower);
    }
    __syncthreads();

    // thread d owns digit d: exclusive scan over warps, tile total
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t c = warp_hist[w * RADIX + tid];
        warp_hist[w * RADIX + tid] = total;
        total += c;
    }
    uint32_t* my_status = status + ((size_t)slot * tiles_per_slot + tile) * RADIX + tid;
    st_volatile_u32(my_status, total | (tile == 0 ? 0x80000000u : 0x40000000u));

    uint32_t dummy;
    const uint32_t gexcl = block_excl_scan_256(ghist[(size_t)slot * MAX_PASS * RADIX + tid], warp_sums, dummy);
    const uint32_t dstart = block_excl_scan_256(total, warp_sums, dummy);

    // decoupled look-back, LOOK predecessors per step with independent loads in flight: the first wave of
    // resident tiles has to walk back over every tile that started with it
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
    digit_start[tid] = dstart;
    scatter_base[tid] = gexcl + excl - dstart;
    __syncthreads();

#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t idx = tile0 + warp * (32 * ITEMS) + j * 32 + lane;
        if (idx < n) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
            const uint32_t pos = digit_start[d] + warp_hist[warp * RADIX + d] + rank[j];
            skeys[pos] = key[j];
            if (KV) svals[pos] = val[j];
        }
    }
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
// segmentation of the sorted records.
//   entry = one distinct (k-mer, sequence) cell; ent_seq (bit 31: first entry of its run),
//           ent_start (index of its first record; count = ent_start[e+1] - ent_start[e]),
//           ent_run (index of its run);  run_start[r] = first entry of run r.
template <typename RecT, bool KV>
__device__ __forceinline__ void seg_flags(const RecT* __restrict__ rec, const uint32_t* __restrict__ val, size_t sbase,
                                          uint32_t i, int idbits, bool& new_ent, bool& new_run, uint32_t& seq) {
    const RecT r = rec[sbase + i];
    if (KV) {
        seq = val[sbase + i];
        if (i == 0) { new_ent = new_run = true; return; }
        const RecT p = rec[sbase + i - 1];
        new_run = p != r;
        new_ent = new_run || val[sbase + i - 1] != seq;
    } else {
        seq = (uint32_t)(r & (((RecT)1 << idbits) - 1));
        if (i == 0) { new_ent = new_run = true; return; }
        const RecT p = rec[sbase + i - 1];
        new_run = (p >> idbits) != (r >> idbits);
        new_ent = p != r;
    }
}

template <typename RecT, bool KV>
__global__ void __launch_bounds__(SEG_THREADS)
seg_count_kernel(const RecT* __restrict__ rec, const uint32_t* __restrict__ val, uint32_t n, int idbits,
                 uint32_t tiles_per_slot, uint2* __restrict__ tile_counts) {
    __shared__ uint32_t se[8], sr[8];
    const int slot = blockIdx.y;
    const size_t sbase = (size_t)slot * n;
    const uint32_t i0 = blockIdx.x * SEG_TILE + threadIdx.x * SEG_ITEMS;
    uint32_t ce = 0, cr = 0;
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; ++j) {
        const uint32_t i = i0 + j;
        if (i < n) {
            bool ne, nr;
            uint32_t seq;
            seg_flags<RecT, KV>(rec, val, sbase, i, idbits, ne, nr, seq);
            ce += ne;
            cr += nr;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ce += __shfl_xor_sync(0xffffffffu, ce, o);
        cr += __shfl_xor_sync(0xffffffffu, cr, o);
    }
    if ((threadIdx.x & 31) == 0) { se[threadIdx.x >> 5] = ce; sr[threadIdx.x >> 5] = cr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t e = 0, r = 0;
        for (int w = 0; w < 8; ++w) { e += se[w]; r += sr[w]; }
        tile_counts[(size_t)slot * tiles_per_slot + blockIdx.x] = make_uint2(e, r);
    }
}

// one CTA per slot: exclusive scan of the tile counts; totals; sentinel ent_start[E] = n
__global__ void __launch_bounds__(1024)
seg_scan_kernel(const uint2* __restrict__ tile_counts, uint2* __restrict__ tile_offs, uint32_t tiles_per_slot,
                uint2* __restrict__ totals, uint32_t* __restrict__ ent_start, uint32_t n, unsigned long long* __restrict__ stat_counters) {
    __shared__ uint32_t we[32], wr[32];
    __shared__ uint32_t carry_e, carry_r;
    const int slot = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_e = 0; carry_r = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < tiles_per_slot; base += 1024) {
        const uint32_t t = base + threadIdx.x;
        uint2 c = make_uint2(0, 0);
        if (t < tiles_per_slot) c = tile_counts[(size_t)slot * tiles_per_slot + t];
        uint32_t ie = c.x, ir = c.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, ie, o), b = __shfl_up_sync(0xffffffffu, ir, o);
            if (lane >= o) { ie += a; ir += b; }
        }
        if (lane == 31) { we[warp] = ie; wr[warp] = ir; }
        __syncthreads();
        uint32_t be = carry_e, br = carry_r, te = 0, tr = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < warp) { be += we[w]; br += wr[w]; }
            te += we[w];
            tr += wr[w];
        }
        if (t < tiles_per_slot) tile_offs[(size_t)slot * tiles_per_slot + t] = make_uint2(be + ie - c.x, br + ir - c.y);
        __syncthreads();
        if (threadIdx.x == 0) { carry_e += te; carry_r += tr; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totals[slot] = make_uint2(carry_e, carry_r);
        ent_start[(size_t)slot * (n + 1) + carry_e] = n;
        if (stat_counters) {
            atomicAdd(&stat_counters[0], (unsigned long long)carry_e);
            atomicAdd(&stat_counters[1], (unsigned long long)carry_r);
        }
    }
}

template <typename RecT, bool KV>
__global__ void __launch_bounds__(SEG_THREADS)
seg_write_kernel(const RecT* __restrict__ rec, const uint32_t* __restrict__ val, uint32_t n, int idbits,
                 uint32_t tiles_per_slot, const uint2* __restrict__ tile_offs, uint32_t* __restrict__ ent_seq,
                 uint32_t* __restrict__ ent_start, uint32_t* __restrict__ ent_run, uint32_t* __restrict__ run_start) {
    __shared__ uint32_t warp_sums[8];
    const int slot = blockIdx.y;
    const size_t sbase = (size_t)slot * n;
    const size_t ebase = (size_t)slot * (n + 1);
    const uint32_t i0 = blockIdx.x * SEG_TILE + threadIdx.x * SEG_ITEMS;
    uint32_t fe = 0, fr = 0, seqs[SEG_ITEMS];
    uint32_t ce = 0, cr = 0;
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; ++j) {
        const uint32_t i = i0 + j;
        seqs[j] = 0;
        if (i < n) {
            bool ne, nr;
            seg_flags<RecT, KV>(rec, val, sbase, i, idbits, ne, nr, seqs[j]);
            fe |= (uint32_t)ne << j;
            fr |= (uint32_t)nr << j;
            ce += ne;
            cr += nr;
        }
    }
    uint32_t total;
    const uint32_t ex = block_excl_scan_256(ce | (cr << 16), warp_sums, total);   // <= 2048 each: 12 bits
    const uint2 off = tile_offs[(size_t)slot * tiles_per_slot + blockIdx.x];
    uint32_t e = off.x + (ex & 0xffffu);
    uint32_t r = off.y + (ex >> 16);
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; ++j) {
        if (fe >> j & 1) {
            const bool head = fr >> j & 1;
            if (head) { run_start[sbase + r] = e; ++r; }
            ent_seq[sbase + e] = seqs[j] | (head ? 0x80000000u : 0u);
            ent_start[ebase + e] = i0 + j;
            ent_run[sbase + e] = r - 1;
            ++e;
        }
    }
}

// ------------------------------------------------------------------------------------------
// accumulate: each entry b of a run is one row of updates K[seq_b][seq_a] += c_a * c_b over the
// entries a <= b of the same run (ids ascend inside a run, so seq_a <= seq_b: packed lower
// triangle, 64-bit index; shared.cpp:97-117 uses int).  One warp per row, lanes over a.
// grid = (ceil(n / ACC_ROWS), slots); CTAs beyond the slot's entry count exit.
template <typename AccT>
__global__ void __launch_bounds__(256)
accumulate_kernel(const uint32_t* __restrict__ ent_seq, const uint32_t* __restrict__ ent_start,
                  const uint32_t* __restrict__ ent_run, const uint32_t* __restrict__ run_start,
                  const uint2* __restrict__ totals, uint32_t n, AccT* __restrict__ K, size_t k_slot_stride,
                  unsigned long long* __restrict__ stat_counters) {
    const int slot = blockIdx.y;
    const uint32_t E = totals[slot].x;
    const uint32_t first = blockIdx.x * ACC_ROWS;
    if (first >= E) return;
    const size_t sbase = (size_t)slot * n;
    const size_t ebase = (size_t)slot * (n + 1);
    AccT* Ks = K + (size_t)slot * k_slot_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t last = min(first + ACC_ROWS, E);
    unsigned long long updates = 0;
    for (uint32_t eb = first + warp; eb < last; eb += 8) {
        const uint32_t sb = ent_seq[sbase + eb] & 0x7fffffffu;
        const uint32_t cb = ent_start[ebase + eb + 1] - ent_start[ebase + eb];
        const uint32_t rs = run_start[sbase + ent_run[sbase + eb]];
        AccT* row = Ks + ((size_t)sb * (sb + 1) >> 1);
        for (uint32_t a = rs + lane; a <= eb; a += 32) {
            const uint32_t sa = ent_seq[sbase + a] & 0x7fffffffu;
            const uint32_t ca = ent_start[ebase + a + 1] - ent_start[ebase + a];
            atomicAdd(row + sa, (AccT)ca * (AccT)cb);
        }
        updates += eb - rs + 1;
    }
    if (stat_counters && lane == 0) atomicAdd(&stat_counters[2], updates);
}

// ------------------------------------------------------------------------------------------
// Row-stationary accumulate (the default path).
//
// Measured on B200 (profiles/r01_atomic_microbench.txt): scattered RED into the 10 GB packed
// triangle runs at 20 G updates/s (DRAM sector read-modify-write), L2-resident at 210 G/s, and
// shared-memory atomics at > 800 G/s.  So the update is re-ordered by OUTPUT ROW: one CTA owns
// row b of K for a whole batch of combinations, keeps it in shared memory (4 B x (b+1), <= 227 KB),
// streams the run prefixes of b's k-mers and flushes the row to HBM once per batch.
//
// seg_finish turns every entry (k-mer, sequence b) into a task (run start, entry) filed under
// row b: task[woff[b] + i], i < row_count[b] (a sequence has at most as many entries as
// windows, so its window range is its task range), and packs (sequence, count) into one word.
__global__ void __launch_bounds__(256)
seg_finish_kernel(const uint32_t* __restrict__ ent_seq, const uint32_t* __restrict__ ent_start,
                  const uint32_t* __restrict__ ent_run, const uint32_t* __restrict__ run_start,
                  const uint2* __restrict__ totals, uint32_t n, uint32_t nseq, int idbits,
                  const uint32_t* __restrict__ woff, uint32_t* __restrict__ row_count,
                  uint32_t* __restrict__ ent_pack, uint2* __restrict__ task) {
    const int slot = blockIdx.y;
    const uint32_t e = blockIdx.x * 256 + threadIdx.x;
    if (e >= totals[slot].x) return;
    const size_t sbase = (size_t)slot * n, ebase = (size_t)slot * (n + 1);
    const uint32_t sb = ent_seq[sbase + e] & 0x7fffffffu;
    const uint32_t cnt = ent_start[ebase + e + 1] - ent_start[ebase + e];
    const uint32_t rs = run_start[sbase + ent_run[sbase + e]];
    ent_pack[sbase + e] = sb | (cnt << idbits);
    const uint32_t pos = atomicAdd(&row_count[(size_t)slot * nseq + sb], 1u);
    // task = (first entry of the run, prefix length - 1 | own count << idbits): same idbits + countbits <= 32
    // condition as ent_pack
    task[sbase + woff[sb] + pos] = make_uint2(rs, (e - rs) | (cnt << idbits));   // length - 1: 0 .. nseq-1 fits idbits
}

// grid = (rows, groups); the slots [group * slots_per_group, +slots_per_group) add into the same K.
// Row N-1 first: the longest rows lead, the short ones fill the tail.
// All (slot, task) pairs of the row are cut into chunks of 32 tasks; a warp takes a chunk, its lanes load
// the 32 task words in one coalesced access, then the warp walks the tasks UNROLL at a time so that
// UNROLL independent entry loads are in flight per lane ahead of the shared-memory atomics.
template <typename AccT, int UNROLL>
__global__ void __launch_bounds__(1024)
accumulate_rows_kernel(const uint32_t* __restrict__ ent_pack, const uint2* __restrict__ task,
                       const uint32_t* __restrict__ row_count, const uint32_t* __restrict__ woff, uint32_t n,
                       uint32_t nseq, int idbits, int slots_per_group, AccT* __restrict__ K, size_t k_group_stride,
                       unsigned long long* __restrict__ stat_counters) {
    extern __shared__ uint32_t row[];
    __shared__ uint32_t chunk_prefix[MAX_BATCH + 1];
    __shared__ uint32_t next_chunk;
    const uint32_t b = nseq - 1 - blockIdx.x;
    const int group = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int s = 0; s < slots_per_group; ++s) {
            chunk_prefix[s] = acc;
            acc += (row_count[(size_t)(group * slots_per_group + s) * nseq + b] + 31) >> 5;
        }
        chunk_prefix[slots_per_group] = acc;
        next_chunk = 0;
    }
    __syncthreads();
    const uint32_t nchunks = chunk_prefix[slots_per_group];
    if (nchunks == 0) return;
    for (uint32_t i = threadIdx.x; i <= b; i += blockDim.x) row[i] = 0;
    __syncthreads();
    const uint32_t idmask = (1u << idbits) - 1;
    const uint32_t wb = woff[b];
    const uint32_t lane_le = 0xffffffffu >> (31 - lane);
    unsigned long long updates = 0;
    int s = 0;
    while (true) {
        // dynamic chunk scheduling: chunk ids only grow, so a warp's slot cursor only moves forward
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(&next_chunk, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= nchunks) break;
        while (chunk_prefix[s + 1] <= c) ++s;
        const int slot = group * slots_per_group + s;
        const uint32_t nt = row_count[(size_t)slot * nseq + b];
        const uint32_t t = ((c - chunk_prefix[s]) << 5) + lane;
        const uint32_t* __restrict__ ep = ent_pack + (size_t)slot * n;
        uint2 q = make_uint2(0, 0);
        uint32_t my_len = 0;
        if (t < nt) {
            q = task[(size_t)slot * n + wb + t];
            my_len = (q.y & idmask) + 1;
        }
        const uint32_t my_cb = q.y >> idbits;
        // load-balanced expansion of the 32 run prefixes over the lanes: P = start of task `lane` in the
        // concatenation, W = total length.  For a window of 32 positions the tasks starting inside it are a
        // bit mask (one REDUX); the task owning position i is the number of starts at or before i, minus 1.
        uint32_t incl = my_len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t P = incl - my_len;
        const uint32_t W = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t delta = q.x - P;                       // entry index = delta[j] + position
        const bool unit = __all_sync(0xffffffffu, my_cb <= 1u);   // every own count is 1: skip one shuffle
        updates += my_len;
        uint32_t started = 0;
        for (uint32_t base = 0; base < W; base += 32 * UNROLL) {
            uint32_t a[UNROLL], cb[UNROLL], p[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t rel = P - (base + 32 * u);
                const uint32_t m = __reduce_or_sync(0xffffffffu, (rel < 32u && my_len) ? (1u << rel) : 0u);
                const uint32_t j = (started + __popc(m & lane_le) - 1) & 31;
                started += __popc(m);
                a[u] = __shfl_sync(0xffffffffu, delta, j) + base + 32 * u + lane;
                cb[u] = unit ? 1u : __shfl_sync(0xffffffffu, my_cb, j);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) p[u] = (base + 32 * u + lane < W) ? ep[a[u]] : 0u;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (base + 32 * u + lane < W) atomicAdd(&row[p[u] & idmask], (p[u] >> idbits) * cb[u]);
        }
    }
    __syncthreads();
    AccT* __restrict__ Krow = K + (size_t)group * k_group_stride + ((size_t)b * (b + 1) >> 1);
    for (uint32_t i = threadIdx.x; i <= b; i += blockDim.x) {
        const uint32_t v = row[i];
        if (v) Krow[i] += (AccT)v;
    }
    if (stat_counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) updates += __shfl_xor_sync(0xffffffffu, updates, o);
        if (lane == 0 && updates) atomicAdd(&stat_counters[2], updates);
    }
}

// ------------------------------------------------------------------------------------------
// normalisation (fastsk_kernel.cpp:96-103): out = K_ij / sqrt(K_ii * K_jj) with IEEE mul, sqrt,
// div (bit-identical to the reference's x86-64 doubles); the diagonal formula K_ii/sqrt(K_ii*K_ii)
// is the same expression.
template <typename T>
__global__ void diag_kernel(const T* __restrict__ K, int64_t n, double* __restrict__ diag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) diag[i] = (double)K[(i * (i + 1) >> 1) + i];
}

__device__ __forceinline__ double norm_entry(double v, double di, double dj) {
    return __ddiv_rn(v, __dsqrt_rn(__dmul_rn(di, dj)));
}

// square block rows [r0, r0+nr) x cols [0, nc) of the symmetric matrix -> out (row-major, ld = nc).
// 32x32 tiles; tiles above the diagonal are read transposed through shared memory so that the
// packed triangle is always read along its contiguous direction.
template <typename T>
__global__ void __launch_bounds__(256)
normalise_block_kernel(const T* __restrict__ K, const double* __restrict__ diag, int64_t r0, int64_t nr, int64_t nc,
                       double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int64_t ti = (int64_t)blockIdx.y * 32, tj = (int64_t)blockIdx.x * 32;   // tile origin (row offset within block, col)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool upper = (r0 + ti + 31) < tj;   // whole tile strictly above the diagonal: read mirrored
    if (!upper) {
        for (int y = ty; y < 32; y += 8) {
            const int64_t i = r0 + ti + y, j = tj + tx;
            if (ti + y < nr && j < nc) {
                const int64_t a = i >= j ? i : j, b = i >= j ? j : i;
                out[(ti + y) * nc + j] = norm_entry((double)K[(a * (a + 1) >> 1) + b], diag[i], diag[j]);
            }
        }
    } else {
        for (int y = ty; y < 32; y += 8) {      // read K[j][i] with threads along i (contiguous)
            const int64_t j = tj + y, i = r0 + ti + tx;
            if (ti + tx < nr && j < nc) tile[y][tx] = norm_entry((double)K[(j * (j + 1) >> 1) + i], diag[i], diag[j]);
        }
        __syncthreads();
        for (int y = ty; y < 32; y += 8) {
            const int64_t j = tj + tx;
            if (ti + y < nr && j < nc) out[(ti + y) * nc + j] = tile[tx][y];
        }
    }
}

template <typename T>
__global__ void normalise_packed_kernel(const T* __restrict__ K, const double* __restrict__ diag, int64_t n,
                                        double* __restrict__ out) {
    const int64_t i = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= i; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = (i * (i + 1) >> 1) + j;
        out[p] = norm_entry((double)K[p], diag[i], diag[j]);
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
__global__ void __launch_bounds__(256)
welford_kernel(AccT* __restrict__ Ks, double* __restrict__ K_hat, int64_t n_pairs, int64_t n_train_pairs, int iter,
               double* __restrict__ block_sums) {
    __shared__ double ws[8];
    const double diter = (double)iter;
    double acc = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
        const double ks = (double)Ks[p];
        Ks[p] = 0;
        double kh = K_hat[p];
        const double delta = __dsub_rn(ks, kh);
        kh = __dadd_rn(kh, __ddiv_rn(delta, diter));
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

__global__ void welford_final_kernel(const double* __restrict__ block_sums, int nblocks, double* __restrict__ out) {
    __shared__ double ws[32];
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
