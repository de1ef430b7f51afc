// Segmentation of the sorted records, register-blocked form (shared.cpp:280-315).
//
// Same outputs as segment_kernel (fsk_kernels.cuh) -- the aligned id stream, and one task per record (task form) or one
// directory entry per (run, block of rows) group (directory form) -- for records that carry the sequence id (32- or 64-bit
// key << idbits | id).  What changed is who does the work: ncu showed segment_kernel issue-bound, ~190 instructions per
// record (profiles/r02_ncu_segment_*): one record per lane means every neighbour comparison is a shuffle and every run
// boundary a ballot.  Here a thread owns 48 bytes of CONSECUTIVE records (12 x 32-bit, loaded as three aligned 16-byte
// vectors), so neighbour comparisons, run starts and the running sum of padded run lengths are register arithmetic, and the
// warp only meets for two scans (last run head: max; padded lengths: add).  The ids go through a shared-memory image of the
// tile's stretch of the id stream -- fill included -- and leave as full 16-byte stores: the separate 0xFF memset of the
// stream is gone, and so is its second pass over the memory.
#pragma once
#include "fsk_kernels.cuh"

namespace fsk {

constexpr int LEAN_THREADS = 256;
__host__ __device__ constexpr int lean_rpt(int rec_bytes) { return 48 / rec_bytes; }                 // records per thread
__host__ __device__ constexpr int lean_tile(int rec_bytes) { return LEAN_THREADS * lean_rpt(rec_bytes); }
__host__ __device__ constexpr int lean_cap(int rec_bytes) { return lean_tile(rec_bytes) * 3 / 2; }   // ids in the image

template <typename RecT, typename IdT, bool DIR, bool HEAVY, bool STATS>
__global__ void __launch_bounds__(LEAN_THREADS, 4)
segment_lean_kernel(const RecT* __restrict__ rec, uint32_t n, uint32_t tiles_per_slot, size_t ids_stride, int idbits, uint32_t nseq,
                    int unit_shift, uint32_t pad_mask, uint32_t* __restrict__ fill, IdT* __restrict__ ids, uint2* __restrict__ task,
                    uint32_t* __restrict__ scan_status /* [slot][tile] */, uint32_t* __restrict__ ticket,
                    uint32_t* __restrict__ unsorted_flag, unsigned long long* __restrict__ stat_counters, uint32_t heavy_tau,
                    uint32_t* __restrict__ heavy_count, uint2* __restrict__ heavy_list, uint32_t heavy_cap,
                    uint32_t* __restrict__ heavy_bits, size_t heavy_bits_stride, uint2* __restrict__ tdir, int dir_bshift,
                    uint32_t dir_nb, int dir_keybits) {
    constexpr int RPT = lean_rpt(sizeof(RecT));
    constexpr int TILE = lean_tile(sizeof(RecT));
    constexpr int CAP = lean_cap(sizeof(RecT));
    constexpr int PER = 16 / sizeof(IdT);
    constexpr int VEC = 16 / sizeof(RecT);                     // records per 16-byte vector
    constexpr int NWARP = LEAN_THREADS / 32;
    extern __shared__ __align__(16) unsigned char lean_smem[];
    IdT* image = reinterpret_cast<IdT*>(lean_smem);
    __shared__ uint32_t s_ticket, s_wmax[NWARP], s_wsum[NWARP], s_carry, s_base, s_p0, s_p1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_ticket = atomicAdd(ticket, 1u); s_p0 = 0xffffffffu; s_p1 = 0; s_carry = 0; }
    {   // the image starts as fill: whatever no record overwrites is the padding between the runs
        uint4* im = reinterpret_cast<uint4*>(image);
        const uint4 ff = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        for (int u = tid; u < CAP / PER; u += LEAN_THREADS) im[u] = ff;
    }
    __syncthreads();
    const uint32_t slot = s_ticket / tiles_per_slot;           // slot-major: the slot's arrays stay L2-resident
    const uint32_t tile = s_ticket - slot * tiles_per_slot;
    const size_t sbase = (size_t)slot * n;
    const RecT* __restrict__ R = rec + sbase;
    // tiles are laid over the slot's records from the 16-byte boundary at or before its first record, so that every
    // thread's three vectors are aligned loads; `mis` virtual records precede record 0
    const int mis = (int)((reinterpret_cast<uintptr_t>(R) & 15u) / sizeof(RecT));
    const int N_ = (int)n;
    const int i0 = (int)(tile * TILE) + tid * RPT - mis;       // index of this thread's first record (may be < 0 or >= n)
    const RecT idmask = (((RecT)1 << idbits) - 1);

    RecT r[RPT];
    {
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(R + i0);
#pragma unroll
        for (int q = 0; q < RPT / VEC; ++q) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (i0 + q * VEC < N_ && i0 + (q + 1) * VEC > 0) v = __ldg(src + q);     // a vector may reach a few records outside
            if (sizeof(RecT) == 4) {                                                   // the slot: those are masked below
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < VEC; ++e) r[q * VEC + e] = (RecT)w[e & 3];
            } else {
                const unsigned long long w[2] = {((unsigned long long)v.y << 32) | v.x, ((unsigned long long)v.w << 32) | v.z};
#pragma unroll
                for (int e = 0; e < VEC; ++e) r[q * VEC + e] = (RecT)w[e & 1];
            }
        }
    }
    // the records just before and after this thread's
    RecT pv = __shfl_up_sync(0xffffffffu, r[RPT - 1], 1), nx = __shfl_down_sync(0xffffffffu, r[0], 1);
    if (lane == 0) pv = (i0 > 0 && i0 - 1 < N_) ? R[i0 - 1] : (RecT)0;
    if (lane == 31) nx = (i0 + RPT < N_ && i0 + RPT >= 0) ? R[i0 + RPT] : (RecT)0;

    uint32_t vb = 0, hb = 0, tl = 0, bt = 0, eq = 0;           // per-record bits: valid, run head, run tail, block-group tail, equals predecessor
    bool unsorted = false;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int i = i0 + j;
        const bool valid = i >= 0 && i < N_;
        const RecT p = j == 0 ? pv : r[j - 1], x = j == RPT - 1 ? nx : r[j + 1];
        const bool first = i == 0, last = i + 1 >= N_;
        const bool head = valid && (first || ((p ^ r[j]) >> idbits) != 0);
        const bool tail = valid && (last || ((x ^ r[j]) >> idbits) != 0);
        unsorted |= valid && !first && p > r[j];
        vb |= (uint32_t)valid << j;
        hb |= (uint32_t)head << j;
        tl |= (uint32_t)tail << j;
        if (DIR) bt |= (uint32_t)(tail || (valid && ((uint32_t)(x & idmask) >> dir_bshift) != ((uint32_t)(r[j] & idmask) >> dir_bshift))) << j;
        if (STATS || !DIR) eq |= (uint32_t)(valid && !first && p == r[j]) << j;
    }
    if (unsorted) *unsorted_flag = 1u;      // the sort must have left the records non-decreasing: see onesweep_kernel

    // ---- scan 1: the last run head at or before every record (max over positions + 1; 0 = none yet)
    const uint32_t lhp = hb ? (uint32_t)(i0 + 31 - __clz(hb)) + 1u : 0u;
    uint32_t inc = lhp;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    uint32_t exm = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exm = 0;
    if (lane == 31) s_wmax[warp] = inc;
    if (warp == 0) {
        // run start of the tile's first record when its run began in an earlier tile: walk back in blocks of 32 (typical runs
        // are short), then lower_bound over the sorted records for very long runs
        const int ifirst = max((int)(tile * TILE) - mis, 0);
        uint32_t p = (uint32_t)ifirst;
        if (ifirst > 0 && ifirst < N_) {
            const RecT r_first = R[ifirst];
            if (((R[ifirst - 1] ^ r_first) >> idbits) == 0) {
                bool found = false;
                for (int step = 0; step < 8 && p > 0 && !found; ++step) {
                    const uint32_t jj = p - 1 - lane;
                    const bool inb = (uint32_t)lane < p;
                    const bool differs = inb && ((R[inb ? jj : 0] ^ r_first) >> idbits) != 0;
                    const uint32_t dm = __ballot_sync(0xffffffffu, differs);
                    if (dm) { p = p - (__ffs(dm) - 1); found = true; }
                    else p = p > 32 ? p - 32 : 0;
                }
                if (!found && p > 0) {
                    uint32_t lo = 0, hi = p;
                    while (lo < hi) {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        if ((R[mid] >> idbits) < (r_first >> idbits)) lo = mid + 1;
                        else hi = mid;
                    }
                    p = lo;
                }
            }
        }
        if (lane == 0) s_carry = p + 1u;
    }
    __syncthreads();
    uint32_t rs_in = max(exm, s_carry);
#pragma unroll
    for (int w = 0; w < NWARP; ++w)
        if (w < warp) rs_in = max(rs_in, s_wmax[w]);
    rs_in -= 1u;                                               // (s_carry >= 1 always)

    // run start of every record, and the padded lengths of the runs that end in this thread's records
    uint32_t rs[RPT], xoff[RPT];
    uint32_t E = 0;
    {
        uint32_t cur = rs_in;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            if ((hb >> j) & 1u) cur = (uint32_t)(i0 + j);
            rs[j] = cur;
            xoff[j] = E;
            if ((tl >> j) & 1u) E += ((uint32_t)(i0 + j) - cur + 1u + pad_mask) & ~pad_mask;
        }
    }
    // ---- scan 2: position of every record in the aligned id stream
    uint32_t einc = E;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, einc, o);
        if (lane >= o) einc += t;
    }
    if (lane == 31) s_wsum[warp] = einc;
    __syncthreads();
    uint32_t ex = einc - E, tile_total = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
        const uint32_t t = s_wsum[w];
        if (w < warp) ex += t;
        tile_total += t;
    }
    uint32_t* __restrict__ st = scan_status + (size_t)slot * tiles_per_slot;
    if (tid == 0 && tile > 0) st_volatile_u32(st + tile, tile_total | 0x40000000u);

    // task form: the tasks' places in their rows' lists (one atomic with return per record), in flight under the look-back
    uint32_t tpos[DIR ? 1 : RPT];
    if (!DIR) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            tpos[j] = 0;
            if ((vb >> j) & 1u) tpos[j] = atomicAdd(&fill[(size_t)slot * nseq + (uint32_t)(r[j] & idmask)], 1u);
        }
    }
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
            s_base = excl;
            st_volatile_u32(st + tile, (excl + tile_total) | 0x80000000u);
        }
    }
    __syncthreads();
    const uint32_t base = s_base + ex;

    // the tile's stretch [p0, p1) of the id stream: from its first record to where the next tile's first record goes
    {
        const int ifirst = max((int)(tile * TILE) - mis, 0), ilast = min((int)((tile + 1) * TILE) - mis, N_) - 1;
        if (i0 <= ifirst && ifirst < i0 + RPT || i0 <= ilast && ilast < i0 + RPT) {
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                if (!((vb >> j) & 1u)) continue;
                const uint32_t xrun = base + xoff[j], off = (uint32_t)(i0 + j) - rs[j];
                if (i0 + j == ifirst) s_p0 = xrun + off;
                if (i0 + j == ilast) s_p1 = ((tl >> j) & 1u) ? xrun + ((off + 1u + pad_mask) & ~pad_mask) : xrun + off + 1u;
            }
        }
    }
    // tasks / directory entries, straight to global memory
    unsigned long long updates = 0;
    uint32_t n_groups = 0;
    if (!DIR || bt || (HEAVY && hb) || STATS) {
        // group ends (task form, statistics): the first record-level tail at or after every record
        uint32_t ge[(!DIR) ? RPT : 1];
        if (!DIR) {
            uint32_t nxt = 0xffffffffu;
#pragma unroll
            for (int j = RPT - 1; j >= 0; --j) {
                const RecT x = j == RPT - 1 ? nx : r[j + 1];
                const bool rtail = ((vb >> j) & 1u) && (i0 + j + 1 >= N_ || x != r[j]);
                if (rtail) nxt = (uint32_t)(i0 + j);
                ge[j] = nxt;
            }
        }
        uint32_t go = 0;                                       // records equal to this one immediately before it
        if (STATS && (eq & 1u)) {                              // the group began before this thread: count back (groups are short)
            int g = i0 - 1;
            while (g >= 0 && R[g] == r[0]) { ++go; --g; }
            go -= 1u;                                          // the loop below adds one for j = 0
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            if (!((vb >> j) & 1u)) continue;
            const uint32_t i = (uint32_t)(i0 + j);
            const uint32_t xrun = base + xoff[j];
            const bool writes = !DIR || ((bt >> j) & 1u);
            uint32_t ln;
            if (DIR) ln = i - rs[j] + 1u;
            else {
                uint32_t g = ge[j];
                if (g == 0xffffffffu) {                        // the group runs past this thread's records (rare)
                    g = (uint32_t)(i0 + RPT - 1);
                    while (g + 1 < n && R[g + 1] == r[j]) ++g;
                }
                ln = g - rs[j] + 1u;
            }
            if (STATS) {
                go = ((eq >> j) & 1u) ? go + 1u : 0u;
                updates += (i - rs[j]) + go + 1u;              // sum over a run = (len^2 + sum of group sizes^2) / 2: its pair updates
                n_groups += go == 0u;
            }
            if (HEAVY && heavy_tau) {
                // a run of more than heavy_tau records is a (nearly) dense column of the count matrix: its d^2/2 updates go to the
                // tensor-core contraction if the batch's list has room, and then the accumulate skips its tasks (segment_kernel)
                const uint32_t far = rs[j] + heavy_tau;
                if ((writes || i == rs[j]) && far < n && ((R[far] ^ r[j]) >> idbits) == 0) {
                    ln |= 0x80000000u;
                    if (i == rs[j]) {
                        const uint32_t at = atomicAdd(heavy_count, 1u);
                        if (at < heavy_cap) {
                            heavy_list[at] = make_uint2(slot, i);
                            const uint32_t xu = xrun >> unit_shift;
                            atomicOr(&heavy_bits[(size_t)slot * heavy_bits_stride + (xu >> 5)], 1u << (xu & 31u));
                        }
                    }
                }
            }
            if (DIR) {
                if (writes) {
                    const uint32_t key = (uint32_t)(r[j] >> idbits);
                    tdir[(((size_t)slot * dir_nb + ((uint32_t)(r[j] & idmask) >> dir_bshift)) << dir_keybits) + key] = make_uint2(xrun >> unit_shift, ln);
                }
            } else {
                task[sbase + tpos[j]] = make_uint2(xrun >> unit_shift, ln);
            }
        }
    }
    __syncthreads();
    // ids -> image -> global, one window of CAP ids at a time (one window unless the tile holds many short runs)
    const uint32_t p0 = s_p0, p1 = s_p1;
    if (p0 != 0xffffffffu) {
        IdT* __restrict__ out = ids + (size_t)slot * ids_stride;
        const uint32_t a0 = p0 & ~(uint32_t)(PER - 1);
        for (uint32_t w0 = a0; w0 < p1; w0 += CAP) {
            if (w0 != a0) {
                __syncthreads();
                uint4* im = reinterpret_cast<uint4*>(image);
                const uint4 ff = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                for (int u = tid; u < CAP / PER; u += LEAN_THREADS) im[u] = ff;
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const uint32_t pos = base + xoff[j] + ((uint32_t)(i0 + j) - rs[j]) - w0;      // wraps below the window
                if (((vb >> j) & 1u) && pos < (uint32_t)CAP) image[pos] = (IdT)(r[j] & idmask);
            }
            __syncthreads();
            const uint32_t nunits = (min(p1 - w0, (uint32_t)CAP) + PER - 1) / PER;
            for (uint32_t u = tid; u < nunits; u += LEAN_THREADS) {
                const uint32_t e0 = w0 + u * PER;
                if (e0 >= p0 && e0 + PER <= p1) {
                    reinterpret_cast<uint4*>(out + e0)[0] = reinterpret_cast<const uint4*>(image)[u];
                } else {                                        // a unit shared with the neighbouring tile: only this tile's cells
#pragma unroll
                    for (int e = 0; e < PER; ++e)
                        if (e0 + e >= p0 && e0 + e < p1) out[e0 + e] = image[u * PER + e];
                }
            }
        }
    }
    if (tile == 0 && tid == 0)      // the unit that stands in for the tasks of runs the tensor cores took: always fill
        reinterpret_cast<uint4*>(ids + (size_t)slot * ids_stride)[ids_stride / PER - 1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (STATS && stat_counters) {
        uint32_t n_runs = __popc(hb);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            updates += __shfl_xor_sync(0xffffffffu, updates, o);
            n_runs += __shfl_xor_sync(0xffffffffu, n_runs, o);
            n_groups += __shfl_xor_sync(0xffffffffu, n_groups, o);
        }
        if (lane == 0) {
            atomicAdd(&stat_counters[0], (unsigned long long)n_groups);
            atomicAdd(&stat_counters[1], (unsigned long long)n_runs);
            atomicAdd(&stat_counters[2], updates);
        }
    }
}

}  // namespace fsk
