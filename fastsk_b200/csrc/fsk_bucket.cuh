// Fused last sort pass + segmentation for keys of two radix digits (9..16 key bits, 32-bit records, 16-bit ids: the
// shape of BASELINE configs[3]).  Replaces the second onesweep pass (shared.cpp:156-191), the gather of the sorted
// records (fastsk_kernel.cpp:233-238) and segment_kernel (shared.cpp:280-315) by ONE kernel:
//
//   the only onesweep pass partitions the records by their HIGH digit (stable), so each of the <= 256 buckets of a slot
//   is contiguous in HBM and holds whole runs: a run is one (bucket, low digit) pair.  A CTA takes one bucket:
//     A  each warp counts the low digits of its contiguous chunk of the bucket (shared-memory histograms),
//     B  a scan gives every run its length, its start in sorted order and its ALIGNED start in the id stream,
//     C  the chunks are read again (from L2) and every record's sequence id goes straight to its sorted, aligned place
//        in a shared-memory image of the bucket's id stream (rank = one shared atomic with return per record, the same
//        optimistic lane-order ranking as onesweep_kernel, verified in D),
//     D  every record of every run files its task (run start, prefix length up to the end of its own group) under its
//        sequence, exactly as segment_kernel does, and the image is copied to HBM in 16-byte pieces.
//   The sorted records themselves are never written: per record 4 B read from HBM, 4 B from L2, 2 B + 8 B written,
//   instead of 4 + 4 (pass two) + 4 + 2 + 8 (segment) + the 0xFF fill of the id stream.
//
// Bucket b of a slot owns the id-stream range starting at roundup64(first sorted position of b) + b * (64 * 2^lo_bits + 64),
// which can never overlap the next bucket's (a bucket's runs need at most 63 pad cells each), so no scan across CTAs is
// needed.  A bucket too large for the shared-memory image (skewed data) builds its image directly in HBM with the same
// code; only the final copy is skipped.
#pragma once
#include "fsk_kernels.cuh"

namespace fsk {

constexpr int BK_THREADS = 512;              // two CTAs per SM: one bucket's phases overlap the other's
constexpr int BK_WARPS = BK_THREADS / 32;
constexpr int BK_ILP = 16;
constexpr int BK_DILP = 16;                 // records (= task-slot atomics in flight) per thread in the filing phase                  // record loads in flight per lane in the two streaming phases

// id-stream cells between the ranges of consecutive buckets beyond the bucket's own records: 63 pad cells for each of its
// 2^lo_bits runs at most, + 63 for rounding the bucket's start up to a 128-byte line
__host__ __device__ constexpr size_t bucket_id_stride(int lo_bits) { return ((size_t)64 << lo_bits) + 64; }
// shared memory: warp histograms, run tables, the image, and the run number of every 64-id block of the image
__host__ __device__ constexpr size_t bucket_smem_bytes(uint32_t image_ids) {
    return (size_t)BK_WARPS * RADIX * 4 + 3 * RADIX * 4 + (size_t)image_ids * 2 + ((size_t)(image_ids / 32 + 8) * 2 + 15) / 16 * 16 + 64;
}

template <bool STATS>
__global__ void __launch_bounds__(BK_THREADS, 2)
bucket_segment_kernel(const uint32_t* __restrict__ rec, uint32_t n, const uint32_t* __restrict__ ghist_hi /* [slot][MAX_PASS][RADIX], pre-offset to the high digit */,
                      int idbits, int lo_shift, int lo_bits, uint32_t nseq, size_t ids_stride, uint32_t pad_mask,
                      uint32_t image_cap /* ids that fit the shared-memory image */, uint32_t* __restrict__ fill,
                      uint16_t* __restrict__ ids, uint2* __restrict__ task, uint32_t* __restrict__ unsorted_flag,
                      unsigned long long* __restrict__ stat_counters) {
    extern __shared__ __align__(16) unsigned char bk_smem[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(bk_smem);            // [BK_WARPS][RADIX]: counts, then sorted-position cursors
    uint32_t* runS = hist + BK_WARPS * RADIX;                         // first sorted position of run d inside the bucket
    uint32_t* runX = runS + RADIX;                                    // aligned start of run d inside the bucket's id range
    uint32_t* runL = runX + RADIX;                                    // length of run d
    uint16_t* image = reinterpret_cast<uint16_t*>(runL + RADIX);      // the bucket's id stream (16-byte aligned)
    uint16_t* block_run = image + image_cap;                          // run that owns the i-th block of pad_mask + 1 ids of the image
    __shared__ uint32_t s_start, s_warp_tot[2][8], s_total_x;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bucket = blockIdx.x, slot = blockIdx.y;
    const uint32_t DL = 1u << lo_bits, dmask = DL - 1u;
    const uint32_t idmask = (1u << idbits) - 1u;
    const uint32_t* __restrict__ gh = ghist_hi + (size_t)slot * MAX_PASS * RADIX;
    const uint32_t nbk = gh[bucket];

    // first sorted position of the bucket: sum of the counts of the lower high digits
    if (tid < 32) {
        uint32_t s = 0;
        for (uint32_t d = lane; d < bucket; d += 32) s += gh[d];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_start = s;
    }
    for (int i = tid; i < BK_WARPS * RADIX; i += BK_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t start = s_start;
    const uint32_t* __restrict__ R = rec + (size_t)slot * n + start;
    // warp w owns records [w * chunk, (w + 1) * chunk) of the bucket, in order (stability)
    const uint32_t chunk = ((nbk + BK_WARPS - 1) / BK_WARPS + 31u) & ~31u;
    const uint32_t c0 = min(nbk, (uint32_t)warp * chunk), c1 = min(nbk, c0 + chunk);

    // A: low-digit histogram of the warp's chunk
    uint32_t* wh = hist + warp * RADIX;
    for (uint32_t i0 = c0; i0 < c1; i0 += 32 * BK_ILP) {          // BK_ILP independent loads in flight per lane
        uint32_t r[BK_ILP];
#pragma unroll
        for (int u = 0; u < BK_ILP; ++u) {
            const uint32_t i = i0 + u * 32 + lane;
            r[u] = i < c1 ? R[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < BK_ILP; ++u)
            if (i0 + u * 32 + lane < c1) atomicAdd(&wh[(r[u] >> lo_shift) & dmask], 1u);
    }
    __syncthreads();

    // B: run lengths, sorted starts, aligned starts; hist[w][d] becomes the sorted position of warp w's first record of d
    uint32_t L = 0;
    if (tid < (int)DL) {
        uint32_t run = 0;
#pragma unroll 8
        for (int w = 0; w < BK_WARPS; ++w) {
            const uint32_t c = hist[w * RADIX + tid];
            hist[w * RADIX + tid] = run;
            run += c;
        }
        L = run;
    }
    {   // two exclusive scans over the <= 256 runs: lengths and padded lengths
        const uint32_t Lp = (L + pad_mask) & ~pad_mask;
        uint32_t a = L, b = Lp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += ta; b += tb; }
        }
        if (warp < 8 && lane == 31) { s_warp_tot[0][warp] = a; s_warp_tot[1][warp] = b; }
        __syncthreads();
        if (tid < (int)DL) {
            uint32_t ba = 0, bb = 0;
            for (int w = 0; w < warp; ++w) { ba += s_warp_tot[0][w]; bb += s_warp_tot[1][w]; }
            runS[tid] = ba + a - L;
            runX[tid] = bb + b - Lp;
            runL[tid] = L;
            if (tid == (int)DL - 1) s_total_x = bb + b;
        }
    }
    __syncthreads();
    const uint32_t total_x = s_total_x;
    if (tid < (int)DL) {
        const uint32_t s0 = runS[tid];
#pragma unroll 8
        for (int w = 0; w < BK_WARPS; ++w) hist[w * RADIX + tid] += s0;
    }
    // id-stream range of this bucket (multiple of 64 ids: 128-byte aligned)
    const size_t xbase = (size_t)((start + 63u) & ~63u) + (size_t)bucket * bucket_id_stride(lo_bits);
    uint16_t* __restrict__ gids = ids + (size_t)slot * ids_stride + xbase;
    const uint32_t xunit0 = (uint32_t)(xbase >> 3);
    const uint32_t blk_shift = 31 - __clz(pad_mask + 1u);              // runs start on multiples of pad_mask + 1 = 2^blk_shift ids
    unsigned long long updates = 0;
    uint32_t groups = 0;
    auto run_end = [&](uint32_t d) -> uint32_t { return d + 1 < DL ? runX[d + 1] : total_x; };   // runX is the scan of the padded lengths

    // The runs [d0, d1) whose padded ids fit the shared-memory image are built together: for the synthetic workload that is
    // the whole bucket at once; a skewed bucket takes several rounds (its records are read again from L2 in each), and a
    // single run longer than the image is built directly in HBM.
    uint32_t d0 = 0;
    while (d0 < DL) {
        const uint32_t base = runX[d0];
        uint32_t d1;
        if (total_x - base <= image_cap) {
            d1 = DL;
        } else {                                                     // largest d1 with run_end(d1 - 1) - base <= image_cap
            uint32_t lo = d0, hi = DL;                               // invariant: runs [d0, lo) fit, run hi - 1 ... may not
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (run_end(mid) - base <= image_cap) lo = mid + 1;
                else hi = mid;
            }
            d1 = lo;
        }
        const bool in_hbm = d1 == d0;                                // run d0 alone exceeds the image
        if (in_hbm) d1 = d0 + 1;
        const uint32_t span = run_end(d1 - 1) - base;
        uint16_t* sid = in_hbm ? gids + base : image;                // generic pointer
        if (!in_hbm && tid >= (int)d0 && tid < (int)d1) {
            const uint32_t X = runX[tid] - base, Xe = run_end(tid) - base;
            for (uint32_t blk = X >> blk_shift; blk < Xe >> blk_shift; ++blk) block_run[blk] = (uint16_t)tid;
        }
        __syncthreads();

        // C: sequence ids of the round's runs to their sorted, aligned places
        for (uint32_t i0 = c0; i0 < c1; i0 += 32 * BK_ILP) {
            uint32_t r[BK_ILP];
#pragma unroll
            for (int u = 0; u < BK_ILP; ++u) {
                const uint32_t i = i0 + u * 32 + lane;
                r[u] = i < c1 ? R[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < BK_ILP; ++u) {                        // in record order: the ranks must follow the input order
                const uint32_t d = (r[u] >> lo_shift) & dmask;
                if (i0 + u * 32 + lane < c1 && d >= d0 && d < d1) {
                    const uint32_t p = atomicAdd(&wh[d], 1u);         // lanes with equal d are served in lane order (verified in D)
                    sid[runX[d] - base + (p - runS[d])] = (uint16_t)(r[u] & idmask);
                }
            }
        }
        __syncthreads();

        // D: tasks.  Record j of a run of sequence b adds the prefix [run start, last record of b's group].
        if (!in_hbm) {
            // every thread walks the image (load balanced): position a belongs to run block_run[a >> blk_shift].  BK_DILP records
            // per thread and step: the global atomics that hand out the task slots are the long-latency part, so all of a step's
            // atomics are in flight together (pv = sequence id before, task slot after)
            for (uint32_t a0 = tid; a0 < span; a0 += BK_DILP * BK_THREADS) {
                uint32_t pv[BK_DILP], ln[BK_DILP];
#pragma unroll
                for (int u = 0; u < BK_DILP; ++u) {
                    const uint32_t a = a0 + u * BK_THREADS;
                    ln[u] = 0;
                    pv[u] = 0;
                    if (a < span) {
                        const uint32_t d = block_run[a >> blk_shift];
                        const uint32_t X = runX[d] - base, Lr = runL[d], j = a - X;
                        if (j < Lr) {
                            const uint32_t id = image[a];
                            uint32_t e = j;
                            while (e + 1 < Lr && image[X + e + 1] == id) ++e;
                            if (e + 1 < Lr && image[X + e + 1] < id) *unsorted_flag = 1u;
                            ln[u] = e + 1;
                            pv[u] = id;
                            if (STATS) { updates += e + 1; groups += (e == j); }
                        } else if (j < ((Lr + 7u) & ~7u)) {
                            image[a] = 0xffffu;      // the rest of the run's last 16-byte unit reads as "no sequence"
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < BK_DILP; ++u)
                    if (ln[u]) pv[u] = atomicAdd(&fill[(size_t)slot * nseq + pv[u]], 1u);
#pragma unroll
                for (int u = 0; u < BK_DILP; ++u)
                    if (ln[u]) {
                        const uint32_t X = runX[block_run[(a0 + u * BK_THREADS) >> blk_shift]];
                        task[(size_t)slot * n + pv[u]] = make_uint2(xunit0 + (X >> 3), ln[u]);
                    }
            }
            __syncthreads();
            // image -> HBM, 16 bytes per thread and step (xbase and the padded run starts are multiples of 8 ids)
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>(image);
            uint4* __restrict__ dst = reinterpret_cast<uint4*>(gids + base);
            const uint32_t nvec = (span + 7u) >> 3;
            for (uint32_t i = tid; i < nvec; i += BK_THREADS) dst[i] = src[i];
        } else {
            const uint32_t Lr = runL[d0];
            for (uint32_t j = tid; j < Lr; j += BK_THREADS) {
                const uint32_t id = sid[j];
                uint32_t e = j;
                while (e + 1 < Lr && sid[e + 1] == id) ++e;
                if (e + 1 < Lr && sid[e + 1] < id) *unsorted_flag = 1u;
                const uint32_t pos = atomicAdd(&fill[(size_t)slot * nseq + id], 1u);
                task[(size_t)slot * n + pos] = make_uint2(xunit0 + (base >> 3), e + 1);
                if (STATS) { updates += e + 1; groups += (e == j); }
            }
            const uint32_t pend = (Lr + 7u) & ~7u;
            if (Lr + tid < pend) sid[Lr + tid] = 0xffffu;
        }
        __syncthreads();                                             // the next round reuses the image and block_run
        d0 = d1;
    }
    if (STATS) {
        uint32_t nruns = (tid < (int)DL && runL[tid] > 0) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            updates += __shfl_xor_sync(0xffffffffu, updates, o);
            groups += __shfl_xor_sync(0xffffffffu, groups, o);
            nruns += __shfl_xor_sync(0xffffffffu, nruns, o);
        }
        if (lane == 0) {
            atomicAdd(&stat_counters[0], (unsigned long long)groups);
            atomicAdd(&stat_counters[1], (unsigned long long)nruns);
            atomicAdd(&stat_counters[2], updates);
        }
    }
}

}  // namespace fsk
