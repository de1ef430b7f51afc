// Dense regime of the accumulate (regime 2): when a combination has few distinct k-mers
// (alphabet^k at most a few thousand: DNA with k = g - m <= 6, the TFBS configurations of the reference) the
// count matrix C[sequence][k-mer] is dense and countAndUpdateTri's K[i][j] += c_i * c_j (shared.cpp:304-327)
// is the contraction K += C * C^T.  Then neither the sort nor the run segmentation is needed:
//
//   dense_count_kernel   C[seq][slot * nks + key] = number of windows of `seq` whose kept characters pack to `key`
//                        (the gather of fastsk_kernel.cpp:224-228 + the counting of shared.cpp:304-315), fp16
//   syrk_tc_kernel       lower-triangle tiles of C * C^T on the 5th-generation tensor cores: TMA (128-byte swizzle)
//                        -> shared memory -> tcgen05.mma kind::f16 with fp32 accumulators in TMEM -> tcgen05.ld ->
//                        integer RED into the packed int64 triangle.  All slots of a batch are one GEMM: the
//                        contraction dimension is (slots x k-mers).
//
// Exactness: counts are integers <= maxwin <= 2048 (exact in fp16), products < 2^22, and the host cuts the
// contraction dimension so that no fp32 accumulator can exceed 2^24 (slots per GEMM <= 2^24 / maxwin^2): every
// intermediate is an exactly representable integer, so the result is bit-identical to the integer sum.
//
// Byte operands (round 2): when no sequence has more than 255 windows every count fits a byte, and the same kernels run
// with U8 = true -- C (and the heavy runs' H) as uint8, tcgen05.mma kind::i8 with int32 accumulators, 128 k-mer columns
// per 128-byte swizzle row instead of 64: half the operand bytes (what bounds these short-K contractions: L2 -> shared
// memory at ~43 B/clk and SM), twice the MMA rate, integer accumulation exact by construction.
//
//   syrk_tc_welford_kernel   variance mode: a CTA owns one tile of one virtual stream's running mean over all the
//                        iterations of a launch group; <true>: the means stay in registers (see there).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "fsk_kernels.cuh"

namespace fsk {

constexpr int DG_TILE = 128;                                     // output tile: 128 x 128 pairs of sequences
constexpr int DG_BK = 64;                                        // fp16 elements per k-block = one 128-byte swizzle row
constexpr int DG_THREADS = 192;                                  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr uint32_t DG_TILE_BYTES = DG_TILE * DG_BK * 2;          // 16 KB
// Two shapes of the same kernel.  NA = 1: one 128 x 128 output tile per CTA, 3 stages x 32 KB, two CTAs per SM (one's epilogue
// runs under the other's MMAs): the default (faster at every N measured in round 2, profiles/r02_gemm_shape_by_n.txt).  NA = 2: two vertically adjacent tiles per CTA share their
// B operand (3 loads feed 8 MMAs per k-block: 24 KB of L2 -> shared-memory traffic per M-MAC instead of 32), 4 stages x 48 KB,
// one CTA per SM, 256 TMEM columns: opt-in (gemm_shape = 2) and the heavy-run contraction.
__host__ __device__ constexpr int dg_stages(int NA) { return NA == 1 ? 3 : 4; }
__host__ __device__ constexpr uint32_t dg_stage_bytes(int NA) { return (uint32_t)(NA + 1) * DG_TILE_BYTES; }
__host__ __device__ constexpr size_t dg_smem(int NA) {
    return (size_t)dg_stages(NA) * dg_stage_bytes(NA) + 1024 /* 1024-byte alignment of the swizzle atom */ + 128;
}
constexpr int DENSE_MAX_KEYS = 4096;

// ------------------------------------------------------------------------------------------
// C = per-sequence k-mer counts of every slot.  One CTA per sequence, one warp per (sequence, slot): the warp counts
// the sequence's windows into its own shared-memory histogram (two 16-bit counters per word: counts are at most
// maxwin <= 2048) and writes the slot's row segment as fp16 (padding columns zero).  Warp-level synchronisation only.
constexpr int DENSE_COUNT_WARPS = 8;
constexpr int DENSE_MAX_SEG = 12;          // at most 12 key bits, so at most 12 stretches of kept characters
// U8: the counts are written as bytes (the caller knows that no sequence has more than 255 windows): operands of the
// kind::i8 contraction of syrk_tc_welford_kernel, half the bytes per k-mer column; ld is then in bytes.
template <typename GwT, int NW, bool U8 = false>
__global__ void __launch_bounds__(DENSE_COUNT_WARPS * 32)
dense_count_kernel(const GwT* __restrict__ gw0, const uint64_t* __restrict__ gw1, const uint32_t* __restrict__ woff,
                   uint32_t nks, size_t ld, int nb, __half* __restrict__ C, const __grid_constant__ BatchSpec spec) {
    using KeyT = typename std::conditional<sizeof(GwT) == 4, uint32_t, uint64_t>::type;
    constexpr uint32_t KB = sizeof(KeyT) * 8;
    extern __shared__ uint32_t hist_all[];
    __shared__ uint32_t seg_all[DENSE_COUNT_WARPS][DENSE_MAX_SEG];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* __restrict__ seg = seg_all[warp];
    const uint32_t words = nks >> 1;
    uint32_t* __restrict__ hist = hist_all + (size_t)warp * words;
    const uint32_t seq = blockIdx.x;
    const uint32_t w0 = woff[seq], nw = woff[seq + 1] - w0;
    // short sequences (at most 96 windows: the TFBS sets): a lane keeps its three windows' g-mer words in registers over all the
    // slots instead of re-reading them per slot
    constexpr int HOLD = 3;
    const bool held = nw <= 32u * HOLD;
    KeyT lo_r[HOLD], hi_r[HOLD];
#pragma unroll
    for (int t = 0; t < HOLD; ++t) {
        const uint32_t p = lane + 32u * t;
        lo_r[t] = held && p < nw ? (KeyT)gw0[w0 + p] : (KeyT)0;
        hi_r[t] = NW == 2 && held && p < nw ? (KeyT)gw1[w0 + p] : (KeyT)0;
    }
    for (int slot = warp; slot < nb; slot += DENSE_COUNT_WARPS) {
        for (uint32_t i = lane; i < words; i += 32) hist[i] = 0;
        const int nseg = spec.nseg[slot];
        if (lane < nseg) seg[lane] = spec.seg[slot][lane];     // the slot's stretch descriptors: broadcast reads below
        __syncwarp();
        if (held) {
            KeyT key[HOLD];
#pragma unroll
            for (int t = 0; t < HOLD; ++t) key[t] = 0;
            uint32_t dst = 0;
            for (int j = 0; j < nseg; ++j) {            // same stretch packing as pack_hist_kernel, three windows at a time
                const uint32_t e = seg[j];
                const uint32_t src = e & 63u, width = (e >> 7) + 1u;
                const KeyT m = (KeyT)((KeyT) ~(KeyT)0 >> (KB - width)) << dst;
                const uint32_t r = (src - dst) & (KB - 1u);
#pragma unroll
                for (int t = 0; t < HOLD; ++t) key[t] |= rotr_key<KeyT>((NW == 2 && (e & 64u)) ? hi_r[t] : lo_r[t], r) & m;
                dst += width;
            }
#pragma unroll
            for (int t = 0; t < HOLD; ++t) {
                const uint32_t kk = (uint32_t)key[t];
                if (lane + 32u * t < nw) atomicAdd(&hist[kk >> 1], 1u << ((kk & 1u) << 4));
            }
        } else
        for (uint32_t p = lane; p < nw; p += 32) {
            const KeyT lo = (KeyT)gw0[w0 + p];
            const KeyT hi = NW == 2 ? (KeyT)gw1[w0 + p] : (KeyT)0;
            KeyT key = 0;
            uint32_t dst = 0;
            for (int j = 0; j < nseg; ++j) {            // same stretch packing as pack_hist_kernel
                const uint32_t e = seg[j];
                const uint32_t src = e & 63u, width = (e >> 7) + 1u;
                const KeyT m = (KeyT)((KeyT) ~(KeyT)0 >> (KB - width)) << dst;
                const uint32_t r = (src - dst) & (KB - 1u);
                key |= rotr_key<KeyT>((NW == 2 && (e & 64u)) ? hi : lo, r) & m;
                dst += width;
            }
            const uint32_t kk = (uint32_t)key;
            atomicAdd(&hist[kk >> 1], 1u << ((kk & 1u) << 4));
        }
        __syncwarp();
        if constexpr (U8) {
            unsigned short* __restrict__ out = reinterpret_cast<unsigned short*>(reinterpret_cast<uint8_t*>(C) + (size_t)seq * ld + (size_t)slot * nks);
            for (uint32_t i = lane; i < words; i += 32) {
                const uint32_t w = hist[i];
                out[i] = (unsigned short)((w & 0xffu) | ((w >> 16) << 8));
            }
        } else {
        __half2* __restrict__ out = reinterpret_cast<__half2*>(C + (size_t)seq * ld + (size_t)slot * nks);
        for (uint32_t i = lane; i < words; i += 32) {
            const uint32_t w = hist[i];
            out[i] = __halves2half2(__ushort2half_rn((unsigned short)(w & 0xffffu)), __ushort2half_rn((unsigned short)(w >> 16)));
        }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a): mbarrier, TMA, tcgen05
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol error must never hang the GPU -- and must not kill the CUDA context of the host process either
// (a trap would).  A wait that runs out raises a device-side flag and returns; every later wait of the kernel returns at once,
// the kernel finishes with garbage, and the host turns the flag into FSK_ECUDA at the build's next check (sort_was_unstable).
__device__ unsigned int fsk_dev_fault = 0;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (*(volatile unsigned int*)&fsk_dev_fault) return;
    for (uint32_t it = 0;; ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        // back off: a spinning single-lane warp (TMA producer, MMA issuer) otherwise takes issue slots from the epilogue warps
        // of its scheduler (ncu: SYNCS + BRA + YIELD were a third of all instructions of the Welford contraction)
        if (it > 2) __nanosleep(it < 64 ? 40 : 200);
        if (it > (1u << 20)) { atomicExch(&fsk_dev_fault, 1u); return; }     // ~0.2 s of back-off: far beyond any legitimate wait
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {   // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::i8: A = B = unsigned bytes, D = int32; UMMA_K = 32 bytes, the same 32-byte descriptor step as kind::f16
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major operand tile staged by TMA with 128-byte swizzle: rows of 128 bytes,
// 8-row groups 1024 bytes apart (stride byte offset), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                       // leading byte offset: unused for swizzled K-major layouts
    d |= (uint64_t)(1024u >> 4) << 32;            // stride byte offset
    d |= (uint64_t)1 << 46;                       // version
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = fp32, A = B = fp16, both K-major, N at bit 17 (>> 3), M at bit 24 (>> 4)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::i8 instruction descriptor: D = int32 (format 2 at bit 4), A = B = uint8 (format 0 at bits 7 and 10), both K-major
__host__ __device__ constexpr uint32_t umma_idesc_u8(int M, int N) {
    return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// grid = (lower-triangle tiles T (T + 1) / 2 in tile_order -- pairs of tile rows for NA = 2 --, groups).  Group `g` contracts the columns [k_begin + g * k_group,
// + klen) of C and adds into K + g * out_group_stride (integer modes: one group; variance mode: one per slot).
// U8: byte operands and int32 accumulators (kind::i8; dense_count_kernel<..., true>): half the operand bytes, twice the MMA rate.
template <int NA, bool U8 = false>
__global__ void __launch_bounds__(DG_THREADS)
syrk_tc_kernel(const __grid_constant__ CUtensorMap tmap, const uint32_t* __restrict__ tile_order, int64_t nseq, uint32_t k_begin,
               uint32_t k_group, uint32_t klen, unsigned long long* __restrict__ K, size_t out_group_stride,
               const uint32_t* __restrict__ klen_dev) {
    // heavy-run contraction: the number of columns is only known on the device (heavy_fill_kernel's list)
    if (klen_dev) {
        constexpr uint32_t R = U8 ? 2 * DG_BK : DG_BK;
        klen = (min(*klen_dev, klen) + R - 1) / R * R;                 // klen carries the capacity of the list
        if (klen == 0) return;
    }
    constexpr int DG_STAGES = dg_stages(NA);
    constexpr uint32_t DG_STAGE_BYTES = dg_stage_bytes(NA);
    constexpr uint32_t DG_TMEM_COLS = NA * 128;                   // 128 lanes x 128 fp32 columns per output tile
    extern __shared__ uint8_t dg_smem_raw[];
    const uint32_t raw = smem_u32(dg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // swizzle atoms need 1024-byte alignment
    const uint32_t bars = base + DG_STAGES * DG_STAGE_BYTES;      // full[S], empty[S], tmem_full, tmem slot
    const uint32_t full0 = bars, empty0 = bars + 8 * DG_STAGES, tmem_full = bars + 16 * DG_STAGES, tmem_slot = tmem_full + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(dg_smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile (I, J), J <= I, in the host's rasterised order: the CTAs resident at once cover a band of 16 tile rows by as
    // many tile columns, so every operand slice TMA fetches is shared by ~16 CTAs through L2
    // (NA = 2: the entry names the PAIR of tile rows 2 P, 2 P + 1; a tile above the diagonal or past the last row computes
    // zeros or masked cells and is skipped by the epilogue)
    const uint32_t ij = tile_order[blockIdx.x];
    const uint32_t I = NA == 1 ? ij >> 16 : (ij >> 16) * 2u, J = ij & 0xffffu;
    const bool diag = NA == 1 && I == J;
    const uint32_t kx0 = k_begin + blockIdx.y * k_group;
    constexpr uint32_t BK = U8 ? 2 * DG_BK : DG_BK;              // elements per k-block: one 128-byte swizzle row either way
    const uint32_t nkb = klen / BK;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < DG_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {   // one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(DG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            const uint32_t tx = diag ? DG_TILE_BYTES : DG_STAGE_BYTES;   // rows past N are zero-filled and still count
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const uint32_t s = kb % DG_STAGES, ph = (kb / DG_STAGES) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);                           // the MMAs that read this stage have completed
                mbar_arrive_expect_tx(full0 + 8 * s, tx);
                const uint32_t a = base + s * DG_STAGE_BYTES;
                const int x = (int)(kx0 + kb * BK);
                tma_load_2d(a, &tmap, full0 + 8 * s, x, (int)(I * DG_TILE));
                if (NA == 2) tma_load_2d(a + DG_TILE_BYTES, &tmap, full0 + 8 * s, x, (int)((I + 1) * DG_TILE));
                if (!diag) tma_load_2d(a + NA * DG_TILE_BYTES, &tmap, full0 + 8 * s, x, (int)(J * DG_TILE));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = U8 ? umma_idesc_u8(DG_TILE, DG_TILE) : umma_idesc_f16(DG_TILE, DG_TILE);
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const uint32_t s = kb % DG_STAGES, ph = (kb / DG_STAGES) & 1u;
                mbar_wait(full0 + 8 * s, ph);                                 // TMA has landed this stage
                tc_fence_after();
                const uint32_t a = base + s * DG_STAGE_BYTES;
                const uint64_t bdesc = umma_desc_sw128(diag ? a : a + NA * DG_TILE_BYTES);
#pragma unroll
                for (int t = 0; t < NA; ++t) {
                    const uint64_t adesc = umma_desc_sw128(a + t * DG_TILE_BYTES);
#pragma unroll
                    for (uint32_t k = 0; k < DG_BK / 16; ++k) {               // UMMA_K = 16 fp16 or 32 bytes = 32 bytes along the swizzled row
                        if constexpr (U8) tc_mma_i8(tmem_base + t * 128u, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0u);
                        else tc_mma_f16(tmem_base + t * 128u, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0u);
                    }
                }
                tc_commit(empty0 + 8 * s);                                    // frees the stage once these MMAs are done
            }
            tc_commit(tmem_full);                                             // accumulator complete
        }
        __syncwarp();
    } else {
        const uint32_t q = (uint32_t)warp & 3u;                               // the TMEM lane quarter this warp can read
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        // The accumulator comes out of TMEM one row per thread; it goes through a padded shared-memory transpose (the pipeline
        // stages are free by now: every MMA has completed) so that the 32 lanes of a warp touch 32 CONSECUTIVE cells of one
        // row of the packed triangle: coalesced 256-byte accesses instead of 32 rows x 8 bytes.
        float* __restrict__ ts = reinterpret_cast<float*>(dg_smem_raw + (base - raw)) + (warp - 2) * (32 * 33);
        const int64_t j0 = (int64_t)J * DG_TILE;
        unsigned long long* __restrict__ Kg = K + (size_t)blockIdx.y * out_group_stride;
#pragma unroll 1
        for (int t = 0; t < NA; ++t) {
        if ((int64_t)(I + t) * DG_TILE >= nseq || J > I + t) continue;      // nothing of this tile is at or below the diagonal
        const int64_t ibase = (int64_t)(I + t) * DG_TILE + q * 32;
#pragma unroll 1
        for (int c0 = 0; c0 < DG_TILE; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + (uint32_t)(t * 128 + c0), v);
#pragma unroll
            for (int c = 0; c < 32; ++c) ts[lane * 33 + c] = __uint_as_float(v[c]);
            __syncwarp();
            const int64_t j = j0 + c0 + lane;
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += 8) {
                {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int64_t i = ibase + r0 + u;
                        const uint32_t cnt = U8 ? __float_as_uint(ts[(r0 + u) * 33 + lane]) : __float2uint_rn(ts[(r0 + u) * 33 + lane]);   // (U8: the int32's bits)
                        if (i < nseq && j <= i && cnt)
                            atomicAdd(&Kg[(size_t)(i * (i + 1) / 2 + j)], (unsigned long long)cnt);   // RED.ADD.64: the tile is owned by this CTA
                    }
                }
            }
            __syncwarp();
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(DG_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// Variance mode at batch speed (fastsk_kernel.cpp:188-262): grid = (lower-triangle tiles, virtual streams of the round).  A
// CTA owns one 128 x 128 tile of ONE stream's running mean and walks the stream's slots -- consecutive iterations, speculated
// past a possible stop -- IN ORDER: per slot, 4 x (nks / 64) MMAs build the tile of this iteration's partial kernel in TMEM
// and the epilogue applies the Welford step to the tile of the mean (which the same CTA touched a few microseconds ago: L2
// hits) and adds up delta * delta2.  Two TMEM accumulators: the MMAs of slot d + 1 run under the epilogue of slot d.  One
// launch does what used to be one launch + one host round trip per iteration.
constexpr int DW_STAGES = 2, DW_STAGES_REGS = 4;
constexpr uint32_t DW_STAGE_BYTES = 2 * DG_TILE_BYTES;
constexpr int DW_EPI_WARPS = 8;                                   // two epilogue warps per TMEM lane quarter, 64 columns each
constexpr int DW_THREADS = 64 + 32 * DW_EPI_WARPS;                // streamed form: warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int DW_THREADS_REGS = 128 + 32 * DW_EPI_WARPS;         // register form: one producer warpgroup (TMA, MMA, TMEM allocation, idle) + two epilogue warpgroups
constexpr uint32_t DW_TS_BYTES = DW_EPI_WARPS * 32 * 33 * 4;      // one padded 32 x 32 fp32 transpose buffer per epilogue warp
__host__ __device__ constexpr size_t dw_smem(bool regs) { return (size_t)(regs ? DW_STAGES_REGS : DW_STAGES) * DW_STAGE_BYTES + DW_TS_BYTES + 1024 + 256; }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// REGS: the 64 cells a thread owns (its row of the TMEM lane quarter x the 64 columns of its half) stay in REGISTERS for all
// the slots of the launch: the means are read once before the first slot and written once after the last, and a slot costs
// one TMEM read and eight fp64 operations per cell -- no shared-memory transpose, no L2 round trip per iteration (the
// streaming form moved 2 x 8 bytes per cell and iteration: 6.8 GB for 50 iterations of EP300, all of the kernel's time).
// The means are then stored STRIP-MAJOR inside a tile: cell (row 32 q + lane, column 64 half + c) at
// ((2 q + half) * 64 + c) * 32 + lane, so that the one load and the one store are coalesced (welford_untile_kernel knows).
// One CTA per SM; 128 of an epilogue thread's registers are means, so the epilogue warpgroups take 232 registers and the
// producer warpgroup gives its own back (setmaxnreg): with the 168 registers a 12-warp CTA starts with, ptxas ran the eight
// dependent fp64 operations of one cell after the other (r2s35: fp64 pipe 33 %, stall_wait 3.5 per issue); with room for
// eight cells in flight the two epilogue warps of a scheduler keep the fp64 pipe busy.  !REGS is the streaming form.
// U8 (register form only): byte operands and int32 accumulators (kind::i8) when no sequence has more than 255 windows --
// the operand tiles, which at 256 k-mer columns per slot cost more shared-memory fill time (128 KB per tile and slot at
// ~43 B/clk per SM) than the Welford arithmetic, halve; and the int32 count becomes a double with one fp64 addition
// (2^52 + v built from the bits, minus 2^52) instead of a quarter-rate conversion.
template <bool REGS, bool U8 = false>
__global__ void __launch_bounds__(REGS ? DW_THREADS_REGS : DW_THREADS, REGS ? 1 : 2)
syrk_tc_welford_kernel(const __grid_constant__ CUtensorMap tmap, const uint32_t* __restrict__ tile_order, int64_t nseq, uint32_t nks,
                       const WelfordSpec* __restrict__ wf) {
    extern __shared__ uint8_t dw_smem_raw[];
    constexpr int S = REGS ? DW_STAGES_REGS : DW_STAGES;          // operand stages (one CTA per SM has room for more)
    const uint32_t raw = smem_u32(dw_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t ts_base = base + S * DW_STAGE_BYTES;
    const uint32_t bars = ts_base + DW_TS_BYTES;                  // full[S], empty[S], tmem_full[2], tmem_empty[2], tmem slot
    const uint32_t full0 = bars, empty0 = bars + 8 * S, tfull0 = bars + 16 * S, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(dw_smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ij = tile_order[blockIdx.x];
    const uint32_t I = ij >> 16, J = ij & 0xffffu;
    const bool diag = I == J;
    const int g = blockIdx.y;
    const uint32_t depth = wf->depth[g], slot0 = wf->slot0[g];
    static_assert(REGS || !U8, "byte operands: register form only");
    constexpr uint32_t BK = U8 ? 2 * DG_BK : DG_BK;              // elements per k-block: one 128-byte swizzle row either way
    const uint32_t nkb = nks / BK;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, DW_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    constexpr int EPI0 = REGS ? 4 : 2;                            // first epilogue warp
    if (warp < EPI0) {                                            // the producer warps: all of their code inside this branch,
        if constexpr (REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");   // so that ptxas budgets each side
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            const uint32_t tx = diag ? DG_TILE_BYTES : DW_STAGE_BYTES;
            const uint32_t total = depth * nkb;
            for (uint32_t t = 0; t < total; ++t) {
                const uint32_t s = t % S, ph = (t / S) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                mbar_arrive_expect_tx(full0 + 8 * s, tx);
                const uint32_t a = base + s * DW_STAGE_BYTES;
                const int x = (int)((slot0 + t / nkb) * nks + (t % nkb) * BK);
                tma_load_2d(a, &tmap, full0 + 8 * s, x, (int)(I * DG_TILE));
                if (!diag) tma_load_2d(a + DG_TILE_BYTES, &tmap, full0 + 8 * s, x, (int)(J * DG_TILE));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = U8 ? umma_idesc_u8(DG_TILE, DG_TILE) : umma_idesc_f16(DG_TILE, DG_TILE);
            uint32_t t = 0;
            for (uint32_t d = 0; d < depth; ++d) {
                const uint32_t buf = d & 1u;
                mbar_wait(tempty0 + 8 * buf, ((d >> 1) & 1u) ^ 1u);          // the epilogue has drained this accumulator
                tc_fence_after();
                for (uint32_t kb = 0; kb < nkb; ++kb, ++t) {
                    const uint32_t s = t % S, ph = (t / S) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t a = base + s * DW_STAGE_BYTES;
                    const uint64_t adesc = umma_desc_sw128(a), bdesc = umma_desc_sw128(diag ? a : a + DG_TILE_BYTES);
#pragma unroll
                    for (uint32_t k = 0; k < DG_BK / 16; ++k) {          // four 32-byte steps along the swizzled row
                        if constexpr (U8) tc_mma_i8(tmem_base + buf * 128u, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0u);
                        else tc_mma_f16(tmem_base + buf * 128u, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0u);
                    }
                    tc_commit(empty0 + 8 * s);
                }
                tc_commit(tfull0 + 8 * buf);
            }
        }
        __syncwarp();
    }
    } else {
        if constexpr (REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
        // The accumulator comes out of TMEM one row per thread; it goes through a padded shared-memory transpose so that the 32
        // lanes of a warp touch 32 CONSECUTIVE cells of one row of the packed triangle (coalesced 256-byte accesses; streaming a
        // row per thread instead measured 40 % slower: 32 sectors per load instruction).  The 32 cells a lane owns in a
        // 32-column chunk are fetched together and BEFORE the wait for the accumulator / the TMEM read of the chunk, so the L2
        // round trip of the mean (the same CTA wrote it one slot ago) runs under the MMAs and the transpose.
        // Eight epilogue warps: a warp reads the TMEM lane quarter (warp % 4) -- the hardware's rule -- and the two warps of a
        // quarter take 64 columns each.  The running means of this path live TILE-MAJOR (tile (I, J) = 128 x 128 consecutive
        // cells, row-major): the cell of row r, column c of the tile is at a compile-time offset from the warp's base pointer,
        // so the inner loop has no address arithmetic at all (the packed triangle cost ~30 of 53 instructions per cell: ncu,
        // profiles/r02_ncu_welford_*).  welford_untile_kernel adds the tiles into the packed triangle once, at the end.
        const uint32_t q = (uint32_t)warp & 3u, half = (uint32_t)(warp - EPI0) >> 2;
        float* __restrict__ ts = reinterpret_cast<float*>(dw_smem_raw + (ts_base - raw)) + (warp - EPI0) * (32 * 33);
        const int64_t j0 = (int64_t)J * DG_TILE + half * 64;
        const int64_t ibase = (int64_t)I * DG_TILE + q * 32;
        const int64_t n_train = wf->n_train;
        if constexpr (REGS) {
            const size_t tcell = ((size_t)I * (I + 1) / 2 + J) * (size_t)(DG_TILE * DG_TILE) + (size_t)((q * 2 + half) * 64) * 32 + lane;
            const double* __restrict__ kin = wf->khat_in[g] + tcell;
            double* __restrict__ kout = wf->khat_out[g] + tcell;
            const int64_t i = ibase + lane;                               // this thread's row of the kernel matrix
            // cells of the strip that count for the variance: training rows, columns up to the diagonal.  Warp-uniform cases:
            // all 32 x 64 of them, none of them (test rows -- most tiles when the test set is large -- and strips above the
            // diagonal: the mean is still updated, the variance terms are not even computed), or some (select, no branches)
            const bool whole = ibase + 31 < n_train && j0 + 63 <= ibase;
            const bool none = ibase >= n_train || j0 > ibase + 31;
            const int ncount = i < n_train ? (int)max((int64_t)0, min((int64_t)64, i - j0 + 1)) : 0;   // columns c < ncount count
            double m[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) m[c] = kin[c * 32];
            for (uint32_t d = 0; d < depth; ++d) {
                const uint32_t buf = d & 1u;
                const double diter = (double)(wf->iter0[g] + (int32_t)d);
                const double riter = __ddiv_rn(1.0, diter);
                double acc8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc8[u] = 0.0;
                mbar_wait(tfull0 + 8 * buf, (d >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(tmem_base + ((q * 32u) << 16) + (uint32_t)(buf * 128 + half * 64 + c0), v);
                    if (c0 == 48) {                                           // last read of this accumulator: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
                    }
                    // eight cells at a time, stage by stage: eight independent chains for the fp64 pipe
#pragma unroll
                    for (int h = 0; h < 16; h += 8) {
                        double ks[8], delta[8], t[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            ks[u] = U8 ? __dsub_rn(__hiloint2double(0x43300000, (int)v[h + u]), 4503599627370496.0)   // (2^52 + v) - 2^52
                                       : (double)__uint_as_float(v[h + u]);                                          // integers below 2^24: exact
#pragma unroll
                        for (int u = 0; u < 8; ++u) delta[u] = __dsub_rn(ks[u], m[c0 + h + u]);
#pragma unroll
                        for (int u = 0; u < 8; ++u) t[u] = __dmul_rn(delta[u], riter);            // div_by_iter, staged
#pragma unroll
                        for (int u = 0; u < 8; ++u) t[u] = __fma_rn(__fma_rn(-t[u], diter, delta[u]), riter, t[u]);
#pragma unroll
                        for (int u = 0; u < 8; ++u) m[c0 + h + u] = __dadd_rn(m[c0 + h + u], t[u]);
                        if (none) continue;
#pragma unroll
                        for (int u = 0; u < 8; ++u) t[u] = __dmul_rn(delta[u], __dsub_rn(ks[u], m[c0 + h + u]));
                        if (whole) {
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc8[u] = __dadd_rn(acc8[u], t[u]);
                        } else {
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc8[u] = __dadd_rn(acc8[u], c0 + h + u < ncount ? t[u] : 0.0);
                        }
                    }
                }
                double acc = __dadd_rn(__dadd_rn(__dadd_rn(acc8[0], acc8[1]), __dadd_rn(acc8[2], acc8[3])),
                                       __dadd_rn(__dadd_rn(acc8[4], acc8[5]), __dadd_rn(acc8[6], acc8[7])));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
                if (lane == 0) wf->sums[(size_t)(slot0 + d) * wf->sums_stride + (size_t)blockIdx.x * 8 + half * 4 + q] = acc;
            }
#pragma unroll
            for (int c = 0; c < 64; ++c) kout[c * 32] = m[c];
        } else {
        // every cell of the warp's 32 x 64 strip is inside the triangle and the matrix (all but the diagonal and last tiles)
        const bool inside = ibase + 31 < nseq && j0 + 63 <= ibase;
        const bool all_train = ibase + 31 < n_train;
        const size_t tcell = ((size_t)I * (I + 1) / 2 + J) * (size_t)(DG_TILE * DG_TILE) + (size_t)(q * 32) * DG_TILE + half * 64 + lane;
        const double* __restrict__ kin = wf->khat_in[g] + tcell;
        double* __restrict__ kout = wf->khat_out[g] + tcell;
        for (uint32_t d = 0; d < depth; ++d) {
            const uint32_t buf = d & 1u;
            const double* __restrict__ ksrc = d == 0 ? kin : kout;
            const double diter = (double)(wf->iter0[g] + (int32_t)d);
            const double riter = __ddiv_rn(1.0, diter);
            double acc = 0.0;
            // rows r0 .. r0 + 15 of the column this lane owns in chunk c0 (cells outside the triangle hold harmless garbage)
            auto fetch = [&](int c0, int r0, double (&k0)[16]) {
#pragma unroll
                for (int u = 0; u < 16; ++u) k0[u] = ksrc[(r0 + u) * DG_TILE + c0];
            };
            double k0[16];
            fetch(0, 0, k0);
            mbar_wait(tfull0 + 8 * buf, (d >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + (uint32_t)(buf * 128 + half * 64 + c0), v);
                    if (c0 == 32) {                                           // last read of this accumulator: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) ts[lane * 33 + c] = __uint_as_float(v[c]);
                }
                __syncwarp();
                const int64_t j = j0 + c0 + lane;
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += 16) {
                    if (inside && all_train) {
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const double ks = (double)ts[(r0 + u) * 33 + lane];   // an integer below 2^24: exact in fp32 and in fp64
                            const double delta = __dsub_rn(ks, k0[u]);
                            const double nh = __dadd_rn(k0[u], div_by_iter(delta, diter, riter));
                            kout[(r0 + u) * DG_TILE + c0] = nh;
                            acc = __dadd_rn(acc, __dmul_rn(delta, __dsub_rn(ks, nh)));
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const int64_t i = ibase + r0 + u;
                            const double ks = (double)ts[(r0 + u) * 33 + lane];
                            const double delta = __dsub_rn(ks, k0[u]);
                            const double nh = __dadd_rn(k0[u], div_by_iter(delta, diter, riter));
                            kout[(r0 + u) * DG_TILE + c0] = nh;
                            if (i < n_train && j <= i) acc = __dadd_rn(acc, __dmul_rn(delta, __dsub_rn(ks, nh)));   // (n_train <= nseq)
                        }
                    }
                    // the next 16 rows (or the next chunk's first 16) are in flight under this half's stores and the next TMEM read
                    if (r0 == 0) fetch(c0, 16, k0);
                    else if (c0 == 0) fetch(32, 0, k0);
                }
                __syncwarp();
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            if (lane == 0) wf->sums[(size_t)(slot0 + d) * wf->sums_stride + (size_t)blockIdx.x * 8 + half * 4 + q] = acc;
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// dst (packed lower triangle) += src (tile-major running mean of one stream): Ksfinal += K_hat (fastsk_kernel.cpp:296-313).
// strips = 0: tiles stored row-major; 1: strip-major (syrk_tc_welford_kernel<true>), transposed through shared memory so
// that both the reads and the updates of the triangle's rows are coalesced.
__global__ void __launch_bounds__(256)
welford_untile_kernel(double* __restrict__ dst, const double* __restrict__ src, const uint32_t* __restrict__ tile_order, int64_t nseq, int strips) {
    __shared__ double sm[64 * 33];
    const uint32_t ij = tile_order[blockIdx.x];
    const int64_t I = ij >> 16, J = ij & 0xffffu;
    const double* __restrict__ t = src + ((size_t)I * (I + 1) / 2 + J) * (size_t)(DG_TILE * DG_TILE);
    if (!strips) {
        for (int e = threadIdx.x; e < DG_TILE * DG_TILE; e += 256) {
            const int64_t i = I * DG_TILE + (e >> 7), j = J * DG_TILE + (e & 127);
            if (i < nseq && j <= i) {
                const size_t p = (size_t)(i * (i + 1) / 2 + j);
                dst[p] = __dadd_rn(dst[p], t[e]);
            }
        }
        return;
    }
    for (int strip = 0; strip < 8; ++strip) {
        const int q = strip >> 1, half = strip & 1;
        __syncthreads();
        for (int e = threadIdx.x; e < 2048; e += 256) sm[(e >> 5) * 33 + (e & 31)] = t[strip * 2048 + e];   // [column][row]
        __syncthreads();
        for (int e = threadIdx.x; e < 2048; e += 256) {
            const int r = e >> 6, c = e & 63;
            const int64_t i = I * DG_TILE + q * 32 + r, j = J * DG_TILE + half * 64 + c;
            if (i < nseq && j <= i) {
                const size_t p = (size_t)(i * (i + 1) / 2 + j);
                dst[p] = __dadd_rn(dst[p], sm[c * 33 + r]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Heavy runs of the sparse regime ("dense contraction when the heavy-run C^T C update is dense").  A run of
// d records costs d^2 / 2 shared-memory updates on the row path but only one column of a tensor-core contraction:
// segment_kernel files the tasks of runs longer than heavy_tau empty and lists the runs; here each listed run becomes one
// fp16 column of H[sequence][column] (count of the sequence's records in the run), and syrk_tc_kernel adds H H^T to K.
// Break-even on B200: d > ~0.05 N (1.3e12 updates/s against 6e14 MAC/s).

// zero the first roundup(*count) columns of H (whole 128-byte k-blocks of the contraction).  OutT: __half, or uint8_t when
// no sequence has more than 255 windows (byte operands, syrk_tc_kernel<NA, true>); ld in elements
template <typename OutT>
__global__ void heavy_zero_kernel(OutT* __restrict__ H, size_t ld, int64_t nseq, const uint32_t* __restrict__ count) {
    constexpr uint32_t R = 128 / sizeof(OutT), PER16 = 16 / sizeof(OutT);
    const uint32_t cols = (min(*count, (uint32_t)ld) + R - 1) / R * R;               // ld = capacity of the list
    if (cols == 0) return;
    const uint32_t vec = cols / PER16;                               // 16-byte pieces per row
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nseq * vec; i += (int64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(H + (size_t)(i / vec) * ld)[i % vec] = z;
}

// one CTA per listed run (grid-stride): find the run's end in the sorted records, then every group of equal sequence ids
// writes its size into the run's column.  `total` accumulates the number of heavy runs since upload (statistics).
template <typename RecT, typename OutT>
__global__ void __launch_bounds__(256)
heavy_fill_kernel(const RecT* __restrict__ rec, uint32_t n, int idbits, const uint2* __restrict__ list,
                  const uint32_t* __restrict__ count, OutT* __restrict__ H, size_t ld, unsigned long long* __restrict__ total) {
    __shared__ uint32_t s_end;
    const uint32_t nh = min(*count, (uint32_t)ld);                    // runs beyond the capacity stayed on the sparse path
    if (blockIdx.x == 0 && threadIdx.x == 0 && total) atomicAdd(total, (unsigned long long)nh);
    const RecT idmask = ((RecT)1 << idbits) - 1;
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        const uint2 e = list[h];
        const RecT* __restrict__ R = rec + (size_t)e.x * n;
        const uint32_t rs = e.y;
        if (threadIdx.x == 0) {                                      // first record after the run: the records are sorted by key
            const RecT key = R[rs] >> idbits;
            uint32_t lo = rs + 1, hi = n;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if ((R[mid] >> idbits) == key) lo = mid + 1;
                else hi = mid;
            }
            s_end = lo;
        }
        __syncthreads();
        const uint32_t d = s_end - rs;
        for (uint32_t t = threadIdx.x; t < d; t += blockDim.x) {
            const uint32_t id = (uint32_t)(R[rs + t] & idmask);
            if (t == 0 || (uint32_t)(R[rs + t - 1] & idmask) != id) {
                uint32_t c = 1;
                while (t + c < d && (uint32_t)(R[rs + t + c] & idmask) == id) ++c;
                if constexpr (sizeof(OutT) == 1) H[(size_t)id * ld + h] = (OutT)c;
                else H[(size_t)id * ld + h] = __uint2half_rn(c);
            }
        }
        __syncthreads();
    }
}

}  // namespace fsk
