// gram_tc.cu -- the stacked Gram contraction of the solver on the 5th-generation tensor cores
// (fp32 contexts, kp = 32, Kc a multiple of 128):
//
//     Out[Kc x kp] = sum_rows A[row, 0:Kc]^T B[row, 0:kp],      B = columns [bcol, bcol + kp) of A
//
// (mm(a,b,c,k,l) of ffm.cpp:41-45 for all cross pairs at once: Q~^T Q1 of gd_cross, ffm.cpp:663-670,
// and Q1^T Q1 of cg, ffm.cpp:767-771).  At Kc = 256, kp = 32 the product needs 8192 FMAs per 1 KB row:
// 43 TFMA/s at HBM speed, more than the SIMT pipes deliver -- the previous SIMT kernel ran the
// 4.1 GB pass of the KDD12 shape at 1.8 TB/s.  On tcgen05 the MMA work is a quarter of the HBM time.
//
// The reduction runs over ROWS, so both operands are MN-major (the M / N index is the contiguous
// one in memory).  A block of 32 rows is brought in by TMA as Kc/32 boxes of [32 columns x 32 rows]
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): box c holds rows r at r*128 bytes, which is the canonical
// MN-major layout of a 32-bit UMMA operand (layout type SWIZZLE_128B_BASE32B: atoms of 4 K-rows x
// 128 bytes; stride between 32-column atoms LBO = 4 KB, between 4-row groups SBO = 512 bytes).
// B needs no load of its own: it is box bcol/32 of the same block.
//
// Precision: 3xTF32 (hi*hi + hi*lo + lo*hi, hi = cvt.rna.tf32, lo = x - hi exact): the split is done
// in shared memory by four warps between the TMA and the MMA (element-wise on the swizzled image, so
// it is layout-agnostic).  Accumulators live in TMEM (fp32) for 512 rows at a time, are then added
// to fp64 registers by the epilogue warps (two accumulator sets ping-pong), and CTAs combine with
// fp64 atomics: Gram entries agree with an fp64 computation to ~1e-6.
//
// Per CTA (persistent, one per SM, 320 threads): warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-5 hi/lo split, warps 6-9 epilogue (TMEM lane quadrant = warp % 4).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ocffm {

namespace {

constexpr int BR = 32;                  // rows per block (4 MMA K-steps of 8)
constexpr int BOXB = BR * 128;          // bytes of one [32 columns x 32 rows] box
constexpr int FLUSH = 16;               // row blocks per TMEM accumulation (512 rows)
constexpr int kThreadsG = 320;
constexpr int KPG = 32;                 // padded latent dimension served here

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// MN-major TF32 operand.  32-bit MN-major operands have ONE legal shared-memory layout on sm_100: the
// 128-byte swizzle with a 32-byte base (UMMA layout type 1, Swizzle<2,5,2>; TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 32 elements (128 bytes) along M/N x 4 K-rows, the
// 32-byte chunk index of a row XORed with (row mod 4).  LBO = stride between 32-element atoms along
// M/N, SBO = stride between groups of 4 K-rows (512 bytes: a K = 8 instruction spans two groups);
// descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFFu);
    d |= uint64_t((uint32_t(BOXB) >> 4) & 0x3FFFu) << 16;   // leading byte offset
    d |= uint64_t(512u >> 4) << 32;                          // stride byte offset
    d |= uint64_t(1) << 46;
    d |= uint64_t(1) << 61;                                  // SWIZZLE_128B_BASE32B
    return d;
}
// D = F32, A = B = TF32, both MN-major (bits 15, 16), M = 128, N = 32
constexpr uint32_t kIdescG = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             (uint32_t(KPG >> 3) << 17) | (uint32_t(128 >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdescG), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory plan: STAGES x (raw -> hi image, lo image), each Kc/32 boxes
template <int NH>
struct Plan {
    static constexpr int kBoxes = NH * 4;                 // Kc / 32
    static constexpr int kStageBytes = kBoxes * BOXB;     // 16 KB per 128 columns
    static constexpr int kStages = NH == 1 ? 6 : 3;
    static constexpr int kTmemCols = NH * KPG * 2 <= 32 ? 32 : (NH * KPG * 2 <= 64 ? 64 : (NH * KPG * 2 <= 128 ? 128 : 256));
};
struct Bars {
    uint64_t full[6], split[6], empty[6], tfull[2], tempty[2];
    uint32_t tmem_base;
};

template <int NH>
__global__ void __launch_bounds__(kThreadsG, 1)
k_gram_tc(const __grid_constant__ CUtensorMap map_a, uint32_t nblocks, uint32_t bbox, double *__restrict__ Out64) {
    pdl_enter();
    using P = Plan<NH>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    unsigned char *hi_base = smem_raw + pad;                                   // [kStages][kStageBytes]
    unsigned char *lo_base = hi_base + size_t(P::kStages) * P::kStageBytes;    // [kStages][kStageBytes]
    Bars &bar = *reinterpret_cast<Bars *>(lo_base + size_t(P::kStages) * P::kStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's row blocks: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t my_blocks = blockIdx.x < nblocks ? (nblocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P::kStages; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.split[s], 4); mbar_init(&bar.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bar.tfull[s], 1); mbar_init(&bar.tempty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)),
                     "r"(uint32_t(P::kTmemCols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bar.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (uint32_t i = 0; i < my_blocks; ++i) {
                const uint32_t s = i % P::kStages, ph = (i / P::kStages) & 1u;
                const int r0 = int((blockIdx.x + i * gridDim.x) * BR);
                mbar_wait(&bar.empty[s], ph ^ 1u);
                mbar_expect_tx(&bar.full[s], uint32_t(P::kStageBytes));
                for (int c = 0; c < P::kBoxes; ++c)
                    tma_load_2d(hi_base + size_t(s) * P::kStageBytes + size_t(c) * BOXB, &map_a, &bar.full[s], c * 32, r0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (uint32_t i = 0; i < my_blocks; ++i) {
                const uint32_t s = i % P::kStages, ph = (i / P::kStages) & 1u;
                const uint32_t round = i / FLUSH, acc = round & 1u, first = (i % FLUSH) == 0;
                if (first) {
                    mbar_wait(&bar.tempty[acc], ((round >> 1) & 1u) ^ 1u);   // epilogue drained this set
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(&bar.split[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t hi = smem_u32(hi_base + size_t(s) * P::kStageBytes);
                const uint32_t lo = smem_u32(lo_base + size_t(s) * P::kStageBytes);
#pragma unroll
                for (int ks = 0; ks < BR / 8; ++ks) {
                    const uint32_t koff = uint32_t(ks) * 1024u;
                    const uint64_t bh = make_desc_mn(hi + bbox * BOXB + koff), bl = make_desc_mn(lo + bbox * BOXB + koff);
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        const uint32_t d = tmem_base + (acc * NH + h) * KPG;
                        const uint64_t ah = make_desc_mn(hi + uint32_t(h) * 4u * BOXB + koff);
                        const uint64_t al = make_desc_mn(lo + uint32_t(h) * 4u * BOXB + koff);
                        umma_tf32(d, ah, bh, (first && ks == 0) ? 0u : 1u);
                        umma_tf32(d, ah, bl, 1u);
                        umma_tf32(d, al, bh, 1u);
                    }
                }
                umma_commit(&bar.empty[s]);                       // stage reusable once these MMAs retire
                if ((i % FLUSH) == FLUSH - 1 || i + 1 == my_blocks) umma_commit(&bar.tfull[acc]);
            }
        }
    } else if (warp < 6) {
        // ===== hi / lo split, element-wise on the swizzled image =====
        const int t = threadIdx.x - 64;   // 0..127
        for (uint32_t i = 0; i < my_blocks; ++i) {
            const uint32_t s = i % P::kStages, ph = (i / P::kStages) & 1u;
            mbar_wait(&bar.full[s], ph);
            float4 *hp = reinterpret_cast<float4 *>(hi_base + size_t(s) * P::kStageBytes);
            float4 *lp = reinterpret_cast<float4 *>(lo_base + size_t(s) * P::kStageBytes);
#pragma unroll 4
            for (int e = t; e < P::kStageBytes / 16; e += 128) {
                const float4 x = hp[e];
                float4 h, l;
                uint32_t u;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.x)); h.x = __uint_as_float(u); l.x = x.x - h.x;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.y)); h.y = __uint_as_float(u); l.y = x.y - h.y;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.z)); h.z = __uint_as_float(u); l.z = x.z - h.z;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.w)); h.w = __uint_as_float(u); l.w = x.w - h.w;
                hp[e] = h;
                lp[e] = l;
            }
            // generic-proxy writes must be visible to the tensor core's (async proxy) reads
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.split[s]);
        }
    } else {
        // ===== epilogue: TMEM -> fp64 registers every FLUSH blocks, fp64 atomics at the end =====
        const int q = warp & 3;
        double accd[NH][KPG];
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int c = 0; c < KPG; ++c) accd[h][c] = 0.0;
        const uint32_t rounds = (my_blocks + FLUSH - 1) / FLUSH;
        for (uint32_t rd = 0; rd < rounds; ++rd) {
            const uint32_t acc = rd & 1u;
            mbar_wait(&bar.tfull[acc], (rd >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                uint32_t r[32];
                tmem_ld32(tmem_base + (acc * NH + h) * KPG + (uint32_t(q * 32) << 16), r);
#pragma unroll
                for (int c = 0; c < KPG; ++c) accd[h][c] += double(__uint_as_float(r[c]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.tempty[acc]);
        }
        if (my_blocks) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                double *o = Out64 + size_t(h * 128 + q * 32 + lane) * KPG;
#pragma unroll
                for (int c = 0; c < KPG; ++c) atomicAdd(o + c, accd[h][c]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(P::kTmemCols))
                     : "memory");
    }
}

// colsum64[c] += sum_rows B[row, c]; wsum64[c] += sum_rows w[row] B[row, c]   (mv of ffm.cpp:660-661)
__global__ void __launch_bounds__(256)
k_colsum_w(const float *__restrict__ B, uint32_t ldb, uint32_t row0, uint32_t row1, const float *__restrict__ wvec,
           double *__restrict__ colsum64, double *__restrict__ wsum64) {
    pdl_enter();
    __shared__ double sh[2][8][KPG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double cs = 0, ws = 0;
    const uint64_t stride = uint64_t(gridDim.x) * 8;
    for (uint64_t r = uint64_t(row0) + uint64_t(blockIdx.x) * 8 + warp; r < row1; r += stride) {
        const float v = __ldg(B + r * ldb + lane);
        cs += double(v);
        if (wvec) ws += double(__ldg(wvec + r)) * double(v);
    }
    sh[0][warp][lane] = cs;
    sh[1][warp][lane] = ws;
    __syncthreads();
    if (warp == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w][lane]; b += sh[1][w][lane]; }
        atomicAdd(colsum64 + lane, a);
        if (wsum64) atomicAdd(wsum64 + lane, b);
    }
}


// ---------------------------------------------------------------------------------------------
// Row GEMM on tcgen05: C[M x 32] = A[M x Ka] * B[Ka x 32]   (T = P~ Gstack of gd_cross,
// ffm.cpp:663-670: Ka = Kc = 128 / 256, A streamed once from HBM, B the 16-32 KB Gram stack).
// A tiles are K-major ([128 rows x 32 floats] boxes, SWIZZLE_128B), B is used as it lies in memory
// ([Ka x 32] row-major = MN-major operand, one 128-byte atom per K row, 32-byte-base swizzle), both split into hi / lo
// in shared memory (3xTF32).  Persistent CTAs, 4-stage ring of K-blocks, two TMEM accumulators.
// ---------------------------------------------------------------------------------------------
constexpr int RT = 128;                      // rows per tile
constexpr int RSTAGES = 4;
constexpr int RSTAGE_BYTES = RT * 128;       // one K-block of a row tile: [128 rows x 32 floats]
struct RBars {
    uint64_t full[RSTAGES], split[RSTAGES], empty[RSTAGES], tfull[2], tempty[2], bfull, bready;
    uint32_t tmem_base;
};
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {   // K-major, SWIZZLE_128B, 8-row groups of 1 KB
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFFu);
    d |= uint64_t(1024u >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
// D = F32, A = TF32 K-major, B = TF32 MN-major (bit 16), M = 128, N = 32
constexpr uint32_t kIdescR = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | (uint32_t(KPG >> 3) << 17) |
                             (uint32_t(128 >> 4) << 24);
__device__ __forceinline__ void umma_tf32_r(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdescR), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void split_image(float4 *hp, float4 *lp, int n16, int t) {
#pragma unroll 4
    for (int e = t; e < n16; e += 128) {
        const float4 x = hp[e];
        float4 h, l;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.x)); h.x = __uint_as_float(u); l.x = x.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.y)); h.y = __uint_as_float(u); l.y = x.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.z)); h.z = __uint_as_float(u); l.z = x.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.w)); h.w = __uint_as_float(u); l.w = x.w - h.w;
        hp[e] = h;
        lp[e] = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int NKB>   // K-blocks of 32: Ka = 32 * NKB
__global__ void __launch_bounds__(kThreadsG, 1)
k_rowgemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, uint32_t ntiles,
             uint64_t M, float *__restrict__ C) {
    pdl_enter();
    constexpr int BBYTES = NKB * 32 * 128;   // B image: Ka rows of 128 bytes
    extern __shared__ unsigned char smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    unsigned char *a_hi = smem_raw + pad;                       // [RSTAGES][RSTAGE_BYTES]
    unsigned char *a_lo = a_hi + RSTAGES * RSTAGE_BYTES;
    unsigned char *b_hi = a_lo + RSTAGES * RSTAGE_BYTES;        // [BBYTES]
    unsigned char *b_lo = b_hi + BBYTES;
    RBars &bar = *reinterpret_cast<RBars *>(b_lo + BBYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RSTAGES; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.split[s], 4); mbar_init(&bar.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bar.tfull[s], 1); mbar_init(&bar.tempty[s], 4); }
        mbar_init(&bar.bfull, 1);
        mbar_init(&bar.bready, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)),
                     "r"(uint32_t(64))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bar.tmem_base;

    if (warp == 0) {
        if (lane == 0 && my_tiles) {
            mbar_expect_tx(&bar.bfull, uint32_t(BBYTES));
            tma_load_2d(b_hi, &map_b, &bar.bfull, 0, 0);
            uint32_t it = 0;
            for (uint32_t t = 0; t < my_tiles; ++t) {
                const int r0 = int((blockIdx.x + t * gridDim.x) * RT);
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const uint32_t s = it % RSTAGES, ph = (it / RSTAGES) & 1u;
                    mbar_wait(&bar.empty[s], ph ^ 1u);
                    mbar_expect_tx(&bar.full[s], uint32_t(RSTAGE_BYTES));
                    tma_load_2d(a_hi + size_t(s) * RSTAGE_BYTES, &map_a, &bar.full[s], kb * 32, r0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && my_tiles) {
            mbar_wait(&bar.bready, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bh0 = smem_u32(b_hi), bl0 = smem_u32(b_lo);
            uint32_t it = 0;
            for (uint32_t t = 0; t < my_tiles; ++t) {
                const uint32_t acc = t & 1u;
                mbar_wait(&bar.tempty[acc], ((t >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * KPG;
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const uint32_t s = it % RSTAGES, ph = (it / RSTAGES) & 1u;
                    mbar_wait(&bar.split[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t ah = make_desc_k(smem_u32(a_hi + size_t(s) * RSTAGE_BYTES));
                    const uint64_t al = make_desc_k(smem_u32(a_lo + size_t(s) * RSTAGE_BYTES));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = uint64_t((ks * 8 * 4) >> 4);                 // 32 bytes per K-step inside the row
                        const uint32_t boff = uint32_t(kb * 4 + ks) * 1024u;              // 8 K-rows of B per step
                        const uint64_t bh = make_desc_mn(bh0 + boff), bl = make_desc_mn(bl0 + boff);
                        umma_tf32_r(d, ah + adv, bh, (kb | ks) != 0 ? 1u : 0u);
                        umma_tf32_r(d, ah + adv, bl, 1u);
                        umma_tf32_r(d, al + adv, bh, 1u);
                    }
                    umma_commit(&bar.empty[s]);
                }
                umma_commit(&bar.tfull[acc]);
            }
        }
    } else if (warp < 6) {
        const int t128 = threadIdx.x - 64;
        if (my_tiles) {
            mbar_wait(&bar.bfull, 0);
            split_image(reinterpret_cast<float4 *>(b_hi), reinterpret_cast<float4 *>(b_lo), BBYTES / 16, t128);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.bready);
        }
        const uint32_t total = my_tiles * NKB;
        for (uint32_t it = 0; it < total; ++it) {
            const uint32_t s = it % RSTAGES, ph = (it / RSTAGES) & 1u;
            mbar_wait(&bar.full[s], ph);
            split_image(reinterpret_cast<float4 *>(a_hi + size_t(s) * RSTAGE_BYTES),
                        reinterpret_cast<float4 *>(a_lo + size_t(s) * RSTAGE_BYTES), RSTAGE_BYTES / 16, t128);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.split[s]);
        }
    } else {
        const int q = warp & 3;
        for (uint32_t t = 0; t < my_tiles; ++t) {
            const uint32_t acc = t & 1u;
            mbar_wait(&bar.tfull[acc], (t >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[32];
            tmem_ld32(tmem_base + acc * KPG + (uint32_t(q * 32) << 16), r);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.tempty[acc]);
            const uint64_t row = uint64_t(blockIdx.x + t * gridDim.x) * RT + uint64_t(q * 32 + lane);
            if (row < M) {
                float4 *o = reinterpret_cast<float4 *>(C + row * KPG);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    o[c] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                                       __uint_as_float(r[4 * c + 3]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(64)) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_g() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        OC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) throw Error(-3, "cuTensorMapEncodeTiled is not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <int NH>
void launch_gram_tc(const CUtensorMap &map, uint32_t nblocks, uint32_t bbox, double *Out64, cudaStream_t s) {
    using P = Plan<NH>;
    const size_t smem = size_t(2) * P::kStages * P::kStageBytes + sizeof(Bars) + 1024;
    static bool configured = false;
    if (!configured) {
        OC_CUDA(cudaFuncSetAttribute(k_gram_tc<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = true;
    }
    const unsigned grid = unsigned(std::min<uint32_t>(uint32_t(kSMs), nblocks));
    OC_LAUNCH((k_gram_tc<NH>), grid, kThreadsG, smem, s, map, nblocks, bbox, Out64);
}

}  // namespace

bool gram_tc_supported(uint32_t Kc, int kp, uint32_t lda) {
    const char *e = getenv("OCFFM_GRAM_TC");   // read per call (a handful per outer iteration): tests flip it
    const int on = e ? atoi(e) : 1;
    return on && kp == KPG && (Kc == 128 || Kc == 256) && lda % 4 == 0;
}

void gram_stack_tc(const float *A, uint32_t lda, uint32_t Kc, uint32_t bcol, uint32_t row0, uint32_t row1,
                   const float *wvec, double *Out64, double *colsum64, double *wsum64, cudaStream_t s) {
    if (row1 <= row0) return;
    const uint64_t rows = row1 - row0;
    CUtensorMap m;
    const cuuint64_t dims[2] = {Kc, rows};
    const cuuint64_t strides[1] = {cuuint64_t(lda) * sizeof(float)};
    const cuuint32_t box[2] = {32, BR};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn_g()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(A + size_t(row0) * lda),
                                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-3, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
    const uint32_t nblocks = uint32_t((rows + BR - 1) / BR);
    const uint32_t bbox = bcol / 32;
    switch (Kc / 128) {
        case 1: launch_gram_tc<1>(m, nblocks, bbox, Out64, s); break;
        case 2: launch_gram_tc<2>(m, nblocks, bbox, Out64, s); break;
        default: throw Error(-6, "gram_stack_tc: Kc must be 128 or 256");
    }
    if (colsum64) {
        const unsigned blocks = unsigned(std::min<uint64_t>((rows + 7) / 8, uint64_t(kSMs) * 8));
        OC_LAUNCH(k_colsum_w, blocks, 256, 0, s, A + bcol, lda, row0, row1, wvec, colsum64, wsum64);
    }
}

bool rowgemm_tc_supported(uint32_t Ka, int kp, uint32_t lda, uint64_t M) {
    const char *e = getenv("OCFFM_ROWGEMM_TC");
    const int on = e ? atoi(e) : 1;
    return on && kp == KPG && (Ka == 128 || Ka == 256) && lda % 4 == 0 && M >= 4096;
}

template <int NKB>
static void launch_rowgemm_tc(const CUtensorMap &ma, const CUtensorMap &mb, uint32_t ntiles, uint64_t M, float *C,
                              cudaStream_t s) {
    const size_t smem = size_t(2) * RSTAGES * RSTAGE_BYTES + size_t(2) * NKB * 32 * 128 + sizeof(RBars) + 1024;
    static bool configured = false;
    if (!configured) {
        OC_CUDA(cudaFuncSetAttribute(k_rowgemm_tc<NKB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = true;
    }
    const unsigned grid = unsigned(std::min<uint32_t>(uint32_t(kSMs), ntiles));
    OC_LAUNCH((k_rowgemm_tc<NKB>), grid, kThreadsG, smem, s, ma, mb, ntiles, M, C);
}

void rowgemm_tc(const float *A, uint32_t lda, uint32_t Ka, const float *B, float *C, uint64_t M, cudaStream_t s) {
    if (!M) return;
    CUtensorMap ma, mb;
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t dims[2] = {Ka, M};
        const cuuint64_t strides[1] = {cuuint64_t(lda) * sizeof(float)};
        const cuuint32_t box[2] = {32, RT};
        const CUresult r = encode_fn_g()(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(A), dims, strides, box,
                                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw Error(-3, "cuTensorMapEncodeTiled(A) failed with code " + std::to_string(int(r)));
    }
    {
        const cuuint64_t dims[2] = {KPG, Ka};
        const cuuint64_t strides[1] = {cuuint64_t(KPG) * sizeof(float)};
        const cuuint32_t box[2] = {32, Ka};
        const CUresult r = encode_fn_g()(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(B), dims, strides, box,
                                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw Error(-3, "cuTensorMapEncodeTiled(B) failed with code " + std::to_string(int(r)));
    }
    const uint32_t ntiles = uint32_t((M + RT - 1) / RT);
    if (Ka == 128) launch_rowgemm_tc<4>(ma, mb, ntiles, M, C, s);
    else if (Ka == 256) launch_rowgemm_tc<8>(ma, mb, ntiles, M, C, s);
    else throw Error(-6, "rowgemm_tc: Ka must be 128 or 256");
}

}  // namespace ocffm
