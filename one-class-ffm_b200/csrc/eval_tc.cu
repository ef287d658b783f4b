// eval_tc.cu -- the scoring GEMM of the evaluator on the 5th-generation tensor cores (fp32
// contexts): Z = P~va * Q~va^T as 3xTF32 (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM), with
// the top-80 selection fused behind it so scores never reach HBM.
//
// Per CTA (one per SM, 192 threads): a 128-row tile of test rows against a range of items.
//   warp 0        TMA producer: per K-block of 32 floats loads A_hi, A_lo [128 x 32] and B_hi, B_lo
//                 [128 x 32] with cp.async.bulk.tensor (SWIZZLE_128B) into a 3-stage smem ring
//   warp 1        allocates TMEM (2 accumulators x 128 columns) and issues tcgen05.mma
//                 cta_group::1 kind::tf32, M=128 N=128 K=8, three products per K-step;
//                 tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 2..5    epilogue: each warp owns one TMEM lane quadrant = 32 rows, one row per thread;
//                 tcgen05.ld 32 columns at a time, add the item bias, keep what beats the row's
//                 current 80th score in a per-row candidate buffer (global, L2 resident); a warp
//                 compacts a row's buffer to its exact top-80 (score desc, id asc) whenever it
//                 could overflow, which also tightens the row's threshold.
// Output: the same per-(row, item-range) sorted partial lists as the SIMT kernel (eval.cu), merged
// by k_merge_topk.  Exactness: an item is dropped only if 80 better ones are already known.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ocffm {

namespace {

constexpr int TOP = 80;
constexpr int TM = 128, TN = 128, BK = 32;          // tile rows, tile items, floats per K-block
constexpr int STAGES = 3;
constexpr int NACC = 4;                             // TMEM accumulators (4 x 128 = all 512 columns)
constexpr int CAP = 256;                            // candidate slots per (row, item range)
constexpr int KPL = CAP / 32;                       // sort keys per lane
constexpr uint32_t STAGE_BYTES = 4u * (TM * BK * 4u);   // A_hi, A_lo, B_hi, B_lo
constexpr int kThreadsTC = 192;

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
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// same load delivered to the same shared-memory offsets (data and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                               int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0),
        "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row atoms of 1024 B (SBO), sm_100 version bit
__device__ __forceinline__ uint64_t make_desc(const void *smem_tile) {
    uint64_t d = 0;
    d |= uint64_t((smem_u32(smem_tile) >> 4) & 0x3FFFu);
    d |= uint64_t(1024u >> 4) << 32;   // stride byte offset between 8-row groups
    d |= uint64_t(1) << 46;            // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;            // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(TN >> 3) << 17) |
                            (uint32_t(TM >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
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

// (score, id) packed into one descending-sortable 64-bit key: larger score first, then smaller id
__device__ __forceinline__ uint64_t pack_key(float sc, uint32_t id) {
    uint32_t u = __float_as_uint(sc);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (uint64_t(u) << 32) | uint64_t(~id);
}
__device__ __forceinline__ float key_score(uint64_t k) {
    const uint32_t u = uint32_t(k >> 32);
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ uint32_t key_id(uint64_t k) { return ~uint32_t(k); }
// float <-> monotone uint32 (0 is below every finite score): the cross-CTA row thresholds
__device__ __forceinline__ uint32_t ord_of(float sc) {
    const uint32_t u = __float_as_uint(sc);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// Warp-cooperative compaction of one row's candidate buffer (count <= CAP) to its exact top-80 in
// (score desc, id asc) order: an in-register bitonic sort of KPL keys per lane (element
// e = r*32 + lane), fully unrolled so every register index is static.  The sorted survivors are
// written back to the front of the buffer; returns their number, *th = the 80th score if full.
__device__ __noinline__ int compact_row(float *cs, uint32_t *ci, int count, float *th) {
    const int lane = threadIdx.x & 31;
    uint64_t key[KPL];
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const int e = r * 32 + lane;
        key[r] = e < count ? pack_key(cs[e], ci[e]) : 0ull;
    }
#pragma unroll
    for (int k = 2; k <= CAP; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int dr = j >> 5;
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    if ((r & dr) == 0) {
                        const bool desc = ((r * 32) & k) == 0;      // lane bits are below k here
                        const uint64_t a = key[r], b = key[r | dr];
                        const bool swap = desc ? (a < b) : (a > b);
                        key[r] = swap ? b : a;
                        key[r | dr] = swap ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, key[r], j);
                    const int e = r * 32 + lane;
                    const bool desc = (e & k) == 0;
                    const bool lower = (lane & j) == 0;
                    const bool take_max = desc == lower;
                    const uint64_t mx = key[r] > other ? key[r] : other;
                    const uint64_t mn = key[r] > other ? other : key[r];
                    key[r] = take_max ? mx : mn;
                }
            }
        }
    }
    const int n = min(count, TOP);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int e = r * 32 + lane;
        if (e < n) { cs[e] = key_score(key[r]); ci[e] = key_id(key[r]); }
    }
    const float t80 = key_score(__shfl_sync(0xffffffffu, key[2], TOP - 1 - 64));
    if (n == TOP) *th = t80;
    __syncwarp();
    return n;
}

// v[c] for a run-time c without putting v[] into local memory
__device__ __forceinline__ float pick32(const float (&v)[32], int c) {
    float x = v[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) x = (c == i) ? v[i] : x;
    return x;
}

// Two shared-memory plans.  Streaming: A and B K-blocks share a 3-stage ring (any Kc).
// Resident (Kc <= 128): the CTA's A tile (hi + lo, all K-blocks, 128 KB) is loaded once and only
// B streams through the ring -- half the L2->SM operand traffic, which is what bounds this kernel.
constexpr int MAX_RES_KB = 4;
template <bool RESA>
struct SmemTC;
template <>
struct SmemTC<false> {
    float a_hi[STAGES][TM * BK];
    float a_lo[STAGES][TM * BK];
    float b_hi[STAGES][TN * BK];
    float b_lo[STAGES][TN * BK];
    __align__(16) float bias[4][32];   // one private 32-column slice per epilogue warp: no cross-warp barrier
    uint64_t full[STAGES], empty[STAGES], tfull[NACC], tempty[NACC], afull;
    uint32_t tmem_base;
};
template <>
struct SmemTC<true> {
    float a_hi[MAX_RES_KB][TM * BK];
    float a_lo[MAX_RES_KB][TM * BK];
    float b_hi[STAGES][TN * BK];
    float b_lo[STAGES][TN * BK];
    __align__(16) float bias[4][32];   // one private 32-column slice per epilogue warp: no cross-warp barrier
    uint64_t full[STAGES], empty[STAGES], tfull[NACC], tempty[NACC], afull;
    uint32_t tmem_base;
};

// MC = 2: the kernel runs as clusters of two CTAs (two row tiles, same item range).  Rank 0 loads
// every B_hi K-block, rank 1 every B_lo K-block, each with TMA multicast into BOTH CTAs' rings, so
// the L2->SM operand stream -- the bound of this kernel -- is halved.  A stage is refilled only
// after both CTAs' MMAs have retired it: the stage-empty barriers count two multicast commits.
template <bool RESA, int MC>
__global__ void __launch_bounds__(kThreadsTC, 1)
k_score_topk_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                uint32_t Kc, const float *__restrict__ bt, uint32_t row0, uint32_t row1, uint32_t n_ranked,
                uint32_t items_per_split, uint32_t nsplit, const uint8_t *__restrict__ cold,
                float *__restrict__ cand_score, uint32_t *__restrict__ cand_id,
                float *__restrict__ part_score, uint32_t *__restrict__ part_id, uint32_t *row_thr,
                uint32_t dbg) {
    // dbg (OCFFM_TC_DEBUG, measurements only): bit 0 = no row is live (no appends / compactions),
    // bit 1 = the epilogue does not even read the accumulators
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operand tiles must start on 1024-byte boundaries of the shared window
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    SmemTC<RESA> &sm = *reinterpret_cast<SmemTC<RESA> *>(smem_raw + pad);
    constexpr uint32_t kStageBytes = RESA ? 2u * (TN * BK * 4u) : STAGE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t u0 = row0 + blockIdx.x * TM;
    const uint32_t j_lo = blockIdx.y * items_per_split;
    const uint32_t j_hi = min(n_ranked, j_lo + items_per_split);
    const uint32_t ntiles = j_hi > j_lo ? (j_hi - j_lo + TN - 1) / TN : 0;
    const uint32_t nkb = Kc / BK;

    const uint32_t crank = MC > 1 ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], MC); }
        for (int s = 0; s < NACC; ++s) { mbar_init(&sm.tfull[s], 1); mbar_init(&sm.tempty[s], 4); }
        mbar_init(&sm.afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(uint32_t(NACC * TN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC > 1) cluster_sync_all();   // the partner's barriers exist before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (RESA && ntiles) {
                mbar_expect_tx(&sm.afull, nkb * 2u * (TM * BK * 4u));
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    tma_load_2d(sm.a_hi[kb], &map_a_hi, &sm.afull, int(kb * BK), int(u0));
                    tma_load_2d(sm.a_lo[kb], &map_a_lo, &sm.afull, int(kb * BK), int(u0));
                }
            }
            uint32_t it = 0;
            for (uint32_t t = 0; t < ntiles; ++t) {
                const int j0 = int(j_lo + t * TN);
                for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&sm.empty[s], ph ^ 1u);
                    mbar_expect_tx(&sm.full[s], kStageBytes);
                    if (!RESA) {
                        tma_load_2d(sm.a_hi[s], &map_a_hi, &sm.full[s], int(kb * BK), int(u0));
                        tma_load_2d(sm.a_lo[s], &map_a_lo, &sm.full[s], int(kb * BK), int(u0));
                    }
                    if (MC == 1) {
                        tma_load_2d(sm.b_hi[s], &map_b_hi, &sm.full[s], int(kb * BK), j0);
                        tma_load_2d(sm.b_lo[s], &map_b_lo, &sm.full[s], int(kb * BK), j0);
                    } else if (crank == 0) {
                        tma_load_2d_mc(sm.b_hi[s], &map_b_hi, &sm.full[s], int(kb * BK), j0, uint16_t(0x3));
                    } else {
                        tma_load_2d_mc(sm.b_lo[s], &map_b_lo, &sm.full[s], int(kb * BK), j0, uint16_t(0x3));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            if (RESA && ntiles) mbar_wait(&sm.afull, 0);
            uint32_t it = 0;
            for (uint32_t t = 0; t < ntiles; ++t) {
                const uint32_t acc = t % NACC, aph = (t / NACC) & 1u;
                mbar_wait(&sm.tempty[acc], aph ^ 1u);     // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * TN;
                for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&sm.full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t as = RESA ? kb : s;
                    const uint64_t ah = make_desc(sm.a_hi[as]), al = make_desc(sm.a_lo[as]);
                    const uint64_t bh = make_desc(sm.b_hi[s]), bl = make_desc(sm.b_lo[s]);
#pragma unroll
                    for (uint32_t k = 0; k < BK / 8; ++k) {
                        const uint64_t adv = uint64_t((k * 8 * 4) >> 4);   // 32 bytes per K-step
                        umma_tf32(d, ah + adv, bh + adv, (kb | k) != 0 ? 1u : 0u);
                        umma_tf32(d, ah + adv, bl + adv, 1u);
                        umma_tf32(d, al + adv, bh + adv, 1u);
                    }
                    // smem stage reusable once these MMAs retire (in both CTAs when clustered)
                    if (MC == 1) umma_commit(&sm.empty[s]);
                    else umma_commit_mc(&sm.empty[s], uint16_t(0x3));
                }
                umma_commit(&sm.tfull[acc]);              // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
        const int q = warp & 3, ew = warp - 2;
        const uint32_t row = u0 + uint32_t(q) * 32u + uint32_t(lane);
        const bool live = row < row1 && !(cold && cold[row]) && !(dbg & 1u);
        const size_t slot = (size_t(row) * nsplit + blockIdx.y);
        float *cs = cand_score + slot * CAP;
        uint32_t *ci = cand_id + slot * CAP;
        int count = 0;
        float th = live ? -3.0e38f : INFINITY;   // finite floor: -inf (out-of-range columns) never passes
        for (uint32_t t = 0; t < ntiles; ++t) {
            const uint32_t acc = t % NACC, aph = (t / NACC) & 1u;
            const uint32_t j0 = j_lo + t * TN;
            // Another CTA working on a different item range of the same row may already know 80 items
            // better than anything seen here: its 80th score is a valid bound for this row too.
            if (live) {
                const uint32_t g = __ldcg(row_thr + row);
                if (g) th = fmaxf(th, ord_to_float(g));
            }
            mbar_wait(&sm.tfull[acc], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + acc * TN + (uint32_t(q * 32) << 16);
            float bnext;   // item bias of the next 32 columns, fetched one chunk ahead
            {
                const uint32_t j = j0 + uint32_t(lane);
                bnext = j < j_hi ? bt[j] : -INFINITY;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < ((dbg & 2u) ? 0 : TN); c0 += 32) {
                // make room: a chunk can add at most 32 candidates to a row
                uint32_t need = __ballot_sync(0xffffffffu, count > CAP - 32);
                while (need) {
                    const int l = __ffs(need) - 1;
                    need &= need - 1;
                    float *rcs = reinterpret_cast<float *>(__shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(cs), l));
                    uint32_t *rci = reinterpret_cast<uint32_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(ci), l));
                    const int rc = __shfl_sync(0xffffffffu, count, l);
                    float nth = -3.0e38f;
                    const int n = compact_row(rcs, rci, rc, &nth);
                    if (lane == l) {
                        count = n;
                        if (n == TOP) {
                            th = fmaxf(th, nth);
                            atomicMax(row_thr + row, ord_of(th));   // publish to the row's other item ranges
                        }
                    }
                    __syncwarp();
                }
                __syncwarp();
                sm.bias[ew][lane] = bnext;   // -inf past the item range: such a column can never pass
                __syncwarp();
                if (c0 + 32 < TN) {
                    const uint32_t j = j0 + uint32_t(c0 + 32 + lane);
                    bnext = j < j_hi ? bt[j] : -INFINITY;
                }
                uint32_t r[32];
                tmem_ld32(taddr + uint32_t(c0), r);
                float v[32];
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(&sm.bias[ew][c]);
                    v[c + 0] = __uint_as_float(r[c + 0]) + b4.x;
                    v[c + 1] = __uint_as_float(r[c + 1]) + b4.y;
                    v[c + 2] = __uint_as_float(r[c + 2]) + b4.z;
                    v[c + 3] = __uint_as_float(r[c + 3]) + b4.w;
                }
                // branch-free filter: one compare + one predicated OR per column builds the lane's
                // mask of passing columns; the (per lane) rare appends are done afterwards
                uint32_t m = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (v[c] >= th) m |= 1u << c;
                while (m) {
                    const int c = __ffs(m) - 1;
                    m &= m - 1;
                    cs[count] = pick32(v, c);
                    ci[count] = j0 + uint32_t(c0 + c);
                    ++count;
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tempty[acc]);
        }
        // final exact top-80 of every row of this warp, written as the sorted partial list
        for (int l = 0; l < 32; ++l) {
            float *rcs = reinterpret_cast<float *>(__shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(cs), l));
            uint32_t *rci = reinterpret_cast<uint32_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(ci), l));
            const int rc = __shfl_sync(0xffffffffu, count, l);
            const uint32_t rrow = __shfl_sync(0xffffffffu, row, l);
            if (rrow >= row1) continue;
            float unused = 0.f;
            const int n = compact_row(rcs, rci, rc, &unused);
            const size_t o = (size_t(rrow) * nsplit + blockIdx.y) * TOP;
            for (int s = lane; s < TOP; s += 32) {
                part_id[o + s] = s < n ? rci[s] : 0xffffffffu;
                part_score[o + s] = s < n ? rcs[s] : 0.f;
            }
            __syncwarp();
        }
    }
    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC > 1) cluster_sync_all();   // nobody leaves while the partner may still signal its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(NACC * TN)) : "memory");
    }
}

// X -> (hi, lo): hi = round-to-nearest TF32 of x, lo = x - hi (exact in fp32)
__global__ void __launch_bounds__(256)
k_split_tf32(const float *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo, uint64_t n) {
    pdl_enter();
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const float v = x[i];
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hf = __uint_as_float(h);
        hi[i] = hf;
        lo[i] = v - hf;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

CUtensorMap make_map(const float *base, uint64_t rows, uint32_t Kc) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {Kc, rows};
    const cuuint64_t strides[1] = {cuuint64_t(Kc) * sizeof(float)};
    const cuuint32_t box[2] = {BK, TM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-3, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
    return m;
}

}  // namespace

bool score_topk_tc_supported(uint32_t Kc) { return Kc % BK == 0 && Kc >= BK; }
size_t score_topk_tc_cand_slots() { return CAP; }

uint32_t score_topk_tc_splits(uint32_t rows, uint32_t n_ranked) {
    const uint32_t tiles_u = (rows + TM - 1) / TM;
    const uint32_t item_tiles = (n_ranked + TN - 1) / TN;
    uint32_t ns = (4 * kSMs + tiles_u - 1) / std::max(1u, tiles_u);    // ~4 waves of one CTA per SM
    ns = std::min(ns, std::max(1u, item_tiles / 16));
    return std::max(1u, std::min(ns, 32u));
}

void split_tf32(const float *x, float *hi, float *lo, uint64_t n, cudaStream_t s) {
    if (!n) return;
    OC_LAUNCH(k_split_tf32, unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(kSMs) * 16)), 256, 0, s, x, hi, lo, n);
}

void score_topk_tc(const float *Phi, const float *Plo, uint64_t p_rows, const float *Qhi, const float *Qlo,
                   uint64_t q_rows, uint32_t Kc, const float *bt, uint32_t row0, uint32_t row1,
                   uint32_t n_ranked, const uint8_t *cold, uint32_t nsplit, float *cand_score,
                   uint32_t *cand_id, float *part_score, uint32_t *part_id, uint32_t *row_thr, cudaStream_t s) {
    if (row1 <= row0) return;
    OC_CUDA(cudaMemsetAsync(row_thr + row0, 0, size_t(row1 - row0) * sizeof(uint32_t), s));
    const CUtensorMap ma_hi = make_map(Phi, p_rows, Kc), ma_lo = make_map(Plo, p_rows, Kc);
    const CUtensorMap mb_hi = make_map(Qhi, q_rows, Kc), mb_lo = make_map(Qlo, q_rows, Kc);
    const uint32_t item_tiles = (n_ranked + TN - 1) / TN;
    const uint32_t per = ((item_tiles + nsplit - 1) / nsplit) * TN;
    unsigned tiles = unsigned((uint64_t(row1 - row0) + TM - 1) / TM);
    static int use_mc = -1;
    if (use_mc < 0) {
        const char *e = getenv("OCFFM_EVAL_MC");
        use_mc = e ? atoi(e) : 2;
    }
    const int mc = (use_mc >= 2 && tiles >= 2) ? 2 : 1;
    if (mc == 2) tiles = (tiles + 1) / 2 * 2;   // a padding CTA (all rows dead) keeps the pair complete
    const bool resa = Kc / BK <= MAX_RES_KB;
    const size_t smem = (resa ? sizeof(SmemTC<true>) : sizeof(SmemTC<false>)) + 1024;
    static int dbg = -1;
    if (dbg < 0) {
        const char *e = getenv("OCFFM_TC_DEBUG");
        dbg = e ? atoi(e) : 0;
    }
    void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, uint32_t, const float *, uint32_t, uint32_t,
                 uint32_t, uint32_t, uint32_t, const uint8_t *, float *, uint32_t *, float *, uint32_t *, uint32_t *,
                 uint32_t) =
        resa ? (mc == 2 ? k_score_topk_tc<true, 2> : k_score_topk_tc<true, 1>)
             : (mc == 2 ? k_score_topk_tc<false, 2> : k_score_topk_tc<false, 1>);
    OC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles, nsplit);
    cfg.blockDim = dim3(kThreadsTC);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(mc);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OC_CUDA(cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, Kc, bt, row0, row1, n_ranked, per, nsplit,
                               cold, cand_score, cand_id, part_score, part_id, row_thr, uint32_t(dbg)));
    count_launch();
}

}  // namespace ocffm
