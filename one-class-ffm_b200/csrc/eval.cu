// eval.cu -- full user x item scoring fused with warp-level top-80 (north_star subsystem 4).
//
// One CTA owns a tile of TU=64 test rows and sweeps every ranked item in tiles of TI=128:
// the score tile Z = P~va_tile * Q~va_tile^T (+ bt) is produced by a register-tiled SIMT GEMM
// (4 rows x 8 items per thread) into shared memory and consumed in place by the top-k stage, so
// the m_t x n score matrix (43 GB at the KKBox shape) never exists in HBM.  Each warp then owns
// 8 rows and keeps, per row, a sorted list of the best 80 (score, id) pairs in shared memory:
// lanes compare the tile against the row's current 80th score, ballot the survivors and insert
// them in ascending item order, which reproduces the reference's repeated first-argmax
// (ffm.cpp:1029-1046, 1074-1108): equal scores rank by lower item id.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace ocffm {

namespace {

constexpr int kThreads = 256;
constexpr int TU = 64, TI = 128, TOP = 80;
__constant__ double c_gain[TOP];   // 1 / log2(rank + 2)
bool g_gain_ready[64] = {false};

template <typename T>
struct TopList {
    T *score;        // [TOP] sorted by (score desc, id asc)
    uint32_t *id;    // [TOP]
};

// Warp-cooperative insertion of the candidates flagged in this round: lane l proposes (v, item)
// when cand is set.  The list order is (score desc, item id asc), which is exactly the order in
// which the reference's repeated first-argmax (ffm.cpp:1033, 1078) emits items, so candidates
// may arrive in any item order.  n = current length; once full, (th, th_id) mirror entry TOP-1.
template <typename T>
__device__ __forceinline__ void topk_insert(TopList<T> L, int &n, T &th, uint32_t &th_id, T v,
                                            uint32_t item, bool cand) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t m = __ballot_sync(0xffffffffu, cand);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const T cv = __shfl_sync(0xffffffffu, v, src);
        const uint32_t ci = __shfl_sync(0xffffffffu, item, src);
        if (n == TOP && !(cv > th || (cv == th && ci < th_id))) continue;
        int pos = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int s = int(lane) + 32 * q;
            bool ahead = false;
            if (s < n) {
                const T es = L.score[s];
                ahead = es > cv || (es == cv && L.id[s] < ci);
            }
            pos += __popc(__ballot_sync(0xffffffffu, ahead));
        }
        const int last = min(n, TOP - 1);
        T ts[3];
        uint32_t ti[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int s = int(lane) + 32 * q;
            if (s > pos && s <= last) { ts[q] = L.score[s - 1]; ti[q] = L.id[s - 1]; }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int s = int(lane) + 32 * q;
            if (s > pos && s <= last) { L.score[s] = ts[q]; L.id[s] = ti[q]; }
        }
        if (lane == 0) { L.score[pos] = cv; L.id[pos] = ci; }
        __syncwarp();
        n = min(n + 1, TOP);
        if (n == TOP) { th = L.score[TOP - 1]; th_id = L.id[TOP - 1]; }
    }
}

// XOR swizzle of the transposed operand tiles: element (k, r) lives at k*LD + (r ^ swz(k)).
// It permutes aligned groups of 4 rows, so the 16-byte reads of the micro-kernel stay aligned
// and conflict-free while the transposing stores of a warp (4 rows x 8 k-quads) hit 32 banks.
__device__ __forceinline__ constexpr int swz(int k) { return ((k >> 2) & 7) << 2; }

// Slow path of the fused scorer, one instance in the binary: the 128 scores of one row of the
// current tile (staged by the owning warp in `scratch`) are offered to the row's list.
template <typename T>
__device__ __noinline__ void topk_offer_row(const T *scratch, T *lscore_u, uint32_t *lid_u, int *lcnt_u,
                                            uint32_t j0, uint32_t j_hi) {
    const uint32_t lane = threadIdx.x & 31u;
    TopList<T> L{lscore_u, lid_u};
    int n = *lcnt_u;
    T th = n == TOP ? lscore_u[TOP - 1] : T(0);
    uint32_t th_id = n == TOP ? lid_u[TOP - 1] : 0u;
    for (int q = 0; q < TI / 32; ++q) {
        const uint32_t item = j0 + q * 32 + lane;
        const T v = scratch[q * 32 + lane];
        const bool cand = item < j_hi && (n < TOP || v > th || (v == th && item < th_id));
        topk_insert(L, n, th, th_id, v, item, cand);
    }
    if (lane == 0) *lcnt_u = n;
    __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
k_score_topk(const T *__restrict__ Pva, const T *__restrict__ Qva, uint32_t Kc,
             const T *__restrict__ bt, uint32_t row0, uint32_t row1, uint32_t n_ranked,
             uint32_t items_per_split, uint32_t nsplit, const uint8_t *__restrict__ cold,
             T *__restrict__ part_score, uint32_t *__restrict__ part_id) {
    pdl_enter();
    constexpr int BK = 32;
    constexpr int NA = TU * (BK / 4) / kThreads;   // float4 loads per thread for the row tile (2)
    constexpr int NB = TI * (BK / 4) / kThreads;   // ... for the item tile (4)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *As = reinterpret_cast<T *>(smem_raw);                 // [BK][TU] swizzled
    T *Bs = As + BK * TU;                                    // [BK][TI] swizzled
    T *lscore = Bs + BK * TI;                                // [TU][TOP]
    uint32_t *lid = reinterpret_cast<uint32_t *>(lscore + TU * TOP);  // [TU][TOP]
    int *lcnt = reinterpret_cast<int *>(lid + TU * TOP);     // [TU]
    T *scratch = reinterpret_cast<T *>(lcnt + TU);           // [8 warps][TI]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = tid % 16, tu = tid / 16;                  // 16 x 16 thread grid, 4 rows x 8 items each
    const uint64_t u0 = uint64_t(row0) + uint64_t(blockIdx.x) * TU;
    const uint32_t j_lo = blockIdx.y * items_per_split;
    const uint32_t j_hi = min(n_ranked, j_lo + items_per_split);
    if (tid < TU) lcnt[tid] = 0;

    const uint32_t nk = (Kc + BK - 1) / BK;
    const uint32_t ntiles = j_hi > j_lo ? (j_hi - j_lo + TI - 1) / TI : 0;
    const uint32_t nchunks = ntiles * nk;
    V4<T> ra[NA], rb[NB];
    auto prefetch = [&](uint32_t chunk) {
        const uint32_t j0 = j_lo + (chunk / nk) * TI, k0 = (chunk % nk) * BK;
#pragma unroll
        for (int x = 0; x < NA; ++x) {
            const int e = tid + x * kThreads, r = e / (BK / 4), c = (e % (BK / 4)) * 4;
            ra[x] = (u0 + r < row1 && k0 + c < Kc) ? ldg4(Pva + (u0 + r) * Kc + k0 + c) : zero4<T>();
        }
#pragma unroll
        for (int x = 0; x < NB; ++x) {
            const int e = tid + x * kThreads, r = e / (BK / 4), c = (e % (BK / 4)) * 4;
            rb[x] = (j0 + r < j_hi && k0 + c < Kc) ? ldg4(Qva + uint64_t(j0 + r) * Kc + k0 + c) : zero4<T>();
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int x = 0; x < NA; ++x) {
            const int e = tid + x * kThreads, r = e / (BK / 4), c = (e % (BK / 4)) * 4;
            const int rs = r ^ swz(c);
            As[(c + 0) * TU + rs] = ra[x].x;
            As[(c + 1) * TU + rs] = ra[x].y;
            As[(c + 2) * TU + rs] = ra[x].z;
            As[(c + 3) * TU + rs] = ra[x].w;
        }
#pragma unroll
        for (int x = 0; x < NB; ++x) {
            const int e = tid + x * kThreads, r = e / (BK / 4), c = (e % (BK / 4)) * 4;
            const int rs = r ^ swz(c);
            Bs[(c + 0) * TI + rs] = rb[x].x;
            Bs[(c + 1) * TI + rs] = rb[x].y;
            Bs[(c + 2) * TI + rs] = rb[x].z;
            Bs[(c + 3) * TI + rs] = rb[x].w;
        }
    };

    T acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = T(0);
    if (nchunks) prefetch(0);
    for (uint32_t chunk = 0; chunk < nchunks; ++chunk) {
        __syncthreads();            // previous chunk fully consumed
        stash();
        __syncthreads();
        if (chunk + 1 < nchunks) prefetch(chunk + 1);   // global latency hidden behind the FMAs below
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const V4<T> a = ld4(As + kk * TU + ((tu * 4) ^ swz(kk)));
            const V4<T> b0 = ld4(Bs + kk * TI + ((ti * 4) ^ swz(kk)));
            const V4<T> b1 = ld4(Bs + kk * TI + ((64 + ti * 4) ^ swz(kk)));
            const T av[4] = {a.x, a.y, a.z, a.w};
            const T bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] += av[i] * bv[j];
        }
        if ((chunk + 1) % nk != 0) continue;
        // ---- tile finished: top-k straight from the accumulator registers ----------------------
        // warp w holds tile rows 8w..8w+7: lanes 0-15 rows 8w..8w+3, lanes 16-31 rows 8w+4..8w+7
        const uint32_t j0 = j_lo + (chunk / nk) * TI;
        uint32_t item8[8];
        T b8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            item8[j] = j0 + (j < 4 ? ti * 4 + j : 64 + ti * 4 + (j - 4));
            b8[j] = item8[j] < j_hi ? bt[item8[j]] : T(0);
        }
        T *my_scratch = scratch + warp * TI;
#pragma unroll
        for (int hr = 0; hr < 8; ++hr) {
            const int u = warp * 8 + hr;
            if (u0 + u < row1 && !(cold && cold[u0 + u])) {
                // cheap register-only pre-check: does any score of this row beat the row's 80th?
                const int n = lcnt[u];
                const T th = n == TOP ? lscore[u * TOP + TOP - 1] : T(0);
                const uint32_t th_id = n == TOP ? lid[u * TOP + TOP - 1] : 0u;
                const bool mine = (lane >> 4) == (hr >> 2);
                bool any = false;
                T v8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v8[j] = acc[hr & 3][j] + b8[j];
                    any |= item8[j] < j_hi && (n < TOP || v8[j] > th || (v8[j] == th && item8[j] < th_id));
                }
                if (__ballot_sync(0xffffffffu, mine && any)) {
                    if (mine) {
                        st4(my_scratch + ti * 4, V4<T>{v8[0], v8[1], v8[2], v8[3]});
                        st4(my_scratch + 64 + ti * 4, V4<T>{v8[4], v8[5], v8[6], v8[7]});
                    }
                    __syncwarp();
                    topk_offer_row(my_scratch, lscore + u * TOP, lid + u * TOP, lcnt + u, j0, j_hi);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = T(0);
    }
    __syncthreads();
    for (int uu = 0; uu < TU / 8; ++uu) {
        const int u = warp * (TU / 8) + uu;
        if (u0 + u >= row1) break;
        const int n = lcnt[u];
        const size_t o = ((u0 + u) * nsplit + blockIdx.y) * TOP;
        for (int s = lane; s < TOP; s += 32) {
            part_id[o + s] = s < n ? lid[u * TOP + s] : 0xffffffffu;
            part_score[o + s] = s < n ? lscore[u * TOP + s] : T(0);
        }
    }
}

// merge the nsplit (<= 32) sorted partial lists of one row: lane l walks list l, 80 rounds of a
// warp arg-max under (score desc, id asc)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_merge_topk(const T *__restrict__ part_score, const uint32_t *__restrict__ part_id, uint32_t nsplit,
             uint32_t row0, uint32_t row1, uint32_t *__restrict__ ids) {
    pdl_enter();
    const uint64_t row = uint64_t(row0) + ((uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (row >= row1) return;
    const size_t o = (row * nsplit + lane) * TOP;
    int head = 0;
    for (int rank = 0; rank < TOP; ++rank) {
        uint32_t id = 0xffffffffu;
        T sc = T(0);
        if (lane < nsplit && head < TOP) {
            id = part_id[o + head];
            sc = part_score[o + head];
        }
        T bs = sc;
        uint32_t bi = id;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const T os = __shfl_xor_sync(0xffffffffu, bs, off);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            const bool take = oi != 0xffffffffu && (bi == 0xffffffffu || os > bs || (os == bs && oi < bi));
            if (take) { bs = os; bi = oi; }
        }
        if (bi != 0xffffffffu && bi == id) ++head;
        if (lane == 0) ids[row * TOP + rank] = bi;
    }
}

template <typename T>
__global__ void __launch_bounds__(32)
k_vector_topk(const T *__restrict__ z, uint32_t n_ranked, uint32_t *__restrict__ ids80) {
    pdl_enter();
    __shared__ T sc[TOP];
    __shared__ uint32_t id[TOP];
    TopList<T> L{sc, id};
    const uint32_t lane = threadIdx.x;
    int n = 0;
    T th = T(0);
    uint32_t th_id = 0;
    for (uint32_t j0 = 0; j0 < n_ranked; j0 += 32) {
        const uint32_t item = j0 + lane;
        const T v = item < n_ranked ? z[item] : T(0);
        const bool cand = item < n_ranked && (n < TOP || v > th || (v == th && item < th_id));
        topk_insert(L, n, th, th_id, v, item, cand);
    }
    __syncwarp();
    for (int s = lane; s < TOP; s += 32) ids80[s] = s < n ? id[s] : 0xffffffffu;
}

// one warp per test row: P@K / nDCG@K pieces (prec_k, ndcg: ffm.cpp:1018-1128) and ploss (982-986)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_eval_metrics(const uint32_t *__restrict__ ids, const uint32_t *__restrict__ cold_ids80,
               const uint8_t *__restrict__ cold, const uint32_t *__restrict__ lab_rowptr,
               const uint32_t *__restrict__ lab_idx, uint32_t row0, uint32_t row1,
               const T *__restrict__ Pva, const T *__restrict__ Qva, uint32_t Kc,
               const T *__restrict__ at, const T *__restrict__ bt, const T *__restrict__ popular,
               uint32_t n_items, uint32_t n_ranked, double *__restrict__ acc64) {
    pdl_enter();
    const uint64_t row = uint64_t(row0) + ((uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (row >= row1) return;
    const bool is_cold = cold && cold[row];
    const uint32_t *my = is_cold ? cold_ids80 : ids + row * TOP;
    const uint32_t lb = lab_rowptr[row], nl = lab_rowptr[row + 1] - lb;
    const uint32_t ranked = min(uint32_t(TOP), n_ranked);
    double hits[5] = {0, 0, 0, 0, 0}, dcg[5] = {0, 0, 0, 0, 0}, idcg[5] = {0, 0, 0, 0, 0};
    const int cut[5] = {5, 10, 20, 40, 80};
    for (uint32_t rank = lane; rank < ranked; rank += 32) {
        const uint32_t item = my[rank];
        bool hit = false;
        if (item != 0xffffffffu)
            for (uint32_t t = 0; t < nl; ++t)
                if (lab_idx[lb + t] == item) { hit = true; break; }
        const double g = c_gain[rank];
#pragma unroll
        for (int s = 0; s < 5; ++s)
            if (int(rank) < cut[s]) {
                if (hit) { hits[s] += 1.0; dcg[s] += g; }
                if (nl > rank) idcg[s] += g;
            }
    }
    // ploss: (1 - z_j - at_i)^2 over the row's labels that fall inside the score vector
    double pl = 0;
    const double ati = double(at[row]);
    for (uint32_t t = lane; t < nl; t += 32) {
        const uint32_t j = lab_idx[lb + t];
        if (is_cold) {
            if (j < n_ranked) { const double e = 1.0 - double(popular[j]) - ati; pl += e * e; }
        } else if (j < n_items) {
            T z = bt[j];
            const T *p = Pva + row * Kc, *q = Qva + uint64_t(j) * Kc;
            for (uint32_t c = 0; c < Kc; c += 4) z += dot4(ldg4(p + c), ldg4(q + c));
            const double e = 1.0 - double(z) - ati;
            pl += e * e;
        }
    }
    pl = warp_sum(pl);
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        hits[s] = warp_sum(hits[s]);
        dcg[s] = warp_sum(dcg[s]);
        idcg[s] = warp_sum(idcg[s]);
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            atomicAdd(acc64 + s, hits[s]);
            atomicAdd(acc64 + 5 + s, dcg[s] / idcg[s]);   // 0/0 = NaN exactly like the reference
        }
        atomicAdd(acc64 + 10, pl);
    }
}

void ensure_gain_table() {
    int dev = 0;
    OC_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && g_gain_ready[dev]) return;
    double h[TOP];
    for (int i = 0; i < TOP; ++i) h[i] = 1.0 / log2(double(i) + 2.0);
    OC_CUDA(cudaMemcpyToSymbol(c_gain, h, sizeof(h)));
    if (dev < 64) g_gain_ready[dev] = true;
}

template <typename T>
size_t score_smem_bytes() {
    return sizeof(T) * (32 * TU + 32 * TI + TU * TOP + 8 * TI) + sizeof(uint32_t) * TU * TOP +
           sizeof(int) * TU + 16;
}

}  // namespace

uint32_t score_topk_splits(uint32_t rows, uint32_t n_ranked) {
    // enough (row tile, item range) units for ~4 waves of 2 CTAs/SM, ranges of >= 8 item tiles
    const uint32_t tiles_u = (rows + TU - 1) / TU;
    const uint32_t item_tiles = (n_ranked + TI - 1) / TI;
    uint32_t ns = (8 * kSMs + tiles_u - 1) / std::max(1u, tiles_u);
    ns = std::min(ns, std::max(1u, item_tiles / 8));
    return std::max(1u, std::min(ns, 32u));
}

template <typename T>
void score_topk(const T *Pva, const T *Qva, uint32_t Kc, const T *bt, uint32_t row0, uint32_t row1,
                uint32_t n_ranked, const uint8_t *cold, uint32_t nsplit, T *part_score,
                uint32_t *part_id, uint32_t *ids, cudaStream_t s) {
    if (row1 <= row0) return;
    const size_t smem = score_smem_bytes<T>();
    OC_CUDA(cudaFuncSetAttribute(k_score_topk<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const uint32_t item_tiles = (n_ranked + TI - 1) / TI;
    const uint32_t per = ((item_tiles + nsplit - 1) / nsplit) * TI;
    const dim3 grid(unsigned((uint64_t(row1 - row0) + TU - 1) / TU), nsplit);
    OC_LAUNCH((k_score_topk<T>), grid, kThreads, smem, s, Pva, Qva, Kc, bt, row0, row1, n_ranked, per,
              nsplit, cold, part_score, part_id);
    merge_topk<T>(part_score, part_id, nsplit, row0, row1, ids, s);
}

template <typename T>
void merge_topk(const T *part_score, const uint32_t *part_id, uint32_t nsplit, uint32_t row0,
                uint32_t row1, uint32_t *ids, cudaStream_t s) {
    if (row1 <= row0) return;
    const unsigned blocks = unsigned((uint64_t(row1 - row0) * 32 + kThreads - 1) / kThreads);
    OC_LAUNCH((k_merge_topk<T>), blocks, kThreads, 0, s, part_score, part_id, nsplit, row0, row1, ids);
}

template <typename T>
void vector_topk(const T *z, uint32_t n_ranked, uint32_t *ids80, cudaStream_t s) {
    OC_LAUNCH((k_vector_topk<T>), 1, 32, 0, s, z, n_ranked, ids80);
}

template <typename T>
void eval_metrics(const uint32_t *ids, const uint32_t *cold_ids80, const uint8_t *cold,
                  const uint32_t *lab_rowptr, const uint32_t *lab_idx, uint32_t row0, uint32_t row1,
                  const T *Pva, const T *Qva, uint32_t Kc, const T *at, const T *bt,
                  const T *popular, uint32_t n_items, uint32_t n_ranked, double *acc64,
                  cudaStream_t s) {
    if (row1 <= row0) return;
    ensure_gain_table();
    const unsigned blocks = unsigned((uint64_t(row1 - row0) * 32 + kThreads - 1) / kThreads);
    OC_LAUNCH((k_eval_metrics<T>), blocks, kThreads, 0, s, ids, cold_ids80, cold, lab_rowptr, lab_idx,
              row0, row1, Pva, Qva, Kc, at, bt, popular, n_items, n_ranked, acc64);
}

#define OC_INSTANTIATE(T)                                                                          \
    template void score_topk<T>(const T *, const T *, uint32_t, const T *, uint32_t, uint32_t,     \
                                uint32_t, const uint8_t *, uint32_t, T *, uint32_t *, uint32_t *,  \
                                cudaStream_t);                                                     \
    template void vector_topk<T>(const T *, uint32_t, uint32_t *, cudaStream_t);                   \
    template void merge_topk<T>(const T *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t *, \
                                cudaStream_t);                                                     \
    template void eval_metrics<T>(const uint32_t *, const uint32_t *, const uint8_t *,             \
                                  const uint32_t *, const uint32_t *, uint32_t, uint32_t,          \
                                  const T *, const T *, uint32_t, const T *, const T *, const T *, \
                                  uint32_t, uint32_t, double *, cudaStream_t);

OC_INSTANTIATE(float)
OC_INSTANTIATE(double)

}  // namespace ocffm
