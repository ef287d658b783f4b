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
    T *score;        // [TOP]
    uint32_t *id;    // [TOP]
};

// Insert the candidates of one 32-wide batch (lane `l` proposes value v for item `item_base + l`
// when cand is set) into the sorted list; n = current length, th = score[TOP-1] once full.
template <typename T>
__device__ __forceinline__ void topk_insert_batch(TopList<T> L, int &n, T &th, T v, bool cand,
                                                  uint32_t item_base) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t m = __ballot_sync(0xffffffffu, cand);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const T cv = __shfl_sync(0xffffffffu, v, src);
        if (n == TOP && !(cv > th)) continue;
        // slot = number of kept entries with score >= cv (they all have smaller item ids)
        int pos = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int s = int(lane) + 32 * q;
            pos += __popc(__ballot_sync(0xffffffffu, s < n && L.score[s] >= cv));
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
        if (lane == 0) { L.score[pos] = cv; L.id[pos] = item_base + uint32_t(src); }
        __syncwarp();
        n = min(n + 1, TOP);
        if (n == TOP) th = L.score[TOP - 1];
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_score_topk(const T *__restrict__ Pva, const T *__restrict__ Qva, uint32_t Kc,
             const T *__restrict__ bt, uint32_t row0, uint32_t row1, uint32_t n_ranked,
             const uint8_t *__restrict__ cold, uint32_t *__restrict__ ids) {
    constexpr int BK = 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *As = reinterpret_cast<T *>(smem_raw);                 // [BK][TU + 4]
    T *Bs = As + BK * (TU + 4);                              // [BK][TI + 4]
    T *Sc = Bs + BK * (TI + 4);                              // [TU][TI]
    T *lscore = Sc + TU * TI;                                // [TU][TOP]
    uint32_t *lid = reinterpret_cast<uint32_t *>(lscore + TU * TOP);  // [TU][TOP]
    int *lcnt = reinterpret_cast<int *>(lid + TU * TOP);     // [TU]
    T *lthr = reinterpret_cast<T *>(lcnt + TU);              // [TU]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = tid % 16, tu = tid / 16;                  // 16 x 16 thread grid
    const uint64_t u0 = uint64_t(row0) + uint64_t(blockIdx.x) * TU;
    if (tid < TU) { lcnt[tid] = 0; lthr[tid] = T(0); }
    __syncthreads();

    for (uint32_t j0 = 0; j0 < n_ranked; j0 += TI) {
        T acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = T(0);
        for (uint32_t k0 = 0; k0 < Kc; k0 += BK) {
            for (int e = tid; e < TU * (BK / 4); e += kThreads) {
                const int r = e / (BK / 4), c = (e % (BK / 4)) * 4;
                V4<T> v = zero4<T>();
                if (u0 + r < row1 && k0 + c < Kc) v = ldg4(Pva + (u0 + r) * Kc + k0 + c);
                As[(c + 0) * (TU + 4) + r] = v.x;
                As[(c + 1) * (TU + 4) + r] = v.y;
                As[(c + 2) * (TU + 4) + r] = v.z;
                As[(c + 3) * (TU + 4) + r] = v.w;
            }
            for (int e = tid; e < TI * (BK / 4); e += kThreads) {
                const int r = e / (BK / 4), c = (e % (BK / 4)) * 4;
                V4<T> v = zero4<T>();
                if (j0 + r < n_ranked && k0 + c < Kc) v = ldg4(Qva + uint64_t(j0 + r) * Kc + k0 + c);
                Bs[(c + 0) * (TI + 4) + r] = v.x;
                Bs[(c + 1) * (TI + 4) + r] = v.y;
                Bs[(c + 2) * (TI + 4) + r] = v.z;
                Bs[(c + 3) * (TI + 4) + r] = v.w;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < BK; ++kk) {
                const V4<T> a = ld4(As + kk * (TU + 4) + tu * 4);
                const V4<T> b0 = ld4(Bs + kk * (TI + 4) + ti * 4);
                const V4<T> b1 = ld4(Bs + kk * (TI + 4) + 64 + ti * 4);
                const T av[4] = {a.x, a.y, a.z, a.w};
                const T bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] += av[i] * bv[j];
            }
            __syncthreads();
        }
        // scores (+ item bias) into the shared tile
        {
            V4<T> bb0 = zero4<T>(), bb1 = zero4<T>();
            const uint32_t ja = j0 + ti * 4, jb = j0 + 64 + ti * 4;
            T b8[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b8[j] = (ja + j < n_ranked) ? bt[ja + j] : T(0);
                b8[4 + j] = (jb + j < n_ranked) ? bt[jb + j] : T(0);
            }
            (void)bb0; (void)bb1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                V4<T> s0 = {acc[i][0] + b8[0], acc[i][1] + b8[1], acc[i][2] + b8[2], acc[i][3] + b8[3]};
                V4<T> s1 = {acc[i][4] + b8[4], acc[i][5] + b8[5], acc[i][6] + b8[6], acc[i][7] + b8[7]};
                st4(Sc + (tu * 4 + i) * TI + ti * 4, s0);
                st4(Sc + (tu * 4 + i) * TI + 64 + ti * 4, s1);
            }
        }
        __syncthreads();
        // warp-level top-k: warp w owns rows w*8 .. w*8+7 of the tile
        for (int uu = 0; uu < TU / 8; ++uu) {
            const int u = warp * (TU / 8) + uu;
            if (u0 + u >= row1) break;
            if (cold && cold[u0 + u]) continue;
            TopList<T> L{lscore + u * TOP, lid + u * TOP};
            int n = lcnt[u];
            T th = lthr[u];
#pragma unroll
            for (int q = 0; q < TI / 32; ++q) {
                const uint32_t item = j0 + q * 32 + lane;
                const T v = Sc[u * TI + q * 32 + lane];
                const bool cand = item < n_ranked && (n < TOP || v > th);
                topk_insert_batch(L, n, th, v, cand, j0 + q * 32);
            }
            if (lane == 0) { lcnt[u] = n; lthr[u] = th; }
        }
        __syncthreads();
    }
    for (int uu = 0; uu < TU / 8; ++uu) {
        const int u = warp * (TU / 8) + uu;
        if (u0 + u >= row1) break;
        const int n = lcnt[u];
        for (int s = lane; s < TOP; s += 32)
            ids[(u0 + u) * TOP + s] = s < n ? lid[u * TOP + s] : 0xffffffffu;
    }
}

template <typename T>
__global__ void __launch_bounds__(32)
k_vector_topk(const T *__restrict__ z, uint32_t n_ranked, uint32_t *__restrict__ ids80) {
    __shared__ T sc[TOP];
    __shared__ uint32_t id[TOP];
    TopList<T> L{sc, id};
    const uint32_t lane = threadIdx.x;
    int n = 0;
    T th = T(0);
    for (uint32_t j0 = 0; j0 < n_ranked; j0 += 32) {
        const uint32_t item = j0 + lane;
        const T v = item < n_ranked ? z[item] : T(0);
        const bool cand = item < n_ranked && (n < TOP || v > th);
        topk_insert_batch(L, n, th, v, cand, j0);
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
    return sizeof(T) * (32 * (TU + 4) + 32 * (TI + 4) + TU * TI + TU * TOP) +
           sizeof(uint32_t) * TU * TOP + sizeof(int) * TU + sizeof(T) * TU + 16;
}

}  // namespace

template <typename T>
void score_topk(const T *Pva, const T *Qva, uint32_t Kc, const T *bt, uint32_t row0, uint32_t row1,
                uint32_t n_ranked, const uint8_t *cold, uint32_t *ids, cudaStream_t s) {
    if (row1 <= row0) return;
    const size_t smem = score_smem_bytes<T>();
    static bool attr_set[2] = {false, false};
    if (!attr_set[sizeof(T) == 8]) {
        OC_CUDA(cudaFuncSetAttribute(k_score_topk<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(smem)));
        attr_set[sizeof(T) == 8] = true;
    }
    const unsigned blocks = unsigned((uint64_t(row1 - row0) + TU - 1) / TU);
    OC_LAUNCH((k_score_topk<T>), blocks, kThreads, smem, s, Pva, Qva, Kc, bt, row0, row1, n_ranked,
              cold, ids);
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
                                uint32_t, const uint8_t *, uint32_t *, cudaStream_t);              \
    template void vector_topk<T>(const T *, uint32_t, uint32_t *, cudaStream_t);                   \
    template void eval_metrics<T>(const uint32_t *, const uint32_t *, const uint8_t *,             \
                                  const uint32_t *, const uint32_t *, uint32_t, uint32_t,          \
                                  const T *, const T *, uint32_t, const T *, const T *, const T *, \
                                  uint32_t, uint32_t, double *, cudaStream_t);

OC_INSTANTIATE(float)
OC_INSTANTIATE(double)

}  // namespace ocffm
