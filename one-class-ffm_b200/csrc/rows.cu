// rows.cu -- row-parallel sparse kernels of the solver (north_star subsystems 1 and 3).
//
// Mapping shared by every kernel here: the latent dimension is padded to kp = 4*G (G a power of
// two <= 32) and a *lane group* of G consecutive lanes owns one row (or one bounded chunk of a
// row of Omega); each lane keeps 4 consecutive latent components in registers, so one group
// reads or writes a k-wide embedding row with one fully coalesced 16-byte (fp32) access per lane
// and a warp works on 32/G rows at once.  Dots over k are group shuffles; X^T(.) scatters are
// 16-byte REDs (red.global.add.v4.f32 on sm_100a) so no thread-private D x k scratch exists
// (the reference's G_/Hv_ buffers, ffm.cpp:555-557, 672-674, 759).
#include <map>

#include "common.cuh"
#include "kernels.h"

namespace ocffm {

namespace {

constexpr int kThreads = 256;
#ifndef OC_GATHER_MINB
#define OC_GATHER_MINB 4   // <= 64 registers: 4 CTAs (32 warps) per SM; the uncapped build drifted to 82
#endif

template <int G>
__device__ __forceinline__ uint32_t group_mask() {
    if constexpr (G == 32) {
        return 0xffffffffu;
    } else {
        const uint32_t lane = threadIdx.x & 31u;
        return ((1u << G) - 1u) << (lane & ~uint32_t(G - 1));
    }
}

template <int G, typename T>
__device__ __forceinline__ T gsum(T v, uint32_t mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

// destination row of a scatter contribution to feature `idx` (hot features: a per-warp replica)
template <typename T>
__device__ __forceinline__ T *scatter_row(const CsrView<T> &X, T *Out, uint32_t idx, uint32_t kp) {
    if (X.hot_slot) {
        const int hs = X.hot_slot[idx];
        if (hs >= 0) {
            const uint32_t rep = (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) & (kHotReplicas - 1);
            return X.shadow + (size_t(hs) * kHotReplicas + rep) * kp;
        }
    }
    return Out + size_t(idx) * kp;
}

// L2-only loads (ld.global.cg) for data written earlier in the same kernel by other CTAs
template <typename T>
__device__ __forceinline__ V4<T> ldcg4(const T *p);
template <>
__device__ __forceinline__ V4<float> ldcg4<float>(const float *p) {
    const float4 v = __ldcg(reinterpret_cast<const float4 *>(p));
    return {v.x, v.y, v.z, v.w};
}
template <>
__device__ __forceinline__ V4<double> ldcg4<double>(const double *p) {
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    return {a.x, a.y, b.x, b.y};
}

template <typename T>
__device__ __forceinline__ V4<T> scale4(const V4<T> &v, T s) {
    return {v.x * s, v.y * s, v.z * s, v.w * s};
}

// ---------------------------------------------------------------------------------------------
template <typename T, int G>
__global__ void __launch_bounds__(kThreads)
k_spmm_rows(CsrView<T> X, const T *__restrict__ A, T *__restrict__ C, uint32_t ldc) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    const uint64_t g = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    const uint32_t lg = threadIdx.x % G;
    if (g >= uint64_t(X.row1 - X.row0)) return;
    const uint32_t row = X.row0 + uint32_t(g);
    V4<T> acc = zero4<T>();
    const uint32_t e = X.rowptr[row + 1];
    for (uint32_t t = X.rowptr[row]; t < e; ++t)
        fma4(acc, X.val[t], ldg4(A + size_t(X.idx[t]) * kp + lg * 4));
    st4(C + size_t(row) * ldc + lg * 4, acc);
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads)
k_spmm_update(CsrView<T> X, const T *__restrict__ S, T *__restrict__ XS, T *__restrict__ P1,
              uint32_t ldp, const T *__restrict__ Q1side, T *__restrict__ gap, T *__restrict__ a1) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    const uint64_t g = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    const uint32_t lg = threadIdx.x % G;
    if (g >= uint64_t(X.row1 - X.row0)) return;
    const uint32_t mask = group_mask<G>();
    const uint32_t row = X.row0 + uint32_t(g);
    V4<T> acc = zero4<T>();
    const uint32_t e = X.rowptr[row + 1];
    for (uint32_t t = X.rowptr[row]; t < e; ++t)
        fma4(acc, X.val[t], ldg4(S + size_t(X.idx[t]) * kp + lg * 4));
    st4(XS + size_t(row) * kp + lg * 4, acc);
    T *p = P1 + size_t(row) * ldp + lg * 4;
    V4<T> pv = ld4(p);
    pv.x += acc.x; pv.y += acc.y; pv.z += acc.z; pv.w += acc.w;
    st4(p, pv);
    if (Q1side) {
        const T d = gsum<G>(dot4(acc, ldg4(Q1side + size_t(row) * kp + lg * 4)), mask);
        if (lg == 0) {
            gap[row] = d;
            a1[row] += d;
        }
    }
}

// G == 8: the partial dot products d[0..7] a lane holds for 8 gathered rows -> lane l returns the
// full sum of d[l] (butterfly reduce-scatter: 4 + 2 + 1 = 7 shuffles instead of 8 x 3)
template <typename T>
__device__ __forceinline__ T reduce_scatter8(const T (&d)[8], uint32_t lg, uint32_t mask) {
    const bool b2 = (lg & 4u) != 0, b1 = (lg & 2u) != 0, b0 = (lg & 1u) != 0;
    T e[4], f[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T send = b2 ? d[i] : d[i + 4], keep = b2 ? d[i + 4] : d[i];
        e[i] = keep + __shfl_xor_sync(mask, send, 4, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T send = b1 ? e[i] : e[i + 2], keep = b1 ? e[i + 2] : e[i];
        f[i] = keep + __shfl_xor_sync(mask, send, 2, 8);
    }
    const T send = b0 ? f[0] : f[1], keep = b0 ? f[1] : f[0];
    return keep + __shfl_xor_sync(mask, send, 1, 8);
}

// ---------------------------------------------------------------------------------------------
template <typename T, int G>
__global__ void __launch_bounds__(kThreads, OC_GATHER_MINB)
k_grad_cross(OmegaView<T> Y, CsrView<T> X, const T *__restrict__ Q1, uint32_t ldq,
             const T *__restrict__ Tm, const T *__restrict__ a1, const T *__restrict__ oQ,
             const T *__restrict__ bQ, T w, T r, T *__restrict__ Gout) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    const uint64_t item = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (item >= Y.n_items) return;
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    const uint32_t row = Y.wi_row[item], beg = Y.wi_beg[item], c = Y.wi_cnt[item];
    const uint32_t cnt = c & 0x7fffffffu;
    const bool first = (c >> 31) != 0;
    const T omw = T(1) - w, cst = w * (T(1) - r);
    constexpr int U = G < 8 ? G : 8;
    V4<T> pk = zero4<T>();
    const T *qbase = Q1 + lg * 4;
    for (uint32_t base = 0; base < cnt; base += G) {
        const bool ok = base + lg < cnt;
        const uint32_t t = beg + base + lg;
        const uint32_t j = ok ? Y.idx[t] : 0u;
        const T sc = ok ? omw * Y.yt[t] - cst : T(0);
        // Short rows take the same U-wide batches as full ones: entries past cnt gather row 0 with
        // a zero coefficient, so every gather of a batch is in flight at once (a one-at-a-time
        // tail would pay one L2 round trip per entry, and most item rows are all tail).
#pragma unroll
        for (int l0 = 0; l0 < G; l0 += U) {
            if (base + l0 >= cnt) break;
            V4<T> q[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                q[u] = ldg4(qbase + size_t(__shfl_sync(mask, j, l0 + u, G)) * ldq);
#pragma unroll
            for (int u = 0; u < U; ++u) fma4(pk, __shfl_sync(mask, sc, l0 + u, G), q[u]);
        }
    }
    if (first) {
        const T zi = a1[row] - r;
        const V4<T> t1 = ldg4(Tm + size_t(row) * kp + lg * 4);
        const V4<T> o = ldg4(oQ + lg * 4), b = ldg4(bQ + lg * 4);
        pk.x += w * (t1.x + zi * o.x + b.x);
        pk.y += w * (t1.y + zi * o.y + b.y);
        pk.z += w * (t1.z + zi * o.z + b.z);
        pk.w += w * (t1.w + zi * o.w + b.w);
    }
    const uint32_t xb = X.diagonal ? row : X.rowptr[row], xe = X.diagonal ? row + 1 : X.rowptr[row + 1];
    for (uint32_t t = xb; t < xe; ++t)
        red4(scatter_row(X, Gout, X.identity ? row : X.idx[t], kp) + lg * 4, scale4(pk, X.val[t]));
}

// One work item of the hs_cross row pass (ffm.cpp:715-738): phi = X_i V, tau = X_i (V QTQ) on the
// row's first item, ka = sum_j (phi . q_j) q_j over the item's pairs, Hv += X_i^T ((1-w) ka + w tau);
// returns this lane's share of V . (X^T z) = phi . z.  COH: V / VQ were written earlier in the SAME
// kernel by other CTAs (persistent CG) and must be read past the non-coherent L1.
template <typename T, int G, bool COH>
__device__ __forceinline__ T hess_cross_item(const OmegaView<T> &Y, const CsrView<T> &X, const T *__restrict__ Q1,
                                             uint32_t ldq, const T *V, const T *VQ, T w, T *Hv, uint32_t item,
                                             int notau) {
    constexpr uint32_t kp = 4 * G;
    constexpr int U = G < 8 ? G : 8;   // gathers kept in flight per lane
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    const uint32_t row = Y.wi_row[item], beg = Y.wi_beg[item], c = Y.wi_cnt[item];
    const uint32_t cnt = c & 0x7fffffffu;
    const bool first = (c >> 31) != 0 && !notau;   // notau: w * tau is already in Hv
    const uint32_t xb = X.diagonal ? row : X.rowptr[row], xe = X.diagonal ? row + 1 : X.rowptr[row + 1];
    uint32_t j = lg < cnt ? Y.idx[beg + lg] : 0u;
    V4<T> phi = zero4<T>(), tau = zero4<T>();
    for (uint32_t t = xb; t < xe; ++t) {
        const size_t off = size_t(X.identity ? row : X.idx[t]) * kp + lg * 4;
        const T v = X.val[t];
        fma4(phi, v, COH ? ldcg4(V + off) : ldg4(V + off));
        if (first) fma4(tau, v, COH ? ldcg4(VQ + off) : ldg4(VQ + off));
    }
    V4<T> ka = zero4<T>();
    const T *qbase = Q1 + lg * 4;
    for (uint32_t base = 0; base < cnt; base += G) {
        const uint32_t nb = base + G + lg;
        const uint32_t jn = nb < cnt ? Y.idx[beg + nb] : 0u;   // next batch of column ids
#pragma unroll
        for (int l0 = 0; l0 < G; l0 += U) {
            if (base + l0 >= cnt) break;   // entries past cnt: row 0, zero coefficient (see k_grad_cross)
            V4<T> q[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                q[u] = ldg4(qbase + size_t(__shfl_sync(mask, j, l0 + u, G)) * ldq);
            if constexpr (G == 8 && sizeof(T) == 4) {   // fp64 would spill: it keeps the plain sums
                T d[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) d[u] = dot4(phi, q[u]);
                T mine = reduce_scatter8(d, lg, mask);   // phi . q of gathered row lg
                if (base + lg >= cnt) mine = T(0);
#pragma unroll
                for (int u = 0; u < 8; ++u) fma4(ka, __shfl_sync(mask, mine, u, 8), q[u]);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    T sdot = gsum<G>(dot4(phi, q[u]), mask);
                    if (base + l0 + u >= cnt) sdot = T(0);
                    fma4(ka, sdot, q[u]);
                }
            }
        }
        j = jn;
    }
    const T omw = T(1) - w;
    V4<T> z = {omw * ka.x + w * tau.x, omw * ka.y + w * tau.y, omw * ka.z + w * tau.z,
               omw * ka.w + w * tau.w};
    for (uint32_t t = xb; t < xe; ++t)
        red4(scatter_row(X, Hv, X.identity ? row : X.idx[t], kp) + lg * 4, scale4(z, X.val[t]));
    return dot4(phi, z);
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads, OC_GATHER_MINB)
k_hess_cross(OmegaView<T> Y, CsrView<T> X, const T *__restrict__ Q1, uint32_t ldq,
             const T *__restrict__ V, const T *__restrict__ VQ, T w, T *__restrict__ Hv, Gate gate,
             double *__restrict__ dot_out, int notau) {
    pdl_enter();
    if (!gate_open(gate)) return;
    const uint64_t item = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    T vhv = T(0);   // this lane's share of V . (X^T z) = (X_i V) . z
    if (item < Y.n_items) vhv = hess_cross_item<T, G, false>(Y, X, Q1, ldq, V, VQ, w, Hv, uint32_t(item), notau);
    if (dot_out) warp_add_slot(double(vhv), dot_out, kDotSlots);
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads, OC_GATHER_MINB)
k_sddmm_add(OmegaView<T> Y, const T *__restrict__ Uown, uint32_t ldu, const T *__restrict__ Vo,
            uint32_t ldv) {
    pdl_enter();
    const uint64_t item = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (item >= Y.n_items) return;
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    const uint32_t row = Y.wi_row[item], beg = Y.wi_beg[item];
    const uint32_t cnt = Y.wi_cnt[item] & 0x7fffffffu;
    if (cnt == 0) return;
    constexpr int U = G < 8 ? G : 8;
    const V4<T> u = ldg4(Uown + size_t(row) * ldu + lg * 4);
    const T *vbase = Vo + lg * 4;
    for (uint32_t base = 0; base < cnt; base += G) {
        const bool ok = base + lg < cnt;
        const uint32_t t = beg + base + lg;
        const uint32_t j = ok ? Y.idx[t] : 0u;
        T mine = T(0);
#pragma unroll
        for (int l0 = 0; l0 < G; l0 += U) {
            if (base + l0 >= cnt) break;   // entries past cnt gather row 0 and are not stored
            V4<T> q[U];
#pragma unroll
            for (int x = 0; x < U; ++x)
                q[x] = ldg4(vbase + size_t(__shfl_sync(mask, j, l0 + x, G)) * ldv);
            if constexpr (G == 8) {
                T d[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) d[x] = dot4(u, q[x]);
                mine = reduce_scatter8(d, lg, mask);
            } else {
#pragma unroll
                for (int x = 0; x < U; ++x) {
                    const T s = gsum<G>(dot4(u, q[x]), mask);
                    if (int(lg) == l0 + x) mine = s;
                }
            }
        }
        if (ok) Y.yt[t] += mine;
    }
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads)
k_ytilde_rowsum(OmegaView<T> Y, T *__restrict__ ysum) {
    pdl_enter();
    const uint64_t item = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (item >= Y.n_items) return;
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    const uint32_t row = Y.wi_row[item], beg = Y.wi_beg[item];
    const uint32_t cnt = Y.wi_cnt[item] & 0x7fffffffu;
    T acc = T(0);
    for (uint32_t o = lg; o < cnt; o += G) acc += Y.yt[beg + o];
    acc = gsum<G>(acc, mask);
    if (lg == 0 && cnt) atomicAdd(ysum + row, acc);
}

template <typename T, int G, int MODE>
__global__ void __launch_bounds__(kThreads)
k_side_rows(OmegaView<T> Y, CsrView<T> X, const T *__restrict__ Q1, const T *__restrict__ a1,
            const T *__restrict__ sa1, const T *__restrict__ ysum, const double *__restrict__ bsum,
            const T *__restrict__ V, T w, T r, T n1, T *__restrict__ Out, Gate gate,
            double *__restrict__ dot_out) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    if (!gate_open(gate)) return;
    const uint64_t g = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    T vhv = T(0);
    if (g < uint64_t(X.row1 - X.row0)) {
        const uint32_t lg = threadIdx.x % G;
        const uint32_t mask = group_mask<G>();
        const uint32_t row = X.row0 + uint32_t(g);
        const T cnt = T(Y.rowptr[row + 1] - Y.rowptr[row]);
        const V4<T> q = ldg4(Q1 + size_t(row) * kp + lg * 4);
        const uint32_t xb = X.rowptr[row], xe = X.rowptr[row + 1];
        T z;
        if (MODE == 0) {
            // z_i = w (n1 (a_i - r) + sum(b) + sa_i) + sum_Omega_i ((1-w) ytilde - w (1-r))
            z = w * (n1 * (a1[row] - r) + T(*bsum) + sa1[row]) + (T(1) - w) * ysum[row] -
                w * (T(1) - r) * cnt;
        } else {
            // d_i (q_i . X_i V),  d_i = (1-w) |Omega_i| + w n1
            T acc = T(0);
            for (uint32_t t = xb; t < xe; ++t)
                acc += X.val[t] * dot4(q, ldg4(V + size_t(X.idx[t]) * kp + lg * 4));
            const T sdot = gsum<G>(acc, mask);
            z = sdot * ((T(1) - w) * cnt + w * n1);
            if (lg == 0) vhv = sdot * z;   // V . X^T (q z) = (q . X_i V) z, once per row
        }
        for (uint32_t t = xb; t < xe; ++t)
            red4(scatter_row(X, Out, X.idx[t], kp) + lg * 4, scale4(q, X.val[t] * z));
    }
    if (MODE == 1 && dot_out) warp_add_slot(double(vhv), dot_out, kDotSlots);
}

// block-wide fp64 sum + deterministic last-block combine (same scheme as dense.cu's finish_sum)
__device__ __forceinline__ void finish_sum_rows(double local, SolveScalars *sc, double *out) {
    __shared__ bool is_last;
    local = block_sum(local);
    if (threadIdx.x == 0) {
        sc->partials[blockIdx.x] = local;
        __threadfence();
        const unsigned t = atomicInc(&sc->counter[0], gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double sum = 0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x)
            sum += reinterpret_cast<volatile double *>(sc->partials)[i];
        sum = block_sum(sum);
        if (threadIdx.x == 0) *out = sum;
    }
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads)
k_side_diag_iter(OmegaView<T> Y, CsrView<T> X, const T *__restrict__ Q1, T *__restrict__ V,
                 const T *__restrict__ R, T *__restrict__ Hv, const T *__restrict__ freq, T lambda, T w, T n1,
                 int it, SolveScalars *sc) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    if (!gate_open(Gate{sc, it})) return;
    const T beta = it > 0 ? T(sc->r2[it] / sc->r2[it - 1]) : T(0);
    const uint32_t rows = X.row1 - X.row0;
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    double local = 0;
    const uint64_t groups_total = uint64_t(gridDim.x) * blockDim.x / G;
    for (uint64_t g = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G; g < rows; g += groups_total) {
        const uint32_t row = X.row0 + uint32_t(g);
        const uint32_t f = X.idx[row];          // rowptr[row] == row on a diagonal field
        const T x = X.val[row];
        const size_t off = size_t(f) * kp + lg * 4;
        V4<T> v = ld4(V + off);
        if (it > 0) {
            const V4<T> r = ld4(R + off);
            v.x = r.x + beta * v.x; v.y = r.y + beta * v.y; v.z = r.z + beta * v.z; v.w = r.w + beta * v.w;
            st4(V + off, v);
        }
        const V4<T> q = ldg4(Q1 + size_t(row) * kp + lg * 4);
        const T cnt = T(Y.rowptr[row + 1] - Y.rowptr[row]);
        const T z = gsum<G>(dot4(q, v), mask) * x * ((T(1) - w) * cnt + w * n1) * x;
        const T c = freq ? lambda * freq[f] : lambda;
        const V4<T> h = {c * v.x + q.x * z, c * v.y + q.y * z, c * v.z + q.z * z, c * v.w + q.w * z};
        st4(Hv + off, h);
        local += double(v.x) * h.x + double(v.y) * h.y + double(v.z) * h.z + double(v.w) * h.w;
    }
    finish_sum_rows(local, sc, &sc->vHv[it]);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_ytilde_base(OmegaView<T> Y, const T *__restrict__ a_own, const T *__restrict__ b_oth) {
    pdl_enter();
    constexpr int G = 8;
    const uint64_t item = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (item >= Y.n_items) return;
    const uint32_t lg = threadIdx.x % G;
    const uint32_t row = Y.wi_row[item], beg = Y.wi_beg[item];
    const uint32_t cnt = Y.wi_cnt[item] & 0x7fffffffu;
    const T base = a_own[row] - T(1);
    for (uint32_t o = lg; o < cnt; o += G) Y.yt[beg + o] = base + b_oth[Y.idx[beg + o]];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_ytilde_gap_by_row(OmegaView<T> Y, const T *__restrict__ gap) {
    pdl_enter();
    constexpr int G = 8;
    const uint64_t item = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (item >= Y.n_items) return;
    const uint32_t lg = threadIdx.x % G;
    const uint32_t row = Y.wi_row[item], beg = Y.wi_beg[item];
    const uint32_t cnt = Y.wi_cnt[item] & 0x7fffffffu;
    const T g = gap[row];
    for (uint32_t o = lg; o < cnt; o += G) Y.yt[beg + o] += g;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_ytilde_gap_by_idx(OmegaView<T> Y, const T *__restrict__ gap) {
    pdl_enter();
    const uint64_t b = Y.rowptr[Y.row0], e = Y.rowptr[Y.row1];
    for (uint64_t t = b + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < e;
         t += uint64_t(gridDim.x) * blockDim.x)
        Y.yt[t] += gap[Y.idx[t]];
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads)
k_rowwise_dot(const T *__restrict__ P, const T *__restrict__ Q, uint32_t rows, T *__restrict__ out,
              int accumulate) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    const uint64_t g = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (g >= rows) return;
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    const T d = gsum<G>(dot4(ldg4(P + g * kp + lg * 4), ldg4(Q + g * kp + lg * 4)), mask);
    if (lg == 0) out[g] = accumulate ? out[g] + d : d;
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads)
k_fold_hot(const T *__restrict__ shadow, const uint32_t *__restrict__ hot_feat, uint32_t n_hot,
           T *__restrict__ Out) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    const uint64_t g = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (g >= n_hot) return;
    const uint32_t lg = threadIdx.x % G;
    V4<T> acc = zero4<T>();
    const T *base = shadow + size_t(g) * kHotReplicas * kp + lg * 4;
#pragma unroll 8
    for (int r = 0; r < kHotReplicas; ++r) {
        const V4<T> v = ld4(base + size_t(r) * kp);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    T *dst = Out + size_t(hot_feat[g]) * kp + lg * 4;
    V4<T> o = ld4(dst);
    o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
    st4(dst, o);
}


// ---------------------------------------------------------------------------------------------
// Per-row observed Gram ("Mrow").  The Omega term of hs_cross (ffm.cpp:722-736) for row i is
//   sum_{j in Omega_i} (phi_i . q_j) q_j = M_i phi_i,   M_i = sum_{j in Omega_i} q_j q_j^T  (k x k),
// and Q1 does not change during a half solve (ffm.cpp:744-813 works on one block), so for rows
// with many observed pairs M_i is built ONCE per half solve (k_row_gram) and every CG iteration
// streams kp*kp numbers per row from HBM (k_hess_heavy) instead of gathering |Omega_i| rows of
// Q1 through L2/L1 -- the gather pass was bound by L1TEX wavefronts and load latency, this one
// by HBM bandwidth.  Rows with few pairs stay on the gather path (k_hess_cross on the light list).
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
__device__ __forceinline__ void load_n(const T *p, T (&o)[N]) {
    if constexpr (sizeof(T) == 4 && N == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        o[0] = v.x; o[1] = v.y;
    } else if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 4) {
            const V4<T> v = ldg4(p + i);
            o[i] = v.x; o[i + 1] = v.y; o[i + 2] = v.z; o[i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = __ldg(p + i);
    }
}

// acc[r][c] += a[r] * b[c]; fp32 uses the packed FFMA2 of sm_100 (two FMAs per issue slot)
template <typename T, int TR, int TC>
__device__ __forceinline__ void outer_acc(T (&acc)[TR][TC], const T (&a)[TR], const T (&b)[TC]) {
    if constexpr (sizeof(T) == 4 && TC % 2 == 0) {
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            const float2 ar = make_float2(a[r], a[r]);
#pragma unroll
            for (int c = 0; c < TC; c += 2) {
                const float2 v = __ffma2_rn(ar, make_float2(b[c], b[c + 1]), make_float2(acc[r][c], acc[r][c + 1]));
                acc[r][c] = v.x;
                acc[r][c + 1] = v.y;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int c = 0; c < TC; ++c) acc[r][c] += a[r] * b[c];
    }
}

// One warp per item: item w covers nnz [beg[w], beg[w] + (cnt[w] & 0x7fffffff)) of the heavy row in
// slot[w]; bit 31 of cnt: the row is split over several items, accumulate with REDs into the
// pre-zeroed M[slot] (otherwise the item is the whole row and M[slot] is simply stored).
// Lane (a, b) = (lane / 4, lane % 4) owns the TR x TC tile at (a*TR, b*TC) of the kp x kp matrix.
template <typename T, int KP>
__global__ void __launch_bounds__(kThreads)
k_row_gram(const uint32_t *__restrict__ it_slot, const uint32_t *__restrict__ it_beg,
           const uint32_t *__restrict__ it_cnt, uint32_t n_items, const uint32_t *__restrict__ yidx,
           const T *__restrict__ Q1, uint32_t ldq, T *__restrict__ M) {
    pdl_enter();
    constexpr int TR = KP / 8, TC = KP / 4, U = 4;
    const uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (item >= n_items) return;
    const uint32_t slot = it_slot[item], beg = it_beg[item], c = it_cnt[item];
    const uint32_t cnt = c & 0x7fffffffu;
    const uint32_t a = lane >> 2, b = lane & 3u;
    T acc[TR][TC];
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) acc[r][cc] = T(0);
    const T *qa_base = Q1 + a * TR, *qb_base = Q1 + b * TC;
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t nb = min(32u, cnt - base);
        const uint32_t jl = lane < nb ? yidx[beg + base + lane] : 0u;
        for (uint32_t u0 = 0; u0 < nb; u0 += U) {
            T qa[U][TR], qb[U][TC];
#pragma unroll
            for (int u = 0; u < U; ++u) {   // entries past nb re-read row yidx -> 0 and contribute nothing
                const size_t off = size_t(__shfl_sync(0xffffffffu, jl, (u0 + u) & 31u)) * ldq;
                load_n<T, TR>(qa_base + off, qa[u]);
                load_n<T, TC>(qb_base + off, qb[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u0 + u >= nb) {
#pragma unroll
                    for (int r = 0; r < TR; ++r) qa[u][r] = T(0);
                }
                outer_acc<T, TR, TC>(acc, qa[u], qb[u]);
            }
        }
    }
    T *out = M + size_t(slot) * KP * KP + size_t(a * TR) * KP + b * TC;
    if (c >> 31) {
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int cc = 0; cc < TC; cc += 4)
                red4(out + r * KP + cc, V4<T>{acc[r][cc], acc[r][cc + 1], acc[r][cc + 2], acc[r][cc + 3]});
    } else {
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int cc = 0; cc < TC; cc += 4)
                st4(out + r * KP + cc, V4<T>{acc[r][cc], acc[r][cc + 1], acc[r][cc + 2], acc[r][cc + 3]});
    }
}

// hs_cross for the heavy rows (one warp per row): z_i = (1-w) M_i phi_i + w tau_i, Hv += X_i^T z_i.
// The kp x kp matrix is read with fully coalesced 16-byte loads: a group of G = kp/4 lanes takes one
// matrix row per step, the 32/G groups of the warp take rows g*STEPS + step; after the group
// reduction lane (g, lg < STEPS) owns component g*STEPS + lg of z.
// one heavy row (one warp): returns this lane's share of V . (X^T z)
template <typename T, int KP, bool COH>
__device__ __forceinline__ T hess_heavy_row(const uint32_t *__restrict__ heavy_rows, const CsrView<T> &X,
                                            const T *__restrict__ M, const T *V, const T *VQ, T w, T *Hv,
                                            uint32_t slot, int notau) {
    constexpr int G = KP / 4, NG = 32 / G, STEPS = KP / NG;
    static_assert(STEPS <= G, "every component of z needs an owner lane");
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t row = heavy_rows[slot];
    const uint32_t g = lane / G, lg = lane % G;
    const bool owner = lg < uint32_t(STEPS);
    const uint32_t zi = g * STEPS + (owner ? lg : 0u);
    const uint32_t xb = X.diagonal ? row : X.rowptr[row], xe = X.diagonal ? row + 1 : X.rowptr[row + 1];
    V4<T> phi4 = zero4<T>();
    T phi_o = T(0), tau_o = T(0);
    for (uint32_t t = xb; t < xe; ++t) {
        const size_t off = size_t(X.identity ? row : X.idx[t]) * KP;
        const T v = X.val[t];
        fma4(phi4, v, COH ? ldcg4(V + off + lg * 4) : ldg4(V + off + lg * 4));
        phi_o += v * (COH ? __ldcg(V + off + zi) : __ldg(V + off + zi));
        if (!notau) tau_o += v * (COH ? __ldcg(VQ + off + zi) : __ldg(VQ + off + zi));
    }
    const T *Mi = M + size_t(slot) * KP * KP + size_t(g * STEPS) * KP + lg * 4;
    T d[STEPS];
#pragma unroll
    for (int cstep = 0; cstep < STEPS; ++cstep) d[cstep] = dot4(ldg4(Mi + cstep * KP), phi4);
    T mine = T(0);
#pragma unroll
    for (int cstep = 0; cstep < STEPS; ++cstep) {
        const T sum = gsum<G>(d[cstep], 0xffffffffu);
        if (lg == uint32_t(cstep)) mine = sum;
    }
    T vhv = T(0);
    if (owner) {
        const T z = (T(1) - w) * mine + w * tau_o;
        vhv = phi_o * z;
        for (uint32_t t = xb; t < xe; ++t)
            atomicAdd(scatter_row(X, Hv, X.identity ? row : X.idx[t], KP) + zi, z * X.val[t]);
    }
    return vhv;
}

template <typename T, int KP>
__global__ void __launch_bounds__(kThreads)
k_hess_heavy(const uint32_t *__restrict__ heavy_rows, uint32_t n_heavy, CsrView<T> X,
             const T *__restrict__ M, const T *__restrict__ V, const T *__restrict__ VQ, T w,
             T *__restrict__ Hv, Gate gate, double *__restrict__ dot_out, int notau) {
    pdl_enter();
    if (!gate_open(gate)) return;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    T vhv = T(0);
    if (slot < n_heavy) vhv = hess_heavy_row<T, KP, false>(heavy_rows, X, M, V, VQ, w, Hv, slot, notau);
    if (dot_out) warp_add_slot(double(vhv), dot_out, kDotSlots);
}

// ---------------------------------------------------------------------------------------------
// A whole CG solve of a same-side half (cg, ffm.cpp:744-813 with hs_side, ffm.cpp:594-628) in ONE
// cooperative launch.  The per-iteration kernels of a same-side half are tiny (a few microseconds of
// work on L2-resident vectors), so the 3-4 launches per iteration plus the host's event wait
// dominated; here the CTAs stay resident, iterate with grid-wide barriers and evaluate the stop test
// (g2 * 0.09 < r2, at most 20 iterations) themselves -- the host enqueues one kernel per half solve
// and never waits inside an outer iteration.
//   general field : [V = R + beta V, Hv = 0, reg share of V.Hv] | [rows: Hv += X^T q d (q . X V), data
//                   share of V.Hv] | [alpha; S += alpha V; R -= alpha (Hv + lambda c V); ||R||^2]
//   diagonal field: [rows: direction, Hv (regulariser included), V.Hv -- row-local, no atomics] |
//                   [alpha; S += alpha V; R -= alpha Hv; ||R||^2]
// Every grid-wide sum is "per-CTA partial -> barrier -> every CTA adds the partials in the same
// order", so all CTAs take bit-identical stop decisions.  Vectors written inside the kernel are read
// back with ld.global.cg (L1 is not coherent across SMs).
// ---------------------------------------------------------------------------------------------
constexpr int kPersistMaxBlocks = 592;   // two partial arrays inside SolveScalars::partials

__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned target = epoch * gridDim.x;
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}
// sum of the per-CTA partials, same order in every CTA (valid in all threads)
__device__ __forceinline__ double sum_partials(const double *part) {
    __shared__ double bc;
    double v = 0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(part + i);
    v = block_sum(v);
    if (threadIdx.x == 0) bc = v;
    __syncthreads();
    v = bc;
    __syncthreads();
    return v;
}

// ---- multi-rank: one fp64 scalar summed over the ranks inside the kernel (thread 0 of CTA 0) ----------
__device__ __forceinline__ double peerk_sum(const PeerK &pk, double v, unsigned seq) {
    const int slot = int(seq & 1u);
    const unsigned long long bits = __double_as_longlong(v);
    const uint32_t lo = uint32_t(bits), hi = uint32_t(bits >> 32);
    for (int q = 0; q < pk.nranks; ++q)
        if (q != pk.rank)
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pk.area[q] + slot * pk.nranks + pk.rank),
                         "r"(lo), "r"(seq), "r"(hi), "r"(seq)
                         : "memory");
    double acc = 0;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int q = 0; q < pk.nranks; ++q) {
        if (q == pk.rank) { acc += v; continue; }
        const uint4 *src = pk.area[pk.rank] + slot * pk.nranks + q;
        uint4 x;
        for (uint32_t spin = 1;; ++spin) {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(src) : "memory");
            if (x.y == seq && x.w == seq) break;
            if ((spin & 0x3ffu) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 30ull * 1000 * 1000 * 1000) {   // a peer died: report, end the solve (NaN fails the stop test)
                    atomicExch(pk.error, 1);
                    return __longlong_as_double(0x7ff8000000000000ll);
                }
            }
        }
        acc += __longlong_as_double((unsigned long long)(x.z) << 32 | x.x);
    }
    return acc;
}
// grid-wide (and, on multi-rank contexts, job-wide) sum of the per-CTA partials; same value in every
// thread of every CTA of every rank.  One rank: every CTA adds the partials itself.  Several ranks:
// CTA 0 adds them, exchanges the sum with the peers and publishes the total behind one more barrier.
__device__ __forceinline__ double global_sum(const double *part, const PeerK &pk, unsigned &kseq, SolveScalars *sc,
                                             unsigned *counter, unsigned &epoch) {
    if (pk.nranks <= 1) return sum_partials(part);
    ++kseq;
    if (blockIdx.x == 0) {
        double v = sum_partials(part);
        if (threadIdx.x == 0) {
            v = peerk_sum(pk, v, kseq);
            sc->misc[kseq & 3u] = v;
        }
    }
    grid_barrier(counter, epoch);
    return __ldcg(&sc->misc[kseq & 3u]);
}

template <typename T, int G, bool DIAG>
__global__ void __launch_bounds__(kThreads)
k_cg_side_persist(OmegaView<T> Y, CsrView<T> X, const T *__restrict__ Q1, T *__restrict__ V, T *__restrict__ R,
                  T *__restrict__ S, T *__restrict__ Hv, const T *__restrict__ freq, T lambda, T w, T n1,
                  uint64_t D, SolveScalars *sc, int max_cg, double eps, unsigned *host_iters, uint64_t f0, PeerK pk) {
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    double *part_a = sc->partials, *part_b = sc->partials + kPersistMaxBlocks;
    unsigned *counter = &sc->counter[1];
    unsigned epoch = 0;
    unsigned kseq = pk.nranks > 1 ? *pk.seq : 0u;   // written only at the end of a kernel: every thread reads the same value
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x, nthreads = uint64_t(gridDim.x) * blockDim.x;
    const uint64_t nvec = D * (kp / 4);
    const uint32_t rows = X.row1 - X.row0;
    // the vector passes run on this rank's slice [f0, f0 + D) (f0 = 0 on one rank); row passes index globally
    T *Vs = V + f0 * kp, *Rs = R + f0 * kp, *Ss = S + f0 * kp, *Hvs = Hv + f0 * kp;
    const T *freqs = freq ? freq + f0 : nullptr;
    const double g2 = sc->r2[0];
    double r2 = g2, r2_prev = 0;   // carried in registers: L1 is not coherent with block 0's stores to sc->r2[]
    int it = 0;
    while (g2 * eps < r2 && it < max_cg) {
        const T beta = it > 0 ? T(r2 / r2_prev) : T(0);
        double local = 0;
        if (!DIAG) {
            for (uint64_t i = tid; i < nvec; i += nthreads) {
                V4<T> v = ldcg4(V + i * 4);
                if (it > 0) {
                    const V4<T> r = ldcg4(R + i * 4);
                    v.x = r.x + beta * v.x; v.y = r.y + beta * v.y; v.z = r.z + beta * v.z; v.w = r.w + beta * v.w;
                    st4(V + i * 4, v);
                }
                st4(Hv + i * 4, zero4<T>());
                const T c = freq ? lambda * freq[i / (kp / 4)] : lambda;
                local += double(c) * (double(v.x) * v.x + double(v.y) * v.y + double(v.z) * v.z + double(v.w) * v.w);
            }
            grid_barrier(counter, epoch);
            for (uint64_t g = tid / G; g < rows; g += nthreads / G) {
                const uint32_t row = X.row0 + uint32_t(g);
                const T cnt = T(Y.rowptr[row + 1] - Y.rowptr[row]);
                const V4<T> q = ldg4(Q1 + size_t(row) * kp + lg * 4);
                const uint32_t xb = X.rowptr[row], xe = X.rowptr[row + 1];
                T acc = T(0);
                for (uint32_t t = xb; t < xe; ++t)
                    acc += X.val[t] * dot4(q, ldcg4(V + size_t(X.idx[t]) * kp + lg * 4));
                const T sdot = gsum<G>(acc, mask);
                const T z = sdot * ((T(1) - w) * cnt + w * n1);
                if (lg == 0) local += double(sdot) * double(z);
                for (uint32_t t = xb; t < xe; ++t)
                    red4(Hv + size_t(X.idx[t]) * kp + lg * 4, scale4(q, X.val[t] * z));
            }
        } else {
            for (uint64_t g = tid / G; g < rows; g += nthreads / G) {
                const uint32_t row = X.row0 + uint32_t(g);
                const uint32_t f = X.idx[row];
                const T x = X.val[row];
                const size_t off = size_t(f) * kp + lg * 4;
                V4<T> v = ldcg4(V + off);
                if (it > 0) {
                    const V4<T> r = ldcg4(R + off);
                    v.x = r.x + beta * v.x; v.y = r.y + beta * v.y; v.z = r.z + beta * v.z; v.w = r.w + beta * v.w;
                    st4(V + off, v);
                }
                const V4<T> q = ldg4(Q1 + size_t(row) * kp + lg * 4);
                const T cnt = T(Y.rowptr[row + 1] - Y.rowptr[row]);
                const T z = gsum<G>(dot4(q, v), mask) * x * ((T(1) - w) * cnt + w * n1) * x;
                const T c = freq ? lambda * freq[f] : lambda;
                const V4<T> h = {c * v.x + q.x * z, c * v.y + q.y * z, c * v.z + q.z * z, c * v.w + q.w * z};
                st4(Hv + off, h);
                local += double(v.x) * h.x + double(v.y) * h.y + double(v.z) * h.z + double(v.w) * h.w;
            }
        }
        local = block_sum(local);
        if (threadIdx.x == 0) part_a[blockIdx.x] = local;
        grid_barrier(counter, epoch);
        const double vhv = global_sum(part_a, pk, kseq, sc, counter, epoch);
        const T alpha = T(r2 / vhv);
        local = 0;
        for (uint64_t i = tid; i < nvec; i += nthreads) {
            const V4<T> v = ldcg4(Vs + i * 4);
            V4<T> h = ldcg4(Hvs + i * 4);
            if (!DIAG) {   // Hv holds the data term only: add lambda c_f V here (ffm.cpp:788-790)
                const T c = freqs ? lambda * freqs[i / (kp / 4)] : lambda;
                h.x += c * v.x; h.y += c * v.y; h.z += c * v.z; h.w += c * v.w;
            }
            V4<T> sv = ldcg4(Ss + i * 4), r = ldcg4(Rs + i * 4);
            sv.x += alpha * v.x; sv.y += alpha * v.y; sv.z += alpha * v.z; sv.w += alpha * v.w;
            r.x -= alpha * h.x; r.y -= alpha * h.y; r.z -= alpha * h.z; r.w -= alpha * h.w;
            st4(Ss + i * 4, sv);
            st4(Rs + i * 4, r);
            local += double(r.x) * r.x + double(r.y) * r.y + double(r.z) * r.z + double(r.w) * r.w;
        }
        local = block_sum(local);
        if (threadIdx.x == 0) part_b[blockIdx.x] = local;
        grid_barrier(counter, epoch);
        const double r2n = global_sum(part_b, pk, kseq, sc, counter, epoch);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            sc->vHv[it] = vhv;
            sc->r2[it + 1] = r2n;
        }
        r2_prev = r2;
        r2 = r2n;
        ++it;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->counter[2] = unsigned(it);
        if (host_iters) *host_iters = unsigned(it);   // mapped pinned host memory: no D2H memcpy needed
        if (pk.nranks > 1) *pk.seq = kseq;
    }
}


// The same for a CROSS half (cg with hs_cross, ffm.cpp:706-742, 744-813): per iteration
//   [per feature row: V = R + beta V, Hv = 0, VQ = V QTQ (QTQ in shared memory), reg share of V.Hv] |
//   [hs_cross work items: gathers of Q1 rows, Hv += X^T z, data share of V.Hv] | [step, ||R||^2]
// with three grid-wide barriers.  kp <= 64 (QTQ must fit the static shared memory).
template <typename T, int G>
__global__ void __launch_bounds__(kThreads, OC_GATHER_MINB)
k_cg_cross_persist(OmegaView<T> Y, CsrView<T> X, const T *__restrict__ Q1, uint32_t ldq, const T *__restrict__ QTQ,
                   T *__restrict__ V, T *__restrict__ R, T *__restrict__ S, T *__restrict__ Hv, T *__restrict__ VQ,
                   const T *__restrict__ freq, T lambda, T w, uint64_t D, SolveScalars *sc, int max_cg, double eps,
                   unsigned *host_iters, const uint32_t *__restrict__ heavy_rows, uint32_t n_heavy,
                   const T *__restrict__ Mrow, uint64_t f0, PeerK pk, float *host_phase_ms) {
    // heavy_rows / Mrow: the rows served by their per-row observed Gram blocks (Y is then the light list)
    pdl_enter();
    constexpr uint32_t kp = 4 * G;
    __shared__ __align__(16) T qtq[kp * kp];
    for (uint32_t i = threadIdx.x; i < kp * kp / 4; i += blockDim.x) st4(qtq + i * 4, ldg4(QTQ + i * 4));
    __syncthreads();
    double *part_a = sc->partials, *part_b = sc->partials + kPersistMaxBlocks;
    unsigned *counter = &sc->counter[1];
    unsigned epoch = 0;
    unsigned kseq = pk.nranks > 1 ? *pk.seq : 0u;
    unsigned long long t_b0 = 0, t_rows = 0;   // OCFFM_PROFILE: time of the hs_cross row phases (CTA 0, thread 0)
    const uint32_t lg = threadIdx.x % G;
    const uint32_t mask = group_mask<G>();
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x, nthreads = uint64_t(gridDim.x) * blockDim.x;
    const uint64_t nvec = D * (kp / 4);
    // this rank's feature slice [f0, f0 + D) for the vector passes (f0 = 0 on one rank)
    T *Vs = V + f0 * kp, *Rs = R + f0 * kp, *Ss = S + f0 * kp, *Hvs = Hv + f0 * kp;
    const T *freqs = freq ? freq + f0 : nullptr;
    const double g2 = sc->r2[0];
    double r2 = g2, r2_prev = 0;
    int it = 0;
    while (g2 * eps < r2 && it < max_cg) {
        const T beta = it > 0 ? T(r2 / r2_prev) : T(0);
        double local = 0;
        // ---- A: direction, Hv = 0, VQ = V QTQ; whole lane groups stay together (shuffles inside)
        for (uint64_t fl = tid / G; fl < ((D + (nthreads / G) - 1) / (nthreads / G)) * (nthreads / G); fl += nthreads / G) {
            const bool ok = fl < D;
            const uint64_t f = f0 + (ok ? fl : 0);
            const size_t off = size_t(f) * kp + lg * 4;
            V4<T> v = ldcg4(V + off);
            if (it > 0) {
                const V4<T> r = ldcg4(R + off);
                v.x = r.x + beta * v.x; v.y = r.y + beta * v.y; v.z = r.z + beta * v.z; v.w = r.w + beta * v.w;
            }
            V4<T> o = zero4<T>();
#pragma unroll
            for (int dl = 0; dl < G; ++dl) {
                const T v0 = __shfl_sync(mask, v.x, dl, G), v1 = __shfl_sync(mask, v.y, dl, G);
                const T v2 = __shfl_sync(mask, v.z, dl, G), v3 = __shfl_sync(mask, v.w, dl, G);
                fma4(o, v0, ld4(qtq + (dl * 4 + 0) * kp + lg * 4));
                fma4(o, v1, ld4(qtq + (dl * 4 + 1) * kp + lg * 4));
                fma4(o, v2, ld4(qtq + (dl * 4 + 2) * kp + lg * 4));
                fma4(o, v3, ld4(qtq + (dl * 4 + 3) * kp + lg * 4));
            }
            if (ok) {
                if (it > 0) st4(V + off, v);
                st4(Hv + off, zero4<T>());
                st4(VQ + off, o);
                const T c = freq ? lambda * freq[f] : lambda;
                local += double(c) * (double(v.x) * v.x + double(v.y) * v.y + double(v.z) * v.z + double(v.w) * v.w);
            }
        }
        grid_barrier(counter, epoch);
        if (host_phase_ms && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_b0));
        // ---- B: hs_cross work items
        for (uint64_t item = tid / G; item < Y.n_items; item += nthreads / G)
            local += double(hess_cross_item<T, G, true>(Y, X, Q1, ldq, V, VQ, w, Hv, uint32_t(item), 0));
        if constexpr (kp == 16 || kp == 32) {
            for (uint64_t slot = tid >> 5; slot < n_heavy; slot += nthreads >> 5)
                local += double(hess_heavy_row<T, int(kp), true>(heavy_rows, X, Mrow, V, VQ, w, Hv, uint32_t(slot), 0));
        }
        local = block_sum(local);
        if (threadIdx.x == 0) part_a[blockIdx.x] = local;
        grid_barrier(counter, epoch);
        if (host_phase_ms && blockIdx.x == 0 && threadIdx.x == 0) {   // every CTA has finished the row pass
            unsigned long long t_b1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_b1));
            t_rows += t_b1 - t_b0;
        }
        // ---- C: step
        const double vhv = global_sum(part_a, pk, kseq, sc, counter, epoch);
        const T alpha = T(r2 / vhv);
        local = 0;
        for (uint64_t i = tid; i < nvec; i += nthreads) {
            const V4<T> v = ldcg4(Vs + i * 4);
            V4<T> h = ldcg4(Hvs + i * 4);
            const T c = freqs ? lambda * freqs[i / (kp / 4)] : lambda;
            h.x += c * v.x; h.y += c * v.y; h.z += c * v.z; h.w += c * v.w;
            V4<T> sv = ldcg4(Ss + i * 4), r = ldcg4(Rs + i * 4);
            sv.x += alpha * v.x; sv.y += alpha * v.y; sv.z += alpha * v.z; sv.w += alpha * v.w;
            r.x -= alpha * h.x; r.y -= alpha * h.y; r.z -= alpha * h.z; r.w -= alpha * h.w;
            st4(Ss + i * 4, sv);
            st4(Rs + i * 4, r);
            local += double(r.x) * r.x + double(r.y) * r.y + double(r.z) * r.z + double(r.w) * r.w;
        }
        local = block_sum(local);
        if (threadIdx.x == 0) part_b[blockIdx.x] = local;
        grid_barrier(counter, epoch);
        const double r2n = global_sum(part_b, pk, kseq, sc, counter, epoch);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            sc->vHv[it] = vhv;
            sc->r2[it + 1] = r2n;
        }
        r2_prev = r2;
        r2 = r2n;
        ++it;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->counter[2] = unsigned(it);
        if (host_iters) *host_iters = unsigned(it);   // mapped pinned host memory: no D2H memcpy needed
        if (pk.nranks > 1) *pk.seq = kseq;
        if (host_phase_ms) *host_phase_ms = float(double(t_rows) * 1e-6);
    }
}

inline int persist_ctas_per_sm(const void *kernel) {
    static std::map<const void *, int> cache;
    auto it = cache.find(kernel);
    if (it != cache.end()) return it->second;
    int per_sm = 0;
    OC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
    OC_REQUIRE(per_sm >= 1, "persistent CG kernel does not fit an SM");
    per_sm = std::min(per_sm, kPersistMaxBlocks / kSMs);
    cache[kernel] = per_sm;
    return per_sm;
}

inline unsigned blocks_for(uint64_t groups, int G) {
    const uint64_t threads = groups * uint64_t(G);
    return unsigned((threads + kThreads - 1) / kThreads);
}

}  // namespace

#define OC_DISPATCH_G(kp, ...)                                                      \
    switch (kp) {                                                                   \
        case 4: { constexpr int G = 1; __VA_ARGS__; } break;                        \
        case 8: { constexpr int G = 2; __VA_ARGS__; } break;                        \
        case 16: { constexpr int G = 4; __VA_ARGS__; } break;                       \
        case 32: { constexpr int G = 8; __VA_ARGS__; } break;                       \
        case 64: { constexpr int G = 16; __VA_ARGS__; } break;                      \
        case 128: { constexpr int G = 32; __VA_ARGS__; } break;                     \
        default: throw Error(-6, "padded latent dimension must be 4..128");         \
    }

template <typename T>
void spmm_rows(const CsrView<T> &X, const T *A, T *C, uint32_t ldc, int kp, cudaStream_t s) {
    const uint64_t rows = X.row1 - X.row0;
    if (!rows) return;
    OC_DISPATCH_G(kp, OC_LAUNCH((k_spmm_rows<T, G>), blocks_for(rows, G), kThreads, 0, s, X, A, C, ldc));
}

template <typename T>
void spmm_update(const CsrView<T> &X, const T *S, T *XS, T *P1, uint32_t ldp, const T *Q1side,
                 T *gap, T *a1, int kp, cudaStream_t s) {
    const uint64_t rows = X.row1 - X.row0;
    if (!rows) return;
    OC_DISPATCH_G(kp, OC_LAUNCH((k_spmm_update<T, G>), blocks_for(rows, G), kThreads, 0, s, X, S, XS,
                                P1, ldp, Q1side, gap, a1));
}

template <typename T>
void grad_cross_rows(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, uint32_t ldq,
                     const T *Tm, const T *a1, const T *oQ, const T *bQ, T w, T r, T *G_, int kp,
                     cudaStream_t s) {
    if (!Y.n_items) return;
    OC_DISPATCH_G(kp, OC_LAUNCH((k_grad_cross<T, G>), blocks_for(Y.n_items, G), kThreads, 0, s, Y, X,
                                Q1, ldq, Tm, a1, oQ, bQ, w, r, G_));
}

template <typename T>
void hess_cross_rows(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, uint32_t ldq,
                     const T *V, const T *VQ, T w, T *Hv, int kp, Gate gate, double *dot_out, int notau,
                     cudaStream_t s) {
    if (!Y.n_items) return;
    OC_DISPATCH_G(kp, OC_LAUNCH((k_hess_cross<T, G>), blocks_for(Y.n_items, G), kThreads, 0, s, Y, X,
                                Q1, ldq, V, VQ, w, Hv, gate, dot_out, notau));
}

template <typename T>
void ytilde_rowsum(const OmegaView<T> &Y, T *ysum, int kp, cudaStream_t s) {
    (void)kp;
    if (!Y.n_items) return;
    OC_LAUNCH((k_ytilde_rowsum<T, 8>), blocks_for(Y.n_items, 8), kThreads, 0, s, Y, ysum);
}

template <typename T>
void side_rows(int mode, const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, const T *a1,
               const T *sa1, const T *ysum, const double *bsum, const T *V, T w, T r, T n1, T *Out,
               int kp, Gate gate, double *dot_out, cudaStream_t s) {
    const uint64_t rows = X.row1 - X.row0;
    if (!rows) return;
    if (mode == 0) {
        OC_DISPATCH_G(kp, OC_LAUNCH((k_side_rows<T, G, 0>), blocks_for(rows, G), kThreads, 0, s, Y, X,
                                    Q1, a1, sa1, ysum, bsum, V, w, r, n1, Out, gate, dot_out));
    } else {
        OC_DISPATCH_G(kp, OC_LAUNCH((k_side_rows<T, G, 1>), blocks_for(rows, G), kThreads, 0, s, Y, X,
                                    Q1, a1, sa1, ysum, bsum, V, w, r, n1, Out, gate, dot_out));
    }
}

template <typename T>
void side_diag_iter(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, T *V, const T *R, T *Hv,
                    const T *freq, T lambda, T w, T n1, int kp, int it, SolveScalars *sc, cudaStream_t s) {
    const uint64_t rows = X.row1 - X.row0;
    if (!rows) return;
    OC_DISPATCH_G(kp, {
        const unsigned blocks = std::min<unsigned>(blocks_for(rows, G), unsigned(kSMs) * 8u);
        OC_LAUNCH((k_side_diag_iter<T, G>), blocks, kThreads, 0, s, Y, X, Q1, V, R, Hv, freq, lambda, w, n1, it, sc);
    });
}

template <typename T>
void sddmm_add(const OmegaView<T> &Y, const T *Uown, uint32_t ldu, const T *Vo, uint32_t ldv,
               int kp, cudaStream_t s) {
    if (!Y.n_items) return;
    OC_DISPATCH_G(kp, OC_LAUNCH((k_sddmm_add<T, G>), blocks_for(Y.n_items, G), kThreads, 0, s, Y, Uown,
                                ldu, Vo, ldv));
}

template <typename T>
void ytilde_base(const OmegaView<T> &Y, const T *a_own, const T *b_oth, cudaStream_t s) {
    if (!Y.n_items) return;
    OC_LAUNCH((k_ytilde_base<T>), blocks_for(Y.n_items, 8), kThreads, 0, s, Y, a_own, b_oth);
}

template <typename T>
void ytilde_add_gap(const OmegaView<T> &Y, const T *gap, int by_row, cudaStream_t s) {
    if (by_row) {
        if (!Y.n_items) return;
        OC_LAUNCH((k_ytilde_gap_by_row<T>), blocks_for(Y.n_items, 8), kThreads, 0, s, Y, gap);
    } else {
        if (!Y.nnz_local) return;
        const unsigned blocks = unsigned(std::min<uint64_t>((Y.nnz_local + kThreads - 1) / kThreads,
                                                            uint64_t(kSMs) * 16));
        OC_LAUNCH((k_ytilde_gap_by_idx<T>), blocks, kThreads, 0, s, Y, gap);
    }
}

template <typename T>
void fold_hot(const T *shadow, const uint32_t *hot_feat, uint32_t n_hot, T *Out, int kp, cudaStream_t s) {
    if (!n_hot) return;
    OC_DISPATCH_G(kp, OC_LAUNCH((k_fold_hot<T, G>), blocks_for(n_hot, G), kThreads, 0, s, shadow, hot_feat,
                                n_hot, Out));
}

template <typename T>
void rowwise_dot(const T *P, const T *Q, uint32_t rows, int kp, T *out, int accumulate,
                 cudaStream_t s) {
    if (!rows) return;
    OC_DISPATCH_G(kp, OC_LAUNCH((k_rowwise_dot<T, G>), blocks_for(rows, G), kThreads, 0, s, P, Q, rows,
                                out, accumulate));
}


bool row_gram_supported(int kp) { return kp == 16 || kp == 32; }

template <typename T>
void row_gram(const uint32_t *it_slot, const uint32_t *it_beg, const uint32_t *it_cnt, uint32_t n_items,
              const uint32_t *yidx, const T *Q1, uint32_t ldq, T *M, int kp, cudaStream_t s) {
    if (!n_items) return;
    const unsigned blocks = unsigned((uint64_t(n_items) * 32 + kThreads - 1) / kThreads);
    if (kp == 32)
        OC_LAUNCH((k_row_gram<T, 32>), blocks, kThreads, 0, s, it_slot, it_beg, it_cnt, n_items, yidx, Q1, ldq, M);
    else if (kp == 16)
        OC_LAUNCH((k_row_gram<T, 16>), blocks, kThreads, 0, s, it_slot, it_beg, it_cnt, n_items, yidx, Q1, ldq, M);
    else
        throw Error(-6, "per-row Gram needs a padded latent dimension of 16 or 32");
}

template <typename T>
void hess_heavy_rows(const uint32_t *heavy_rows, uint32_t n_heavy, const CsrView<T> &X, const T *M, const T *V,
                     const T *VQ, T w, T *Hv, int kp, Gate gate, double *dot_out, int notau, cudaStream_t s) {
    if (!n_heavy) return;
    const unsigned blocks = unsigned((uint64_t(n_heavy) * 32 + kThreads - 1) / kThreads);
    if (kp == 32)
        OC_LAUNCH((k_hess_heavy<T, 32>), blocks, kThreads, 0, s, heavy_rows, n_heavy, X, M, V, VQ, w, Hv, gate, dot_out, notau);
    else if (kp == 16)
        OC_LAUNCH((k_hess_heavy<T, 16>), blocks, kThreads, 0, s, heavy_rows, n_heavy, X, M, V, VQ, w, Hv, gate, dot_out, notau);
    else
        throw Error(-6, "per-row Gram needs a padded latent dimension of 16 or 32");
}


template <typename T>
void cg_side_persist(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, T *V, T *R, T *S, T *Hv, const T *freq,
                     T lambda, T w, T n1, uint64_t D, int kp, bool diag, SolveScalars *sc, int max_cg, double eps,
                     unsigned *host_iters, uint64_t f0, const PeerK &pk_, cudaStream_t s) {
    auto launch = [&](auto kernel) {
        // per kernel INSTANCE (the instantiations for different lane-group sizes share one pointer type, so a
        // function-local static would be shared between them): co-resident CTAs per SM of this very kernel
        int per_sm = persist_ctas_per_sm(reinterpret_cast<const void *>(kernel));
        int sms = kSMs;
        int dev = 0;
        OC_CUDA(cudaGetDevice(&dev));
        OC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const unsigned grid = unsigned(std::min(kPersistMaxBlocks, per_sm * sms));
        OmegaView<T> y = Y;
        CsrView<T> x = X;
        PeerK pk = pk_;
        void *args[] = {&y, &x, &Q1, &V, &R, &S, &Hv, &freq, &lambda, &w, &n1, &D, &sc, &max_cg, &eps, &host_iters, &f0, &pk};
        OC_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kernel), dim3(grid), dim3(kThreads), args, 0, s));
        count_launch();
    };
    if (diag) {
        OC_DISPATCH_G(kp, launch(k_cg_side_persist<T, G, true>));
    } else {
        OC_DISPATCH_G(kp, launch(k_cg_side_persist<T, G, false>));
    }
}


template <typename T>
void cg_cross_persist(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, uint32_t ldq, const T *QTQ, T *V, T *R,
                      T *S, T *Hv, T *VQ, const T *freq, T lambda, T w, uint64_t D, int kp, SolveScalars *sc,
                      int max_cg, double eps, unsigned *host_iters, const uint32_t *heavy_rows, uint32_t n_heavy,
                      const T *Mrow, uint64_t f0, const PeerK &pk_, float *host_phase_ms, cudaStream_t s) {
    auto launch = [&](auto kernel) {
        // per kernel INSTANCE (the instantiations for different lane-group sizes share one pointer type, so a
        // function-local static would be shared between them): co-resident CTAs per SM of this very kernel
        int per_sm = persist_ctas_per_sm(reinterpret_cast<const void *>(kernel));
        int sms = kSMs, dev = 0;
        OC_CUDA(cudaGetDevice(&dev));
        OC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const unsigned grid = unsigned(std::min(kPersistMaxBlocks, per_sm * sms));
        OmegaView<T> y = Y;
        CsrView<T> x = X;
        PeerK pk = pk_;
        void *args[] = {&y, &x, &Q1, &ldq, &QTQ, &V, &R, &S, &Hv, &VQ, &freq, &lambda, &w, &D, &sc, &max_cg, &eps, &host_iters,
                        &heavy_rows, &n_heavy, &Mrow, &f0, &pk, &host_phase_ms};
        OC_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kernel), dim3(grid), dim3(kThreads), args, 0, s));
        count_launch();
    };
    switch (kp) {
        case 4: launch(k_cg_cross_persist<T, 1>); break;
        case 8: launch(k_cg_cross_persist<T, 2>); break;
        case 16: launch(k_cg_cross_persist<T, 4>); break;
        case 32: launch(k_cg_cross_persist<T, 8>); break;
        case 64:
            if constexpr (sizeof(T) == 4) { launch(k_cg_cross_persist<T, 16>); break; }
            [[fallthrough]];
        default: throw Error(-6, "cg_cross_persist: padded latent dimension too large for shared memory");
    }
}
bool cg_cross_persist_supported(int kp, size_t elem) { return kp <= 32 || (kp == 64 && elem == 4); }

#define OC_INSTANTIATE(T)                                                                          \
    template void spmm_rows<T>(const CsrView<T> &, const T *, T *, uint32_t, int, cudaStream_t);   \
    template void spmm_update<T>(const CsrView<T> &, const T *, T *, T *, uint32_t, const T *, T *, \
                                 T *, int, cudaStream_t);                                          \
    template void grad_cross_rows<T>(const OmegaView<T> &, const CsrView<T> &, const T *, uint32_t, \
                                     const T *, const T *, const T *, const T *, T, T, T *, int,   \
                                     cudaStream_t);                                                \
    template void hess_cross_rows<T>(const OmegaView<T> &, const CsrView<T> &, const T *, uint32_t, \
                                     const T *, const T *, T, T *, int, Gate, double *, int,       \
                                     cudaStream_t);                                                \
    template void ytilde_rowsum<T>(const OmegaView<T> &, T *, int, cudaStream_t);                  \
    template void side_rows<T>(int, const OmegaView<T> &, const CsrView<T> &, const T *, const T *, \
                               const T *, const T *, const double *, const T *, T, T, T, T *, int, \
                               Gate, double *, cudaStream_t);                                      \
    template void sddmm_add<T>(const OmegaView<T> &, const T *, uint32_t, const T *, uint32_t, int, \
                               cudaStream_t);                                                      \
    template void ytilde_base<T>(const OmegaView<T> &, const T *, const T *, cudaStream_t);        \
    template void side_diag_iter<T>(const OmegaView<T> &, const CsrView<T> &, const T *, T *, const T *, T *, \
                                    const T *, T, T, T, int, int, SolveScalars *, cudaStream_t);   \
    template void ytilde_add_gap<T>(const OmegaView<T> &, const T *, int, cudaStream_t);           \
    template void rowwise_dot<T>(const T *, const T *, uint32_t, int, T *, int, cudaStream_t);     \
    template void fold_hot<T>(const T *, const uint32_t *, uint32_t, T *, int, cudaStream_t);         \
    template void cg_side_persist<T>(const OmegaView<T> &, const CsrView<T> &, const T *, T *, T *, T *, T *,      \
                                     const T *, T, T, T, uint64_t, int, bool, SolveScalars *, int, double,      \
                                     unsigned *, uint64_t, const PeerK &, cudaStream_t);                        \
    template void cg_cross_persist<T>(const OmegaView<T> &, const CsrView<T> &, const T *, uint32_t, const T *, T *, \
                                      T *, T *, T *, T *, const T *, T, T, uint64_t, int, SolveScalars *, int,   \
                                      double, unsigned *, const uint32_t *, uint32_t, const T *, uint64_t,       \
                                      const PeerK &, float *, cudaStream_t);                                     \
    template void row_gram<T>(const uint32_t *, const uint32_t *, const uint32_t *, uint32_t,      \
                              const uint32_t *, const T *, uint32_t, T *, int, cudaStream_t);      \
    template void hess_heavy_rows<T>(const uint32_t *, uint32_t, const CsrView<T> &, const T *,    \
                                     const T *, const T *, T, T *, int, Gate, double *, int, cudaStream_t);

OC_INSTANTIATE(float)
OC_INSTANTIATE(double)

}  // namespace ocffm
