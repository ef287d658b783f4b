// kernels.h -- launcher prototypes of the sm_100a kernels (defined in rows.cu, dense.cu, eval.cu).
// Every launcher is templated on the device real type T (float or double, explicitly
// instantiated) and enqueues on the given stream without synchronising.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ocffm {

// One sparse design matrix X^phi (ImpData::Xs[phi], ffm.cpp:185-257) as SoA CSR on the device.
template <typename T>
struct CsrView {
    const uint32_t *rowptr;  // [rows+1]
    const uint32_t *idx;     // [nnz]
    const T *val;            // [nnz]
    uint32_t row0, row1;     // rows owned by this rank (global row ids)
    // Hot features (skewed fields): X^T(.) contributions to feature f with hot_slot[f] >= 0 go to
    // one of kHotReplicas shadow rows (picked by warp) instead of the single row of G / Hv, so
    // the REDs of a feature that occurs millions of times do not serialise on one L2 address;
    // fold_hot() adds the replicas back.  Null when the field has no hot feature.
    const int16_t *hot_slot; // [D] or nullptr
    T *shadow;               // [n_hot * kHotReplicas * kp]
    // one feature per row (rowptr[i] == i) / that feature is i itself: the row kernels skip the
    // dependent rowptr -> idx loads (id fields; two fewer round trips per work item)
    bool diagonal, identity;
};
constexpr int kHotReplicas = 64;

// The observed pairs Omega in one orientation (U->Y by user or V->Y by item) cut into bounded
// work items so that power-law rows cannot serialise a warp: item w covers nnz
// [beg[w], beg[w] + (cnt[w] & 0x7fffffff)) of row row[w]; bit 31 of cnt marks the first chunk
// of its row (the one that adds the row's non-Omega terms exactly once).
template <typename T>
struct OmegaView {
    const uint32_t *wi_row, *wi_beg, *wi_cnt;
    uint32_t n_items;
    const uint32_t *rowptr;  // [rows+1] (global)
    const uint32_t *idx;     // [nnz] column ids (items for the user orientation, users for the item one)
    T *yt;                   // [nnz] y-tilde cache (Node::val of Y, ffm.cpp:393,400)
    uint32_t row0, row1;     // rows owned by this rank
    uint64_t nnz_local;
};

constexpr int kPeerMaxRanksK = 16;
struct SolveScalars {       // device-resident fp64 scalars of one CG solve (ffm.cpp:761-811)
    double r2[24];          // r2[it] = ||R||^2 before iteration it; r2[0] = g2 = ||G||^2
    double vHv[24];
    double bsum;            // sum(b1) of gd_side (ffm.cpp:551)
    double misc[8];
    double partials[148 * 8];   // per-block partial sums of the deterministic reductions
    unsigned counter[4];
    // V.Hv of iteration it when the Hessian pass accumulates it on the way (fused CG iteration):
    // the sum of these kDotSlots partial sums, spread so that the per-warp REDs do not serialise
    // on one address.  cg_step adds them up (every block, same order).
    double vpart[24][256];
};
constexpr int kDotSlots = 256;

// Device-side gate of the speculatively enqueued CG iteration `it` (ffm.cpp:780: the loop runs
// while g2 * cg_eps < r2): every kernel of an iteration returns at once when the stop test
// already failed, so the host can enqueue iteration it+1 before it has read r2[it+1] back.
struct Gate {
    const SolveScalars *sc;
    int it;   // < 0: always open
};
#ifdef __CUDACC__
__device__ __forceinline__ bool gate_open(const Gate &g) {
    return g.it < 0 || g.sc->r2[0] * 9e-2 < g.sc->r2[g.it];
}
#endif
constexpr Gate kNoGate{nullptr, -1};

// ---- rows.cu ---------------------------------------------------------------------------------
// C[i, 0:kp] = X_i * A   (UTX, ffm.cpp:314-331); C has leading dimension ldc
template <typename T>
void spmm_rows(const CsrView<T> &X, const T *A, T *C, uint32_t ldc, int kp, cudaStream_t s);

// XS = X S ; P1 += XS ; side blocks also gap_i = XS_i . Q1_i, a1_i += gap_i
// (update_side / update_cross, ffm.cpp:405-422, 439-449)
template <typename T>
void spmm_update(const CsrView<T> &X, const T *S, T *XS, T *P1, uint32_t ldp, const T *Q1side,
                 T *gap, T *a1, int kp, cudaStream_t s);

// gd_cross row pass (ffm.cpp:678-700): scatters X_i^T (pk_i + w (T_i + (a_i - r) oQ + bQ)) into G
template <typename T>
void grad_cross_rows(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, uint32_t ldq,
                     const T *Tm, const T *a1, const T *oQ, const T *bQ, T w, T r, T *G, int kp,
                     cudaStream_t s);

// hs_cross row pass (ffm.cpp:715-738): phi = X_i V, tau = X_i (V QTQ), ka = sum_j (phi.q_j) q_j
template <typename T>
void hess_cross_rows(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, uint32_t ldq,
                     const T *V, const T *VQ, T w, T *Hv, int kp, Gate gate, double *dot_out, int notau,
                     cudaStream_t s);   // notau: the w * tau term is already in Hv (rowgemm_dir hv_scale)   // dot_out (optional) += V . Hv, as sum over work items of phi . z

// ysum[row] = sum of y-tilde over the row (first half of gd_side's z_i, ffm.cpp:577-580)
template <typename T>
void ytilde_rowsum(const OmegaView<T> &Y, T *ysum, int kp, cudaStream_t s);

// gd_side (mode 0, ffm.cpp:572-589) / hs_side (mode 1, ffm.cpp:603-624) row pass
template <typename T>
void side_rows(int mode, const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, const T *a1,
               const T *sa1, const T *ysum, const double *bsum, const T *V, T w, T r, T n1, T *Out,
               int kp, Gate gate, double *dot_out, cudaStream_t s);

// Same-side CG iteration on a DIAGONAL field (every row has exactly one feature and the features
// form a permutation, e.g. a user-id / item-id field): hs_side (ffm.cpp:603-624) is then local to
// a row, so cg_dir + side_rows<1> + cg_reg_dot collapse into one pass without atomics:
//   v = (it ? R_f + beta V_f : V_f);  Hv_f = lambda_f v + q_i x d_i (q_i . x v);  vHv += v . Hv_f
template <typename T>
void side_diag_iter(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, T *V, const T *R, T *Hv,
                    const T *freq, T lambda, T w, T n1, int kp, int it, SolveScalars *sc, cudaStream_t s);

// In-kernel scalar all-reduce of the persistent CG kernels on multi-rank contexts: a second, tiny
// peer-memory area (2 slots x nranks 16-byte lines {lo, seq, hi, seq} per rank, same flag-in-data
// protocol as peer.cu) with its OWN sequence counter that lives on the device -- every rank runs the
// same number of reductions per kernel (the stop decisions are bit-identical), so the counters stay
// in step without the host.  nranks <= 1: unused.
struct PeerK {
    uint4 *area[kPeerMaxRanksK];   // every rank's area, mapped here (area[rank] is local)
    int nranks, rank;
    unsigned *seq;                 // device-resident call counter (last used sequence number)
    int *error;                    // mapped host flag: a wait timed out
};

// The whole CG solve of a same-side half in one cooperative launch (single rank, field without hot
// features): S, R, V as left by cg_init; on return sc->r2[1..it], sc->vHv[0..it-1] and sc->counter[2] = it.
template <typename T>
void cg_side_persist(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, T *V, T *R, T *S, T *Hv, const T *freq,
                     T lambda, T w, T n1, uint64_t D, int kp, bool diag, SolveScalars *sc, int max_cg, double eps,
                     unsigned *host_iters, uint64_t f0, const PeerK &pk, cudaStream_t s);
// host_iters: mapped pinned, receives the iteration count.  f0 / D: this rank's feature slice (sliced
// halves of multi-rank contexts: vectors are indexed globally, the vector passes run on [f0, f0 + D))

// ... and of a cross half (QTQ [kp x kp] of this pair; kp <= 32, or 64 in fp32)
bool cg_cross_persist_supported(int kp, size_t elem);
template <typename T>
void cg_cross_persist(const OmegaView<T> &Y, const CsrView<T> &X, const T *Q1, uint32_t ldq, const T *QTQ, T *V, T *R,
                      T *S, T *Hv, T *VQ, const T *freq, T lambda, T w, uint64_t D, int kp, SolveScalars *sc,
                      int max_cg, double eps, unsigned *host_iters, const uint32_t *heavy_rows, uint32_t n_heavy,
                      const T *Mrow, uint64_t f0, const PeerK &pk, float *host_phase_ms, cudaStream_t s);
// n_heavy > 0: Y is the light list, heavy rows use Mrow.  host_phase_ms (mapped pinned, optional): receives the
// time this launch spent in its hs_cross row phases (%globaltimer between the grid barriers, CTA 0)

// y-tilde[t] += U_row . Vo[idx[t]]   (update_cross, ffm.cpp:451-464; also init_y_tilde per pair)
template <typename T>
void sddmm_add(const OmegaView<T> &Y, const T *Uown, uint32_t ldu, const T *Vo, uint32_t ldv,
               int kp, cudaStream_t s);

// y-tilde[t] = a[row] + b[idx[t]] - 1  (base of init_y_tilde, ffm.cpp:393)
template <typename T>
void ytilde_base(const OmegaView<T> &Y, const T *a_own, const T *b_oth, cudaStream_t s);

// y-tilde[t] += gap[row]  (own orientation) / += gap[idx[t]] (other orientation) (ffm.cpp:423-436)
template <typename T>
void ytilde_add_gap(const OmegaView<T> &Y, const T *gap, int by_row, cudaStream_t s);

// out[i] (+)= sum_d P[i,d] Q[i,d]  (add_side, ffm.cpp:352-358)
template <typename T>
void rowwise_dot(const T *P, const T *Q, uint32_t rows, int kp, T *out, int accumulate,
                 cudaStream_t s);

// Out[hot_feat[s], :] += sum_r shadow[(s * kHotReplicas + r), :]  (see CsrView::hot_slot)
template <typename T>
void fold_hot(const T *shadow, const uint32_t *hot_feat, uint32_t n_hot, T *Out, int kp, cudaStream_t s);

// Per-row observed Gram (rows.cu "Mrow"): M[slot] = sum over the nnz of an item of q_j q_j^T, and the
// hs_cross pass of the heavy rows from it (z_i = (1-w) M_i phi_i + w tau_i, Hv += X_i^T z_i).
bool row_gram_supported(int kp);
template <typename T>
void row_gram(const uint32_t *it_slot, const uint32_t *it_beg, const uint32_t *it_cnt, uint32_t n_items,
              const uint32_t *yidx, const T *Q1, uint32_t ldq, T *M, int kp, cudaStream_t s);
template <typename T>
void hess_heavy_rows(const uint32_t *heavy_rows, uint32_t n_heavy, const CsrView<T> &X, const T *M, const T *V,
                     const T *VQ, T w, T *Hv, int kp, Gate gate, double *dot_out, int notau, cudaStream_t s);

// ---- dense.cu --------------------------------------------------------------------------------
// Out64[Kc x kp] += A[rows x Kc]^T B[rows x kp]; colsum64[0:kp] += B^T 1 ; wsum64[0:kp] += B^T wvec
// (mm(a,b,c,k,l) ffm.cpp:41-45 for all cross pairs at once + mv ffm.cpp:660-661)
template <typename T>
void gram_stack(const T *A, uint32_t lda, uint32_t Kc, const T *B, uint32_t ldb, int kp,
                uint32_t row0, uint32_t row1, const T *wvec, double *Out64, double *colsum64,
                double *wsum64, int acc_double, cudaStream_t s);

// gram_tc.cu (fp32, kp = 32, Kc = 128 / 256, B = columns [bcol, bcol + kp) of A): the same contraction on
// tcgen05 (TMA-fed MN-major operands, 3xTF32, fp32 TMEM accumulate flushed to fp64 every 512 rows)
bool gram_tc_supported(uint32_t Kc, int kp, uint32_t lda);
void gram_stack_tc(const float *A, uint32_t lda, uint32_t Kc, uint32_t bcol, uint32_t row0, uint32_t row1,
                   const float *wvec, double *Out64, double *colsum64, double *wsum64, cudaStream_t s);

// row GEMM C[M x 32] = A[M x Ka] B[Ka x 32] on tcgen05 (Ka = 128 / 256, fp32 in / out, 3xTF32)
bool rowgemm_tc_supported(uint32_t Ka, int kp, uint32_t lda, uint64_t M);
void rowgemm_tc(const float *A, uint32_t lda, uint32_t Ka, const float *B, float *C, uint64_t M, cudaStream_t s);

// C[M x kp] = A[M x Ka] B[Ka x kp]   (T = P~ Gstack, ffm.cpp:663-670; VQTQ = V QTQ, ffm.cpp:799)
template <typename T>
void rowgemm(const T *A, uint32_t lda, uint32_t Ka, const T *B, T *C, uint64_t M, int kp, Gate gate,
             cudaStream_t s);

template <typename T>
void convert_from_f64(const double *src, T *dst, uint64_t n, cudaStream_t s);
// model blocks cross the C ABI as fp64 [rows x k]; the device keeps T [rows x ld] (zero padded)
template <typename T>
void pad_from_f64(const double *src, T *dst, uint64_t rows, uint32_t k, uint32_t ld, cudaStream_t s);
template <typename T>
void unpad_to_f64(const T *src, uint32_t ld, double *dst, uint64_t rows, uint32_t k, cudaStream_t s);
// dst[rows x ld] (zero padded beyond k) = U(-scale, scale), a counter-based function of (seed, stream, row, col)
template <typename T>
void init_uniform(T *dst, uint64_t rows, uint32_t k, uint32_t ld, uint64_t seed, uint64_t stream, double scale,
                  cudaStream_t s);

// G += lambda * (freq ? freq[row] : 1) * W ; R = -G ; V = R ; S = 0 ; sc->r2[0] += ||G||^2
template <typename T>
void cg_init(T *G, const T *W, const T *freq, T lambda, T *R, T *V, T *S, uint64_t D, int kp,
             SolveScalars *sc, double *host_g2, cudaStream_t s);   // host_*: optional mapped pinned copy of the scalar
// it > 0: V = R + (r2[it]/r2[it-1]) V ; always Hv = 0
// vv != 0: also sc->vHv[it] += lambda sum_f c_f |V_f|^2 over rows [sum_lo, sum_hi) (the regulariser's
// share of V.Hv when the Hessian pass accumulates the data share itself)
template <typename T>
void cg_dir(T *V, const T *R, T *Hv, uint64_t n, int it, SolveScalars *sc, const T *freq, T lambda, int kp,
            uint64_t sum_lo, uint64_t sum_hi, int vv, cudaStream_t s);
// cross halves: the same direction update folded into VQ = V * QTQ (one pass over V)
template <typename T>
void rowgemm_dir(T *V, const T *R, T *Hv, const T *freq, T lambda, uint64_t sum_lo, uint64_t sum_hi,
                 const T *B, T *C, uint64_t M, int kp, int it, SolveScalars *sc, T hv_scale, cudaStream_t s);
// hv_scale: Hv = hv_scale * C instead of 0 (see DirFuse::hv_scale; then hess_cross_rows runs with notau)
// Hv += lambda * (freq ? freq[row] : 1) * V ; sc->vHv[it] += V . Hv
template <typename T>
void cg_reg_dot(T *Hv, const T *V, const T *freq, T lambda, uint64_t D, int kp, int it,
                SolveScalars *sc, int gated, cudaStream_t s);
// alpha = r2[it]/vHv[it] ; S += alpha V ; R -= alpha (Hv + lambda c_f V) ; sc->r2[it+1] += ||R||^2
// (lambda = 0 when Hv already contains the regulariser)
template <typename T>
void cg_step(T *S, T *R, const T *V, const T *Hv, uint64_t n, int it, SolveScalars *sc, const T *freq,
             T lambda, int kp, int slotted, double *host_r2, cudaStream_t s);   // slotted: V.Hv = sum(sc->vpart[it])
template <typename T>
void axpy(T *y, const T *x, T alpha, uint64_t n, cudaStream_t s);
// dst[i] = src[pos[i]]: refreshes one orientation of the y-tilde cache from the other
template <typename T>
void gather_copy(T *dst, const T *src, const uint32_t *pos, uint64_t n, cudaStream_t s);
// out64 += sum(x) / sum(x*x) / colsum of a [rows x ld] matrix
template <typename T>
void reduce_sum(const T *x, uint64_t n, int square, double *out64, cudaStream_t s);
template <typename T>
void col_sums(const T *A, uint32_t lda, uint32_t cols, uint32_t row0, uint32_t row1, double *out64,
              cudaStream_t s);
// out[i] (+)= A[i, 0:cols] . v   (cache_sasb second mv, ffm.cpp:528,532)
template <typename T>
void matvec_rows(const T *A, uint32_t lda, uint32_t cols, uint32_t rows, const T *v, T *out,
                 cudaStream_t s);
// sum over Omega of (yt^2 - w (yt + 1 - r)^2) into out64 (objective, ffm.cpp:1338-1341 regrouped)
template <typename T>
void omega_objective(const T *yt, uint64_t nnz, T w, T r, double *out64, cudaStream_t s);

// ---- eval.cu ---------------------------------------------------------------------------------
// Fused full-ranking scorer + top-80 (pred_z ffm.cpp:915-923 + the argmax loops of prec_k / ndcg,
// ffm.cpp:1029-1046, 1074-1108): for test rows [row0,row1) scores z = bt + P~va_i . Q~va_j over
// items [0, n_ranked) are produced tile by tile in shared memory and never written to HBM;
// ids[row*80 + rank] gets the 80 best (first index wins ties), UINT32_MAX padded.
// number of item-range splits score_topk wants for `rows` test rows (<= 32)
uint32_t score_topk_splits(uint32_t rows, uint32_t n_ranked);
template <typename T>
void score_topk(const T *Pva, const T *Qva, uint32_t Kc, const T *bt, uint32_t row0, uint32_t row1,
                uint32_t n_ranked, const uint8_t *cold, uint32_t nsplit, T *part_score,
                uint32_t *part_id, uint32_t *ids, cudaStream_t s);
// merge of the per-item-range partial lists (also used behind the tensor-core scorer)
template <typename T>
void merge_topk(const T *part_score, const uint32_t *part_id, uint32_t nsplit, uint32_t row0,
                uint32_t row1, uint32_t *ids, cudaStream_t s);

// ---- eval_tc.cu (fp32 contexts): tcgen05 3xTF32 scorer -------------------------------------------
bool score_topk_tc_supported(uint32_t Kc);
size_t score_topk_tc_cand_slots();
uint32_t score_topk_tc_splits(uint32_t rows, uint32_t n_ranked);
void split_tf32(const float *x, float *hi, float *lo, uint64_t n, cudaStream_t s);
void score_topk_tc(const float *Phi, const float *Plo, uint64_t p_rows, const float *Qhi, const float *Qlo,
                   uint64_t q_rows, uint32_t Kc, const float *bt, uint32_t row0, uint32_t row1,
                   uint32_t n_ranked, const uint8_t *cold, uint32_t nsplit, float *cand_score,
                   uint32_t *cand_id, float *part_score, uint32_t *part_id, uint32_t *row_thr, cudaStream_t s);

// top-80 of a plain score vector (the `popular` ranking shared by all cold rows)
template <typename T>
void vector_topk(const T *z, uint32_t n_ranked, uint32_t *ids80, cudaStream_t s);
// hits / dcg / idcg per row from the top-80 ids and the test labels, summed into acc64[0:5]
// (hits@K), acc64[5:10] (sum of dcg/idcg @K); plus ploss sum in acc64[10] (ffm.cpp:982-986)
template <typename T>
void eval_metrics(const uint32_t *ids, const uint32_t *cold_ids80, const uint8_t *cold,
                  const uint32_t *lab_rowptr, const uint32_t *lab_idx, uint32_t row0, uint32_t row1,
                  const T *Pva, const T *Qva, uint32_t Kc, const T *at, const T *bt,
                  const T *popular, uint32_t n_items, uint32_t n_ranked, double *acc64,
                  cudaStream_t s);

// ---- peer.cu: one-shot all-reduce over NVLink peer memory (small messages of the sharded solver) ---
constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxBlocks = 64;
constexpr int kPeerThreads = 256;
constexpr unsigned long long kPeerTimeoutNs = 120ull * 1000 * 1000 * 1000;   // host-side skew between ranks can be seconds
struct PeerView {
    unsigned char *base[kPeerMaxRanks];   // every rank's staging area, mapped here (base[rank] is local)
    int nranks, rank;
    size_t cap;                           // largest message in bytes; a [slot][src] region is 2*cap
    int *error;                           // local flag: set when a wait timed out
};
// area layout: [2 slots][nranks] line regions of 2*cap bytes, then the barrier block
// (kPeerMaxRanks arrival flags + one block counter)
inline size_t peer_barrier_offset(int nranks, size_t cap) { return size_t(2) * nranks * cap * 2; }
inline size_t peer_area_bytes(int nranks, size_t cap) { return peer_barrier_offset(nranks, cap) + 1024; }
// The same [rows x ld] buffer of every rank, mapped here (buf[rank] is local).
struct PeerBuffers {
    void *buf[kPeerMaxRanks];
};
// All-gather of row slices by direct stores: this rank's rows [row_lo, row_hi) go to the same
// offset of every peer's buffer, then the ranks meet at a flag barrier inside the same kernel, so
// the next kernel of every stream sees the whole matrix.  seq: 1, 2, 3, ... per call.
template <typename T>
void peer_allgather_rows(const PeerView &pv, const PeerBuffers &pb, uint64_t row_lo, uint64_t row_hi,
                         uint32_t ld, unsigned long long seq, cudaStream_t s);
// buf[0:n] <- sum over ranks, in rank order, identical bits everywhere; seq: 1, 2, 3, ... per call
template <typename T>
void peer_allreduce(const PeerView &pv, T *buf, size_t n, unsigned long long seq, cudaStream_t s);

}  // namespace ocffm
