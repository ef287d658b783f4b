// peer.cu -- one-shot all-reduce over NVLink / NVSwitch peer memory for the small messages of the
// multi-GPU solver: the two fp64 scalars of every CG iteration (V.Hv and ||R||^2, ffm.cpp:791-806
// summed over the row shards), the Gram stacks and the G / Hv vectors of the low-cardinality fields.
//
// Every rank owns one staging area (PeerView::base[rank]) that all peers have mapped through CUDA
// IPC.  A call is one kernel per rank and ONE NVLink hop: the message travels as 16-byte lines
// {payload.lo, seq, payload.hi, seq} (8 payload bytes guarded by the call's sequence number in
// both 8-byte halves, the flag-in-data scheme of low-latency collectives), stored straight into
// slot[seq & 1][rank] of every peer's area.  The receiver spins on each line until both flags
// carry seq and sums the nranks contributions in rank order, so there is no fence, no separate
// flag round trip, and the result is bit-identical on all ranks (the device-side CG gate relies
// on that).  Two slots are enough: a peer can only be one call ahead, because call seq+1 cannot
// complete anywhere before this rank has contributed to it.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace ocffm {
namespace {

__device__ __forceinline__ void st_line(uint4 *p, uint32_t lo, uint32_t hi, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi),
                 "r"(flag)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_line(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint4 *lines_of(const PeerView &pv, unsigned char *base, int slot, int src) {
    return reinterpret_cast<uint4 *>(base + (size_t(slot) * pv.nranks + src) * pv.cap * 2);
}

// payload of line i: one double, or floats 2i and 2i+1
__device__ __forceinline__ void pack(const double *buf, uint32_t i, uint32_t, uint32_t &lo, uint32_t &hi) {
    const unsigned long long b = __double_as_longlong(buf[i]);
    lo = uint32_t(b);
    hi = uint32_t(b >> 32);
}
__device__ __forceinline__ void pack(const float *buf, uint32_t i, uint32_t n, uint32_t &lo, uint32_t &hi) {
    lo = __float_as_uint(buf[2 * i]);
    hi = 2 * i + 1 < n ? __float_as_uint(buf[2 * i + 1]) : 0u;
}
struct Acc64 { double v; };
struct Acc32 { float a, b; };
__device__ __forceinline__ void add(Acc64 &s, uint32_t lo, uint32_t hi) {
    s.v += __longlong_as_double((unsigned long long)(hi) << 32 | lo);
}
__device__ __forceinline__ void add(Acc32 &s, uint32_t lo, uint32_t hi) {
    s.a += __uint_as_float(lo);
    s.b += __uint_as_float(hi);
}
__device__ __forceinline__ void unpack(double *buf, uint32_t i, uint32_t, const Acc64 &s) { buf[i] = s.v; }
__device__ __forceinline__ void unpack(float *buf, uint32_t i, uint32_t n, const Acc32 &s) {
    buf[2 * i] = s.a;
    if (2 * i + 1 < n) buf[2 * i + 1] = s.b;
}

template <typename T>
__global__ void __launch_bounds__(kPeerThreads)
k_peer_allreduce(PeerView pv, T *__restrict__ buf, uint32_t n, uint32_t n_lines, uint32_t seq) {
    pdl_enter();
    using Acc = typename std::conditional<sizeof(T) == 8, Acc64, Acc32>::type;
    const int slot = int(seq & 1u);
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t first = blockIdx.x * blockDim.x + threadIdx.x;
    // 1. my lines -> slot[rank] of every peer's area (posted stores, no fence)
    for (uint32_t i = first; i < n_lines; i += stride) {
        uint32_t lo, hi;
        pack(buf, i, n, lo, hi);
        for (int q = 0; q < pv.nranks; ++q)
            if (q != pv.rank) st_line(lines_of(pv, pv.base[q], slot, pv.rank) + i, lo, hi, seq);
    }
    // 2. the same lines of every peer, summed in rank order
    const unsigned long long t0 = global_ns();
    for (uint32_t i = first; i < n_lines; i += stride) {
        Acc acc{};
        for (int q = 0; q < pv.nranks; ++q) {
            uint32_t lo, hi;
            if (q == pv.rank) {
                pack(buf, i, n, lo, hi);
            } else {
                const uint4 *src = lines_of(pv, pv.base[pv.rank], slot, q) + i;
                uint4 v = ld_line(src);
                for (uint32_t spin = 1; v.y != seq || v.w != seq; ++spin) {
                    if ((spin & 0x3ffu) == 0 && global_ns() - t0 > kPeerTimeoutNs) {
                        atomicExch(pv.error, 1);   // a peer died: report, never hang the GPU
                        break;
                    }
                    v = ld_line(src);
                }
                lo = v.x;
                hi = v.z;
            }
            add(acc, lo, hi);
        }
        unpack(buf, i, n, acc);
    }
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Own rows -> every peer's copy (16-byte stores over NVLink), then a barrier: the last block of
// this rank to finish publishes seq in every peer's arrival flags and waits for theirs.
__global__ void __launch_bounds__(kPeerThreads)
k_peer_allgather(PeerView pv, PeerBuffers pb, size_t byte_off, size_t n16, size_t barrier_off,
                 unsigned long long seq) {
    pdl_enter();
    const uint4 *src = reinterpret_cast<const uint4 *>(static_cast<unsigned char *>(pb.buf[pv.rank]) + byte_off);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += size_t(gridDim.x) * blockDim.x) {
        const uint4 v = src[i];
        for (int q = 0; q < pv.nranks; ++q)
            if (q != pv.rank)
                reinterpret_cast<uint4 *>(static_cast<unsigned char *>(pb.buf[q]) + byte_off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    unsigned long long *flags = reinterpret_cast<unsigned long long *>(pv.base[pv.rank] + barrier_off);
    unsigned *counter = reinterpret_cast<unsigned *>(flags + kPeerMaxRanks);
    if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last && int(threadIdx.x) < pv.nranks && int(threadIdx.x) != pv.rank) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long *>(pv.base[threadIdx.x] + barrier_off) + pv.rank, seq);
        const unsigned long long t0 = global_ns();
        // arrival flags only grow, so ">= seq" also holds if a peer were ever a call ahead
        for (uint32_t spin = 1; ld_acquire_sys(flags + threadIdx.x) < seq; ++spin) {
            if ((spin & 0x3ffu) == 0 && global_ns() - t0 > kPeerTimeoutNs) {
                atomicExch(pv.error, 1);
                break;
            }
        }
    }
}

}  // namespace

template <typename T>
void peer_allgather_rows(const PeerView &pv, const PeerBuffers &pb, uint64_t row_lo, uint64_t row_hi,
                         uint32_t ld, unsigned long long seq, cudaStream_t s) {
    const size_t byte_off = size_t(row_lo) * ld * sizeof(T), bytes = size_t(row_hi - row_lo) * ld * sizeof(T);
    OC_REQUIRE(byte_off % 16 == 0 && bytes % 16 == 0, "peer_allgather_rows: rows must be 16-byte multiples");
    const size_t n16 = bytes / 16;
    unsigned blocks = unsigned(std::min<size_t>((n16 + kPeerThreads - 1) / kPeerThreads, size_t(kSMs) * 4));
    if (blocks == 0) blocks = 1;
    OC_LAUNCH(k_peer_allgather, blocks, kPeerThreads, 0, s, pv, pb, byte_off, n16,
              peer_barrier_offset(pv.nranks, pv.cap), seq);
}
template void peer_allgather_rows<float>(const PeerView &, const PeerBuffers &, uint64_t, uint64_t, uint32_t,
                                         unsigned long long, cudaStream_t);
template void peer_allgather_rows<double>(const PeerView &, const PeerBuffers &, uint64_t, uint64_t, uint32_t,
                                          unsigned long long, cudaStream_t);

template <typename T>
void peer_allreduce(const PeerView &pv, T *buf, size_t n, unsigned long long seq, cudaStream_t s) {
    OC_REQUIRE(n * sizeof(T) <= pv.cap, "peer_allreduce: message larger than the staging slot");
    const uint32_t n_lines = uint32_t((n * sizeof(T) + 7) / 8);
    uint32_t blocks = (n_lines + kPeerThreads - 1) / kPeerThreads;
    if (blocks > uint32_t(kPeerMaxBlocks)) blocks = kPeerMaxBlocks;
    if (blocks == 0) blocks = 1;
    // flags are the low 32 bits of the call number, never 0 (the areas start zeroed)
    uint32_t flag = uint32_t(seq & 0xffffffffull);
    OC_LAUNCH((k_peer_allreduce<T>), blocks, kPeerThreads, 0, s, pv, buf, uint32_t(n), n_lines, flag);
}

template void peer_allreduce<float>(const PeerView &, float *, size_t, unsigned long long, cudaStream_t);
template void peer_allreduce<double>(const PeerView &, double *, size_t, unsigned long long, cudaStream_t);

}  // namespace ocffm
