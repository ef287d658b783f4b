// common.cuh -- shared device/host helpers of libocffm_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <cstdlib>
#include <vector>

namespace ocffm {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

#define OC_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            throw ::ocffm::Error(e__ == cudaErrorMemoryAllocation ? -4 : -3,                  \
                                 std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" + \
                                     __FILE__ + ":" + std::to_string(__LINE__) + ")");        \
    } while (0)

#define OC_REQUIRE(cond, msg)                                         \
    do {                                                              \
        if (!(cond)) throw ::ocffm::Error(-1, std::string(msg));      \
    } while (0)

// every kernel launch of the library goes through this counter (ocffm_stats::kernel_launches)
extern thread_local uint64_t *g_launch_counter;
inline void count_launch() {
    if (g_launch_counter) ++*g_launch_counter;
}
// Programmatic dependent launch: every kernel of the library starts with pdl_enter(), which waits
// for the previous grid of the stream to complete (memory visible) and at once lets the NEXT
// kernel's CTAs be scheduled, so the launch latency and CTA ramp of the ~700 short kernels of an
// outer iteration overlap the tail of their predecessor.  OCFFM_PDL=0 launches plainly.
inline bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("OCFFM_PDL");
        return !e || atoi(e) != 0;
    }();
    return on;
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define OC_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                    \
        OC_CUDA(::ocffm::launch_kernel(kernel, (grid), (block), (smem), (stream), __VA_ARGS__)); \
        ::ocffm::count_launch();                                                            \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        if (count == n && p) return;
        release();
        n = count;
        OC_CUDA(cudaMalloc(&p, (count ? count : 1) * sizeof(T)));
    }
    void ensure(size_t count) {
        if (count > n || !p) alloc(count);
    }
    void zero(cudaStream_t s) { OC_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), s)); }
    void upload(const T *h, size_t count, cudaStream_t s) {
        alloc(count);
        if (count) OC_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T> &h, cudaStream_t s) { upload(h.data(), h.size(), s); }
    void download(T *h, size_t count, cudaStream_t s) const {
        if (count) OC_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
constexpr int kSMs = 148;  // B200

template <typename T>
struct V4 {
    T x, y, z, w;
};

template <typename T>
__device__ __forceinline__ V4<T> ld4(const T *p);
template <>
__device__ __forceinline__ V4<float> ld4<float>(const float *p) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    return {v.x, v.y, v.z, v.w};
}
template <>
__device__ __forceinline__ V4<double> ld4<double>(const double *p) {
    const double2 a = reinterpret_cast<const double2 *>(p)[0];
    const double2 b = reinterpret_cast<const double2 *>(p)[1];
    return {a.x, a.y, b.x, b.y};
}
// read-only (non-coherent) path for operands that are constant during the kernel
template <typename T>
__device__ __forceinline__ V4<T> ldg4(const T *p);
template <>
__device__ __forceinline__ V4<float> ldg4<float>(const float *p) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    return {v.x, v.y, v.z, v.w};
}
template <>
__device__ __forceinline__ V4<double> ldg4<double>(const double *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return {a.x, a.y, b.x, b.y};
}
template <typename T>
__device__ __forceinline__ void st4(T *p, const V4<T> &v);
template <>
__device__ __forceinline__ void st4<float>(float *p, const V4<float> &v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v.x, v.y, v.z, v.w);
}
template <>
__device__ __forceinline__ void st4<double>(double *p, const V4<double> &v) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2 *>(p)[1] = make_double2(v.z, v.w);
}

// scatter-add of one lane's 4 consecutive elements.  fp32: one 16-byte RED (sm_90+);
// fp64: four scalar REDs.
__device__ __forceinline__ void red4(float *p, const V4<float> &v) {
    atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v.x, v.y, v.z, v.w));
}
__device__ __forceinline__ void red4(double *p, const V4<double> &v) {
    atomicAdd(p + 0, v.x);
    atomicAdd(p + 1, v.y);
    atomicAdd(p + 2, v.z);
    atomicAdd(p + 3, v.w);
}

template <typename T>
__device__ __forceinline__ V4<T> zero4() {
    return {T(0), T(0), T(0), T(0)};
}
template <typename T>
__device__ __forceinline__ T dot4(const V4<T> &a, const V4<T> &b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
template <typename T>
__device__ __forceinline__ void fma4(V4<T> &acc, T s, const V4<T> &v) {
    acc.x += s * v.x;
    acc.y += s * v.y;
    acc.z += s * v.z;
    acc.w += s * v.w;
}

// sum over the G lanes of a lane group (G is a power of two <= 32, groups are aligned)
template <int G, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
    return group_sum<32>(v);
}

// block-wide fp64 sum, result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum(v);
    return v;
}

// block-wide fp64 sum added to *out by one thread (the order of the blocks is not fixed)
__device__ __forceinline__ void block_add(double local, double *out) {
    local = block_sum(local);
    if (threadIdx.x == 0 && local != 0.0) atomicAdd(out, local);
}

// warp-wide fp64 sum added to one of `slots` partial sums (no block barrier: the row kernels'
// warps finish at very different times)
__device__ __forceinline__ void warp_add_slot(double local, double *part, unsigned slots) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.0)
        atomicAdd(part + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (slots - 1)), local);
}

inline int ceil_div(size_t a, size_t b) { return int((a + b - 1) / b); }

}  // namespace ocffm
