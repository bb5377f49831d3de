// Shared helpers for the HandMvNet B200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace hmv {

typedef __nv_bfloat16 bf16;

// ---- error plumbing: no exception ever crosses the C boundary -------------------------------
void set_error(const std::string& msg);
const char* get_error();

#define HMV_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            hmv::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                           __FILE__ + ":" + std::to_string(__LINE__));                         \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define HMV_CHECK(cond, msg)                                                                   \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            hmv::set_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +             \
                           std::to_string(__LINE__));                                          \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

// ---- scalar conversion helpers used by the templated (bf16 | fp32) kernels -------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---- programmatic dependent launch: every kernel of the path starts with pdl_wait() (a no-op when launched
// normally) so that its launch latency and prologue overlap the tail of the previous kernel in the stream, and calls
// pdl_launch_dependents() right after it, so the NEXT kernel's CTAs may be placed (and run their own prologue: barrier
// init, TMEM allocation, descriptor prefetch, parameter loads) as soon as every CTA of this one is running ----------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// cudaFuncSetAttribute applies to the CURRENT device only: a lazily configured kernel keeps one bit per device
// (a process that drives several GPUs would otherwise launch with the default 48 KB limit on the second one).
inline bool first_use_on_this_device(unsigned long long& mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

bool pdl_enabled();      // false when HMV_NO_PDL=1
bool clusters_enabled(); // false when HMV_CLUSTER=0

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same, launched as thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, unsigned cluster_x, size_t smem,
                                         cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Folded-BN biases of a launch, passed BY VALUE as a __grid_constant__ kernel parameter: the epilogue warps then read them
// through the constant cache (one broadcast LDC per value).  With ~227 KB of shared memory carved out per CTA the L1
// is almost gone, and the former `__ldg(bias + col)` loads went to L2 on every 64-column chunk: ~850 cycles of
// exposed latency per chunk per warp, the largest single cost of the epilogues (measured with HMV_BN_PROF=1,
// profiles/r02/README.md).  1280 floats = conv3 (1024) + the next conv1 (256) of the fused layer3 seam.
constexpr int kBiasBankFloats = 1280;
struct BiasBank { float v[kBiasBankFloats]; };

// Epilogue description shared by the tensor-core and the fp32 implicit-GEMM kernels.
enum OutMode { OUT_BF16_ROWMAJOR = 0, OUT_F32_ROWMAJOR = 1, OUT_F32_NCHW = 2 };
enum ResMode { RES_NONE = 0, RES_BF16 = 1, RES_F32 = 2 };
enum ActMode { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct Epilogue {
    void* out;              // [M, ldc] (row-major modes) or [M/hw, N, hw] (NCHW mode)
    const float* bias;      // [N_alloc] fp32 (never null; zero padded)
    const void* residual;   // [*, res_ld] or null
    int ldc;
    int out_mode;
    int res_mode;
    int res_ld;
    int res_group;          // residual row = (m / res_group) * res_stride + m % res_group
    int res_stride;
    int act;
    int M;                  // valid rows
    int N;                  // valid cols (only enforced in NCHW mode; row-major pitches cover N_alloc)
    int hw;                 // pixels per image (NCHW mode)
};

}  // namespace hmv
