// Shared helpers for the demonet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "../../include/demonet_b200.h"

// measurement aid shared by postprocess.cu and engine.cu (C++ linkage, not part of the C ABI)
int dn_postprocess_timed(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                         const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                         float* out_scores, int64_t* out_labels, int32_t* out_counts, cudaStream_t stream, int iters,
                         float* ms3);

// dn_postprocess with one external event recorded (cudaEventRecordExternal: legal under stream capture) between the
// softmax / decode front and the sort / NMS / merge rounds -- the in-graph timing of dn_engine_profile
int dn_postprocess_marked(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                          const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                          float* out_scores, int64_t* out_labels, int32_t* out_counts, cudaStream_t stream,
                          cudaEvent_t after_front);

// Back-off of mbarrier spin loops (nanoseconds, 0 = plain spin).  A warp that spins on mbarrier.try_wait stays eligible and
// competes for its scheduler's issue slots with the warps doing the arithmetic: the producer warps run several stages
// ahead, so a late wake-up costs them nothing (DN_SLEEP_PRODUCER); DN_SLEEP_CONSUMER is the back-off of the waits that sit
// on the critical path (consumers waiting for data, epilogue warps waiting for an accumulator).
#ifndef DN_SLEEP_PRODUCER
#define DN_SLEEP_PRODUCER 0
#endif
#ifndef DN_SLEEP_CONSUMER
#define DN_SLEEP_CONSUMER 0
#endif

namespace dn {

template <int NS>
__device__ __forceinline__ void spin_backoff() {
    if (NS > 0) asm volatile("nanosleep.u32 %0;" ::"n"(NS));
}

void set_error(const char* fmt, ...);
// uint8 -> fp32 / 255 (transform.cu), used by the engine's uint8 ingest
int u8_to_f32_launch(const unsigned char* src, float* dst, size_t n, cudaStream_t s);
// squeeze-excitation whose pooling was done by the producing depthwise launch (se.cu)
int se_max_pool_slots();
int se_inplace_pooled(void* x, const float* w1, const float* b1, const float* w2t, const float* b2, int B, int HW, int C, int Cs,
                      void* workspace, size_t workspace_bytes, int dw_parts, int dw_slots, int dw_rows, cudaStream_t s,
                      bool apply = true);
float* se_scales_ptr(void* workspace, int B, int C);

#define DN_CHECK_CUDA(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            dn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DN_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define DN_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            dn::set_error(__VA_ARGS__);      \
            return (code);                   \
        }                                    \
    } while (0)

#define DN_CHECK_LAUNCH()                                                                        \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            dn::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DN_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

// ---- per-device launch state ------------------------------------------------------------------
// One process may drive several GPUs (SSDLiteB200 keeps one engine per device), and both the SM count and the
// opt-in dynamic shared-memory limit of a kernel are PER-DEVICE properties: everything cached here is keyed by
// cudaGetDevice() and guarded by a mutex (engines of different devices may run on different host threads).
constexpr int DN_MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < DN_MAX_DEVICES) ? dev : 0;
}

inline int sm_count() {
    static int n[DN_MAX_DEVICES] = {};
    static std::mutex mu;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
        // DN_SM_COUNT=<n>: size the persistent grids for n SMs (measurement aid: with two batches in flight, kernels sized
        // for half the GPU leave room for the other batch's kernel to run beside them)
        const char* v = getenv("DN_SM_COUNT");
        if (v && atoi(v) > 0 && atoi(v) < n[dev]) n[dev] = atoi(v);
    }
    return n[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize of one kernel (one `static SmemOptIn` per call site), raised on the
// CURRENT device when a launch needs more than what that device was configured for.  `occupancy` caches one
// cudaOccupancyMaxActiveBlocksPerMultiprocessor answer per device for the call sites that size their grid by it.
struct SmemOptIn {
    size_t cfg[DN_MAX_DEVICES] = {};
    int occupancy[DN_MAX_DEVICES] = {};
    std::mutex mu;
    template <typename Kernel>
    cudaError_t ensure(Kernel kern, size_t bytes, size_t set_to = 0) {
        if (bytes <= 48 * 1024) return cudaSuccess;
        const int dev = current_device();
        std::lock_guard<std::mutex> lock(mu);
        if (bytes <= cfg[dev]) return cudaSuccess;
        const size_t want = set_to > bytes ? set_to : bytes;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
        if (e == cudaSuccess) cfg[dev] = want;
        return e;
    }
    template <typename Kernel>
    cudaError_t blocks_per_sm(Kernel kern, int threads, size_t smem, int* out) {
        const int dev = current_device();
        std::lock_guard<std::mutex> lock(mu);
        if (occupancy[dev] == 0) {
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
            if (e != cudaSuccess) return e;
            occupancy[dev] = n > 0 ? n : 1;
        }
        *out = occupancy[dev];
        return cudaSuccess;
    }
    // The same for kernels that allocate TMEM: cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that
    // contains tcgen05.alloc (measured on B200 / CUDA 12.9: ncu showed one CTA per SM for dwpw_fused and the tensor-core
    // stem at 31 KB / 68 KB of shared memory), so residency is computed from the kernel's own resources: shared memory
    // (+ 1 KB the driver reserves per CTA), registers, threads and the `tmem_cols` columns every resident CTA holds.
    template <typename Kernel>
    cudaError_t blocks_per_sm_tmem(Kernel kern, int threads, size_t smem, int tmem_cols, int* out) {
        const int dev = current_device();
        std::lock_guard<std::mutex> lock(mu);
        if (occupancy[dev] == 0) {
            cudaFuncAttributes fa;
            cudaError_t e = cudaFuncGetAttributes(&fa, kern);
            if (e != cudaSuccess) return e;
            int smem_sm = 0, regs_sm = 0, thr_sm = 0;
            cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
            cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
            cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
            const size_t per_cta = smem + fa.sharedSizeBytes + 1024;
            const int regs_cta = ((fa.numRegs + 7) & ~7) * ((threads + 31) & ~31);
            int n = (int)((size_t)smem_sm / per_cta);
            if (regs_cta > 0 && regs_sm / regs_cta < n) n = regs_sm / regs_cta;
            if (thr_sm / threads < n) n = thr_sm / threads;
            if (tmem_cols > 0 && 512 / tmem_cols < n) n = 512 / tmem_cols;
            occupancy[dev] = n > 0 ? (n > 32 ? 32 : n) : 1;
        }
        *out = occupancy[dev];
        return cudaSuccess;
    }
};

// ---- programmatic dependent launch ----------------------------------------------------------
// Every kernel of the forward is launched with the programmatic-stream-serialization attribute and starts with
// pdl_trigger() + pdl_wait(): its CTAs may become resident (and run whatever comes before the wait: barrier
// setup, TMEM allocation, descriptor prefetch) while the previous kernel drains, but touch no activation memory
// before that kernel has completed and flushed.  Inside the CUDA graph these become programmatic edges.
// A kernel launched through launch_pdl MUST execute pdl_wait() before its first global access to data another
// kernel writes, and before its first global write.  DN_PDL=0 turns the attribute off (measurement aid).
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* v = getenv("DN_PDL");
        return !(v && atoi(v) == 0);
    }();
    return on;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);          // errors surface in DN_CHECK_LAUNCH
}

// the same with the grid cut into thread-block clusters of `cluster_x` consecutive CTAs (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, unsigned cluster_x,
                               Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = cluster_x;
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);          // errors surface in DN_CHECK_LAUNCH
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

// ---- activations (fp32) --------------------------------------------------------------
// hardswish(x) = x * relu6(x + 3) / 6   (torch.nn.Hardswish; mobilenetv3.py:72,141-142)
__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case DN_ACT_RELU: return fmaxf(v, 0.f);
        case DN_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
        case DN_ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
        default: return v;
    }
}

// ---- 16-bit activation storage ----------------------------------------------------------
// The library is built once per storage type (csrc/build.sh): DN_ACT_FP16=1 -> IEEE half (libdemonet_b200_fp16.so, the
// default: same bytes and the same tcgen05 kind::f16 rate as bf16 with 3 more mantissa bits -- 6.6x lower end-to-end
// error on this network, DESIGN.md "Numerics"), DN_ACT_FP16=0 -> bfloat16 (libdemonet_b200_bf16.so).  Stores saturate
// to +-65504 in the fp16 build (F2FP.SATFINITE, one instruction either way) so that an outlier cannot turn into inf.
#ifndef DN_ACT_FP16
#define DN_ACT_FP16 1
#endif
#if DN_ACT_FP16
typedef __half dn_half_t;
typedef __half2 dn_half2_t;
#define DN_ACT_DTYPE_ID 1
#define DN_TMAP_HALF CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define DN_UMMA_AB_FORMAT 0u        // tcgen05 kind::f16 instruction descriptor, a_format / b_format: 0 = f16, 1 = bf16
#else
typedef __nv_bfloat16 dn_half_t;
typedef __nv_bfloat162 dn_half2_t;
#define DN_ACT_DTYPE_ID 0
#define DN_TMAP_HALF CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define DN_UMMA_AB_FORMAT 1u
#endif

#ifdef __CUDACC__
__device__ __forceinline__ float2 h2_to_float2(uint32_t v) {
#if DN_ACT_FP16
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
#else
    float2 r;
    r.x = __uint_as_float(v << 16);
    r.y = __uint_as_float(v & 0xffff0000u);
    return r;
#endif
}
__device__ __forceinline__ uint32_t float2_to_h2(float a, float b) {
#if DN_ACT_FP16
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
#else
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
#endif
}
// the same pair as a typed value (for __hmax2 / __hmin2 on the packed result)
__device__ __forceinline__ dn_half2_t floats_to_half2(float a, float b) {
    const uint32_t r = float2_to_h2(a, b);
    return *reinterpret_cast<const dn_half2_t*>(&r);
}
__device__ __forceinline__ dn_half2_t half2_const(float v) {
#if DN_ACT_FP16
    return __float2half2_rn(v);
#else
    return __float2bfloat162_rn(v);
#endif
}
__device__ __forceinline__ float half_to_float(dn_half_t v) {
#if DN_ACT_FP16
    return __half2float(v);
#else
    return __bfloat162float(v);
#endif
}
__device__ __forceinline__ dn_half_t float_to_half(float v) {
#if DN_ACT_FP16
    const uint32_t r = float2_to_h2(v, 0.f);
    return *reinterpret_cast<const __half*>(&r);
#else
    return __float2bfloat16_rn(v);
#endif
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    float2 t;
    t = h2_to_float2(v.x); f[0] = t.x; f[1] = t.y;
    t = h2_to_float2(v.y); f[2] = t.x; f[3] = t.y;
    t = h2_to_float2(v.z); f[4] = t.x; f[5] = t.y;
    t = h2_to_float2(v.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 v;
    v.x = float2_to_h2(f[0], f[1]);
    v.y = float2_to_h2(f[2], f[3]);
    v.z = float2_to_h2(f[4], f[5]);
    v.w = float2_to_h2(f[6], f[7]);
    return v;
}
#endif  // __CUDACC__

}  // namespace dn
