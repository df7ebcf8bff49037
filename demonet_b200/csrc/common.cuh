// Shared helpers for the demonet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/demonet_b200.h"

// measurement aid shared by postprocess.cu and engine.cu (C++ linkage, not part of the C ABI)
int dn_postprocess_timed(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                         const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                         float* out_scores, int64_t* out_labels, int32_t* out_counts, cudaStream_t stream, int iters,
                         float* ms3);

namespace dn {

void set_error(const char* fmt, ...);
// uint8 -> fp32 / 255 (transform.cu), used by the engine's uint8 ingest
int u8_to_f32_launch(const unsigned char* src, float* dst, size_t n, cudaStream_t s);
// squeeze-excitation whose pooling was done by the producing depthwise launch (se.cu)
int se_max_pool_slots();
int se_inplace_pooled(void* x, const float* w1, const float* b1, const float* w2t, const float* b2, int B, int HW, int C, int Cs,
                      void* workspace, size_t workspace_bytes, int dw_parts, int dw_slots, int dw_rows, cudaStream_t s);

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

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

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

// ---- bf16 packing --------------------------------------------------------------------
__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) {
    float2 r;
    r.x = __uint_as_float(v << 16);
    r.y = __uint_as_float(v & 0xffff0000u);
    return r;
}
__device__ __forceinline__ uint32_t float2_to_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    float2 t;
    t = bf16x2_to_float2(v.x); f[0] = t.x; f[1] = t.y;
    t = bf16x2_to_float2(v.y); f[2] = t.x; f[3] = t.y;
    t = bf16x2_to_float2(v.z); f[4] = t.x; f[5] = t.y;
    t = bf16x2_to_float2(v.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 v;
    v.x = float2_to_bf16x2(f[0], f[1]);
    v.y = float2_to_bf16x2(f[2], f[3]);
    v.z = float2_to_bf16x2(f[4], f[5]);
    v.w = float2_to_bf16x2(f[6], f[7]);
    return v;
}

}  // namespace dn
