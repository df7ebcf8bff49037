// Bandwidth-bound CUDA-core kernels of the SSDLite backbone for sm_100a:
//   * depthwise k x k stencil (k in {3,5}, stride in {1,2}) + folded BN + activation   (dn_dwconv)
//   * stem: input normalisation + dense 3x3 stride-2 conv + folded BN + activation     (dn_stem_conv)
// Activations are NHWC bf16, 8 channels (16 B) per thread access, fp32 accumulation.
#include <cstdlib>

#include "common.cuh"
#include "dwconv.cuh"

namespace dn {

// ---------------------------------------------------------------------------------------------
// Depthwise conv.  Thread = 8 channels x TW consecutive output columns of one output row.
// Consecutive threads walk the channel vectors of a pixel, then the next column tile, so every
// global access of a warp is a run of contiguous 16-byte vectors; the k-row / k-column overlap
// between neighbouring threads is served by L1.
// Reference: ConvBNActivation(groups=C), demonet/models/mobilenetv2.py:32-55.
// ---------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float dw_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));      // x * relu6(x + 3) / 6
    return v;
}

template <int KS, int S, int TW, int ACT>
__global__ void __launch_bounds__(256)
dwconv_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              uint4* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
    constexpr int PAD = (KS - 1) / 2;
    constexpr int NV = (TW - 1) * S + KS;
    const int CV = C >> 3;
    const int WT = (Wo + TW - 1) / TW;
    const unsigned total = (unsigned)B * Ho * WT * CV;          // < 2^31, checked on the host
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (tid >= total) return;
    const int cv = (int)(tid % (unsigned)CV);
    unsigned r = tid / (unsigned)CV;
    const int wt = (int)(r % (unsigned)WT);
    r /= (unsigned)WT;
    const int oh = (int)(r % (unsigned)Ho);
    const int b = (int)(r / (unsigned)Ho);
    const int ow0 = wt * TW;
    const int iw0 = ow0 * S - PAD;

    // packed fp32 math (sm_100 FFMA2: two fp32 FMAs per instruction), channel pairs (2q, 2q+1)
    float2 acc[TW][4];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + cv * 2);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + cv * 2 + 1);
#pragma unroll
        for (int t = 0; t < TW; ++t) {
            acc[t][0] = make_float2(b0.x, b0.y); acc[t][1] = make_float2(b0.z, b0.w);
            acc[t][2] = make_float2(b1.x, b1.y); acc[t][3] = make_float2(b1.z, b1.w);
        }
    }
    const uint4* xb = x + (long long)b * H * W * CV + cv;
    const bool interior_w = (iw0 >= 0) && (iw0 + NV <= W);      // no column of the window is padding
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
        const int ih = oh * S - PAD + kh;
        if (ih < 0 || ih >= H) continue;
        const uint4* xr = xb + (long long)(ih * W + iw0) * CV;
        // one load + one bf16->fp32 unpack per input vector of the row window; every vector then feeds up to
        // KS taps x TW outputs from registers
        float2 in[NV][4];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int iw = iw0 + i;
            const uint4 v = (interior_w || (iw >= 0 && iw < W)) ? __ldg(xr + i * CV) : make_uint4(0u, 0u, 0u, 0u);
            in[i][0] = h2_to_float2(v.x);
            in[i][1] = h2_to_float2(v.y);
            in[i][2] = h2_to_float2(v.z);
            in[i][3] = h2_to_float2(v.w);
        }
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
            const float4* wp = reinterpret_cast<const float4*>(w + (kh * KS + kw) * C) + cv * 2;
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                  make_float2(w1.z, w1.w)};
#pragma unroll
            for (int t = 0; t < TW; ++t) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[t][q] = __ffma2_rn(in[t * S + kw][q], wv[q], acc[t][q]);
            }
        }
    }
    uint4* yo = y + (((long long)b * Ho + oh) * Wo + ow0) * CV + cv;
#pragma unroll
    for (int t = 0; t < TW; ++t) {
        if (ow0 + t < Wo) {
            uint4 o;
            o.x = float2_to_h2(dw_act<ACT>(acc[t][0].x), dw_act<ACT>(acc[t][0].y));
            o.y = float2_to_h2(dw_act<ACT>(acc[t][1].x), dw_act<ACT>(acc[t][1].y));
            o.z = float2_to_h2(dw_act<ACT>(acc[t][2].x), dw_act<ACT>(acc[t][2].y));
            o.w = float2_to_h2(dw_act<ACT>(acc[t][3].x), dw_act<ACT>(acc[t][3].y));
            yo[(long long)t * CV] = o;
        }
    }
}

template <int KS, int S, int TW>
static int launch_dw(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int Ho,
                     int Wo, int act, cudaStream_t stream) {
    const long long total = (long long)B * Ho * ((Wo + TW - 1) / TW) * (C / 8);
    DN_REQUIRE(total < (1ll << 31) && (long long)B * H * W * (C / 8) < (1ll << 31), DN_ERR_UNSUPPORTED,
               "depthwise problem too large");
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const uint4* xi = (const uint4*)x;
    uint4* yo = (uint4*)y;
    switch (act) {
        case DN_ACT_RELU: launch_pdl(dwconv_kernel<KS, S, TW, DN_ACT_RELU>, blocks, 256, 0, stream, xi, w, bias, yo, B, H, W, C, Ho, Wo); break;
        case DN_ACT_RELU6: launch_pdl(dwconv_kernel<KS, S, TW, DN_ACT_RELU6>, blocks, 256, 0, stream, xi, w, bias, yo, B, H, W, C, Ho, Wo); break;
        case DN_ACT_HSWISH: launch_pdl(dwconv_kernel<KS, S, TW, DN_ACT_HSWISH>, blocks, 256, 0, stream, xi, w, bias, yo, B, H, W, C, Ho, Wo); break;
        default: launch_pdl(dwconv_kernel<KS, S, TW, DN_ACT_NONE>, blocks, 256, 0, stream, xi, w, bias, yo, B, H, W, C, Ho, Wo);
    }
    DN_CHECK_LAUNCH();
    return DN_OK;
}

// ---------------------------------------------------------------------------------------------
// Stem.  Thread = one output pixel, all COUT channels in registers; weights broadcast from smem.
// Reads the caller's fp32 NCHW image, normalises on the fly ((x - mean) / std, zero padding is
// applied AFTER normalisation exactly as conv2d pads the normalised tensor: transform.py:129-138,
// mobilenetv3.py:141-142) and writes NHWC bf16.
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(128)
stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                 uint4* __restrict__ y, int B, int H, int W, int Ho, int Wo, float m0, float m1, float m2, float s0,
                 float s1, float s2, int act) {
    __shared__ __align__(16) float sw[27 * COUT];
    __shared__ float sb[COUT];
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const long long total = (long long)B * Ho * Wo;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int ow = (int)(tid % Wo);
    const int oh = (int)((tid / Wo) % Ho);
    const int b = (int)(tid / ((long long)Wo * Ho));
    float2 acc[COUT / 2];
#pragma unroll
    for (int c = 0; c < COUT / 2; ++c) acc[c] = make_float2(sb[2 * c], sb[2 * c + 1]);
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float* plane = img + ((long long)b * 3 + ci) * H * W;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = oh * 2 - 1 + kh;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = ow * 2 - 1 + kw;
                float v = 0.f;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W)
                    v = __fdiv_rn(__fsub_rn(__ldg(plane + (long long)ih * W + iw), mean[ci]), stdv[ci]);
                const float4* wr = reinterpret_cast<const float4*>(sw + ((ci * 3 + kh) * 3 + kw) * COUT);
                const float2 vv = make_float2(v, v);
#pragma unroll
                for (int c4 = 0; c4 < COUT / 4; ++c4) {
                    const float4 ww = wr[c4];
                    acc[c4 * 2 + 0] = __ffma2_rn(vv, make_float2(ww.x, ww.y), acc[c4 * 2 + 0]);
                    acc[c4 * 2 + 1] = __ffma2_rn(vv, make_float2(ww.z, ww.w), acc[c4 * 2 + 1]);
                }
            }
        }
    }
    uint4* yo = y + tid * (COUT / 8);
#pragma unroll
    for (int v = 0; v < COUT / 8; ++v) {
        uint4 o;
        o.x = float2_to_h2(apply_act(acc[v * 4 + 0].x, act), apply_act(acc[v * 4 + 0].y, act));
        o.y = float2_to_h2(apply_act(acc[v * 4 + 1].x, act), apply_act(acc[v * 4 + 1].y, act));
        o.z = float2_to_h2(apply_act(acc[v * 4 + 2].x, act), apply_act(acc[v * 4 + 2].y, act));
        o.w = float2_to_h2(apply_act(acc[v * 4 + 3].x, act), apply_act(acc[v * 4 + 3].y, act));
        yo[v] = o;
    }
}

// Kernel per layer shape (measured on B200, profiles/r01_dw_*.txt): stride-1 layers with at least 8 rows run on the
// TMA-fed row stream (dwconv_stream.cu, 3.0-4.0 TB/s), stride-2 layers with wide enough channel blocks on its
// stride-2 sibling (dwconv_stream2.cu); of the rest, the TMA-fed shared-memory tiles win on the large maps
// (>= 40x40 outputs) and the register-tiled direct kernel on the small ones.
DwImpl dw_choose(int H, int W, int C, int k, int stride) {
    static const int forced = [] {                  // measurement aid: DN_DW_IMPL=1|2|4|8 forces one kernel where it applies
        const char* e = getenv("DN_DW_IMPL");
        return e ? atoi(e) : 0;
    }();
    const int pad = (k - 1) / 2;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    DwStream sp;
    int tw2 = 0;
    const bool stream_ok = H >= 8 && dw_stream_plan(H, W, C, k, stride, &sp);
    const bool stream2_ok = stride == 2 && dw_stream2_plan(H, W, C, k, &sp, &tw2);
    const DwImpl tiled = (long long)Ho * Wo >= 1600 ? DW_TMA : DW_DIRECT;      // 40x40 outputs and up (r01 A/B table)
    if (forced == DW_DIRECT) return DW_DIRECT;
    if (forced == DW_TMA) return DW_TMA;
    if (forced == DW_STREAM) return stream_ok ? DW_STREAM : tiled;
    if (forced == DW_STREAM2) return stream2_ok ? DW_STREAM2 : (stream_ok ? DW_STREAM : tiled);
    if (forced == 16) return stream_ok ? DW_STREAM : tiled;          // the plan before the stride-2 stream existed
    // stride 2: the row stream wins when a channel block is at least 24 channels wide (48-byte TMA rows) and the map has
    // at least 8 output rows (r01 A/B: 80x80x72 k5 0.162 -> 0.079 ms, 20x20x672 k5 0.084 -> 0.067, 40x40x240 k3
    // 0.062 -> 0.048; 160x160x64, where the block is 8 channels, stays on the tiles: 0.179 vs 0.334)
    if (stream2_ok && sp.CB >= 24 && Ho >= 8) return DW_STREAM2;
    return stream_ok ? DW_STREAM : tiled;
}

}  // namespace dn

using namespace dn;

extern "C" int dn_dwconv(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int k,
                         int stride, int act, void* stream_) {
    DN_REQUIRE(x && w && bias && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(C % 8 == 0, DN_ERR_UNSUPPORTED, "depthwise channels must be a multiple of 8 (got %d)", C);
    DN_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), DN_ERR_UNSUPPORTED,
               "depthwise supports k in {3,5}, stride in {1,2} (got k=%d stride=%d)", k, stride);
    const int pad = (k - 1) / 2;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    cudaStream_t s = (cudaStream_t)stream_;
    const DwImpl impl = dw_choose(H, W, C, k, stride);
    if (impl == DW_STREAM) {
        DwStream sp;
        DN_REQUIRE(dw_stream_plan(H, W, C, k, stride, &sp), DN_ERR_UNSUPPORTED, "no stream plan");
        CUtensorMap tm;
        int rc = dw_stream_make_tmap(&tm, x, B, H, W, C, k, sp);
        if (rc) return rc;
        return dwconv_stream_launch(tm, sp, w, bias, y, B, H, W, C, k, act, s);
    }
    if (impl == DW_STREAM2) {
        DwStream sp;
        int tw = 0;
        DN_REQUIRE(dw_stream2_plan(H, W, C, k, &sp, &tw), DN_ERR_UNSUPPORTED, "no stride-2 stream plan");
        CUtensorMap tm;
        int rc = dw_stream2_make_tmap(&tm, x, B, H, W, C, k, sp);
        if (rc) return rc;
        return dwconv_stream2_launch(tm, sp, tw, w, bias, y, B, H, W, C, k, act, s);
    }
    if (impl == DW_TMA) {
        DwTiling tl;
        int rc = dw_plan(H, W, C, k, stride, &tl);
        if (rc) return rc;
        CUtensorMap tm;
        rc = dw_make_tmap(&tm, x, B, H, W, C, tl);
        if (rc) return rc;
        return dwconv_tma_launch(tm, tl, w, bias, y, B, H, W, C, k, stride, act, s);
    }
    if (k == 3 && stride == 1) return launch_dw<3, 1, 4>(x, w, bias, y, B, H, W, C, Ho, Wo, act, s);
    if (k == 3 && stride == 2) return launch_dw<3, 2, 2>(x, w, bias, y, B, H, W, C, Ho, Wo, act, s);
    if (k == 5 && stride == 1) return launch_dw<5, 1, 4>(x, w, bias, y, B, H, W, C, Ho, Wo, act, s);
    return launch_dw<5, 2, 2>(x, w, bias, y, B, H, W, C, Ho, Wo, act, s);
}

extern "C" int dn_dwconv_plan_info(int H, int W, int C, int k, int stride, int32_t* out8) {
    DN_REQUIRE(out8 != nullptr, DN_ERR_INVALID, "NULL argument");
    DN_REQUIRE(H > 0 && W > 0 && C > 0 && C % 8 == 0 && (k == 3 || k == 5) && (stride == 1 || stride == 2), DN_ERR_INVALID,
               "bad depthwise shape");
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    const DwImpl impl = dw_choose(H, W, C, k, stride);
    out8[0] = (int32_t)impl;
    DwStream sp{};
    int tw = 4;
    if (impl == DW_STREAM) {
        DN_REQUIRE(dw_stream_plan(H, W, C, k, stride, &sp), DN_ERR_UNSUPPORTED, "no stream plan");
    } else if (impl == DW_STREAM2) {
        DN_REQUIRE(dw_stream2_plan(H, W, C, k, &sp, &tw), DN_ERR_UNSUPPORTED, "no stride-2 stream plan");
    } else {
        return DN_OK;
    }
    out8[1] = sp.CB, out8[2] = sp.ncb, out8[3] = sp.nstrip, out8[4] = tw, out8[5] = sp.threads, out8[6] = sp.nst;
    out8[7] = (int32_t)sp.smem;
    return DN_OK;
}

// Depthwise conv followed by squeeze-excitation of its output (InvertedResidual with use_se, mobilenetv3.py:43-96):
// when the layer runs on a row-stream kernel (either stride) and the batch is large enough, the stream leaves the SE
// channel sums in the workspace and the SE pooling pass is skipped (*pooled_out = 1); otherwise dn_dwconv + dn_se_inplace.
extern "C" int dn_dwconv_se(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int k,
                            int stride, int act, const float* se_w1, const float* se_b1, const float* se_w2t,
                            const float* se_b2, int Cs, void* workspace, size_t workspace_bytes, int* pooled_out,
                            void* stream_) {
    DN_REQUIRE(x && w && bias && y && se_w1 && se_b1 && se_w2t && se_b2, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Cs > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), DN_ERR_UNSUPPORTED,
               "depthwise supports k in {3,5}, stride in {1,2} (got k=%d stride=%d)", k, stride);
    const int pad = (k - 1) / 2;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    DN_REQUIRE(workspace != nullptr && workspace_bytes >= dn_se_workspace_bytes(B, Ho * Wo, C), DN_ERR_WORKSPACE,
               "SE workspace too small (%zu < %zu bytes)", workspace_bytes, dn_se_workspace_bytes(B, Ho * Wo, C));
    cudaStream_t s = (cudaStream_t)stream_;
    if (pooled_out) *pooled_out = 0;
    if (C % 8 == 0 && dw_choose(H, W, C, k, stride) == DW_STREAM) {
        DwStream sp;
        DN_REQUIRE(dw_stream_plan(H, W, C, k, stride, &sp), DN_ERR_UNSUPPORTED, "no stream plan");
        CUtensorMap tm;
        int rc = dw_stream_make_tmap(&tm, x, B, H, W, C, k, sp);
        if (rc) return rc;
        DwPool pool{(float*)workspace, se_max_pool_slots(), 0, 0};
        rc = dwconv_stream_launch(tm, sp, w, bias, y, B, H, W, C, k, act, s, &pool);
        if (rc) return rc;
        if (pooled_out) *pooled_out = pool.parts > 0;
        return se_inplace_pooled(y, se_w1, se_b1, se_w2t, se_b2, B, Ho * Wo, C, Cs, workspace, workspace_bytes, pool.parts,
                                 pool.slots, Ho, s);
    }
    if (C % 8 == 0 && dw_choose(H, W, C, k, stride) == DW_STREAM2) {
        DwStream sp;
        int tw = 0;
        DN_REQUIRE(dw_stream2_plan(H, W, C, k, &sp, &tw), DN_ERR_UNSUPPORTED, "no stride-2 stream plan");
        CUtensorMap tm;
        int rc = dw_stream2_make_tmap(&tm, x, B, H, W, C, k, sp);
        if (rc) return rc;
        DwPool pool{(float*)workspace, se_max_pool_slots(), 0, 0};
        rc = dwconv_stream2_launch(tm, sp, tw, w, bias, y, B, H, W, C, k, act, s, &pool);
        if (rc) return rc;
        if (pooled_out) *pooled_out = pool.parts > 0;
        return se_inplace_pooled(y, se_w1, se_b1, se_w2t, se_b2, B, Ho * Wo, C, Cs, workspace, workspace_bytes, pool.parts,
                                 pool.slots, Ho, s);
    }
    int rc = dn_dwconv(x, w, bias, y, B, H, W, C, k, stride, act, stream_);
    if (rc) return rc;
    return dn_se_inplace(y, se_w1, se_b1, se_w2t, se_b2, B, Ho * Wo, C, Cs, workspace, workspace_bytes, stream_);
}

extern "C" int dn_stem_conv(const float* images, const float* w, const float* bias, const float* mean3_host,
                            const float* std3_host, void* y, int B, int H, int W, int Cout, int act, void* stream_) {
    DN_REQUIRE(images && w && bias && y && mean3_host && std3_host, DN_ERR_INVALID, "NULL pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(Cout == 16 || Cout == 32, DN_ERR_UNSUPPORTED, "stem supports 16 or 32 output channels (got %d)", Cout);
    if (stem_can_tma(images, W) && stem_norm_ok(std3_host)) {          // TMA-tiled kernels; the direct kernel below covers unaligned rows
        if (stem_tc_enabled())                                          // tensor-core form (default); DN_STEM=simt: fp32 SIMT tiles
            return stem_tc_launch(images, w, bias, mean3_host, std3_host, y, B, H, W, Cout, act, (cudaStream_t)stream_);
        CUtensorMap tm;
        int rc = stem_make_tmap(&tm, images, B, H, W);
        if (rc) return rc;
        return stem_tma_launch(tm, w, bias, mean3_host, std3_host, y, B, H, W, Cout, act, (cudaStream_t)stream_);
    }
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)B * Ho * Wo;
    const unsigned blocks = (unsigned)((total + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream_;
    if (Cout == 16)
        launch_pdl(stem_conv_kernel<16>, blocks, 128, 0, s, images, w, bias, (uint4*)y, B, H, W, Ho, Wo, mean3_host[0],
                   mean3_host[1], mean3_host[2], std3_host[0], std3_host[1], std3_host[2], act);
    else
        launch_pdl(stem_conv_kernel<32>, blocks, 128, 0, s, images, w, bias, (uint4*)y, B, H, W, Ho, Wo, mean3_host[0],
                   mean3_host[1], mean3_host[2], std3_host[0], std3_host[1], std3_host[2], act);
    DN_CHECK_LAUNCH();
    return DN_OK;
}
