// The small memory-bound operators ssd300_vgg16 needs next to the tensor-core convolutions, sm_100a:
//   dn_conv3x3_first   normalise + dense 3x3 stride-1 conv on the 3-channel image + ReLU  (vgg features[0:2];
//                      GeneralizedRCNNTransform.normalize, transform.py:129-138, folded in like the SSDLite stem)
//   dn_maxpool2d       nn.MaxPool2d(k, s, p, ceil_mode) on NHWC 16-bit activations (vgg features; ceil_mode patched in at
//                      ssd_vgg16.py:36-37; the 3x3 s1 p1 "pool5" of ssd_vgg16.py:84)
//   dn_l2norm_scale    scale_weight * F.normalize(x) over the channels (ssd_vgg16.py:98-100)
#include <float.h>

#include "common.cuh"

namespace dn {

// thread = one output pixel x 16 output channels; taps of the 27 x Cout filter in shared memory
template <int COUT>
__global__ void __launch_bounds__(256)
conv3x3_first_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                     uint4* __restrict__ y, int B, int H, int W, float3 mean, float3 rstd) {
    __shared__ float ws[27 * COUT];
    __shared__ float bs[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) ws[i] = __ldg(w + i);
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = __ldg(bias + i);
    pdl_trigger();
    __syncthreads();
    pdl_wait();
    constexpr int G = COUT / 16;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)B * H * W;
    if (t >= npix * G) return;
    const int grp = (int)(t % G);
    const long long pix = t / G;
    const int x = (int)(pix % W), yy = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = bs[grp * 16 + i];
    const float mu[3] = {mean.x, mean.y, mean.z}, rs[3] = {rstd.x, rstd.y, rstd.z};
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float* plane = img + ((long long)b * 3 + ci) * H * W;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int iy = yy + kh - 1;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ix = x + kw - 1;
                float v = 0.f;                                        // zero padding of the NORMALISED image
                if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
                    v = __fmul_rn(__fsub_rn(__ldg(plane + (long long)iy * W + ix), mu[ci]), rs[ci]);
                const float* wt = ws + ((ci * 3 + kh) * 3 + kw) * COUT + grp * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(v, wt[i], acc[i]);
            }
        }
    }
    float f[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaxf(acc[h * 8 + i], 0.f);
        y[(pix * COUT + grp * 16) / 8 + h] = pack8(f);
    }
}

// thread = one output pixel x 8 channels
__global__ void __launch_bounds__(256)
maxpool_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int C8, int Ho, int Wo, int k, int s, int p) {
    pdl_trigger();
    pdl_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Ho * Wo * C8;
    if (t >= total) return;
    const int c = (int)(t % C8);
    const long long pix = t / C8;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -FLT_MAX;
    for (int kh = 0; kh < k; ++kh) {
        const int iy = oy * s - p + kh;
        if ((unsigned)iy >= (unsigned)H) continue;
        for (int kw = 0; kw < k; ++kw) {
            const int ix = ox * s - p + kw;
            if ((unsigned)ix >= (unsigned)W) continue;
            float f[8];
            unpack8(__ldg(x + (((long long)b * H + iy) * W + ix) * C8 + c), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
        }
    }
    y[t] = pack8(m);
}

// one warp per pixel: y[c] = x[c] / max(||x||_2, eps) * scale[c]
__global__ void __launch_bounds__(256)
l2norm_scale_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, uint4* __restrict__ y, long long npix, int C8, float eps) {
    pdl_trigger();
    pdl_wait();
    const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pix >= npix) return;
    float ss = 0.f;
    for (int c = lane; c < C8; c += 32) {
        float f[8];
        unpack8(__ldg(x + pix * C8 + c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = __fdiv_rn(1.f, fmaxf(sqrtf(ss), eps));
    for (int c = lane; c < C8; c += 32) {
        float f[8];
        unpack8(__ldg(x + pix * C8 + c), f);
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c * 8) + 1);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = f[i] * inv * sc[i];
        y[pix * C8 + c] = pack8(f);
    }
}

}  // namespace dn

using namespace dn;

extern "C" int dn_conv3x3_first(const float* images, const float* w, const float* bias, const float* mean3_host,
                                const float* std3_host, void* y, int B, int H, int W, int Cout, void* stream_) {
    DN_REQUIRE(images && w && bias && mean3_host && std3_host && y, DN_ERR_INVALID, "NULL argument");
    DN_REQUIRE(B > 0 && H > 0 && W > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(Cout == 64, DN_ERR_UNSUPPORTED, "first convolution is built for 64 output channels (VGG), got %d", Cout);
    const float3 mean = make_float3(mean3_host[0], mean3_host[1], mean3_host[2]);
    // (x - mean) / std with the division replaced by the reciprocal only where it is exact enough: the reference's
    // std = 1 / 255 (ssd_vgg16.py:199) -> rstd = 255 up to the rounding of 1 / 255 itself
    const float3 rstd = make_float3(1.f / std3_host[0], 1.f / std3_host[1], 1.f / std3_host[2]);
    const long long threads = (long long)B * H * W * (Cout / 16);
    launch_pdl(conv3x3_first_kernel<64>, (unsigned)ceil_div<long long>(threads, 256), 256, 0, (cudaStream_t)stream_, images, w, bias,
               (uint4*)y, B, H, W, mean, rstd);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

extern "C" int dn_maxpool2d(const void* x, void* y, int B, int H, int W, int C, int k, int stride, int pad, int ceil_mode,
                            void* stream_) {
    DN_REQUIRE(x && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && k > 0 && stride > 0 && pad >= 0 && 2 * pad <= k, DN_ERR_INVALID,
               "bad max-pool shape");
    auto out = [&](int n) {
        int o = ceil_mode ? (n + 2 * pad - k + stride - 1) / stride + 1 : (n + 2 * pad - k) / stride + 1;
        if (ceil_mode && (o - 1) * stride >= n + pad) --o;          // the last window must start inside the input (torch)
        return o;
    };
    const int Ho = out(H), Wo = out(W);
    const long long total = (long long)B * Ho * Wo * (C / 8);
    launch_pdl(maxpool_kernel, (unsigned)ceil_div<long long>(total, 256), 256, 0, (cudaStream_t)stream_, (const uint4*)x, (uint4*)y, B,
               H, W, C / 8, Ho, Wo, k, stride, pad);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

extern "C" int dn_l2norm_scale(const void* x, const float* scale, void* y, int64_t npix, int C, void* stream_) {
    DN_REQUIRE(x && scale && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(npix > 0 && C > 0 && C % 8 == 0, DN_ERR_INVALID, "bad shape");
    launch_pdl(l2norm_scale_kernel, (unsigned)ceil_div<long long>(npix * 32, 256), 256, 0, (cudaStream_t)stream_, (const uint4*)x, scale,
               (uint4*)y, (long long)npix, C / 8, 1e-12f);
    DN_CHECK_LAUNCH();
    return DN_OK;
}
