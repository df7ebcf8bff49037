// The small memory-bound operators ssd300_vgg16 needs next to the tensor-core convolutions, sm_100a:
//   dn_conv3x3_first   normalise + dense 3x3 stride-1 conv on the 3-channel image + ReLU  (vgg features[0:2], fp32 SIMT;
//                      GeneralizedRCNNTransform.normalize, transform.py:129-138, folded in like the SSDLite stem)
//   dn_im2col3x3_first the same layer's tensor-core form: normalised 27-tap rows (padded to 32) for dn_pwconv
//   dn_maxpool2d       nn.MaxPool2d(k, s, p, ceil_mode) on NHWC 16-bit activations (vgg features; ceil_mode patched in at
//                      ssd_vgg16.py:36-37; the 3x3 s1 p1 "pool5" of ssd_vgg16.py:84)
//   dn_l2norm_scale    scale_weight * F.normalize(x) over the channels (ssd_vgg16.py:98-100)
#include <float.h>

#include "common.cuh"

namespace dn {

// thread = two horizontally adjacent output pixels x 32 output channels (blockIdx.y selects the channel half, so every lane
// of a block reads the same filter taps: broadcast LDS.128, one load per 8 FMAs); the 27 x Cout filter lives in shared memory
template <int COUT>
__global__ void __launch_bounds__(256)
conv3x3_first_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                     uint4* __restrict__ y, int B, int H, int W, float3 mean, float3 rstd) {
    __shared__ __align__(16) float ws[27 * COUT];
    __shared__ float bs[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) ws[i] = __ldg(w + i);
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = __ldg(bias + i);
    pdl_trigger();
    __syncthreads();
    pdl_wait();
    const int half = blockIdx.y;
    const int Wp = (W + 1) >> 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * H * Wp) return;
    const int xp = (int)(t % Wp), yy = (int)((t / Wp) % H), b = (int)(t / ((long long)Wp * H));
    const int x0 = 2 * xp;
    float acc[2][32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[0][i] = acc[1][i] = bs[half * 32 + i];
    const float mu[3] = {mean.x, mean.y, mean.z}, rs[3] = {rstd.x, rstd.y, rstd.z};
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float* plane = img + ((long long)b * 3 + ci) * H * W;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int iy = yy + kh - 1;
            float in[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ix = x0 + j - 1;
                in[j] = 0.f;                                          // zero padding of the NORMALISED image
                if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
                    in[j] = __fmul_rn(__fsub_rn(__ldg(plane + (long long)iy * W + ix), mu[ci]), rs[ci]);
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float4* wt = reinterpret_cast<const float4*>(ws + ((ci * 3 + kh) * 3 + kw) * COUT + half * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 wv = wt[q];
                    acc[0][q * 4 + 0] = fmaf(in[kw], wv.x, acc[0][q * 4 + 0]);
                    acc[0][q * 4 + 1] = fmaf(in[kw], wv.y, acc[0][q * 4 + 1]);
                    acc[0][q * 4 + 2] = fmaf(in[kw], wv.z, acc[0][q * 4 + 2]);
                    acc[0][q * 4 + 3] = fmaf(in[kw], wv.w, acc[0][q * 4 + 3]);
                    acc[1][q * 4 + 0] = fmaf(in[kw + 1], wv.x, acc[1][q * 4 + 0]);
                    acc[1][q * 4 + 1] = fmaf(in[kw + 1], wv.y, acc[1][q * 4 + 1]);
                    acc[1][q * 4 + 2] = fmaf(in[kw + 1], wv.z, acc[1][q * 4 + 2]);
                    acc[1][q * 4 + 3] = fmaf(in[kw + 1], wv.w, acc[1][q * 4 + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if (x0 + p >= W) break;
        const long long pix = ((long long)b * H + yy) * W + x0 + p;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(acc[p][h * 8 + i], 0.f);
            y[(pix * COUT + half * 32) / 8 + h] = pack8(f);
        }
    }
}

// im2col of the normalised 3-channel image for the tensor-core form of the first convolution: thread = one pixel, writes
// the 27 taps ((ci*3+kh)*3+kw major, zero padding) + 5 zeros as one 64-byte row of 16-bit values
__global__ void __launch_bounds__(256)
im2col3x3_first_kernel(const float* __restrict__ img, uint4* __restrict__ cols, int B, int H, int W, float3 mean, float3 rstd) {
    pdl_trigger();
    pdl_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * H * W) return;
    const int x = (int)(t % W), yy = (int)((t / W) % H), b = (int)(t / ((long long)W * H));
    const float mu[3] = {mean.x, mean.y, mean.z}, rs[3] = {rstd.x, rstd.y, rstd.z};
    float v[32];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float* plane = img + ((long long)b * 3 + ci) * H * W;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int iy = yy + kh - 1;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ix = x + kw - 1;
                float f = 0.f;
                if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
                    f = __fmul_rn(__fsub_rn(__ldg(plane + (long long)iy * W + ix), mu[ci]), rs[ci]);
                v[(ci * 3 + kh) * 3 + kw] = f;
            }
        }
    }
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) cols[t * 4 + h] = pack8(v + h * 8);
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
    const long long threads = (long long)B * H * ((W + 1) / 2);
    launch_pdl(conv3x3_first_kernel<64>, dim3((unsigned)ceil_div<long long>(threads, 256), Cout / 32), 256, 0, (cudaStream_t)stream_,
               images, w, bias, (uint4*)y, B, H, W, mean, rstd);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

extern "C" int dn_im2col3x3_first(const float* images, const float* mean3_host, const float* std3_host, void* cols, int B, int H,
                                  int W, void* stream_) {
    DN_REQUIRE(images && mean3_host && std3_host && cols, DN_ERR_INVALID, "NULL argument");
    DN_REQUIRE(B > 0 && H > 0 && W > 0, DN_ERR_INVALID, "bad shape");
    const float3 mean = make_float3(mean3_host[0], mean3_host[1], mean3_host[2]);
    const float3 rstd = make_float3(1.f / std3_host[0], 1.f / std3_host[1], 1.f / std3_host[2]);
    const long long threads = (long long)B * H * W;
    launch_pdl(im2col3x3_first_kernel, (unsigned)ceil_div<long long>(threads, 256), 256, 0, (cudaStream_t)stream_, images,
               (uint4*)cols, B, H, W, mean, rstd);
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
