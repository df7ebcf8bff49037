// Squeeze-Excitation for sm_100a, applied in place:  x *= hardsigmoid(fc2(relu(fc1(mean_hw(x))))).
// Reference: SqueezeExcitation, demonet/models/mobilenetv3.py:22-40.
//
// Four launches, each shaped by what bounds it:
//   se_pool_kernel   HBM-bound read of x: grid (pixel chunks, images), thread = 8 channels x a pixel slice;
//                    writes per-chunk channel sums (fixed summation order -> deterministic, no atomics)
//   se_fc1 / se_fc2  latency-bound: the two tiny fully-connected layers, 16 images per group of 8 CTAs that split
//                    the rows of fc1 / the channels of fc2, so that the fp32 weights are read once per group and
//                    no thread issues more than 24 (independent) weight loads
//   se_scale_kernel  HBM-bound read-modify-write of x: same thread <-> channel mapping, the 8 scales of a thread
//                    live in registers
// w1: fp32 [Cs][C]; w2t: fp32 [Cs][C] (fc2 weight TRANSPOSED so that threads read it coalesced).
#include <algorithm>

#include <cooperative_groups.h>

#include "common.cuh"

namespace dn {

constexpr int SE_POOL_THREADS = 256;
constexpr int SE_MAX_CHUNKS = 16;
constexpr int SE_FC_THREADS = 512;
constexpr int SE_NC = 8;                    // CTAs that share the fc work of one image group
constexpr int SE_GI = 16;                   // images per group
constexpr int SE_LD = 24;                   // weight loads a thread keeps in flight
constexpr int SE_SCALE_PX = 8;              // pixels per thread of the scale kernel

struct SePlan {
    int CV;                 // 16-byte channel vectors per pixel
    int rows;               // pixel slices per CTA (threads = CV * rows <= 256)
    int pool_chunks;        // pixel chunks per image of the pooling kernel
    int pool_px;            // pixels per pooling chunk
    int scale_chunks, scale_px;
};

static SePlan se_plan(int B, int HW, int C) {
    SePlan p;
    p.CV = C >> 3;
    p.rows = SE_POOL_THREADS / p.CV;
    if (p.rows > HW) p.rows = HW;
    // pooling: about eight 256-thread CTAs per SM across the batch, at least 4 pixels per thread
    int want = ceil_div(8 * sm_count(), B);
    const int most = HW / (p.rows * 4);
    if (want > most) want = most;
    if (want > SE_MAX_CHUNKS) want = SE_MAX_CHUNKS;
    if (want < 1) want = 1;
    p.pool_px = ceil_div(HW, want);
    p.pool_chunks = ceil_div(HW, p.pool_px);
    p.scale_px = p.rows * SE_SCALE_PX;
    p.scale_chunks = ceil_div(HW, p.scale_px);
    return p;
}

// partial[b][chunk][c] = sum over the chunk's pixels of x[b][p][c]
__global__ void __launch_bounds__(SE_POOL_THREADS)
se_pool_kernel(const uint4* __restrict__ x, float* __restrict__ partial, int HW, int C, int CV, int rows, int chunk_px) {
    extern __shared__ float s_part[];           // [rows][C]
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
    const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
    if (r < rows) {
        const uint4* xb = x + (long long)b * HW * CV + cv;
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int p = p0 + r;
        for (; p + 3 * rows < p1; p += 4 * rows) {          // four independent loads in flight per thread
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(xb + (long long)(p + u * rows) * CV);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int q = 0; q < 8; ++q) s[q] += f[q];
            }
        }
        for (; p < p1; p += rows) {
            float f[8];
            unpack8(__ldg(xb + (long long)p * CV), f);
#pragma unroll
            for (int q = 0; q < 8; ++q) s[q] += f[q];
        }
        float4* dst = reinterpret_cast<float4*>(s_part + r * C + cv * 8);
        dst[0] = make_float4(s[0], s[1], s[2], s[3]);
        dst[1] = make_float4(s[4], s[5], s[6], s[7]);
    }
    __syncthreads();
    float* out = partial + ((long long)b * gridDim.x + chunk) * C;
    for (int c = threadIdx.x; c < C; c += SE_POOL_THREADS) {
        float s = 0.f;
        for (int i = 0; i < rows; ++i) s += s_part[i * C + c];
        out[c] = s;
    }
}

// The two fully-connected layers as two launches over the same grid (SE_NC, image groups of SE_GI): CTA (r, grp) of
// se_fc1_kernel computes hidden rows [r*Cs/NC, ...) of its 16 images, CTA (r, grp) of se_fc2_kernel the scales of
// channels [r*C/NC, ...), so each weight matrix is read once per image group and no thread has more than SE_LD
// independent weight loads to issue.  hidden: fp32 [groups][Cs][SE_GI].
// fc1 of one CTA: hidden rows [rank * Cs / NC, ...) of the 16 images from b0 on, written to `hidden` ([Cs][SE_GI], global
// memory in the two-launch form, the CTA's own shared memory in the cluster form)
__device__ __forceinline__ void se_fc1_body(const float* __restrict__ partial, const float* __restrict__ w1,
                                            const float* __restrict__ b1, float* pooled, float* hidden, int rank, int b0, int B,
                                            int HW, int C, int Cs, int chunks, int dw_parts, int dw_rows) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // pooled means: thread = 4 channels of one image, the chunk sums of a thread are all in flight together
    const int C4 = C >> 2;
    for (int i = tid; i < SE_GI * C4; i += SE_FC_THREADS) {
        const int g = i / C4, c4 = i - g * C4;
        const int b = min(b0 + g, B - 1);
        const float4* src = reinterpret_cast<const float4*>(partial + (long long)b * chunks * C) + c4;
        // sums left by the depthwise row stream (dw_parts > 0): `chunks` is the slot stride, the image holds one slot per
        // CTA share that touched its rows [b * dw_rows, (b + 1) * dw_rows) of the B * dw_rows row stream
        int nch = chunks;
        if (dw_parts > 0) {
            const long long T = (long long)B * dw_rows;
            const long long first = (((long long)b * dw_rows + 1) * dw_parts + T - 1) / T - 1;
            const long long last = (((long long)(b + 1) * dw_rows) * dw_parts + T - 1) / T - 1;
            nch = (int)(last - first + 1);
        }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k0 = 0; k0 < SE_MAX_CHUNKS; k0 += 8) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[k] = (k0 + k < nch) ? __ldg(src + (long long)(k0 + k) * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) a.x += v[k].x, a.y += v[k].y, a.z += v[k].z, a.w += v[k].w;
        }
        const float hw = (float)HW;
        reinterpret_cast<float4*>(pooled)[i] = make_float4(__fdiv_rn(a.x, hw), __fdiv_rn(a.y, hw), __fdiv_rn(a.z, hw), __fdiv_rn(a.w, hw));
    }
    __syncthreads();

    // fc1: work item = (row, half of the images); lanes stride the C inputs
    const int rows_per = ceil_div(Cs, SE_NC);
    const int j_begin = rank * rows_per, j_end = min(Cs, j_begin + rows_per);
    constexpr int GH = SE_GI / 2;
    for (int item = warp; item < 2 * (j_end - j_begin); item += SE_FC_THREADS / 32) {
        const int j = j_begin + (item >> 1), g0 = (item & 1) * GH;
        const float* wr = w1 + (long long)j * C;
        float s[GH];
#pragma unroll
        for (int g = 0; g < GH; ++g) s[g] = 0.f;
        for (int cb = 0; cb < C; cb += SE_LD * 32) {          // one pass for C <= 768: every load of the row in flight
            float wv[SE_LD];
#pragma unroll
            for (int u = 0; u < SE_LD; ++u) {
                const int c = cb + u * 32 + lane;
                wv[u] = (c < C) ? __ldg(wr + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < SE_LD; ++u) {
                const int c = min(cb + u * 32 + lane, C - 1);
#pragma unroll
                for (int g = 0; g < GH; ++g) s[g] = fmaf(wv[u], pooled[(g0 + g) * C + c], s[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < GH; ++g) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
        }
        const float bv = __ldg(b1 + j);
        if (lane == 0) {
            float* dst = hidden + j * SE_GI + g0;
#pragma unroll
            for (int g = 0; g < GH; ++g) dst[g] = fmaxf(s[g] + bv, 0.f);
        }
    }
}

__global__ void __launch_bounds__(SE_FC_THREADS)
se_fc1_kernel(const float* __restrict__ partial, const float* __restrict__ w1, const float* __restrict__ b1,
              float* __restrict__ hidden_g, int B, int HW, int C, int Cs, int chunks, int dw_parts, int dw_rows) {
    extern __shared__ __align__(16) float s_fc[];
    pdl_trigger();
    pdl_wait();
    se_fc1_body(partial, w1, b1, s_fc, hidden_g + (long long)blockIdx.y * Cs * SE_GI, blockIdx.x, blockIdx.y * SE_GI, B, HW, C, Cs,
                chunks, dw_parts, dw_rows);
}

// fc2 + hardsigmoid of one CTA: the scales of channels [rank * C / NC, ...) of the 16 images from b0 on; `hidden` is the
// complete [Cs][SE_GI] block in shared memory, `part` [SE_NC slices][SE_GI][C / SE_NC] scratch
__device__ __forceinline__ void se_fc2_body(const float* hidden, float* part, const float* __restrict__ w2t,
                                            const float* __restrict__ b2, float* __restrict__ scale, int rank, int b0, int B, int C,
                                            int Cs) {
    const int tid = threadIdx.x;
    // thread = (channel of this CTA's share, slice of the Cs inputs)
    const int cper = C / SE_NC;                 // C % 8 == 0
    const int c_begin = rank * cper;
    const int jper = ceil_div(Cs, SE_NC);
    for (int t = tid; t < SE_NC * cper; t += SE_FC_THREADS) {
        const int js = t / cper, cl = t - js * cper;
        const int j0 = js * jper, j1 = min(Cs, j0 + jper);
        float s[SE_GI];
#pragma unroll
        for (int g = 0; g < SE_GI; ++g) s[g] = 0.f;
        for (int jb = j0; jb < j1; jb += SE_LD) {
            float wv[SE_LD];
#pragma unroll
            for (int u = 0; u < SE_LD; ++u) wv[u] = (jb + u < j1) ? __ldg(w2t + (long long)(jb + u) * C + c_begin + cl) : 0.f;
#pragma unroll
            for (int u = 0; u < SE_LD; ++u) {
                if (jb + u < j1) {
                    const float4* h = reinterpret_cast<const float4*>(hidden + (jb + u) * SE_GI);
#pragma unroll
                    for (int g4 = 0; g4 < SE_GI / 4; ++g4) {
                        const float4 hv = h[g4];
                        s[g4 * 4 + 0] = fmaf(wv[u], hv.x, s[g4 * 4 + 0]);
                        s[g4 * 4 + 1] = fmaf(wv[u], hv.y, s[g4 * 4 + 1]);
                        s[g4 * 4 + 2] = fmaf(wv[u], hv.z, s[g4 * 4 + 2]);
                        s[g4 * 4 + 3] = fmaf(wv[u], hv.w, s[g4 * 4 + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < SE_GI; ++g) part[(js * SE_GI + g) * cper + cl] = s[g];
    }
    __syncthreads();
    for (int i = tid; i < SE_GI * cper; i += SE_FC_THREADS) {
        const int g = i / cper, cl = i - g * cper;
        if (b0 + g >= B) continue;
        float s = __ldg(b2 + c_begin + cl);
        for (int js = 0; js < SE_NC; ++js) s += part[(js * SE_GI + g) * cper + cl];
        // hardsigmoid(x) = relu6(x + 3) / 6
        scale[(long long)(b0 + g) * C + c_begin + cl] = __fdiv_rn(fminf(fmaxf(s + 3.f, 0.f), 6.f), 6.f);
    }
}

__global__ void __launch_bounds__(SE_FC_THREADS)
se_fc2_kernel(const float* __restrict__ hidden_g, const float* __restrict__ w2t, const float* __restrict__ b2,
              float* __restrict__ scale, int B, int C, int Cs) {
    extern __shared__ __align__(16) float s_fc[];
    float* hidden = s_fc;                       // [Cs][SE_GI]
    pdl_trigger();
    pdl_wait();
    {
        const float4* src = reinterpret_cast<const float4*>(hidden_g + (long long)blockIdx.y * Cs * SE_GI);
        for (int i = threadIdx.x; i < Cs * SE_GI / 4; i += SE_FC_THREADS) reinterpret_cast<float4*>(hidden)[i] = __ldg(src + i);
    }
    __syncthreads();
    se_fc2_body(hidden, hidden + Cs * SE_GI, w2t, b2, scale, blockIdx.x, blockIdx.y * SE_GI, B, C, Cs);
}

// Both layers in ONE launch: the SE_NC CTAs of an image group form a thread-block cluster.  Each computes its rows of fc1
// into its own shared memory, the cluster synchronises, every CTA pulls the other seven CTAs' rows through distributed
// shared memory (<= 11 KB) and goes on with its channels of fc2 -- the hidden activations never visit global memory and
// the launch gap between the two layers (8 of the 16 SE launches of a V3 forward) is gone.  Same arithmetic in the same
// order as the two-launch form: bit-identical scales.
__global__ void __cluster_dims__(SE_NC, 1, 1) __launch_bounds__(SE_FC_THREADS)
se_fc_cluster_kernel(const float* __restrict__ partial, const float* __restrict__ w1, const float* __restrict__ b1,
                     const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ scale, int B, int HW, int C,
                     int Cs, int chunks, int dw_parts, int dw_rows) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) float s_fc[];
    float* pooled = s_fc;                       // [SE_GI][C]; fc2's `part` scratch afterwards
    float* hidden = s_fc + SE_GI * C;           // [Cs][SE_GI]
    cg::cluster_group cluster = cg::this_cluster();
    pdl_trigger();
    pdl_wait();
    const int rank = (int)cluster.block_rank();
    const int b0 = blockIdx.y * SE_GI;
    se_fc1_body(partial, w1, b1, pooled, hidden, rank, b0, B, HW, C, Cs, chunks, dw_parts, dw_rows);
    cluster.sync();
    const int rows_per = ceil_div(Cs, SE_NC);
    for (int i = threadIdx.x; i < Cs * SE_GI / 4; i += SE_FC_THREADS) {
        const int owner = (i / (SE_GI / 4)) / rows_per;
        if (owner != rank) reinterpret_cast<float4*>(hidden)[i] = reinterpret_cast<const float4*>(cluster.map_shared_rank(hidden, owner))[i];
    }
    cluster.sync();                             // every remote read is done (and the pulled rows are visible CTA-wide)
    se_fc2_body(hidden, pooled, w2t, b2, scale, rank, b0, B, C, Cs);
}

// x[b][p][c] *= scale[b][c]
__global__ void __launch_bounds__(SE_POOL_THREADS)
se_scale_kernel(uint4* __restrict__ x, const float* __restrict__ scale, int HW, int C, int CV, int rows, int chunk_px) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
    if (r >= rows) return;
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + (long long)b * C + cv * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + (long long)b * C + cv * 8) + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint4* xb = x + (long long)b * HW * CV + cv;
    const int p0 = blockIdx.x * chunk_px + r;
    uint4 v[SE_SCALE_PX];
#pragma unroll
    for (int u = 0; u < SE_SCALE_PX; ++u) {
        const int p = p0 + u * rows;
        if (p < HW) v[u] = xb[(long long)p * CV];
    }
#pragma unroll
    for (int u = 0; u < SE_SCALE_PX; ++u) {
        const int p = p0 + u * rows;
        if (p < HW) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] *= sc[q];
            xb[(long long)p * CV] = pack8(f);
        }
    }
}

}  // namespace dn

using namespace dn;

extern "C" size_t dn_se_workspace_bytes(int B, int HW, int C) {
    (void)HW;
    if (B <= 0 || C <= 0) return 0;
    // per-chunk channel sums, scales, hidden activations (Cs <= C) of ceil(B / SE_GI) image groups
    return ((size_t)B * (SE_MAX_CHUNKS + 1) * C + (size_t)ceil_div(B, SE_GI) * SE_GI * C) * sizeof(float);
}

extern "C" int dn_se_inplace(void* x, const float* w1, const float* b1, const float* w2t, const float* b2, int B, int HW,
                             int C, int Cs, void* workspace, size_t workspace_bytes, void* stream_) {
    return dn::se_inplace_pooled(x, w1, b1, w2t, b2, B, HW, C, Cs, workspace, workspace_bytes, 0, 0, 0, (cudaStream_t)stream_, true);
}

namespace dn {

int se_max_pool_slots() { return SE_MAX_CHUNKS; }

// DN_SE_CLUSTER=1: fc1 + fc2 as one cluster launch (read per call).  Bit-identical to the two-launch form and NOT the
// default: with two batches in flight the eight co-scheduled 512-thread CTAs of a cluster wait for eight free SM slots at
// once, and the step is 0.4 % slower (79.4k vs 79.7k img/s, two A/B pairs on one box) although eight launches are gone.
static bool se_fc_cluster() {
    const char* e = getenv("DN_SE_CLUSTER");
    return e && e[0] == '1';
}

// dw_parts > 0: the channel sums are already in the workspace, written by the depthwise row stream that produced x
// (dwconv_stream.cu, POOL) as [B][dw_slots][C] over dw_parts CTA shares of a B * dw_rows row stream; the pooling pass is
// skipped.
// apply == false: stop after fc2 -- the [B][C] scales stay in the workspace (se_scales_ptr) for the project GEMM that
// applies them to its A operand in shared memory (pwconv_tc.cu, SCALE_A); x is not touched.
float* se_scales_ptr(void* workspace, int B, int C) { return (float*)workspace + (size_t)B * SE_MAX_CHUNKS * C; }

int se_inplace_pooled(void* x, const float* w1, const float* b1, const float* w2t, const float* b2, int B, int HW, int C, int Cs,
                      void* workspace, size_t workspace_bytes, int dw_parts, int dw_slots, int dw_rows, cudaStream_t s, bool apply) {
    DN_REQUIRE(x && w1 && b1 && w2t && b2, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && HW > 0 && C > 0 && Cs > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(C % 8 == 0 && C / 8 <= SE_POOL_THREADS && C <= 2048, DN_ERR_UNSUPPORTED,
               "SE channels must be a multiple of 8 and <= 2048");
    DN_REQUIRE(workspace != nullptr && workspace_bytes >= dn_se_workspace_bytes(B, HW, C), DN_ERR_WORKSPACE,
               "SE workspace too small (%zu < %zu bytes)", workspace_bytes, dn_se_workspace_bytes(B, HW, C));
    DN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, DN_ERR_INVALID, "SE workspace must be 16-byte aligned");
    DN_REQUIRE(dw_parts == 0 || (dw_slots >= 1 && dw_slots <= SE_MAX_CHUNKS && dw_rows > 0), DN_ERR_INVALID, "bad pooled layout");
    const SePlan p = se_plan(B, HW, C);
    float* partial = (float*)workspace;
    float* scale = partial + (size_t)B * SE_MAX_CHUNKS * C;
    const size_t pool_smem = (size_t)p.rows * C * sizeof(float);
    const int groups = ceil_div(B, SE_GI);
    float* hidden = scale + (size_t)B * C;
    const size_t fc1_smem = (size_t)SE_GI * C * sizeof(float);
    const size_t fc2_smem = ((size_t)SE_GI * Cs + (size_t)SE_GI * C) * sizeof(float);
    DN_REQUIRE(pool_smem <= 48 * 1024 && fc2_smem <= 160 * 1024, DN_ERR_UNSUPPORTED, "SE block too large for shared memory");
    static SmemOptIn optin_fc1, optin_fc2;
    DN_CHECK_CUDA(optin_fc1.ensure(se_fc1_kernel, fc2_smem, 160 * 1024));
    DN_CHECK_CUDA(optin_fc2.ensure(se_fc2_kernel, fc2_smem, 160 * 1024));
    if (dw_parts == 0) {
        launch_pdl(se_pool_kernel, dim3(p.pool_chunks, B), SE_POOL_THREADS, pool_smem, s, (const uint4*)x, partial, HW, C, p.CV,
                   p.rows, p.pool_px);
        DN_CHECK_LAUNCH();
    }
    if (se_fc_cluster()) {
        static SmemOptIn optin_fcc;
        DN_CHECK_CUDA(optin_fcc.ensure(se_fc_cluster_kernel, fc2_smem, 160 * 1024));
        launch_pdl(se_fc_cluster_kernel, dim3(SE_NC, groups), SE_FC_THREADS, fc2_smem, s, (const float*)partial, w1, b1, w2t, b2, scale,
                   B, HW, C, Cs, dw_parts ? dw_slots : p.pool_chunks, dw_parts, dw_rows);
        DN_CHECK_LAUNCH();
    } else {
        launch_pdl(se_fc1_kernel, dim3(SE_NC, groups), SE_FC_THREADS, fc1_smem, s, (const float*)partial, w1, b1, hidden, B, HW, C, Cs,
                   dw_parts ? dw_slots : p.pool_chunks, dw_parts, dw_rows);
        DN_CHECK_LAUNCH();
        launch_pdl(se_fc2_kernel, dim3(SE_NC, groups), SE_FC_THREADS, fc2_smem, s, (const float*)hidden, w2t, b2, scale, B, C, Cs);
        DN_CHECK_LAUNCH();
    }
    if (!apply) return DN_OK;
    launch_pdl(se_scale_kernel, dim3(p.scale_chunks, B), SE_POOL_THREADS, 0, s, (uint4*)x, (const float*)scale, HW, C, p.CV, p.rows,
               p.scale_px);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

}  // namespace dn
