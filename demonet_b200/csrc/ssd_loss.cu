// Training-side operators of the SSD detectors, sm_100a: default-box matching and the multibox loss with
// hard-negative mining, value and gradient with respect to the head outputs.
//   dn_ssd_match      box_ops.box_iou + SSDMatcher per image   generalized_ssd.py:326-335, _utils.py:227-362
//   dn_match_quality  SSDMatcher on a caller-supplied M x N quality matrix (the reference's own call signature)
//   dn_ssd_loss       SSD.compute_loss                          generalized_ssd.py:210-269, BoxCoder.encode_single _utils.py:83-127
// Index work (matches) is bit-exact against the reference's CPU results: the IoU is computed with the reference's
// operation order in round-to-nearest fp32 without contraction, arg-max ties resolve to the first maximum and, when two
// ground-truth boxes claim the same default box, to the larger ground-truth index (index_put_ order on the CPU).  The loss
// sums are accumulated in double in a fixed order (deterministic; the reference's fp32 pairwise sums agree to ~1e-6).
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace dn {

constexpr int SL_THREADS = 256;
constexpr int SL_MINE_THREADS = 512;

__device__ __forceinline__ float sl_area(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

// torchvision.ops.boxes.box_iou: inter / (area1 + area2 - inter), wh = (min(rb) - max(lt)).clamp(min=0)
__device__ __forceinline__ float sl_iou(const float4 a, const float area_a, const float4 b, const float area_b) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

struct QualityIoU {                       // quality(g, p) from the boxes
    const float4* gt;                     // this image's ground-truth boxes
    const float4* anchors;
    __device__ __forceinline__ float operator()(int g, int p) const {
        const float4 a = __ldg(gt + g), b = __ldg(anchors + p);
        return sl_iou(a, sl_area(a), b, sl_area(b));
    }
};
struct QualityMatrix {                    // quality(g, p) = q[g][p]
    const float* q;
    int P;
    __device__ __forceinline__ float operator()(int g, int p) const { return __ldg(q + (long long)g * P + p); }
};

// Matcher.__call__ with low == high == thresh (_utils.py:283-323) followed by SSDMatcher's forced matches (:350-362).
// One CTA per image.  best_p: int32 scratch of this image's M ground-truth boxes.
template <typename Q>
__device__ __forceinline__ void sl_match_image(const Q& quality, int M, int P, float thresh, int64_t* __restrict__ matches,
                                               int* __restrict__ best_p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // max over the ground truth (dim 0): first maximum
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        long long idx = -1;
        if (M > 0) {
            float best = quality(0, p);
            int bi = 0;
            for (int g = 1; g < M; ++g) {
                const float v = quality(g, p);
                if (v > best) best = v, bi = g;
            }
            idx = best < thresh ? -1 : bi;
        }
        matches[p] = idx;
    }
    // max over the default boxes (dim 1) per ground-truth box: first maximum
    for (int g = warp; g < M; g += nwarps) {
        float best = -FLT_MAX;
        int bp = 0x7fffffff;
        for (int p = lane; p < P; p += 32) {
            const float v = quality(g, p);
            if (v > best || bp == 0x7fffffff) best = v, bp = p;          // ascending p per lane: strict > keeps the first
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (op != 0x7fffffff && (bp == 0x7fffffff || ov > best || (ov == best && op < bp))) best = ov, bp = op;
        }
        if (lane == 0) best_p[g] = bp;
    }
    __syncthreads();
    // matches[highest_quality_pred_foreach_gt] = arange(M): sequential, so the larger ground-truth index wins a shared box
    if (threadIdx.x == 0)
        for (int g = 0; g < M; ++g) matches[best_p[g]] = g;
}

__global__ void __launch_bounds__(SL_THREADS)
ssd_match_kernel(const float4* __restrict__ gt_boxes, const int* __restrict__ gt_offsets, const float4* __restrict__ anchors, int P,
                 float thresh, int64_t* __restrict__ matches, int* __restrict__ best_p) {
    const int b = blockIdx.x;
    const int g0 = gt_offsets[b], M = gt_offsets[b + 1] - g0;
    QualityIoU q{gt_boxes + g0, anchors};
    sl_match_image(q, M, P, thresh, matches + (long long)b * P, best_p + g0);
}

__global__ void __launch_bounds__(SL_THREADS)
match_quality_kernel(const float* __restrict__ quality, int M, int P, float thresh, int64_t* __restrict__ matches,
                     int* __restrict__ best_p) {
    QualityMatrix q{quality, P};
    sl_match_image(q, M, P, thresh, matches, best_p);
}

// ---- loss ------------------------------------------------------------------------------------------------------------
struct BoxWeights {
    float wx, wy, ww, wh;
};

// smooth_l1_loss(beta = 1): 0.5 d^2 if |d| < 1 else |d| - 0.5
__device__ __forceinline__ float sl_smooth_l1(float d) {
    const float a = fabsf(d);
    return a < 1.f ? 0.5f * d * d : a - 0.5f;
}

// BoxCoder.encode_single (_utils.py:83-127) of one matched pair, operation order kept
__device__ __forceinline__ void sl_encode(const float4 gt, const float4 an, const BoxWeights bw, float* t) {
    const float ex_w = __fsub_rn(an.z, an.x), ex_h = __fsub_rn(an.w, an.y);
    const float ex_cx = __fadd_rn(an.x, __fmul_rn(0.5f, ex_w)), ex_cy = __fadd_rn(an.y, __fmul_rn(0.5f, ex_h));
    const float gt_w = __fsub_rn(gt.z, gt.x), gt_h = __fsub_rn(gt.w, gt.y);
    const float gt_cx = __fadd_rn(gt.x, __fmul_rn(0.5f, gt_w)), gt_cy = __fadd_rn(gt.y, __fmul_rn(0.5f, gt_h));
    t[0] = __fdiv_rn(__fmul_rn(bw.wx, __fsub_rn(gt_cx, ex_cx)), ex_w);
    t[1] = __fdiv_rn(__fmul_rn(bw.wy, __fsub_rn(gt_cy, ex_cy)), ex_h);
    t[2] = __fmul_rn(bw.ww, logf(__fdiv_rn(gt_w, ex_w)));
    t[3] = __fmul_rn(bw.wh, logf(__fdiv_rn(gt_h, ex_h)));
}

// warp per default box: cross entropy against the matched label (0 = background) and the box loss of foreground boxes.
// WITH_GRAD: second pass, writes d loss / d cls_logits and d loss / d bbox_regression (unit upstream gradient).
template <bool WITH_GRAD>
__global__ void __launch_bounds__(SL_THREADS)
ssd_anchor_loss_kernel(const float* __restrict__ cls_logits, const float4* __restrict__ bbox_regression,
                       const float4* __restrict__ anchors, const float4* __restrict__ gt_boxes,
                       const int64_t* __restrict__ gt_labels, const int* __restrict__ gt_offsets,
                       const int64_t* __restrict__ matches, int B, int P, int K, BoxWeights bw, float* __restrict__ ce,
                       int* __restrict__ ctarget, float* __restrict__ box_loss, const unsigned char* __restrict__ weight,
                       const float* __restrict__ losses, float* __restrict__ grad_cls, float4* __restrict__ grad_reg) {
    const int lane = threadIdx.x & 31;
    const long long a = (long long)blockIdx.x * (SL_THREADS / 32) + (threadIdx.x >> 5);
    if (a >= (long long)B * P) return;
    const int b = (int)(a / P), p = (int)(a - (long long)b * P);
    const long long m = matches[a];
    const int g = gt_offsets[b] + (int)(m >= 0 ? m : 0);
    const int ct = m >= 0 ? (int)gt_labels[g] : 0;
    const float* x = cls_logits + a * K;
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, __ldg(x + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += expf(__ldg(x + k) - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const bool label_ok = (unsigned)ct < (unsigned)K;
    if (!WITH_GRAD) {
        if (lane == 0) {
            // log_softmax(x)[ct] = x[ct] - max - log(sum exp(x - max)); a label outside [0, K) poisons the loss
            ce[a] = label_ok ? -((__ldg(x + ct) - mx) - logf(s)) : NAN;
            ctarget[a] = ct;
            float bl = 0.f;
            if (m >= 0) {
                float t[4];
                sl_encode(__ldg(gt_boxes + g), __ldg(anchors + p), bw, t);
                const float4 r = __ldg(bbox_regression + a);
                bl = sl_smooth_l1(r.x - t[0]) + sl_smooth_l1(r.y - t[1]) + sl_smooth_l1(r.z - t[2]) + sl_smooth_l1(r.w - t[3]);
            }
            box_loss[a] = bl;
        }
    } else {
        const float inv_n = 1.f / losses[2];
        const float wgt = (float)weight[a] * inv_n;
        if (grad_cls) {
            float* gc = grad_cls + a * K;
            const float inv_s = 1.f / s;
            for (int k = lane; k < K; k += 32) {
                float v = 0.f;
                if (wgt != 0.f) v = wgt * (expf(__ldg(x + k) - mx) * inv_s - (k == ct ? 1.f : 0.f));
                gc[k] = v;
            }
        }
        if (grad_reg && lane == 0) {
            float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m >= 0) {
                float t[4];
                sl_encode(__ldg(gt_boxes + g), __ldg(anchors + p), bw, t);
                const float4 r = __ldg(bbox_regression + a);
                const float d[4] = {r.x - t[0], r.y - t[1], r.z - t[2], r.w - t[3]};
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = (fabsf(d[j]) < 1.f ? d[j] : (d[j] > 0.f ? 1.f : -1.f)) * inv_n;
                gr = make_float4(o[0], o[1], o[2], o[3]);
            }
            grad_reg[a] = gr;
        }
    }
}

// order-preserving map float -> uint32 (larger float = larger key; -inf is the smallest key)
__device__ __forceinline__ uint32_t sl_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ double sl_block_sum(double v, double* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

// Hard-negative mining of one image (generalized_ssd.py:252-262): the ceil(ratio * #foreground) largest losses among
// negative_loss (foreground = -inf) are the background sample; rank ties go to the lower index (a stable descending sort).
// CTA per image.  Writes weight[p] = (foreground) + (background sample) and the image's partial sums.
__global__ void __launch_bounds__(SL_MINE_THREADS)
ssd_mine_kernel(const float* __restrict__ ce, const int* __restrict__ ctarget, const float* __restrict__ box_loss,
                const int64_t* __restrict__ matches, int P, float ratio, unsigned char* __restrict__ weight,
                double* __restrict__ partial) {
    extern __shared__ uint32_t sl_keys[];                  // [P]
    __shared__ int hist[256];
    __shared__ double red[SL_MINE_THREADS / 32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_remaining, s_nfg, s_nmatched;
    const int b = blockIdx.x;
    ce += (long long)b * P, ctarget += (long long)b * P, box_loss += (long long)b * P, matches += (long long)b * P;
    weight += (long long)b * P;
    if (threadIdx.x == 0) s_nfg = 0, s_nmatched = 0;
    __syncthreads();
    int nfg = 0, nm = 0;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const bool fg = ctarget[p] > 0;
        nfg += fg, nm += matches[p] >= 0;
        sl_keys[p] = sl_key(fg ? -INFINITY : ce[p]);
    }
    atomicAdd(&s_nfg, nfg);                                 // integer sums: order does not matter
    atomicAdd(&s_nmatched, nm);
    __syncthreads();
    // ranks r with float(r) < ratio * #foreground
    const float kf = ratio * (float)s_nfg;
    int k = kf > 0.f ? (kf >= (float)P ? P : (int)ceilf(kf)) : 0;
    if (k > P) k = P;
    // 4-pass radix select of the k-th largest key
    if (threadIdx.x == 0) s_prefix = 0u, s_remaining = k;
    __syncthreads();
    if (k > 0 && k < P) {
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int p = threadIdx.x; p < P; p += blockDim.x) {
                const uint32_t key = sl_keys[p];
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int rem = s_remaining, bin = 255;
                for (; bin > 0; --bin) {
                    if (hist[bin] >= rem) break;
                    rem -= hist[bin];
                }
                s_remaining = rem;                          // how many keys of this bin are still to be taken
                s_prefix = prefix | ((uint32_t)bin << shift);
            }
            __syncthreads();
        }
    }
    const uint32_t T = s_prefix;
    const int take_eq = s_remaining;
    // keys equal to the threshold: the first take_eq by index (one warp walks the row)
    if (k > 0 && k < P) {
        if (threadIdx.x < 32) {
            int taken = 0;
            for (int p0 = 0; p0 < P; p0 += 32) {
                const int p = p0 + (int)threadIdx.x;
                const bool eq = p < P && sl_keys[p] == T;
                const uint32_t bal = __ballot_sync(0xffffffffu, eq);
                if (eq) {
                    const int r = taken + __popc(bal & ((1u << threadIdx.x) - 1u));
                    if (r >= take_eq) sl_keys[p] = T - 1u;          // not sampled: push below the threshold (T > 0 here:
                                                                    // every finite loss and -inf map to keys >= 0x007fffff)
                }
                taken += __popc(bal);
            }
        }
        __syncthreads();
    }
    double cls_sum = 0.0, box_sum = 0.0;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const bool fg = ctarget[p] > 0;
        const bool bg = k >= P ? true : (k > 0 && sl_keys[p] >= T);
        const float c = ce[p];
        // the reference adds cls_loss[foreground] and cls_loss[background]: a foreground box that ranks into the
        // background sample (only when there are fewer negatives than ratio * #foreground) counts twice
        cls_sum += (fg ? (double)c : 0.0) + (bg ? (double)c : 0.0);
        box_sum += (double)box_loss[p];
        weight[p] = (unsigned char)((fg ? 1 : 0) + (bg ? 1 : 0));
    }
    cls_sum = sl_block_sum(cls_sum, red);
    box_sum = sl_block_sum(box_sum, red);
    if (threadIdx.x == 0) {
        partial[3 * b + 0] = box_sum;
        partial[3 * b + 1] = cls_sum;
        partial[3 * b + 2] = (double)s_nmatched;
    }
}

// losses = {bbox_regression, classification, N}: sums over the images in index order, N = max(1, #matched)
__global__ void ssd_loss_finish_kernel(const double* __restrict__ partial, int B, float* __restrict__ losses) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double bs = 0.0, cs = 0.0, n = 0.0;
    for (int b = 0; b < B; ++b) bs += partial[3 * b], cs += partial[3 * b + 1], n += partial[3 * b + 2];
    const float N = n < 1.0 ? 1.f : (float)n;
    losses[0] = (float)bs / N;
    losses[1] = (float)cs / N;
    losses[2] = N;
}

static inline size_t sl_align(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace dn

using namespace dn;

extern "C" size_t dn_ssd_loss_workspace_bytes(int B, int P, int G) {
    if (B < 0 || P < 0 || G < 0) return 0;
    const size_t bp = (size_t)B * P;
    // best_p [G] | ce [B,P] | ctarget [B,P] | box_loss [B,P] | weight [B,P] | partial [B,3] double
    return sl_align((size_t)(G > 0 ? G : 1) * 4) + 3 * sl_align(bp * 4) + sl_align(bp) + sl_align((size_t)B * 3 * 8) + 256;
}

extern "C" int dn_ssd_match(const float* gt_boxes, const int32_t* gt_offsets, const float* anchors, int B, int P, int G,
                            float iou_thresh, int64_t* matched_idxs, void* workspace, size_t workspace_bytes, void* stream) {
    DN_REQUIRE(B >= 0 && P > 0 && G >= 0, DN_ERR_INVALID, "dn_ssd_match: bad sizes B=%d P=%d G=%d", B, P, G);
    DN_REQUIRE(gt_offsets && anchors && matched_idxs && (G == 0 || gt_boxes), DN_ERR_INVALID, "dn_ssd_match: null pointer");
    DN_REQUIRE(workspace && workspace_bytes >= sl_align((size_t)(G > 0 ? G : 1) * 4), DN_ERR_WORKSPACE,
               "dn_ssd_match: workspace too small");
    if (B == 0) return DN_OK;
    ssd_match_kernel<<<B, SL_THREADS, 0, (cudaStream_t)stream>>>((const float4*)gt_boxes, gt_offsets, (const float4*)anchors, P,
                                                                 iou_thresh, matched_idxs, (int*)workspace);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

extern "C" int dn_match_quality(const float* quality, int M, int P, float thresh, int64_t* matches, void* workspace,
                                size_t workspace_bytes, void* stream) {
    // Matcher.__call__ raises on an empty matrix (_utils.py:298-307)
    DN_REQUIRE(M > 0, DN_ERR_INVALID, "No ground-truth boxes available for one of the images during training");
    DN_REQUIRE(P > 0, DN_ERR_INVALID, "No proposal boxes available for one of the images during training");
    DN_REQUIRE(quality && matches, DN_ERR_INVALID, "dn_match_quality: null pointer");
    DN_REQUIRE(workspace && workspace_bytes >= (size_t)M * 4, DN_ERR_WORKSPACE, "dn_match_quality: workspace too small");
    match_quality_kernel<<<1, SL_THREADS, 0, (cudaStream_t)stream>>>(quality, M, P, thresh, matches, (int*)workspace);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

extern "C" int dn_ssd_loss(const float* cls_logits, const float* bbox_regression, const float* anchors, const float* gt_boxes,
                           const int64_t* gt_labels, const int32_t* gt_offsets, const int64_t* matched_idxs, int B, int P, int K,
                           int G, float neg_to_pos_ratio, const float* box_weights4_host, float* losses, float* grad_cls,
                           float* grad_reg, void* workspace, size_t workspace_bytes, void* stream) {
    DN_REQUIRE(B > 0 && P > 0 && K > 0 && G >= 0, DN_ERR_INVALID, "dn_ssd_loss: bad sizes B=%d P=%d K=%d G=%d", B, P, K, G);
    DN_REQUIRE(cls_logits && bbox_regression && anchors && gt_offsets && matched_idxs && losses && box_weights4_host,
               DN_ERR_INVALID, "dn_ssd_loss: null pointer");
    DN_REQUIRE(G == 0 || (gt_boxes && gt_labels), DN_ERR_INVALID, "dn_ssd_loss: null ground truth");
    DN_REQUIRE((size_t)P * 4 <= 200 * 1024, DN_ERR_UNSUPPORTED, "dn_ssd_loss: %d default boxes per image exceed the mining kernel's shared memory", P);
    DN_REQUIRE(workspace && workspace_bytes >= dn_ssd_loss_workspace_bytes(B, P, G), DN_ERR_WORKSPACE,
               "dn_ssd_loss: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bp = (size_t)B * P;
    unsigned char* w = (unsigned char*)workspace + sl_align((size_t)(G > 0 ? G : 1) * 4);
    float* ce = (float*)w;
    w += sl_align(bp * 4);
    int* ctarget = (int*)w;
    w += sl_align(bp * 4);
    float* box_loss = (float*)w;
    w += sl_align(bp * 4);
    unsigned char* weight = w;
    w += sl_align(bp);
    double* partial = (double*)w;
    const BoxWeights bw{box_weights4_host[0], box_weights4_host[1], box_weights4_host[2], box_weights4_host[3]};
    const unsigned grid = (unsigned)((bp + SL_THREADS / 32 - 1) / (SL_THREADS / 32));
    ssd_anchor_loss_kernel<false><<<grid, SL_THREADS, 0, s>>>(cls_logits, (const float4*)bbox_regression, (const float4*)anchors,
                                                             (const float4*)gt_boxes, gt_labels, gt_offsets, matched_idxs, B, P, K, bw,
                                                             ce, ctarget, box_loss, nullptr, nullptr, nullptr, nullptr);
    DN_CHECK_LAUNCH();
    const size_t smem = (size_t)P * 4;
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(ssd_mine_kernel, smem));
    ssd_mine_kernel<<<B, SL_MINE_THREADS, smem, s>>>(ce, ctarget, box_loss, matched_idxs, P, neg_to_pos_ratio, weight, partial);
    DN_CHECK_LAUNCH();
    ssd_loss_finish_kernel<<<1, 32, 0, s>>>(partial, B, losses);
    DN_CHECK_LAUNCH();
    if (grad_cls || grad_reg) {
        ssd_anchor_loss_kernel<true><<<grid, SL_THREADS, 0, s>>>(cls_logits, (const float4*)bbox_regression, (const float4*)anchors,
                                                                (const float4*)gt_boxes, gt_labels, gt_offsets, matched_idxs, B, P, K,
                                                                bw, nullptr, nullptr, nullptr, weight, losses, grad_cls,
                                                                (float4*)grad_reg);
        DN_CHECK_LAUNCH();
    }
    return DN_OK;
}
