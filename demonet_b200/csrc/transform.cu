// Input / output transforms either side of the hot path, sm_100a:
//   * dn_resize_bilinear : fixed-size bilinear resize of one CHW image (fp32 in [0,1] or uint8) to fp32 [C,Ho,Wo]
//                          = torch.nn.functional.interpolate(mode='bilinear', align_corners=False) as called by
//                          _resize_image_and_masks (demonet/models/transform.py:27-53); a uint8 source is first
//                          converted exactly like ToTensor (x / 255 in fp32)
//   * dn_u8_to_f32       : uint8 -> fp32 / 255, the ToTensor conversion, for the uint8 ingest of the engine
//   * dn_rescale_boxes   : boxes *= (orig / resized) per image = resize_boxes (transform.py:278-292) as applied by
//                          GeneralizedRCNNTransform.postprocess (transform.py:228-247)
//   * dn_detections_to_coco : padded detections -> compact COCO result rows (coco_eval.py:76-98,162-164)
// All three are HBM-bound elementwise kernels: 16-byte accesses where the layout allows, one pass over the data.
#include <algorithm>

#include "common.cuh"

namespace dn {

__device__ __forceinline__ float px_load(const float* p, long long i) { return __ldg(p + i); }
__device__ __forceinline__ float px_load(const unsigned char* p, long long i) { return __fdiv_rn((float)__ldg(p + i), 255.f); }

// thread = one output pixel, all C channels (C <= 4); source index and weights follow ATen's
// area_pixel_compute_source_index (align_corners = false): src = scale * (dst + 0.5) - 0.5, clamped at 0
template <typename T>
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int H, int W, int Ho, int Wo, float sh,
                       float sw) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= Wo) return;
    // one rounding, like the fused multiply-add both ATen builds (CPU -mfma, CUDA fmad) contract this expression to:
    // at coordinates of a few hundred pixels a second rounding moves the weights by 1e-5
    const float fy = fmaxf(__fmaf_rn(sh, (float)oy + 0.5f, -0.5f), 0.f);
    const float fx = fmaxf(__fmaf_rn(sw, (float)ox + 0.5f, -0.5f), 0.f);
    const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
    for (int c = 0; c < C; ++c) {
        const long long base = (long long)c * H * W;
        const float p00 = px_load(src, base + (long long)y0 * W + x0), p01 = px_load(src, base + (long long)y0 * W + x1);
        const float p10 = px_load(src, base + (long long)y1 * W + x0), p11 = px_load(src, base + (long long)y1 * W + x1);
        const float top = __fadd_rn(__fmul_rn(lx0, p00), __fmul_rn(lx1, p01));
        const float bot = __fadd_rn(__fmul_rn(lx0, p10), __fmul_rn(lx1, p11));
        dst[((long long)c * Ho + oy) * Wo + ox] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    }
}

// 16 pixels per thread: one 16-byte load, four 16-byte stores
__global__ void __launch_bounds__(256)
u8_to_f32_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, long long n16, const unsigned char* __restrict__ src_tail,
                 float* __restrict__ dst_tail, int tail) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n16) {
        const uint4 v = __ldg(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 o;
            o.x = __fdiv_rn((float)(w[q] & 0xffu), 255.f);
            o.y = __fdiv_rn((float)((w[q] >> 8) & 0xffu), 255.f);
            o.z = __fdiv_rn((float)((w[q] >> 16) & 0xffu), 255.f);
            o.w = __fdiv_rn((float)(w[q] >> 24), 255.f);
            dst[i * 4 + q] = o;
        }
    }
    if (i < tail) dst_tail[i] = __fdiv_rn((float)__ldg(src_tail + i), 255.f);
}

__global__ void __launch_bounds__(256)
rescale_boxes_kernel(float4* __restrict__ boxes, const float2* __restrict__ ratio_hw, int D, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 r = __ldg(ratio_hw + i / D);          // (ratio_height, ratio_width)
    float4 b = boxes[i];
    b.x = __fmul_rn(b.x, r.y);
    b.y = __fmul_rn(b.y, r.x);
    b.z = __fmul_rn(b.z, r.y);
    b.w = __fmul_rn(b.w, r.x);
    boxes[i] = b;
}

// Detection sink: padded per-image detections -> compact COCO result rows in image order, the
// device-side form of CocoEvaluator.prepare_for_coco_detection (demonet/data/coco_eval.py:76-98) with
// convert_to_xywh (coco_eval.py:162-164).  CTA b sums counts[0..b) itself (B is a few thousand at most), so one
// launch does scan + compaction deterministically and without atomics.
__global__ void __launch_bounds__(256)
coco_rows_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const long long* __restrict__ labels,
                 const int* __restrict__ counts, const long long* __restrict__ image_ids, int B, int D,
                 long long* __restrict__ out_image_id, long long* __restrict__ out_category, float4* __restrict__ out_bbox,
                 float* __restrict__ out_score, long long* __restrict__ out_total) {
    __shared__ long long s_warp[8];
    __shared__ long long s_off;
    const int b = blockIdx.x;
    long long part = 0;
    for (int i = threadIdx.x; i < b; i += blockDim.x) part += min(max(counts[i], 0), D);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += s_warp[w];
        s_off = t;
    }
    __syncthreads();
    const long long off = s_off;
    const int n = min(max(counts[b], 0), D);
    const long long id = image_ids[b];
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        const float4 q = boxes[(long long)b * D + d];
        out_image_id[off + d] = id;
        out_category[off + d] = labels[(long long)b * D + d];
        out_bbox[off + d] = make_float4(q.x, q.y, __fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));      // xywh
        out_score[off + d] = scores[(long long)b * D + d];
    }
    if (b == B - 1 && threadIdx.x == 0) *out_total = off + n;
}

int u8_to_f32_launch(const unsigned char* src, float* dst, size_t n, cudaStream_t s) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const long long n16 = aligned ? (long long)(n / 16) : 0;
    const int tail = (int)(n - (size_t)n16 * 16);
    if (!aligned) {          // rare: plain per-byte kernel through the tail path, 2^20 elements per launch
        for (size_t off = 0; off < n; off += (1u << 20)) {
            const int cnt = (int)std::min<size_t>(1u << 20, n - off);
            u8_to_f32_kernel<<<ceil_div(cnt, 256), 256, 0, s>>>(nullptr, nullptr, 0, src + off, dst + off, cnt);
        }
    } else {
        const long long threads = std::max<long long>(n16, tail);
        u8_to_f32_kernel<<<(unsigned)ceil_div<long long>(threads, 256), 256, 0, s>>>(
            (const uint4*)src, (float4*)dst, n16, src + (size_t)n16 * 16, dst + (size_t)n16 * 16, tail);
    }
    DN_CHECK_LAUNCH();
    return DN_OK;
}

}  // namespace dn

using namespace dn;

extern "C" int dn_resize_bilinear(const void* src, int src_is_u8, int C, int H, int W, float* dst, int Ho, int Wo,
                                  void* stream_) {
    DN_REQUIRE(src && dst, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, DN_ERR_INVALID, "bad shape");
    cudaStream_t s = (cudaStream_t)stream_;
    const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;      // area_pixel_compute_scale, align_corners = false
    const dim3 grid(ceil_div(Wo, 256), Ho);
    if (src_is_u8)
        resize_bilinear_kernel<unsigned char><<<grid, 256, 0, s>>>((const unsigned char*)src, dst, C, H, W, Ho, Wo, sh, sw);
    else
        resize_bilinear_kernel<float><<<grid, 256, 0, s>>>((const float*)src, dst, C, H, W, Ho, Wo, sh, sw);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

extern "C" int dn_u8_to_f32(const uint8_t* src, float* dst, size_t n, void* stream_) {
    DN_REQUIRE(src && dst, DN_ERR_INVALID, "NULL tensor pointer");
    if (n == 0) return DN_OK;
    return u8_to_f32_launch(src, dst, n, (cudaStream_t)stream_);
}

extern "C" int dn_rescale_boxes(float* boxes, const float* ratio_hw, int B, int D, void* stream_) {
    DN_REQUIRE(boxes && ratio_hw, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && D > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(((reinterpret_cast<uintptr_t>(boxes) & 15) | (reinterpret_cast<uintptr_t>(ratio_hw) & 7)) == 0, DN_ERR_INVALID,
               "boxes must be 16-byte aligned, ratios 8-byte aligned");
    const long long n = (long long)B * D;
    rescale_boxes_kernel<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, (cudaStream_t)stream_>>>((float4*)boxes,
                                                                                                 (const float2*)ratio_hw, D, n);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

extern "C" int dn_detections_to_coco(const float* boxes, const float* scores, const int64_t* labels, const int32_t* counts,
                                     const int64_t* image_ids, int B, int D, int64_t* out_image_id, int64_t* out_category_id,
                                     float* out_bbox_xywh, float* out_score, int64_t* out_total, void* stream_) {
    DN_REQUIRE(boxes && scores && labels && counts && image_ids && out_image_id && out_category_id && out_bbox_xywh &&
                   out_score && out_total,
               DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && D > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(out_bbox_xywh)) & 15) == 0, DN_ERR_INVALID,
               "box arrays must be 16-byte aligned");
    coco_rows_kernel<<<B, 256, 0, (cudaStream_t)stream_>>>((const float4*)boxes, scores, (const long long*)labels, counts,
                                                            (const long long*)image_ids, B, D, (long long*)out_image_id,
                                                            (long long*)out_category_id, (float4*)out_bbox_xywh, out_score,
                                                            (long long*)out_total);
    DN_CHECK_LAUNCH();
    return DN_OK;
}
