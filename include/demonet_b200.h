/* demonet_b200 -- C ABI of the B200-native (sm_100a) SSDLite inference hot path.
 *
 * The reference (zhiqwang/demonet) has NO plugin / FFI layer ("There are no extra compiled
 * components in DEMONET", README.md:18); its hot path runs entirely inside torch / torchvision
 * operators.  This header is therefore the boundary a reference maintainer would bind with
 * ctypes: each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes, no torch types; every data pointer is CALLER-OWNED DEVICE memory
 *     unless the name ends in `_host`;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); no host
 *     synchronisation, no allocation per call (workspaces are sized by a query and passed in);
 *   - return value 0 = ok, negative = dn_status; dn_last_error() gives a thread-local message;
 *   - activations are NHWC bf16 (row-major [B*H*W, C], C % 8 == 0), head outputs / scores / boxes fp32.
 */
#ifndef DEMONET_B200_H_
#define DEMONET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DN_ABI_VERSION 4          /* 3: dn_postprocess_scored, dn_engine_get_stats, activation dtype builds
                                     4: dn_ssd_match, dn_match_quality, dn_ssd_loss (training-side operators) */

typedef enum {
    DN_OK = 0,
    DN_ERR_INVALID = -1,        /* bad argument (maps to ValueError on the Python side)           */
    DN_ERR_CUDA = -2,           /* CUDA runtime / driver error (RuntimeError)                      */
    DN_ERR_UNSUPPORTED = -3,    /* shape / option the engine cannot honour (NotImplementedError)   */
    DN_ERR_WORKSPACE = -4       /* workspace too small                                             */
} dn_status;

typedef enum { DN_ACT_NONE = 0, DN_ACT_RELU = 1, DN_ACT_RELU6 = 2, DN_ACT_HSWISH = 3 } dn_act;

const char* dn_last_error(void);
int dn_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Stage-level entry points (also what the parity tests call)
 * ---------------------------------------------------------------------------------------- */

/* Depthwise k x k convolution + folded BatchNorm + activation.
 * Replaces ConvBNActivation with groups == channels: demonet/models/mobilenetv2.py:32-55 as used at
 * mobilenetv3.py:80-83, mobilenetv2.py:87, ssd_mobilenetv3.py:31-32,48-49, backbone.py:106, and the
 * biased depthwise of SeperableConv2d, box_head.py:30.
 * x: bf16 [B,H,W,C]; w: fp32 [k*k, C] (tap-major, BN folded); bias: fp32 [C]; y: bf16 [B,Ho,Wo,C],
 * Ho = (H + 2*(k/2) - k)/stride + 1.  k in {3,5}, stride in {1,2}, C % 8 == 0. */
int dn_dwconv(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int k,
              int stride, int act, void* stream);
/* Host-only: which kernel dn_dwconv runs for a layer shape and how it is tiled (no device work; usable without a GPU).
 * out8 = { impl (1 direct, 2 TMA tiles, 4 stride-1 row stream, 8 stride-2 row stream), channels per block, column blocks
 * per CTA, column strips across the width, output columns per thread, threads per CTA, ring stages, dynamic shared
 * memory bytes }; the last six are 0 for impl 1 / 2. */
int dn_dwconv_plan_info(int H, int W, int C, int k, int stride, int32_t* out8);

/* Pointwise (1x1) convolution as a GEMM on the tcgen05 tensor cores, with folded BatchNorm / bias,
 * activation and the residual add fused in the epilogue.
 * Replaces the 1x1 ConvBNActivation / Conv2d(+BN) / Conv2d(bias): mobilenetv3.py:75-77,88-89,149-152,
 * mobilenetv2.py:84,89-90,166, ssd_mobilenetv3.py:35,44-45,52-53, box_head.py:33,55-56, and the
 * `result += input` of InvertedResidual.forward (mobilenetv3.py:95-99, mobilenetv2.py:96-100).
 * x: bf16 [M,K] (M = B*H*W); w: bf16 [N,K]; bias: fp32 [N]; residual: bf16 [M,N] or NULL.
 * Output element (m,n) is written at  y + (m / hw) * out_batch_stride + (m % hw) * out_row_stride + n
 * (element units), as bf16 or, when out_fp32 != 0, fp32 -- this is how the head's
 * view/permute/reshape/cat (generalized_ssd.py:66-74, box_head.py:107-146) becomes a no-op.
 * K % 8 == 0.  impl: 0 = tcgen05/TMA kernel (product path), 1 = SIMT self-check kernel. */
int dn_pwconv(const void* x, const void* w, const float* bias, const void* residual, void* y, int M, int K,
              int N, int act, int out_fp32, int hw, int64_t out_batch_stride, int64_t out_row_stride,
              int impl, void* stream);
/* Host-only: the tile plan dn_pwconv uses for a GEMM shape (no device work; usable without a GPU).
 * out8 = { block_n (columns per output tile), n_tiles, ring stages, TMEM columns (two accumulator buffers), dynamic shared
 * memory bytes, weight-stationary (K <= 128: the weight tile stays in shared memory), pair mode (DN_PW_PAIR: 0 off,
 * 1 TMA multicast, 2 cta_group::2), resident CTAs per SM }. */
int dn_pwconv_plan_info(long long M, int K, int N, int32_t* out8);

/* Fused pointwise expand (1x1 conv + folded BN + act) -> depthwise k x k (+ folded BN + act): the expanded tensor stays
 * in shared memory.  Replaces InvertedResidual.block[0:2] (mobilenetv3.py:75-83, mobilenetv2.py:84-87) where the shape
 * is supported -- this build: K = 16 input channels, N = 64 expanded channels, 3x3 stride 2 (MobileNetV3 block 2);
 * DN_ERR_UNSUPPORTED otherwise.  x: bf16 [B,H,W,K]; w_pw: bf16 [N,K]; b_pw: fp32 [N]; w_dw: fp32 [k*k,N]; b_dw: fp32 [N];
 * y: bf16 [B,Ho,Wo,N].  Results equal dn_pwconv followed by dn_dwconv up to the fp32 summation order of the stencil. */
int dn_pwdw_fused(const void* x, const void* w_pw, const float* b_pw, const float* w_dw, const float* b_dw, void* y,
                  int B, int H, int W, int K, int N, int ksize, int stride, int act_pw, int act_dw, void* stream);

/* Fused depthwise k x k (+ folded BN + act) -> pointwise project (1x1 conv + folded BN, no activation) (+ residual = the
 * block input): the depthwise output stays in shared memory.  Replaces InvertedResidual.block[1:] and `result += input`
 * (mobilenetv3.py:80-99, mobilenetv2.py:87-100) where the shape is supported -- this build: C = N = 16 channels, 3x3
 * stride 1 (MobileNetV3 block 1); DN_ERR_UNSUPPORTED otherwise.  x: bf16 [B,H,W,C]; w_dw: fp32 [k*k,C]; b_dw: fp32 [C];
 * w_pw: bf16 [N,C]; b_pw: fp32 [N]; y: bf16 [B,H,W,N], y != x; residual != 0 adds x. */
int dn_dwpw_fused(const void* x, const float* w_dw, const float* b_dw, const void* w_pw, const float* b_pw, void* y,
                  int B, int H, int W, int C, int N, int ksize, int stride, int act_dw, int residual, void* stream);

/* Stem: input normalisation + dense 3x3 stride-2 convolution + folded BN + activation.
 * Replaces GeneralizedRCNNTransform.normalize (transform.py:129-138) followed by the first
 * ConvBNActivation (mobilenetv3.py:141-142, mobilenetv2.py:157).
 * images: fp32 NCHW [B,3,H,W] in [0,1]; w: fp32 [27, Cout] ((ci*3+kh)*3+kw major); y: bf16 NHWC. */
int dn_stem_conv(const float* images, const float* w, const float* bias, const float* mean3_host,
                 const float* std3_host, void* y, int B, int H, int W, int Cout, int act, void* stream);

/* Squeeze-Excitation applied in place: x *= hardsigmoid(fc2(relu(fc1(avgpool(x))))).
 * Replaces SqueezeExcitation.forward, mobilenetv3.py:22-40.
 * x: bf16 [B,HW,C] (in/out); w1: fp32 [Cs,C] (fc1); b1: fp32 [Cs]; w2t: fp32 [Cs,C] (fc2 weight
 * TRANSPOSED); b2: fp32 [C].  workspace: device scratch of dn_se_workspace_bytes(B, HW, C) bytes
 * (per-chunk channel sums and the [B,C] scales), 16-byte aligned. */
size_t dn_se_workspace_bytes(int B, int HW, int C);
int dn_se_inplace(void* x, const float* w1, const float* b1, const float* w2t, const float* b2, int B, int HW,
                  int C, int Cs, void* workspace, size_t workspace_bytes, void* stream);

/* Squeeze-excitation folded into the project GEMM behind it (InvertedResidual.block[-2:], mobilenetv3.py:85-89): the
 * scales hardsigmoid(fc2(relu(fc1(avgpool(x))))) are computed as in dn_se_inplace, but instead of rewriting x the 1x1
 * convolution multiplies its A operand by them in shared memory, rounding to the storage type exactly like the in-place
 * pass -- the result equals dn_se_inplace followed by dn_pwconv(act = none) bit for bit, with one read and one write
 * of x less.  x: [B,HW,C] (NOT modified); w_pw: [N,C]; b_pw: fp32 [N]; residual: [B*HW,N] or NULL; y: [B*HW,N];
 * workspace as for dn_se_inplace. */
int dn_se_project(const void* x, const float* se_w1, const float* se_b1, const float* se_w2t, const float* se_b2,
                  const void* w_pw, const float* b_pw, const void* residual, void* y, int B, int HW, int C, int Cs, int N,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Depthwise convolution (dn_dwconv) followed by the squeeze-excitation of its output (dn_se_inplace), the pair an
 * InvertedResidual with use_se runs between its expand and project convolutions (mobilenetv3.py:43-96).  When the
 * layer runs on the stride-1 row-stream kernel and the batch is large enough for every image to be covered by at most
 * 16 CTA shares, the depthwise launch leaves the SE channel sums in the workspace and the SE pooling pass over y is
 * skipped (*pooled_out = 1, may be NULL); otherwise the two calls above run back to back.  Arguments as for
 * dn_dwconv and dn_se_inplace; workspace of dn_se_workspace_bytes(B, Ho * Wo, C) bytes. */
int dn_dwconv_se(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int k,
                 int stride, int act, const float* se_w1, const float* se_b1, const float* se_w2t, const float* se_b2,
                 int Cs, void* workspace, size_t workspace_bytes, int* pooled_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * the operators ssd300_vgg16 (demonet/models/ssd_vgg16.py:139-213) needs beyond the SSDLite path
 * ---------------------------------------------------------------------------------------- */

/* Dense 3x3 convolution + bias + activation as an implicit GEMM on the tcgen05 tensor cores (no im2col buffer: one 4-D TMA
 * box per tap, out-of-bounds zero fill = the convolution's padding).  Replaces nn.Conv2d(C, N, 3, stride, padding,
 * dilation) + ReLU of the VGG16 features and the SSD extra blocks (ssd_vgg16.py:30-109, conv6: dilation 6) and the dense
 * SSD heads (SSDScoringHead / SSDClassificationHead / SSDRegressionHead, generalized_ssd.py:38-92).
 * x: [B,H,W,C] 16-bit NHWC, C % 64 == 0; w: 16-bit [9][N][C] (tap = kh * 3 + kw major); bias fp32 [N];
 * y: [B,Ho,Wo,N] 16-bit, or fp32 with the head addressing of dn_pwconv when out_fp32 != 0 (out_*_stride = 0: dense).
 * stride 1 or 2; pad 0 or = dilation; act: none / relu / relu6. */
int dn_conv3x3(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C, int N, int stride, int pad,
               int dilation, int act, int out_fp32, int64_t out_batch_stride, int64_t out_row_stride, void* stream);

/* Input normalisation (transform.py:129-138) + the first VGG convolution (3 -> 64, 3x3, stride 1, padding 1) + ReLU.
 * images: fp32 NCHW [B,3,H,W] in [0,1]; w: fp32 [27][64] ((ci*3+kh)*3+kw major); y: 16-bit NHWC [B,H,W,64]. */
int dn_conv3x3_first(const float* images, const float* w, const float* bias, const float* mean3_host, const float* std3_host,
                     void* y, int B, int H, int W, int Cout, void* stream);

/* The tensor-core form of the same layer: writes the normalised 3x3 neighbourhood of every pixel as one row of 32 16-bit
 * values (27 taps, (ci*3+kh)*3+kw major, zero padding at the border, 5 trailing zeros) -- cols: [B*H*W, 32] -- so that
 * dn_pwconv(cols, w16 [64][32], bias, relu) is the convolution (K = 32, one tcgen05 k-block).  The image values are rounded
 * to the 16-bit activation type once, like the input of every later layer. */
int dn_im2col3x3_first(const float* images, const float* mean3_host, const float* std3_host, void* cols, int B, int H, int W,
                       void* stream);

/* nn.MaxPool2d(k, stride, pad, ceil_mode) on 16-bit NHWC activations (vgg features; ceil_mode patched in at
 * ssd_vgg16.py:36-37; the 3x3 stride-1 "pool5" of ssd_vgg16.py:84). */
int dn_maxpool2d(const void* x, void* y, int B, int H, int W, int C, int k, int stride, int pad, int ceil_mode, void* stream);

/* scale_weight * F.normalize(x) over the channels of every pixel (ssd_vgg16.py:98-100, eps 1e-12).  x, y: [npix, C]. */
int dn_l2norm_scale(const void* x, const float* scale, void* y, int64_t npix, int C, void* stream);

/* ------------------------------------------------------------------------------------------
 * Transforms either side of the path (GeneralizedRCNNTransform, demonet/models/transform.py)
 * ---------------------------------------------------------------------------------------- */

/* Fixed-size bilinear resize of ONE image, = torch.nn.functional.interpolate(image[None], size=(Ho,Wo),
 * mode='bilinear', align_corners=False) as called by _resize_image_and_masks (transform.py:27-53).
 * src: CHW, fp32 in [0,1] (src_is_u8 = 0) or uint8 (src_is_u8 = 1: converted like ToTensor, x / 255, first);
 * dst: fp32 [C,Ho,Wo] (e.g. row b of the engine's [B,3,S,S] input batch). */
int dn_resize_bilinear(const void* src, int src_is_u8, int C, int H, int W, float* dst, int Ho, int Wo, void* stream);

/* dst[i] = src[i] / 255 in fp32 (ToTensor), n elements. */
int dn_u8_to_f32(const uint8_t* src, float* dst, size_t n, void* stream);

/* boxes[b,d,:] *= (rw, rh, rw, rh) with ratio_hw[b] = (rh, rw) = original size / network size, computed by the
 * caller in fp32 exactly as resize_boxes does (transform.py:278-292; applied by postprocess, transform.py:228-247).
 * boxes: fp32 [B,D,4] in place; ratio_hw: fp32 [B,2] device. */
int dn_rescale_boxes(float* boxes, const float* ratio_hw, int B, int D, void* stream);

/* Detection sink: padded per-image detections (the outputs of dn_postprocess / dn_engine_forward) -> compact COCO
 * result rows in image order, i.e. CocoEvaluator.prepare_for_coco_detection (demonet/data/coco_eval.py:76-98) with
 * convert_to_xywh (coco_eval.py:162-164: w = xmax - xmin, h = ymax - ymin in fp32) on the device.
 * boxes fp32 [B,D,4] xyxy, scores fp32 [B,D], labels int64 [B,D], counts int32 [B], image_ids int64 [B].
 * Outputs (capacity B*D rows): out_image_id int64, out_category_id int64, out_bbox_xywh fp32 [.,4], out_score fp32,
 * out_total int64 [1] = number of rows written (sum of counts). */
int dn_detections_to_coco(const float* boxes, const float* scores, const int64_t* labels, const int32_t* counts,
                          const int64_t* image_ids, int B, int D, int64_t* out_image_id, int64_t* out_category_id,
                          float* out_bbox_xywh, float* out_score, int64_t* out_total, void* stream);

typedef struct {
    int32_t num_priors;          /* P                                                          */
    int32_t num_classes;         /* K, including background column 0                           */
    int32_t image_h, image_w;    /* clip window (clip_boxes_to_image)                          */
    float score_thresh;          /* compared in fp32: score > thresh                           */
    double nms_thresh;           /* compared in double: (double)iou > thresh                   */
    int32_t topk_candidates;     /* per-class top-k before NMS; <= 0 means "no top-k" (legacy) */
    int32_t detections_per_img;  /* D                                                          */
    float min_box_size;          /* legacy remove_small_boxes(min_size); < 0 disables          */
    float box_weights[4];        /* BoxCoder weights (10,10,5,5)                               */
    float bbox_xform_clip;       /* log(1000/16)                                               */
} dn_postprocess_params;

/* Fused softmax + box decode + clip + score threshold + per-class top-k + per-class NMS + top-D.
 * Replaces SSD.postprocess_detections (generalized_ssd.py:351-397) and, with topk_candidates <= 0
 * and min_box_size = 1e-2, the legacy PostProcess.forward (box_head.py:323-381).
 * cls_logits fp32 [B,P,K]; bbox_regression fp32 [B,P,4]; anchors fp32 [P,4] xyxy pixels.
 * Outputs (fixed shape, padded): boxes fp32 [B,D,4], scores fp32 [B,D], labels int64 [B,D],
 * counts int32 [B]; rows >= counts[b] are zero.  Detections are in descending score order, ties
 * broken by (class, rank within class). */
size_t dn_postprocess_workspace_bytes(int B, const dn_postprocess_params* p);
int dn_postprocess(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                   const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                   float* out_scores, int64_t* out_labels, int32_t* out_counts, void* stream);

/* The same post-processing entered where the reference already holds softmax scores and decoded, clipped boxes
 * (generalized_ssd.py:361-363): per-class threshold / top-k (:368-382), batched NMS (:389) and keep[:D] (:390-396) run
 * through exactly the kernels dn_postprocess and the engine use (class sort, warp / CTA NMS with lazy rounds, top-D
 * merge); only the softmax + decode arithmetic is skipped.  This is the parity entry of SURVEY.md 8(a) N1: fed the
 * reference's own scores and boxes, (out_priors, out_labels) must equal the (anchor, class) identity of
 * _batched_nms_vanilla(...)[:detections_per_img] bit for bit, and out_scores / out_boxes are copies of the inputs.
 * scores fp32 [B,P,K] (column 0 = background, ignored); boxes fp32 [B,P,4] xyxy.
 * out_priors (may be NULL): int32 [B,D], anchor index of every detection, -1 in the padding.
 * out_rounds (may be NULL): int32 [B], the lazy-NMS round (0, 1, 2) that made the image final -- lets a test prove
 * that the deeper rounds were exercised. */
int dn_postprocess_scored(const float* scores, const float* boxes, int B, const dn_postprocess_params* p, void* workspace,
                          size_t workspace_bytes, float* out_boxes, float* out_scores, int64_t* out_labels,
                          int32_t* out_counts, int32_t* out_priors, int32_t* out_rounds, void* stream);

/* measurement aid: same as dn_postprocess, each of its 3 kernels launched `iters` times between CUDA
 * events; ms3_host[0..3) = mean ms of softmax+decode, per-class top-k+NMS, top-D merge.  Synchronises. */
int dn_postprocess_profile(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                           const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                           float* out_scores, int64_t* out_labels, int32_t* out_counts, int iters, float* ms3_host,
                           void* stream);

/* torchvision.ops.batched_nms with per-class ("vanilla") semantics, bit-exact against the CPU
 * kernel (the call at generalized_ssd.py:389 / box_head.py:374): boxes fp32 [n,4], scores fp32 [n],
 * idxs int64 [n] with 0 <= idx < 4096 and fewer than 4096 KEPT boxes per class (nkeep_out = -1 / -2
 * otherwise).  keep_out int64 [n] receives
 * the kept indices in descending score order (ties: lower index first), nkeep_out int64 [1]. */
size_t dn_batched_nms_workspace_bytes(int64_t n);
int dn_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n, double iou_threshold,
                   void* workspace, size_t workspace_bytes, int64_t* keep_out, int64_t* nkeep_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-side operators: default-box matching and the multibox loss
 * ---------------------------------------------------------------------------------------- */

/* Workspace of the three entry points below for B images, P default boxes per image and G ground-truth boxes in total. */
size_t dn_ssd_loss_workspace_bytes(int B, int P, int G);

/* box_ops.box_iou(targets['boxes'], anchors) followed by SSDMatcher(iou_thresh) for every image of a batch
 * (generalized_ssd.py:326-335; _utils.py:283-323 Matcher.__call__ with low == high threshold, :350-362 the forced match
 * of each ground-truth box to its best default box).  gt_boxes: fp32 [G,4] xyxy, the images' boxes back to back;
 * gt_offsets: int32 [B+1] (device), image b owns rows gt_offsets[b] .. gt_offsets[b+1]; anchors: fp32 [P,4] (shared by the
 * images, anchor_utils.py:110-126); matched_idxs: int64 [B,P] = index of the matched ground-truth box within its image,
 * or -1.  An image without boxes gets -1 everywhere (generalized_ssd.py:329-332).  Bit-exact against the reference's CPU
 * result: first maximum on arg-max ties, the larger ground-truth index when two boxes claim one default box. */
int dn_ssd_match(const float* gt_boxes, const int32_t* gt_offsets, const float* anchors, int B, int P, int G,
                 float iou_thresh, int64_t* matched_idxs, void* workspace, size_t workspace_bytes, void* stream);

/* SSDMatcher.__call__(match_quality_matrix) (_utils.py:350-362) on a caller-supplied fp32 [M,P] matrix (one image);
 * matches: int64 [P].  DN_ERR_INVALID with the reference's message for an empty matrix (_utils.py:298-307). */
int dn_match_quality(const float* quality, int M, int P, float thresh, int64_t* matches, void* workspace,
                     size_t workspace_bytes, void* stream);

/* SSD.compute_loss (generalized_ssd.py:210-269): smooth-L1 on BoxCoder.encode_single targets (_utils.py:83-127) of the
 * matched boxes, softmax cross entropy against the matched labels (0 = background), hard-negative mining (the
 * ceil(neg_to_pos_ratio * #foreground) largest background losses per image), both sums divided by N = max(1, #matched).
 * cls_logits: fp32 [B,P,K]; bbox_regression: fp32 [B,P,4]; gt_labels: int64 [G]; matched_idxs: int64 [B,P] (dn_ssd_match).
 * losses (device fp32 [3]) receives {bbox_regression, classification, N}.  grad_cls (fp32 [B,P,K]) / grad_reg (fp32 [B,P,4]),
 * when not NULL, receive d(classification)/d(cls_logits) and d(bbox_regression loss)/d(bbox_regression).  Sums are
 * accumulated in double in a fixed order (deterministic).  Equal losses at the mining cut go to the lower index. */
int dn_ssd_loss(const float* cls_logits, const float* bbox_regression, const float* anchors, const float* gt_boxes,
                const int64_t* gt_labels, const int32_t* gt_offsets, const int64_t* matched_idxs, int B, int P, int K,
                int G, float neg_to_pos_ratio, const float* box_weights4_host, float* losses, float* grad_cls,
                float* grad_reg, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Engine: the whole forward (SSD.forward eval branch, generalized_ssd.py:271-349) as one
 * pre-planned launch sequence replayed from a CUDA graph.
 * ---------------------------------------------------------------------------------------- */
typedef enum {
    DN_OP_STEM = 0,      /* normalise + dense 3x3 s2 conv                      */
    DN_OP_DW = 1,        /* depthwise conv                                      */
    DN_OP_PW = 2,        /* pointwise GEMM                                      */
    DN_OP_SE = 3,        /* squeeze-excitation, in place on in_buf              */
    DN_OP_PWDW = 4,      /* fused pointwise expand + depthwise (dn_pwdw_fused): in_buf [H,W,c_in] -> out_buf
                            [h_out,w_out,c_out]; w_off/b_off = expand weights / bias, w2_off/b2_off = depthwise
                            weights / bias, act = expand activation, act2 = depthwise activation             */
    DN_OP_NOP = 5,       /* placeholder of a layer that was fused into its predecessor                       */
    DN_OP_DWPW = 6       /* fused depthwise + pointwise project (dn_dwpw_fused): w_off/b_off = depthwise weights /
                            bias, w2_off/b2_off = project weights / bias, act = depthwise activation, res_buf = in_buf
                            for `result += input` or DN_BUF_NONE                                              */
} dn_op_kind;

#define DN_BUF_NONE (-1)
#define DN_BUF_IMAGES (-2)      /* the caller's fp32 NCHW images          */

typedef struct {
    int32_t kind, act;
    int32_t in_buf, out_buf, res_buf;     /* arena buffer ids                              */
    int32_t h_in, w_in, c_in;
    int32_t h_out, w_out, c_out;
    int32_t ksize, stride;
    int32_t c_mid;                         /* SE squeeze width                              */
    int32_t out_fp32;
    int32_t lane;                          /* launch lane: 0 = main chain, > 0 = a side branch (head) that may run
                                              concurrently; its tensors never share arena buffers with other lanes */
    int32_t act2;                          /* second activation of a fused op               */
    int32_t se_fold;                       /* squeeze-excitation folded into its project GEMM: set on the DN_OP_SE (it then
                                              only produces the [B][C] scales) AND on the DN_OP_PW directly behind it (which
                                              applies them to its A operand in shared memory); x itself is never rescaled */
    int64_t w_off, b_off, w2_off, b2_off;  /* byte offsets into the weight blob             */
    int64_t out_batch_stride, out_row_stride, out_offset;   /* PW output addressing (elements) */
} dn_op;

typedef struct {
    int64_t elems_per_image;
    int32_t elem_bytes;                    /* 2 = bf16, 4 = fp32 */
    int32_t reserved;
} dn_buf;

typedef struct {
    int32_t image_h, image_w;
    float image_mean[3], image_std[3];
    int32_t n_ops, n_bufs;
    const dn_op* ops_host;
    const dn_buf* bufs_host;
    int32_t logits_buf, bbox_buf;          /* arena ids of fp32 [P,K] and [P,4] per image    */
    const float* anchors_host;             /* fp32 [P,4]                                      */
    dn_postprocess_params post;
    int32_t gemm_impl;                     /* 0 = tcgen05 (product), 1 = SIMT self-check      */
    int32_t use_cuda_graph;
    int32_t pipeline_slots;                /* 0 / 1: one forward at a time (strict stream semantics).  n = 2..4: consecutive
                                              dn_engine_forward calls go round n complete engine instances on
                                              engine-owned streams and overlap; see dn_engine_join                     */
    int32_t reserved;
} dn_model_desc;

typedef struct dn_engine dn_engine;

int dn_engine_create(dn_engine** out, const dn_model_desc* desc, int max_batch);
int dn_engine_destroy(dn_engine* e);
/* (re)upload the packed weight blob (BN already folded, layouts as documented in DESIGN.md) */
int dn_engine_load_weights(dn_engine* e, const void* blob_host, size_t bytes);
/* images_dev: fp32 [B,3,H,W]; outputs as in dn_postprocess */
int dn_engine_forward(dn_engine* e, const float* images_dev, int B, float* out_boxes, float* out_scores,
                      int64_t* out_labels, int32_t* out_counts, void* stream);
/* Pipeline mode only (pipeline_slots = n >= 2): dn_engine_forward returns with the forward enqueued on an engine-owned
 * stream (ordered behind the work already on `stream`), NOT yet ordered before later work on `stream`.
 * dn_engine_join makes `stream` wait for every forward issued so far; dn_engine_join_previous only for the OLDEST
 * forward in flight -- the one issued n - 1 calls before the most recent one, whose slot the next call reuses (n = 2: the
 * forward before the most recent one) -- so that a consumer of batch i-n+1 runs while batches i-n+2 .. i compute.
 * Both are no-ops without pipeline mode.  The host entry points keep their plain contract in either mode. */
int dn_engine_join(dn_engine* e, void* stream);
int dn_engine_join_previous(dn_engine* e, void* stream);
/* same, taking PINNED HOST buffers: H2D copy, forward, D2H copy.  The forward and the D2H copies are
 * enqueued on `stream`; the H2D copy runs on an engine-owned stream that `stream` waits for, into one of two
 * staging buffers, so back-to-back calls overlap the upload of call i+1 with the forward of call i.
 * The results (and images_host) are safe to touch once `stream` has been synchronised. */
int dn_engine_forward_host(dn_engine* e, const float* images_host, int B, float* out_boxes_host,
                           float* out_scores_host, int64_t* out_labels_host, int32_t* out_counts_host,
                           void* stream);
/* same with uint8 [B,3,H,W] pixels (0..255) in pinned host memory: a quarter of the PCIe bytes; the ToTensor
 * conversion x / 255 (exact in fp32) runs on the device in front of the stem (dn_u8_to_f32). */
int dn_engine_forward_host_u8(dn_engine* e, const uint8_t* images_host, int B, float* out_boxes_host,
                              float* out_scores_host, int64_t* out_labels_host, int32_t* out_counts_host,
                              void* stream);
/* device pointers of intermediate arena buffers (valid after a forward), for stage-by-stage parity; in pipeline mode
 * dn_engine_copy_buffer reads the arena of the slot that ran the forward issued last */
int dn_engine_buffer(dn_engine* e, int buf_id, void** ptr_out, int64_t* elems_per_image_out);
/* enqueue a device-to-device copy of the first `bytes` bytes of an arena buffer into dst_dev */
int dn_engine_copy_buffer(dn_engine* e, int buf_id, void* dst_dev, size_t bytes, void* stream);
/* Per-launch device time measured inside a CUDA graph: the plan is captured on one stream with an event-record node
 * between consecutive ops, replayed `iters` times, and the mean milliseconds between consecutive events are written to
 * ms_out_host[0 .. n_ops + 3): the n_ops layers in plan order (0 for layers fused into their predecessor), then
 * softmax + decode, then class sort + NMS + top-D merge (all rounds), then the cost of an EMPTY bracket (two event nodes
 * with nothing in between, about 2.6 us on B200), which has already been subtracted from every entry.  Every op runs
 * once per replay, in plan order, on the cache state the real step leaves it.  Synchronises; a measurement aid. */
int dn_engine_profile(dn_engine* e, const float* images_dev, int B, int iters, float* ms_out_host, void* stream);
/* What the forward issued last actually ran (parity tests assert that the benchmarked configuration -- fused blocks,
 * pooled squeeze-excitation, graph replay, two slots -- is the one being compared with the oracle). */
typedef struct {
    int32_t launches_per_forward;
    int32_t fused_pwdw, fused_dwpw;      /* fused expand+depthwise / depthwise+project launches in the plan          */
    int32_t se_layers, se_pooled;        /* squeeze-excitation layers / those whose pooling the depthwise launch did */
    int32_t se_folded, reserved;         /* ... / those whose scaling pass is folded into the project GEMM           */
    int32_t pipeline_slots, last_slot;   /* 1..4 engine instances; the slot of the forward issued last              */
    int32_t act_dtype;                   /* storage type of the activations: 0 = bf16, 1 = fp16                      */
    int64_t forwards, graph_replays;     /* forwards enqueued so far / those that were one cudaGraphLaunch           */
} dn_engine_stats;
int dn_engine_get_stats(dn_engine* e, dn_engine_stats* out);
/* number of kernel launches one forward enqueues (for bench.py's gpu_launches) */
int dn_engine_launches_per_forward(dn_engine* e);
size_t dn_engine_device_bytes(dn_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* DEMONET_B200_H_ */
