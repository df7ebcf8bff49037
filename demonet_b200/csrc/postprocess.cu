// Fused SSD post-processing for sm_100a: softmax + box decode + clip, per-(image,class) threshold /
// top-k / greedy NMS, and the per-image top-D merge.  Replaces SSD.postprocess_detections
// (demonet/models/generalized_ssd.py:351-397), the legacy PostProcess.forward
// (demonet/models/box_head.py:323-381) and torchvision.ops.batched_nms as called from both.
//
// Bit-exactness contract, all in fp32 WITHOUT fused multiply-add:
//   area = (x2-x1)*(y2-y1);  inter = max(0,xx2-xx1)*max(0,yy2-yy1);
//   iou  = inter / ((area_i + area_j) - inter);   suppress iff (double)iou > thr   (strict)
//   candidates visited in stable descending-score order; NaN IoU never suppresses.
// `(double)iou > thr` is evaluated as `iou >= thr_up`, thr_up = the smallest float whose double
// value exceeds thr (computed on the host) -- identical for every float iou, NaN included.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace dn {

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t orderable(float f) {      // monotone float -> uint map
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}
// ascending sort of this key == (score descending, index ascending)
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t idx) {
    return ((unsigned long long)(~orderable(score)) << 32) | idx;
}
__device__ __forceinline__ float key_score(unsigned long long k) { return from_orderable(~(uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_index(unsigned long long k) { return (uint32_t)k; }

__device__ __forceinline__ float box_area(const float4& b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
// true iff box j is suppressed by kept box i
__device__ __forceinline__ bool iou_exceeds(const float4& a, float aarea, const float4& b, float barea,
                                            float thr_up) {
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
    float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    float ovr = __fdiv_rn(inter, uni);
    return ovr >= thr_up;
}

// Same predicate as iou_exceeds, but most pairs are decided by an approximate quotient: a pair is
// handed to the exact IEEE division only when the approximation lies within 4e-6 (relative) of the
// threshold (rcp + one multiply are accurate to ~2e-7, the exact quotient to 6e-8), so the result is
// bit-identical to the division for every input (NaN/inf fall through to the exact path).
__device__ __forceinline__ bool iou_exceeds_fast(const float4& a, float aarea, const float4& b, float barea,
                                                 float thr_up, float thr_lo, float thr_hi) {
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
    float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(uni));
    float q = __fmul_rn(inter, r);
    if (q < thr_lo) return false;
    if (q > thr_hi && q < 3.0e38f) return true;
    return __fdiv_rn(inter, uni) >= thr_up;
}

// in-smem bitonic sort (ascending) of n_pad (power of two) 64-bit keys by the whole CTA
__device__ void bitonic_sort_u64(unsigned long long* keys, int n_pad) {
    if (n_pad <= 64) {                 // one warp is enough: no block-wide barriers
        if (threadIdx.x < 32) {
            for (int k = 2; k <= n_pad; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = threadIdx.x; t < (n_pad >> 1); t += 32) {
                        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        int hi = lo | j;
                        unsigned long long a = keys[lo], b = keys[hi];
                        bool up = ((lo & k) == 0);
                        if ((a > b) == up) {
                            keys[lo] = b;
                            keys[hi] = a;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        return;
    }
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int hi = lo | j;
                unsigned long long a = keys[lo], b = keys[hi];
                bool up = ((lo & k) == 0);
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Greedy-NMS consumer.  Candidates arrive in visiting order in chunks of <= 32 (one per lane of
// warp 0); the kept list (boxes + areas) lives in shared memory.  Every warp of the CTA helps with
// the candidate-vs-kept test; warp 0 resolves the 32x32 dependencies inside the chunk with ballots.
// A candidate is kept iff no EARLIER KEPT box overlaps it by more than the threshold -- exactly the
// greedy loop of the CPU kernel.
// ---------------------------------------------------------------------------------------------
struct NmsState {
    float4* kept_box;        // smem [max_keep]
    float* kept_area;        // smem [max_keep]
    float4* chunk_box;       // smem [32]
    float* chunk_area;       // smem [32]
    unsigned int* dead;      // smem word
    int* n_kept;             // smem word
};

// cand_cnt <= 32 candidates already staged in st.chunk_box/area by the caller (and synced).
// Returns (to every thread) a 32-bit mask of the candidates that were kept AND fit under max_keep;
// appends them to the kept list.  Must be called by all threads of the CTA.
__device__ unsigned int nms_consume_chunk(NmsState& st, int cand_cnt, float thr_up, int max_keep) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nk = *st.n_kept;
    // phase A: candidate (lane) vs kept boxes (strided over warps)
    bool dead = false;
    if (lane < cand_cnt) {
        const float4 cb = st.chunk_box[lane];
        const float ca = st.chunk_area[lane];
        for (int k = warp; k < nk; k += nwarps) {
            if (iou_exceeds(st.kept_box[k], st.kept_area[k], cb, ca, thr_up)) {
                dead = true;
                break;
            }
        }
    }
    unsigned int dead_bits = __ballot_sync(0xffffffffu, dead);
    if (lane == 0 && dead_bits) atomicOr(st.dead, dead_bits);
    __syncthreads();
    // phase B: warp 0 resolves the chunk
    __shared__ unsigned int s_result;
    if (warp == 0) {
        unsigned int valid = (cand_cnt >= 32) ? 0xffffffffu : ((1u << cand_cnt) - 1u);
        unsigned int alive = valid & ~(*st.dead);
        float4 mb = make_float4(0.f, 0.f, 0.f, 0.f);
        float ma = 0.f;
        if (lane < cand_cnt) {
            mb = st.chunk_box[lane];
            ma = st.chunk_area[lane];
        }
        for (int i = 0; i < cand_cnt; ++i) {
            if (!((alive >> i) & 1u)) continue;          // warp-uniform
            float4 ib;
            ib.x = __shfl_sync(0xffffffffu, mb.x, i);
            ib.y = __shfl_sync(0xffffffffu, mb.y, i);
            ib.z = __shfl_sync(0xffffffffu, mb.z, i);
            ib.w = __shfl_sync(0xffffffffu, mb.w, i);
            float ia = __shfl_sync(0xffffffffu, ma, i);
            bool sup = (lane > i) && ((alive >> lane) & 1u) && iou_exceeds(ib, ia, mb, ma, thr_up);
            alive &= ~__ballot_sync(0xffffffffu, sup);
        }
        int room = max_keep - nk;
        int rank = __popc(alive & ((1u << lane) - 1u));
        bool mine = ((alive >> lane) & 1u) && rank < room;
        if (mine) {
            st.kept_box[nk + rank] = mb;
            st.kept_area[nk + rank] = ma;
        }
        unsigned int taken = __ballot_sync(0xffffffffu, mine);
        __syncwarp();                                    // every lane has read *st.dead before lane 0 clears it
        if (lane == 0) {
            *st.n_kept = nk + __popc(taken);
            *st.dead = 0u;
            s_result = taken;
        }
    }
    __syncthreads();
    return s_result;
}

// ---------------------------------------------------------------------------------------------
// P1: softmax over classes + box decode + clip.  One CTA = 32 consecutive priors of one image.
//   scores_t[b][k-1][p] = softmax(logits[b][p][:])[k]   (class-major so that P2 reads contiguously)
//   boxes[b][p]         = clip(decode(bbox[b][p], anchors[p]))
// ---------------------------------------------------------------------------------------------
constexpr int P1_ROWS = 32;
constexpr int P1_THREADS = 256;
constexpr int P1_MAX_PER_LANE = 3;     // register path of the softmax covers K <= 96 (91 COCO / 21 VOC classes)
// Per-image histogram of the foreground scores over a monotone key (float bits >> 19: 16 bins per
// binade, from 2^-24 up to 1.0).  It only steers how deep the lazy NMS rounds go -- any threshold is
// exact -- so the resolution (4.4 % in score) is irrelevant for correctness.  16 bins per binade is the measured
// optimum (r01): finer bins cost more global atomics in the per-CTA flush, coarser ones serialise the
// shared-memory atomics of a warp on the same bin.
constexpr int HIST_SHIFT = 19;
constexpr int HIST_BASE = 0x33800000 >> HIST_SHIFT;
constexpr int HIST_BINS = (0x3F800000 >> HIST_SHIFT) - HIST_BASE + 2;
__device__ __forceinline__ int hist_bin(float s) {
    int b = (int)(__float_as_uint(s) >> HIST_SHIFT) - HIST_BASE;
    return b < 0 ? 0 : (b >= HIST_BINS ? HIST_BINS - 1 : b);
}
__device__ __forceinline__ float hist_bin_lower_edge(int b) {
    return b <= 0 ? 0.f : __uint_as_float((uint32_t)(b + HIST_BASE) << HIST_SHIFT);
}

// SCORED: `logits` already holds softmax scores and `bbox` decoded, clipped boxes (dn_postprocess_scored -- the
// reference's own tensors at generalized_ssd.py:361-363); the kernel then only re-lays them out (class-major scores,
// float4 boxes) and builds the histogram, so that everything downstream is the very code the engine runs.
template <bool SCORED>
__global__ void __launch_bounds__(P1_THREADS)
softmax_decode_kernel(const float* __restrict__ logits, const float* __restrict__ bbox,
                      const float* __restrict__ anchors, float* __restrict__ scores_t,
                      float4* __restrict__ boxes, int* __restrict__ hist, int P, int K, dn_postprocess_params prm) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_tile[];                    // [P1_ROWS][ld], ld odd -> conflict-free column reads
    __shared__ int s_hist[HIST_BINS];
    for (int i = threadIdx.x; i < HIST_BINS; i += P1_THREADS) s_hist[i] = 0;
    const int ld = K | 1;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * P1_ROWS;
    const int rows = min(P1_ROWS, P - p0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // decode + clip (BoxCoder.decode_single, _utils.py:187-224; clip_boxes_to_image)
    if (SCORED) {
        if (threadIdx.x < rows) {
            const int p = p0 + threadIdx.x;
            boxes[(size_t)b * P + p] = reinterpret_cast<const float4*>(bbox)[(size_t)b * P + p];
        }
    } else if (threadIdx.x < rows) {
        const int p = p0 + threadIdx.x;
        const float4 r = reinterpret_cast<const float4*>(bbox)[(size_t)b * P + p];
        const float4 a = reinterpret_cast<const float4*>(anchors)[p];
        const float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);
        const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
        const float dx = __fdiv_rn(r.x, prm.box_weights[0]), dy = __fdiv_rn(r.y, prm.box_weights[1]);
        const float dw = fminf(__fdiv_rn(r.z, prm.box_weights[2]), prm.bbox_xform_clip);
        const float dh = fminf(__fdiv_rn(r.w, prm.box_weights[3]), prm.bbox_xform_clip);
        const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
        const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
        const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
        float4 o;
        o.x = fminf(fmaxf(__fsub_rn(pcx, hw), 0.f), (float)prm.image_w);
        o.y = fminf(fmaxf(__fsub_rn(pcy, hh), 0.f), (float)prm.image_h);
        o.z = fminf(fmaxf(__fadd_rn(pcx, hw), 0.f), (float)prm.image_w);
        o.w = fminf(fmaxf(__fadd_rn(pcy, hh), 0.f), (float)prm.image_h);
        boxes[(size_t)b * P + p] = o;
    }

    // softmax, one warp per row; a lane keeps its (up to P1_MAX_PER_LANE) elements in registers and writes
    // the finished probabilities to shared memory once
    for (int r = warp; r < rows; r += P1_THREADS / 32) {
        const float* src = logits + ((size_t)b * P + p0 + r) * K;
        float* dst = s_tile + r * ld;
        if (SCORED) {
            for (int k = lane; k < K; k += 32) dst[k] = src[k];
        } else if (K <= 32 * P1_MAX_PER_LANE) {
            float v[P1_MAX_PER_LANE];
            float m = -FLT_MAX;
#pragma unroll
            for (int i = 0; i < P1_MAX_PER_LANE; ++i) {
                const int k = lane + 32 * i;
                v[i] = (k < K) ? src[k] : -FLT_MAX;
                m = fmaxf(m, v[i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < P1_MAX_PER_LANE; ++i) {
                // ex2.approx: |rel err| <= 2^-21 * |x - m|, i.e. <= 1e-7 absolute on a score
                v[i] = (lane + 32 * i < K) ? __expf(v[i] - m) : 0.f;
                s += v[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float inv = __fdiv_rn(1.f, s);
#pragma unroll
            for (int i = 0; i < P1_MAX_PER_LANE; ++i)
                if (lane + 32 * i < K) dst[lane + 32 * i] = v[i] * inv;
        } else {
            float m = -FLT_MAX;
            for (int k = lane; k < K; k += 32) {
                float v = src[k];
                dst[k] = v;
                m = fmaxf(m, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s = 0.f;
            for (int k = lane; k < K; k += 32) {
                float e = __expf(dst[k] - m);
                dst[k] = e;
                s += e;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float inv = __fdiv_rn(1.f, s);
            for (int k = lane; k < K; k += 32) dst[k] *= inv;
        }
    }
    __syncthreads();
    // transposed, coalesced store of classes 1..K-1 (+ histogram of the scores above the threshold)
    for (int k = 1 + warp; k < K; k += P1_THREADS / 32) {
        if (lane < rows) {
            const float sv = s_tile[lane * ld + k];
            scores_t[((size_t)b * (K - 1) + (k - 1)) * P + p0 + lane] = sv;
            if (sv > prm.score_thresh) atomicAdd(&s_hist[hist_bin(sv)], 1);
        }
    }
    __syncthreads();
    int* gh = hist + (size_t)b * HIST_BINS;
    for (int i = threadIdx.x; i < HIST_BINS; i += P1_THREADS) {
        const int c = s_hist[i];
        if (c) atomicAdd(gh + i, c);
    }
}

// One warp per image: score thresholds of the lazy-NMS rounds.  Round r processes every candidate whose
// score is >= thr[b][r], chosen so that about targets[r] candidates qualify (0 = everything).
constexpr int NMS_ROUNDS = 3;
struct RoundTargets {
    int t[NMS_ROUNDS];
};
__global__ void __launch_bounds__(32)
pick_thresholds_kernel(const int* __restrict__ hist, float* __restrict__ thr, RoundTargets targets) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x, lane = threadIdx.x;
    const int* gh = hist + (size_t)b * HIST_BINS;
    constexpr int PER = (HIST_BINS + 31) / 32;
    // lane l owns bins [l*PER, (l+1)*PER); suffix = candidates in all higher lanes
    int mine = 0;
    for (int i = 0; i < PER; ++i) {
        const int bin = lane * PER + i;
        if (bin < HIST_BINS) mine += gh[bin];
    }
    int above = 0;
    for (int l = 31; l >= 0; --l) {
        const int v = __shfl_sync(0xffffffffu, mine, l);
        if (l > lane) above += v;
    }
    for (int r = 0; r < NMS_ROUNDS; ++r) {
        const int target = targets.t[r];
        float t = 0.f;                                   // everything
        const bool here = target > 0 && above < target && above + mine >= target;
        if (here) {
            int acc = above;
            for (int i = PER - 1; i >= 0; --i) {
                const int bin = lane * PER + i;
                if (bin >= HIST_BINS) continue;
                acc += gh[bin];
                if (acc >= target) {
                    t = hist_bin_lower_edge(bin);
                    break;
                }
            }
        }
        // at most one lane found a crossing; otherwise fewer than `target` candidates exist -> 0
        for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (lane == 0) thr[b * NMS_ROUNDS + r] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// P2a: one CTA per (class, image): score threshold, stable descending sort, top-k cut.  Writes the
// class's candidate list (keys = score desc, prior asc) to global memory.
// ---------------------------------------------------------------------------------------------
constexpr int P2_THREADS = 256;

struct Entry {                        // one kept detection of a class list
    float score;
    int prior;
};

__global__ void __launch_bounds__(P2_THREADS)
class_sort_kernel(const float* __restrict__ scores_t, const float4* __restrict__ boxes,
                  unsigned long long* __restrict__ cand_keys, int* __restrict__ cand_counts, int P, int K, int cap,
                  float score_thresh, int topk, float min_box_size, const float* __restrict__ thr,
                  const int* __restrict__ done, int round, int B) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(s_raw);              // [next_pow2(P)]
    __shared__ int s_n;
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31;
    // round 0: gridDim.y == B, one image per CTA.  Later rounds run on a few image slots per class and skip the
    // images that are done (usually all of them), instead of launching B x (K-1) CTAs that exit at once.
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
    if (round > 0 && done[b]) continue;
    const float* sc = scores_t + ((size_t)b * (K - 1) + c) * P;
    const float4* bx = boxes + (size_t)b * P;
    const float round_thr = thr[b * NMS_ROUNDS + round];      // lazy NMS: only the top of the image this round
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // threshold (fp32 compare, generalized_ssd.py:371) [+ legacy remove_small_boxes, box_head.py:370]
    auto consider = [&](int p, float s) {               // called by all lanes of a warp together
        bool pass = false;
        if (p < P) {
            pass = (s > score_thresh) && (s >= round_thr);
            if (pass && min_box_size >= 0.f) {
                const float4 q = bx[p];
                pass = (__fsub_rn(q.z, q.x) >= min_box_size) && (__fsub_rn(q.w, q.y) >= min_box_size);
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, pass);
        if (m == 0u) return;                             // the common case in the early rounds
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_n, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) keys[base + __popc(m & ((1u << lane) - 1u))] = make_key(s, (uint32_t)p);
    };
    if ((P & 1) == 0) {
        // rows are 8-byte aligned when P is even: 8-byte loads, eight of them in flight per thread (the slot order
        // inside `keys` is irrelevant, the list is sorted below)
        constexpr int NV = 8;
        const float2* sc2 = reinterpret_cast<const float2*>(sc);
        const int P2n = P >> 1;
        for (int q0 = 0; q0 < P2n; q0 += NV * P2_THREADS) {
            float2 sv[NV];
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int i = q0 + u * P2_THREADS + threadIdx.x;
                sv[u] = (i < P2n) ? __ldg(sc2 + i) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                if (q0 + u * P2_THREADS >= P2n) break;   // uniform
                const int i = q0 + u * P2_THREADS + threadIdx.x;
                const int p = (i < P2n) ? 2 * i : P;
                consider(p, sv[u].x);
                consider(p + (i < P2n ? 1 : 0), sv[u].y);
            }
        }
    } else {
        for (int q0 = 0; q0 < P; q0 += 4 * P2_THREADS) {
            float sv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {                // four independent loads in flight per thread
                const int p = q0 + u * P2_THREADS + threadIdx.x;
                sv[u] = (p < P) ? sc[p] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (q0 + u * P2_THREADS >= P) break;     // uniform
                consider(q0 + u * P2_THREADS + threadIdx.x, sv[u]);
            }
        }
    }
    __syncthreads();
    const int n = s_n;
    const size_t slot = (size_t)b * (K - 1) + c;
    if (n == 0) {
        if (threadIdx.x == 0) cand_counts[slot] = 0;
    } else {
        int n_pad = 32;
        while (n_pad < n) n_pad <<= 1;
        for (int i = n + threadIdx.x; i < n_pad; i += P2_THREADS) keys[i] = ~0ull;
        __syncthreads();
        bitonic_sort_u64(keys, n_pad);                       // (score desc, prior asc)
        const int m = (topk > 0 && topk < n) ? topk : n;     // top-k, generalized_ssd.py:376-378
        unsigned long long* dst = cand_keys + slot * cap;
        for (int i = threadIdx.x; i < m; i += P2_THREADS) dst[i] = keys[i];
        if (threadIdx.x == 0) cand_counts[slot] = m;
    }
    __syncthreads();                                         // keys / s_n are reused by the next image of this CTA
    }
}

// ---------------------------------------------------------------------------------------------
// Lazy NMS.  Round r only looks at the candidates of an image whose score is >= thr[b][r] (see
// pick_thresholds_kernel).  Those candidates are a PREFIX of every class's full sorted top-k list,
// greedy NMS on a prefix equals the full greedy NMS restricted to it, and every unprocessed candidate
// scores strictly below every processed one -- so once D detections survive among the processed ones
// they are exactly the reference's keep[:detections_per_img].  If fewer survive, the next round goes
// deeper (the last round processes everything).
// ---------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;
constexpr int NMS_WARP_MAX = 96;       // class lists up to this length run on one warp, longer ones on a CTA

// ---------------------------------------------------------------------------------------------
// P2b: greedy NMS, one WARP per (image, class) over this round's candidate list; stops at D kept.
// Lane = one candidate of the current 32-wide chunk; the kept list lives in this warp's slice of
// shared memory.  A candidate is kept iff no earlier kept box of its class overlaps it by more than
// the threshold (the greedy loop of the CPU kernel).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
class_nms_warp_kernel(const unsigned long long* __restrict__ cand_keys, const int* __restrict__ prefix,
                      const float4* __restrict__ boxes, Entry* __restrict__ out_entries, int* __restrict__ out_counts,
                      const int* __restrict__ done, int B, int P, int K, int cap, float thr_up, int D, int round,
                      int img_slots) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nc = K - 1;
    // warp = (image slot, class); round 0 has one slot per image, later rounds a few slots that walk the images
    // and skip the ones that are done
    const long long wg = (long long)blockIdx.x * warps + warp;
    if (wg >= (long long)img_slots * nc) return;
    const int cls = (int)(wg % nc);
    for (int b = (int)(wg / nc); b < B; b += img_slots) {
    if (round > 0 && done[b]) continue;
    const long long prob = (long long)b * nc + cls;
    // per-warp slices: kept boxes [D], chunk boxes [32], kept areas [D], chunk areas [32]
    float4* kept_box = reinterpret_cast<float4*>(s_raw) + (size_t)warp * (D + 32);
    float4* cbox = kept_box + D;
    float* kept_area = reinterpret_cast<float*>(reinterpret_cast<float4*>(s_raw) + (size_t)warps * (D + 32)) + (size_t)warp * (D + 32);
    float* carea = kept_area + D;
    const float margin = fabsf(thr_up) * 4e-6f + 1e-30f;
    const float thr_lo = thr_up - margin, thr_hi = thr_up + margin;
    const unsigned long long* keys = cand_keys + (size_t)prob * cap;
    const float4* bx = boxes + (size_t)b * P;
    Entry* dst = out_entries + (size_t)prob * D;
    const int q = prefix[prob];
    if (q > NMS_WARP_MAX) continue;                   // long lists are handled by class_nms_cta_kernel
    int nk = 0;
    for (int c0 = 0; c0 < q && nk < D; c0 += 32) {
        const int cnt = min(32, q - c0);
        unsigned long long kk = 0ull;
        float4 mb = make_float4(0.f, 0.f, 0.f, 0.f);
        float ma = 0.f;
        if (lane < cnt) {
            kk = keys[c0 + lane];
            mb = bx[key_index(kk)];
            ma = box_area(mb);
        }
        // candidate vs the kept list: independent tests (no early-out inside a group of 4 -> ILP)
        bool dead = lane >= cnt;
        int k = 0;
        for (; k + 4 <= nk; k += 4) {
            const bool d0 = iou_exceeds_fast(kept_box[k], kept_area[k], mb, ma, thr_up, thr_lo, thr_hi);
            const bool d1 = iou_exceeds_fast(kept_box[k + 1], kept_area[k + 1], mb, ma, thr_up, thr_lo, thr_hi);
            const bool d2 = iou_exceeds_fast(kept_box[k + 2], kept_area[k + 2], mb, ma, thr_up, thr_lo, thr_hi);
            const bool d3 = iou_exceeds_fast(kept_box[k + 3], kept_area[k + 3], mb, ma, thr_up, thr_lo, thr_hi);
            dead = dead || d0 || d1 || d2 || d3;
            if ((k & 31) == 28 && __all_sync(0xffffffffu, dead)) break;
        }
        for (; k < nk; ++k) dead = dead || iou_exceeds_fast(kept_box[k], kept_area[k], mb, ma, thr_up, thr_lo, thr_hi);
        unsigned int alive = ~__ballot_sync(0xffffffffu, dead);
        // dependencies inside the chunk: lane j first computes, in parallel, which earlier candidates of
        // the chunk would suppress it; the serial part is then one ballot per surviving candidate
        cbox[lane] = mb;
        carea[lane] = ma;
        __syncwarp();
        unsigned int sup_by = 0u;
        for (int i = 0; i < cnt; ++i) {
            const bool hit = (i < lane) && iou_exceeds_fast(cbox[i], carea[i], mb, ma, thr_up, thr_lo, thr_hi);
            sup_by |= hit ? (1u << i) : 0u;
        }
        __syncwarp();
        for (int i = 0; i < cnt; ++i) {
            if (!((alive >> i) & 1u)) continue;          // warp-uniform
            alive &= ~__ballot_sync(0xffffffffu, (sup_by >> i) & 1u);
        }
        const int rank = __popc(alive & ((1u << lane) - 1u));
        const bool mine = ((alive >> lane) & 1u) && (nk + rank < D);
        if (mine) {
            kept_box[nk + rank] = mb;
            kept_area[nk + rank] = ma;
            Entry e;
            e.score = key_score(kk);
            e.prior = (int)key_index(kk);
            dst[nk + rank] = e;
        }
        nk += __popc(__ballot_sync(0xffffffffu, mine));
        __syncwarp();
    }
    if (lane == 0) out_counts[prob] = nk;
    __syncwarp();
    }
}

// Long class lists (a dominant class easily holds a few hundred candidates) would make their warp the
// straggler of the whole launch, so they get a full CTA: every warp tests the 32-candidate chunk against a
// slice of the kept list (nms_consume_chunk), warp 0 resolves the chunk.
constexpr int NMS_CTA_THREADS = 256;
constexpr int NMS_CTA_CLASS_SLOTS = 4;      // CTAs per image of class_nms_cta_kernel (classes are dealt round-robin)
constexpr int LATE_ROUND_SLOTS = 8;         // image slots per class of the sort / warp-NMS launches of rounds >= 1
__global__ void __launch_bounds__(NMS_CTA_THREADS)
class_nms_cta_kernel(const unsigned long long* __restrict__ cand_keys, const int* __restrict__ prefix,
                     const float4* __restrict__ boxes, Entry* __restrict__ out_entries, int* __restrict__ out_counts,
                     const int* __restrict__ done, int P, int K, int cap, float thr_up, int D, int round) {
    pdl_trigger();
    pdl_wait();
    const int nc = K - 1;
    // CTA = (image, class residue): walks the classes c = blockIdx.y, blockIdx.y + gridDim.y, ... of its image and
    // only works on the long lists -- a grid of B x (K-1) CTAs that nearly all exit at once costs more than the NMS
    const int b = blockIdx.x;
    if (round > 0 && done[b]) return;
    extern __shared__ __align__(16) unsigned char s_raw[];
    float4* kept_box = reinterpret_cast<float4*>(s_raw);               // [D]
    float* kept_area = reinterpret_cast<float*>(kept_box + D);         // [D]
    __shared__ int s_nkept;
    __shared__ unsigned int s_dead;
    __shared__ float4 s_chunk_box[32];
    __shared__ float s_chunk_area[32];
    __shared__ unsigned long long s_chunk_key[32];
    const int lane = threadIdx.x & 31;
    const float4* bx = boxes + (size_t)b * P;
    for (int cls = blockIdx.y; cls < nc; cls += gridDim.y) {
    const int prob = b * nc + cls;
    const int q = prefix[prob];
    if (q <= NMS_WARP_MAX) continue;                   // uniform
    if (threadIdx.x == 0) {
        s_nkept = 0;
        s_dead = 0u;
    }
    __syncthreads();
    NmsState st{kept_box, kept_area, s_chunk_box, s_chunk_area, &s_dead, &s_nkept};
    const unsigned long long* keys = cand_keys + (size_t)prob * cap;
    Entry* dst = out_entries + (size_t)prob * D;
    for (int c0 = 0; c0 < q; c0 += 32) {
        const int cnt = min(32, q - c0);
        if (threadIdx.x < cnt) {
            const unsigned long long kk = keys[c0 + threadIdx.x];
            const float4 v = bx[key_index(kk)];
            s_chunk_key[threadIdx.x] = kk;
            s_chunk_box[threadIdx.x] = v;
            s_chunk_area[threadIdx.x] = box_area(v);
        }
        __syncthreads();
        const int nk_before = s_nkept;
        const unsigned int taken = nms_consume_chunk(st, cnt, thr_up, D);
        if (threadIdx.x < 32 && ((taken >> lane) & 1u)) {
            const unsigned long long kk = s_chunk_key[lane];
            Entry e;
            e.score = key_score(kk);
            e.prior = (int)key_index(kk);
            dst[nk_before + __popc(taken & ((1u << lane) - 1u))] = e;
        }
        if (s_nkept >= D) break;             // uniform: written before the closing barrier of the consumer
        __syncthreads();
    }
    if (threadIdx.x == 0) out_counts[prob] = s_nkept;
    __syncthreads();                         // the next class of this CTA resets s_nkept / the chunk buffers
    }
}

// ---------------------------------------------------------------------------------------------
// P3: one warp per image merges the (K-1) per-class kept lists (each already in descending
// order) and emits the first D by (score desc, class asc, rank asc) -- keep[:detections_per_img].
// The result is final when D detections survived or every class list was processed completely;
// otherwise done[b] stays 0 and the next round goes deeper.
// ---------------------------------------------------------------------------------------------
constexpr int P3_MAX_PER_LANE = 8;     // supports K-1 <= 256 classes
constexpr int P3_THREADS = 256;
constexpr int P3_SORT_CAP = 2048;      // kept entries per image that are staged and sorted in shared memory (rank < 4096,
                                       // class < 256 and slot < 4096 share the low 32 key bits)

__global__ void __launch_bounds__(P3_THREADS)
merge_topd_kernel(const Entry* __restrict__ entries, const int* __restrict__ counts, const float* __restrict__ thr,
                  const float4* __restrict__ boxes, float4* __restrict__ out_boxes, float* __restrict__ out_scores,
                  long long* __restrict__ out_labels, int* __restrict__ out_counts, int* __restrict__ out_priors,
                  int* __restrict__ done, int P, int K, int D, int round) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    if (round > 0 && done[b]) return;
    __shared__ Entry s_ent[P3_SORT_CAP];
    __shared__ int s_off[32 * P3_MAX_PER_LANE + 1];
    const int nc = K - 1;
    const int lane = threadIdx.x & 31;
    for (int c = threadIdx.x; c < nc; c += P3_THREADS) s_off[c + 1] = counts[b * nc + c];
    __syncthreads();
    if (threadIdx.x == 0) {                       // exclusive scan of the per-class kept counts (nc <= 256)
        int acc = 0;
        s_off[0] = 0;
        for (int c = 1; c <= nc; ++c) {
            acc += s_off[c];
            s_off[c] = acc;
        }
    }
    __syncthreads();
    const int total = s_off[nc];
    const bool incomplete = thr[b * NMS_ROUNDS + round] > 0.f;      // candidates below the round threshold exist
    if (total < D && incomplete) {
        if (threadIdx.x == 0) done[b] = 0;
        return;
    }
    // Usual case (the kept entries of the image fit in shared memory): stage them, sort 64-bit keys
    // (score desc | class asc | rank-in-class asc | slot) with the whole CTA and emit the first D in parallel.
    __shared__ __align__(16) unsigned long long s_key[P3_SORT_CAP];      // aliased by s_sel_prior in the serial path
    if (total <= P3_SORT_CAP) {
        int n_pad = 32;
        while (n_pad < total) n_pad <<= 1;
        for (int f = threadIdx.x; f < n_pad; f += P3_THREADS) {
            unsigned long long key = ~0ull;
            if (f < total) {
                int lo = 0, hi = nc;              // class whose [off, off + cnt) contains f
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= f) lo = mid;
                    else hi = mid;
                }
                const int rank = f - s_off[lo];
                const Entry e = entries[((size_t)b * nc + lo) * D + rank];
                s_ent[f] = e;
                key = ((unsigned long long)(~orderable(e.score)) << 32) |
                      ((unsigned long long)lo << 24) | ((unsigned long long)rank << 12) | (unsigned long long)f;
            }
            s_key[f] = key;
        }
        __syncthreads();
        bitonic_sort_u64(s_key, n_pad);
        const int nsel = min(D, total);
        for (int i = threadIdx.x; i < D; i += P3_THREADS) {
            if (i < nsel) {
                const unsigned long long key = s_key[i];
                const Entry e = s_ent[(int)(key & 0xfffu)];
                out_scores[(size_t)b * D + i] = e.score;
                out_labels[(size_t)b * D + i] = (long long)((int)((key >> 24) & 0xffu) + 1);
                out_boxes[(size_t)b * D + i] = boxes[(size_t)b * P + e.prior];
                if (out_priors) out_priors[(size_t)b * D + i] = e.prior;
            } else {
                out_scores[(size_t)b * D + i] = 0.f;
                out_labels[(size_t)b * D + i] = 0;
                out_boxes[(size_t)b * D + i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (out_priors) out_priors[(size_t)b * D + i] = -1;
            }
        }
        if (threadIdx.x == 0) {
            done[b] = 1;
            out_counts[b] = nsel;
        }
        return;
    }
    // Rare case (deep rounds of the stress configuration): serial k-way merge by one warp straight from global memory
    const bool staged = false;
    int* s_sel_prior = reinterpret_cast<int*>(s_key);       // detections_per_img <= 4096 = 2 * P3_SORT_CAP ints
    __shared__ int s_nsel;
    if (threadIdx.x < 32) {
    if (lane == 0) done[b] = 1;
    int pos[P3_MAX_PER_LANE], cnt[P3_MAX_PER_LANE];
    const Entry* list[P3_MAX_PER_LANE];
    Entry head[P3_MAX_PER_LANE];
#pragma unroll
    for (int i = 0; i < P3_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        pos[i] = 0;
        cnt[i] = (c < nc) ? s_off[c + 1] - s_off[c] : 0;
        list[i] = (c < nc) ? (staged ? s_ent + s_off[c] : entries + ((size_t)b * nc + c) * D) : entries;
        head[i].score = 0.f;
        head[i].prior = 0;
        if (cnt[i] > 0) head[i] = list[i][0];
    }
    int d = 0;
    for (; d < D; ++d) {
        unsigned long long best = 0ull;
#pragma unroll
        for (int i = 0; i < P3_MAX_PER_LANE; ++i) {
            if (pos[i] < cnt[i]) {
                unsigned long long k = ((unsigned long long)orderable(head[i].score) << 32) |
                                       (0xffffffffu - (uint32_t)(lane + 32 * i));
                best = (k > best) ? k : best;
            }
        }
        unsigned long long w = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long t = __shfl_xor_sync(0xffffffffu, w, o);
            w = (t > w) ? t : w;
        }
        if (w == 0ull) break;
        const int cls = (int)(0xffffffffu - (uint32_t)w);
        if ((cls & 31) == lane) {
            const int i = cls >> 5;
            Entry e = head[0];
#pragma unroll
            for (int q = 1; q < P3_MAX_PER_LANE; ++q)
                if (q == i) e = head[q];
            out_scores[(size_t)b * D + d] = e.score;
            out_labels[(size_t)b * D + d] = (long long)(cls + 1);
            s_sel_prior[d] = e.prior;              // the box gather happens after the serial loop, in parallel
#pragma unroll
            for (int q = 0; q < P3_MAX_PER_LANE; ++q) {
                if (q == i) {
                    pos[q] += 1;
                    if (pos[q] < cnt[q]) head[q] = list[q][pos[q]];
                }
            }
        }
    }
    if (lane == 0) {
        out_counts[b] = d;
        s_nsel = d;
    }
    }
    __syncthreads();
    const int nsel = s_nsel;
    for (int i = threadIdx.x; i < D; i += P3_THREADS) {
        if (i < nsel) {
            out_boxes[(size_t)b * D + i] = boxes[(size_t)b * P + s_sel_prior[i]];
            if (out_priors) out_priors[(size_t)b * D + i] = s_sel_prior[i];
        } else {
            out_scores[(size_t)b * D + i] = 0.f;
            out_labels[(size_t)b * D + i] = 0;
            out_boxes[(size_t)b * D + i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (out_priors) out_priors[(size_t)b * D + i] = -1;
        }
    }
}

// parity aid (dn_postprocess_scored): which lazy round made image b final (0, 1 or 2)
__global__ void mark_rounds_kernel(const int* __restrict__ done, int* __restrict__ rounds, int B, int round) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && (round == 0 || rounds[b] == round) ) rounds[b] = done[b] ? round : round + 1;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static float threshold_up(double thr) {
    // smallest float f with (double)f > thr
    float f = (float)thr;
    if ((double)f > thr) {
        float g = nextafterf(f, -INFINITY);
        while ((double)g > thr) {
            f = g;
            g = nextafterf(f, -INFINITY);
        }
        return f;
    }
    while (!((double)f > thr)) f = nextafterf(f, INFINITY);
    return f;
}

static int next_pow2(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

// rounds of the lazy NMS: process about this many top-scoring candidates per image, go deeper only
// for the images where fewer than D detections survived (-1 = everything)
static RoundTargets round_targets(int D) {
    RoundTargets r;
    r.t[0] = D * 4 > 1024 ? D * 4 : 1024;
    r.t[1] = r.t[0] * 8;
    r.t[2] = 0;               // everything
    return r;
}

static int cand_cap(const dn_postprocess_params* p) {
    return (p->topk_candidates > 0 && p->topk_candidates < p->num_priors) ? p->topk_candidates : p->num_priors;
}

struct PostLayout {
    size_t scores_off, boxes_off, cand_off, ccount_off, hist_off, thr_off, entries_off, counts_off, done_off, total;
};
static PostLayout post_layout(int B, const dn_postprocess_params* p) {
    PostLayout L;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t P = p->num_priors, K = p->num_classes, D = p->detections_per_img, cap = cand_cap(p);
    size_t off = 0;
    L.scores_off = off; off = al(off + (size_t)B * (K - 1) * P * sizeof(float));
    L.boxes_off = off;  off = al(off + (size_t)B * P * sizeof(float4));
    L.cand_off = off;   off = al(off + (size_t)B * (K - 1) * cap * sizeof(unsigned long long));
    L.ccount_off = off; off = al(off + (size_t)B * (K - 1) * sizeof(int));
    L.hist_off = off;   off = al(off + (size_t)B * HIST_BINS * sizeof(int));
    L.thr_off = off;    off = al(off + (size_t)B * NMS_ROUNDS * sizeof(float));
    L.entries_off = off; off = al(off + (size_t)B * (K - 1) * D * sizeof(Entry));
    L.counts_off = off; off = al(off + (size_t)B * (K - 1) * sizeof(int));
    L.done_off = off;   off = al(off + (size_t)B * sizeof(int));
    L.total = off;
    return L;
}

static int nms_warps_per_cta(int D) {
    int w = (int)((160 * 1024) / ((size_t)(D + 32) * 20));
    return w >= 8 ? 8 : (w < 1 ? 1 : w);
}

}  // namespace dn

using namespace dn;

extern "C" size_t dn_postprocess_workspace_bytes(int B, const dn_postprocess_params* p) {
    if (!p || B <= 0) return 0;
    return post_layout(B, p).total;
}

static int validate_post(const dn_postprocess_params* p) {
    DN_REQUIRE(p != nullptr, DN_ERR_INVALID, "postprocess params are NULL");
    DN_REQUIRE(p->num_priors > 0 && p->num_classes >= 2, DN_ERR_INVALID, "need num_priors > 0 and num_classes >= 2");
    DN_REQUIRE(p->num_classes - 1 <= 32 * P3_MAX_PER_LANE, DN_ERR_UNSUPPORTED, "at most %d foreground classes",
               32 * P3_MAX_PER_LANE);
    DN_REQUIRE(p->detections_per_img > 0 && p->detections_per_img <= 4096, DN_ERR_UNSUPPORTED,
               "detections_per_img must be in [1,4096]");
    DN_REQUIRE(p->num_priors <= 16384, DN_ERR_UNSUPPORTED, "at most 16384 priors per image");
    DN_REQUIRE(p->num_classes - 1 <= SEL_THREADS, DN_ERR_UNSUPPORTED, "at most %d foreground classes", SEL_THREADS);
    return DN_OK;
}

static int postprocess_impl(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                            const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                            float* out_scores, int64_t* out_labels, int32_t* out_counts, cudaStream_t stream, int iters,
                            float* ms3, bool scored = false, int32_t* out_priors = nullptr, int32_t* out_rounds = nullptr,
                            cudaEvent_t after_front = nullptr) {
    int rc = validate_post(p);
    if (rc) return rc;
    DN_REQUIRE(B > 0, DN_ERR_INVALID, "batch must be positive");
    DN_REQUIRE(cls_logits && bbox_regression && (anchors || scored) && out_boxes && out_scores && out_labels && out_counts,
               DN_ERR_INVALID, "NULL tensor pointer");
    const PostLayout L = post_layout(B, p);
    DN_REQUIRE(workspace && workspace_bytes >= L.total, DN_ERR_WORKSPACE, "workspace too small: need %zu bytes", L.total);
    const int P = p->num_priors, K = p->num_classes, D = p->detections_per_img;
    unsigned char* ws = (unsigned char*)workspace;
    float* scores_t = (float*)(ws + L.scores_off);
    float4* boxes = (float4*)(ws + L.boxes_off);
    Entry* entries = (Entry*)(ws + L.entries_off);
    int* counts = (int*)(ws + L.counts_off);
    unsigned long long* cand = (unsigned long long*)(ws + L.cand_off);
    int* ccount = (int*)(ws + L.ccount_off);
    int* hist = (int*)(ws + L.hist_off);
    float* thr = (float*)(ws + L.thr_off);
    int* done = (int*)(ws + L.done_off);
    const int cap = cand_cap(p);
    const size_t smem1 = (size_t)P1_ROWS * (K + 1) * sizeof(float);
    static SmemOptIn optin_p1, optin_p1s, optin_sort, optin_warp, optin_cta;
    DN_CHECK_CUDA(optin_p1.ensure(softmax_decode_kernel<false>, smem1));
    DN_CHECK_CUDA(optin_p1s.ensure(softmax_decode_kernel<true>, smem1));
    const size_t smem_sort = (size_t)next_pow2(P) * 8;
    const int nms_warps = nms_warps_per_cta(D);
    const size_t smem_nms = (size_t)nms_warps * (D + 32) * 20;
    DN_CHECK_CUDA(optin_sort.ensure(class_sort_kernel, smem_sort));
    DN_CHECK_CUDA(optin_cta.ensure(class_nms_cta_kernel, (size_t)D * 20));
    DN_CHECK_CUDA(optin_warp.ensure(class_nms_warp_kernel, smem_nms));
    const float thr_up = threshold_up(p->nms_thresh);
    const RoundTargets targets = round_targets(D);
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (ms3) {
        DN_CHECK_CUDA(cudaEventCreate(&ev[0]));
        DN_CHECK_CUDA(cudaEventCreate(&ev[1]));
    }
    // phase 0: softmax + decode (+ histogram, round thresholds); phase 1: per-round sort and NMS; phase 2: merges.
    // (the profiling variant re-runs each phase `iters` times, which is idempotent)
    for (int phase = 0; phase < 3; ++phase) {
        if (ms3) DN_CHECK_CUDA(cudaEventRecord(ev[0], stream));
        for (int it = 0; it < iters; ++it) {
            if (phase == 0 || !ms3) {
                DN_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)B * HIST_BINS * sizeof(int), stream));
                dim3 grid(ceil_div(P, P1_ROWS), B);
                if (scored)
                    softmax_decode_kernel<true><<<grid, P1_THREADS, smem1, stream>>>(cls_logits, bbox_regression, anchors,
                                                                                    scores_t, boxes, hist, P, K, *p);
                else
                    softmax_decode_kernel<false><<<grid, P1_THREADS, smem1, stream>>>(cls_logits, bbox_regression, anchors,
                                                                                     scores_t, boxes, hist, P, K, *p);
                DN_CHECK_LAUNCH();
                launch_pdl(pick_thresholds_kernel, B, 32, 0, stream, (const int*)hist, thr, targets);
                DN_CHECK_LAUNCH();
                if (after_front) DN_CHECK_CUDA(cudaEventRecordWithFlags(after_front, stream, cudaEventRecordExternal));
            }
            for (int r = 0; r < NMS_ROUNDS; ++r) {
                if (phase == 1 || !ms3) {
                    const int slots = (r == 0) ? B : std::min(B, LATE_ROUND_SLOTS);
                    dim3 grid(K - 1, slots);
                    launch_pdl(class_sort_kernel, grid, P2_THREADS, smem_sort, stream, scores_t, boxes, cand, ccount, P, K, cap,
                               p->score_thresh, p->topk_candidates, p->min_box_size, thr, done, r, B);
                    DN_CHECK_LAUNCH();
                    launch_pdl(class_nms_warp_kernel, (unsigned)ceil_div<long long>((long long)slots * (K - 1), nms_warps),
                               nms_warps * 32, smem_nms, stream, cand, ccount, boxes, entries, counts, done, B, P, K, cap, thr_up,
                               D, r, slots);
                    DN_CHECK_LAUNCH();
                    launch_pdl(class_nms_cta_kernel, dim3(B, std::min(K - 1, NMS_CTA_CLASS_SLOTS)), NMS_CTA_THREADS, (size_t)D * 20,
                               stream, cand, ccount, boxes, entries, counts, done, P, K, cap, thr_up, D, r);
                    DN_CHECK_LAUNCH();
                }
                if (phase == 2 || !ms3) {
                    launch_pdl(merge_topd_kernel, B, P3_THREADS, 0, stream, entries, counts, thr, boxes, (float4*)out_boxes,
                               out_scores, (long long*)out_labels, out_counts, out_priors, done, P, K, D, r);
                    DN_CHECK_LAUNCH();
                    // parity aid: after round r, rounds[b] = r + 1 for every image that is still not final
                    if (out_rounds)
                        mark_rounds_kernel<<<ceil_div(B, 256), 256, 0, stream>>>(done, out_rounds, B, r);
                }
            }
            if (!ms3) break;
        }
        if (ms3) {
            DN_CHECK_CUDA(cudaEventRecord(ev[1], stream));
            DN_CHECK_CUDA(cudaEventSynchronize(ev[1]));
            DN_CHECK_CUDA(cudaEventElapsedTime(&ms3[phase], ev[0], ev[1]));
            ms3[phase] /= iters;
        } else {
            break;
        }
    }
    if (ms3) {
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
    }
    return DN_OK;
}

extern "C" int dn_postprocess(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                              const dn_postprocess_params* p, void* workspace, size_t workspace_bytes,
                              float* out_boxes, float* out_scores, int64_t* out_labels, int32_t* out_counts,
                              void* stream_) {
    return postprocess_impl(cls_logits, bbox_regression, anchors, B, p, workspace, workspace_bytes, out_boxes, out_scores,
                            out_labels, out_counts, (cudaStream_t)stream_, 1, nullptr);
}

extern "C" int dn_postprocess_scored(const float* scores, const float* boxes, int B, const dn_postprocess_params* p,
                                     void* workspace, size_t workspace_bytes, float* out_boxes, float* out_scores,
                                     int64_t* out_labels, int32_t* out_counts, int32_t* out_priors, int32_t* out_rounds,
                                     void* stream_) {
    return postprocess_impl(scores, boxes, nullptr, B, p, workspace, workspace_bytes, out_boxes, out_scores, out_labels,
                            out_counts, (cudaStream_t)stream_, 1, nullptr, true, out_priors, out_rounds);
}

int dn_postprocess_marked(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                          const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                          float* out_scores, int64_t* out_labels, int32_t* out_counts, cudaStream_t stream,
                          cudaEvent_t after_front) {
    return postprocess_impl(cls_logits, bbox_regression, anchors, B, p, workspace, workspace_bytes, out_boxes, out_scores,
                            out_labels, out_counts, stream, 1, nullptr, false, nullptr, nullptr, after_front);
}

// measurement aid used by dn_engine_profile / dn_postprocess_profile: mean ms of P1, P2, P3
int dn_postprocess_timed(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                         const dn_postprocess_params* p, void* workspace, size_t workspace_bytes, float* out_boxes,
                         float* out_scores, int64_t* out_labels, int32_t* out_counts, cudaStream_t stream, int iters,
                         float* ms3) {
    return postprocess_impl(cls_logits, bbox_regression, anchors, B, p, workspace, workspace_bytes, out_boxes, out_scores,
                            out_labels, out_counts, stream, iters, ms3);
}

extern "C" int dn_postprocess_profile(const float* cls_logits, const float* bbox_regression, const float* anchors, int B,
                                      const dn_postprocess_params* p, void* workspace, size_t workspace_bytes,
                                      float* out_boxes, float* out_scores, int64_t* out_labels, int32_t* out_counts,
                                      int iters, float* ms3_host, void* stream_) {
    DN_REQUIRE(iters > 0 && ms3_host, DN_ERR_INVALID, "bad iters / NULL output");
    return postprocess_impl(cls_logits, bbox_regression, anchors, B, p, workspace, workspace_bytes, out_boxes, out_scores,
                            out_labels, out_counts, (cudaStream_t)stream_, iters, ms3_host);
}

// =============================================================================================
// dn_batched_nms: torchvision.ops.batched_nms, per-class ("vanilla") semantics, arbitrary n.
//   1. global stable sort of (score desc, index asc) keys  -> visiting order
//   2. one CTA per class streams the sorted keys, picks its members in order and runs the same
//      greedy consumer; kept flags are written per rank
//   3. order-preserving compaction of the flagged ranks -> keep[] (already in descending score)
// =============================================================================================
namespace dn {

constexpr int BN_MAX_LABEL = 4096;
constexpr int BN_MAX_KEEP = 4096;
constexpr int BN_THREADS = 256;

__global__ void bnms_prepare_kernel(const float* __restrict__ scores, const long long* __restrict__ idxs,
                                    long long n, long long n_pad, unsigned long long* __restrict__ keys,
                                    int* __restrict__ hist, int* __restrict__ err) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pad; i += (long long)gridDim.x * blockDim.x) {
        if (i < n) {
            keys[i] = make_key(scores[i], (uint32_t)i);
            long long l = idxs[i];
            if (l < 0 || l >= BN_MAX_LABEL) atomicExch(err, 1);
            else atomicAdd(&hist[l], 1);
        } else {
            keys[i] = ~0ull;
        }
    }
}

// one bitonic (k, j) step over global memory
__global__ void bnms_bitonic_step_kernel(unsigned long long* __restrict__ keys, long long n_pad, long long k, long long j) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (n_pad >> 1); t += (long long)gridDim.x * blockDim.x) {
        long long lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        long long hi = lo | j;
        unsigned long long a = keys[lo], b = keys[hi];
        bool up = ((lo & k) == 0);
        if ((a > b) == up) {
            keys[lo] = b;
            keys[hi] = a;
        }
    }
}

// every (k, j) step with j < BN_TILE is tile-local: phases k_lo..k_hi run inside shared memory
constexpr int BN_TILE = 4096;
__global__ void __launch_bounds__(1024)
bnms_bitonic_tile_kernel(unsigned long long* __restrict__ keys, long long k_lo, long long k_hi) {
    __shared__ unsigned long long s[BN_TILE];
    const long long base = (long long)blockIdx.x * BN_TILE;
    for (int i = threadIdx.x; i < BN_TILE; i += blockDim.x) s[i] = keys[base + i];
    __syncthreads();
    for (long long k = k_lo; k <= k_hi; k <<= 1) {
        const int j0 = (int)((k >> 1) < (BN_TILE / 2) ? (k >> 1) : (BN_TILE / 2));
        for (int j = j0; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < BN_TILE / 2; t += blockDim.x) {
                int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int hi = lo | j;
                unsigned long long a = s[lo], b = s[hi];
                bool up = (((base + lo) & k) == 0);
                if ((a > b) == up) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < BN_TILE; i += blockDim.x) keys[base + i] = s[i];
}

__global__ void __launch_bounds__(BN_THREADS)
bnms_class_kernel(const float4* __restrict__ boxes, const long long* __restrict__ idxs,
                  const unsigned long long* __restrict__ keys, long long n, const int* __restrict__ hist,
                  unsigned char* __restrict__ kept_flag, int* __restrict__ err, float thr_up) {
    const int cls = blockIdx.x;
    const int members = hist[cls];
    if (members == 0) return;
    extern __shared__ __align__(16) unsigned char s_raw[];
    float4* kept_box = reinterpret_cast<float4*>(s_raw);                    // [BN_MAX_KEEP]
    float* kept_area = reinterpret_cast<float*>(kept_box + BN_MAX_KEEP);    // [BN_MAX_KEEP]
    __shared__ int s_nkept, s_cnt;
    __shared__ unsigned int s_dead;
    __shared__ float4 s_chunk_box[32];
    __shared__ float s_chunk_area[32];
    __shared__ long long s_chunk_rank[32];
    __shared__ int s_warp_cnt[BN_THREADS / 32];
    if (threadIdx.x == 0) {
        s_nkept = 0;
        s_dead = 0u;
        s_cnt = 0;
    }
    __syncthreads();
    NmsState st{kept_box, kept_area, s_chunk_box, s_chunk_area, &s_dead, &s_nkept};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int seen = 0;
    // stream ranks in tiles of BN_THREADS; members of this class are compacted IN ORDER into chunks of 32
    for (long long r0 = 0; r0 < n && seen < members; r0 += BN_THREADS) {
        const long long r = r0 + threadIdx.x;
        bool mine = false;
        uint32_t idx = 0;
        if (r < n) {
            idx = key_index(keys[r]);
            mine = (idxs[idx] == (long long)cls);
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, mine);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < BN_THREADS / 32; ++w) {
            int cw = s_warp_cnt[w];
            if (w < warp) before += cw;
            total += cw;
        }
        const int my_pos = before + __popc(bal & ((1u << lane) - 1u));    // order-preserving position in this tile
        __syncthreads();                                                   // s_warp_cnt is rewritten next tile
        // feed the tile's members to the consumer in groups that complete 32-wide chunks
        int fed = 0;
        while (fed < total) {
            // s_cnt is read ONCE per iteration, here, behind the barrier that closed the previous one: thread 0
            // rewrites it below, possibly before a slower warp would get to a second read
            const int cur = s_cnt;
            const int room = 32 - cur;
            const int take = min(room, total - fed);
            if (mine && my_pos >= fed && my_pos < fed + take) {
                const int slot = cur + (my_pos - fed);
                const float4 q = boxes[idx];
                s_chunk_box[slot] = q;
                s_chunk_area[slot] = box_area(q);
                s_chunk_rank[slot] = r;
            }
            __syncthreads();
            const int cnt = cur + take;
            fed += take;
            const bool last_of_class = (seen + fed == members);
            if (cnt == 32 || last_of_class) {
                const unsigned int taken = nms_consume_chunk(st, cnt, thr_up, BN_MAX_KEEP);
                if (threadIdx.x < 32 && ((taken >> lane) & 1u)) kept_flag[s_chunk_rank[lane]] = 1;
                if (threadIdx.x == 0) s_cnt = 0;
                if (s_nkept >= BN_MAX_KEEP && !last_of_class) {             // uniform: kept list is full
                    if (threadIdx.x == 0) atomicExch(err, 2);
                    return;
                }
            } else {
                if (threadIdx.x == 0) s_cnt = cnt;
            }
            __syncthreads();
        }
        seen += total;
    }
}

__global__ void __launch_bounds__(1024)
bnms_compact_kernel(const unsigned long long* __restrict__ keys, const unsigned char* __restrict__ kept_flag,
                    long long n, long long* __restrict__ keep_out, long long* __restrict__ nkeep_out,
                    const int* __restrict__ err) {
    __shared__ int s_warp[32];
    __shared__ long long s_base;
    if (*err != 0) {
        if (threadIdx.x == 0) *nkeep_out = -(long long)(*err);
        return;
    }
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long r0 = 0; r0 < n; r0 += 1024) {
        const long long r = r0 + threadIdx.x;
        const bool f = (r < n) && kept_flag[r];
        const unsigned int bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            int cw = s_warp[w];
            if (w < warp) before += cw;
            total += cw;
        }
        if (f) keep_out[s_base + before + __popc(bal & ((1u << lane) - 1u))] = (long long)key_index(keys[r]);
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *nkeep_out = s_base;
}

struct BnmsLayout {
    size_t keys_off, hist_off, flag_off, err_off, total;
    long long n_pad;
};
static BnmsLayout bnms_layout(long long n) {
    BnmsLayout L;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    long long n_pad = BN_TILE;
    while (n_pad < n) n_pad <<= 1;
    L.n_pad = n_pad;
    size_t off = 0;
    L.keys_off = off; off = al(off + (size_t)n_pad * 8);
    L.hist_off = off; off = al(off + (size_t)BN_MAX_LABEL * sizeof(int));
    L.flag_off = off; off = al(off + (size_t)(n > 0 ? n : 1));
    L.err_off = off;  off = al(off + sizeof(int));
    L.total = off;
    return L;
}

}  // namespace dn

extern "C" size_t dn_batched_nms_workspace_bytes(int64_t n) { return bnms_layout(n < 0 ? 0 : n).total; }

extern "C" int dn_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n,
                              double iou_threshold, void* workspace, size_t workspace_bytes, int64_t* keep_out,
                              int64_t* nkeep_out, void* stream_) {
    DN_REQUIRE(n >= 0 && n < (1ll << 31), DN_ERR_INVALID, "n out of range");
    DN_REQUIRE(nkeep_out != nullptr, DN_ERR_INVALID, "nkeep_out is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) {
        DN_CHECK_CUDA(cudaMemsetAsync(nkeep_out, 0, sizeof(int64_t), stream));
        return DN_OK;
    }
    DN_REQUIRE(boxes && scores && idxs && keep_out, DN_ERR_INVALID, "NULL tensor pointer");
    const BnmsLayout L = bnms_layout(n);
    DN_REQUIRE(workspace && workspace_bytes >= L.total, DN_ERR_WORKSPACE, "workspace too small: need %zu bytes", L.total);
    unsigned char* ws = (unsigned char*)workspace;
    unsigned long long* keys = (unsigned long long*)(ws + L.keys_off);
    int* hist = (int*)(ws + L.hist_off);
    unsigned char* flag = ws + L.flag_off;
    int* err = (int*)(ws + L.err_off);
    DN_CHECK_CUDA(cudaMemsetAsync(ws + L.hist_off, 0, L.total - L.hist_off, stream));
    const int blocks = (int)std::min<long long>(ceil_div<long long>(L.n_pad, 256), 148 * 8);
    bnms_prepare_kernel<<<blocks, 256, 0, stream>>>(scores, (const long long*)idxs, n, L.n_pad, keys, hist, err);
    DN_CHECK_LAUNCH();
    // bitonic sort: phases k <= BN_TILE in one in-smem launch; then, per larger k, global steps
    // while j >= BN_TILE followed by one in-smem launch for the remaining j
    const unsigned tiles = (unsigned)(L.n_pad / BN_TILE);
    bnms_bitonic_tile_kernel<<<tiles, 1024, 0, stream>>>(keys, 2, BN_TILE);
    DN_CHECK_LAUNCH();
    for (long long k = 2 * (long long)BN_TILE; k <= L.n_pad; k <<= 1) {
        for (long long j = k >> 1; j >= BN_TILE; j >>= 1) {
            bnms_bitonic_step_kernel<<<blocks, 256, 0, stream>>>(keys, L.n_pad, k, j);
            DN_CHECK_LAUNCH();
        }
        bnms_bitonic_tile_kernel<<<tiles, 1024, 0, stream>>>(keys, k, k);
        DN_CHECK_LAUNCH();
    }
    const size_t smem = (size_t)BN_MAX_KEEP * (sizeof(float4) + sizeof(float));
    static SmemOptIn optin_bnms;
    DN_CHECK_CUDA(optin_bnms.ensure(bnms_class_kernel, smem));
    bnms_class_kernel<<<BN_MAX_LABEL, BN_THREADS, smem, stream>>>((const float4*)boxes, (const long long*)idxs, keys, n,
                                                                hist, flag, err, threshold_up(iou_threshold));
    DN_CHECK_LAUNCH();
    bnms_compact_kernel<<<1, 1024, 0, stream>>>(keys, flag, n, (long long*)keep_out, (long long*)nkeep_out, err);
    DN_CHECK_LAUNCH();
    return DN_OK;
}
