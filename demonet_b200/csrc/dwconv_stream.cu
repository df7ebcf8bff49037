// Depthwise k x k stride-1 stencil as a TMA-fed ROW STREAM for sm_100a.
// Reference: ConvBNActivation(groups=C), demonet/models/mobilenetv2.py:32-55 (see dn_dwconv).
//
// CTA c works on channel block c % ncblk for its whole life (taps and bias staged once); the output rows of all
// images form one long stream per channel block and the CTAs of a block own equal contiguous shares of it, so the
// GPU is balanced to within one row, a CTA pays the k-1 halo rows once per image it touches, and the CTAs of the
// different channel blocks walk the same images at the same time (their interleaved reads meet in L2).  Inside a CTA, warp 0 is the producer: one thread issues cp.async.bulk.tensor.4d loads of k input
// rows x the full padded width x CB channels into a ring of shared-memory stages (zero padding = the tensor
// map's out-of-bounds fill, no boundary branches), running ahead of the consumers by the depth of the ring, so
// the bytes in flight are set by shared memory rather than by registers.  Consumer thread = 4 channels x TW
// output columns, walking down the rows: each staged input vector is read and unpacked to fp32 once and
// scattered into a ring of k output-row accumulators in registers (packed FFMA2); the row whose last input row
// just arrived is activated, packed to bf16 and stored, and its slot restarts from the bias.
#include <cuda.h>

#include "common.cuh"
#include "dwconv.cuh"

namespace dn {

__device__ __forceinline__ uint32_t dws_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ACT>
__device__ __forceinline__ float dws_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));      // x * relu6(x + 3) / 6
    return v;
}

template <int NS = DN_SLEEP_CONSUMER>
__device__ __forceinline__ void dws_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) break;
        spin_backoff<NS>();
    }
}

// taps are re-read from shared memory at every use: `volatile` keeps ptxas from hoisting all k*k of them into
// registers (25 x 4 for a 5x5), which costs the occupancy this kernel lives on
__device__ __forceinline__ float4 dws_lds128(uint32_t addr) {
    float4 f;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(addr));
    return f;
}

__device__ __forceinline__ uint2 dws_lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// one staged pixel of a thread: NP packed bf16 pairs (8 or 4 bytes)
template <int NP>
__device__ __forceinline__ void dws_lds_px(uint32_t addr, uint32_t* v) {
    if (NP == 2) {
        const uint2 t = dws_lds64(addr);
        v[0] = t.x, v[NP - 1] = t.y;
    } else {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[0]) : "r"(addr));
    }
}
// one tap of a thread's channels: NP float pairs (16 or 8 bytes)
template <int NP>
__device__ __forceinline__ void dws_lds_tap(uint32_t addr, float2* wv) {
    if (NP == 2) {
        const float4 f = dws_lds128(addr);
        wv[0] = make_float2(f.x, f.y), wv[NP - 1] = make_float2(f.z, f.w);
    } else {
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(wv[0].x), "=f"(wv[0].y) : "r"(addr));
    }
}

constexpr int DWS_MAX_STAGES = 8;

// CH = channels per consumer thread (4 or 2).  Every tap is read from shared memory once per input row and thread and
// feeds TW * CH / 2 packed FMAs: the 5x5 kernel runs CH = 2, TW = 4 (8-byte tap loads, 4 FFMA2 each) because with
// CH = 4, TW = 2 (16-byte tap loads, 4 FFMA2 each) the tap traffic alone oversubscribes the shared-memory pipe 2x.
//
// POOL: the kernel also leaves the per-channel sums of its outputs for the squeeze-excitation that follows (SE pools the
// depthwise output, mobilenetv3.py:22-40), saving that layer's read of the tensor.  Every CTA share of an image's rows
// writes one slot pool[b][slot][c], slot = this CTA's index among the CTAs that touch image b (se_fc1_kernel derives
// the slot count of an image from the same split), summed in a fixed order: deterministic, no atomics.
template <int KS, int TW, int CH, int ACT, int NT, bool POOL>
__global__ void __launch_bounds__(NT, NT <= 160 ? 3 : 1)
dwconv_stream_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ w, const float* __restrict__ bias,
                     uint32_t* __restrict__ y, DwStream sp, int B, int H, int W, int C, float* __restrict__ pool, int pool_stride) {
    constexpr int P = KS / 2, TWIN = TW + KS - 1, NP = CH / 2;
    extern __shared__ __align__(128) unsigned char dws_smem[];
    __shared__ uint64_t full[DWS_MAX_STAGES], empty[DWS_MAX_STAGES];
    unsigned char* stages = dws_smem;
    float* wsm = reinterpret_cast<float*>(dws_smem + (size_t)sp.nst * sp.stage_stride);     // [KS*KS][CB]
    float* psm = wsm + KS * KS * sp.CB;                                                      // POOL: [ncb][CB]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_cons_warps = (blockDim.x >> 5) - 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < sp.nst; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dws_u32(&full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dws_u32(&empty[s])), "r"(n_cons_warps) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // taps of this CTA's channel block, staged once by ALL threads and BEFORE the programmatic-dependency wait (weights are
    // not produced by the previous kernel, so this runs while that kernel drains).  Eight independent loads per thread and
    // round: one dependent load per round had every CTA spend 5-10 us here (ncu, 10 x 10 x 480 k5: 10 % of the samples on
    // the store behind the load).  smem layout [channel group][tap][CH]: a thread's k*k taps are contiguous, so every tap
    // is base + immediate offset (no per-tap address registers); lanes are k*k*16 bytes apart = 4 banks mod 32,
    // conflict-free for LDS.128
    {
        const float* wsrc = w + (blockIdx.x % sp.ncblk) * sp.CB;
        const int total = KS * KS * sp.CB;
        for (int i0 = threadIdx.x; i0 < total; i0 += 8 * blockDim.x) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * blockDim.x;
                v[u] = (i < total) ? __ldg(wsrc + (i / sp.CB) * C + (i % sp.CB)) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < total) {
                    const int c = i % sp.CB, t = i / sp.CB;
                    wsm[((c / CH) * (KS * KS) + t) * CH + (c % CH)] = v[u];
                }
            }
        }
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();                 // the barriers are set up and the taps staged; nothing above touches activation memory

    // this CTA's channel block and its share of that block's output-row stream (B * H rows)
    // (wide maps: the width is cut into sp.nstrip column strips and a CTA also keeps its strip for its whole life)
    const int ncs = sp.ncblk * sp.nstrip;
    const int cblk = blockIdx.x % sp.ncblk;
    const int col0 = ((blockIdx.x / sp.ncblk) % sp.nstrip) * sp.ncb * TW;       // first output column of the strip
    const long long T = (long long)B * H;
    const int part = blockIdx.x / ncs, parts = gridDim.x / ncs;
    const long long g_begin = T * part / parts, g_end = T * (part + 1) / parts;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t seq = 0;
            for (long long g = g_begin; g < g_end;) {
                const int b = (int)(g / H), o0 = (int)(g - (long long)b * H);
                const int o1 = (int)min((long long)H, o0 + (g_end - g));
                const int groups = (o1 - o0 + 2 * P + KS - 1) / KS;
                for (int gi = 0; gi < groups; ++gi, ++seq) {
                    const uint32_t slot = seq % sp.nst, round = seq / sp.nst;
                    if (round > 0) dws_wait<DN_SLEEP_PRODUCER>(dws_u32(&empty[slot]), (round - 1) & 1u);
                    const uint32_t bar = dws_u32(&full[slot]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)sp.stage_bytes)
                                 : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
                        "[%2];" ::"r"(dws_u32(stages + (size_t)slot * sp.stage_stride)),
                        "l"(&tmap_x), "r"(bar), "r"(cblk * sp.CB), "r"(col0 - P), "r"(o0 - P + gi * KS), "r"(b)
                        : "memory");
                }
                g += o1 - o0;
            }
        }
        return;
    }

    // ---- consumers ----
    const int ct = threadIdx.x - 32;
    const int nqb = sp.CB / CH;                      // channel groups (CH channels) per block
    const int q = ct % nqb, cb = ct / nqb;
    const bool active = cb < sp.ncb;
    const int ow0 = col0 + cb * TW;
    const int cw = C >> 1;                           // 32-bit words per pixel of the output
    const int row_bytes = sp.IW * sp.CB * 2;
    const int n_cons = n_cons_warps * 32;

    const uint32_t wq = dws_u32(wsm + q * (KS * KS) * CH);
    float2 bv[NP];
#pragma unroll
    for (int h = 0; h < NP; ++h) bv[h] = __ldg(reinterpret_cast<const float2*>(bias + cblk * sp.CB + (active ? q : 0) * CH) + h);
    const uint32_t stage0 = dws_u32(stages) + (uint32_t)((cb * TW * sp.CB + q * CH) * 2);
    const uint32_t cstep = (uint32_t)sp.CB * 2;     // bytes between staged pixels
    float2 acc[KS][TW][NP];                         // ring of k output rows (slot = row mod k)
#pragma unroll
    for (int s2 = 0; s2 < KS; ++s2)
#pragma unroll
        for (int p = 0; p < TW; ++p)
#pragma unroll
            for (int h = 0; h < NP; ++h) acc[s2][p][h] = bv[h];
    float2 ps[NP];                                  // POOL: this thread's output sums over the current image share
#pragma unroll
    for (int h = 0; h < NP; ++h) ps[h] = make_float2(0.f, 0.f);
    uint32_t seq = 0;
    for (long long g = g_begin; g < g_end;) {
        const int b = (int)(g / H), o0 = (int)(g - (long long)b * H);
        const int o1 = (int)min((long long)H, o0 + (g_end - g));
        const int n_steps = o1 - o0 + 2 * P;
        const int groups = (n_steps + KS - 1) / KS;
        uint32_t* yrow = y + ((long long)b * H * W) * cw + (cblk * nqb + q) * NP + ((long long)o0 * W + ow0) * cw;

        for (int gi = 0; gi < groups; ++gi, ++seq) {
            const uint32_t slot = seq % sp.nst, round = seq / sp.nst;
            dws_wait(dws_u32(&full[slot]), round & 1u);
            if (active) {
                const uint32_t st = stage0 + slot * (uint32_t)sp.stage_stride;
                uint32_t raw[TWIN][NP];
#pragma unroll
                for (int j = 0; j < TWIN; ++j) dws_lds_px<NP>(st + j * cstep, raw[j]);
#pragma unroll
                for (int u = 0; u < KS; ++u) {
                    const int i = gi * KS + u;
                    if (i < n_steps) {
                        const int ih = o0 - P + i;
                        float2 in[TWIN][NP];
#pragma unroll
                        for (int j = 0; j < TWIN; ++j)
#pragma unroll
                            for (int h = 0; h < NP; ++h) in[j][h] = h2_to_float2(raw[j][h]);
                        if (u + 1 < KS) {                              // next staged row, in flight during this row's math
#pragma unroll
                            for (int j = 0; j < TWIN; ++j) dws_lds_px<NP>(st + (u + 1) * (uint32_t)row_bytes + j * cstep, raw[j]);
                        }
                        constexpr int NEWEST = KS - 1;                 // output row that receives its first contribution
                        if ((unsigned)ih < (unsigned)H) {
#pragma unroll
                            for (int m = 0; m < KS; ++m) {             // output row o0 + i - 2P + m takes kernel row 2P - m
                                const int kh = 2 * P - m;
                                const int sl = (u + 1 + m) % KS;
#pragma unroll
                                for (int kw = 0; kw < KS; ++kw) {
                                    float2 wv[NP];
                                    dws_lds_tap<NP>(wq + (kh * KS + kw) * (CH * 4), wv);
#pragma unroll
                                    for (int p = 0; p < TW; ++p) {
                                        const bool first = (m == NEWEST) && (kw == 0);      // starts from the bias
#pragma unroll
                                        for (int h = 0; h < NP; ++h)
                                            acc[sl][p][h] = __ffma2_rn(in[p + kw][h], wv[h], first ? bv[h] : acc[sl][p][h]);
                                    }
                                }
                            }
                        } else {                                       // a row above / below the image contributes nothing
#pragma unroll
                            for (int p = 0; p < TW; ++p)
#pragma unroll
                                for (int h = 0; h < NP; ++h) acc[(u + 1 + NEWEST) % KS][p][h] = bv[h];
                        }
                        if (i >= 2 * P) {                              // output row o0 + i - 2P is complete
                            const int sl = (u + 1) % KS;
#pragma unroll
                            for (int p = 0; p < TW; ++p) {
                                if (ow0 + p < W) {
                                    uint32_t v[NP];
#pragma unroll
                                    for (int h = 0; h < NP; ++h) {
                                        v[h] = float2_to_h2(dws_act<ACT>(acc[sl][p][h].x), dws_act<ACT>(acc[sl][p][h].y));
                                        // the STORED (bf16) values are pooled, as a separate pass over y would: their fp32
                                        // sums are (all but) exact, so the result does not depend on how the rows are shared
                                        if (POOL) ps[h] = __fadd2_rn(ps[h], h2_to_float2(v[h]));
                                    }
                                    if (NP == 2) *reinterpret_cast<uint2*>(yrow + (long long)p * cw) = make_uint2(v[0], v[NP - 1]);
                                    else yrow[(long long)p * cw] = v[0];
                                }
                            }
                            yrow += (long long)W * cw;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dws_u32(&empty[slot])) : "memory");
        }
        if (POOL) {
            // column blocks -> channel sums of this share, in a fixed order
            if (active) {
#pragma unroll
                for (int h = 0; h < NP; ++h) {
                    *reinterpret_cast<float2*>(psm + cb * sp.CB + q * CH + 2 * h) = ps[h];
                    ps[h] = make_float2(0.f, 0.f);
                }
            }
            __syncwarp();                           // bar.sync is warp-aligned: rejoin the idle lanes first
            asm volatile("bar.sync 1, %0;" ::"r"(n_cons) : "memory");
            // first CTA share that holds a row of image b: smallest part whose end lies beyond row b * H
            const long long first = (((long long)b * H + 1) * parts + T - 1) / T - 1;
            float* dst = pool + ((long long)b * pool_stride + (part - first)) * C + cblk * sp.CB;
            for (int c = ct; c < sp.CB; c += n_cons) {
                float s = 0.f;
                for (int j = 0; j < sp.ncb; ++j) s += psm[j * sp.CB + c];
                dst[c] = s;
            }
            __syncwarp();
            asm volatile("bar.sync 1, %0;" ::"r"(n_cons) : "memory");
        }
        g += o1 - o0;
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
template <int KS>
static constexpr int dws_tw() { return 4; }
template <int KS>
static constexpr int dws_ch() { return KS == 3 ? 4 : 2; }      // channels per consumer thread

// One candidate plan: the width cut into nstrip column strips.  small_only: the consumers must fit the 160-thread variant
// (three CTAs per SM); the channel block with the most consumer threads wins, the wider one on ties.
static bool dws_plan_strips(int W, int C, int k, int nstrip, bool small_only, DwStream* sp) {
    const int TW = (k == 3) ? dws_tw<3>() : dws_tw<5>();
    const int CH = (k == 3) ? dws_ch<3>() : dws_ch<5>();
    const int P = k / 2;
    const int ncb = ((W + nstrip - 1) / nstrip + TW - 1) / TW;     // column blocks per strip
    const int IW = ncb * TW + 2 * P;
    if (IW > 256 || (nstrip - 1) * ncb * TW >= W) return false;
    // channel block: consumers (CB/4 x ncb threads) should fill about four warps -- small CTAs, several per SM
    int best = 0, best_thr = 0;
    for (int cb = 8; cb <= C; cb += 8) {
        if (C % cb) continue;
        const int thr = (cb / CH) * ncb;
        if (thr > (small_only ? 128 : DWS_MAX_THREADS - 32)) break;
        const bool good = thr >= 96 && thr <= 128, best_good = best_thr >= 96 && best_thr <= 128;
        if (!best || (good && !best_good) || (good == best_good)) best = cb, best_thr = thr;     // later (larger) wins ties
    }
    if (!best) return false;
    sp->CB = best;
    sp->ncb = ncb;
    sp->IW = IW;
    sp->nstrip = nstrip;
    sp->ncblk = C / best;
    sp->stage_bytes = k * IW * best * 2;
    sp->stage_stride = (sp->stage_bytes + 127) & ~127;
    int nst = (60 * 1024) / sp->stage_stride;             // ~60 KB of ring per CTA, three CTAs per SM
    nst = nst < 3 ? 3 : (nst > DWS_MAX_STAGES ? DWS_MAX_STAGES : nst);
    sp->nst = nst;
    sp->threads = 32 + (((best / CH) * ncb + 31) / 32) * 32;
    sp->smem = (size_t)nst * sp->stage_stride + (size_t)k * k * best * 4;
    return sp->smem <= 200 * 1024;
}

// Whole rows per CTA wherever that gives small CTAs (<= 4 consumer warps, three CTAs per SM).  Wide maps (>= 64 columns)
// where it does not -- the row exceeds the 256-pixel TMA box (256 x 256 x 32 of V2 @ 512 ran on the tiles at 4.1 TB/s) or
// even the narrowest channel block needs the 288-thread variant (one CTA per SM) -- are cut into 2 / 4 / 8 column strips:
// the split with the most consumer threads per CTA wins (256 x 256 x 32: four strips of 64 columns x 32 channels, 5.3 TB/s).
// A strip re-reads k - 1 halo columns (3-6 %, L2 hits).  What counts for the stride-1 stream is threads per SM, not the
// width of the channel block: 128 x 128 x 144 in whole rows of 16 channels (128 consumers) runs at 4.4 TB/s, as four strips
// of 48 channels (96 consumers) at 4.0.
bool dw_stream_plan(int H, int W, int C, int k, int stride, DwStream* sp) {
    if (stride != 1 || (k != 3 && k != 5) || C % 8 != 0 || H < 4) return false;
    DwStream whole{};
    const bool whole_ok = dws_plan_strips(W, C, k, 1, false, &whole);
    if (const char* f = getenv("DN_DWS_NSTRIP")) {            // measurement aid: force the strip count on wide maps
        DwStream cand{};
        if (W >= 64 && atoi(f) > 1 && dws_plan_strips(W, C, k, atoi(f), true, &cand)) {
            *sp = cand;
            return true;
        }
    }
    if (W >= 64 && dn_dw_strips() && !(whole_ok && whole.threads <= 160)) {
        DwStream pick{};
        int pick_thr = 0;
        for (int nstrip = 2; nstrip <= 8; nstrip *= 2) {
            DwStream cand{};
            if (!dws_plan_strips(W, C, k, nstrip, true, &cand)) continue;
            const int thr = (cand.CB / ((k == 3) ? dws_ch<3>() : dws_ch<5>())) * cand.ncb;
            // more consumer threads first; at equal threads the wider channel block (256 x 256 x 32: two strips of 16 channels
            // 0.955 ms, four strips of 32 channels 0.816 ms, both 128 consumers)
            if (thr > pick_thr || (thr == pick_thr && cand.CB > pick.CB)) pick = cand, pick_thr = thr;
        }
        if (pick_thr >= 64) {
            *sp = pick;
            return true;
        }
    }
    if (whole_ok) *sp = whole;
    return whole_ok;
}

int dw_stream_make_tmap(CUtensorMap* map, const void* x, int B, int H, int W, int C, int k, const DwStream& sp) {
    DwTiling tl{};
    tl.CB = sp.CB, tl.IWT = sp.IW, tl.IHT = k;
    return dw_make_tmap(map, x, B, H, W, C, tl);
}

template <int KS, int ACT, int NT, bool POOL>
static int dws_launch_p(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H, int W,
                        int C, DwPool* pool, bool probe, cudaStream_t stream) {
    auto kern = dwconv_stream_kernel<KS, dws_tw<KS>(), dws_ch<KS>(), ACT, NT, POOL>;
    const size_t smem = sp.smem + (POOL ? (size_t)sp.ncb * sp.CB * sizeof(float) : 0);
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(kern, smem > 48 * 1024 ? smem : 48 * 1024 + 1, 200 * 1024));
    int per_sm = 0;
    DN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, sp.threads, smem));
    if (per_sm < 1) per_sm = 1;
    // a multiple of the (channel block, column strip) count, at most one CTA per output row of a block
    const int ncs = sp.ncblk * sp.nstrip;
    long long parts = (long long)per_sm * sm_count() / ncs;
    if (parts < 1) parts = 1;
    if (parts > (long long)B * H) parts = (long long)B * H;
    if (POOL) {
        // CTA shares that can touch one image: every share holds at least floor(B * H / parts) rows
        const long long rmin = (long long)B * H / parts;
        const long long slots = (H + rmin - 1) / rmin + 1;
        pool->parts = slots <= pool->max_slots ? (int)parts : 0;
        pool->slots = (int)slots;
        if (probe) return DN_OK;
    }
    const long long grid = parts * ncs;
    launch_pdl(kern, (unsigned)grid, sp.threads, smem, stream, tm, w, bias, (uint32_t*)y, sp, B, H, W, C,
               POOL ? pool->partial : (float*)nullptr, POOL ? pool->slots : 0);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

template <int KS, int ACT, int NT>
static int dws_launch_n(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H, int W,
                        int C, DwPool* pool, cudaStream_t stream) {
    if (pool && pool->partial && sp.nstrip == 1) {       // (column strips do not pool: a slot is one CTA share of whole rows)
        // the pooled variant has its own occupancy: see whether its split keeps the slots of an image within bounds
        int rc = dws_launch_p<KS, ACT, NT, true>(tm, sp, w, bias, y, B, H, W, C, pool, true, stream);
        if (rc) return rc;
        if (pool->parts > 0) return dws_launch_p<KS, ACT, NT, true>(tm, sp, w, bias, y, B, H, W, C, pool, false, stream);
    }
    if (pool) pool->parts = 0;
    return dws_launch_p<KS, ACT, NT, false>(tm, sp, w, bias, y, B, H, W, C, nullptr, false, stream);
}

template <int KS, int ACT>
static int dws_launch_a(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H, int W,
                        int C, DwPool* pool, cudaStream_t stream) {
    // small CTAs (<= 4 consumer warps) are compiled for three per SM; wide rows with few channels need up to 8
    if (sp.threads <= 160) return dws_launch_n<KS, ACT, 160>(tm, sp, w, bias, y, B, H, W, C, pool, stream);
    return dws_launch_n<KS, ACT, DWS_MAX_THREADS>(tm, sp, w, bias, y, B, H, W, C, pool, stream);
}

template <int KS>
static int dws_launch_k(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H, int W,
                        int C, int act, DwPool* pool, cudaStream_t stream) {
    switch (act) {
        case DN_ACT_NONE: return dws_launch_a<KS, DN_ACT_NONE>(tm, sp, w, bias, y, B, H, W, C, pool, stream);
        case DN_ACT_RELU: return dws_launch_a<KS, DN_ACT_RELU>(tm, sp, w, bias, y, B, H, W, C, pool, stream);
        case DN_ACT_RELU6: return dws_launch_a<KS, DN_ACT_RELU6>(tm, sp, w, bias, y, B, H, W, C, pool, stream);
        case DN_ACT_HSWISH: return dws_launch_a<KS, DN_ACT_HSWISH>(tm, sp, w, bias, y, B, H, W, C, pool, stream);
    }
    DN_REQUIRE(false, DN_ERR_INVALID, "bad activation %d", act);
}

int dwconv_stream_launch(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H,
                         int W, int C, int k, int act, cudaStream_t stream, DwPool* pool) {
    DN_REQUIRE((long long)sp.ncblk * sp.nstrip * B * H < (1ll << 40), DN_ERR_UNSUPPORTED, "depthwise problem too large");
    if (k == 3) return dws_launch_k<3>(tm, sp, w, bias, y, B, H, W, C, act, pool, stream);
    return dws_launch_k<5>(tm, sp, w, bias, y, B, H, W, C, act, pool, stream);
}

}  // namespace dn
