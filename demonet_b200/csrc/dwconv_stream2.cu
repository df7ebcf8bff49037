// Depthwise k x k STRIDE-2 stencil as a TMA-fed row stream for sm_100a (the stride-1 stream is dwconv_stream.cu).
// Reference: ConvBNActivation(groups=C, stride=2), demonet/models/mobilenetv2.py:32-55 (see dn_dwconv).
//
// Same structure as the stride-1 kernel: CTA = one channel block for its whole life and an equal contiguous share of
// that block's OUTPUT-row stream (all images); warp 0 produces, one thread issues cp.async.bulk.tensor.4d loads of
// G = 2R input rows x the padded width x CB channels into a ring of shared-memory stages (zero padding = the tensor
// map's out-of-bounds fill).  Consumer thread = 2 channels x TW output columns.  Input row t of a share feeds the
// output rows j = t/2 - m with kernel row kh = t%2 + 2m, so R = (k+1)/2 output rows are open at any time; they live in
// a register ring whose slot indices are compile-time constants because the row loop is unrolled over one period
// (G rows).  An output row is stored when its last input row (t = 2j + k - 1, an even step) has been added.
#include <cuda.h>

#include "common.cuh"
#include "dwconv.cuh"

namespace dn {

__device__ __forceinline__ uint32_t dw2_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ACT>
__device__ __forceinline__ float dw2_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));      // x * relu6(x + 3) / 6
    return v;
}

template <int NS = DN_SLEEP_CONSUMER>
__device__ __forceinline__ void dw2_wait(uint32_t bar, uint32_t parity) {
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
// volatile: the taps are re-read at every use instead of being hoisted into k*k*2 registers
__device__ __forceinline__ float2 dw2_lds_tap(uint32_t addr) {
    float2 f;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f.x), "=f"(f.y) : "r"(addr));
    return f;
}
__device__ __forceinline__ uint32_t dw2_lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

constexpr int DW2_MAX_STAGES = 8;
constexpr int DW2_THREADS = 160;            // producer warp + four consumer warps, three CTAs per SM

// POOL: per-share channel sums of the stored outputs for the squeeze-excitation behind this layer, exactly as in the
// stride-1 stream (dwconv_stream.cu): pool[b][slot][c], slot = index of this CTA among the shares of image b.
template <int KS, int TW, int ACT, bool POOL>
__global__ void __launch_bounds__(DW2_THREADS, 3)
dwconv_stream2_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ w, const float* __restrict__ bias,
                      uint32_t* __restrict__ y, DwStream sp, int B, int H, int W, int C, int Ho, int Wo,
                      float* __restrict__ pool, int pool_stride) {
    constexpr int P = KS / 2, R = (KS + 1) / 2, G = 2 * R, TWIN = 2 * TW + KS - 2;
    extern __shared__ __align__(128) unsigned char dw2_smem[];
    __shared__ uint64_t full[DW2_MAX_STAGES], empty[DW2_MAX_STAGES];
    unsigned char* stages = dw2_smem;
    float* wsm = reinterpret_cast<float*>(dw2_smem + (size_t)sp.nst * sp.stage_stride);     // [CB/2][KS*KS][2]
    float* psm = wsm + KS * KS * sp.CB;                                                      // POOL: [ncb][CB]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_cons_warps = (blockDim.x >> 5) - 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < sp.nst; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw2_u32(&full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dw2_u32(&empty[s])), "r"(n_cons_warps) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // taps of this CTA's channel block, staged once by all threads BEFORE the programmatic-dependency wait, eight
    // independent loads per thread and round (see dwconv_stream.cu): [channel pair][tap][2] so that a thread's taps are
    // base + immediate
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
                    wsm[((c >> 1) * (KS * KS) + t) * 2 + (c & 1)] = v[u];
                }
            }
        }
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();

    // this CTA's channel block and its share of that block's output-row stream (B * Ho rows)
    // (wide maps: sp.nstrip column strips, as in the stride-1 stream)
    const int ncs = sp.ncblk * sp.nstrip;
    const int cblk = blockIdx.x % sp.ncblk;
    const int col0 = ((blockIdx.x / sp.ncblk) % sp.nstrip) * sp.ncb * TW;       // first output column of the strip
    const long long T = (long long)B * Ho;
    const int part = blockIdx.x / ncs, parts = gridDim.x / ncs;
    const long long g_begin = T * part / parts, g_end = T * (part + 1) / parts;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t seq = 0;
            for (long long g = g_begin; g < g_end;) {
                const int b = (int)(g / Ho), o0 = (int)(g - (long long)b * Ho);
                const int o1 = (int)min((long long)Ho, o0 + (g_end - g));
                const int n_steps = 2 * (o1 - o0) + KS - 2;
                const int groups = (n_steps + G - 1) / G;
                for (int gi = 0; gi < groups; ++gi, ++seq) {
                    const uint32_t slot = seq % sp.nst, round = seq / sp.nst;
                    if (round > 0) dw2_wait<DN_SLEEP_PRODUCER>(dw2_u32(&empty[slot]), (round - 1) & 1u);
                    const uint32_t bar = dw2_u32(&full[slot]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)sp.stage_bytes)
                                 : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
                        "[%2];" ::"r"(dw2_u32(stages + (size_t)slot * sp.stage_stride)),
                        "l"(&tmap_x), "r"(bar), "r"(cblk * sp.CB), "r"(2 * col0 - P), "r"(2 * o0 - P + gi * G), "r"(b)
                        : "memory");
                }
                g += o1 - o0;
            }
        }
        return;
    }

    // ---- consumers ----
    const int ct = threadIdx.x - 32;
    const int nqb = sp.CB >> 1;                      // channel pairs per block
    const int q = ct % nqb, cb = ct / nqb;
    const bool active = cb < sp.ncb;
    const int ow0 = col0 + cb * TW;
    const int cw = C >> 1;                           // 32-bit words per output pixel
    const int row_bytes = sp.IW * sp.CB * 2;
    const int n_cons = n_cons_warps * 32;

    const uint32_t wq = dw2_u32(wsm + q * (KS * KS) * 2);
    const float2 bv = __ldg(reinterpret_cast<const float2*>(bias + cblk * sp.CB + (active ? q : 0) * 2));
    const uint32_t stage0 = dw2_u32(stages) + (uint32_t)((2 * cb * TW * sp.CB + q * 2) * 2);
    const uint32_t cstep = (uint32_t)sp.CB * 2;     // bytes between staged pixels
    float2 acc[R][TW];                              // ring of the open output rows (slot = row mod R)
#pragma unroll
    for (int s2 = 0; s2 < R; ++s2)
#pragma unroll
        for (int p = 0; p < TW; ++p) acc[s2][p] = bv;
    float2 ps = make_float2(0.f, 0.f);              // POOL: this thread's output sums over the current image share
    uint32_t seq = 0;
    for (long long g = g_begin; g < g_end;) {
        const int b = (int)(g / Ho), o0 = (int)(g - (long long)b * Ho);
        const int o1 = (int)min((long long)Ho, o0 + (g_end - g));
        const int nrows = o1 - o0;
        const int n_steps = 2 * nrows + KS - 2;
        const int groups = (n_steps + G - 1) / G;
        uint32_t* ybase = y + (((long long)b * Ho + o0) * Wo + ow0) * cw + cblk * nqb + q;

        for (int gi = 0; gi < groups; ++gi, ++seq) {
            const uint32_t slot = seq % sp.nst, round = seq / sp.nst;
            dw2_wait(dw2_u32(&full[slot]), round & 1u);
            if (active) {
                const uint32_t st = stage0 + slot * (uint32_t)sp.stage_stride;
                uint32_t raw[TWIN];
#pragma unroll
                for (int j = 0; j < TWIN; ++j) raw[j] = dw2_lds32(st + j * cstep);
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const int t = gi * G + u;
                    if (t < n_steps) {
                        const int par = u & 1, jn = u >> 1;           // compile-time after unrolling
                        const int ih = 2 * o0 - P + t;
                        float2 in[TWIN];
#pragma unroll
                        for (int j = 0; j < TWIN; ++j) in[j] = h2_to_float2(raw[j]);
                        if (u + 1 < G) {                               // next staged row, in flight during this row's math
#pragma unroll
                            for (int j = 0; j < TWIN; ++j) raw[j] = dw2_lds32(st + (u + 1) * (uint32_t)row_bytes + j * cstep);
                        }
                        if ((unsigned)ih < (unsigned)H) {
#pragma unroll
                            for (int m = 0; m < R; ++m) {              // output row t/2 - m takes kernel row par + 2m
                                const int kh = par + 2 * m;
                                if (kh < KS) {
                                    const int sl = (jn - m + R) % R;
#pragma unroll
                                    for (int kw = 0; kw < KS; ++kw) {
                                        const float2 wv = dw2_lds_tap(wq + (kh * KS + kw) * 8);
#pragma unroll
                                        for (int p = 0; p < TW; ++p) {
                                            const bool first = (kh == 0) && (kw == 0);          // starts from the bias
                                            acc[sl][p] = __ffma2_rn(in[2 * p + kw], wv, first ? bv : acc[sl][p]);
                                        }
                                    }
                                }
                            }
                        } else if (par == 0) {                         // a padding row opens the newest output row
#pragma unroll
                            for (int p = 0; p < TW; ++p) acc[jn % R][p] = bv;
                        }
                        if (par == 0) {                                // output row t/2 - (k-1)/2 is complete
                            const int jc = gi * R + jn - (KS - 1) / 2;
                            if (jc >= 0 && jc < nrows) {
                                const int sl = (jn - (KS - 1) / 2 + 2 * R) % R;
                                uint32_t* yrow = ybase + (long long)jc * Wo * cw;
#pragma unroll
                                for (int p = 0; p < TW; ++p)
                                    if (ow0 + p < Wo) {
                                        const uint32_t v = float2_to_h2(dw2_act<ACT>(acc[sl][p].x), dw2_act<ACT>(acc[sl][p].y));
                                        yrow[(long long)p * cw] = v;
                                        if (POOL) ps = __fadd2_rn(ps, h2_to_float2(v));      // the stored values
                                    }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dw2_u32(&empty[slot])) : "memory");
        }
        if (POOL) {
            if (active) {
                *reinterpret_cast<float2*>(psm + cb * sp.CB + q * 2) = ps;
                ps = make_float2(0.f, 0.f);
            }
            __syncwarp();                           // bar.sync is warp-aligned: rejoin the idle lanes first
            asm volatile("bar.sync 1, %0;" ::"r"(n_cons) : "memory");
            const long long first = (((long long)b * Ho + 1) * parts + T - 1) / T - 1;       // first share with a row of image b
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
static bool dw2_plan_strips(int Wo, int C, int k, int nstrip, DwStream* sp, int* tw_out) {
    const int R = (k + 1) / 2, G = 2 * R;
    const int TW = Wo >= 16 ? 4 : 2;
    const int ncb = ((Wo + nstrip - 1) / nstrip + TW - 1) / TW;       // column blocks per strip
    const int IW = 2 * ncb * TW + k - 2;
    if (IW > 256 || (nstrip - 1) * ncb * TW >= Wo) return false;
    // channel block: consumers (CB/2 x ncb threads) should fill the four consumer warps
    int best = 0, best_thr = 0;
    for (int cb = 8; cb <= C; cb += 8) {
        if (C % cb) continue;
        const int thr = (cb / 2) * ncb;
        if (thr > DW2_THREADS - 32) break;
        if (thr >= best_thr) best = cb, best_thr = thr;
    }
    if (!best || best_thr < 64) return false;        // too few consumers: the tiled / direct kernels do better
    sp->CB = best;
    sp->ncb = ncb;
    sp->IW = IW;
    sp->nstrip = nstrip;
    sp->ncblk = C / best;
    sp->stage_bytes = G * IW * best * 2;
    sp->stage_stride = (sp->stage_bytes + 127) & ~127;
    int nst = (64 * 1024) / sp->stage_stride;             // ~64 KB of ring per CTA, three CTAs per SM
    nst = nst < 2 ? 2 : (nst > DW2_MAX_STAGES ? DW2_MAX_STAGES : nst);
    sp->nst = nst;
    sp->threads = DW2_THREADS;
    sp->smem = (size_t)nst * sp->stage_stride + (size_t)k * k * best * 4;
    if (tw_out) *tw_out = TW;
    return sp->smem <= 72 * 1024;
}

// Whole output rows per CTA, or -- on wide maps (>= 64 output columns) where that leaves a channel block narrower than 32
// channels or does not fit the 256-pixel TMA box -- the first split into 2 / 4 / 8 column strips with >= 32-channel blocks
// (see dw_stream_plan).
bool dw_stream2_plan(int H, int W, int C, int k, DwStream* sp, int* tw_out) {
    if ((k != 3 && k != 5) || C % 8 != 0) return false;
    const int P = k / 2;
    const int Ho = (H + 2 * P - k) / 2 + 1, Wo = (W + 2 * P - k) / 2 + 1;
    if (Ho < 4) return false;
    DwStream whole{};
    int tw = 0;
    const bool whole_ok = dw2_plan_strips(Wo, C, k, 1, &whole, &tw);
    const int want = C < 32 ? C : 32;
    if (Wo >= 64 && dn_dw_strips() && !(whole_ok && whole.CB >= want)) {
        for (int nstrip = 2; nstrip <= 8; nstrip *= 2) {
            DwStream cand{};
            int twc = 0;
            if (dw2_plan_strips(Wo, C, k, nstrip, &cand, &twc) && cand.CB >= want) {
                *sp = cand;
                if (tw_out) *tw_out = twc;
                return true;
            }
        }
    }
    if (whole_ok) {
        *sp = whole;
        if (tw_out) *tw_out = tw;
    }
    return whole_ok;
}

int dw_stream2_make_tmap(CUtensorMap* map, const void* x, int B, int H, int W, int C, int k, const DwStream& sp) {
    DwTiling tl{};
    tl.CB = sp.CB, tl.IWT = sp.IW, tl.IHT = 2 * ((k + 1) / 2);
    return dw_make_tmap(map, x, B, H, W, C, tl);
}

template <int KS, int TW, int ACT, bool POOL>
static int dw2_launch_p(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H, int W,
                        int C, int Ho, int Wo, DwPool* pool, bool probe, cudaStream_t stream) {
    auto kern = dwconv_stream2_kernel<KS, TW, ACT, POOL>;
    const size_t smem = sp.smem + (POOL ? (size_t)sp.ncb * sp.CB * sizeof(float) : 0);
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(kern, smem > 48 * 1024 ? smem : 48 * 1024 + 1, 76 * 1024));
    int per_sm = 0;
    DN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, sp.threads, smem));
    if (per_sm < 1) per_sm = 1;
    const int ncs = sp.ncblk * sp.nstrip;
    long long parts = (long long)per_sm * sm_count() / ncs;
    if (parts < 1) parts = 1;
    if (parts > (long long)B * Ho) parts = (long long)B * Ho;
    if (POOL) {
        // CTA shares that can touch one image: every share holds at least floor(B * Ho / parts) rows
        const long long rmin = (long long)B * Ho / parts;
        const long long slots = (Ho + rmin - 1) / rmin + 1;
        pool->parts = slots <= pool->max_slots ? (int)parts : 0;
        pool->slots = (int)slots;
        if (probe) return DN_OK;
    }
    const long long grid = parts * ncs;
    launch_pdl(kern, (unsigned)grid, sp.threads, smem, stream, tm, w, bias, (uint32_t*)y, sp, B, H, W, C, Ho, Wo,
               POOL ? pool->partial : (float*)nullptr, POOL ? pool->slots : 0);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

template <int KS, int TW, int ACT>
static int dw2_launch_t(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H, int W,
                        int C, int Ho, int Wo, DwPool* pool, cudaStream_t stream) {
    if (pool && pool->partial && sp.nstrip == 1) {       // (column strips do not pool)
        int rc = dw2_launch_p<KS, TW, ACT, true>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, pool, true, stream);
        if (rc) return rc;
        if (pool->parts > 0) return dw2_launch_p<KS, TW, ACT, true>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, pool, false, stream);
    }
    if (pool) pool->parts = 0;
    return dw2_launch_p<KS, TW, ACT, false>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, nullptr, false, stream);
}

template <int KS, int TW>
static int dw2_launch_a(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H, int W,
                        int C, int Ho, int Wo, int act, DwPool* pool, cudaStream_t stream) {
    switch (act) {
        case DN_ACT_NONE: return dw2_launch_t<KS, TW, DN_ACT_NONE>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, pool, stream);
        case DN_ACT_RELU: return dw2_launch_t<KS, TW, DN_ACT_RELU>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, pool, stream);
        case DN_ACT_RELU6: return dw2_launch_t<KS, TW, DN_ACT_RELU6>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, pool, stream);
        case DN_ACT_HSWISH: return dw2_launch_t<KS, TW, DN_ACT_HSWISH>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, pool, stream);
    }
    DN_REQUIRE(false, DN_ERR_INVALID, "bad activation %d", act);
}

int dwconv_stream2_launch(const CUtensorMap& tm, const DwStream& sp, int tw, const float* w, const float* bias, void* y, int B,
                          int H, int W, int C, int k, int act, cudaStream_t stream, DwPool* pool) {
    const int P = k / 2;
    const int Ho = (H + 2 * P - k) / 2 + 1, Wo = (W + 2 * P - k) / 2 + 1;
    DN_REQUIRE((long long)sp.ncblk * sp.nstrip * B * Ho < (1ll << 40), DN_ERR_UNSUPPORTED, "depthwise problem too large");
    if (k == 3) return tw == 4 ? dw2_launch_a<3, 4>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, act, pool, stream)
                               : dw2_launch_a<3, 2>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, act, pool, stream);
    return tw == 4 ? dw2_launch_a<5, 4>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, act, pool, stream)
                   : dw2_launch_a<5, 2>(tm, sp, w, bias, y, B, H, W, C, Ho, Wo, act, pool, stream);
}

}  // namespace dn
