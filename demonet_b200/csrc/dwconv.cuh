// Tiling of the TMA-fed depthwise kernel (dwconv_tma.cu), shared with the engine.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace dn {

constexpr int DW_THREADS = 256;

struct DwTiling {
    int CB;                 // channels per CTA chunk (multiple of 8, divides C, <= 64)
    int THo, TWo;           // output tile
    int spr;                // strips (TW output columns) per tile row
    int IHT, IWT;           // input halo tile
    int tiles_x, tiles_y, chunks;
};

int dw_plan(int H, int W, int C, int k, int stride, DwTiling* tl);
int dw_make_tmap(CUtensorMap* map, const void* x, int B, int H, int W, int C, const DwTiling& tl);
int dwconv_tma_launch(const CUtensorMap& tm, const DwTiling& tl, const float* w, const float* bias, void* y, int B, int H,
                      int W, int C, int k, int stride, int act, cudaStream_t stream);

// TMA-fed row stream (dwconv_stream.cu), stride 1
constexpr int DWS_MAX_THREADS = 288;        // producer warp + up to 8 consumer warps
struct DwStream {
    int CB;                 // channels per block (multiple of 8, divides C)
    int ncb;                // column blocks (TW output columns each) of one column strip
    int IW;                 // staged row width in pixels (ncb * TW + k - 1)
    int nstrip;             // column strips across the width (1 = a CTA stages whole rows); a strip is ncb * TW output
                            // columns wide, so that wide maps still get >= 64-byte channel blocks per staged pixel
    int ncblk;              // channel blocks
    int nst;                // ring stages (k input rows each)
    int stage_bytes, stage_stride;
    int threads;
    size_t smem;
};
// measurement aid: DN_DW_STRIPS=0 restores whole-row CTAs on wide maps (read per call: the tests flip it inside one process)
inline bool dn_dw_strips() {
    const char* e = getenv("DN_DW_STRIPS");
    return !(e && e[0] == '0');
}
bool dw_stream_plan(int H, int W, int C, int k, int stride, DwStream* sp);
int dw_stream_make_tmap(CUtensorMap* map, const void* x, int B, int H, int W, int C, int k, const DwStream& sp);
// Pooling side output of a row-stream launch for the squeeze-excitation that follows (dwconv_stream.cu, se.cu).
// In: partial (fp32 [B][max_slots][C] workspace, NULL = no pooling), max_slots.  Out: parts = CTA shares per channel
// block of the launch (0 = the launch did not pool: too many shares per image), slots = slot stride of `partial`.
struct DwPool {
    float* partial;
    int max_slots;
    int parts, slots;
};
int dwconv_stream_launch(const CUtensorMap& tm, const DwStream& sp, const float* w, const float* bias, void* y, int B, int H,
                         int W, int C, int k, int act, cudaStream_t stream, DwPool* pool = nullptr);
// TMA-fed row stream, stride 2 (dwconv_stream2.cu): same DwStream plan record (stage = 2 * ((k + 1) / 2) input rows),
// tw = output columns per consumer thread (4 or 2)
bool dw_stream2_plan(int H, int W, int C, int k, DwStream* sp, int* tw_out);
int dw_stream2_make_tmap(CUtensorMap* map, const void* x, int B, int H, int W, int C, int k, const DwStream& sp);
int dwconv_stream2_launch(const CUtensorMap& tm, const DwStream& sp, int tw, const float* w, const float* bias, void* y, int B,
                          int H, int W, int C, int k, int act, cudaStream_t stream, DwPool* pool = nullptr);
// which depthwise kernel a layer shape runs on
enum DwImpl { DW_DIRECT = 1, DW_TMA = 2, DW_STREAM = 4, DW_STREAM2 = 8 };
DwImpl dw_choose(int H, int W, int C, int k, int stride);

// TMA-tiled stem (stem_tma.cu)
bool stem_can_tma(const void* images, int W);
bool stem_norm_ok(const float* std3);
int stem_make_tmap(CUtensorMap* map, const float* images, int B, int H, int W);
// tensor-core stem (stem_tc.cu): im2col rows built by the threads, hi / lo fp16 split, tcgen05 GEMM; same preconditions
bool stem_tc_enabled();
int stem_tc_launch(const float* images, const float* w, const float* bias, const float* mean3, const float* std3, void* y, int B,
                   int H, int W, int Cout, int act, cudaStream_t stream);
int stem_tma_launch(const CUtensorMap& tm, const float* w, const float* bias, const float* mean3, const float* std3, void* y,
                    int B, int H, int W, int Cout, int act, cudaStream_t stream);

}  // namespace dn
