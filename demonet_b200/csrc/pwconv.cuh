// Epilogue shared by the pointwise-GEMM kernels: folded-BN bias, activation, residual add and the
// strided output addressing that turns the head's view/permute/reshape/cat into plain stores.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dn {

struct PwEpilogue {
    const float* bias;                 // [N]
    const dn_half_t* residual;     // [M, N] or nullptr
    const float* a_scale = nullptr;    // [M / hw][a_scale_c] per-(image, input channel) scale of the A operand (folded SE)
    int a_scale_c = 0;                 // channels of the scale vector; K is a multiple of it (pixel-packed layers: K = p * C)
    void* y;
    int N;
    int act;
    int out_fp32;
    int hw;                            // rows per image
    long long out_batch_stride;        // elements
    long long out_row_stride;          // elements

    __device__ __forceinline__ long long row_offset(int m) const {
        const int b = m / hw, r = m - b * hw;
        return (long long)b * out_batch_stride + (long long)r * out_row_stride;
    }
    // y = act(acc + bias[n]) (+ residual[m][n]), single rounding at the end
    __device__ __forceinline__ float finish(int m, int n, float acc) const {
        float v = apply_act(acc + __ldg(bias + n), act);
        if (residual) v += half_to_float(residual[(long long)m * N + n]);
        return v;
    }
    __device__ __forceinline__ void store(int m, int n, float acc) const {
        const float v = finish(m, n, acc);
        const long long o = row_offset(m) + n;
        if (out_fp32) reinterpret_cast<float*>(y)[o] = v;
        else reinterpret_cast<dn_half_t*>(y)[o] = float_to_half(v);
    }
};

int pwconv_simt(const void* x, const void* w, const PwEpilogue& ep, int M, int K, int N, cudaStream_t stream);
int pwconv_tc(const void* x, const void* w, const PwEpilogue& ep, int M, int K, int N, cudaStream_t stream);
// pieces of pwconv_tc that the engine caches per layer
int make_tmap_h16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, int box_cols);
// m_plan: the row count the tile shape is planned for (the engine passes its max-batch M so that the weight tensor
// map built at load time and every later launch agree)
// pair: the layer runs as clusters of two CTAs that share every weight k-block by TMA multicast (pwconv_tc.cu, PAIR mode);
// the weight tensor map's box is then block_n / 2 rows
void pwconv_tc_plan(long long m_plan, int K, int N, int* block_n, int* n_tiles, int* stages, int* tmem_cols, size_t* smem_bytes,
                    int* w_stationary = nullptr, int* pair = nullptr, int pair_force = -1);
// pair_planned: the `pair` answer of the plan the weight tensor map was built with (-1: plan again)
int pwconv_tc_launch(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap* ty, const PwEpilogue& ep, int M, long long m_plan, int K,
                     int N, cudaStream_t stream, int pair_planned = -1);

// fused pointwise-expand -> depthwise (pwdw_fused.cu)
bool pwdw_fused_supported(int H, int W, int K, int N, int ksize, int stride);
int pwdw_fused_make_tmaps(CUtensorMap* tx, CUtensorMap* tw, const void* x, const void* w_pw, int B, int H, int W);
int pwdw_fused_launch(const CUtensorMap& tx, const CUtensorMap& tw, const float* b_pw, const float* w_dw, const float* b_dw, void* y,
                      int B, int H, int W, int act_pw, int act_dw, cudaStream_t stream);

// fused depthwise -> pointwise project (+ residual) (dwpw_fused.cu)
bool dwpw_fused_supported(int H, int W, int C, int N, int ksize, int stride);
int dwpw_fused_make_tmaps(CUtensorMap* tx, CUtensorMap* tw, const void* x, const void* w_pw, int B, int H, int W);
int dwpw_fused_launch(const CUtensorMap& tx, const CUtensorMap& tw, const float* w_dw, const float* b_dw, const float* b_pw, void* y,
                      int B, int H, int W, int act_dw, int residual, cudaStream_t stream);

}  // namespace dn
