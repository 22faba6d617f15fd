// tcgen05 (5th-gen tensor core) implicit-GEMM convolution interface -- see conv_tc.cu.
#pragma once
#include "common.cuh"

namespace ddn {

struct TcPlanes { const __nv_bfloat16* hi; const __nv_bfloat16* lo; };   // x ~= hi + lo (lo unused in single-pass bf16)

bool tc_available();
// 3x3 (pad == dil) or 1x1 (pad 0), stride 1, Cin and Cout multiples of 64
bool tc_conv_supported(int Cin, int Cout, int k, int stride, int pad, int dil, int H, int W);
size_t tc_weight_ws_bytes();                       // staging for one conv's packed bf16 weights
size_t tc_workspace_bytes(size_t max_act_elems);   // staging for the fp32-tensor wrappers below
int tc_bn_partial_blocks(int N, int H, int W);     // rows of the BN partial-sum buffer a forward conv writes
int tc_split(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, int precision, cudaStream_t st);

// plane-level entry points (what the network engine calls)
int tc_conv_planes(TcPlanes in, const float* w_oihw, float* out, const float* addend, float* bn_partial,
                   int N, int H, int W, int Cin, int Cout, int k, int dil, int dgrad, int precision,
                   void* wws, size_t wws_bytes, cudaStream_t st);
int tc_wgrad_planes(TcPlanes x, TcPlanes dy, float* dw, int N, int H, int W, int Cin, int Cout, int k, int dil,
                    int precision, float* dwp, cudaStream_t st);

// fp32-tensor wrappers (single-operator C ABI)
int tc_conv_forward(const float* x_nhwc, const float* w_oihw, float* y_nhwc, int N, int H, int W, int Cin, int Cout,
                    int k, int pad, int dil, int precision, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_conv_backward(const float* x_nhwc, const float* w_oihw, const float* dy_nhwc, float* dx_nhwc, const float* dx_addend,
                     float* dw_oihw, int N, int H, int W, int Cin, int Cout, int k, int pad, int dil, int precision,
                     void* ws, size_t ws_bytes, float* dwp_scratch, cudaStream_t st);

}  // namespace ddn
