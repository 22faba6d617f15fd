// tcgen05 (5th-gen tensor core) implicit-GEMM convolution interface -- see conv_tc.cu.
#pragma once
#include "common.cuh"

namespace ddn {

bool tc_available();
// 3x3 (pad == dil) or 1x1 (pad 0), stride 1, Cin and Cout multiples of 64
bool tc_conv_supported(int Cin, int Cout, int k, int stride, int pad, int dil, int H, int W);
size_t tc_workspace_bytes(size_t max_act_elems);
int tc_conv_forward(const float* x_nhwc, const float* w_oihw, float* y_nhwc, int N, int H, int W, int Cin, int Cout,
                    int k, int pad, int dil, int precision, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_conv_backward(const float* x_nhwc, const float* w_oihw, const float* dy_nhwc, float* dx_nhwc, const float* dx_addend,
                     float* dw_oihw, int N, int H, int W, int Cin, int Cout, int k, int pad, int dil, int precision,
                     void* ws, size_t ws_bytes, float* dwp_scratch, cudaStream_t st);

}  // namespace ddn
