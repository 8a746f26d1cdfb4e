// Winograd F(2x2, 3x3) transform passes as CUDA kernels (bodies: wino_passes.cuh, shared with the host emulation of
// the CPU test suite).  Reference: model/TSNet.py:10-49 (ResnetBlock: ReflectionPad2d(1) + 3x3 conv + InstanceNorm).
#include "wino_passes.cuh"
#include "host_util.h"
#include "../../include/tsnet_b200.h"

namespace tsnet {

__global__ void __launch_bounds__(256) wino_weight_kernel(const float* __restrict__ w, int Cout, int Cin,
                                                          float* __restrict__ u) {
  wino_weight_body(w, Cout, Cin, u, blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x);
}

__global__ void __launch_bounds__(256, 2) wino_input_kernel(const WinoInArgs a) {
  wino_input_body(a, blockIdx.x, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256, 3) wino_output_kernel(const WinoOutArgs a) {
  wino_output_body(a, blockIdx.x, threadIdx.x, blockDim.x);
}

// called by tsnet_build_taps (elementwise.cu) for mode TSNET_TAPS_WINO
int launch_wino_input(const WinoInArgs& a, cudaStream_t stream) {
  TSNET_ARG_CHECK(a.H % 2 == 0 && a.W % 2 == 0 && a.H >= 4 && a.W >= 4, "build_taps(WINO): H, W must be even, >= 4");
  TSNET_ARG_CHECK(a.C % 4 == 0, "build_taps(WINO): C %d must be a multiple of 4", a.C);
  TSNET_ARG_CHECK(a.hi && a.lo, "build_taps(WINO): needs the hi / lo destination");
  const unsigned rows = static_cast<unsigned>(a.B) * (a.H / 2);
  wino_input_kernel<<<rows, 256, 0, stream>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}

}  // namespace tsnet

using namespace tsnet;

extern "C" int tsnet_wino_weight_transform(const float* w_oihw, int Cout, int Cin, float* u_out, void* stream) {
  TSNET_ARG_CHECK(w_oihw && u_out && Cout > 0 && Cin > 0, "wino_weight_transform: bad argument");
  const size_t total = static_cast<size_t>(Cout) * Cin;
  wino_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, Cout, Cin, u_out);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_wino_output(const float* m, int B, int H, int W, int C, const float* bias, const float* addend,
                                 long long addend_rows, float* y_raw, float* stats_partial, void* stream) {
  TSNET_ARG_CHECK(m && y_raw, "wino_output: null argument");
  TSNET_ARG_CHECK(H % 2 == 0 && W % 2 == 0 && (W / 2) % kWinoRun == 0, "wino_output: W/2 = %d must be a multiple of %d",
                  W / 2, kWinoRun);
  TSNET_ARG_CHECK(C % 4 == 0, "wino_output: C %d must be a multiple of 4", C);
  TSNET_ARG_CHECK(!addend || addend_rows > 0, "wino_output: addend needs addend_rows > 0");
  WinoOutArgs a;
  a.m = m; a.bias = bias; a.addend = addend; a.y = y_raw; a.stats = stats_partial;
  a.B = B; a.H = H; a.W = W; a.C = C; a.addend_rows = addend_rows;
  const unsigned rows = static_cast<unsigned>(B) * (H / 2);
  wino_output_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}
