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

template <int CS, int PS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) wino_bridge_kernel(const WinoBridgeArgs a) {
  extern __shared__ __align__(16) uint8_t bridge_smem[];
  float* s_y = reinterpret_cast<float*>(bridge_smem);
  double* s_part = reinterpret_cast<double*>(bridge_smem + static_cast<size_t>(a.H) * a.W * PS * 4);
  float* s_mr = reinterpret_cast<float*>(s_part + THREADS * 8);
  wino_bridge_phase_a<CS, PS>(a, blockIdx.x, threadIdx.x, THREADS, s_y, s_part);
  __syncthreads();
  wino_bridge_phase_s2<CS, PS>(a, blockIdx.x, threadIdx.x, THREADS, s_part, s_mr);
  __syncthreads();
  // a boundary without residual / act_out / correlation outputs needs no in-place pass: phase C normalises on read
  const bool fold = !a.residual && !a.act_out && !a.corr_hi;
  if (!fold) {
    wino_bridge_phase_b<CS, PS>(a, blockIdx.x, threadIdx.x, THREADS, s_y, s_mr);
    __syncthreads();
    wino_bridge_phase_n<CS, PS>(a, blockIdx.x, threadIdx.x, THREADS, s_y);
  }
  wino_bridge_phase_c<CS, PS>(a, blockIdx.x, threadIdx.x, THREADS, s_y, s_mr, fold);
}

template <int CS, int PS, int THREADS, int MINB>
static int launch_wino_bridge(const WinoBridgeArgs& a, cudaStream_t stream) {
  const size_t smem = wino_bridge_smem_bytes<CS, PS>(a.H, a.W, THREADS);
  TSNET_ARG_CHECK(smem <= 227 * 1024, "wino_bridge: %d x %d image needs %zu B of shared memory", a.H, a.W, smem);
  static int smem_attr[kMaxDevices] = {0};
  TSNET_CUDA_CHECK(ensure_dyn_smem(wino_bridge_kernel<CS, PS, THREADS, MINB>, static_cast<int>(smem), smem_attr));
  const unsigned blocks = static_cast<unsigned>(a.B) * (a.C / CS);
  wino_bridge_kernel<CS, PS, THREADS, MINB><<<blocks, THREADS, smem, stream>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}

// called by tsnet_build_taps (elementwise.cu) for mode TSNET_TAPS_WINO
int launch_wino_input(const WinoInArgs& a, cudaStream_t stream) {
  TSNET_ARG_CHECK(a.H % 2 == 0 && a.W % 2 == 0 && a.H >= 4 && a.W >= 4, "build_taps(WINO): H, W must be even, >= 4");
  TSNET_ARG_CHECK(a.C % 4 == 0, "build_taps(WINO): C %d must be a multiple of 4", a.C);
  TSNET_ARG_CHECK(a.Cp_total % 64 == 0, "build_taps(WINO): Cp_total %d must be a multiple of 64", a.Cp_total);
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
  TSNET_ARG_CHECK(C % kWinoMSlab == 0, "wino_output: C %d must be a multiple of %d", C, kWinoMSlab);
  TSNET_ARG_CHECK(!addend || addend_rows > 0, "wino_output: addend needs addend_rows > 0");
  WinoOutArgs a;
  a.m = m; a.bias = bias; a.addend = addend; a.y = y_raw; a.stats = stats_partial;
  a.B = B; a.H = H; a.W = W; a.C = C; a.addend_rows = addend_rows;
  const unsigned rows = static_cast<unsigned>(B) * (H / 2);
  wino_output_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_wino_bridge(const tsnet_wino_bridge_desc* d, const float* m, const float* bias, const float* addend,
                                 const float* residual, float* act_out, float* mean_rstd_out, uint16_t* v_hi,
                                 uint16_t* v_lo, void* stream) {
  TSNET_ARG_CHECK(d && m && v_hi && v_lo, "wino_bridge: null argument");
  TSNET_ARG_CHECK(d->H % 2 == 0 && d->W % 2 == 0 && d->H >= 4 && d->W >= 4, "wino_bridge: H, W must be even, >= 4");
  TSNET_ARG_CHECK(d->C > 0 && d->C % kWinoMSlab == 0, "wino_bridge: C %d must be a multiple of %d", d->C, kWinoMSlab);
  TSNET_ARG_CHECK(d->variant == 0 || d->variant == 1, "wino_bridge: variant %d", d->variant);
  TSNET_ARG_CHECK(d->Cp_total % 64 == 0 && d->c_off % 4 == 0 && d->c_off + d->C <= d->Cp_total,
                  "wino_bridge: operand channel window does not fit (Cp_total must be a multiple of 64)");
  TSNET_ARG_CHECK(!addend || d->addend_rows > 0, "wino_bridge: addend needs addend_rows > 0");
  const int actC = d->act_C_total > 0 ? d->act_C_total : d->C;
  TSNET_ARG_CHECK(!act_out || (actC % 4 == 0 && d->act_c_off % 4 == 0 && d->act_c_off + d->C <= actC),
                  "wino_bridge: act_out channel window does not fit");
  WinoBridgeArgs a;
  a.m = m; a.bias = bias; a.addend = addend; a.residual = residual; a.act_out = act_out;
  a.mean_rstd_out = mean_rstd_out; a.hi = v_hi; a.lo = v_lo;
  a.B = d->B; a.H = d->H; a.W = d->W; a.C = d->C; a.relu = d->relu; a.Cp_total = d->Cp_total; a.c_off = d->c_off;
  a.fmt = d->fmt; a.act_C_total = actC; a.act_c_off = d->act_c_off; a.addend_rows = d->addend_rows;
  a.scale = d->scale == 0.f ? 1.f : d->scale;
  a.eps = d->eps == 0.f ? 1e-5f : d->eps;
  TSNET_ARG_CHECK((d->corr_hi == nullptr) == (d->corr_lo == nullptr) && (d->corr_hi == nullptr) == (d->corr_rank == nullptr) &&
                      (d->corr_hi == nullptr) == (d->corr_ssq == nullptr),
                  "wino_bridge: corr_hi / corr_lo / corr_rank / corr_ssq come together");
  a.corr_hi = d->corr_hi; a.corr_lo = d->corr_lo; a.corr_rank = d->corr_rank; a.corr_ssq = d->corr_ssq;
  a.corr_scale = d->corr_scale == 0.f ? 1.f : d->corr_scale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d->variant == 1 ? launch_wino_bridge<16, 24, 256, 2>(a, st) : launch_wino_bridge<32, 32, 512, 1>(a, st);
}
