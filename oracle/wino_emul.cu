// TEST INFRASTRUCTURE (not part of the product path): host emulation of the Winograd transform passes.
// The kernel bodies of wacv23_tsnet_b200/csrc/wino_passes.cuh are plain functions of (block, thread); here they are
// executed by nested loops on the CPU so that `pytest -m "not gpu"` checks their index math, layouts and statistics
// against torch before any GPU time is spent.  Built by oracle/Makefile into oracle/_build/libwino_emul.so.
#include "../wacv23_tsnet_b200/csrc/wino_passes.cuh"

using namespace tsnet;

extern "C" void wino_emul_weight(const float* w, int Cout, int Cin, float* u) {
  const size_t total = static_cast<size_t>(Cout) * Cin;
  for (size_t i = 0; i < total; ++i) wino_weight_body(w, Cout, Cin, u, i);
}

extern "C" void wino_emul_input(const float* raw, const float* mean_rstd, const float* residual, float* act_out,
                                uint16_t* hi, uint16_t* lo, int B, int H, int W, int C, int relu, int Cp_total,
                                int c_off, int fmt, int act_C_total, int act_c_off, float scale, int nthreads) {
  WinoInArgs a;
  a.raw = raw; a.mean_rstd = mean_rstd; a.residual = residual; a.act_out = act_out; a.hi = hi; a.lo = lo;
  a.B = B; a.H = H; a.W = W; a.C = C; a.relu = relu; a.Cp_total = Cp_total; a.c_off = c_off; a.fmt = fmt;
  a.act_C_total = act_C_total; a.act_c_off = act_c_off; a.scale = scale;
  const int blocks = B * (H / 2);
  for (int blk = 0; blk < blocks; ++blk)
    for (int t = 0; t < nthreads; ++t) wino_input_body(a, blk, t, nthreads);
}

extern "C" void wino_emul_output(const float* m, const float* bias, const float* addend, long long addend_rows, float* y,
                                 float* stats, int B, int H, int W, int C, int nthreads) {
  WinoOutArgs a;
  a.m = m; a.bias = bias; a.addend = addend; a.y = y; a.stats = stats; a.B = B; a.H = H; a.W = W; a.C = C;
  a.addend_rows = addend_rows;
  const int blocks = B * (H / 2);
  for (int blk = 0; blk < blocks; ++blk)
    for (int t = 0; t < nthreads; ++t) wino_output_body(a, blk, t, nthreads);
}

template <int CS, int PS>
static void emul_bridge(const WinoBridgeArgs& a, int nthreads) {
  const size_t bytes = wino_bridge_smem_bytes<CS, PS>(a.H, a.W, nthreads);
  uint8_t* smem = new uint8_t[bytes + 16];
  float* s_y = reinterpret_cast<float*>(smem);
  double* s_part = reinterpret_cast<double*>(smem + static_cast<size_t>(a.H) * a.W * PS * 4);
  float* s_mr = reinterpret_cast<float*>(s_part + static_cast<size_t>(nthreads) * 8);
  const int blocks = a.B * (a.C / CS);
  const bool fold = !a.residual && !a.act_out && !a.corr_hi;
  for (int blk = 0; blk < blocks; ++blk) {   // phases separated by block-wide barriers in the kernel
    for (int t = 0; t < nthreads; ++t) wino_bridge_phase_a<CS, PS>(a, blk, t, nthreads, s_y, s_part);
    for (int t = 0; t < nthreads; ++t) wino_bridge_phase_s2<CS, PS>(a, blk, t, nthreads, s_part, s_mr);
    if (!fold) {
      for (int t = 0; t < nthreads; ++t) wino_bridge_phase_b<CS, PS>(a, blk, t, nthreads, s_y, s_mr);
      for (int t = 0; t < nthreads; ++t) wino_bridge_phase_n<CS, PS>(a, blk, t, nthreads, s_y);
    }
    for (int t = 0; t < nthreads; ++t) wino_bridge_phase_c<CS, PS>(a, blk, t, nthreads, s_y, s_mr, fold);
  }
  delete[] smem;
}

extern "C" void wino_emul_bridge(const float* m, const float* bias, const float* addend, long long addend_rows,
                                 const float* residual, float* act_out, float* mean_rstd_out, uint16_t* hi, uint16_t* lo,
                                 int B, int H, int W, int C, int relu, int Cp_total, int c_off, int fmt, int act_C_total,
                                 int act_c_off, float scale, float eps, int nthreads, uint16_t* corr_hi,
                                 uint16_t* corr_lo, const uint16_t* corr_rank, float* corr_ssq, float corr_scale,
                                 int variant) {
  WinoBridgeArgs a;
  a.corr_hi = corr_hi; a.corr_lo = corr_lo; a.corr_rank = corr_rank; a.corr_ssq = corr_ssq; a.corr_scale = corr_scale;
  a.m = m; a.bias = bias; a.addend = addend; a.residual = residual; a.act_out = act_out; a.mean_rstd_out = mean_rstd_out;
  a.hi = hi; a.lo = lo; a.B = B; a.H = H; a.W = W; a.C = C; a.relu = relu; a.Cp_total = Cp_total; a.c_off = c_off;
  a.fmt = fmt; a.act_C_total = act_C_total; a.act_c_off = act_c_off; a.addend_rows = addend_rows; a.scale = scale;
  a.eps = eps;
  if (variant == 1) emul_bridge<16, 24>(a, nthreads);
  else emul_bridge<32, 32>(a, nthreads);
}
