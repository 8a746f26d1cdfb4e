// HBM-bound helper kernels of the TS-Net forward: weight packing, InstanceNorm statistics reduction,
// tap-source construction (normalise + ReLU + residual + pad / upsample / parity split + hi/lo split),
// encoder stem input construction, L2 normalisation for the correlation, output head, and a plain
// fp32 direct convolution used for on-device validation.
#include "sm100_prims.cuh"
#include "host_util.h"
#include "wino_passes.cuh"
#include "../../include/tsnet_b200.h"
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace tsnet {

__device__ __forceinline__ int reflect_idx(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int fold_kw,
                                   int Cp, int Cout_pad, float scale, int fmt, uint16_t* __restrict__ hi,
                                   uint16_t* __restrict__ lo) {
  const int taps = fold_kw ? KH : KH * KW;
  const size_t K = static_cast<size_t>(taps) * Cp;
  const size_t total = static_cast<size_t>(Cout_pad) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i / K);
    const int kcol = static_cast<int>(i - static_cast<size_t>(o) * K);
    const int tap = kcol / Cp;
    const int j = kcol - tap * Cp;
    int r, s, c;
    bool valid;
    if (fold_kw) {
      const int fc = fold_kw > 1 ? fold_kw : Cin;  // channel stride of a folded horizontal tap
      r = tap; s = j / fc; c = j - s * fc; valid = s < KW && c < Cin;
    } else {
      r = tap / KW; s = tap - r * KW; c = j; valid = c < Cin;
    }
    float v = 0.f;
    if (valid && o < Cout) v = w[((static_cast<size_t>(o) * Cin + c) * KH + r) * KW + s] * scale;
    uint16_t h, l;
    split16(v, fmt, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// InstanceNorm statistics: [B*nsub, C, 2] (sum, centred M2 of 32 pixels) -> [B, C, 2] (mean, rstd)
// block = (32 channels) x (8 segments).  All partials cover 32 pixels, so the merge is
//   mean = (sum_k sum_k) / N ;  M2 = sum_k [ M2_k + sum_k^2 / 32 ] - N mean^2
// evaluated in fp64 in ONE pass with a fixed summation order (bit-reproducible), one sqrt per (b, c).
// ------------------------------------------------------------------------------------------------
constexpr int kInSegs = 16;
__global__ void __launch_bounds__(32 * kInSegs) instnorm_reduce_kernel(const float* __restrict__ part, int nsub, int C,
                                                                       float eps, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int seg = threadIdx.y;
  __shared__ double s_sum[kInSegs][32], s_q[kInSegs][32];
  const int per = (nsub + kInSegs - 1) / kInSegs;
  const int k0 = seg * per, k1 = min(nsub, k0 + per);
  const float* p = part + (static_cast<size_t>(b) * nsub) * C * 2 + static_cast<size_t>(min(c, C - 1)) * 2;
  // single pass: S = sum_k sum_k,  Q = sum_k (M2_k + sum_k^2 / 32) = sum of squares;  M2 = Q - S^2 / N.
  // The subtraction is done in fp64 on fp32-precision inputs: exact to ~1e-16 Q, i.e. far below fp32 resolution
  // of the variance even when mean^2 >> var.
  double acc_s = 0.0, acc_q = 0.0;
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const float2 sm = *reinterpret_cast<const float2*>(p + static_cast<size_t>(k) * C * 2);
    const double sk = static_cast<double>(sm.x);
    acc_s += sk;
    acc_q += static_cast<double>(sm.y) + sk * sk * (1.0 / 32.0);
  }
  s_sum[seg][threadIdx.x] = acc_s;
  s_q[seg][threadIdx.x] = acc_q;
  __syncthreads();
  if (seg == 0 && c < C) {
    double total = 0.0, q = 0.0;
#pragma unroll
    for (int s = 0; s < kInSegs; ++s) {
      total += s_sum[s][threadIdx.x];
      q += s_q[s][threadIdx.x];
    }
    const double n = 32.0 * nsub;
    const double mean = total / n;
    const double var = fmax(q / n - mean * mean, 0.0);  // biased, as nn.InstanceNorm2d
    float2 r;
    r.x = static_cast<float>(mean);
    r.y = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    *reinterpret_cast<float2*>(out + (static_cast<size_t>(b) * C + c) * 2) = r;
  }
}

// ------------------------------------------------------------------------------------------------
// tap-source builder.  One thread = one destination pixel x 8 channels.
// ------------------------------------------------------------------------------------------------
struct TapsArgs {
  const float* raw;
  const float* mean_rstd;
  const float* residual;
  float* act_out;
  uint16_t* hi;
  uint16_t* lo;
  int B, H, W, C, mode, relu, Cp_total, c_off, fmt;
  float scale;
  int Hd, Wd, planes;  // destination geometry
  int act_C_total, act_c_off, avg_n;
  // residual given as a channel concatenation of two tensors (torch.cat([a, t], 1) never materialised): channels
  // [0, res_split) from `residual` [B, H, W, res_split], the rest from residual2 [res2_batch, H, W, C - res_split] with
  // the sample index taken modulo res2_batch (one target feature map shared by all sources)
  const float* residual2;
  int res_split, res2_batch;
};

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void fetch_act8(const TapsArgs& a, int b, int y, int x, int c, const float (&mean)[8],
                                           const float (&rstd)[8], float (&v)[8]) {
  const size_t off = ((static_cast<size_t>(b) * a.H + y) * a.W + x) * a.C + c;
  load8(a.raw + off, v);
  if (a.mean_rstd) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean[j]) * rstd[j];
  }
  if (a.relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (a.residual) {
    float r[8];
    if (a.residual2 == nullptr) {
      load8(a.residual + off, r);
    } else {
      const size_t pix = (static_cast<size_t>(b) * a.H + y) * a.W + x;
      if (c < a.res_split) {
        load8(a.residual + pix * a.res_split + c, r);
      } else {
        const size_t pix2 = (static_cast<size_t>(b % a.res2_batch) * a.H + y) * a.W + x;
        load8(a.residual2 + pix2 * (a.C - a.res_split) + (c - a.res_split), r);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += r[j];
  }
}

// grid.x = destination rows (b, plane, yd); the 256 threads of a block sweep (xd, channel group) of that row, so the
// per-thread index math is 32-bit and, whenever 256 % (C/8) == 0, every thread keeps ONE channel group for the whole
// row and loads its InstanceNorm statistics once.
// MODE is a template parameter so that the simple modes are not compiled with the register budget of the 4-tap
// bilinear one (80 registers, 37 % occupancy for every mode when it was a run-time switch).
template <int MODE>
__global__ void __launch_bounds__(256, MODE == TSNET_TAPS_UP2REFLECT1 ? 3 : (MODE == TSNET_TAPS_SAME ? 4 : 5)) build_taps_kernel(const TapsArgs a) {
  const int cg = a.C / 8;
  int rowid = blockIdx.x;
  const int yd = rowid % a.Hd;
  rowid /= a.Hd;
  const int plane = rowid % a.planes;
  const int b = rowid / a.planes;
  const bool fixed_c = (blockDim.x % cg) == 0;
  const int per_row = a.Wd * cg;

  float mean[8], rstd[8];
  int c_loaded = -1;
  for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
    const int xd = idx / cg;
    const int c = (idx - xd * cg) * 8;
    if (a.mean_rstd && c != c_loaded) {
      const float* mr = a.mean_rstd + (static_cast<size_t>(b) * a.C + c) * 2;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float4 t = *reinterpret_cast<const float4*>(mr + j * 2);
        mean[j] = t.x; rstd[j] = t.y; mean[j + 1] = t.z; rstd[j + 1] = t.w;
      }
      if (fixed_c) c_loaded = c;
    }
    float v[8];
    bool interior = false;  // destination pixel that owns the (unique) act_out write of its source pixel
    if constexpr (MODE == TSNET_TAPS_SAME) {
      fetch_act8(a, b, yd, xd, c, mean, rstd, v);
      if (a.avg_n > 1) {  // mean over sources: samples b, B + b, 2B + b, ...
        for (int i = 1; i < a.avg_n; ++i) {
          const int bi = i * a.B + b;
          if (a.mean_rstd) {
            const float* mr = a.mean_rstd + (static_cast<size_t>(bi) * a.C + c) * 2;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const float4 t = *reinterpret_cast<const float4*>(mr + j * 2);
              mean[j] = t.x; rstd[j] = t.y; mean[j + 1] = t.z; rstd[j + 1] = t.w;
            }
            c_loaded = -1;
          }
          float u[8];
          fetch_act8(a, bi, yd, xd, c, mean, rstd, u);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += u[j];
        }
        const float nf = static_cast<float>(a.avg_n);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __fdiv_rn(v[j], nf);
      }
      interior = true;
    } else if constexpr (MODE == TSNET_TAPS_REFLECT1) {
      const int y = reflect_idx(yd - 1, a.H), x = reflect_idx(xd - 1, a.W);
      fetch_act8(a, b, y, x, c, mean, rstd, v);
      interior = (yd >= 1 && yd <= a.H && xd >= 1 && xd <= a.W);
    } else if constexpr (MODE == TSNET_TAPS_S2ZERO) {
      const int y = 2 * yd + (plane >> 1) - 1, x = 2 * xd + (plane & 1) - 1;
      if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
        fetch_act8(a, b, y, x, c, mean, rstd, v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
    } else {  // TSNET_TAPS_UP2REFLECT1
      const int uy = reflect_idx(yd - 1, 2 * a.H), ux = reflect_idx(xd - 1, 2 * a.W);
      // ATen area_pixel_compute_source_index(scale = 0.5, align_corners = false): max(0.5*(d+0.5)-0.5, 0)
      const float sy = fmaxf(0.5f * (uy + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (ux + 0.5f) - 0.5f, 0.f);
      const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
      const int y1 = min(y0 + 1, a.H - 1), x1 = min(x0 + 1, a.W - 1);
      const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
      float v00[8], v01[8], v10[8], v11[8];
      fetch_act8(a, b, y0, x0, c, mean, rstd, v00);
      fetch_act8(a, b, y0, x1, c, mean, rstd, v01);
      fetch_act8(a, b, y1, x0, c, mean, rstd, v10);
      fetch_act8(a, b, y1, x1, c, mean, rstd, v11);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t0 = __fadd_rn(__fmul_rn(hx, v00[j]), __fmul_rn(lx, v01[j]));
        const float t1 = __fadd_rn(__fmul_rn(hx, v10[j]), __fmul_rn(lx, v11[j]));
        v[j] = __fadd_rn(__fmul_rn(hy, t0), __fmul_rn(ly, t1));
      }
    }
    if (a.act_out && interior) {
      const int y = MODE == TSNET_TAPS_SAME ? yd : yd - 1, x = MODE == TSNET_TAPS_SAME ? xd : xd - 1;
      float* o = a.act_out + ((static_cast<size_t>(b) * a.H + y) * a.W + x) * a.act_C_total + a.act_c_off + c;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (a.hi) {
      uint16_t h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split16(v[j] * a.scale, a.fmt, h[j], l[j]);
      const size_t d = (((static_cast<size_t>(b) * a.planes + plane) * a.Hd + yd) * a.Wd + xd) * a.Cp_total +
                       a.c_off + c;
      uint4 ph, pl;
      ph.x = h[0] | (uint32_t(h[1]) << 16); ph.y = h[2] | (uint32_t(h[3]) << 16);
      ph.z = h[4] | (uint32_t(h[5]) << 16); ph.w = h[6] | (uint32_t(h[7]) << 16);
      pl.x = l[0] | (uint32_t(l[1]) << 16); pl.y = l[2] | (uint32_t(l[3]) << 16);
      pl.z = l[4] | (uint32_t(l[5]) << 16); pl.w = l[6] | (uint32_t(l[7]) << 16);
      *reinterpret_cast<uint4*>(a.hi + d) = ph;
      *reinterpret_cast<uint4*>(a.lo + d) = pl;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// UP2REFLECT1 specialisation: thread = one SOURCE pixel x 4 channels -> the 2 x 2 block of up-sampled pixels around it
// (and their reflect-pad duplicates on the border).  The generic kernel fetches + normalises 4 source values per
// destination pixel (16 per 2 x 2 block); here the 3 x 3 neighbourhood is fetched + normalised once (9 per block).
// Every destination value is computed with exactly the generic kernel's expression (ATen's source index / weights,
// same operation order), so the two are bit-identical.  grid.x = (b, source row); threads sweep (source column, c4).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fetch_act4(const TapsArgs& a, int b, int y, int x, int c, const float4& mean,
                                           const float4& rstd, float (&v)[4]) {
  const size_t off = ((static_cast<size_t>(b) * a.H + y) * a.W + x) * a.C + c;
  const float4 t = *reinterpret_cast<const float4*>(a.raw + off);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  if (a.mean_rstd) {
    v[0] = (v[0] - mean.x) * rstd.x; v[1] = (v[1] - mean.y) * rstd.y;
    v[2] = (v[2] - mean.z) * rstd.z; v[3] = (v[3] - mean.w) * rstd.w;
  }
  if (a.relu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (a.residual) {
    const float4 r = *reinterpret_cast<const float4*>(a.residual + off);
    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
  }
}

__global__ void __launch_bounds__(256, 4) build_taps_up2_kernel(const TapsArgs a) {
  const int cg = a.C / 4;
  const int ys = blockIdx.x % a.H;
  const int b = blockIdx.x / a.H;
  const int per_row = a.W * cg;
  const int H2 = 2 * a.H, W2 = 2 * a.W;
  const int yr[3] = {max(ys - 1, 0), ys, min(ys + 1, a.H - 1)};
  for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
    const int xs = idx / cg;
    const int c = (idx - xs * cg) * 4;
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), rstd = make_float4(1.f, 1.f, 1.f, 1.f);
    if (a.mean_rstd) {
      const float* mr = a.mean_rstd + (static_cast<size_t>(b) * a.C + c) * 2;
      const float4 t0 = *reinterpret_cast<const float4*>(mr), t1 = *reinterpret_cast<const float4*>(mr + 4);
      mean = make_float4(t0.x, t0.z, t1.x, t1.z);
      rstd = make_float4(t0.y, t0.w, t1.y, t1.w);
    }
    const int xr[3] = {max(xs - 1, 0), xs, min(xs + 1, a.W - 1)};
    float nb[3][3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) fetch_act4(a, b, yr[i], xr[j], c, mean, rstd, nb[i][j]);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int uy = 2 * ys + dy;
      // ATen area_pixel_compute_source_index(scale = 0.5, align_corners = false): max(0.5*(d+0.5)-0.5, 0)
      const float sy = fmaxf(0.5f * (uy + 0.5f) - 0.5f, 0.f);
      const int y0 = static_cast<int>(sy), y1 = min(y0 + 1, a.H - 1);
      const float ly = sy - y0, hy = 1.f - ly;
      (void)y1;
      // ATen's (y0, y1) are rows (ys-1, ys) for the even and (ys, ys+1) for the odd up-sampled row, i.e. nb rows
      // (dy, dy+1); on the border the clamped neighbour differs from ATen's y1 only where its weight ly is exactly 0
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ux = 2 * xs + dx;
        const float sx = fmaxf(0.5f * (ux + 0.5f) - 0.5f, 0.f);
        const int x0 = static_cast<int>(sx);
        const float lx = sx - x0, hx = 1.f - lx;
        uint16_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v00 = nb[dy][dx][j], v01 = nb[dy][dx + 1][j], v10 = nb[dy + 1][dx][j], v11 = nb[dy + 1][dx + 1][j];
          const float t0 = __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01));
          const float t1 = __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11));
          const float v = __fadd_rn(__fmul_rn(hy, t0), __fmul_rn(ly, t1));
          split16(v * a.scale, a.fmt, h[j], l[j]);
        }
        const uint2 ph = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
        const uint2 pl = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
        // destination rows / columns holding up-sampled pixel (uy, ux): uy + 1, plus the reflect-pad duplicates
        // (row 0 <- uy = 1, row 2H+1 <- uy = 2H-2; same for columns)
        int yd[2] = {uy + 1, -1}, xd[2] = {ux + 1, -1};
        if (uy == 1) yd[1] = 0;
        if (uy == H2 - 2) yd[1] = H2 + 1;
        if (ux == 1) xd[1] = 0;
        if (ux == W2 - 2) xd[1] = W2 + 1;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          if (yd[p] < 0) continue;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (xd[r] < 0) continue;
            const size_t d = ((static_cast<size_t>(b) * a.Hd + yd[p]) * a.Wd + xd[r]) * a.Cp_total + a.c_off + c;
            *reinterpret_cast<uint2*>(a.hi + d) = ph;
            *reinterpret_cast<uint2*>(a.lo + d) = pl;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stem tap source: [B, H+6, W, Cp], kw folded into channels.
// block = one destination row (b, yd) x 64 pixels.  Phase 1 stages the Cin input channels of the 70 source pixels
// (64 + 6 halo, reflect-indexed) in shared memory, pixel-major, with coalesced plane reads (CoordConv channels are
// generated, the /255 is applied).  Phase 2: thread = destination pixel x 8 folded channels = 8 consecutive floats of
// the staging buffer; consecutive threads write consecutive 16-byte pieces (fully coalesced hi / lo stores).
// lbl_kind: 0 = fp32 one-hot planes [B, Clbl, H, W]; 1 = uint8 class-index map [B, H, W] (utils/misc.py:50-67 vl2ch).
// ------------------------------------------------------------------------------------------------
constexpr int kStemTW = 64;

// img_kind: 0 = fp32 planes (mean-subtracted BGR, what the reference's datasets emit); 1 = uint8 BGR planes, the
// dataset's `image -= mean` (dataset/dataset_video_face.py:329, :401) is then applied here: (float(u8) - mean_c) / 255.
__global__ void __launch_bounds__(256) stem_taps_kernel(const void* __restrict__ img, int Cimg, int img_kind,
                                                        float mean0, float mean1, float mean2, float img_div,
                                                        const void* __restrict__ lbl, int Clbl, int lbl_kind, int B,
                                                        int H, int W, int Cp, int fmt, float scale,
                                                        uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  // pixel-major staging buffer [kStemTW + 6 source pixels][Cin]: folded channel j = s * Cin + c of destination pixel
  // px is source pixel px + s, channel c, i.e. element px * Cin + j -- the folded row of a pixel is a contiguous
  // window of 7 * Cin floats, no (tap, channel) tables needed
  extern __shared__ float s_src[];
  const int Cin = Cimg + Clbl + 3;
  const int Hd = H + 6;
  const int xt = blockIdx.x * kStemTW;
  const int yd = blockIdx.y;
  const int b = blockIdx.z;
  const int ys = reflect_idx(yd - 3, H);
  const size_t plane = static_cast<size_t>(H) * W;
  // Encoder.coord_conv (model/TSNet.py:117-122): t = idx / (n-1); 2*t - 1; r = sqrt(x^2 + y^2), separate roundings
  const float yy = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(static_cast<float>(ys), static_cast<float>(H - 1))), -1.f);
  constexpr int SW = kStemTW + 6;
  for (int i = threadIdx.x; i < Cin * SW; i += blockDim.x) {
    const int c = i / SW, k = i - c * SW;  // consecutive threads read consecutive pixels of one input plane
    const int xs = reflect_idx(xt + k - 3, W);
    float v;
    if (c < Cimg) {
      const size_t off = (static_cast<size_t>(b) * Cimg + c) * plane + static_cast<size_t>(ys) * W + xs;
      if (img_kind == 0) {
        v = __fdiv_rn(static_cast<const float*>(img)[off], img_div);
      } else {
        const float mc = c == 0 ? mean0 : (c == 1 ? mean1 : mean2);
        v = __fdiv_rn(__fadd_rn(static_cast<float>(static_cast<const uint8_t*>(img)[off]), -mc), img_div);
      }
    } else if (c < Cimg + Clbl) {
      if (lbl_kind == 0)
        v = static_cast<const float*>(lbl)[(static_cast<size_t>(b) * Clbl + (c - Cimg)) * plane +
                                           static_cast<size_t>(ys) * W + xs];
      else
        v = static_cast<const uint8_t*>(lbl)[static_cast<size_t>(b) * plane + static_cast<size_t>(ys) * W + xs] ==
                    (c - Cimg) ? 1.f : 0.f;
    } else {
      const float xx = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(static_cast<float>(xs), static_cast<float>(W - 1))), -1.f);
      const int kk = c - Cimg - Clbl;
      v = kk == 0 ? xx : (kk == 1 ? yy : __fsqrt_rn(__fadd_rn(__fmul_rn(xx, xx), __fmul_rn(yy, yy))));
    }
    s_src[k * Cin + c] = v;
  }
  __syncthreads();
  const int cg = Cp / 8;
  const int jmax = 7 * Cin;
  const size_t row_base = ((static_cast<size_t>(b) * Hd + yd) * W + xt) * Cp;
  for (int i = threadIdx.x; i < kStemTW * cg; i += blockDim.x) {
    const int px = i / cg, g = i - px * cg;
    if (xt + px >= W) break;
    const float* src = s_src + px * Cin + g * 8;
    uint16_t h[8], l[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float v = g * 8 + jj < jmax ? src[jj] : 0.f;
      split16(v * scale, fmt, h[jj], l[jj]);
    }
    uint4 ph, pl;
    ph.x = h[0] | (uint32_t(h[1]) << 16); ph.y = h[2] | (uint32_t(h[3]) << 16);
    ph.z = h[4] | (uint32_t(h[5]) << 16); ph.w = h[6] | (uint32_t(h[7]) << 16);
    pl.x = l[0] | (uint32_t(l[1]) << 16); pl.y = l[2] | (uint32_t(l[3]) << 16);
    pl.z = l[4] | (uint32_t(l[5]) << 16); pl.w = l[6] | (uint32_t(l[7]) << 16);
    *reinterpret_cast<uint4*>(hi + row_base + static_cast<size_t>(i) * 8) = ph;
    *reinterpret_cast<uint4*>(lo + row_base + static_cast<size_t>(i) * 8) = pl;
  }
}

// ------------------------------------------------------------------------------------------------
// F.normalize(dim = channel) + hi/lo split. One warp = one pixel; C % 128 == 0, C <= 1024.
// ------------------------------------------------------------------------------------------------
// rank (optional, [B, HW] uint16): row `pix` of sample b is written to row b*HW + rank[pix] -- the class-sorted
// operand order of the correlation kernel (tsnet_corr_prepare).
__global__ void __launch_bounds__(256) l2norm_split_kernel(const float* __restrict__ fea, size_t npix, int HW, int C,
                                                           int fmt, float scale, const uint16_t* __restrict__ rank,
                                                           uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x / 32) + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int lane = threadIdx.x & 31;
  const float* p = fea + pix * C;
  float4 v[8];
  float ss = 0.f;
  const int nk = C / 128;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < nk) {
      v[k] = *reinterpret_cast<const float4*>(p + k * 128 + lane * 4);
      ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(__fsqrt_rn(ss), 1e-12f);  // F.normalize eps
  // x * (scale / denom) instead of (x / denom) * scale: one IEEE division per pixel instead of one per element (the
  // kernel was issue-bound on the division sequences, 65 % SM throughput at 3 TB/s); the <= 1 ulp (2^-24) difference is
  // below the 2^-22 resolution of the hi/lo operands it feeds
  const float rs = __fdiv_rn(scale, denom);
  const size_t drow = rank ? (pix / HW) * HW + rank[pix] : pix;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < nk) {
      const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
      uint16_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split16(e[j] * rs, fmt, h[j], l[j]);
      const size_t d = drow * C + k * 128 + lane * 4;
      *reinterpret_cast<uint2*>(hi + d) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
      *reinterpret_cast<uint2*>(lo + d) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Correlation operands WITHOUT the normalisation: x * scale split into hi / lo at the row's sorted rank, plus
// rnorm = 1 / max(||x||_2, 1e-12) (F.normalize's clamp) at the same rank; the tile kernel applies rnorm_t * rnorm_s as
// a scale of the similarity.  One warp = one pixel; no per-element division (l2norm_split was issue-bound on them).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) corr_operands_kernel(const float* __restrict__ fea, size_t npix, int HW, int C,
                                                            int fmt, float scale, const uint16_t* __restrict__ rank,
                                                            uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                            float* __restrict__ rnorm) {
  const size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x / 32) + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int lane = threadIdx.x & 31;
  const float* p = fea + pix * C;
  float4 v[8];
  float ss = 0.f;
  const int nk = C / 128;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < nk) {
      v[k] = *reinterpret_cast<const float4*>(p + k * 128 + lane * 4);
      ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const size_t drow = rank ? (pix / HW) * HW + rank[pix] : pix;
  if (lane == 0) rnorm[drow] = __fdiv_rn(1.f, fmaxf(__fsqrt_rn(ss), 1e-12f));
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < nk) {
      const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
      uint16_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split16(e[j] * scale, fmt, h[j], l[j]);
      const size_t d = drow * C + k * 128 + lane * 4;
      *reinterpret_cast<uint2*>(hi + d) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
      *reinterpret_cast<uint2*>(lo + d) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
    }
  }
}

// rnorm of the sources from the per-slab partial sums of squares the bridge pass wrote: [B][slabs][HW] -> [B * HW] at the
// sorted rank.  One thread per pixel, fixed summation order.
__global__ void __launch_bounds__(256) corr_norms_kernel(const float* __restrict__ ssq_part, size_t npix, int HW,
                                                         int slabs, const uint16_t* __restrict__ rank,
                                                         float* __restrict__ rnorm) {
  const size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (pix >= npix) return;
  const size_t b = pix / HW, p = pix - b * HW;
  float ss = 0.f;
  for (int k = 0; k < slabs; ++k) ss += ssq_part[(b * slabs + k) * HW + p];
  rnorm[b * HW + (rank ? rank[pix] : p)] = __fdiv_rn(1.f, fmaxf(__fsqrt_rn(ss), 1e-12f));
}

// ------------------------------------------------------------------------------------------------
// output head: reflect-pad 3, 7x7 conv Cin -> 3, tanh, optional pose compositing, NCHW store.
// block = 32 x 16 output pixels, 128 threads; thread = 4 vertically adjacent pixels x 3 outputs.
// Channels go through shared memory in chunks of 8 (pixel stride 12 floats: conflict-free LDS.128).  For every
// column tap the thread keeps a 10-row sliding window in registers, so each activation LDS feeds 84 FMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadTW = 32, kHeadTH = 16, kHeadCC = 8, kHeadPS = 12, kHeadPPT = 4;

__global__ void __launch_bounds__(128) head_conv_kernel(const float* __restrict__ act,
                                                        const float* __restrict__ mean_rstd, int relu, int B, int H,
                                                        int W, int Cin, const float* __restrict__ w,
                                                        const float* __restrict__ bias,
                                                        int fore_x0, int fore_x1, float fill0, float fill1,
                                                        float fill2, float* __restrict__ out) {
  __shared__ __align__(16) float s_in[(kHeadTH + 6) * (kHeadTW + 6) * kHeadPS];
  __shared__ __align__(16) float s_w[49 * kHeadCC * 4];  // [tap][c][4] (3 outputs + pad)
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kHeadTW, y0 = blockIdx.y * kHeadTH;
  const int tx = threadIdx.x % kHeadTW, ty = (threadIdx.x / kHeadTW) * kHeadPPT;
  float acc[kHeadPPT][3];
#pragma unroll
  for (int p = 0; p < kHeadPPT; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0.f;
  for (int cc = 0; cc < Cin; cc += kHeadCC) {
    __syncthreads();
    for (int i = threadIdx.x; i < (kHeadTH + 6) * (kHeadTW + 6) * (kHeadCC / 4); i += blockDim.x) {
      const int c4 = i % (kHeadCC / 4);
      const int p = i / (kHeadCC / 4);
      const int px = p % (kHeadTW + 6), py = p / (kHeadTW + 6);
      const int ys = reflect_idx(y0 + py - 3, H), xs = reflect_idx(x0 + px - 3, W);
      float4 v =
          *reinterpret_cast<const float4*>(act + ((static_cast<size_t>(b) * H + ys) * W + xs) * Cin + cc + c4 * 4);
      if (mean_rstd) {  // InstanceNorm (+ ReLU) of the raw conv output applied while staging (model/TSNet.py:149-150)
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(mean_rstd + (static_cast<size_t>(b) * Cin + cc + c4 * 4) * 2));
        const float4 m1 = __ldg(reinterpret_cast<const float4*>(mean_rstd + (static_cast<size_t>(b) * Cin + cc + c4 * 4) * 2 + 4));
        v.x = (v.x - m0.x) * m0.y; v.y = (v.y - m0.z) * m0.w; v.z = (v.z - m1.x) * m1.y; v.w = (v.w - m1.z) * m1.w;
      }
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(&s_in[p * kHeadPS + c4 * 4]) = v;
    }
    for (int i = threadIdx.x; i < 49 * kHeadCC; i += blockDim.x) {
      const int c = i % kHeadCC, tap = i / kHeadCC;
      float4 wv;
      wv.x = w[(static_cast<size_t>(0) * Cin + cc + c) * 49 + tap];
      wv.y = w[(static_cast<size_t>(1) * Cin + cc + c) * 49 + tap];
      wv.z = w[(static_cast<size_t>(2) * Cin + cc + c) * 49 + tap];
      wv.w = 0.f;
      *reinterpret_cast<float4*>(&s_w[i * 4]) = wv;
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < 7; ++s) {
#pragma unroll
      for (int c4 = 0; c4 < kHeadCC / 4; ++c4) {
        float4 col[kHeadPPT + 6];
#pragma unroll
        for (int rr = 0; rr < kHeadPPT + 6; ++rr)
          col[rr] = *reinterpret_cast<const float4*>(&s_in[((ty + rr) * (kHeadTW + 6) + tx + s) * kHeadPS + c4 * 4]);
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          const float* wp = &s_w[((r * 7 + s) * kHeadCC + c4 * 4) * 4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + j * 4);
#pragma unroll
            for (int p = 0; p < kHeadPPT; ++p) {
              const float4 cv = col[p + r];
              const float e = j == 0 ? cv.x : (j == 1 ? cv.y : (j == 2 ? cv.z : cv.w));
              acc[p][0] = fmaf(e, wv.x, acc[p][0]);
              acc[p][1] = fmaf(e, wv.y, acc[p][1]);
              acc[p][2] = fmaf(e, wv.z, acc[p][2]);
            }
          }
        }
      }
    }
  }
  const int x = x0 + tx;
  const size_t plane = static_cast<size_t>(H) * W;
#pragma unroll
  for (int p = 0; p < kHeadPPT; ++p) {
    const int y = y0 + ty + p;
    if (x < W && y < H) {
      float o[3] = {tanhf(acc[p][0] + bias[0]), tanhf(acc[p][1] + bias[1]), tanhf(acc[p][2] + bias[2])};
      if (fore_x1 > fore_x0) {
        const float fore = (x >= fore_x0 && x < fore_x1) ? 1.f : 0.f;
        const float fill[3] = {fill0, fill1, fill2};
#pragma unroll
        for (int k = 0; k < 3; ++k) o[k] = __fadd_rn(__fmul_rn(o[k], fore), __fmul_rn(fill[k], 1.f - fore));
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) out[(static_cast<size_t>(b) * 3 + k) * plane + static_cast<size_t>(y) * W + x] = o[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// demo post-processing (demo/demo_face.py:96-105, 194-199): per image and channel, re-normalise the generated frame
// to the reference statistics of the source video, add the dataset mean, clamp, x255, BGR -> RGB, uint8.
//   y = (x - mean_c) / std_c * ref_std_c + ref_mean_c ;  y = clamp(y + img_mean_c, 0, 1) * 255 ;  out[.., 2 - c] = (u8) y
// mean / std (unbiased, torch.std) of the generated frame come from tsnet_plane_stats (fp64, fixed order): given the
// same statistics the byte output equals the reference's fp32 arithmetic bit for bit (every operation individually
// rounded, truncating uint8 conversion).  grid = (pixel chunks, 3, B).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) postprocess_u8_kernel(const float* __restrict__ x, int HW,
                                                             const float* __restrict__ gen_mean_std,
                                                             const float* __restrict__ ref_mean,
                                                             const float* __restrict__ ref_std, float m0, float m1,
                                                             float m2, uint8_t* __restrict__ out) {
  const int c = blockIdx.y, b = blockIdx.z;
  const float* p = x + (static_cast<size_t>(b) * 3 + c) * HW;
  const float mean = gen_mean_std[(b * 3 + c) * 2], stdv = gen_mean_std[(b * 3 + c) * 2 + 1];
  const float rs = ref_std[c], rm = ref_mean[c];
  const float im = c == 0 ? m0 : (c == 1 ? m1 : m2);
  uint8_t* o = out + static_cast<size_t>(b) * HW * 3 + (2 - c);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float y = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(p[i], -mean), stdv), rs), rm);
    y = __fadd_rn(y, im);
    y = __fmul_rn(fminf(fmaxf(y, 0.f), 1.f), 255.f);
    o[static_cast<size_t>(i) * 3] = static_cast<uint8_t>(y);  // astype('uint8'): truncation
  }
}

// ------------------------------------------------------------------------------------------------
// validation-only direct convolution, fp32 FMA, NHWC. One thread = one output element.
// ------------------------------------------------------------------------------------------------
__global__ void direct_conv_kernel(const float* __restrict__ x, int B, int H, int W, int Cin,
                                   const float* __restrict__ w, const float* __restrict__ bias, int Cout, int K,
                                   int stride, int pad, int reflect, int Ho, int Wo, float* __restrict__ y) {
  const size_t total = static_cast<size_t>(B) * Ho * Wo * Cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);
    size_t p = i / Cout;
    const int xo = static_cast<int>(p % Wo);
    p /= Wo;
    const int yo = static_cast<int>(p % Ho);
    const int b = static_cast<int>(p / Ho);
    float acc = 0.f;
    for (int r = 0; r < K; ++r) {
      int yi = yo * stride + r - pad;
      if (reflect) yi = reflect_idx(yi, H);
      if (yi < 0 || yi >= H) continue;
      for (int s = 0; s < K; ++s) {
        int xi = xo * stride + s - pad;
        if (reflect) xi = reflect_idx(xi, W);
        if (xi < 0 || xi >= W) continue;
        const float* xp = x + ((static_cast<size_t>(b) * H + yi) * W + xi) * Cin;
        const float* wp = w + static_cast<size_t>(co) * Cin * K * K + r * K + s;
        for (int c = 0; c < Cin; ++c) acc = fmaf(xp[c], wp[static_cast<size_t>(c) * K * K], acc);
      }
    }
    y[i] = acc + (bias ? bias[co] : 0.f);
  }
}

int launch_wino_input(const WinoInArgs& a, cudaStream_t stream);  // winograd.cu

static inline int grid_for(size_t total, int block, int max_blocks = 148 * 16) {
  size_t g = (total + block - 1) / block;
  return static_cast<int>(g < static_cast<size_t>(max_blocks) ? (g ? g : 1) : max_blocks);
}

}  // namespace tsnet

using namespace tsnet;

namespace tsnet {
std::atomic<long long>& launch_counter() {
  static std::atomic<long long> n{0};
  return n;
}
}  // namespace tsnet

extern "C" int tsnet_abi_version(void) { return TSNET_ABI_VERSION; }
extern "C" long long tsnet_launch_count(void) { return tsnet::launch_counter().load(); }
extern "C" const char* tsnet_last_error(void) { return last_error_buf(); }
extern "C" int tsnet_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int tsnet_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int KH, int KW, int fold_kw, int Cp,
                                      int Cout_pad, float scale, int fmt, uint16_t* w_hi, uint16_t* w_lo,
                                      void* stream) {
  TSNET_ARG_CHECK(w_oihw && w_hi && w_lo, "pack_conv_weight: null argument");
  TSNET_ARG_CHECK(fold_kw <= 1 || fold_kw >= Cin, "pack_conv_weight: fold_kw %d < Cin %d", fold_kw, Cin);
  TSNET_ARG_CHECK(Cp % 64 == 0 && Cp >= (fold_kw ? KW * (fold_kw > 1 ? fold_kw : Cin) : Cin),
                  "pack_conv_weight: Cp %d too small", Cp);
  TSNET_ARG_CHECK(Cout_pad >= Cout, "pack_conv_weight: Cout_pad");
  const size_t total = static_cast<size_t>(Cout_pad) * (fold_kw ? KH : KH * KW) * Cp;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, Cout, Cin, KH, KW, fold_kw, Cp, Cout_pad, scale, fmt, w_hi, w_lo);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_instnorm_reduce(const float* stats_partial, int B, int HW, int C, float eps, float* mean_rstd,
                                     void* stream) {
  TSNET_ARG_CHECK(stats_partial && mean_rstd, "instnorm_reduce: null argument");
  TSNET_ARG_CHECK(HW % 32 == 0, "instnorm_reduce: HW %d", HW);
  dim3 grid((C + 31) / 32, B), block(32, kInSegs);
  instnorm_reduce_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(stats_partial, HW / 32, C, eps,
                                                                                mean_rstd);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_build_taps(const tsnet_taps_desc* d, const float* raw, const float* mean_rstd,
                                const float* residual, float* act_out, uint16_t* taps_hi, uint16_t* taps_lo,
                                void* stream) {
  TSNET_ARG_CHECK(d && raw, "build_taps: null argument");
  TSNET_ARG_CHECK(d->C % 8 == 0, "build_taps: C %d must be a multiple of 8", d->C);
  TSNET_ARG_CHECK((taps_hi == nullptr) == (taps_lo == nullptr), "build_taps: hi/lo must both be given or both NULL");
  TSNET_ARG_CHECK(!taps_hi || (d->Cp_total % 8 == 0 && d->c_off % 8 == 0 && d->c_off + d->C <= d->Cp_total),
                  "build_taps: channel window [%d, %d) does not fit Cp_total %d", d->c_off, d->c_off + d->C,
                  d->Cp_total);
  TSNET_ARG_CHECK(!(act_out && (d->mode == TSNET_TAPS_S2ZERO || d->mode == TSNET_TAPS_UP2REFLECT1)),
                  "build_taps: act_out is not available in mode %d", d->mode);
  TapsArgs a;
  a.raw = raw; a.mean_rstd = mean_rstd; a.residual = residual; a.act_out = act_out; a.hi = taps_hi; a.lo = taps_lo;
  a.B = d->B; a.H = d->H; a.W = d->W; a.C = d->C; a.mode = d->mode; a.relu = d->relu; a.Cp_total = d->Cp_total;
  a.c_off = d->c_off; a.fmt = d->fmt; a.scale = d->scale == 0.f ? 1.f : d->scale;
  a.act_C_total = d->act_C_total > 0 ? d->act_C_total : d->C;
  a.act_c_off = d->act_c_off;
  a.avg_n = d->avg_n > 1 ? d->avg_n : 1;
  a.residual2 = d->residual2; a.res_split = d->res_split; a.res2_batch = d->res2_batch;
  TSNET_ARG_CHECK(!d->residual2 || (residual && d->res_split > 0 && d->res_split < d->C && d->res_split % 8 == 0 &&
                                    (d->C - d->res_split) % 8 == 0 && d->res2_batch > 0 && d->avg_n <= 1 &&
                                    (d->mode == TSNET_TAPS_SAME || d->mode == TSNET_TAPS_REFLECT1 ||
                                     d->mode == TSNET_TAPS_S2ZERO)),
                  "build_taps: two-part residual needs mode SAME / REFLECT1 / S2ZERO, res_split %% 8 == 0, res2_batch > 0");
  TSNET_ARG_CHECK(a.avg_n == 1 || d->mode == TSNET_TAPS_SAME, "build_taps: avg_n needs mode SAME");
  TSNET_ARG_CHECK(a.act_c_off % 4 == 0 && a.act_C_total % 4 == 0 && a.act_c_off + d->C <= a.act_C_total,
                  "build_taps: act_out channel window does not fit");
  if (d->mode == TSNET_TAPS_WINO) {  // Winograd input transform (wino_passes.cuh / winograd.cu)
    TSNET_ARG_CHECK(a.avg_n == 1, "build_taps: avg_n needs mode SAME");
    WinoInArgs w;
    w.raw = raw; w.mean_rstd = mean_rstd; w.residual = residual; w.act_out = act_out; w.hi = taps_hi; w.lo = taps_lo;
    w.B = d->B; w.H = d->H; w.W = d->W; w.C = d->C; w.relu = d->relu; w.Cp_total = d->Cp_total; w.c_off = d->c_off;
    w.fmt = d->fmt; w.act_C_total = a.act_C_total; w.act_c_off = a.act_c_off; w.scale = a.scale;
    return launch_wino_input(w, static_cast<cudaStream_t>(stream));
  }
  a.planes = 1;
  switch (d->mode) {
    case TSNET_TAPS_SAME: a.Hd = d->H; a.Wd = d->W; break;
    case TSNET_TAPS_REFLECT1: a.Hd = d->H + 2; a.Wd = d->W + 2; break;
    case TSNET_TAPS_S2ZERO:
      TSNET_ARG_CHECK(d->H % 2 == 0 && d->W % 2 == 0, "build_taps: stride-2 needs even H, W");
      a.Hd = d->H / 2 + 1; a.Wd = d->W / 2 + 1; a.planes = 4; break;
    case TSNET_TAPS_UP2REFLECT1: a.Hd = 2 * d->H + 2; a.Wd = 2 * d->W + 2; break;
    default: return set_error(-1, "build_taps: unknown mode %d", d->mode);
  }
  const unsigned rows = static_cast<unsigned>(a.B) * a.planes * a.Hd;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d->mode) {
    case TSNET_TAPS_SAME: build_taps_kernel<TSNET_TAPS_SAME><<<rows, 256, 0, st>>>(a); break;
    case TSNET_TAPS_REFLECT1: build_taps_kernel<TSNET_TAPS_REFLECT1><<<rows, 256, 0, st>>>(a); break;
    case TSNET_TAPS_S2ZERO: build_taps_kernel<TSNET_TAPS_S2ZERO><<<rows, 256, 0, st>>>(a); break;
    default:
      // quad kernel (9 instead of 16 normalised fetches per 2 x 2 block); the generic one stays for odd shapes / tests
      if (a.hi && !act_out && a.avg_n == 1 && d->C % 4 == 0 && d->H >= 2 && d->W >= 2 &&
          (d->flags & TSNET_TAPS_GENERIC_UP2) == 0)
        build_taps_up2_kernel<<<static_cast<unsigned>(a.B) * a.H, 256, 0, st>>>(a);
      else
        build_taps_kernel<TSNET_TAPS_UP2REFLECT1><<<rows, 256, 0, st>>>(a);
      break;
  }
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_stem_taps(const void* img_nchw, int Cimg, int img_kind, const float* img_mean3_host,
                               float img_div, const void* lbl, int Clbl, int lbl_kind, int B, int H, int W, int Cp,
                               int fmt, float scale, uint16_t* taps_hi, uint16_t* taps_lo, void* stream) {
  TSNET_ARG_CHECK(lbl && taps_hi && taps_lo, "stem_taps: null argument");
  TSNET_ARG_CHECK((img_nchw != nullptr) == (Cimg > 0), "stem_taps: img pointer / Cimg mismatch");
  TSNET_ARG_CHECK(img_kind == 0 || (img_kind == 1 && Cimg == 3 && img_mean3_host),
                  "stem_taps: img_kind %d (uint8 images need Cimg == 3 and a host pointer to the 3 channel means)", img_kind);
  TSNET_ARG_CHECK(lbl_kind == 0 || lbl_kind == 1, "stem_taps: lbl_kind %d", lbl_kind);
  TSNET_ARG_CHECK(Cp % 64 == 0 && Cp >= 7 * (Cimg + Clbl + 3) && Cp <= 1024, "stem_taps: Cp %d out of range", Cp);
  TSNET_ARG_CHECK(Cimg + Clbl + 3 <= 127, "stem_taps: too many input channels");
  TSNET_ARG_CHECK(W % kStemTW == 0 && W >= 4 && H >= 4, "stem_taps: W %d must be a multiple of %d", W, kStemTW);
  dim3 grid(W / kStemTW, H + 6, B);
  const size_t smem = static_cast<size_t>(Cimg + Clbl + 3) * (kStemTW + 6) * sizeof(float);
  const float m0 = img_kind == 1 ? img_mean3_host[0] : 0.f, m1 = img_kind == 1 ? img_mean3_host[1] : 0.f,
              m2 = img_kind == 1 ? img_mean3_host[2] : 0.f;
  stem_taps_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      img_nchw, Cimg, img_kind, m0, m1, m2, img_div == 0.f ? 1.f : img_div, lbl, Clbl, lbl_kind, B, H, W, Cp, fmt,
      scale == 0.f ? 1.f : scale, taps_hi, taps_lo);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_l2norm_split(const float* fea, int B, int HW, int C, int fmt, float scale, const uint16_t* rank,
                                  uint16_t* out_hi, uint16_t* out_lo, void* stream) {
  TSNET_ARG_CHECK(fea && out_hi && out_lo, "l2norm_split: null argument");
  TSNET_ARG_CHECK(C % 128 == 0 && C <= 1024, "l2norm_split: C %d must be a multiple of 128, <= 1024", C);
  const size_t npix = static_cast<size_t>(B) * HW;
  l2norm_split_kernel<<<static_cast<unsigned>((npix + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      fea, npix, HW, C, fmt, scale == 0.f ? 1.f : scale, rank, out_hi, out_lo);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_corr_operands(const float* fea, int B, int HW, int C, int fmt, float scale, const uint16_t* rank,
                                   uint16_t* out_hi, uint16_t* out_lo, float* rnorm, void* stream) {
  TSNET_ARG_CHECK(fea && out_hi && out_lo && rnorm, "corr_operands: null argument");
  TSNET_ARG_CHECK(C % 128 == 0 && C <= 1024, "corr_operands: C %d must be a multiple of 128, <= 1024", C);
  const size_t npix = static_cast<size_t>(B) * HW;
  corr_operands_kernel<<<static_cast<unsigned>((npix + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      fea, npix, HW, C, fmt, scale == 0.f ? 1.f : scale, rank, out_hi, out_lo, rnorm);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_corr_norms(const float* ssq_part, int B, int HW, int slabs, const uint16_t* rank, float* rnorm,
                                void* stream) {
  TSNET_ARG_CHECK(ssq_part && rnorm && slabs > 0, "corr_norms: bad argument");
  const size_t npix = static_cast<size_t>(B) * HW;
  corr_norms_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ssq_part, npix, HW, slabs, rank, rnorm);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_head_conv_tanh(const float* act_nhwc, const float* mean_rstd, int relu, int B, int H, int W,
                                    int Cin, const float* w_oihw, const float* bias, int fore_x0, int fore_x1,
                                    const float* fill3, float* out_nchw, void* stream) {
  TSNET_ARG_CHECK(act_nhwc && w_oihw && bias && out_nchw, "head_conv: null argument");
  TSNET_ARG_CHECK(Cin % kHeadCC == 0, "head_conv: Cin %d must be a multiple of %d", Cin, kHeadCC);
  TSNET_ARG_CHECK(H >= 4 && W >= 4, "head_conv: image too small for reflect pad 3");
  TSNET_ARG_CHECK(fore_x1 <= fore_x0 || fill3, "head_conv: compositing needs fill3 (host pointer to 3 floats)");
  dim3 grid((W + kHeadTW - 1) / kHeadTW, (H + kHeadTH - 1) / kHeadTH, B);
  const float f0 = fill3 ? fill3[0] : 0.f, f1 = fill3 ? fill3[1] : 0.f, f2 = fill3 ? fill3[2] : 0.f;
  head_conv_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(act_nhwc, mean_rstd, relu, B, H, W, Cin,
                                                                          w_oihw, bias, fore_x0, fore_x1, f0, f1, f2,
                                                                          out_nchw);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_postprocess_u8(const float* rec_nchw, int B, int H, int W, const float* gen_mean_std,
                                    const float* ref_mean3, const float* ref_std3, const float* img_mean3_host,
                                    uint8_t* out_hwc_rgb, void* stream) {
  TSNET_ARG_CHECK(rec_nchw && gen_mean_std && ref_mean3 && ref_std3 && img_mean3_host && out_hwc_rgb,
                  "postprocess_u8: null argument");
  TSNET_ARG_CHECK(H * W >= 2, "postprocess_u8: image too small");
  const int chunks = (H * W + 256 * 8 - 1) / (256 * 8);
  dim3 grid(chunks, 3, B);
  postprocess_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rec_nchw, H * W, gen_mean_std, ref_mean3, ref_std3, img_mean3_host[0], img_mean3_host[1], img_mean3_host[2],
      out_hwc_rgb);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_direct_conv_fp32(const float* x_nhwc, int B, int H, int W, int Cin, const float* w_oihw,
                                      const float* bias, int Cout, int K, int stride, int pad, int reflect,
                                      float* y_nhwc, void* stream) {
  TSNET_ARG_CHECK(x_nhwc && w_oihw && y_nhwc, "direct_conv: null argument");
  const int Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * Cout;
  direct_conv_kernel<<<grid_for(total, 256, 148 * 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nhwc, B, H, W, Cin, w_oihw, bias, Cout, K, stride, pad, reflect, Ho, Wo, y_nhwc);
  TSNET_LAUNCH_CHECK();
  return 0;
}
