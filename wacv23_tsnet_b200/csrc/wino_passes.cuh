// Winograd F(2x2, 3x3) transform passes of the 32 x 32 ResnetBlock convolutions (model/TSNet.py:10-49 of the
// reference: ReflectionPad2d(1) + Conv2d(dim, dim, 3) + InstanceNorm [+ ReLU | + x]).
//
//   Y = A^T [ sum_c (G g G^T) .* (B^T d B) ] A          (Lavin & Gray; d = 4 x 4 input tile, Y = 2 x 2 outputs)
//
//   wino_weight_body   U[p][o][c]   = (G g[o][c] G^T)[i][j], p = 4 i + j           (once per weight version)
//   wino_input_body    V[b][p][ty][tx][c] = (B^T d B)[i][j] of the normalised / activated / reflect-padded input,
//                      written as the 16-bit hi / lo operands of the 16 plane GEMMs     (HBM-bound pass "T")
//   wino_output_body   y[b][2ty+a][2tx+e][o] = (A^T M A)[a][e] + bias (+ addend), plus the InstanceNorm partial
//                      statistics of y                                                  (HBM-bound pass "I")
//
// The bodies are plain __host__ __device__ functions of (block, thread) indices without shared memory or shuffles:
// the same source is compiled into the CUDA kernels (winograd.cu) and into a host emulation used by the CPU test
// suite (oracle/wino_emul.cu), so the index math is checked on the CPU box before any GPU time is spent.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#if defined(__CUDACC__)
#define TSNET_HD __host__ __device__ __forceinline__
#else
#define TSNET_HD inline
#endif

namespace tsnet {

struct f4 {
  float x, y, z, w;
};
TSNET_HD f4 f4_add(const f4& a, const f4& b) { return f4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
TSNET_HD f4 f4_sub(const f4& a, const f4& b) { return f4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }

TSNET_HD f4 ld_f4(const float* p) {
#if defined(__CUDA_ARCH__)
  const float4 t = *reinterpret_cast<const float4*>(p);
  return f4{t.x, t.y, t.z, t.w};
#else
  return f4{p[0], p[1], p[2], p[3]};
#endif
}
TSNET_HD void st_f4(float* p, const f4& v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
#else
  p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
#endif
}

// x = hi + lo in the 16-bit operand format (same function as split16 of sm100_prims.cuh; host + device).
// fp16 saturates at the largest finite value instead of producing inf - inf = NaN for |x| > 65504.
TSNET_HD void wino_split16(float x, int fmt, uint16_t& hi, uint16_t& lo) {
  if (fmt == 1) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
  } else {
    x = x > 65504.f ? 65504.f : (x < -65504.f ? -65504.f : x);
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
  }
}

TSNET_HD int wino_reflect(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

constexpr int kWinoRun = 8;  // tiles one thread walks along a tile row (= 32 output pixels = one statistics partial)

// Winograd-domain layouts in HBM (both chosen so that a pass that owns an (image, channel slab) pair streams contiguous
// memory, and so that the plane GEMMs read / write whole 128-byte lines):
//   V (16-bit hi and lo operands of the plane GEMMs)  [B][16 planes][Cp / 64 K blocks][T tiles][64]   (K-block-major)
//   M (fp32 results of the plane GEMMs)               [16 planes][B][C / 32 slabs][T tiles][32]       (slab-major)
constexpr int kWinoMSlab = 32;
TSNET_HD size_t wino_v_index(int b, int p, int T, int Cp_total, int tile, int cc) {
  return (((static_cast<size_t>(b) * 16 + p) * (Cp_total >> 6) + (cc >> 6)) * T + tile) * 64 + (cc & 63);
}
TSNET_HD size_t wino_m_index(int B, int C, int T, int p, int b, int tile, int c) {
  return (((static_cast<size_t>(p) * B + b) * (C / kWinoMSlab) + c / kWinoMSlab) * T + tile) * kWinoMSlab +
         (c % kWinoMSlab);
}

// ------------------------------------------------------------------------------------------------
// weights: one thread per (o, c).  fp64 arithmetic, fp32 result [16][Cout][Cin].
// ------------------------------------------------------------------------------------------------
TSNET_HD void wino_weight_body(const float* w_oihw, int Cout, int Cin, float* u, size_t idx) {
  const size_t total = static_cast<size_t>(Cout) * Cin;
  if (idx >= total) return;
  const float* g = w_oihw + idx * 9;
  double gg[4][3];
  for (int s = 0; s < 3; ++s) {
    const double g0 = g[0 * 3 + s], g1 = g[1 * 3 + s], g2 = g[2 * 3 + s];
    gg[0][s] = g0;
    gg[1][s] = 0.5 * (g0 + g1 + g2);
    gg[2][s] = 0.5 * (g0 - g1 + g2);
    gg[3][s] = g2;
  }
  for (int i = 0; i < 4; ++i) {
    const double u0 = gg[i][0], u1 = 0.5 * (gg[i][0] + gg[i][1] + gg[i][2]), u2 = 0.5 * (gg[i][0] - gg[i][1] + gg[i][2]),
                 u3 = gg[i][2];
    u[(static_cast<size_t>(i * 4 + 0)) * total + idx] = static_cast<float>(u0);
    u[(static_cast<size_t>(i * 4 + 1)) * total + idx] = static_cast<float>(u1);
    u[(static_cast<size_t>(i * 4 + 2)) * total + idx] = static_cast<float>(u2);
    u[(static_cast<size_t>(i * 4 + 3)) * total + idx] = static_cast<float>(u3);
  }
}

// ------------------------------------------------------------------------------------------------
// pass T: input transform.  Same producer options as build_taps (tsnet_taps_desc): InstanceNorm, ReLU, residual,
// fp32 act_out.  block = one tile row (b, ty); a thread owns 4 channels and walks kWinoRun consecutive tiles, carrying
// the two overlapping (already row-transformed) columns in registers, so every activation is read and normalised once
// per tile row.
// ------------------------------------------------------------------------------------------------
struct WinoInArgs {
  const float* raw;        // fp32 [B, H, W, C]
  const float* mean_rstd;  // [B, C, 2] or null
  const float* residual;   // fp32 [B, H, W, C] or null
  float* act_out;          // fp32 [B, H, W, act_C_total] window [act_c_off, +C) or null
  uint16_t* hi;            // V layout (wino_v_index), channel window [c_off, +C) of Cp_total
  uint16_t* lo;
  int B, H, W, C, relu, Cp_total, c_off, fmt, act_C_total, act_c_off;
  float scale;
};

TSNET_HD f4 wino_fetch(const WinoInArgs& a, int b, int y, int x, int c, const f4& mean, const f4& rstd) {
  const size_t off = ((static_cast<size_t>(b) * a.H + y) * a.W + x) * a.C + c;
  f4 v = ld_f4(a.raw + off);
  if (a.mean_rstd) {
    v.x = (v.x - mean.x) * rstd.x; v.y = (v.y - mean.y) * rstd.y;
    v.z = (v.z - mean.z) * rstd.z; v.w = (v.w - mean.w) * rstd.w;
  }
  if (a.relu) {
    v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
    v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
  }
  if (a.residual) v = f4_add(v, ld_f4(a.residual + off));
  return v;
}

TSNET_HD void wino_store_plane(const WinoInArgs& a, size_t d, const f4& v) {
  uint16_t h[4], l[4];
  wino_split16(v.x * a.scale, a.fmt, h[0], l[0]);
  wino_split16(v.y * a.scale, a.fmt, h[1], l[1]);
  wino_split16(v.z * a.scale, a.fmt, h[2], l[2]);
  wino_split16(v.w * a.scale, a.fmt, h[3], l[3]);
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<uint2*>(a.hi + d) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
  *reinterpret_cast<uint2*>(a.lo + d) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
#else
  for (int j = 0; j < 4; ++j) {
    a.hi[d + j] = h[j];
    a.lo[d + j] = l[j];
  }
#endif
}

// Store the 16 planes of one tile with ONE running (hi, lo) pointer pair that advances by the plane stride: keeps the
// compiler from materialising 32 destination pointers at once (176 bytes of spills in the bridge kernel otherwise).
struct WinoPlaneWriter {
  uint16_t* ph;
  uint16_t* pl;
  size_t pstride;
  float scale;
  int fmt;
};
TSNET_HD void wino_write_next(WinoPlaneWriter& w, const f4& v) {
  uint16_t h[4], l[4];
  wino_split16(v.x * w.scale, w.fmt, h[0], l[0]);
  wino_split16(v.y * w.scale, w.fmt, h[1], l[1]);
  wino_split16(v.z * w.scale, w.fmt, h[2], l[2]);
  wino_split16(v.w * w.scale, w.fmt, h[3], l[3]);
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<uint2*>(w.ph) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
  *reinterpret_cast<uint2*>(w.pl) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
  w.ph += w.pstride;
  w.pl += w.pstride;
  asm volatile("" : "+l"(w.ph), "+l"(w.pl));  // sequential dependence: no precomputed pointer table
#else
  for (int j = 0; j < 4; ++j) {
    w.ph[j] = h[j];
    w.pl[j] = l[j];
  }
  w.ph += w.pstride;
  w.pl += w.pstride;
#endif
}

// column transform B^T d of one input column (rows ys[0..3]) -> t[0..3]; writes act_out for the two rows the tile row
// owns (r = 1, 2) when this column is owned by the calling thread
TSNET_HD void wino_in_column(const WinoInArgs& a, int b, const int (&ys)[4], int x, bool own, int c, const f4& mean,
                             const f4& rstd, f4 (&t)[4]) {
  const f4 d0 = wino_fetch(a, b, ys[0], x, c, mean, rstd);
  const f4 d1 = wino_fetch(a, b, ys[1], x, c, mean, rstd);
  const f4 d2 = wino_fetch(a, b, ys[2], x, c, mean, rstd);
  const f4 d3 = wino_fetch(a, b, ys[3], x, c, mean, rstd);
  if (a.act_out && own) {
    st_f4(a.act_out + ((static_cast<size_t>(b) * a.H + ys[1]) * a.W + x) * a.act_C_total + a.act_c_off + c, d1);
    st_f4(a.act_out + ((static_cast<size_t>(b) * a.H + ys[2]) * a.W + x) * a.act_C_total + a.act_c_off + c, d2);
  }
  t[0] = f4_sub(d0, d2);
  t[1] = f4_add(d1, d2);
  t[2] = f4_sub(d2, d1);
  t[3] = f4_sub(d1, d3);
}

TSNET_HD void wino_input_body(const WinoInArgs& a, int block, int thread, int nthreads) {
  const int TH = a.H / 2, TW = a.W / 2;
  const int ty = block % TH, b = block / TH;
  const int cg = a.C / 4;
  const int runs = (TW + kWinoRun - 1) / kWinoRun;
  const int ys[4] = {wino_reflect(2 * ty - 1, a.H), 2 * ty, 2 * ty + 1, wino_reflect(2 * ty + 2, a.H)};
  for (int idx = thread; idx < runs * cg; idx += nthreads) {
    const int run = idx / cg, c = (idx - run * cg) * 4;
    f4 mean = f4{0.f, 0.f, 0.f, 0.f}, rstd = f4{1.f, 1.f, 1.f, 1.f};
    if (a.mean_rstd) {
      const float* mr = a.mean_rstd + (static_cast<size_t>(b) * a.C + c) * 2;
      const f4 t0 = ld_f4(mr), t1 = ld_f4(mr + 4);
      mean = f4{t0.x, t0.z, t1.x, t1.z};
      rstd = f4{t0.y, t0.w, t1.y, t1.w};
    }
    const int tx0 = run * kWinoRun, tx1 = tx0 + kWinoRun < TW ? tx0 + kWinoRun : TW;
    // columns x in [2 tx0, 2 tx1) are owned by this thread (their act_out is written here, once)
    f4 t[4][4];  // t[s][i]: row-transformed column s of the current tile
    wino_in_column(a, b, ys, wino_reflect(2 * tx0 - 1, a.W), false, c, mean, rstd, t[0]);
    wino_in_column(a, b, ys, 2 * tx0, true, c, mean, rstd, t[1]);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int tx = tx0; tx < tx1; ++tx) {
      wino_in_column(a, b, ys, 2 * tx + 1, true, c, mean, rstd, t[2]);
      const int x3 = 2 * tx + 2;
      wino_in_column(a, b, ys, wino_reflect(x3, a.W), x3 < 2 * tx1 && x3 < a.W, c, mean, rstd, t[3]);
      const size_t pstride = static_cast<size_t>(TH) * TW * a.Cp_total;  // one plane of one image
      const size_t d0 = wino_v_index(b, 0, TH * TW, a.Cp_total, ty * TW + tx, a.c_off + c);
      WinoPlaneWriter w{a.hi + d0, a.lo + d0, pstride, a.scale, a.fmt};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int i = 0; i < 4; ++i) {  // planes p = 4 i + j in order
        wino_write_next(w, f4_sub(t[0][i], t[2][i]));
        wino_write_next(w, f4_add(t[1][i], t[2][i]));
        wino_write_next(w, f4_sub(t[2][i], t[1][i]));
        wino_write_next(w, f4_sub(t[1][i], t[3][i]));
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int i = 0; i < 4; ++i) {  // the next tile starts two columns to the right
        t[0][i] = t[2][i];
        t[1][i] = t[3][i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pass I: output transform + bias (+ addend) + InstanceNorm partial statistics.
// M fp32 in the slab-major layout (wino_m_index); y fp32 [B, H, W, C]; stats [B * H*W/32, C, 2] = (sum, centred M2) of 32 pixels:
// a thread's kWinoRun = 8 tiles along one tile row are exactly one partial (2 rows x 16 pixels).
// ------------------------------------------------------------------------------------------------
struct WinoOutArgs {
  const float* m;
  const float* bias;    // [C] or null
  const float* addend;  // fp32 [addend_rows, C] or null: y[pixel] += addend[pixel % addend_rows]
  float* y;
  float* stats;         // or null
  int B, H, W, C;
  long long addend_rows;
};

TSNET_HD void wino_stat_merge(float v0, float v1, float v2, float v3, int t, float& S, float& M2) {
  const float s4 = (v0 + v1) + (v2 + v3);
  const float m4 = 0.25f * s4;
  const float e0 = v0 - m4, e1 = v1 - m4, e2 = v2 - m4, e3 = v3 - m4;
  const float q4 = (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
  if (t == 0) {
    S = s4;
    M2 = q4;
  } else {  // Chan merge of (n = 4 t, S, M2) with (4, s4, q4)
    const float n = 4.f * static_cast<float>(t);
    const float dl = m4 - S / n;
    M2 = M2 + q4 + dl * dl * (n * 4.f / (n + 4.f));
    S = S + s4;
  }
}

TSNET_HD void wino_output_body(const WinoOutArgs& a, int block, int thread, int nthreads) {
  const int TH = a.H / 2, TW = a.W / 2;
  const int ty = block % TH, b = block / TH;
  const int cg = a.C / 4;
  const int runs = TW / kWinoRun;
  const size_t ptile = static_cast<size_t>(a.B) * TH * TW;  // tiles per plane
  for (int idx = thread; idx < runs * cg; idx += nthreads) {
    const int run = idx / cg, c = (idx - run * cg) * 4;
    f4 bias = f4{0.f, 0.f, 0.f, 0.f};
    if (a.bias) bias = ld_f4(a.bias + c);
    float S[4] = {0.f, 0.f, 0.f, 0.f}, M2[4] = {0.f, 0.f, 0.f, 0.f};
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int t = 0; t < kWinoRun; ++t) {
      const int tx = run * kWinoRun + t;
      const float* mp = a.m + wino_m_index(a.B, a.C, TH * TW, 0, b, ty * TW + tx, c);
      f4 z[2][4];  // A^T M: rows a = 0, 1; columns j
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 4; ++j) {
        const f4 m0 = ld_f4(mp + (0 * 4 + j) * ptile * a.C);
        const f4 m1 = ld_f4(mp + (1 * 4 + j) * ptile * a.C);
        const f4 m2 = ld_f4(mp + (2 * 4 + j) * ptile * a.C);
        const f4 m3 = ld_f4(mp + (3 * 4 + j) * ptile * a.C);
        z[0][j] = f4_add(f4_add(m0, m1), m2);
        z[1][j] = f4_sub(f4_sub(m1, m2), m3);
      }
      f4 o[2][2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int r = 0; r < 2; ++r) {
        o[r][0] = f4_add(f4_add(f4_add(z[r][0], z[r][1]), z[r][2]), bias);
        o[r][1] = f4_add(f4_sub(f4_sub(z[r][1], z[r][2]), z[r][3]), bias);
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int re = 0; re < 4; ++re) {
        const int r = re >> 1, e = re & 1;
        const size_t pix = (static_cast<size_t>(b) * a.H + 2 * ty + r) * a.W + 2 * tx + e;
        if (a.addend)
          o[r][e] = f4_add(o[r][e], ld_f4(a.addend + (pix % static_cast<size_t>(a.addend_rows)) * a.C + c));
        st_f4(a.y + pix * a.C + c, o[r][e]);
      }
      if (a.stats) {
        wino_stat_merge(o[0][0].x, o[0][1].x, o[1][0].x, o[1][1].x, t, S[0], M2[0]);
        wino_stat_merge(o[0][0].y, o[0][1].y, o[1][0].y, o[1][1].y, t, S[1], M2[1]);
        wino_stat_merge(o[0][0].z, o[0][1].z, o[1][0].z, o[1][1].z, t, S[2], M2[2]);
        wino_stat_merge(o[0][0].w, o[0][1].w, o[1][0].w, o[1][1].w, t, S[3], M2[3]);
      }
    }
    if (a.stats) {
      const size_t part = (static_cast<size_t>(b) * TH + ty) * runs + run;  // = b * (H*W/32) + ty * runs + run
      float* sp = a.stats + (part * a.C + c) * 2;
      st_f4(sp, f4{S[0], M2[0], S[1], M2[1]});
      st_f4(sp + 4, f4{S[2], M2[2], S[3], M2[3]});
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bridge pass "I+T": output transform of layer k, InstanceNorm, [ReLU | + residual], input transform of layer k+1 in ONE
// pass over HBM -- M[k] is read once, V[k+1] is written once; the fp32 conv output, its statistics partials and the
// separate instnorm_reduce launch disappear.  One CTA = one image x a slab of kBridgeCS channels: the whole H x W x CS
// conv output lives in shared memory (128 KB at 32 x 32 x 32), so the per-(image, channel) statistics are
// CTA-local.  (A 16-channel variant with two CTAs per SM was measured slower: 204 vs 164 us on the 32-sample layer.)
// Phases (block-wide barriers between them; the host emulation runs each phase for every thread in turn):
//   A  y = A^T M A + bias (+ addend)                      -> shared memory
//   S  per-channel sum / sum of squares in fp64, fixed order -> mean, 1/sqrt(var + eps)   (S1 partials, S2 merge)
//   B  v = (y - mean) * rstd ; ReLU ; + residual ; act_out  -> shared memory (in place) and optional fp32 output
//   C  reflect pad + B^T d B + hi/lo split                  -> the 16 operand planes of the next plane GEMMs
// ------------------------------------------------------------------------------------------------
// Two variants (template parameters CS = channels per CTA, PS = shared-memory pixel stride in floats):
//   <32, 32>  one M slab per CTA: phase A streams 32 KB contiguous per plane; 128 KB image buffer, 512 threads, 1 CTA / SM
//   <16, 24>  half a slab per CTA (64-byte pieces): 96 KB image buffer (16 channels + 8 pad so that the 64-byte groups
//             the lanes of a warp touch alternate bank halves), 256 threads, 2 CTAs / SM whose read-heavy (A) and
//             write-heavy (C) phases overlap
constexpr int kBridgeCSDefault = kWinoMSlab;

struct WinoBridgeArgs {
  const float* m;         // fp32, slab-major (wino_m_index)
  const float* bias;      // [C] or null
  const float* addend;    // fp32 [addend_rows, C] or null
  const float* residual;  // fp32 [B, H, W, C] or null
  float* act_out;         // fp32 [B, H, W, act_C_total] window [act_c_off, +C) or null
  float* mean_rstd_out;   // [B, C, 2] or null (tests / diagnostics)
  uint16_t* hi;           // V layout (wino_v_index), channel window [c_off, +C) of Cp_total
  uint16_t* lo;
  int B, H, W, C, relu, Cp_total, c_off, fmt, act_C_total, act_c_off;
  long long addend_rows;
  float scale, eps;
  // optional: the activations also leave as the UN-NORMALISED operand rows of the correlation (tsnet_corr_tiles with
  // rnorm): hi / lo [B * H*W, C] at the row's sorted rank (corr_rank: [B, H*W] position -> rank) plus this slab's partial
  // sum of squares per pixel, corr_ssq [B][C / 32][H*W] (summed and inverted by tsnet_corr_norms)
  uint16_t* corr_hi;
  uint16_t* corr_lo;
  const uint16_t* corr_rank;
  float* corr_ssq;
  float corr_scale;
};

// shared-memory layout: float y[H * W * PS]; double part[nthreads][8] (sum, sum of squares of the thread's 4 channels);
// float mr[CS * 2]
template <int kBridgeCS, int kBridgePS>
TSNET_HD size_t wino_bridge_smem_bytes(int H, int W, int nthreads) {
  return static_cast<size_t>(H) * W * kBridgePS * 4 + static_cast<size_t>(nthreads) * 8 * 8 + kBridgeCS * 2 * 4;
}

template <int kBridgeCS, int kBridgePS>
TSNET_HD void wino_bridge_phase_a(const WinoBridgeArgs& a, int block, int thread, int nthreads, float* s_y,
                                  double* s_part) {
  const int slabs = a.C / kBridgeCS;
  const int b = block / slabs, slab = block - b * slabs;
  const int TH = a.H / 2, TW = a.W / 2, T = TH * TW;
  const size_t ptile = static_cast<size_t>(a.B) * T;
  constexpr int CQ = kBridgeCS / 4;
  // statistics: nthreads is a multiple of CQ, so every unit of a thread has the same 4 channels (cq = thread % CQ); the
  // thread's sum / sum of squares over its output pixels are kept in fp64 and merged in phase S in a fixed order
  double ds[4] = {0.0, 0.0, 0.0, 0.0}, dq[4] = {0.0, 0.0, 0.0, 0.0};
  for (int u = thread; u < T * CQ; u += nthreads) {
    const int tile = u / CQ, cq = u - tile * CQ;
    const int ty = tile / TW, tx = tile - ty * TW;
    const int c = slab * kBridgeCS + cq * 4;
    const float* mp = a.m + wino_m_index(a.B, a.C, T, 0, b, tile, c);
    f4 bias = f4{0.f, 0.f, 0.f, 0.f};
    if (a.bias) bias = ld_f4(a.bias + c);
    f4 z[2][4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) {
      const f4 m0 = ld_f4(mp + (0 * 4 + j) * ptile * a.C);
      const f4 m1 = ld_f4(mp + (1 * 4 + j) * ptile * a.C);
      const f4 m2 = ld_f4(mp + (2 * 4 + j) * ptile * a.C);
      const f4 m3 = ld_f4(mp + (3 * 4 + j) * ptile * a.C);
      z[0][j] = f4_add(f4_add(m0, m1), m2);
      z[1][j] = f4_sub(f4_sub(m1, m2), m3);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int re = 0; re < 4; ++re) {
      const int r = re >> 1, e = re & 1;
      f4 o = e == 0 ? f4_add(f4_add(f4_add(z[r][0], z[r][1]), z[r][2]), bias)
                    : f4_add(f4_sub(f4_sub(z[r][1], z[r][2]), z[r][3]), bias);
      const int pix = (2 * ty + r) * a.W + 2 * tx + e;
      if (a.addend) {
        const size_t gp = static_cast<size_t>(b) * a.H * a.W + pix;
        o = f4_add(o, ld_f4(a.addend + (gp % static_cast<size_t>(a.addend_rows)) * a.C + c));
      }
      st_f4(s_y + static_cast<size_t>(pix) * kBridgePS + cq * 4, o);
      const double ox = o.x, oy = o.y, oz = o.z, ow = o.w;
      ds[0] += ox; dq[0] += ox * ox;
      ds[1] += oy; dq[1] += oy * oy;
      ds[2] += oz; dq[2] += oz * oz;
      ds[3] += ow; dq[3] += ow * ow;
    }
  }
  double* sp = s_part + static_cast<size_t>(thread) * 8;
  for (int k = 0; k < 4; ++k) {
    sp[2 * k] = ds[k];
    sp[2 * k + 1] = dq[k];
  }
}

template <int kBridgeCS, int kBridgePS>
TSNET_HD void wino_bridge_phase_s2(const WinoBridgeArgs& a, int block, int thread, int nthreads, const double* s_part,
                                   float* s_mr) {
  if (thread >= kBridgeCS) return;
  constexpr int CQ = kBridgeCS / 4;
  const int cq = thread >> 2, comp = thread & 3;  // channel `thread` = component comp of the threads with t % CQ == cq
  double s = 0.0, q = 0.0;
  for (int t = cq; t < nthreads; t += CQ) {        // fixed order: deterministic
    s += s_part[static_cast<size_t>(t) * 8 + 2 * comp];
    q += s_part[static_cast<size_t>(t) * 8 + 2 * comp + 1];
  }
  const double n = static_cast<double>(a.H) * a.W;
  const double mean = s / n;
  double var = q / n - mean * mean;  // biased, as nn.InstanceNorm2d
  var = var > 0.0 ? var : 0.0;
  const float mf = static_cast<float>(mean);
#if defined(__CUDA_ARCH__)
  const float rf = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
#else
  const float rf = static_cast<float>(1.0 / __builtin_sqrt(var + static_cast<double>(a.eps)));
#endif
  s_mr[thread * 2 + 0] = mf;
  s_mr[thread * 2 + 1] = rf;
  if (a.mean_rstd_out) {
    const int slabs = a.C / kBridgeCS;
    const int b = block / slabs, slab = block - b * slabs;
    float* o = a.mean_rstd_out + (static_cast<size_t>(b) * a.C + slab * kBridgeCS + thread) * 2;
    o[0] = mf;
    o[1] = rf;
  }
}

template <int kBridgeCS, int kBridgePS>
TSNET_HD void wino_bridge_phase_b(const WinoBridgeArgs& a, int block, int thread, int nthreads, float* s_y,
                                  const float* s_mr) {
  const int slabs = a.C / kBridgeCS;
  const int b = block / slabs, slab = block - b * slabs;
  const int HW = a.H * a.W;
  constexpr int CQ = kBridgeCS / 4;
  const int total = HW * CQ;
  constexpr int kPB = 4;
  // kPB units per trip: the (dependent-latency) residual loads of a trip are issued together
  for (int u0 = thread; u0 < total; u0 += kPB * nthreads) {
    f4 res[kPB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < kPB; ++k) {
      const int u = u0 + k * nthreads;
      res[k] = f4{0.f, 0.f, 0.f, 0.f};
      if (a.residual && u < total) {
        const int pix = u / CQ, cq = u - pix * CQ;
        res[k] = ld_f4(a.residual + (static_cast<size_t>(b) * HW + pix) * a.C + slab * kBridgeCS + cq * 4);
      }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < kPB; ++k) {
      const int u = u0 + k * nthreads;
      if (u >= total) continue;
      const int pix = u / CQ, cq = u - pix * CQ;
      float* sp = s_y + static_cast<size_t>(pix) * kBridgePS + cq * 4;
      f4 v = ld_f4(sp);
      const f4 m01 = ld_f4(s_mr + cq * 8), m23 = ld_f4(s_mr + cq * 8 + 4);  // (mean, rstd) x 4 channels
      v.x = (v.x - m01.x) * m01.y; v.y = (v.y - m01.z) * m01.w;
      v.z = (v.z - m23.x) * m23.y; v.w = (v.w - m23.z) * m23.w;
      if (a.relu) {
        v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
        v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
      }
      if (a.residual) v = f4_add(v, res[k]);
      st_f4(sp, v);
      if (a.act_out)
        st_f4(a.act_out + (static_cast<size_t>(b) * HW + pix) * a.act_C_total + a.act_c_off + slab * kBridgeCS + cq * 4, v);
      if (a.corr_hi) {
        const size_t row = static_cast<size_t>(b) * HW + a.corr_rank[static_cast<size_t>(b) * HW + pix];
        WinoPlaneWriter w{a.corr_hi + row * a.C + slab * kBridgeCS + cq * 4, a.corr_lo + row * a.C + slab * kBridgeCS + cq * 4,
                          0, a.corr_scale, a.fmt};
        wino_write_next(w, v);
      }
    }
  }
}

// phase N (after B; reads only): this slab's sum of squares of every pixel, channel order rotated by the pixel index so
// that the lanes of a warp (consecutive pixels, 128 bytes apart) hit different banks
template <int kBridgeCS, int kBridgePS>
TSNET_HD void wino_bridge_phase_n(const WinoBridgeArgs& a, int block, int thread, int nthreads, const float* s_y) {
  if (!a.corr_ssq) return;
  const int slabs = a.C / kBridgeCS;
  const int b = block / slabs, slab = block - b * slabs;
  const int HW = a.H * a.W;
  for (int pix = thread; pix < HW; pix += nthreads) {
    float ss = 0.f;
    for (int k = 0; k < kBridgeCS; ++k) {
      const float v = s_y[static_cast<size_t>(pix) * kBridgePS + ((k + pix) & (kBridgeCS - 1))];
      ss += v * v;
    }
    a.corr_ssq[(static_cast<size_t>(b) * slabs + slab) * HW + pix] = ss;
  }
}

template <int kBridgeCS, int kBridgePS>
// `fold` = phase B was skipped (no residual, act_out or correlation outputs): normalisation (+ ReLU) happens here, on read
TSNET_HD void wino_bridge_phase_c(const WinoBridgeArgs& a, int block, int thread, int nthreads, const float* s_y,
                                  const float* s_mr, bool fold) {
  const int slabs = a.C / kBridgeCS;
  const int b = block / slabs, slab = block - b * slabs;
  const int TH = a.H / 2, TW = a.W / 2, T = TH * TW;
  constexpr int CQ = kBridgeCS / 4;
  const size_t pstride = static_cast<size_t>(T) * a.Cp_total;
  for (int u = thread; u < T * CQ; u += nthreads) {
    const int tile = u / CQ, cq = u - tile * CQ;
    const int ty = tile / TW, tx = tile - ty * TW;
    const int ys[4] = {wino_reflect(2 * ty - 1, a.H), 2 * ty, 2 * ty + 1, wino_reflect(2 * ty + 2, a.H)};
    const int xs[4] = {wino_reflect(2 * tx - 1, a.W), 2 * tx, 2 * tx + 1, wino_reflect(2 * tx + 2, a.W)};
    f4 mean = f4{0.f, 0.f, 0.f, 0.f}, rstd = f4{1.f, 1.f, 1.f, 1.f};
    if (fold) {
      const f4 m01 = ld_f4(s_mr + cq * 8), m23 = ld_f4(s_mr + cq * 8 + 4);
      mean = f4{m01.x, m01.z, m23.x, m23.z};
      rstd = f4{m01.y, m01.w, m23.y, m23.w};
    }
    f4 t[4][4];  // t[s][i] = (B^T d)[i][s]
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int sx = 0; sx < 4; ++sx) {
      const float* col = s_y + static_cast<size_t>(xs[sx]) * kBridgePS + cq * 4;
      f4 d[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int r = 0; r < 4; ++r) {
        f4 v = ld_f4(col + static_cast<size_t>(ys[r]) * a.W * kBridgePS);
        if (fold) {  // same expression as phase B
          v.x = (v.x - mean.x) * rstd.x; v.y = (v.y - mean.y) * rstd.y;
          v.z = (v.z - mean.z) * rstd.z; v.w = (v.w - mean.w) * rstd.w;
          if (a.relu) {
            v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
            v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
          }
        }
        d[r] = v;
      }
      const f4 d0 = d[0], d1 = d[1], d2 = d[2], d3 = d[3];
      t[sx][0] = f4_sub(d0, d2);
      t[sx][1] = f4_add(d1, d2);
      t[sx][2] = f4_sub(d2, d1);
      t[sx][3] = f4_sub(d1, d3);
    }
    const size_t d0 = wino_v_index(b, 0, T, a.Cp_total, tile, a.c_off + slab * kBridgeCS + cq * 4);
    WinoPlaneWriter w{a.hi + d0, a.lo + d0, pstride, a.scale, a.fmt};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {  // planes p = 4 i + j in order
      wino_write_next(w, f4_sub(t[0][i], t[2][i]));
      wino_write_next(w, f4_add(t[1][i], t[2][i]));
      wino_write_next(w, f4_sub(t[2][i], t[1][i]));
      wino_write_next(w, f4_sub(t[1][i], t[3][i]));
    }
  }
}

}  // namespace tsnet
