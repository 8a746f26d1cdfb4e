// Fused mask-aware correlation -> softmax(100 x) -> expected source coordinate -> bilinear warp -> mean over sources.
// (model/TSNet.py:319-366, :392 of the reference.)  The hw x hw similarity matrix lives only in TMEM.
//
// Work item = (sample b, tile of 128 target positions).  For every source i and every chunk of 256 source
// positions the tensor cores compute S = T_hat[128 x C] . S_hat_i[256 x C]^T (3-term hi/lo split).  As in the
// conv GEMM, tcgen05's truncating fp32 accumulation is kept short: every 2 K-blocks (24 MMAs) the partial sum in
// one of two TMEM buffers is promoted to fp32 REGISTER accumulators of the eight softmax warps (one thread owns
// one target row x 128 of the 256 columns).  When a chunk is complete the same threads run an online softmax with a
// 2-channel "V" (the source coordinates) over it while the tensor cores work on the next chunk; the two column
// halves of a row are merged through shared memory at the end of each source.
// (Promoting after EVERY K-block was measured: no accuracy gain -- 1.23e-5 vs 1.25e-5 grid error against fp64 -- while
// the TMEM -> register traffic, 128 KB per promotion at ~64 B/clk, then exceeds the 1536 clk of MMA work it must
// hide behind: tensor pipe 37 % active instead of ~55 %.)
// After the last source the eight warps gather the 4 bilinear taps per (row, source) from the UN-normalised fp32
// source features and write the source mean.
//
// warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-11 = softmax + gather.
#include "sm100_prims.cuh"
#include "host_util.h"
#include "../../include/tsnet_b200.h"
#include <math.h>

namespace tsnet {

constexpr int kCorrM = 128;       // target rows per work item
constexpr int kCorrN = 256;       // source columns per chunk
constexpr int kCorrNC = 128;      // columns owned by one thread
constexpr int kCorrK = 64;        // K block (one 128 B swizzle row)
constexpr int kCorrChunkKb = 2;   // K-blocks (24 MMAs) accumulated in TMEM before promotion to registers
constexpr int kCorrThreads = 384;
constexpr int kCorrEpiThreads = 256;
constexpr int kCorrMaxSrc = 12;
constexpr int kCorrMaxHW = 1024;
constexpr int kCorrABytes = kCorrM * kCorrK * 2;                    // 16 KB
constexpr int kCorrBBytes = kCorrN * kCorrK * 2;                    // 32 KB
constexpr int kCorrStageBytes = 2 * kCorrABytes + 2 * kCorrBBytes;  // 96 KB
constexpr int kCorrStages = 2;

struct alignas(64) CorrArgs {
  CUtensorMap t_hi, t_lo, s_hi, s_lo;  // [B*hw, C] box {64, 128} and [n_src*B*hw, C] box {64, 256}
  const float* src_fea[kCorrMaxSrc];
  const void* src_bbox[kCorrMaxSrc];
  const void* tar_bbox;
  const float* coord_table;  // h values (y) then w values (x)
  float* out_mean;
  float* out_grids;
  int B, n_src, C, h, w, hw, tiles_per_img, num_items;
  int bbox_h, bbox_w, bbox_dtype;
  int split, fmt;
  float logit_scale;  // temperature / operand_scale
};

struct CorrSmemTail {
  uint64_t full_bar[kCorrStages], empty_bar[kCorrStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad[15];
  alignas(16) float mask[kCorrMaxHW];  // nearest-down-sampled source mask of the current source
  float cx[kCorrMaxHW];              // x coordinate of source position s
  float cy[kCorrMaxHW];              // y coordinate of source position s
  float2 grid[kCorrMaxSrc][kCorrM];  // expected coordinate per (source, row)
  float4 merge[kCorrM];              // softmax state of the upper column half
};

__device__ __forceinline__ float read_mask(const void* bbox, int dtype, int b, int bh, int bw, int h, int w, int pos) {
  // F.interpolate(mode='nearest'): src = min(floor(dst * (in / out)), in - 1), float scale (ATen)
  const int y = pos / w, x = pos - y * w;
  const int sy = min(static_cast<int>(floorf(y * (static_cast<float>(bh) / h))), bh - 1);
  const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(bw) / w))), bw - 1);
  const size_t off = (static_cast<size_t>(b) * bh + sy) * bw + sx;
  return dtype == 0 ? static_cast<float>(static_cast<const uint8_t*>(bbox)[off])
                    : static_cast<const float*>(bbox)[off];
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 2^x for x <= 0 (softmax weights): MUFU.EX2, results below 2^-126 flush to zero (they are < 1e-38 of the row maximum)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kCorrThreads, 1) corr_warp_kernel(const __grid_constant__ CorrArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  CorrSmemTail& tl = *reinterpret_cast<CorrSmemTail*>(smem + kCorrStages * kCorrStageBytes);

  const int warp = threadIdx.x >> 5;
  const int num_kb = args.C / kCorrK;
  const int chunks = args.hw / kCorrN;

  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&args.t_hi);
    tma_prefetch_desc(&args.s_hi);
    if (args.split) {
      tma_prefetch_desc(&args.t_lo);
      tma_prefetch_desc(&args.s_lo);
    }
  }
  if (warp == 1 && lane_id() == 0) {
    for (int s = 0; s < kCorrStages; ++s) {
      mbar_init(&tl.full_bar[s], 1);
      mbar_init(&tl.empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tl.tmem_full[a], 1);
      mbar_init(&tl.tmem_empty[a], 8);  // one arrive per softmax warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tl.tmem_base, 2 * kCorrN);  // two partial-sum buffers
  for (int s = threadIdx.x; s < args.hw; s += blockDim.x) {
    const int sy = s / args.w, sx = s - sy * args.w;
    tl.cx[s] = args.coord_table[args.h + sx];
    tl.cy[s] = args.coord_table[sy];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tl.tmem_base;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane_id() == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t stage_tx = args.split ? kCorrStageBytes : (kCorrABytes + kCorrBBytes);
        for (int item = blockIdx.x; item < args.num_items; item += gridDim.x) {
          const int b = item / args.tiles_per_img;
          const int mt = item - b * args.tiles_per_img;
          const int trow = b * args.hw + mt * kCorrM;
          for (int i = 0; i < args.n_src; ++i) {
            for (int ch = 0; ch < chunks; ++ch) {
              const int srow = (i * args.B + b) * args.hw + ch * kCorrN;
              for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&tl.empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + stage * kCorrStageBytes;
                mbar_arrive_expect_tx(&tl.full_bar[stage], stage_tx);
                tma_load_2d(st, &args.t_hi, &tl.full_bar[stage], kb * kCorrK, trow);
                tma_load_2d(st + 2 * kCorrABytes, &args.s_hi, &tl.full_bar[stage], kb * kCorrK, srow);
                if (args.split) {
                  tma_load_2d(st + kCorrABytes, &args.t_lo, &tl.full_bar[stage], kb * kCorrK, trow);
                  tma_load_2d(st + 2 * kCorrABytes + kCorrBBytes, &args.s_lo, &tl.full_bar[stage], kb * kCorrK, srow);
                }
                if (++stage == kCorrStages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane_id() == 0) {
        const uint32_t idesc = make_idesc_f16(kCorrM, kCorrN, args.fmt);
        int stage = 0;
        uint32_t phase = 0;
        int cc = 0;  // partial-accumulator counter -> TMEM buffer + phase
        for (int item = blockIdx.x; item < args.num_items; item += gridDim.x) {
          for (int ic = 0; ic < args.n_src * chunks; ++ic) {
            for (int kb0 = 0; kb0 < num_kb; kb0 += kCorrChunkKb, ++cc) {
              const int buf = cc & 1;
              const uint32_t buf_phase = (cc >> 1) & 1;
              mbar_wait(&tl.tmem_empty[buf], buf_phase ^ 1);
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + buf * kCorrN;
              const int kb1 = min(num_kb, kb0 + kCorrChunkKb);
              for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&tl.full_bar[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + stage * kCorrStageBytes);
                const uint64_t a_hi = make_desc_kmajor_sw128(st);
                const uint64_t a_lo = make_desc_kmajor_sw128(st + kCorrABytes);
                const uint64_t b_hi = make_desc_kmajor_sw128(st + 2 * kCorrABytes);
                const uint64_t b_lo = make_desc_kmajor_sw128(st + 2 * kCorrABytes + kCorrBBytes);
#pragma unroll
                for (int k = 0; k < kCorrK / 16; ++k) {
                  const uint32_t off = k * 32;
                  umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_hi, off), idesc, ((kb - kb0) | k) != 0);
                  if (args.split) {
                    umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_lo, off), idesc, 1);
                    umma_f16(d_tmem, desc_advance_k(a_lo, off), desc_advance_k(b_hi, off), idesc, 1);
                  }
                }
                umma_commit(&tl.empty_bar[stage]);
                if (++stage == kCorrStages) { stage = 0; phase ^= 1; }
              }
              umma_commit(&tl.tmem_full[buf]);
            }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax + gather =====================
    const int q = warp & 3;            // TMEM lane quarter
    const int half = (warp - 4) >> 2;  // column half of every 256-column chunk
    const int row = q * 32 + lane_id();
    const int et = threadIdx.x - 128;  // 0..255 among the softmax threads
    constexpr float kLog2e = 1.4426950408889634f;
    int it = 0;
    for (int item = blockIdx.x; item < args.num_items; item += gridDim.x) {
      const int b = item / args.tiles_per_img;
      const int mt = item - b * args.tiles_per_img;
      const int tpos = mt * kCorrM + row;  // target position inside the image
      const float m_t = read_mask(args.tar_bbox, args.bbox_dtype, b, args.bbox_h, args.bbox_w, args.h, args.w, tpos);
      // (T*mt).(S*ms) + (T*(1-mt)).(S*(1-ms)) == (T.S) * (mt*ms + (1-mt)*(1-ms)) = (T.S) * (wa*ms + wb);
      // exact for binary masks (mismatched pairs get logit 0, not -inf, as in the reference)
      const float wa = 2.f * m_t - 1.f, wb = 1.f - m_t;
      for (int i = 0; i < args.n_src; ++i) {
        epi_bar_sync();  // previous source: mask and merge buffer no longer in use
        for (int p = et; p < args.hw; p += kCorrEpiThreads)
          tl.mask[p] = read_mask(args.src_bbox[i], args.bbox_dtype, b, args.bbox_h, args.bbox_w, args.h, args.w, p);
        epi_bar_sync();
        float run_max = -INFINITY, run_sum = 0.f, gx = 0.f, gy = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
          // ---- promote the partial sums of this chunk into registers
          float acc[kCorrNC];
#pragma unroll
          for (int j = 0; j < kCorrNC; ++j) acc[j] = 0.f;
          for (int kb0 = 0; kb0 < num_kb; kb0 += kCorrChunkKb, ++it) {
            const int buf = it & 1;
            const uint32_t buf_phase = (it >> 1) & 1;
            mbar_wait(&tl.tmem_full[buf], buf_phase);
            tc_fence_after();
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kCorrN + half * kCorrNC;
#pragma unroll
            for (int c0 = 0; c0 < kCorrNC; c0 += 32) {
              float v[32];
              tmem_ld_32x32(t0 + c0, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[c0 + j] += v[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(&tl.tmem_empty[buf]);
          }
          // ---- online softmax with the source coordinates as V (tables read as float4: 3 LDS.128 per 4 columns)
#pragma unroll
          for (int c0 = 0; c0 < kCorrNC; c0 += 32) {
            const int s0 = ch * kCorrN + half * kCorrNC + c0;
            float gmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 mk = *reinterpret_cast<const float4*>(&tl.mask[s0 + j]);
              const float mv[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float wgt = fmaf(wa, mv[u], wb);
                acc[c0 + j + u] = (acc[c0 + j + u] * args.logit_scale) * wgt;
                gmax = fmaxf(gmax, acc[c0 + j + u]);
              }
            }
            const float new_max = fmaxf(run_max, gmax);
            const float corr = ex2_approx((run_max - new_max) * kLog2e);  // 2^(-inf) = 0 on the first group
            run_sum *= corr; gx *= corr; gy *= corr;
            run_max = new_max;
            const float mb = new_max * kLog2e;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 vx = *reinterpret_cast<const float4*>(&tl.cx[s0 + j]);
              const float4 vy = *reinterpret_cast<const float4*>(&tl.cy[s0 + j]);
              const float xs[4] = {vx.x, vx.y, vx.z, vx.w}, ys[4] = {vy.x, vy.y, vy.z, vy.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float p = ex2_approx(fmaf(acc[c0 + j + u], kLog2e, -mb));  // argument <= 0
                run_sum += p;
                gx = fmaf(p, xs[u], gx);
                gy = fmaf(p, ys[u], gy);
              }
            }
          }
        }
        // ---- merge the two column halves of the row
        if (half == 1) tl.merge[row] = make_float4(run_max, run_sum, gx, gy);
        epi_bar_sync();
        if (half == 0) {
          const float4 o = tl.merge[row];
          const float m = fmaxf(run_max, o.x);
          const float c0 = ex2_approx((run_max - m) * kLog2e), c1 = ex2_approx((o.x - m) * kLog2e);
          const float l = run_sum * c0 + o.y * c1;
          const float2 g = make_float2((gx * c0 + o.z * c1) / l, (gy * c0 + o.w * c1) / l);
          tl.grid[i][row] = g;
          if (args.out_grids)
            *reinterpret_cast<float2*>(args.out_grids + ((static_cast<size_t>(i) * args.B + b) * args.hw + tpos) * 2) = g;
        }
      }
      epi_bar_sync();  // all grids of this item visible
      // ---- gather: softmax warp e handles rows e*16 .. e*16+15; lane owns channels {128 k + 4 lane .. +3} ----
      const int nk = args.C / 128;
      const float n_srcf = static_cast<float>(args.n_src);
      const int e = warp - 4;
      for (int r = 0; r < (args.out_mean ? 16 : 0); ++r) {
        const int grow = e * 16 + r;
        const int pos = mt * kCorrM + grow;
        float4 accv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) accv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < args.n_src; ++i) {
          const float2 g = tl.grid[i][grow];
          // F.grid_sample(bilinear, zeros, align_corners=False): ix = ((x + 1) * W - 1) / 2
          const float ix = ((g.x + 1.f) * args.w - 1.f) * 0.5f;
          const float iy = ((g.y + 1.f) * args.h - 1.f) * 0.5f;
          const float fx = floorf(ix), fy = floorf(iy);
          const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
          const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
          float wts[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};  // nw, ne, sw, se
          const float* base = args.src_fea[i] + static_cast<size_t>(b) * args.hw * args.C + lane_id() * 4;
          const float* tp[4];
#pragma unroll
          for (int tap = 0; tap < 4; ++tap) {
            const int xx = x0 + (tap & 1), yy = y0 + (tap >> 1);
            const bool in = xx >= 0 && xx < args.w && yy >= 0 && yy < args.h;
            if (!in) wts[tap] = 0.f;  // zeros padding: out-of-range taps contribute nothing
            const int xc = min(max(xx, 0), args.w - 1), yc = min(max(yy, 0), args.h - 1);
            tp[tap] = base + static_cast<size_t>(yc * args.w + xc) * args.C;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < nk) {
              const float4 f0 = __ldg(reinterpret_cast<const float4*>(tp[0] + k * 128));
              const float4 f1 = __ldg(reinterpret_cast<const float4*>(tp[1] + k * 128));
              const float4 f2 = __ldg(reinterpret_cast<const float4*>(tp[2] + k * 128));
              const float4 f3 = __ldg(reinterpret_cast<const float4*>(tp[3] + k * 128));
              float4 res;
              res.x = fmaf(f3.x, wts[3], fmaf(f2.x, wts[2], fmaf(f1.x, wts[1], f0.x * wts[0])));
              res.y = fmaf(f3.y, wts[3], fmaf(f2.y, wts[2], fmaf(f1.y, wts[1], f0.y * wts[0])));
              res.z = fmaf(f3.z, wts[3], fmaf(f2.z, wts[2], fmaf(f1.z, wts[1], f0.z * wts[0])));
              res.w = fmaf(f3.w, wts[3], fmaf(f2.w, wts[2], fmaf(f1.w, wts[1], f0.w * wts[0])));
              accv[k].x += res.x; accv[k].y += res.y; accv[k].z += res.z; accv[k].w += res.w;
            }
          }
        }
        float* o = args.out_mean + (static_cast<size_t>(b) * args.hw + pos) * args.C + lane_id() * 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < nk)
            *reinterpret_cast<float4*>(o + k * 128) =
                make_float4(accv[k].x / n_srcf, accv[k].y / n_srcf, accv[k].z / n_srcf, accv[k].w / n_srcf);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kCorrN);
  }
}

constexpr int kCorrSmemBytes = kCorrStages * kCorrStageBytes + 1024 + static_cast<int>(sizeof(CorrSmemTail));
static_assert(kCorrSmemBytes <= 227 * 1024, "corr_warp shared memory budget");

// ------------------------------------------------------------------------------------------------
// K2: bilinear grid_sample of the n source feature maps at the warp grids + mean over sources, written as the
// operand (hi / lo tap source) of the decoder's map_conv -- "grid_sample fused with the following conv's load"
// (model/TSNet.py:366, :392, :163).  One warp = one target position; lane owns channels {128 k + 4 lane .. +3}.
// Runs with the whole L1 available (K1 leaves ~4 KB), which is what the 4-tap gather wants.
// ------------------------------------------------------------------------------------------------
struct WarpTapsArgs {
  const float* src_fea[kCorrMaxSrc];
  const float* grids;   // [n, B, hw, 2]
  float* out_mean;      // fp32 [B, hw, C] or null
  uint16_t* hi;         // [B, hw, Cp_total] or null
  uint16_t* lo;
  int B, n_src, C, h, w, Cp_total, c_off, fmt;
  float scale;
};

__global__ void __launch_bounds__(256) warp_mean_taps_kernel(const WarpTapsArgs a) {
  const int hw = a.h * a.w;
  const size_t gpos = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gpos >= static_cast<size_t>(a.B) * hw) return;
  const int b = static_cast<int>(gpos / hw);
  const int lane = threadIdx.x & 31;
  const int nk = a.C / 128;
  float4 accv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) accv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < a.n_src; ++i) {
    const float2 g = *reinterpret_cast<const float2*>(a.grids + (static_cast<size_t>(i) * a.B * hw + gpos) * 2);
    // F.grid_sample(bilinear, zeros, align_corners=False): ix = ((x + 1) * W - 1) / 2
    const float ix = ((g.x + 1.f) * a.w - 1.f) * 0.5f;
    const float iy = ((g.y + 1.f) * a.h - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
    float wts[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};  // nw, ne, sw, se
    const float* base = a.src_fea[i] + static_cast<size_t>(b) * hw * a.C + lane * 4;
    const float* tp[4];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int xx = x0 + (tap & 1), yy = y0 + (tap >> 1);
      if (!(xx >= 0 && xx < a.w && yy >= 0 && yy < a.h)) wts[tap] = 0.f;  // zeros padding
      const int xc = min(max(xx, 0), a.w - 1), yc = min(max(yy, 0), a.h - 1);
      tp[tap] = base + static_cast<size_t>(yc * a.w + xc) * a.C;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < nk) {
        const float4 f0 = __ldg(reinterpret_cast<const float4*>(tp[0] + k * 128));
        const float4 f1 = __ldg(reinterpret_cast<const float4*>(tp[1] + k * 128));
        const float4 f2 = __ldg(reinterpret_cast<const float4*>(tp[2] + k * 128));
        const float4 f3 = __ldg(reinterpret_cast<const float4*>(tp[3] + k * 128));
        accv[k].x += fmaf(f3.x, wts[3], fmaf(f2.x, wts[2], fmaf(f1.x, wts[1], f0.x * wts[0])));
        accv[k].y += fmaf(f3.y, wts[3], fmaf(f2.y, wts[2], fmaf(f1.y, wts[1], f0.y * wts[0])));
        accv[k].z += fmaf(f3.z, wts[3], fmaf(f2.z, wts[2], fmaf(f1.z, wts[1], f0.z * wts[0])));
        accv[k].w += fmaf(f3.w, wts[3], fmaf(f2.w, wts[2], fmaf(f1.w, wts[1], f0.w * wts[0])));
      }
    }
  }
  const float nf = static_cast<float>(a.n_src);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < nk) {
      const float v[4] = {accv[k].x / nf, accv[k].y / nf, accv[k].z / nf, accv[k].w / nf};
      const int c = k * 128 + lane * 4;
      if (a.out_mean) *reinterpret_cast<float4*>(a.out_mean + gpos * a.C + c) = make_float4(v[0], v[1], v[2], v[3]);
      if (a.hi) {
        uint16_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split16(v[j] * a.scale, a.fmt, h[j], l[j]);
        const size_t d = gpos * a.Cp_total + a.c_off + c;
        *reinterpret_cast<uint2*>(a.hi + d) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
        *reinterpret_cast<uint2*>(a.lo + d) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
      }
    }
  }
}

}  // namespace tsnet

using namespace tsnet;

extern "C" int tsnet_warp_mean_taps(const float* const* src_fea, int n_src, const float* grids, int B, int h, int w,
                                    int C, float* out_mean, uint16_t* taps_hi, uint16_t* taps_lo, int Cp_total,
                                    int c_off, int fmt, float scale, void* stream) {
  TSNET_ARG_CHECK(src_fea && grids && (out_mean || taps_hi), "warp_mean_taps: null argument");
  TSNET_ARG_CHECK(n_src >= 1 && n_src <= kCorrMaxSrc, "warp_mean_taps: n_src %d (max %d)", n_src, kCorrMaxSrc);
  TSNET_ARG_CHECK(C % 128 == 0 && C <= 1024, "warp_mean_taps: C %d must be a multiple of 128, <= 1024", C);
  TSNET_ARG_CHECK((taps_hi == nullptr) == (taps_lo == nullptr), "warp_mean_taps: hi/lo must both be given or both NULL");
  TSNET_ARG_CHECK(!taps_hi || (Cp_total % 4 == 0 && c_off % 4 == 0 && c_off + C <= Cp_total),
                  "warp_mean_taps: channel window does not fit");
  WarpTapsArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < n_src; ++i) {
    TSNET_ARG_CHECK(src_fea[i], "warp_mean_taps: null source %d", i);
    a.src_fea[i] = src_fea[i];
  }
  a.grids = grids; a.out_mean = out_mean; a.hi = taps_hi; a.lo = taps_lo;
  a.B = B; a.n_src = n_src; a.C = C; a.h = h; a.w = w; a.Cp_total = Cp_total; a.c_off = c_off; a.fmt = fmt;
  a.scale = scale == 0.f ? 1.f : scale;
  const size_t rows = static_cast<size_t>(B) * h * w;
  warp_mean_taps_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  TSNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" size_t tsnet_corr_warp_workspace_bytes(const tsnet_corr_desc*) { return 0; }

extern "C" int tsnet_corr_warp_fwd(const tsnet_corr_desc* d, const uint16_t* tar_hi, const uint16_t* tar_lo,
                                   const uint16_t* src_hi, const uint16_t* src_lo, const float* const* src_fea,
                                   const void* tar_bbox, const void* const* src_bbox, const float* coord_table,
                                   float* out_mean, float* out_grids, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  TSNET_ARG_CHECK(d && tar_hi && src_hi && tar_bbox && src_bbox && coord_table && (out_mean || out_grids),
                  "corr_warp: null argument");
  TSNET_ARG_CHECK(!out_mean || src_fea, "corr_warp: out_mean needs the source features");
  TSNET_ARG_CHECK(!d->split || (tar_lo && src_lo), "corr_warp: split mode needs the lo operands");
  TSNET_ARG_CHECK(d->n_src >= 1 && d->n_src <= kCorrMaxSrc, "corr_warp: n_src %d (max %d)", d->n_src, kCorrMaxSrc);
  const int hw = d->h * d->w;
  TSNET_ARG_CHECK(hw % kCorrN == 0 && hw <= kCorrMaxHW, "corr_warp: h*w = %d must be a multiple of 256, <= %d", hw,
                  kCorrMaxHW);
  TSNET_ARG_CHECK(d->C % 128 == 0 && d->C <= 1024, "corr_warp: C %d must be a multiple of 128, <= 1024", d->C);
  TSNET_ARG_CHECK(d->bbox_dtype == 0 || d->bbox_dtype == 1, "corr_warp: bbox_dtype %d", d->bbox_dtype);

  CorrArgs a;
  memset(&a, 0, sizeof(a));
  {
    const uint64_t dims[2] = {(uint64_t)d->C, (uint64_t)d->B * hw};
    const uint64_t str[1] = {(uint64_t)d->C * 2};
    const uint32_t box[2] = {64, kCorrM};
    int r = encode_tmap_u16_sw128(&a.t_hi, tar_hi, 2, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.t_lo, tar_lo, 2, dims, str, box))) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d->C, (uint64_t)d->n_src * d->B * hw};
    const uint64_t str[1] = {(uint64_t)d->C * 2};
    const uint32_t box[2] = {64, kCorrN};
    int r = encode_tmap_u16_sw128(&a.s_hi, src_hi, 2, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.s_lo, src_lo, 2, dims, str, box))) return r;
  }
  for (int i = 0; i < d->n_src; ++i) {
    TSNET_ARG_CHECK((!out_mean || src_fea[i]) && src_bbox[i], "corr_warp: null source %d", i);
    a.src_fea[i] = out_mean ? src_fea[i] : nullptr;
    a.src_bbox[i] = src_bbox[i];
  }
  a.tar_bbox = tar_bbox;
  a.coord_table = coord_table;
  a.out_mean = out_mean;
  a.out_grids = out_grids;
  a.B = d->B; a.n_src = d->n_src; a.C = d->C; a.h = d->h; a.w = d->w; a.hw = hw;
  a.tiles_per_img = hw / kCorrM;
  a.num_items = d->B * a.tiles_per_img;
  a.bbox_h = d->bbox_h; a.bbox_w = d->bbox_w; a.bbox_dtype = d->bbox_dtype;
  a.split = d->split; a.fmt = d->fmt;
  // operand_scale is a power of two, so folding it into the temperature is exact up to one rounding of the product
  a.logit_scale = d->temperature / (d->operand_scale == 0.f ? 1.f : d->operand_scale);

  static bool attr_set = false;
  if (!attr_set) {
    TSNET_CUDA_CHECK(cudaFuncSetAttribute(corr_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kCorrSmemBytes));
    attr_set = true;
  }
  const int grid = a.num_items < num_sms() ? a.num_items : num_sms();
  corr_warp_kernel<<<grid, kCorrThreads, kCorrSmemBytes, static_cast<cudaStream_t>(stream)>>>(a);
  TSNET_CUDA_CHECK(cudaGetLastError());
  return 0;
}
