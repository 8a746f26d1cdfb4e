// Mask-aware correlation -> softmax(100 x) -> expected source coordinate -> bilinear warp -> mean over sources
// (model/TSNet.py:319-366, :392 of the reference).  The hw x hw similarity matrix lives only in TMEM.
//
// Three kernels, all enqueued on the caller's stream:
//
//  P  corr_prepare_kernel (one block per mask: class sort; the last block of a sample builds its work list)
//     masks -> class-sorted order + tile classes + work list
//     The reference's similarity is (T.S) * (mt*ms + (1-mt)(1-ms)): for the {0,1} bbox masks every pair whose classes
//     differ has logit EXACTLY 0.  Positions of every map are therefore stably sorted by mask class (1, soft, 0); a
//     128-row target tile and a 256-column source chunk that are class-pure with different classes need no tensor
//     work at all: their softmax contribution is the closed form (max 0, weight count, sum of coordinates) and is
//     written here.  Everything else goes on the work list.  (sort = 0 keeps the raster order.)
//  K1 corr_tile2_kernel    (persistent, 74 CTA PAIRS, default)  work item = (sample, 256 target rows, source, 256 source
//     columns): tcgen05.mma.cta_group::2, M = 256 across the pair -- see the comment above the kernel.
//     corr_tile_kernel     (persistent, <= 148 CTAs, tsnet_corr_desc.one_cta)  work item = (sample, 128 target rows, source,
//     256 source columns).  Common to both:
//     S = T_hat[128 x C] . S_hat[256 x C]^T on tcgen05 (3-term hi/lo split, fp32 accumulate in TMEM, two 128 x 256
//     accumulators so the tensor pipe runs on the next item while eight softmax warps read the finished one:
//     thread = one row x 128 columns), mask weight as one FMA, softmax partial state (max, sum, sum p.x, sum p.y) with
//     the source coordinates as V, written per (row, source, column half).
//     Measured on the B200: accumulating all of K = 512 in TMEM (96 MMAs) costs no accuracy here -- the truncating
//     accumulate acts as a 1e-6 relative temperature change (warp grids 1.06e-5 from fp64 vs 0.95e-5 when the partial
//     sums are promoted to registers every 2 K-blocks; the fp32 reference itself is 0.65e-5 off) -- so `chunk_kb`
//     defaults to the whole K and the register promotion of the conv GEMM is kept only as an option.
//  K2 corr_finish_kernel   (one warp per target position)  merge the partial states in fixed order -> warp grid ->
//     4-tap bilinear gather of the UN-normalised fp32 source features -> mean over sources -> fp32 output and / or the
//     hi/lo operand of the decoder's map_conv ("grid_sample fused with the following conv's load").
//
// warp roles in K1: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = work-list prefix, 4-11 = softmax.
#include "sm100_prims.cuh"
#include "host_util.h"
#include "../../include/tsnet_b200.h"
#include <math.h>
#include <stdlib.h>

namespace tsnet {

constexpr int kCorrM = 128;       // target rows per work item
constexpr int kCorrN = 256;       // source columns per work item
constexpr int kCorrNC = 128;      // columns owned by one thread
constexpr int kCorrK = 64;        // K block (one 128 B swizzle row)
constexpr int kCorrThreads = 384;
constexpr int kCorrMaxSrc = 12;
constexpr int kCorrMaxHW = 1024;
constexpr int kCorrMaxB = 1024;
constexpr int kCorrABytes = kCorrM * kCorrK * 2;                    // 16 KB
constexpr int kCorrBBytes = kCorrN * kCorrK * 2;                    // 32 KB
constexpr int kCorrStageBytes = 2 * kCorrABytes + 2 * kCorrBBytes;  // 96 KB
constexpr int kCorrStages = 2;
constexpr int kCorr2StageBytes = 4 * kCorrABytes;  // 2-CTA kernel: A hi/lo (128 rows) + B-half hi/lo (128 columns) = 64 KB
constexpr int kCorr2Stages = 3;
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------------------------
struct CorrWs {
  size_t rank, maskv, cxs, cys, cls, sums, items, counts, done, state, total;
};
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
static CorrWs corr_ws_layout(int B, int n, int hw) {
  CorrWs L;
  const size_t NM = static_cast<size_t>(n + 1) * B, runs = hw / kCorrM, chunks = hw / kCorrN;
  size_t o = 0;
  L.rank = o;   o = align256(o + NM * hw * sizeof(uint16_t));           // position -> sorted rank, per map
  L.maskv = o;  o = align256(o + NM * hw * sizeof(float));              // mask value at sorted rank
  L.cxs = o;    o = align256(o + static_cast<size_t>(n) * B * hw * 4);  // x coordinate of the source position at rank
  L.cys = o;    o = align256(o + static_cast<size_t>(n) * B * hw * 4);
  L.cls = o;    o = align256(o + NM * runs);                            // class of every 128-run: 0, 1, 2 = mixed
  L.sums = o;   o = align256(o + NM * runs * 2 * sizeof(float));                // (sum x, sum y) of every 128-run
  L.items = o;  o = align256(o + static_cast<size_t>(B) * n * runs * chunks * 4);
  L.counts = o; o = align256(o + static_cast<size_t>(B) * 4);
  L.done = o;   o = align256(o + static_cast<size_t>(B) * 4);             // sorted-map counter per sample (prepare)
  L.state = o;  o = align256(o + static_cast<size_t>(n) * B * hw * 2 * chunks * sizeof(float4));
  L.total = o;
  return L;
}

__device__ __forceinline__ float read_mask(const void* bbox, int dtype, int b, int bh, int bw, int h, int w, int pos) {
  // F.interpolate(mode='nearest'): src = min(floor(dst * (in / out)), in - 1), float scale (ATen)
  const int y = pos / w, x = pos - y * w;
  const int sy = min(static_cast<int>(floorf(y * (static_cast<float>(bh) / h))), bh - 1);
  const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(bw) / w))), bw - 1);
  const size_t off = (static_cast<size_t>(b) * bh + sy) * bw + sx;
  return dtype == 0 ? static_cast<float>(static_cast<const uint8_t*>(bbox)[off])
                    : static_cast<const float*>(bbox)[off];
}

// 2^x for x <= 0 (softmax weights): MUFU.EX2, results below 2^-126 flush to zero (they are < 1e-38 of the row maximum)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------------------
// P: masks -> sorted order, tile classes, work list, closed-form states of the skipped tiles
// ---------------------------------------------------------------------------------------------------------------
struct PrepArgs {
  const void* tar_bbox;
  const void* src_bbox[kCorrMaxSrc];
  const float* coord_table;  // h values (y) then w values (x)
  uint16_t* rank;
  float* maskv;
  float* cxs;
  float* cys;
  uint8_t* cls;
  float* sums;
  int* items;
  int* counts;
  int* done;
  float4* state;
  int B, n_src, h, w, hw, bbox_h, bbox_w, bbox_dtype, sort;
  int pair;  // 1: work items cover 256 target rows (two adjacent 128-row tiles) -- the 2-CTA tile kernel
};

// class sort of ONE map (q = 0: target, q >= 1: source q - 1) of sample b by the 1024 threads of a block
__device__ __forceinline__ void corr_sort_map(const PrepArgs& a, const int b, const int q) {
  __shared__ int wc[3][32];
  __shared__ float sv[kCorrMaxHW];
  __shared__ uint16_t sp[kCorrMaxHW];
  __shared__ float wsx[32], wsy[32];
  __shared__ uint8_t w1[32], w0[32];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int hw = a.hw, runs = hw / kCorrM;
  const bool active = t < hw;  // hw % 256 == 0: warps are entirely active or entirely idle
  const uint32_t lt = (1u << lane) - 1u;
  const size_t map = q == 0 ? b : static_cast<size_t>(a.B) + static_cast<size_t>(q - 1) * a.B + b;
  const void* bbox = q == 0 ? a.tar_bbox : a.src_bbox[q - 1];
  const float v = active ? read_mask(bbox, a.bbox_dtype, b, a.bbox_h, a.bbox_w, a.h, a.w, t) : 0.f;
  // sort key: exact ones first, soft values, exact zeros last (stable); sort = 0 keeps the raster order
  const int key = !active ? 3 : (a.sort ? (v == 1.f ? 0 : (v == 0.f ? 2 : 1)) : 0);
  const uint32_t m0 = __ballot_sync(0xffffffffu, key == 0), m1 = __ballot_sync(0xffffffffu, key == 1),
                 m2 = __ballot_sync(0xffffffffu, key == 2);
  if (lane == 0) {
    wc[0][warp] = __popc(m0);
    wc[1][warp] = __popc(m1);
    wc[2][warp] = __popc(m2);
  }
  __syncthreads();
  int before = 0, tot0 = 0, tot1 = 0;
  for (int ww = 0; ww < 32; ++ww) {
    const int c0 = wc[0][ww], c1 = wc[1][ww], c2 = wc[2][ww];
    tot0 += c0;
    tot1 += c1;
    if (ww < warp) before += key == 0 ? c0 : (key == 1 ? c1 : c2);
  }
  if (active) {
    const uint32_t mine = key == 0 ? m0 : (key == 1 ? m1 : m2);
    const int r = before + __popc(mine & lt) + (key == 1 ? tot0 : (key == 2 ? tot0 + tot1 : 0));
    a.rank[map * hw + t] = static_cast<uint16_t>(r);
    sv[r] = v;
    sp[r] = static_cast<uint16_t>(t);
  }
  __syncthreads();
  // ---- sorted order: thread t = rank t
  float v2 = 0.f, cx = 0.f, cy = 0.f;
  if (active) {
    v2 = sv[t];
    a.maskv[map * hw + t] = v2;
    if (q > 0) {
      const int pos = sp[t], y = pos / a.w, x = pos - y * a.w;
      cx = a.coord_table[a.h + x];
      cy = a.coord_table[y];
      const size_t o = (map - a.B) * hw + t;
      a.cxs[o] = cx;
      a.cys[o] = cy;
    }
  }
  const bool all1 = __all_sync(0xffffffffu, v2 == 1.f), all0 = __all_sync(0xffffffffu, v2 == 0.f);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    cx += __shfl_xor_sync(0xffffffffu, cx, o);
    cy += __shfl_xor_sync(0xffffffffu, cy, o);
  }
  if (lane == 0) {
    w1[warp] = all1;
    w0[warp] = all0;
    wsx[warp] = cx;
    wsy[warp] = cy;
  }
  __syncthreads();
  if (t < runs) {
    const bool o1 = w1[4 * t] && w1[4 * t + 1] && w1[4 * t + 2] && w1[4 * t + 3];
    const bool o0 = w0[4 * t] && w0[4 * t + 1] && w0[4 * t + 2] && w0[4 * t + 3];
    a.cls[map * runs + t] = o1 ? 1 : (o0 ? 0 : 2);
    // coordinate sums of every 128-run (closed form of the skipped tiles), fixed summation order
    a.sums[(map * runs + t) * 2 + 0] = ((wsx[4 * t] + wsx[4 * t + 1]) + wsx[4 * t + 2]) + wsx[4 * t + 3];
    a.sums[(map * runs + t) * 2 + 1] = ((wsy[4 * t] + wsy[4 * t + 1]) + wsy[4 * t + 2]) + wsy[4 * t + 3];
  }
}

// work list + closed-form states of the skipped (row tile, column chunk) pairs of sample b (whole block)
__device__ __forceinline__ void corr_plan_sample(const PrepArgs& a, const int b) {
  __shared__ uint8_t skip_sm[kCorrMaxSrc * 32];
  __shared__ int kc[32];
  __shared__ float2 sums_sm[kCorrMaxSrc * 8];  // (sum x, sum y) of every 128-run of every source map of this sample
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t < 32) kc[t] = 0;
  __syncthreads();
  const int hw = a.hw, runs = hw / kCorrM, chunks = hw / kCorrN, NS = 2 * chunks;
  const uint32_t lt = (1u << lane) - 1u;
  const int rtiles = a.pair ? runs / 2 : runs;                    // row tiles (pairs of 128-row tiles in pair mode)
  const int per_src = rtiles * chunks, ncand = a.n_src * per_src;  // <= 12 * 8 * 4 = 384 candidates
  if (t < a.n_src * runs) {  // one coalesced read instead of a dependent global load per skipped tile below
    const int i = t / runs, u = t - i * runs;
    const size_t smap = static_cast<size_t>(a.B) + static_cast<size_t>(i) * a.B + b;
    sums_sm[i * 8 + u] = __ldcg(reinterpret_cast<const float2*>(a.sums + (smap * runs + u) * 2));  // written by other blocks
  }
  bool keep = false;
  int code = 0;
  if (t < ncand) {  // candidates (source, row tile, column chunk) in lexicographic order
    const int i = t / per_src, rem = t - i * per_src, mt = rem / chunks, ch = rem - mt * chunks;
    const size_t smap = static_cast<size_t>(a.B) + static_cast<size_t>(i) * a.B + b;
    int rowc = __ldcg(a.cls + static_cast<size_t>(b) * runs + (a.pair ? 2 * mt : mt));
    if (a.pair && __ldcg(a.cls + static_cast<size_t>(b) * runs + 2 * mt + 1) != rowc) rowc = 2;
    const int c0 = __ldcg(a.cls + smap * runs + 2 * ch), c1 = __ldcg(a.cls + smap * runs + 2 * ch + 1);
    const int colc = c0 == c1 ? c0 : 2;
    const bool skip = (rowc == 1 && colc == 0) || (rowc == 0 && colc == 1);
    skip_sm[t] = skip;
    keep = !skip;
    code = mt | (ch << 4) | (i << 8);
  }
  const uint32_t km = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) kc[warp] = __popc(km);
  __syncthreads();
  int kbefore = 0, ktot = 0;
  for (int ww = 0; ww < 32; ++ww) {
    if (ww < warp) kbefore += kc[ww];
    ktot += kc[ww];
  }
  if (keep) a.items[static_cast<size_t>(b) * ncand + kbefore + __popc(km & lt)] = code;
  if (t == 0) a.counts[b] = ktot;
  // ---- skipped tiles: all 256 logits are exactly 0 -> per column half: max 0, weight 128, coordinate sums
  const int rows_per_item = a.pair ? 2 * kCorrM : kCorrM;
  for (int c = 0; c < ncand; ++c) {
    if (!skip_sm[c]) continue;
    const int i = c / per_src, rem = c - i * per_src, mt = rem / chunks, ch = rem - mt * chunks;
    for (int e = t; e < 2 * rows_per_item; e += blockDim.x) {
      const int row = e % rows_per_item, half = e / rows_per_item;
      const float2 sm = sums_sm[i * 8 + 2 * ch + half];
      const size_t r = (static_cast<size_t>(i) * a.B + b) * hw + mt * rows_per_item + row;
      a.state[r * NS + 2 * ch + half] = make_float4(0.f, static_cast<float>(kCorrNC), sm.x, sm.y);
    }
  }
}

// ONE launch per forward: grid (B, n_src + 1), block (b, q) sorts map q of sample b; the block that finishes LAST for a
// sample (a per-sample counter in the workspace, zeroed by a memset node ahead of the launch and reset here) builds that
// sample's work list.  The result does not depend on which block that is.  (A second plan launch cost 13 us of device
// time plus the launch gap for 1 MB of data; one block per sample sorting its maps serially was slower still: 40 us.)
__global__ void __launch_bounds__(1024) corr_prepare_kernel(const PrepArgs a) {
  __shared__ int is_last;
  const int b = blockIdx.x;
  corr_sort_map(a, b, blockIdx.y);
  __threadfence();   // this block's rank / class / sum tables are visible device-wide before the counter moves
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&a.done[b], 1);
    is_last = prev == a.n_src;
    if (is_last) a.done[b] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();   // acquire side: the other blocks' tables
  corr_plan_sample(a, b);
}

// ---------------------------------------------------------------------------------------------------------------
// K1: tensor-core tiles of the work list -> partial softmax states
// ---------------------------------------------------------------------------------------------------------------
struct alignas(64) CorrArgs {
  CUtensorMap t_hi, t_lo, s_hi, s_lo;  // [B*hw, C] box {64, 128} and [n_src*B*hw, C] box {64, 256}; rows in sorted order
  const float* maskv;
  const float* cxs;
  const float* cys;
  const int* items;
  const int* counts;
  float4* state;
  // un-normalised operands (tsnet_corr_operands / tsnet_wino_bridge): 1 / max(||x||, 1e-12) of every target row and
  // source column at its sorted rank; F.normalize (model/TSNet.py:319, :339) is then applied as a per-row x per-column
  // scale of the similarity inside the softmax FMA.  Null = the operands are already normalised (tsnet_l2norm_split).
  const float* rn_t;
  const float* rn_s;
  int B, n_src, C, hw, ncand, NS;
  int split, fmt, chunk_kb;
  float k2;  // temperature / operand_scale * log2(e)
};

struct CorrSmemTail {
  uint64_t full_bar[3], empty_bar[3], tmem_full[2], tmem_empty[2];  // (3 = max stages of the two tile kernels)
  uint32_t tmem_base;
  uint32_t pad[15];
  // per item (double-buffered), for the 256 source columns: mask * rnorm, x, y, rnorm
  alignas(16) float tab[2][4][kCorrN];
  int pref[kCorrMaxB + 1];              // exclusive prefix of the per-sample work-list lengths
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

#define TSNET_R32(v, c) c(v[0]), c(v[1]), c(v[2]), c(v[3]), c(v[4]), c(v[5]), c(v[6]), c(v[7]), c(v[8]), c(v[9]),    \
    c(v[10]), c(v[11]), c(v[12]), c(v[13]), c(v[14]), c(v[15]), c(v[16]), c(v[17]), c(v[18]), c(v[19]), c(v[20]),    \
    c(v[21]), c(v[22]), c(v[23]), c(v[24]), c(v[25]), c(v[26]), c(v[27]), c(v[28]), c(v[29]), c(v[30]), c(v[31])
// issue a 32-lane x 32-column TMEM load without waiting for it (registers are valid after tmem_ld32_wait on them)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : TSNET_R32(v, "=f")
      : "r"(taddr)
      : "memory");
}
// waits for ALL outstanding TMEM loads of the thread; the operands tie the consumers of `v` to the wait
__device__ __forceinline__ void tmem_ld32_wait(float* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : TSNET_R32(v, "+f")::"memory");
}
__device__ __forceinline__ void reg_fence32(float* v) { asm volatile("" : TSNET_R32(v, "+f")::"memory"); }

__device__ __forceinline__ void decode_item(const CorrArgs& a, const int* pref, int g, int& b, int& mt, int& i,
                                            int& ch) {
  int lo = 0, hi = a.B;  // invariant: pref[lo] <= g < pref[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pref[mid] <= g) lo = mid; else hi = mid;
  }
  b = lo;
  const int code = a.items[static_cast<size_t>(b) * a.ncand + (g - pref[b])];
  mt = code & 15;
  ch = (code >> 4) & 15;
  i = code >> 8;
}

__global__ void __launch_bounds__(kCorrThreads, 1) corr_tile_kernel(const __grid_constant__ CorrArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  CorrSmemTail& tl = *reinterpret_cast<CorrSmemTail*>(smem + kCorrStages * kCorrStageBytes);

  const int warp = threadIdx.x >> 5;
  const int num_kb = args.C / kCorrK;

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
  if (warp == 2) tmem_alloc(&tl.tmem_base, 2 * kCorrN);  // two 128 x 256 fp32 accumulators
  if (warp == 3) {
    int carry = 0;
    if (lane_id() == 0) tl.pref[0] = 0;
    for (int base = 0; base < args.B; base += 32) {
      const int idx = base + lane_id();
      int v = idx < args.B ? args.counts[idx] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (static_cast<int>(lane_id()) >= o) v += u;
      }
      if (idx < args.B) tl.pref[idx + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tl.tmem_base;
  const int total = tl.pref[args.B];

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane_id() == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t stage_tx = args.split ? kCorrStageBytes : (kCorrABytes + kCorrBBytes);
        for (int g = blockIdx.x; g < total; g += gridDim.x) {
          int b, mt, i, ch;
          decode_item(args, tl.pref, g, b, mt, i, ch);
          const int trow = b * args.hw + mt * kCorrM;
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
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      {  // whole warp, converged; one elected lane issues (see sm100_prims.cuh)
        const uint32_t idesc = make_idesc_f16(kCorrM, kCorrN, args.fmt);
        int stage = 0;
        uint32_t phase = 0;
        int cc = 0;  // partial-accumulator counter -> TMEM buffer + phase
        for (int g = blockIdx.x; g < total; g += gridDim.x) {
          for (int kb0 = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++cc) {
            const int buf = cc & 1;
            const uint32_t buf_phase = (cc >> 1) & 1;
            mbar_wait(&tl.tmem_empty[buf], buf_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * kCorrN;
            const int kb1 = min(num_kb, kb0 + args.chunk_kb);
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&tl.full_bar[stage], phase);
              tc_fence_after();
              const uint32_t st = smem_u32(smem + stage * kCorrStageBytes);
              const uint64_t a_hi = make_desc_kmajor_sw128(st);
              const uint64_t a_lo = make_desc_kmajor_sw128(st + kCorrABytes);
              const uint64_t b_hi = make_desc_kmajor_sw128(st + 2 * kCorrABytes);
              const uint64_t b_lo = make_desc_kmajor_sw128(st + 2 * kCorrABytes + kCorrBBytes);
              if (elect_one()) {  // one election per K block: the 12 MMAs are issued back to back by the leader
#pragma unroll
                for (int k = 0; k < kCorrK / 16; ++k) {
                  const uint32_t off = k * 32;
                  umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_hi, off), idesc, ((kb - kb0) | k) != 0);
                  if (args.split) {
                    umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_lo, off), idesc, 1);
                    umma_f16(d_tmem, desc_advance_k(a_lo, off), desc_advance_k(b_hi, off), idesc, 1);
                  }
                }
              }
              __syncwarp();
              umma_commit_elect(&tl.empty_bar[stage]);
              if (++stage == kCorrStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_elect(&tl.tmem_full[buf]);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax =====================
    const int q = warp & 3;            // TMEM lane quarter
    const int half = (warp - 4) >> 2;  // column half of the 256-column chunk
    const int row = q * 32 + lane_id();
    const int et = threadIdx.x - 128;  // 0..255 among the softmax threads
    int it = 0, par = 0;
    for (int g = blockIdx.x; g < total; g += gridDim.x, par ^= 1) {
      int b, mt, i, ch;
      decode_item(args, tl.pref, g, b, mt, i, ch);
      // ---- column tables of this item (the table of item k is last read before the barrier of item k+1, so
      //      writing buffer `par` again at item k+2 is safe)
      {
        const size_t sb = (static_cast<size_t>(i) * args.B + b) * args.hw + ch * kCorrN + et;
        const float rs = args.rn_s ? args.rn_s[sb] : 1.f;
        tl.tab[par][0][et] = args.maskv[static_cast<size_t>(args.B) * args.hw + sb] * rs;
        tl.tab[par][1][et] = args.cxs[sb];
        tl.tab[par][2][et] = args.cys[sb];
        tl.tab[par][3][et] = rs;
      }
      const float m_t = args.maskv[static_cast<size_t>(b) * args.hw + mt * kCorrM + row];
      const float r_t = args.rn_t ? args.rn_t[static_cast<size_t>(b) * args.hw + mt * kCorrM + row] : 1.f;
      // (T*mt).(S*ms) + (T*(1-mt)).(S*(1-ms)) == (T.S) * (mt*ms + (1-mt)*(1-ms)) = (T.S) * (wa*ms + wb);
      // exact for binary masks (mismatched pairs get logit 0, not -inf, as in the reference).  k2 folds the temperature,
      // the operand scale and log2(e): logits are kept in log2 units.
      // with un-normalised operands the weight of column j is r_t * r_s[j] * (wa' * ms_j + wb') =
      // wa * (ms_j * r_s[j]) + wb * r_s[j]  (tables 0 and 3)
      const float wa = (2.f * m_t - 1.f) * args.k2 * r_t, wb = (1.f - m_t) * args.k2 * r_t;
      epi_bar_sync();

      // ---- accumulator -> registers (one TMEM read when the whole K is accumulated in TMEM)
      float acc[kCorrNC];
      for (int kb0 = 0, p = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++p, ++it) {
        const int buf = it & 1;
        const uint32_t buf_phase = (it >> 1) & 1;
        mbar_wait(&tl.tmem_full[buf], buf_phase);
        tc_fence_after();
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kCorrN + half * kCorrNC;
        if (p == 0) {
#pragma unroll
          for (int c0 = 0; c0 < kCorrNC; c0 += 32) tmem_ld32_issue(t0 + c0, acc + c0);
          tmem_ld32_wait(acc);
#pragma unroll
          for (int c0 = 32; c0 < kCorrNC; c0 += 32) reg_fence32(acc + c0);
        } else {
#pragma unroll
          for (int c0 = 0; c0 < kCorrNC; c0 += 32) {
            float v[32];
            tmem_ld32_issue(t0 + c0, v);
            tmem_ld32_wait(v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c0 + j] += v[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&tl.tmem_empty[buf]);
      }

      // ---- logits (log2 units) and their maximum
      const float* tmask = &tl.tab[par][0][half * kCorrNC];
      const float* tcx = &tl.tab[par][1][half * kCorrNC];
      const float* tcy = &tl.tab[par][2][half * kCorrNC];
      const float* trn = &tl.tab[par][3][half * kCorrNC];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kCorrNC; j += 4) {
        const float4 mk = *reinterpret_cast<const float4*>(tmask + j);
        const float4 rk = *reinterpret_cast<const float4*>(trn + j);
        acc[j + 0] *= fmaf(wa, mk.x, wb * rk.x);
        acc[j + 1] *= fmaf(wa, mk.y, wb * rk.y);
        acc[j + 2] *= fmaf(wa, mk.z, wb * rk.z);
        acc[j + 3] *= fmaf(wa, mk.w, wb * rk.w);
        mx = fmaxf(mx, fmaxf(fmaxf(acc[j], acc[j + 1]), fmaxf(acc[j + 2], acc[j + 3])));
      }
      // ---- softmax partial state with the source coordinates as V
      float sum = 0.f, gx = 0.f, gy = 0.f;
#pragma unroll
      for (int j = 0; j < kCorrNC; j += 4) {
        const float4 vx = *reinterpret_cast<const float4*>(tcx + j);
        const float4 vy = *reinterpret_cast<const float4*>(tcy + j);
        const float xs[4] = {vx.x, vx.y, vx.z, vx.w}, ys[4] = {vy.x, vy.y, vy.z, vy.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float p = ex2_approx(acc[j + u] - mx);  // argument <= 0
          sum += p;
          gx = fmaf(p, xs[u], gx);
          gy = fmaf(p, ys[u], gy);
        }
      }
      const size_t r = (static_cast<size_t>(i) * args.B + b) * args.hw + mt * kCorrM + row;
      args.state[r * args.NS + 2 * ch + half] = make_float4(mx, sum, gx, gy);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kCorrN);
  }
}

// 2-CTA variant (cta_group::2): a CTA pair takes 256 target rows x 256 source columns; each CTA loads its own 128
// rows of T and only HALF of the S chunk, the leader issues M = 256 MMAs that read both halves.  Operand delivery per
// SM drops from 96 KB to 64 KB per K block (the 1-CTA kernel needs 62 B/clk/SM from L2 to keep the tensor pipe busy).
__global__ void __launch_bounds__(kCorrThreads, 1) corr_tile2_kernel(const __grid_constant__ CorrArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  CorrSmemTail& tl = *reinterpret_cast<CorrSmemTail*>(smem + kCorr2Stages * kCorr2StageBytes);
  const uint32_t rank = cluster_ctarank();  // 0 = leader: issues the MMAs of the pair and owns the full / tmem_empty barriers

  const int warp = threadIdx.x >> 5;
  const int num_kb = args.C / kCorrK;

  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&args.t_hi);
    tma_prefetch_desc(&args.s_hi);
    if (args.split) {
      tma_prefetch_desc(&args.t_lo);
      tma_prefetch_desc(&args.s_lo);
    }
  }
  if (warp == 1 && lane_id() == 0) {
    for (int s = 0; s < kCorr2Stages; ++s) {
      mbar_init(&tl.full_bar[s], 1);   // leader's copy is the one in use: its producer arrives, both CTAs' TMA add bytes
      mbar_init(&tl.empty_bar[s], 1);  // armed in both CTAs by the leader's multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tl.tmem_full[a], 1);    // multicast commit
      mbar_init(&tl.tmem_empty[a], 16);  // leader's copy: one arrive per softmax warp of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(&tl.tmem_base, 2 * kCorrN);  // two 128 x 256 fp32 accumulators in each CTA
  if (warp == 3) {
    int carry = 0;
    if (lane_id() == 0) tl.pref[0] = 0;
    for (int base = 0; base < args.B; base += 32) {
      const int idx = base + lane_id();
      int v = idx < args.B ? args.counts[idx] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (static_cast<int>(lane_id()) >= o) v += u;
      }
      if (idx < args.B) tl.pref[idx + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM exist before any remote arrive / 2-CTA MMA
  tc_fence_after();
  const uint32_t tmem_base = tl.tmem_base;
  const int total = tl.pref[args.B];
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane_id() == 0) {
        int stage = 0;
        uint32_t phase = 0;
        // per stage and CTA: A = own 128 rows (hi, lo), B = own HALF (128 of the 256 columns) of the source chunk
        const uint32_t stage_tx = 2u * (args.split ? kCorr2StageBytes : kCorr2StageBytes / 2);  // bytes of BOTH CTAs
        for (int g = pair_id; g < total; g += num_pairs) {
          int b, mt, i, ch;
          decode_item(args, tl.pref, g, b, mt, i, ch);
          const int trow = b * args.hw + (2 * mt + static_cast<int>(rank)) * kCorrM;
          const int srow = (i * args.B + b) * args.hw + ch * kCorrN + static_cast<int>(rank) * (kCorrN / 2);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&tl.empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * kCorr2StageBytes;
            if (rank == 0) mbar_arrive_expect_tx(&tl.full_bar[stage], stage_tx);
            tma_load_2d_2sm(st, &args.t_hi, &tl.full_bar[stage], kb * kCorrK, trow);
            tma_load_2d_2sm(st + 2 * kCorrABytes, &args.s_hi, &tl.full_bar[stage], kb * kCorrK, srow);
            if (args.split) {
              tma_load_2d_2sm(st + kCorrABytes, &args.t_lo, &tl.full_bar[stage], kb * kCorrK, trow);
              tma_load_2d_2sm(st + 3 * kCorrABytes, &args.s_lo, &tl.full_bar[stage], kb * kCorrK, srow);
            }
            if (++stage == kCorr2Stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      // ===================== MMA issuer (leader CTA only: M = 256 across the pair) =====================
      {  // whole warp, converged; one elected lane issues (see sm100_prims.cuh)
        const uint32_t idesc = make_idesc_f16(2 * kCorrM, kCorrN, args.fmt);
        int stage = 0;
        uint32_t phase = 0;
        int cc = 0;  // partial-accumulator counter -> TMEM buffer + phase
        for (int g = pair_id; g < total; g += num_pairs) {
          for (int kb0 = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++cc) {
            const int buf = cc & 1;
            const uint32_t buf_phase = (cc >> 1) & 1;
            mbar_wait(&tl.tmem_empty[buf], buf_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * kCorrN;
            const int kb1 = min(num_kb, kb0 + args.chunk_kb);
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&tl.full_bar[stage], phase);
              tc_fence_after();
              const uint32_t st = smem_u32(smem + stage * kCorr2StageBytes);
              const uint64_t a_hi = make_desc_kmajor_sw128(st);
              const uint64_t a_lo = make_desc_kmajor_sw128(st + kCorrABytes);
              const uint64_t b_hi = make_desc_kmajor_sw128(st + 2 * kCorrABytes);
              const uint64_t b_lo = make_desc_kmajor_sw128(st + 3 * kCorrABytes);
              if (elect_one()) {  // one election per K block: the 12 MMAs are issued back to back by the leader
#pragma unroll
                for (int k = 0; k < kCorrK / 16; ++k) {
                  const uint32_t off = k * 32;
                  umma_f16_2sm(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_hi, off), idesc, ((kb - kb0) | k) != 0);
                  if (args.split) {
                    umma_f16_2sm(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_lo, off), idesc, 1);
                    umma_f16_2sm(d_tmem, desc_advance_k(a_lo, off), desc_advance_k(b_hi, off), idesc, 1);
                  }
                }
              }
              __syncwarp();
              if (elect_one()) umma_commit_2sm(&tl.empty_bar[stage]);
              __syncwarp();
              if (++stage == kCorr2Stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit_2sm(&tl.tmem_full[buf]);
            __syncwarp();
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax =====================
    const int q = warp & 3;            // TMEM lane quarter
    const int half = (warp - 4) >> 2;  // column half of the 256-column chunk
    const int row = q * 32 + lane_id();
    const int et = threadIdx.x - 128;  // 0..255 among the softmax threads
    int it = 0, par = 0;
    for (int g = pair_id; g < total; g += num_pairs, par ^= 1) {
      int b, mt, i, ch;
      decode_item(args, tl.pref, g, b, mt, i, ch);
      // ---- column tables of this item (the table of item k is last read before the barrier of item k+1, so
      //      writing buffer `par` again at item k+2 is safe)
      {
        const size_t sb = (static_cast<size_t>(i) * args.B + b) * args.hw + ch * kCorrN + et;
        const float rs = args.rn_s ? args.rn_s[sb] : 1.f;
        tl.tab[par][0][et] = args.maskv[static_cast<size_t>(args.B) * args.hw + sb] * rs;
        tl.tab[par][1][et] = args.cxs[sb];
        tl.tab[par][2][et] = args.cys[sb];
        tl.tab[par][3][et] = rs;
      }
      const int mrow = (2 * mt + static_cast<int>(rank)) * kCorrM + row;  // this CTA's 128 of the item's 256 rows
      const float m_t = args.maskv[static_cast<size_t>(b) * args.hw + mrow];
      const float r_t = args.rn_t ? args.rn_t[static_cast<size_t>(b) * args.hw + mrow] : 1.f;
      // (T*mt).(S*ms) + (T*(1-mt)).(S*(1-ms)) == (T.S) * (mt*ms + (1-mt)*(1-ms)) = (T.S) * (wa*ms + wb);
      // exact for binary masks (mismatched pairs get logit 0, not -inf, as in the reference).  k2 folds the temperature,
      // the operand scale and log2(e): logits are kept in log2 units.
      // with un-normalised operands the weight of column j is r_t * r_s[j] * (wa' * ms_j + wb') =
      // wa * (ms_j * r_s[j]) + wb * r_s[j]  (tables 0 and 3)
      const float wa = (2.f * m_t - 1.f) * args.k2 * r_t, wb = (1.f - m_t) * args.k2 * r_t;
      epi_bar_sync();

      // ---- accumulator -> registers (one TMEM read when the whole K is accumulated in TMEM)
      float acc[kCorrNC];
      for (int kb0 = 0, p = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++p, ++it) {
        const int buf = it & 1;
        const uint32_t buf_phase = (it >> 1) & 1;
        mbar_wait(&tl.tmem_full[buf], buf_phase);
        tc_fence_after();
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kCorrN + half * kCorrNC;
        if (p == 0) {
#pragma unroll
          for (int c0 = 0; c0 < kCorrNC; c0 += 32) tmem_ld32_issue(t0 + c0, acc + c0);
          tmem_ld32_wait(acc);
#pragma unroll
          for (int c0 = 32; c0 < kCorrNC; c0 += 32) reg_fence32(acc + c0);
        } else {
#pragma unroll
          for (int c0 = 0; c0 < kCorrNC; c0 += 32) {
            float v[32];
            tmem_ld32_issue(t0 + c0, v);
            tmem_ld32_wait(v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c0 + j] += v[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) {
          if (rank == 0) mbar_arrive(&tl.tmem_empty[buf]);
          else mbar_arrive_remote_cta(map_to_cta(smem_u32(&tl.tmem_empty[buf]), 0));
        }
      }

      // ---- logits (log2 units) and their maximum
      const float* tmask = &tl.tab[par][0][half * kCorrNC];
      const float* tcx = &tl.tab[par][1][half * kCorrNC];
      const float* tcy = &tl.tab[par][2][half * kCorrNC];
      const float* trn = &tl.tab[par][3][half * kCorrNC];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kCorrNC; j += 4) {
        const float4 mk = *reinterpret_cast<const float4*>(tmask + j);
        const float4 rk = *reinterpret_cast<const float4*>(trn + j);
        acc[j + 0] *= fmaf(wa, mk.x, wb * rk.x);
        acc[j + 1] *= fmaf(wa, mk.y, wb * rk.y);
        acc[j + 2] *= fmaf(wa, mk.z, wb * rk.z);
        acc[j + 3] *= fmaf(wa, mk.w, wb * rk.w);
        mx = fmaxf(mx, fmaxf(fmaxf(acc[j], acc[j + 1]), fmaxf(acc[j + 2], acc[j + 3])));
      }
      // ---- softmax partial state with the source coordinates as V
      float sum = 0.f, gx = 0.f, gy = 0.f;
#pragma unroll
      for (int j = 0; j < kCorrNC; j += 4) {
        const float4 vx = *reinterpret_cast<const float4*>(tcx + j);
        const float4 vy = *reinterpret_cast<const float4*>(tcy + j);
        const float xs[4] = {vx.x, vx.y, vx.z, vx.w}, ys[4] = {vy.x, vy.y, vy.z, vy.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float p = ex2_approx(acc[j + u] - mx);  // argument <= 0
          sum += p;
          gx = fmaf(p, xs[u], gx);
          gy = fmaf(p, ys[u], gy);
        }
      }
      const size_t r = (static_cast<size_t>(i) * args.B + b) * args.hw + mrow;
      args.state[r * args.NS + 2 * ch + half] = make_float4(mx, sum, gx, gy);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal its barriers / read its tile
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * kCorrN);
  }
}

constexpr int kCorrSmemBytes = kCorrStages * kCorrStageBytes + 1024 + static_cast<int>(sizeof(CorrSmemTail));
static_assert(kCorrSmemBytes <= 227 * 1024, "corr_tile shared memory budget");
constexpr int kCorr2SmemBytes = kCorr2Stages * kCorr2StageBytes + 1024 + static_cast<int>(sizeof(CorrSmemTail));
static_assert(kCorr2SmemBytes <= 227 * 1024, "corr_tile2 shared memory budget");

// ---------------------------------------------------------------------------------------------------------------
// K2: (partial states -> warp grid) -> bilinear grid_sample of the n source feature maps + mean over sources, written
// as fp32 and / or as the operand (hi / lo tap source) of the decoder's map_conv (model/TSNet.py:359-366, :392, :163).
// One warp = one target position; lane owns channels {128 k + 4 lane .. +3}.  Runs with the whole L1 available
// (K1 leaves ~20 KB), which is what the 4-tap gather wants.  kFromStates = false: grids are given
// (tsnet_warp_mean_taps).
// ---------------------------------------------------------------------------------------------------------------
struct WarpTapsArgs {
  const float* src_fea[kCorrMaxSrc];
  const float* grids;        // [n, B, hw, 2] (kFromStates = false)
  const float4* state;       // [n, B, hw (sorted target rank), NS] (kFromStates = true)
  const uint16_t* rank_t;    // [B, hw] target position -> sorted rank
  float* out_grids;          // [n, B, hw, 2] or null (kFromStates = true)
  float* out_mean;           // fp32 [B, hw, C] or null
  uint16_t* hi;              // [B, hw, Cp_total] or null
  uint16_t* lo;
  int B, n_src, C, h, w, Cp_total, c_off, fmt, NS, csplit;
  float scale;
};

template <bool kFromStates>
__global__ void __launch_bounds__(256) warp_mean_taps_kernel(const WarpTapsArgs a) {
  const int hw = a.h * a.w;
  // warp = (target position, channel slab): csplit warps share a position so that more loads are in flight
  const size_t wid = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t gpos = wid / a.csplit;
  const int slab = static_cast<int>(wid - gpos * a.csplit);
  if (gpos >= static_cast<size_t>(a.B) * hw) return;
  const int b = static_cast<int>(gpos / hw);
  const int lane = threadIdx.x & 31;
  const int nk = a.C / 128 / a.csplit;  // 128-channel chunks of this warp (<= 4)
  const int cbase = slab * nk * 128 + lane * 4;
  const bool gather = a.out_mean != nullptr || a.hi != nullptr;

  // ---- phase A: the warp grid of every source.  Lanes 8j..8j+7 work on source 4*grp + j: with the partial states,
  // lane 8j+k loads state k and the 8 lanes merge them in a fixed butterfly order.
  float gxr[3], gyr[3];
  const int sub = lane >> 3, k8 = lane & 7;
#pragma unroll
  for (int grp = 0; grp < 3; ++grp) {
    gxr[grp] = 0.f;
    gyr[grp] = 0.f;
    const int i = grp * 4 + sub;
    if (grp * 4 >= a.n_src) continue;
    if (kFromStates) {
      float4 s = make_float4(-INFINITY, 0.f, 0.f, 0.f);
      if (i < a.n_src && k8 < a.NS) {
        const int trank = a.rank_t[gpos];
        s = a.state[((static_cast<size_t>(i) * a.B + b) * hw + trank) * a.NS + k8];
      }
      float m = s.x;
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float f = ex2_approx(s.x - m);  // 2^(-inf) = 0 for unused lanes (m is finite whenever the source exists)
      float l = s.y * f, sx = s.z * f, sy = s.w * f;
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) {
        l += __shfl_xor_sync(0xffffffffu, l, o);
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
      }
      gxr[grp] = sx / l;  // (kept as true divisions: the warp grid is an output, warp_grid2d_list)
      gyr[grp] = sy / l;
      if (a.out_grids && slab == 0 && k8 == 0 && i < a.n_src)
        *reinterpret_cast<float2*>(a.out_grids + (static_cast<size_t>(i) * a.B * hw + gpos) * 2) =
            make_float2(gxr[grp], gyr[grp]);
    } else if (i < a.n_src) {
      const float2 g = *reinterpret_cast<const float2*>(a.grids + (static_cast<size_t>(i) * a.B * hw + gpos) * 2);
      gxr[grp] = g.x;
      gyr[grp] = g.y;
    }
  }
  if (!gather) return;

  // ---- phase B: 4-tap gather of every source, accumulated in registers
  float4 accv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) accv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < a.n_src; ++i) {
    const int grp = i >> 2, src_lane = (i & 3) * 8;  // lane 8j holds the canonical value of its group
    const float gx = __shfl_sync(0xffffffffu, grp == 0 ? gxr[0] : (grp == 1 ? gxr[1] : gxr[2]), src_lane);
    const float gy = __shfl_sync(0xffffffffu, grp == 0 ? gyr[0] : (grp == 1 ? gyr[1] : gyr[2]), src_lane);
    // F.grid_sample(bilinear, zeros, align_corners=False): ix = ((x + 1) * W - 1) / 2
    const float ix = ((gx + 1.f) * a.w - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * a.h - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
    float wts[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};  // nw, ne, sw, se
    const float* base = a.src_fea[i] + static_cast<size_t>(b) * hw * a.C + cbase;
    const float* tp[4];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int xx = x0 + (tap & 1), yy = y0 + (tap >> 1);
      if (!(xx >= 0 && xx < a.w && yy >= 0 && yy < a.h)) wts[tap] = 0.f;  // zeros padding
      const int xc = min(max(xx, 0), a.w - 1), yc = min(max(yy, 0), a.h - 1);
      tp[tap] = base + static_cast<size_t>(yc * a.w + xc) * a.C;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
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
  // mean over sources (torch.stack(...).mean(1), model/TSNet.py:392) as a multiplication by 1/n: exact for n = 2^k,
  // within 1 ulp of the reference's division otherwise (8 IEEE division sequences per lane made this kernel issue-bound)
  const float rn = 1.f / static_cast<float>(a.n_src);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < nk) {
      const float v[4] = {accv[k].x * rn, accv[k].y * rn, accv[k].z * rn, accv[k].w * rn};
      const int c = cbase + k * 128;
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

// 2-CTA tile kernel (cta_group::2) is the default; tsnet_corr_desc.one_cta selects the 1-CTA kernel (tests compare the
// two).  tsnet_corr_prepare and tsnet_corr_tiles read the same descriptor field, so the work-list granularity always
// matches the kernel that consumes the list.
static bool corr_use_2cta(const tsnet_corr_desc* d) { return d->one_cta == 0; }

static int corr_desc_check(const tsnet_corr_desc* d) {
  TSNET_ARG_CHECK(d, "corr: null descriptor");
  TSNET_ARG_CHECK(d->n_src >= 1 && d->n_src <= kCorrMaxSrc, "corr: n_src %d (max %d)", d->n_src, kCorrMaxSrc);
  TSNET_ARG_CHECK(d->B >= 1 && d->B <= kCorrMaxB, "corr: B %d (max %d per call)", d->B, kCorrMaxB);
  const int hw = d->h * d->w;
  TSNET_ARG_CHECK(hw % kCorrN == 0 && hw <= kCorrMaxHW, "corr: h*w = %d must be a multiple of 256, <= %d", hw,
                  kCorrMaxHW);
  TSNET_ARG_CHECK(d->C % 128 == 0 && d->C <= 1024, "corr: C %d must be a multiple of 128, <= 1024", d->C);
  TSNET_ARG_CHECK(d->bbox_dtype == 0 || d->bbox_dtype == 1, "corr: bbox_dtype %d", d->bbox_dtype);
  return 0;
}

}  // namespace tsnet

using namespace tsnet;

extern "C" size_t tsnet_corr_workspace_bytes(const tsnet_corr_desc* d) {
  if (corr_desc_check(d)) return 0;
  return corr_ws_layout(d->B, d->n_src, d->h * d->w).total;
}

extern "C" const uint16_t* tsnet_corr_rank_table(const tsnet_corr_desc* d, const void* workspace, int which) {
  if (corr_desc_check(d) || !workspace) return nullptr;
  const CorrWs L = corr_ws_layout(d->B, d->n_src, d->h * d->w);
  const uint16_t* r = reinterpret_cast<const uint16_t*>(static_cast<const uint8_t*>(workspace) + L.rank);
  return which == 0 ? r : r + static_cast<size_t>(d->B) * d->h * d->w;
}

extern "C" int tsnet_corr_prepare(const tsnet_corr_desc* d, const void* tar_bbox, const void* const* src_bbox,
                                  const float* coord_table, void* workspace, size_t workspace_bytes, void* stream) {
  if (int r = corr_desc_check(d)) return r;
  TSNET_ARG_CHECK(tar_bbox && src_bbox && coord_table && workspace, "corr_prepare: null argument");
  const int hw = d->h * d->w;
  const CorrWs L = corr_ws_layout(d->B, d->n_src, hw);
  TSNET_ARG_CHECK(workspace_bytes >= L.total, "corr_prepare: workspace %zu B < %zu B", workspace_bytes, L.total);
  TSNET_ARG_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "corr_prepare: workspace must be 256 B aligned");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  a.tar_bbox = tar_bbox;
  for (int i = 0; i < d->n_src; ++i) {
    TSNET_ARG_CHECK(src_bbox[i], "corr_prepare: null source mask %d", i);
    a.src_bbox[i] = src_bbox[i];
  }
  a.coord_table = coord_table;
  a.rank = reinterpret_cast<uint16_t*>(ws + L.rank);
  a.maskv = reinterpret_cast<float*>(ws + L.maskv);
  a.cxs = reinterpret_cast<float*>(ws + L.cxs);
  a.cys = reinterpret_cast<float*>(ws + L.cys);
  a.cls = ws + L.cls;
  a.sums = reinterpret_cast<float*>(ws + L.sums);
  a.items = reinterpret_cast<int*>(ws + L.items);
  a.counts = reinterpret_cast<int*>(ws + L.counts);
  a.done = reinterpret_cast<int*>(ws + L.done);
  a.state = reinterpret_cast<float4*>(ws + L.state);
  a.B = d->B; a.n_src = d->n_src; a.h = d->h; a.w = d->w; a.hw = hw;
  a.bbox_h = d->bbox_h; a.bbox_w = d->bbox_w; a.bbox_dtype = d->bbox_dtype;
  a.sort = d->sort;
  a.pair = corr_use_2cta(d) ? 1 : 0;
  TSNET_CUDA_CHECK(cudaMemsetAsync(a.done, 0, static_cast<size_t>(d->B) * sizeof(int), static_cast<cudaStream_t>(stream)));
  corr_prepare_kernel<<<dim3(d->B, d->n_src + 1), 1024, 0, static_cast<cudaStream_t>(stream)>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsnet_corr_tiles(const tsnet_corr_desc* d, const uint16_t* tar_hi, const uint16_t* tar_lo,
                                const uint16_t* src_hi, const uint16_t* src_lo, const float* rnorm_t,
                                const float* rnorm_s, void* workspace, size_t workspace_bytes, void* stream) {
  if (int r = corr_desc_check(d)) return r;
  TSNET_ARG_CHECK(tar_hi && src_hi && workspace, "corr_tiles: null argument");
  TSNET_ARG_CHECK(!d->split || (tar_lo && src_lo), "corr_tiles: split mode needs the lo operands");
  TSNET_ARG_CHECK((rnorm_t == nullptr) == (rnorm_s == nullptr), "corr_tiles: rnorm_t / rnorm_s come together");
  const int hw = d->h * d->w;
  const CorrWs L = corr_ws_layout(d->B, d->n_src, hw);
  TSNET_ARG_CHECK(workspace_bytes >= L.total, "corr_tiles: workspace %zu B < %zu B", workspace_bytes, L.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);

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
  const bool two_cta = corr_use_2cta(d);
  {
    const uint64_t dims[2] = {(uint64_t)d->C, (uint64_t)d->n_src * d->B * hw};
    const uint64_t str[1] = {(uint64_t)d->C * 2};
    const uint32_t box[2] = {64, two_cta ? (uint32_t)(kCorrN / 2) : (uint32_t)kCorrN};  // 2-CTA: each CTA loads half a chunk
    int r = encode_tmap_u16_sw128(&a.s_hi, src_hi, 2, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.s_lo, src_lo, 2, dims, str, box))) return r;
  }
  a.maskv = reinterpret_cast<const float*>(ws + L.maskv);
  a.cxs = reinterpret_cast<const float*>(ws + L.cxs);
  a.cys = reinterpret_cast<const float*>(ws + L.cys);
  a.items = reinterpret_cast<const int*>(ws + L.items);
  a.counts = reinterpret_cast<const int*>(ws + L.counts);
  a.state = reinterpret_cast<float4*>(ws + L.state);
  a.rn_t = rnorm_t;
  a.rn_s = rnorm_s;
  a.B = d->B; a.n_src = d->n_src; a.C = d->C; a.hw = hw;
  a.ncand = d->n_src * (hw / kCorrM) * (hw / kCorrN) / (two_cta ? 2 : 1);
  a.NS = 2 * (hw / kCorrN);
  a.split = d->split; a.fmt = d->fmt;
  a.chunk_kb = d->chunk_kb > 0 ? d->chunk_kb : d->C / kCorrK;  // default: whole K accumulated in TMEM (header comment)
  // operand_scale is a power of two, so folding it into the temperature is exact up to one rounding of the product
  a.k2 = d->temperature / (d->operand_scale == 0.f ? 1.f : d->operand_scale) * kLog2e;

  const int max_items = d->B * a.ncand;
  if (two_cta) {
    static int attr2[kMaxDevices] = {0};
    TSNET_CUDA_CHECK(ensure_dyn_smem(corr_tile2_kernel, kCorr2SmemBytes, attr2));
    const int pairs = max_items < num_sms() / 2 ? max_items : num_sms() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kCorrThreads);
    cfg.dynamicSmemBytes = kCorr2SmemBytes;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ++launch_counter();
    TSNET_CUDA_CHECK(cudaLaunchKernelEx(&cfg, corr_tile2_kernel, a));
    return 0;
  }
  static int attr1[kMaxDevices] = {0};
  TSNET_CUDA_CHECK(ensure_dyn_smem(corr_tile_kernel, kCorrSmemBytes, attr1));
  const int grid = max_items < num_sms() ? max_items : num_sms();
  corr_tile_kernel<<<grid, kCorrThreads, kCorrSmemBytes, static_cast<cudaStream_t>(stream)>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}

static int launch_warp_taps(WarpTapsArgs& a, bool from_states, void* stream) {
  const size_t warps = static_cast<size_t>(a.B) * a.h * a.w * a.csplit;
  const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
  if (from_states)
    warp_mean_taps_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  else
    warp_mean_taps_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}

static int fill_warp_taps(WarpTapsArgs& a, const float* const* src_fea, int n_src, int B, int h, int w, int C,
                          float* out_mean, uint16_t* taps_hi, uint16_t* taps_lo, int Cp_total, int c_off, int fmt,
                          float scale, bool need_fea) {
  TSNET_ARG_CHECK(n_src >= 1 && n_src <= kCorrMaxSrc, "warp_mean_taps: n_src %d (max %d)", n_src, kCorrMaxSrc);
  TSNET_ARG_CHECK(C % 128 == 0 && C <= 1024, "warp_mean_taps: C %d must be a multiple of 128, <= 1024", C);
  TSNET_ARG_CHECK((taps_hi == nullptr) == (taps_lo == nullptr), "warp_mean_taps: hi/lo must both be given or both NULL");
  TSNET_ARG_CHECK(!taps_hi || (Cp_total % 4 == 0 && c_off % 4 == 0 && c_off + C <= Cp_total),
                  "warp_mean_taps: channel window does not fit");
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < n_src && need_fea; ++i) {
    TSNET_ARG_CHECK(src_fea && src_fea[i], "warp_mean_taps: null source %d", i);
    a.src_fea[i] = src_fea[i];
  }
  a.out_mean = out_mean; a.hi = taps_hi; a.lo = taps_lo;
  a.B = B; a.n_src = n_src; a.C = C; a.h = h; a.w = w; a.Cp_total = Cp_total; a.c_off = c_off; a.fmt = fmt;
  a.scale = scale == 0.f ? 1.f : scale;
  a.csplit = (C / 128) % 2 == 0 ? 2 : 1;  // warps per position (each takes C / csplit channels, at most 512)
  TSNET_ARG_CHECK(C / 128 / a.csplit <= 4, "warp_mean_taps: C %d not supported (128, 256, 384, 512, 768, 1024)", C);
  return 0;
}

extern "C" int tsnet_corr_finish(const tsnet_corr_desc* d, const float* const* src_fea, float* out_mean,
                                 float* out_grids, uint16_t* taps_hi, uint16_t* taps_lo, int Cp_total, int c_off,
                                 float taps_scale, void* workspace, size_t workspace_bytes, void* stream) {
  if (int r = corr_desc_check(d)) return r;
  TSNET_ARG_CHECK(workspace && (out_mean || out_grids || taps_hi), "corr_finish: null argument");
  const int hw = d->h * d->w;
  const CorrWs L = corr_ws_layout(d->B, d->n_src, hw);
  TSNET_ARG_CHECK(workspace_bytes >= L.total, "corr_finish: workspace %zu B < %zu B", workspace_bytes, L.total);
  const uint8_t* ws = static_cast<const uint8_t*>(workspace);
  WarpTapsArgs a;
  if (int r = fill_warp_taps(a, src_fea, d->n_src, d->B, d->h, d->w, d->C, out_mean, taps_hi, taps_lo, Cp_total, c_off,
                             d->fmt, taps_scale, out_mean || taps_hi))
    return r;
  a.state = reinterpret_cast<const float4*>(ws + L.state);
  a.rank_t = reinterpret_cast<const uint16_t*>(ws + L.rank);
  a.out_grids = out_grids;
  a.NS = 2 * (hw / kCorrN);
  return launch_warp_taps(a, true, stream);
}

extern "C" int tsnet_corr_warp_fwd(const tsnet_corr_desc* d, const uint16_t* tar_hi, const uint16_t* tar_lo,
                                   const uint16_t* src_hi, const uint16_t* src_lo, const float* rnorm_t,
                                   const float* rnorm_s, const float* const* src_fea, float* out_mean,
                                   float* out_grids, uint16_t* taps_hi, uint16_t* taps_lo, int Cp_total, int c_off,
                                   float taps_scale, void* workspace, size_t workspace_bytes, void* stream) {
  if (int r = tsnet_corr_tiles(d, tar_hi, tar_lo, src_hi, src_lo, rnorm_t, rnorm_s, workspace, workspace_bytes, stream))
    return r;
  return tsnet_corr_finish(d, src_fea, out_mean, out_grids, taps_hi, taps_lo, Cp_total, c_off, taps_scale, workspace,
                           workspace_bytes, stream);
}

extern "C" int tsnet_warp_mean_taps(const float* const* src_fea, int n_src, const float* grids, int B, int h, int w,
                                    int C, float* out_mean, uint16_t* taps_hi, uint16_t* taps_lo, int Cp_total,
                                    int c_off, int fmt, float scale, void* stream) {
  TSNET_ARG_CHECK(src_fea && grids && (out_mean || taps_hi), "warp_mean_taps: null argument");
  WarpTapsArgs a;
  if (int r = fill_warp_taps(a, src_fea, n_src, B, h, w, C, out_mean, taps_hi, taps_lo, Cp_total, c_off, fmt, scale,
                             true))
    return r;
  a.grids = grids;
  return launch_warp_taps(a, false, stream);
}
