// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   y_raw[b, y, x, co] = bias[co] + out_scale * sum_{tap, c} taps[b*planes + plane(tap), y + dy(tap), x + dx(tap), c]
//                                                            * w[co, tap*Cp + c]
//
// A operand  = 128 output pixels x 64 channels of one tap, fetched by ONE 4-D TMA box from the padded
//              NHWC "tap source" (padding / upsampling / parity split were applied by its producer);
// B operand  = BLOCK_N x 64 slice of the K-major packed weight, 2-D TMA box;
// both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly.
// fp32-faithful mode: operands are stored as 16-bit (hi, lo) pairs and every K step issues
//   D += Ahi*Bhi ; D += Ahi*Blo ; D += Alo*Bhi      (fp32 accumulation in TMEM).
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-11 = accumulate + epilogue.
//
// Accumulation precision: tcgen05.mma adds into the fp32 TMEM accumulator with truncation, and the error grows
// linearly with the number of accumulating instructions (measured: 2e-5 relative at K = 4608).  The K loop is
// therefore cut into chunks of `chunk_kb` K-blocks; each chunk accumulates from zero into one of two TMEM
// buffers, and the 8 accumulate warps add the finished chunk into fp32 REGISTER accumulators with round-to-nearest
// (each thread owns one output pixel x BLOCK_N/2 channels).  The MMA pipe runs ahead by one chunk.  At the end of
// the tile the same warps apply scale + bias, store the row, and produce the InstanceNorm partial statistics.
#include "sm100_prims.cuh"
#include "host_util.h"
#include "../../include/tsnet_b200.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace tsnet {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 x 16-bit = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;   // 4 control warps + 8 accumulate/epilogue warps
constexpr int kAccWarps = 8;
constexpr int kSmemBudget = 200 * 1024;

struct alignas(64) ConvGemmArgs {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  const float* bias;
  const float* addend;
  float* y;
  float* stats;
  float out_scale;
  int addend_rows;
  int num_m_tiles, num_n_tiles, tiles_per_img, wtiles_per_row, rows_per_tile, Wt;
  int m_tile_begin;  // this launch covers pixel tiles [m_tile_begin, m_tile_begin + num_m_tiles) (tail-wave split)
  int Cout, num_taps, kc_per_tap, planes, split, fmt, chunk_kb;
  // batched GEMM over `batch_planes` independent (A plane, weight matrix, output plane) triples in ONE launch: the 16
  // Winograd-domain contractions of a 3x3 convolution (tsnet_wino_gemm_fwd).  1 for ordinary convolutions.
  // plane p reads tap-source plane index img * planes + tap_plane + p, weight rows [p * b_plane_rows, ...) and writes
  // y + p * y_plane_stride.
  int batch_planes, b_plane_rows;
  long long y_plane_stride;
  // Winograd-domain layouts (tsnet_wino_gemm_fwd): the A operand is K-block-major, [B * 16][Cp / 64][TH][TW][64] (5-D
  // tensor map), and M is written slab-major, [16][B][Cout / 32][tiles per image][32], so that the transform passes, which
  // own (image, channel slab) pairs, stream contiguous memory (y_slab_tiles = tiles per image; 0 = plain [M, Cout])
  int a_kblock_major, y_slab_tiles;
  // MMA issue order inside a K block (split mode).  0: per k step hi*hi, hi*lo, lo*hi.  1 ("small terms first"): the
  // four hi*lo, then the four lo*hi, then the four hi*hi MMAs: tcgen05 truncates the fp32 accumulator after every MMA,
  // and the error of a truncation scales with the accumulator's magnitude at that moment -- the 2^-11-sized terms of
  // the first K block of a chunk are then accumulated while the accumulator is still small.
  int small_first;
  // direct stem input (conv_gemm_vr_kernel<true>, tsnet_stem_conv_fwd): the kw-folded halo tile is GENERATED in shared
  // memory by three producer warps from the raw NCHW network inputs (no tap source in HBM)
  const void* in_img;
  const void* in_lbl;
  int in_Cimg, in_Clbl, in_img_kind, in_lbl_kind;
  float in_mean[3], in_div, in_scale;
  // fused InstanceNorm epilogue (FUSED kernel variant; needs 8 tiles per image = one 8-CTA cluster per image)
  const float* f_residual;   // fp32 [B, H, W, Cout] or null
  float* f_act_out;          // fp32, channel window [f_act_c_off, +Cout) of f_act_C_total, or null
  uint16_t* f_taps_hi;       // destination tap source (hi / lo) or null
  uint16_t* f_taps_lo;
  int f_relu, f_mode, f_act_C_total, f_act_c_off, f_taps_Cp, f_taps_c_off, f_H, f_W;
  float f_act_scale, f_eps;
  int8_t tap_dy[TSNET_MAX_TAPS + 7], tap_dx[TSNET_MAX_TAPS + 7], tap_plane[TSNET_MAX_TAPS + 7];
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KB
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // hi + lo of both operands
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 6 ? 6 : (kSmemBudget / kStageBytes);
  // all 512 TMEM columns as chunk accumulators (2 / 4 / 8 for BLOCK_N 256 / 128 / 64): with narrow tiles the MMA pipe
  // runs several chunks ahead, so the accumulate warps' per-tile epilogue overlaps tensor work
  static constexpr int kTmemBufs = 512 / BLOCK_N;
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  // fused-IN variant: + per-warp partials [4][N] float2, per-CTA partials [2][N] double2 (read by peers over DSMEM),
  // (mean, rstd) [N] float2
  static constexpr int kFusedExtra = 4 * BLOCK_N * 8 + 2 * BLOCK_N * 16 + BLOCK_N * 8;
  static constexpr int kSmemBytesFused = kSmemBytes + kFusedExtra;
  static_assert(kSmemBytesFused <= 227 * 1024, "fused conv shared memory budget");
};

// ---- cluster helpers: sm100_prims.cuh ----
__device__ __forceinline__ double2 ld_cluster_double2(uint32_t cluster_addr) {
  double2 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Lane j of the warp ends with sum over the 32 lanes of v[j] (transpose-reduce butterfly, 31 shuffles).
__device__ __forceinline__ float warp_col_sums(float (&v)[32]) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = upper ? v[i] : v[i + o];
      const float keep = upper ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// Tile enumeration.  Plain variant: tile = blockIdx.x + k * gridDim.x.  FUSED variant: the 8 CTAs of a cluster take the
// 8 pixel tiles of ONE image for the same channel slab, so that the InstanceNorm statistics of that (image, slab)
// are complete inside the cluster: item = cluster_id + k * num_clusters, (img, n_tile) = item / % num_n_tiles.
template <bool FUSED>
__device__ __forceinline__ bool tile_at(const ConvGemmArgs& args, int k, int& m_tile, int& n_tile, int& plane) {
  if constexpr (FUSED) {
    const int item = static_cast<int>(blockIdx.x >> 3) + k * static_cast<int>(gridDim.x >> 3);
    if (item >= (args.num_m_tiles >> 3) * args.num_n_tiles) return false;
    const int img = item / args.num_n_tiles;
    n_tile = item - img * args.num_n_tiles;
    m_tile = img * 8 + static_cast<int>(cluster_ctarank());
    plane = 0;
    return true;
  } else {
    // order: plane, pixel tile, channel slab (fastest) -- the slabs of one A tile run concurrently, so the second read
    // of the tile is an L2 hit, and the weights of one plane stay L2-resident while the plane is swept
    const int tile = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
    if (tile >= args.batch_planes * args.num_m_tiles * args.num_n_tiles) return false;
    const int mt = tile / args.num_n_tiles;
    n_tile = tile - mt * args.num_n_tiles;
    plane = mt / args.num_m_tiles;
    m_tile = args.m_tile_begin + (mt - plane * args.num_m_tiles);
    return true;
  }
}

// 32 consecutive fp32 of one output row as four 256-bit stores (STG.E.256): one full 32 B sector per lane and request
// (16-byte pieces from 32 different rows touch 32 half-sectors per instruction)
__device__ __forceinline__ void store_row32(float* p, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p + j), "f"(v[j]), "f"(v[j + 1]),
                 "f"(v[j + 2]), "f"(v[j + 3]), "f"(v[j + 4]), "f"(v[j + 5]), "f"(v[j + 6]), "f"(v[j + 7])
                 : "memory");
}

template <int BLOCK_N, bool FUSED>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmArgs args) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + Cfg::kTmemBufs;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + Cfg::kTmemBufs);
  uint64_t* stats_ready = tmem_empty + Cfg::kTmemBufs + 1;  // fused variant: 8 arrivals (one per CTA of the cluster) per item
  // fused-variant scratch behind the barrier block
  uint8_t* fx = smem + Cfg::kStages * Cfg::kStageBytes + 256;
  float2* s_part = reinterpret_cast<float2*>(fx);                                  // [4][BLOCK_N]
  double2* s_cta = reinterpret_cast<double2*>(fx + 4 * BLOCK_N * 8);               // [2][BLOCK_N]
  float2* s_mr = reinterpret_cast<float2*>(fx + 4 * BLOCK_N * 8 + 2 * BLOCK_N * 16);  // [BLOCK_N]

  const int warp = threadIdx.x >> 5;
  const int num_kb = args.num_taps * args.kc_per_tap;

  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&args.a_hi);
    tma_prefetch_desc(&args.b_hi);
    if (args.split) {
      tma_prefetch_desc(&args.a_lo);
      tma_prefetch_desc(&args.b_lo);
    }
  }
  if (warp == 1 && lane_id() == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < Cfg::kTmemBufs; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kAccWarps);  // one arrive per accumulate warp
    }
    if constexpr (FUSED) mbar_init(stats_ready, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  if constexpr (FUSED) cluster_sync_all();  // every CTA's barriers exist before any remote arrive

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane_id() == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = args.split ? Cfg::kStageBytes : (Cfg::kABytes + Cfg::kBBytes);
      int m_tile, n_tile, plane;
      for (int k = 0; tile_at<FUSED>(args, k, m_tile, n_tile, plane); ++k) {
        const int img = m_tile / args.tiles_per_img;
        const int t = m_tile - img * args.tiles_per_img;
        const int ty = t / args.wtiles_per_row;
        const int tx = t - ty * args.wtiles_per_row;
        const int y0 = ty * args.rows_per_tile;
        const int x0 = tx * args.Wt;
        const int brow = plane * args.b_plane_rows + n_tile * BLOCK_N;
        for (int tap = 0; tap < args.num_taps; ++tap) {
          const int cy = y0 + args.tap_dy[tap];
          const int cx = x0 + args.tap_dx[tap];
          const int cn = img * args.planes + args.tap_plane[tap] + plane;
          for (int kc = 0; kc < args.kc_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * Cfg::kStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
            const int kb = tap * args.kc_per_tap + kc;
            if (args.a_kblock_major) tma_load_5d(st, &args.a_hi, &full_bar[stage], 0, cx, cy, kc, cn);
            else tma_load_4d(st, &args.a_hi, &full_bar[stage], kc * kBlockK, cx, cy, cn);
            tma_load_2d(st + 2 * Cfg::kABytes, &args.b_hi, &full_bar[stage], kb * kBlockK, brow);
            if (args.split) {
              if (args.a_kblock_major) tma_load_5d(st + Cfg::kABytes, &args.a_lo, &full_bar[stage], 0, cx, cy, kc, cn);
              else tma_load_4d(st + Cfg::kABytes, &args.a_lo, &full_bar[stage], kc * kBlockK, cx, cy, cn);
              tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &args.b_lo, &full_bar[stage], kb * kBlockK, brow);
            }
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {  // whole warp, converged; one elected lane issues (see sm100_prims.cuh)
      const uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N, args.fmt);
      int stage = 0;
      uint32_t phase = 0;
      int cc = 0;  // global chunk counter -> TMEM buffer + phase
      int m_tile, n_tile, plane;
      for (int kt = 0; tile_at<FUSED>(args, kt, m_tile, n_tile, plane); ++kt) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++cc) {
          const int buf = cc % Cfg::kTmemBufs;
          const uint32_t buf_phase = (cc / Cfg::kTmemBufs) & 1;
          mbar_wait(&tmem_empty[buf], buf_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
          const int kb1 = min(num_kb, kb0 + args.chunk_kb);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * Cfg::kStageBytes);
            const uint64_t a_hi = make_desc_kmajor_sw128(st);
            const uint64_t a_lo = make_desc_kmajor_sw128(st + Cfg::kABytes);
            const uint64_t b_hi = make_desc_kmajor_sw128(st + 2 * Cfg::kABytes);
            const uint64_t b_lo = make_desc_kmajor_sw128(st + 2 * Cfg::kABytes + Cfg::kBBytes);
            if (elect_one()) {  // one election per K block: the 12 MMAs are issued back to back by the leader
              if (args.split && args.small_first) {
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_f16(d_tmem, desc_advance_k(a_hi, k * kUmmaK * 2), desc_advance_k(b_lo, k * kUmmaK * 2), idesc,
                           ((kb - kb0) | k) != 0);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_f16(d_tmem, desc_advance_k(a_lo, k * kUmmaK * 2), desc_advance_k(b_hi, k * kUmmaK * 2), idesc, 1);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_f16(d_tmem, desc_advance_k(a_hi, k * kUmmaK * 2), desc_advance_k(b_hi, k * kUmmaK * 2), idesc, 1);
              } else {
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  const uint32_t off = k * kUmmaK * 2;
                  umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_hi, off), idesc, ((kb - kb0) | k) != 0);
                  if (args.split) {
                    umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_lo, off), idesc, 1);
                    umma_f16(d_tmem, desc_advance_k(a_lo, off), desc_advance_k(b_hi, off), idesc, 1);
                  }
                }
              }
            }
            __syncwarp();
            umma_commit_elect(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit_elect(&tmem_full[buf]);
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== accumulate + epilogue =====================
    constexpr int NC = BLOCK_N / 2;     // columns owned by one thread
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;   // which half of the BLOCK_N columns
    const int row = q * 32 + lane_id();
    int cc = 0;
    int m_tile, n_tile, plane;
    for (int kt = 0; tile_at<FUSED>(args, kt, m_tile, n_tile, plane); ++kt) {
      float acc[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[j] = 0.f;
      for (int kb0 = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++cc) {
        const int buf = cc % Cfg::kTmemBufs;
        const uint32_t buf_phase = (cc / Cfg::kTmemBufs) & 1;
        mbar_wait(&tmem_full[buf], buf_phase);
        tc_fence_after();
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BLOCK_N + half * NC;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 32) {
          float v[32];
          tmem_ld_32x32(t0 + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += v[j];  // fp32 round-to-nearest promotion
        }
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&tmem_empty[buf]);
      }
      const size_t gm = static_cast<size_t>(m_tile) * kBlockM + row;
      const float* arow = args.addend ? args.addend + (gm % args.addend_rows) * args.Cout : nullptr;
      if constexpr (FUSED) {
        // ================= fused epilogue: InstanceNorm (+ReLU, +residual) + tap building =================
        const int et = threadIdx.x - 128;            // 0..255 among the accumulate threads
        const int nb0 = n_tile * BLOCK_N;            // first channel of this slab
        const int buf = kt & 1;
        // (a) scale + bias (+ addend)
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 32) {
          const int n0 = nb0 + half * NC + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            acc[c0 + j] = fmaf(acc[c0 + j], args.out_scale, args.bias ? __ldg(args.bias + n0 + j) : 0.f);
          if (arow) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(arow + n0 + j));
              acc[c0 + j] += t4.x; acc[c0 + j + 1] += t4.y; acc[c0 + j + 2] += t4.z; acc[c0 + j + 3] += t4.w;
            }
          }
          // (b) per-column (sum, centred M2) over this warp's 32 pixels
          float t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = acc[c0 + j];
          const float colsum = warp_col_sums(t);
          const float mean_l = colsum * (1.f / 32.f);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dlt = acc[c0 + j] - __shfl_sync(0xffffffffu, mean_l, j);
            t[j] = dlt * dlt;
          }
          const float m2 = warp_col_sums(t);
          s_part[q * BLOCK_N + half * NC + c0 + lane_id()] = make_float2(colsum, m2);
        }
        epi_bar();
        // (c) CTA partial (128 pixels) per column, fixed order over the 4 lane quarters, fp64
        if (et < BLOCK_N) {
          double n = 0.0, mean = 0.0, m2 = 0.0;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const float2 pm = s_part[qq * BLOCK_N + et];
            const double nb = 32.0, mb = static_cast<double>(pm.x) * (1.0 / 32.0), nn = n + nb, dl = mb - mean;
            mean += dl * nb / nn;
            m2 += static_cast<double>(pm.y) + dl * dl * n * nb / nn;
            n = nn;
          }
          s_cta[buf * BLOCK_N + et] = make_double2(mean, m2);
        }
        epi_bar();
        // (d) publish to the cluster: one release-arrive on every CTA's stats_ready (count 8), then wait for ours
        if (et == 0) {
          const uint32_t local = smem_u32(stats_ready);
#pragma unroll
          for (uint32_t r = 0; r < 8; ++r) mbar_arrive_remote(map_to_cta(local, r));
        }
        mbar_wait_cluster(stats_ready, kt & 1);
        // (e) merge the 8 tiles of the image in rank order (identical in every CTA) -> mean, rstd
        if (et < BLOCK_N) {
          const uint32_t local = smem_u32(&s_cta[buf * BLOCK_N + et]);
          double n = 0.0, mean = 0.0, m2 = 0.0;
#pragma unroll
          for (uint32_t r = 0; r < 8; ++r) {
            const double2 pm = ld_cluster_double2(map_to_cta(local, r));
            const double nb = 128.0, nn = n + nb, dl = pm.x - mean;
            mean += dl * nb / nn;
            m2 += pm.y + dl * dl * n * nb / nn;
            n = nn;
          }
          const double var = m2 / n;  // biased, as nn.InstanceNorm2d
          s_mr[et] = make_float2(static_cast<float>(mean),
                                 static_cast<float>(1.0 / sqrt(var + static_cast<double>(args.f_eps))));
        }
        epi_bar();
        // (f) normalise, ReLU, residual, then write what the next layer reads
        const int tpi = m_tile & 7;
        const int img = m_tile >> 3;
        const int py = tpi * args.rows_per_tile + row / args.Wt;   // pixel of this thread inside the image
        const int px = row % args.Wt;
        const size_t pix = (static_cast<size_t>(img) * args.f_H + py) * args.f_W + px;
        const int ncol = nb0 + half * NC;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 mr = s_mr[half * NC + c0 + j];
            float v = (acc[c0 + j] - mr.x) * mr.y;
            if (args.f_relu) v = fmaxf(v, 0.f);
            acc[c0 + j] = v;
          }
          if (args.f_residual) {
            const float* rp = args.f_residual + pix * args.Cout + ncol + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(rp + j));
              acc[c0 + j] += t4.x; acc[c0 + j + 1] += t4.y; acc[c0 + j + 2] += t4.z; acc[c0 + j + 3] += t4.w;
            }
          }
          if (args.f_act_out) {
            float* op = args.f_act_out + pix * args.f_act_C_total + args.f_act_c_off + ncol + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(op + j) = make_float4(acc[c0 + j], acc[c0 + j + 1], acc[c0 + j + 2], acc[c0 + j + 3]);
          }
        }
        if (args.f_taps_hi) {
          // destination pixels: (py, px) itself, plus its mirror images in the reflect-pad ring (mode REFLECT1)
          int ys[2], xs[2], nys = 1, nxs = 1, Hd = args.f_H, Wd = args.f_W;
          ys[0] = py; xs[0] = px;
          if (args.f_mode == TSNET_TAPS_REFLECT1) {
            Hd += 2; Wd += 2;
            ys[0] = py + 1; xs[0] = px + 1;
            if (py == 1) ys[nys++] = 0;
            if (py == args.f_H - 2) ys[nys++] = args.f_H + 1;
            if (px == 1) xs[nxs++] = 0;
            if (px == args.f_W - 2) xs[nxs++] = args.f_W + 1;
          }
#pragma unroll
          for (int c0 = 0; c0 < NC; c0 += 8) {
            uint16_t h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split16(acc[c0 + j] * args.f_act_scale, args.fmt, h[j], l[j]);
            uint4 ph, pl;
            ph.x = h[0] | (uint32_t(h[1]) << 16); ph.y = h[2] | (uint32_t(h[3]) << 16);
            ph.z = h[4] | (uint32_t(h[5]) << 16); ph.w = h[6] | (uint32_t(h[7]) << 16);
            pl.x = l[0] | (uint32_t(l[1]) << 16); pl.y = l[2] | (uint32_t(l[3]) << 16);
            pl.z = l[4] | (uint32_t(l[5]) << 16); pl.w = l[6] | (uint32_t(l[7]) << 16);
            for (int a = 0; a < nys; ++a)
              for (int b2 = 0; b2 < nxs; ++b2) {
                const size_t d = ((static_cast<size_t>(img) * Hd + ys[a]) * Wd + xs[b2]) * args.f_taps_Cp +
                                 args.f_taps_c_off + ncol + c0;
                *reinterpret_cast<uint4*>(args.f_taps_hi + d) = ph;
                *reinterpret_cast<uint4*>(args.f_taps_lo + d) = pl;
              }
          }
        }
        continue;
      }
      // ---- tile epilogue: scale + bias, store, InstanceNorm partial statistics ----
      float* yrow = args.y + static_cast<size_t>(plane) * args.y_plane_stride + gm * args.Cout;
      // slab-major M (Winograd): [image][Cout / 32][tiles per image][32]
      const size_t simg = args.y_slab_tiles ? gm / args.y_slab_tiles : 0;
      const size_t stile = args.y_slab_tiles ? gm - simg * args.y_slab_tiles : 0;
      float* ybase = args.y + static_cast<size_t>(plane) * args.y_plane_stride +
                     simg * static_cast<size_t>(args.Cout) * args.y_slab_tiles + stile * 32;
      float* srow = args.stats ? args.stats + (static_cast<size_t>(m_tile) * 4 + q) * args.Cout * 2 : nullptr;
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 32) {
        const int n0 = n_tile * BLOCK_N + half * NC + c0;
        if (n0 < args.Cout) {  // padded output channels are skipped (warp-uniform)
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = fmaf(acc[c0 + j], args.out_scale, args.bias ? __ldg(args.bias + n0 + j) : 0.f);
          if (arow) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(arow + n0 + j));
              v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
            }
          }
          store_row32(args.y_slab_tiles ? ybase + static_cast<size_t>(n0 >> 5) * args.y_slab_tiles * 32 : yrow + n0, v);
          if (srow) {
            // per-column (sum, centred M2) over this warp's 32 pixels
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = v[j];
            const float colsum = warp_col_sums(t);  // lane j: sum of column j
            const float mean_l = colsum * (1.f / 32.f);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float mj = __shfl_sync(0xffffffffu, mean_l, j);
              const float dlt = v[j] - mj;
              t[j] = dlt * dlt;
            }
            const float m2 = warp_col_sums(t);
            *reinterpret_cast<float2*>(srow + (n0 + lane_id()) * 2) = make_float2(colsum, m2);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (FUSED) cluster_sync_all();  // no CTA leaves while a peer may still read its shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant of the BLOCK_N = 256 kernel (tcgen05.mma.cta_group::2; the default whenever the pixel tiles pair up).
// A CTA pair takes two adjacent pixel tiles (M = 256) of one 256-channel slab: each CTA TMA-loads its own A tile and only
// HALF of the weight tile (128 of the 256 rows); the leader issues the MMAs of both SMs.  Operand delivery per SM drops
// from 96 KB to 64 KB per K block (three 64 KB stages instead of two 96 KB ones) -- the same change took the
// correlation tile kernel from 80 % to 92 % tensor-pipe activity.  Chunked accumulation, promotion and the epilogue
// are those of conv_gemm_kernel<256, false>; per-element accumulation order is unchanged (bit-identical results).
// ------------------------------------------------------------------------------------------------
constexpr int kG2N = 256;
constexpr int kG2ABytes = kBlockM * kBlockK * 2;        // 16 KB: one A tile (hi or lo)
constexpr int kG2BBytes = (kG2N / 2) * kBlockK * 2;     // 16 KB: this CTA's half of the weight tile (hi or lo)
constexpr int kG2StageBytes = 2 * kG2ABytes + 2 * kG2BBytes;  // 64 KB
constexpr int kG2Stages = 3;
constexpr int kG2SmemBytes = kG2Stages * kG2StageBytes + 1024 + 256;

__global__ void __launch_bounds__(kGemmThreads, 1) conv_gemm2_kernel(const __grid_constant__ ConvGemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kG2Stages * kG2StageBytes);
  uint64_t* empty_bar = full_bar + kG2Stages;
  uint64_t* tmem_full = empty_bar + kG2Stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int num_kb = args.num_taps * args.kc_per_tap;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_pairs = args.num_m_tiles >> 1;
  const int num_items = args.batch_planes * m_pairs * args.num_n_tiles;  // order: plane, tile pair, channel slab

  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&args.a_hi);
    tma_prefetch_desc(&args.b_hi);
    if (args.split) {
      tma_prefetch_desc(&args.a_lo);
      tma_prefetch_desc(&args.b_lo);
    }
  }
  if (warp == 1 && lane_id() == 0) {
    for (int s = 0; s < kG2Stages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's copy is used: its producer arrives, both CTAs' TMA add their bytes
      mbar_init(&empty_bar[s], 1);  // armed in both CTAs by the leader's multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);               // multicast commit
      mbar_init(&tmem_empty[a], 2 * kAccWarps);  // leader's copy: the accumulate warps of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_base_smem, 2 * kG2N);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================== TMA producer (both CTAs) =====================
      if (lane_id() == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t stage_tx = 2u * (args.split ? kG2StageBytes : kG2StageBytes / 2);  // bytes of BOTH CTAs
        for (int item = pair_id; item < num_items; item += num_pairs) {
          const int mpa = item / args.num_n_tiles, n_tile = item - mpa * args.num_n_tiles;
          const int plane = mpa / m_pairs, mp = mpa - plane * m_pairs;
          const int m_tile = args.m_tile_begin + 2 * mp + static_cast<int>(rank);
          const int img = m_tile / args.tiles_per_img;
          const int t = m_tile - img * args.tiles_per_img;
          const int ty = t / args.wtiles_per_row;
          const int tx = t - ty * args.wtiles_per_row;
          const int y0 = ty * args.rows_per_tile;
          const int x0 = tx * args.Wt;
          const int brow = plane * args.b_plane_rows + n_tile * kG2N + static_cast<int>(rank) * (kG2N / 2);
          for (int tap = 0; tap < args.num_taps; ++tap) {
            const int cy = y0 + args.tap_dy[tap];
            const int cx = x0 + args.tap_dx[tap];
            const int cn = img * args.planes + args.tap_plane[tap] + plane;
            for (int kc = 0; kc < args.kc_per_tap; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* st = smem + stage * kG2StageBytes;
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
              const int kb = tap * args.kc_per_tap + kc;
              if (args.a_kblock_major) tma_load_5d_2sm(st, &args.a_hi, &full_bar[stage], 0, cx, cy, kc, cn);
              else tma_load_4d_2sm(st, &args.a_hi, &full_bar[stage], kc * kBlockK, cx, cy, cn);
              tma_load_2d_2sm(st + 2 * kG2ABytes, &args.b_hi, &full_bar[stage], kb * kBlockK, brow);
              if (args.split) {
                if (args.a_kblock_major)
                  tma_load_5d_2sm(st + kG2ABytes, &args.a_lo, &full_bar[stage], 0, cx, cy, kc, cn);
                else tma_load_4d_2sm(st + kG2ABytes, &args.a_lo, &full_bar[stage], kc * kBlockK, cx, cy, cn);
                tma_load_2d_2sm(st + 2 * kG2ABytes + kG2BBytes, &args.b_lo, &full_bar[stage], kb * kBlockK, brow);
              }
              if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      // ===================== MMA issuer (leader CTA: M = 256 across the pair) =====================
      const uint32_t idesc = make_idesc_f16(2 * kBlockM, kG2N, args.fmt);
      int stage = 0;
      uint32_t phase = 0;
      int cc = 0;
      for (int item = pair_id; item < num_items; item += num_pairs) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++cc) {
          const int buf = cc & 1;
          const uint32_t buf_phase = (cc >> 1) & 1;
          mbar_wait(&tmem_empty[buf], buf_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kG2N;
          const int kb1 = min(num_kb, kb0 + args.chunk_kb);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * kG2StageBytes);
            const uint64_t a_hi = make_desc_kmajor_sw128(st);
            const uint64_t a_lo = make_desc_kmajor_sw128(st + kG2ABytes);
            const uint64_t b_hi = make_desc_kmajor_sw128(st + 2 * kG2ABytes);
            const uint64_t b_lo = make_desc_kmajor_sw128(st + 2 * kG2ABytes + kG2BBytes);
            if (elect_one()) {
              if (args.split && args.small_first) {
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_f16_2sm(d_tmem, desc_advance_k(a_hi, k * kUmmaK * 2), desc_advance_k(b_lo, k * kUmmaK * 2), idesc,
                               ((kb - kb0) | k) != 0);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_f16_2sm(d_tmem, desc_advance_k(a_lo, k * kUmmaK * 2), desc_advance_k(b_hi, k * kUmmaK * 2), idesc, 1);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_f16_2sm(d_tmem, desc_advance_k(a_hi, k * kUmmaK * 2), desc_advance_k(b_hi, k * kUmmaK * 2), idesc, 1);
              } else {
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  const uint32_t off = k * kUmmaK * 2;
                  umma_f16_2sm(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_hi, off), idesc, ((kb - kb0) | k) != 0);
                  if (args.split) {
                    umma_f16_2sm(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_lo, off), idesc, 1);
                    umma_f16_2sm(d_tmem, desc_advance_k(a_lo, off), desc_advance_k(b_hi, off), idesc, 1);
                  }
                }
              }
              umma_commit_2sm(&empty_bar[stage]);  // frees the stage in BOTH CTAs once these MMAs have read it
            }
            __syncwarp();
            if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) umma_commit_2sm(&tmem_full[buf]);
          __syncwarp();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== accumulate + epilogue (each CTA: its own 128 pixels x 256 channels) =====================
    constexpr int NC = kG2N / 2;
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane_id();
    int cc = 0;
    for (int item = pair_id; item < num_items; item += num_pairs) {
      const int mpa = item / args.num_n_tiles, n_tile = item - mpa * args.num_n_tiles;
      const int plane = mpa / m_pairs, mp = mpa - plane * m_pairs;
      const int m_tile = args.m_tile_begin + 2 * mp + static_cast<int>(rank);
      float acc[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[j] = 0.f;
      for (int kb0 = 0; kb0 < num_kb; kb0 += args.chunk_kb, ++cc) {
        const int buf = cc & 1;
        const uint32_t buf_phase = (cc >> 1) & 1;
        mbar_wait(&tmem_full[buf], buf_phase);
        tc_fence_after();
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kG2N + half * NC;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 32) {
          float v[32];
          tmem_ld_32x32(t0 + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += v[j];  // fp32 round-to-nearest promotion
        }
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) {
          if (rank == 0) mbar_arrive(&tmem_empty[buf]);
          else mbar_arrive_remote_cta(map_to_cta(smem_u32(&tmem_empty[buf]), 0));
        }
      }
      const size_t gm = static_cast<size_t>(m_tile) * kBlockM + row;
      const float* arow = args.addend ? args.addend + (gm % args.addend_rows) * args.Cout : nullptr;
      float* yrow = args.y + static_cast<size_t>(plane) * args.y_plane_stride + gm * args.Cout;
      // slab-major M (Winograd): [image][Cout / 32][tiles per image][32]
      const size_t simg = args.y_slab_tiles ? gm / args.y_slab_tiles : 0;
      const size_t stile = args.y_slab_tiles ? gm - simg * args.y_slab_tiles : 0;
      float* ybase = args.y + static_cast<size_t>(plane) * args.y_plane_stride +
                     simg * static_cast<size_t>(args.Cout) * args.y_slab_tiles + stile * 32;
      float* srow = args.stats ? args.stats + (static_cast<size_t>(m_tile) * 4 + q) * args.Cout * 2 : nullptr;
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 32) {
        const int n0 = n_tile * kG2N + half * NC + c0;
        if (n0 < args.Cout) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = fmaf(acc[c0 + j], args.out_scale, args.bias ? __ldg(args.bias + n0 + j) : 0.f);
          if (arow) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(arow + n0 + j));
              v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
            }
          }
          store_row32(args.y_slab_tiles ? ybase + static_cast<size_t>(n0 >> 5) * args.y_slab_tiles * 32 : yrow + n0, v);
          if (srow) {
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = v[j];
            const float colsum = warp_col_sums(t);
            const float mean_l = colsum * (1.f / 32.f);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float mj = __shfl_sync(0xffffffffu, mean_l, j);
              const float dlt = v[j] - mj;
              t[j] = dlt * dlt;
            }
            const float m2 = warp_col_sums(t);
            *reinterpret_cast<float2*>(srow + (n0 + lane_id()) * 2) = make_float2(colsum, m2);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers or read its operand tiles
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * kG2N);
  }
}

static int launch_conv_gemm2(const ConvGemmArgs& a, cudaStream_t stream) {
  static int smem_attr[kMaxDevices] = {0};
  TSNET_CUDA_CHECK(ensure_dyn_smem(conv_gemm2_kernel, kG2SmemBytes, smem_attr));
  const int items = a.batch_planes * (a.num_m_tiles / 2) * a.num_n_tiles;
  const int pairs = items < num_sms() / 2 ? items : num_sms() / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kG2SmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ++launch_counter();
  TSNET_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm2_kernel, a));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Vertical-reuse variant for the kw-folded 7x7 stems (Cout = 64, one 64-channel K block per vertical tap).
// In the plain kernel every tap re-fetches its own 128-pixel A tile and the weights are re-fetched per tile: 338 KB from
// L2 for 84 MMAs of N = 64 (2688 tensor cycles) -- L2-bound at a third of the tensor rate (1.97 ms for 96 samples).
// Here the tile is 8 rows x 16 pixels: ONE TMA box of (8 + taps - 1) rows x 16 px x 64 ch is the halo of all vertical
// taps -- tap dy is the same shared-memory tile shifted by dy rows = dy * 2048 B, a multiple of the 1024 B swizzle
// atom, so it is just a different UMMA descriptor start address -- and the packed weights of all taps (16 KB per tap)
// stay resident in shared memory for the whole kernel.  L2 traffic per tile: 56 KB.
// Accumulation order per output element is identical to the plain kernel (chunks of chunk_kb taps promoted to
// registers), so the two kernels are bit-identical.
// ------------------------------------------------------------------------------------------------
constexpr int kVrRows = 8, kVrW = 16, kVrN = 64;
constexpr int kVrBTap = kVrN * kBlockK * 2;  // 8 KB: one tap of the packed weight (hi or lo)
// all 512 TMEM columns as 8 chunk accumulators: the MMA pipe runs up to two tiles ahead of the accumulate warps, so
// their per-tile epilogue (store + statistics) overlaps tensor work instead of stalling it after 2 chunks
constexpr int kVrTmemBufs = 8;

constexpr int kStemFold = 8;  // direct mode: channels per folded horizontal tap (Cin <= 8, zero padded): one 16-byte chunk

// Direct-input producer of conv_gemm_vr_kernel<true>: 96 threads (warps 0, 2, 3) build the (8 + 6) x 16 pixel halo tile of
// one output tile in the 128-byte-swizzled K-major layout a TMA box {64, 16, 14} of the materialised tap source would
// have produced.  Row R = r * 16 + px of the tile holds, for s = 0..6, the 8-channel vector of source pixel
// (reflect(y0 - 3 + r), reflect(x0 - 3 + px + s)) in 16-byte chunk s ^ (R & 7); chunk 7 ^ (R & 7) stays zero.  Every
// source pixel vector (torch.cat([img / 255, lbl]) + CoordConv channels, model/TSNet.py:107-125, :312; same rounding
// sequence as stem_taps_kernel) is computed once and stored to the <= 7 rows that use it.
// SPEC = true: the channel layout (NIMG image + NLBL label channels, input kinds IK / LK) is a compile-time constant, so
// the per-channel selection below folds away -- the generic version executed ~940 instructions per source pixel and made
// the producers, not the tensor pipe, the bound of the kernel.
template <bool SPEC, int NIMG, int NLBL, int IK, int LK>
__device__ __forceinline__ void stem_direct_fill(const ConvGemmArgs& args, uint8_t* a_hi, uint8_t* a_lo, int img, int y0,
                                                 int x0, int pt) {
  const int H = args.f_H, W = args.f_W;
  const size_t plane = static_cast<size_t>(H) * W;
  constexpr int SW = kVrW + 6, SR = kVrRows + 6;
  constexpr int NP = (SR * SW + 95) / 96;  // source pixels per producer thread (4)
  constexpr int NRAW = kStemFold - 3;      // image + label channels (<= 5)
  const int n_img = SPEC ? NIMG : args.in_Cimg, n_lbl = SPEC ? NLBL : args.in_Clbl, n_il = n_img + n_lbl;
  const int img_kind = SPEC ? IK : args.in_img_kind, lbl_kind = SPEC ? LK : args.in_lbl_kind;
  // ---- all global loads of this thread's pixels first (one exposed latency per tile instead of one per pixel)
  float raw[NP][NRAW];
  int ysv[NP], xsv[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int idx = pt + k * 96;
    const int r = idx / SW, sx = idx - r * SW;
    int ys = y0 - 3 + r, xs = x0 - 3 + sx;
    ys = ys < 0 ? -ys : ys; ys = ys >= H ? 2 * H - 2 - ys : ys;
    xs = xs < 0 ? -xs : xs; xs = xs >= W ? 2 * W - 2 - xs : xs;
    ysv[k] = ys; xsv[k] = xs;
    const size_t pix = static_cast<size_t>(ys) * W + xs;
    int cls = -1;
    if (idx < SR * SW && lbl_kind != 0)
      cls = static_cast<const uint8_t*>(args.in_lbl)[static_cast<size_t>(img) * plane + pix];
#pragma unroll
    for (int c = 0; c < NRAW; ++c) {
      float q = 0.f;
      if (idx < SR * SW) {
        if (c < n_img) {
          const size_t off = (static_cast<size_t>(img) * n_img + c) * plane + pix;
          if (img_kind == 0) q = static_cast<const float*>(args.in_img)[off];
          else q = static_cast<float>(static_cast<const uint8_t*>(args.in_img)[off]);
        } else if (c < n_il) {
          if (lbl_kind == 0)
            q = static_cast<const float*>(args.in_lbl)[(static_cast<size_t>(img) * n_lbl + (c - n_img)) * plane + pix];
          else
            q = cls == c - n_img ? 1.f : 0.f;
        }
      }
      raw[k][c] = q;
    }
  }
  // ---- channel vector, split, stores
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int idx = pt + k * 96;
    if (idx >= SR * SW) break;
    const int r = idx / SW, sx = idx - r * SW;
    // Encoder.coord_conv: t = idx / (n - 1); 2 t - 1; r = sqrt(x^2 + y^2), separate roundings
    const float yy = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(static_cast<float>(ysv[k]), static_cast<float>(H - 1))), -1.f);
    const float xx = __fadd_rn(__fmul_rn(2.f, __fdiv_rn(static_cast<float>(xsv[k]), static_cast<float>(W - 1))), -1.f);
    const float rr = __fsqrt_rn(__fadd_rn(__fmul_rn(xx, xx), __fmul_rn(yy, yy)));
    float v[kStemFold];
#pragma unroll
    for (int c = 0; c < kStemFold; ++c) {  // channel c of cat[img / div, lbl, x, y, r, 0...] (warp-uniform branches)
      float q = 0.f;
      if (c < NRAW && c < n_img) {
        float x = raw[k][c < NRAW ? c : 0];
        if (img_kind != 0) x = __fadd_rn(x, -args.in_mean[c < 3 ? c : 2]);
        q = __fdiv_rn(x, args.in_div);
      } else if (c < NRAW && c < n_il) {
        q = raw[k][c < NRAW ? c : 0];
      } else if (c == n_il) {
        q = xx;
      } else if (c == n_il + 1) {
        q = yy;
      } else if (c == n_il + 2) {
        q = rr;
      }
      v[c] = q;
    }
    uint16_t h[kStemFold], l[kStemFold];
#pragma unroll
    for (int j = 0; j < kStemFold; ++j) split16(v[j] * args.in_scale, args.fmt, h[j], l[j]);
    uint4 ph, pl;
    ph.x = h[0] | (uint32_t(h[1]) << 16); ph.y = h[2] | (uint32_t(h[3]) << 16);
    ph.z = h[4] | (uint32_t(h[5]) << 16); ph.w = h[6] | (uint32_t(h[7]) << 16);
    pl.x = l[0] | (uint32_t(l[1]) << 16); pl.y = l[2] | (uint32_t(l[3]) << 16);
    pl.z = l[4] | (uint32_t(l[5]) << 16); pl.w = l[6] | (uint32_t(l[7]) << 16);
#pragma unroll
    for (int sft = 0; sft < 7; ++sft) {
      const int px = sx - sft;
      if (px >= 0 && px < kVrW) {
        const int R = r * kVrW + px;
        const int off = R * 128 + ((sft ^ (R & 7)) << 4);
        *reinterpret_cast<uint4*>(a_hi + off) = ph;
        if (args.split) *reinterpret_cast<uint4*>(a_lo + off) = pl;
      }
    }
  }
}

template <bool DIRECT>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_gemm_vr_kernel(const __grid_constant__ ConvGemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int taps = args.num_taps;
  const int a_bytes = (kVrRows + taps - 1) * kVrW * 128;  // one halo tile (hi or lo); multiple of 2048
  uint8_t* b_sm = smem;                                  // [hi: taps x 8 KB][lo: taps x 8 KB]
  uint8_t* a_sm = smem + 2 * taps * kVrBTap;             // 2 buffers x [hi][lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_sm + 4 * a_bytes);
  uint64_t* b_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = bars + 3;
  uint64_t* tmem_full = bars + 5;
  uint64_t* tmem_empty = bars + 5 + kVrTmemBufs;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 5 + 2 * kVrTmemBufs);

  const int warp = threadIdx.x >> 5;
  const int num_tiles = args.num_m_tiles;
  const int tiles_x = args.wtiles_per_row;  // W / 16

  if (warp == 0 && lane_id() == 0) {
    if (!DIRECT) tma_prefetch_desc(&args.a_hi);
    tma_prefetch_desc(&args.b_hi);
    if (args.split) {
      if (!DIRECT) tma_prefetch_desc(&args.a_lo);
      tma_prefetch_desc(&args.b_lo);
    }
  }
  if (warp == 1 && lane_id() == 0) {
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], DIRECT ? 3 : 1);  // direct mode: one arrive per producer warp (0, 2, 3)
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kVrTmemBufs; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kAccWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_smem, kVrTmemBufs * kVrN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp < 4) {
    // control warp group: 40 registers per thread; the direct-input producers (loads + conversions) get 120 (4 x 32 x 120 + 8 x 32 x 192 = the 64512 registers the CTA was launched with) -- the
    // accumulate warps of this N = 64 kernel hold only 32 accumulators and make do with 192 instead of 224
    if constexpr (DIRECT) asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (DIRECT && warp != 1) {
      // ===================== direct-input producers (warps 0, 2, 3) =====================
      const int pt = (warp == 0 ? 0 : (warp - 1) * 32) + static_cast<int>(lane_id());
      if (warp == 0 && lane_id() == 0) {
        mbar_arrive_expect_tx(b_full, (args.split ? 2u : 1u) * taps * kVrBTap);
        for (int t = 0; t < taps; ++t) {
          tma_load_2d(b_sm + t * kVrBTap, &args.b_hi, b_full, t * kBlockK, 0);
          if (args.split) tma_load_2d(b_sm + (taps + t) * kVrBTap, &args.b_lo, b_full, t * kBlockK, 0);
        }
      }
      // the unused 8th chunk of every row (folded channels 56..63) is zero for the whole kernel, in all four tiles
      for (int i = pt; i < 4 * (kVrRows + 6) * kVrW; i += 96) {
        const int tile_i = i / ((kVrRows + 6) * kVrW), R = i - tile_i * ((kVrRows + 6) * kVrW);
        *reinterpret_cast<uint4*>(a_sm + tile_i * a_bytes + R * 128 + ((7 ^ (R & 7)) << 4)) = make_uint4(0, 0, 0, 0);
      }
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int m_tile = args.m_tile_begin + tile;
        const int img = m_tile / args.tiles_per_img;
        const int t = m_tile - img * args.tiles_per_img;
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        mbar_wait(&a_empty[buf], ph ^ 1);   // the MMAs of the tile that used this buffer have read it
        uint8_t* st = a_sm + buf * 2 * a_bytes;
        // (warp-uniform dispatch on the input description; the two face-configuration encoders are specialised)
        const int key = args.in_Cimg * 100 + args.in_Clbl * 10 + args.in_img_kind * 2 + args.in_lbl_kind;
        if (key == 320) stem_direct_fill<true, 3, 2, 0, 0>(args, st, st + a_bytes, img, ty * kVrRows, tx * kVrW, pt);
        else if (key == 323) stem_direct_fill<true, 3, 2, 1, 1>(args, st, st + a_bytes, img, ty * kVrRows, tx * kVrW, pt);
        else if (key == 20) stem_direct_fill<true, 0, 2, 0, 0>(args, st, st + a_bytes, img, ty * kVrRows, tx * kVrW, pt);
        else if (key == 21) stem_direct_fill<true, 0, 2, 0, 1>(args, st, st + a_bytes, img, ty * kVrRows, tx * kVrW, pt);
        else stem_direct_fill<false, 0, 0, 0, 0>(args, st, st + a_bytes, img, ty * kVrRows, tx * kVrW, pt);
        fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&a_full[buf]);
      }
    } else if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane_id() == 0) {
        mbar_arrive_expect_tx(b_full, (args.split ? 2u : 1u) * taps * kVrBTap);
        for (int t = 0; t < taps; ++t) {
          tma_load_2d(b_sm + t * kVrBTap, &args.b_hi, b_full, t * kBlockK, 0);
          if (args.split) tma_load_2d(b_sm + (taps + t) * kVrBTap, &args.b_lo, b_full, t * kBlockK, 0);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
          const int buf = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          const int m_tile = args.m_tile_begin + tile;
          const int img = m_tile / args.tiles_per_img;
          const int t = m_tile - img * args.tiles_per_img;
          const int ty = t / tiles_x, tx = t - ty * tiles_x;
          mbar_wait(&a_empty[buf], ph ^ 1);
          uint8_t* st = a_sm + buf * 2 * a_bytes;
          mbar_arrive_expect_tx(&a_full[buf], (args.split ? 2u : 1u) * a_bytes);
          const int cn = img * args.planes + args.tap_plane[0];
          tma_load_4d(st, &args.a_hi, &a_full[buf], 0, tx * kVrW + args.tap_dx[0], ty * kVrRows + args.tap_dy[0], cn);
          if (args.split)
            tma_load_4d(st + a_bytes, &args.a_lo, &a_full[buf], 0, tx * kVrW + args.tap_dx[0],
                        ty * kVrRows + args.tap_dy[0], cn);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      {  // whole warp, converged; one elected lane issues (see sm100_prims.cuh)
        const uint32_t idesc = make_idesc_f16(kBlockM, kVrN, args.fmt);
        mbar_wait(b_full, 0);
        tc_fence_after();
        const uint32_t b_base = smem_u32(b_sm);
        int it = 0, cc = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
          const int buf = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&a_full[buf], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_sm + buf * 2 * a_bytes);
          for (int t0 = 0; t0 < taps; t0 += args.chunk_kb, ++cc) {
            const int tb = cc % kVrTmemBufs;
            const uint32_t tb_phase = (cc / kVrTmemBufs) & 1;
            mbar_wait(&tmem_empty[tb], tb_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + tb * kVrN;
            const int t1 = min(taps, t0 + args.chunk_kb);
            for (int t = t0; t < t1; ++t) {
              const uint32_t a_off = static_cast<uint32_t>(args.tap_dy[t] - args.tap_dy[0]) * (kVrW * 128);
              const uint64_t a_hi = make_desc_kmajor_sw128(a_base + a_off);
              const uint64_t a_lo = make_desc_kmajor_sw128(a_base + a_bytes + a_off);
              const uint64_t b_hi = make_desc_kmajor_sw128(b_base + t * kVrBTap);
              const uint64_t b_lo = make_desc_kmajor_sw128(b_base + (taps + t) * kVrBTap);
              if (elect_one()) {  // one election per K block: the 12 MMAs are issued back to back by the leader
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  const uint32_t off = k * kUmmaK * 2;
                  umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_hi, off), idesc, ((t - t0) | k) != 0);
                  if (args.split) {
                    umma_f16(d_tmem, desc_advance_k(a_hi, off), desc_advance_k(b_lo, off), idesc, 1);
                    umma_f16(d_tmem, desc_advance_k(a_lo, off), desc_advance_k(b_hi, off), idesc, 1);
                  }
                }
              }
              __syncwarp();
            }
            umma_commit_elect(&tmem_full[tb]);
          }
          umma_commit_elect(&a_empty[buf]);  // the halo tile may be overwritten once these MMAs have read it
        }
      }
    }
  } else {
    if constexpr (DIRECT) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== accumulate + epilogue =====================
    constexpr int NC = kVrN / 2;
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane_id();
    int cc = 0;
    // the single channel slab never changes: keep its bias in registers (a per-tile __ldg was 20 % of the warps' time)
    float bias_r[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) bias_r[j] = (args.bias && half * NC + j < args.Cout) ? __ldg(args.bias + half * NC + j) : 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      float acc[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[j] = 0.f;
      for (int t0 = 0; t0 < taps; t0 += args.chunk_kb, ++cc) {
        const int tb = cc % kVrTmemBufs;
        const uint32_t tb_phase = (cc / kVrTmemBufs) & 1;
        mbar_wait(&tmem_full[tb], tb_phase);
        tc_fence_after();
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tb * kVrN + half * NC, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += v[j];  // fp32 round-to-nearest promotion
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&tmem_empty[tb]);
      }
      const int m_tile = args.m_tile_begin + tile;
      const int img = m_tile / args.tiles_per_img;
      const int t = m_tile - img * args.tiles_per_img;
      const int ty = t / tiles_x, tx = t - ty * tiles_x;
      const int y = ty * kVrRows + (row >> 4), x = tx * kVrW + (row & 15);
      const size_t gm = (static_cast<size_t>(img) * args.f_H + y) * args.f_W + x;
      const float* arow = args.addend ? args.addend + (gm % args.addend_rows) * args.Cout : nullptr;
      float* yrow = args.y + gm * args.Cout;
      float* srow = args.stats ? args.stats + (static_cast<size_t>(m_tile) * 4 + q) * args.Cout * 2 : nullptr;
      const int n0 = half * NC;
      if (n0 < args.Cout) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(acc[j], args.out_scale, bias_r[j]);
        if (arow) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(arow + n0 + j));
            v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
          }
        }
        store_row32(yrow + n0, v);
        if (srow) {
          float tt[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) tt[j] = v[j];
          const float colsum = warp_col_sums(tt);
          const float mean_l = colsum * (1.f / 32.f);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float mj = __shfl_sync(0xffffffffu, mean_l, j);
            const float dlt = v[j] - mj;
            tt[j] = dlt * dlt;
          }
          const float m2 = warp_col_sums(tt);
          *reinterpret_cast<float2*>(srow + (n0 + lane_id()) * 2) = make_float2(colsum, m2);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kVrTmemBufs * kVrN);
  }
}

static int launch_conv_gemm_vr(const ConvGemmArgs& a, cudaStream_t stream, bool direct = false) {
  const int a_bytes = (kVrRows + a.num_taps - 1) * kVrW * 128;
  const int smem_bytes = 2 * a.num_taps * kVrBTap + 4 * a_bytes + 1024 + 256;
  const int grid = a.num_m_tiles < num_sms() ? a.num_m_tiles : num_sms();
  if (direct) {
    static int smem_attr_d[kMaxDevices] = {0};
    TSNET_CUDA_CHECK(ensure_dyn_smem(conv_gemm_vr_kernel<true>, smem_bytes, smem_attr_d));
    conv_gemm_vr_kernel<true><<<grid, kGemmThreads, smem_bytes, stream>>>(a);
  } else {
    static int smem_attr[kMaxDevices] = {0};
    TSNET_CUDA_CHECK(ensure_dyn_smem(conv_gemm_vr_kernel<false>, smem_bytes, smem_attr));
    conv_gemm_vr_kernel<false><<<grid, kGemmThreads, smem_bytes, stream>>>(a);
  }
  TSNET_LAUNCH_CHECK();
  return 0;
}

template <int BLOCK_N>
static int launch_conv_gemm(const ConvGemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static int smem_attr[kMaxDevices] = {0};
  TSNET_CUDA_CHECK(ensure_dyn_smem(conv_gemm_kernel<BLOCK_N, false>, Cfg::kSmemBytes, smem_attr));
  const int tiles = a.batch_planes * a.num_m_tiles * a.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_gemm_kernel<BLOCK_N, false><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(a);
  TSNET_LAUNCH_CHECK();
  return 0;
}

// Fused-InstanceNorm variant: clusters of 8 CTAs, one cluster per (image, channel slab) item.
template <int BLOCK_N>
static int launch_conv_gemm_fused(const ConvGemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static int smem_attr[kMaxDevices] = {0};
  TSNET_CUDA_CHECK(ensure_dyn_smem(conv_gemm_kernel<BLOCK_N, true>, Cfg::kSmemBytesFused, smem_attr));
  const int items = (a.num_m_tiles / 8) * a.num_n_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytesFused;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 8;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // The persistent loop needs every cluster to be CO-RESIDENT: a cluster lives inside one GPC, and GPCs of 16/18/20 SMs
  // hold two 8-CTA clusters each (16 on a B200, not 148/8 = 18).  Ask the runtime instead of guessing.
  static int max_clusters_dev[kMaxDevices] = {0};
  int& max_clusters = max_clusters_dev[current_device()];
  if (max_clusters == 0) {
    cfg.gridDim = dim3(num_sms() / 8 * 8);
    int n = 0;
    TSNET_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<BLOCK_N, true>, &cfg));
    max_clusters = n > 0 ? n : 1;
  }
  const int clusters = items < max_clusters ? items : max_clusters;
  cfg.gridDim = dim3(clusters * 8);
  ++launch_counter();
  TSNET_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BLOCK_N, true>, a));
  return 0;
}

}  // namespace tsnet

using namespace tsnet;

extern "C" int tsnet_conv_gemm_fwd(const tsnet_conv_desc* d, const uint16_t* taps_hi, const uint16_t* taps_lo,
                                   const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, float* y_raw,
                                   float* stats_partial, void* stream) {
  TSNET_ARG_CHECK(d && taps_hi && w_hi && (y_raw || d->fuse_in), "conv_gemm: null argument");
  TSNET_ARG_CHECK(!d->split || (taps_lo && w_lo), "conv_gemm: split mode needs the lo operands");
  TSNET_ARG_CHECK(d->block_n == 64 || d->block_n == 128 || d->block_n == 256, "conv_gemm: block_n %d", d->block_n);
  TSNET_ARG_CHECK(d->Cp > 0 && d->Cp % 64 == 0, "conv_gemm: Cp %d must be a multiple of 64", d->Cp);
  TSNET_ARG_CHECK(d->Cout > 0 && d->Cout % 32 == 0, "conv_gemm: Cout %d must be a multiple of 32", d->Cout);
  TSNET_ARG_CHECK(d->Cout_pad % d->block_n == 0 && d->Cout_pad >= d->Cout, "conv_gemm: Cout_pad %d", d->Cout_pad);
  TSNET_ARG_CHECK(d->num_taps >= 1 && d->num_taps <= TSNET_MAX_TAPS, "conv_gemm: num_taps %d", d->num_taps);
  TSNET_ARG_CHECK((d->H * d->W) % kBlockM == 0, "conv_gemm: H*W = %d must be a multiple of 128", d->H * d->W);
  const int Wt = d->W < kBlockM ? d->W : kBlockM;
  TSNET_ARG_CHECK(kBlockM % Wt == 0 && d->W % Wt == 0, "conv_gemm: unsupported width %d", d->W);
  const int rows = kBlockM / Wt;
  TSNET_ARG_CHECK(d->H % rows == 0, "conv_gemm: H %d not a multiple of tile rows %d", d->H, rows);
  for (int t = 0; t < d->num_taps; ++t) {
    TSNET_ARG_CHECK(d->tap_plane[t] >= 0 && d->tap_plane[t] < d->planes, "conv_gemm: tap %d plane", t);
    TSNET_ARG_CHECK(d->tap_dy[t] >= 0 && d->tap_dy[t] + d->H <= d->Hp, "conv_gemm: tap %d dy out of range", t);
    TSNET_ARG_CHECK(d->tap_dx[t] >= 0 && d->tap_dx[t] + d->W <= d->Wp, "conv_gemm: tap %d dx out of range", t);
  }

  ConvGemmArgs a;
  memset(&a, 0, sizeof(a));
  {
    const uint64_t dims[4] = {(uint64_t)d->Cp, (uint64_t)d->Wp, (uint64_t)d->Hp, (uint64_t)d->B * d->planes};
    const uint64_t str[3] = {(uint64_t)d->Cp * 2, (uint64_t)d->Wp * d->Cp * 2, (uint64_t)d->Hp * d->Wp * d->Cp * 2};
    const uint32_t box[4] = {64, (uint32_t)Wt, (uint32_t)rows, 1};
    int r = encode_tmap_u16_sw128(&a.a_hi, taps_hi, 4, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.a_lo, taps_lo, 4, dims, str, box))) return r;
  }
  {
    const uint64_t K = (uint64_t)d->num_taps * d->Cp;
    const uint64_t dims[2] = {K, (uint64_t)d->Cout_pad};
    const uint64_t str[1] = {K * 2};
    const uint32_t box[2] = {64, (uint32_t)d->block_n};
    int r = encode_tmap_u16_sw128(&a.b_hi, w_hi, 2, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.b_lo, w_lo, 2, dims, str, box))) return r;
  }
  a.bias = bias;
  a.addend = d->addend;
  a.addend_rows = d->addend_rows;
  TSNET_ARG_CHECK(!d->addend || d->addend_rows > 0, "conv_gemm: addend needs addend_rows > 0");
  a.y = y_raw;
  a.stats = stats_partial;
  a.out_scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  a.tiles_per_img = d->H * d->W / kBlockM;
  a.num_m_tiles = d->B * a.tiles_per_img;
  a.num_n_tiles = d->Cout_pad / d->block_n;
  a.Wt = Wt;
  a.rows_per_tile = rows;
  a.wtiles_per_row = d->W / Wt;
  a.Cout = d->Cout;
  a.num_taps = d->num_taps;
  a.kc_per_tap = d->Cp / 64;
  a.planes = d->planes;
  a.split = d->split;
  a.fmt = d->fmt;
  a.chunk_kb = d->split ? 2 : 6;  // <= 24 accumulating MMAs per TMEM chunk before promotion to registers
  a.batch_planes = 1;
  a.b_plane_rows = 0;
  a.y_plane_stride = 0;
  for (int t = 0; t < d->num_taps; ++t) {
    a.tap_dy[t] = d->tap_dy[t];
    a.tap_dx[t] = d->tap_dx[t];
    a.tap_plane[t] = d->tap_plane[t];
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d->fuse_in) {
    TSNET_ARG_CHECK(a.tiles_per_img == 8, "conv_gemm: fuse_in needs H*W == 1024 (8 tiles per image), got %d tiles",
                    a.tiles_per_img);
    TSNET_ARG_CHECK(d->Cout % d->block_n == 0, "conv_gemm: fuse_in needs Cout %% block_n == 0");
    TSNET_ARG_CHECK(d->fuse_mode == TSNET_TAPS_SAME || d->fuse_mode == TSNET_TAPS_REFLECT1, "conv_gemm: fuse_mode %d",
                    d->fuse_mode);
    TSNET_ARG_CHECK((d->fuse_taps_hi == nullptr) == (d->fuse_taps_lo == nullptr), "conv_gemm: fuse taps hi/lo");
    TSNET_ARG_CHECK(d->fuse_taps_hi || d->fuse_act_out, "conv_gemm: fuse_in without any output");
    TSNET_ARG_CHECK(!d->fuse_taps_hi || (d->fuse_taps_Cp % 8 == 0 && d->fuse_taps_c_off % 8 == 0 &&
                                         d->fuse_taps_c_off + d->Cout <= d->fuse_taps_Cp),
                    "conv_gemm: fused tap window does not fit");
    TSNET_ARG_CHECK(!d->fuse_act_out || (d->fuse_act_c_off % 4 == 0 && d->fuse_act_c_off + d->Cout <= d->fuse_act_C_total),
                    "conv_gemm: fused act_out window does not fit");
    TSNET_ARG_CHECK(d->H >= 4 && d->W >= 4, "conv_gemm: fused reflect pad needs H, W >= 4");
    a.f_residual = d->fuse_residual;
    a.f_act_out = d->fuse_act_out;
    a.f_taps_hi = d->fuse_taps_hi;
    a.f_taps_lo = d->fuse_taps_lo;
    a.f_relu = d->fuse_relu;
    a.f_mode = d->fuse_mode;
    a.f_act_C_total = d->fuse_act_C_total;
    a.f_act_c_off = d->fuse_act_c_off;
    a.f_taps_Cp = d->fuse_taps_Cp;
    a.f_taps_c_off = d->fuse_taps_c_off;
    a.f_H = d->H;
    a.f_W = d->W;
    a.f_act_scale = d->fuse_act_scale == 0.f ? 1.f : d->fuse_act_scale;
    a.f_eps = d->fuse_eps == 0.f ? 1e-5f : d->fuse_eps;
    switch (d->block_n) {
      case 64: return launch_conv_gemm_fused<64>(a, s);
      case 128: return launch_conv_gemm_fused<128>(a, s);
      default: return launch_conv_gemm_fused<256>(a, s);
    }
  }
  // ---- vertical-reuse kernel for the kw-folded stems (see conv_gemm_vr_kernel)
  {
    const bool vr_enabled = (d->flags & TSNET_CONV_NO_VR) == 0;  // (tests compare the two kernels)
    bool vr = vr_enabled && !d->fuse_in && d->block_n == 64 && d->Cout_pad == 64 && d->Cp == 64 && d->num_taps > 1 &&
              d->W % kVrW == 0 && d->H % kVrRows == 0;
    for (int t = 0; vr && t < d->num_taps; ++t)
      vr = d->tap_dx[t] == d->tap_dx[0] && d->tap_plane[t] == d->tap_plane[0] && d->tap_dy[t] == d->tap_dy[0] + t;
    const int a_bytes = (kVrRows + d->num_taps - 1) * kVrW * 128;
    if (vr && 2 * d->num_taps * kVrBTap + 4 * a_bytes + 1280 <= 227 * 1024) {
      const uint64_t dims[4] = {(uint64_t)d->Cp, (uint64_t)d->Wp, (uint64_t)d->Hp, (uint64_t)d->B * d->planes};
      const uint64_t str[3] = {(uint64_t)d->Cp * 2, (uint64_t)d->Wp * d->Cp * 2, (uint64_t)d->Hp * d->Wp * d->Cp * 2};
      const uint32_t box[4] = {64, (uint32_t)kVrW, (uint32_t)(kVrRows + d->num_taps - 1), 1};
      int r = encode_tmap_u16_sw128(&a.a_hi, taps_hi, 4, dims, str, box);
      if (r) return r;
      if (d->split && (r = encode_tmap_u16_sw128(&a.a_lo, taps_lo, 4, dims, str, box))) return r;
      a.wtiles_per_row = d->W / kVrW;
      a.f_H = d->H;
      a.f_W = d->W;
      return launch_conv_gemm_vr(a, s);
    }
  }
  // ---- BLOCK_N = 256 launches use the 2-CTA kernel (cta_group::2: half the weight-tile traffic per SM) whenever the
  // pixel tiles pair up; TSNET_CONV_ONE_CTA forces the 1-CTA kernel (the two are bit-identical; tests compare them)
  auto launch_256 = [&](ConvGemmArgs& x) -> int {
    if ((d->flags & TSNET_CONV_ONE_CTA) == 0 && x.num_m_tiles >= 2 && x.num_m_tiles % 2 == 0 &&
        d->Cout_pad % kG2N == 0) {
      ConvGemmArgs y = x;
      const uint64_t K = (uint64_t)d->num_taps * d->Cp;
      const uint64_t dims[2] = {K, (uint64_t)d->Cout_pad};
      const uint64_t str[1] = {K * 2};
      const uint32_t box[2] = {64, (uint32_t)(kG2N / 2)};  // each CTA of the pair loads half of the weight tile
      int r = encode_tmap_u16_sw128(&y.b_hi, w_hi, 2, dims, str, box);
      if (r) return r;
      if (d->split && (r = encode_tmap_u16_sw128(&y.b_lo, w_lo, 2, dims, str, box))) return r;
      return launch_conv_gemm2(y, s);
    }
    return launch_conv_gemm<256>(x, s);
  };
  // ---- tail-wave split.  The persistent grid runs ceil(tiles / SMs) waves; when the last wave is mostly empty
  // (e.g. 1536 tiles on 148 SMs = 10.4 waves -> 11), the whole waves keep BLOCK_N and the remaining pixel tiles are
  // computed by a second launch with BLOCK_N / 2 (twice as many half-cost tiles): 10.5 waves instead of 11.
  const int sms = num_sms();
  const int tiles = a.num_m_tiles * a.num_n_tiles;
  const bool tail_split = (d->flags & TSNET_CONV_NO_TAIL_SPLIT) == 0;  // (tests compare the two launch plans)
  if (tail_split && d->block_n >= 128 && tiles > sms && tiles % sms != 0 && d->Cout_pad % (d->block_n / 2) == 0) {
    const int full_waves = tiles / sms;
    const int main_m = full_waves * sms / a.num_n_tiles;
    const int tail_m = a.num_m_tiles - main_m;
    const int tail_tiles = tail_m * a.num_n_tiles * 2;
    const float split_cost = static_cast<float>((main_m * a.num_n_tiles + sms - 1) / sms) +
                             0.525f * static_cast<float>((tail_tiles + sms - 1) / sms) + 0.05f;
    if (main_m > 0 && tail_m > 0 && split_cost < static_cast<float>((tiles + sms - 1) / sms)) {
      ConvGemmArgs t = a;
      const int bn2 = d->block_n / 2;
      {
        const uint64_t K = (uint64_t)d->num_taps * d->Cp;
        const uint64_t dims[2] = {K, (uint64_t)d->Cout_pad};
        const uint64_t str[1] = {K * 2};
        const uint32_t box[2] = {64, (uint32_t)bn2};
        int r = encode_tmap_u16_sw128(&t.b_hi, w_hi, 2, dims, str, box);
        if (r) return r;
        if (d->split && (r = encode_tmap_u16_sw128(&t.b_lo, w_lo, 2, dims, str, box))) return r;
      }
      t.m_tile_begin = main_m;
      t.num_m_tiles = tail_m;
      t.num_n_tiles = d->Cout_pad / bn2;
      a.num_m_tiles = main_m;
      int r = d->block_n == 256 ? launch_256(a) : launch_conv_gemm<128>(a, s);
      if (r) return r;
      return bn2 == 128 ? launch_conv_gemm<128>(t, s) : launch_conv_gemm<64>(t, s);
    }
  }
  switch (d->block_n) {
    case 64: return launch_conv_gemm<64>(a, s);
    case 128: return launch_conv_gemm<128>(a, s);
    default: return launch_256(a);
  }
}

// ------------------------------------------------------------------------------------------------
// Winograd F(2x2, 3x3): the 16 plane contractions  M[p] = V[p] (tiles x C) . U[p]^T (C x Cout)  as ONE batched launch
// of the BLOCK_N = 256 kernel (2-CTA pairs whenever the tile count is even).  V comes from tsnet_build_taps(mode
// TSNET_TAPS_WINO), U from tsnet_wino_weight_transform + tsnet_pack_conv_weight, M goes to tsnet_wino_output.
// ------------------------------------------------------------------------------------------------
extern "C" int tsnet_wino_gemm_fwd(const tsnet_wino_gemm_desc* d, const uint16_t* v_hi, const uint16_t* v_lo,
                                   const uint16_t* u_hi, const uint16_t* u_lo, float* m_out, void* stream) {
  TSNET_ARG_CHECK(d && v_hi && u_hi && m_out, "wino_gemm: null argument");
  TSNET_ARG_CHECK(!d->split || (v_lo && u_lo), "wino_gemm: split mode needs the lo operands");
  TSNET_ARG_CHECK(d->B >= 1 && d->TH >= 1 && d->TW >= 1, "wino_gemm: geometry");
  TSNET_ARG_CHECK(d->C > 0 && d->C % 64 == 0, "wino_gemm: C %d must be a multiple of 64", d->C);
  TSNET_ARG_CHECK(d->Cout > 0 && d->Cout % kG2N == 0, "wino_gemm: Cout %d must be a multiple of %d", d->Cout, kG2N);
  const int tiles = d->TH * d->TW;
  TSNET_ARG_CHECK(tiles % kBlockM == 0, "wino_gemm: TH*TW = %d must be a multiple of 128", tiles);
  const int Wt = d->TW < kBlockM ? d->TW : kBlockM;
  TSNET_ARG_CHECK(kBlockM % Wt == 0 && d->TW % Wt == 0, "wino_gemm: unsupported tile-row width %d", d->TW);
  const int rows = kBlockM / Wt;
  TSNET_ARG_CHECK(d->TH % rows == 0, "wino_gemm: TH %d not a multiple of %d", d->TH, rows);

  ConvGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.tiles_per_img = tiles / kBlockM;
  a.num_m_tiles = d->B * a.tiles_per_img;
  // tile width: 256 (CTA pairs) whenever that gives every SM work; small batches (the demos' one frame per forward:
  // 16 planes x 2 pixel tiles x Cout/256 slabs = 32 .. 64 items) fall back to 128- / 64-wide 1-CTA tiles.  The
  // accumulation order per element does not depend on the tile width, so results are bit-identical.
  int bn = kG2N;
  while (bn > 64 && 16 * a.num_m_tiles * (d->Cout / bn) < num_sms()) bn >>= 1;
  const bool two_cta = bn == kG2N && (d->flags & TSNET_CONV_ONE_CTA) == 0 && a.num_m_tiles % 2 == 0;
  {  // V, K-block-major: [B * 16][C / 64][TH][TW][64]
    const uint64_t dims[5] = {64, (uint64_t)d->TW, (uint64_t)d->TH, (uint64_t)d->C / 64, (uint64_t)d->B * 16};
    const uint64_t str[4] = {128, (uint64_t)d->TW * 128, (uint64_t)tiles * 128, (uint64_t)(d->C / 64) * tiles * 128};
    const uint32_t box[5] = {64, (uint32_t)Wt, (uint32_t)rows, 1, 1};
    int r = encode_tmap_u16_sw128(&a.a_hi, v_hi, 5, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.a_lo, v_lo, 5, dims, str, box))) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d->C, (uint64_t)16 * d->Cout};
    const uint64_t str[1] = {(uint64_t)d->C * 2};
    const uint32_t box[2] = {64, two_cta ? (uint32_t)(kG2N / 2) : (uint32_t)bn};
    int r = encode_tmap_u16_sw128(&a.b_hi, u_hi, 2, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.b_lo, u_lo, 2, dims, str, box))) return r;
  }
  a.y = m_out;
  a.out_scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  a.num_n_tiles = d->Cout / bn;
  a.Wt = Wt;
  a.rows_per_tile = rows;
  a.wtiles_per_row = d->TW / Wt;
  a.Cout = d->Cout;
  a.num_taps = 1;
  a.kc_per_tap = d->C / 64;
  a.planes = 16;
  a.split = d->split;
  a.fmt = d->fmt;
  a.chunk_kb = d->chunk_kb > 0 ? d->chunk_kb : (d->split ? 2 : 6);
  a.batch_planes = 16;
  a.b_plane_rows = d->Cout;
  a.y_plane_stride = static_cast<long long>(d->B) * tiles * d->Cout;
  a.a_kblock_major = 1;
  a.y_slab_tiles = tiles;
  a.small_first = (d->flags & TSNET_CONV_SMALL_FIRST) ? 1 : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (two_cta) return launch_conv_gemm2(a, s);
  return bn == 256 ? launch_conv_gemm<256>(a, s) : (bn == 128 ? launch_conv_gemm<128>(a, s) : launch_conv_gemm<64>(a, s));
}

// ------------------------------------------------------------------------------------------------
// Encoder stem without a materialised operand: ReflectionPad2d(3) + Conv2d(Cin, 64, 7) (model/TSNet.py:66) straight
// from the raw NCHW network inputs.  The kw-folded halo tile of every 8 x 16 output tile is built in shared memory by
// producer warps (conv_gemm_vr_kernel<true>): torch.cat([img / 255, lbl]) (:312), the uint8 -> float / mean-subtract /
// one-hot staging of the datasets, and Encoder.coord_conv (:107-125) never exist in HBM.
// ------------------------------------------------------------------------------------------------
extern "C" int tsnet_stem_conv_fwd(const tsnet_stem_conv_desc* d, const void* img_nchw, const void* lbl,
                                   const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, float* y_raw,
                                   float* stats_partial, void* stream) {
  TSNET_ARG_CHECK(d && lbl && w_hi && y_raw, "stem_conv: null argument");
  TSNET_ARG_CHECK(!d->split || w_lo, "stem_conv: split mode needs the lo weights");
  TSNET_ARG_CHECK((img_nchw != nullptr) == (d->Cimg > 0), "stem_conv: img pointer / Cimg mismatch");
  TSNET_ARG_CHECK(d->Cimg >= 0 && d->Clbl >= 1 && d->Cimg + d->Clbl + 3 <= kStemFold,
                  "stem_conv: Cimg + Clbl + 3 = %d channels do not fit the %d-channel folded tap (use tsnet_stem_taps + "
                  "tsnet_conv_gemm_fwd)", d->Cimg + d->Clbl + 3, kStemFold);
  TSNET_ARG_CHECK(d->img_kind == 0 || (d->img_kind == 1 && d->Cimg == 3), "stem_conv: img_kind %d", d->img_kind);
  TSNET_ARG_CHECK(d->lbl_kind == 0 || d->lbl_kind == 1, "stem_conv: lbl_kind %d", d->lbl_kind);
  TSNET_ARG_CHECK(d->Cout == kVrN, "stem_conv: Cout %d (the stem has %d output channels)", d->Cout, kVrN);
  TSNET_ARG_CHECK(d->H % kVrRows == 0 && d->W % kVrW == 0 && d->H >= 4 && d->W >= 4,
                  "stem_conv: H %d / W %d must be multiples of %d / %d", d->H, d->W, kVrRows, kVrW);
  ConvGemmArgs a;
  memset(&a, 0, sizeof(a));
  {  // packed weight [64, 7 * 64]: column r * 64 + s * 8 + c (tsnet_pack_conv_weight with fold_kw = 8)
    const uint64_t K = 7ull * 64;
    const uint64_t dims[2] = {K, (uint64_t)kVrN};
    const uint64_t str[1] = {K * 2};
    const uint32_t box[2] = {64, (uint32_t)kVrN};
    int r = encode_tmap_u16_sw128(&a.b_hi, w_hi, 2, dims, str, box);
    if (r) return r;
    if (d->split && (r = encode_tmap_u16_sw128(&a.b_lo, w_lo, 2, dims, str, box))) return r;
  }
  a.bias = bias;
  a.y = y_raw;
  a.stats = stats_partial;
  a.out_scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  a.tiles_per_img = (d->H / kVrRows) * (d->W / kVrW);
  a.num_m_tiles = d->B * a.tiles_per_img;
  a.num_n_tiles = 1;
  a.wtiles_per_row = d->W / kVrW;
  a.Cout = d->Cout;
  a.num_taps = 7;
  a.kc_per_tap = 1;
  a.planes = 1;
  a.split = d->split;
  a.fmt = d->fmt;
  a.chunk_kb = d->split ? 2 : 6;
  a.batch_planes = 1;
  for (int t = 0; t < 7; ++t) a.tap_dy[t] = static_cast<int8_t>(t);
  a.f_H = d->H;
  a.f_W = d->W;
  a.in_img = img_nchw;
  a.in_lbl = lbl;
  a.in_Cimg = d->Cimg; a.in_Clbl = d->Clbl; a.in_img_kind = d->img_kind; a.in_lbl_kind = d->lbl_kind;
  a.in_mean[0] = d->img_mean[0]; a.in_mean[1] = d->img_mean[1]; a.in_mean[2] = d->img_mean[2];
  a.in_div = d->img_div == 0.f ? 1.f : d->img_div;
  a.in_scale = d->act_scale == 0.f ? 1.f : d->act_scale;
  return launch_conv_gemm_vr(a, static_cast<cudaStream_t>(stream), true);
}
