// Train-mode branches INSIDE TSNet.forward() (forward-only, no backward):
//   model/TSNet.py:327-331   reference statistics of the target image (per sample and channel, unbiased std)
//   model/TSNet.py:372-390   image-space warp: F.unfold(src_img, 8, 8) -> grid_sample with the 32 x 32 warp grid ->
//                            F.fold, i.e. every 8 x 8 pixel patch of the target is the bilinear mix of the four source
//                            patches around the expected source coordinate; per-image colour re-normalisation;
//                            warp loss 10 * L1 against the target image
//   model/TSNet.py:402-405   alignment loss 1 - mean cosine similarity of the two branch means (face variant)
//   model/TSNet_pose.py:395-396  foreground compositing of the warped image (pose variant)
// All reductions use fp64 accumulators with a fixed order (bit-reproducible, no atomics).
#include "host_util.h"
#include "../../include/tsnet_b200.h"
#include <math.h>

namespace tsnet {

constexpr int kPlaneThreads = 1024;

// fixed-order block reduction of two doubles (blockDim.x = 1024)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh /* [64] */) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[warp] = a;
    sh[32 + warp] = b;
  }
  __syncthreads();
  if (warp == 0) {
    a = sh[lane];
    b = sh[32 + lane];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      sh[0] = a;
      sh[32] = b;
    }
  }
  __syncthreads();
  a = sh[0];
  b = sh[32];
  __syncthreads();
}

// one block per plane: (mean, unbiased std) of x / div over `n` contiguous elements  (tensor.mean / tensor.std)
__global__ void __launch_bounds__(kPlaneThreads) plane_stats_kernel(const float* __restrict__ x, int n, float div,
                                                                    float* __restrict__ out) {
  __shared__ double sh[64];
  const float* p = x + static_cast<size_t>(blockIdx.x) * n;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < n; i += kPlaneThreads) {
    const double v = static_cast<double>(div == 1.f ? p[i] : __fdiv_rn(p[i], div));
    s += v;
    q += v * v;
  }
  block_sum2(s, q, sh);
  if (threadIdx.x == 0) {
    const double mean = s / n;
    const double var = fmax((q - s * mean) / (n - 1), 0.0);
    out[blockIdx.x * 2 + 0] = static_cast<float>(mean);
    out[blockIdx.x * 2 + 1] = static_cast<float>(sqrt(var));
  }
}

// image-space warp.  grid = (h, B, n_src); block = W threads x 8 rows... one block handles one row of cells (8 image
// rows x W columns x 3 channels); thread = (pixel column X, row ky of the cell), loops over the 3 channels.
struct ImgWarpArgs {
  const float* src_img[12];  // raw NCHW [B, 3, H, W]
  float div[12];             // 255 or 1 (use_prev)
  const float* grids;        // [n, B, h, w, 2]
  float* out;                // [n, B, 3, H, W]
  int B, H, W, h, w, n_src;
};

__global__ void __launch_bounds__(256) image_warp_kernel(const ImgWarpArgs a) {
  const int ty = blockIdx.x, b = blockIdx.y, i = blockIdx.z;
  const int down_y = a.H / a.h, down_x = a.W / a.w;
  const float* img = a.src_img[i];
  const float div = a.div[i];
  const size_t plane = static_cast<size_t>(a.H) * a.W;
  for (int idx = threadIdx.x; idx < down_y * a.W; idx += blockDim.x) {
    const int ky = idx / a.W, X = idx - ky * a.W;
    const int tx = X / down_x, kx = X - tx * down_x;
    const float2 g = *reinterpret_cast<const float2*>(
        a.grids + (((static_cast<size_t>(i) * a.B + b) * a.h + ty) * a.w + tx) * 2);
    // F.grid_sample(bilinear, zeros, align_corners=False) on the [B, 3*64, h, w] unfolded image
    const float ix = ((g.x + 1.f) * a.w - 1.f) * 0.5f;
    const float iy = ((g.y + 1.f) * a.h - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
    float wts[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};  // nw, ne, sw, se
    size_t off[4];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int xx = x0 + (tap & 1), yy = y0 + (tap >> 1);
      if (!(xx >= 0 && xx < a.w && yy >= 0 && yy < a.h)) wts[tap] = 0.f;
      const int xc = min(max(xx, 0), a.w - 1), yc = min(max(yy, 0), a.h - 1);
      off[tap] = static_cast<size_t>(yc * down_y + ky) * a.W + xc * down_x + kx;
    }
    const int Y = ty * down_y + ky;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = img + (static_cast<size_t>(b) * 3 + c) * plane;
      float v = 0.f;
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        const float s = div == 1.f ? pl[off[tap]] : __fdiv_rn(pl[off[tap]], div);
        v += s * wts[tap];
      }
      a.out[((static_cast<size_t>(i) * a.B + b) * 3 + c) * plane + static_cast<size_t>(Y) * a.W + X] = v;
    }
  }
}

// one block per (source, sample, channel) plane: out = (x - gen_mean) / gen_std * ref_std + ref_mean, optional pose
// compositing, in place; l1[plane] = sum |out - tar / tar_div|
__global__ void __launch_bounds__(kPlaneThreads) warp_renorm_l1_kernel(float* __restrict__ warp,
                                                                       const float* __restrict__ gen_stats,
                                                                       const float* __restrict__ ref_stats,
                                                                       const float* __restrict__ tar, float tar_div,
                                                                       int B, int H, int W, int fore_x0, int fore_x1,
                                                                       float fill0, float fill1, float fill2,
                                                                       double* __restrict__ l1) {
  __shared__ double sh[64];
  const int pl = blockIdx.x;            // (i * B + b) * 3 + c
  const int c = pl % 3, b = (pl / 3) % B;
  const int n = H * W;
  const float gm = gen_stats[pl * 2], gs = gen_stats[pl * 2 + 1];
  const float rm = ref_stats[(b * 3 + c) * 2], rs = ref_stats[(b * 3 + c) * 2 + 1];
  const float fill = c == 0 ? fill0 : (c == 1 ? fill1 : fill2);
  float* p = warp + static_cast<size_t>(pl) * n;
  const float* t = tar + static_cast<size_t>(b * 3 + c) * n;
  double s = 0.0, unused = 0.0;
  for (int i = threadIdx.x; i < n; i += kPlaneThreads) {
    float v = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p[i], gm), gs), rs), rm);
    if (fore_x1 > fore_x0) {
      const int x = i % W;
      const float f = (x >= fore_x0 && x < fore_x1) ? 1.f : 0.f;
      v = __fadd_rn(__fmul_rn(v, f), __fmul_rn(fill, 1.f - f));
    }
    p[i] = v;
    s += fabs(static_cast<double>(v) - static_cast<double>(__fdiv_rn(t[i], tar_div)));
  }
  block_sum2(s, unused, sh);
  if (threadIdx.x == 0) l1[pl] = s;
}

// cosine similarity of two fp32 NHWC maps along C (F.cosine_similarity, eps = 1e-8): one warp per position, one block
// of 8 positions writes its partial sum (fixed order)
__global__ void __launch_bounds__(256) cosine_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             int rows, int C, double* __restrict__ partial) {
  __shared__ double sh[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  double cs = 0.0;
  if (r < rows) {
    const float* px = x + static_cast<size_t>(r) * C;
    const float* py = y + static_cast<size_t>(r) * C;
    float dot = 0.f, nx = 0.f, ny = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 u = *reinterpret_cast<const float4*>(px + c), v = *reinterpret_cast<const float4*>(py + c);
      dot += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
      nx += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
      ny += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
      nx += __shfl_xor_sync(0xffffffffu, nx, o);
      ny += __shfl_xor_sync(0xffffffffu, ny, o);
    }
    // ATen: w12 / sqrt(clamp_min(w1 * w2, eps^2))
    cs = static_cast<double>(dot / sqrtf(fmaxf(nx * ny, 1e-16f)));
  }
  if (lane == 0) sh[warp] = cs;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += sh[k];
    partial[blockIdx.x] = s;
  }
}

// out[0] = loss_warp = sum_i 10 * mean |warp_i - tar| ; out[1] = loss_align = 1 - mean cos (if cos_n > 0)
__global__ void __launch_bounds__(kPlaneThreads) finalize_losses_kernel(const double* __restrict__ l1, int n_src,
                                                                        int planes_per_src, double elems_per_src,
                                                                        const double* __restrict__ cosp, int cos_blocks,
                                                                        double cos_n, float* __restrict__ out) {
  __shared__ double sh[64];
  double lw = 0.0;
  for (int i = 0; i < n_src; ++i) {
    double s = 0.0, z = 0.0;
    for (int k = threadIdx.x; k < planes_per_src; k += kPlaneThreads) s += l1[i * planes_per_src + k];
    block_sum2(s, z, sh);
    lw += 10.0 * static_cast<double>(static_cast<float>(s / elems_per_src));  // each l1_loss is an fp32 tensor
  }
  double cs = 0.0, z = 0.0;
  for (int k = threadIdx.x; k < cos_blocks; k += kPlaneThreads) cs += cosp[k];
  block_sum2(cs, z, sh);
  if (threadIdx.x == 0) {
    out[0] = static_cast<float>(lw);
    out[1] = cos_n > 0 ? 1.f - static_cast<float>(cs / cos_n) : 0.f;
  }
}

}  // namespace tsnet

using namespace tsnet;

extern "C" int tsnet_plane_stats(const float* x, int planes, int n, float div, float* mean_std, void* stream) {
  TSNET_ARG_CHECK(x && mean_std && planes > 0 && n > 1, "plane_stats: bad argument");
  plane_stats_kernel<<<planes, kPlaneThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, n, div == 0.f ? 1.f : div,
                                                                                     mean_std);
  TSNET_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t tsnet_train_extras_workspace_bytes(int B, int n_src, int h, int w) {
  const size_t planes = static_cast<size_t>(n_src) * B * 3;
  const size_t cos_blocks = (static_cast<size_t>(B) * h * w + 7) / 8;
  return planes * 2 * sizeof(float) + static_cast<size_t>(B) * 3 * 2 * sizeof(float) + 256 +
         (planes + cos_blocks) * sizeof(double) + 256;
}

extern "C" int tsnet_train_extras_fwd(const float* const* src_img, const float* src_div, int n_src,
                                      const float* tar_img, float tar_div, const float* grids, int B, int H, int W,
                                      int h, int w, const float* pg_mean, const float* sg_mean, int C, int fore_x0,
                                      int fore_x1, const float* fill3, float* warp_out, float* losses2,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  TSNET_ARG_CHECK(src_img && src_div && tar_img && grids && warp_out && losses2 && workspace,
                  "train_extras: null argument");
  TSNET_ARG_CHECK(n_src >= 1 && n_src <= 12, "train_extras: n_src %d", n_src);
  TSNET_ARG_CHECK(H % h == 0 && W % w == 0 && H / h == W / w, "train_extras: %dx%d image vs %dx%d grid", H, W, h, w);
  TSNET_ARG_CHECK((pg_mean == nullptr) == (sg_mean == nullptr), "train_extras: pg/sg means must come together");
  TSNET_ARG_CHECK(!pg_mean || C % 4 == 0, "train_extras: C %d", C);
  TSNET_ARG_CHECK(fore_x1 <= fore_x0 || fill3, "train_extras: compositing needs fill3 (host pointer to 3 floats)");
  TSNET_ARG_CHECK(workspace_bytes >= tsnet_train_extras_workspace_bytes(B, n_src, h, w), "train_extras: workspace");
  TSNET_ARG_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "train_extras: workspace must be 256 B aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int planes = n_src * B * 3;
  const int cos_blocks = pg_mean ? (B * h * w + 7) / 8 : 0;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* gen_stats = reinterpret_cast<float*>(ws);
  float* ref_stats = gen_stats + static_cast<size_t>(planes) * 2;
  size_t off = (static_cast<size_t>(planes) * 2 + static_cast<size_t>(B) * 3 * 2) * sizeof(float);
  off = (off + 255) & ~size_t(255);
  double* l1 = reinterpret_cast<double*>(ws + off);
  double* cosp = l1 + planes;

  ImgWarpArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < n_src; ++i) {
    TSNET_ARG_CHECK(src_img[i], "train_extras: null source %d", i);
    a.src_img[i] = src_img[i];
    a.div[i] = src_div[i] == 0.f ? 1.f : src_div[i];
  }
  a.grids = grids; a.out = warp_out; a.B = B; a.H = H; a.W = W; a.h = h; a.w = w; a.n_src = n_src;
  image_warp_kernel<<<dim3(h, B, n_src), 256, 0, s>>>(a);
  TSNET_LAUNCH_CHECK();
  plane_stats_kernel<<<planes, kPlaneThreads, 0, s>>>(warp_out, H * W, 1.f, gen_stats);
  TSNET_LAUNCH_CHECK();
  plane_stats_kernel<<<B * 3, kPlaneThreads, 0, s>>>(tar_img, H * W, tar_div == 0.f ? 1.f : tar_div, ref_stats);
  TSNET_LAUNCH_CHECK();
  const float f0 = fill3 ? fill3[0] : 0.f, f1 = fill3 ? fill3[1] : 0.f, f2 = fill3 ? fill3[2] : 0.f;
  warp_renorm_l1_kernel<<<planes, kPlaneThreads, 0, s>>>(warp_out, gen_stats, ref_stats, tar_img,
                                                         tar_div == 0.f ? 1.f : tar_div, B, H, W, fore_x0, fore_x1, f0,
                                                         f1, f2, l1);
  TSNET_LAUNCH_CHECK();
  if (pg_mean) {
    cosine_partial_kernel<<<cos_blocks, 256, 0, s>>>(pg_mean, sg_mean, B * h * w, C, cosp);
    TSNET_LAUNCH_CHECK();
  }
  finalize_losses_kernel<<<1, kPlaneThreads, 0, s>>>(l1, n_src, B * 3, static_cast<double>(B) * 3 * H * W, cosp,
                                                     cos_blocks, pg_mean ? static_cast<double>(B) * h * w : 0.0,
                                                     losses2);
  TSNET_LAUNCH_CHECK();
  return 0;
}
