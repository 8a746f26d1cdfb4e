/*
 * libtsnet_sm100.so -- C ABI of the B200-native TS-Net forward hot path.
 *
 * The reference (nihaomiao/WACV23_TSNet) has no FFI: its hot path is the body of
 * TSNet.forward() (model/TSNet.py:309-407, model/TSNet_pose.py:325-417) expressed as stock torch ops.
 * Each entry point below replaces a group of those ops; the reference lines are cited per function.
 * The Python class wacv23_tsnet_b200.model.TSNet.TSNet (same surface as the reference class) binds
 * these through ctypes (wacv23_tsnet_b200/lib.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a BORROWED DEVICE pointer (the caller -- PyTorch -- owns the memory and keeps
 *     it alive until the stream has drained); descriptors (`*_desc`) are host structs;
 *   - functions only ENQUEUE work on `stream` (a cudaStream_t passed as void*); they never
 *     synchronise and never allocate: all workspaces are caller-provided;
 *   - return 0 on success, <0 for argument errors, >0 = cudaError_t; the message is available from
 *     tsnet_last_error() (thread-local);
 *   - there is NO CPU fallback: without an sm_100a device the launch fails and the error is returned.
 *
 * Data layouts in HBM
 *   activation ("act")   fp32 NHWC  [B, H, W, C]
 *   tap source ("taps")  16-bit hi and lo planes, NHWC with padding already applied:
 *                        [B * planes, Hp, Wp, Cp], Cp % 64 == 0; value = hi + lo (3-term split operand)
 *   packed weight        16-bit hi and lo, K-major [Cout_pad, num_taps * Cp]
 *   instance statistics  partial  [B * H*W/32, C, 2] = (sum, centred M2) per 32-pixel run
 *                        reduced  [B, C, 2]          = (mean, 1/sqrt(var + eps))
 */
#ifndef TSNET_B200_H_
#define TSNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSNET_ABI_VERSION 3

/* 16-bit operand format of the tensor-core path */
#define TSNET_FMT_FP16 0
#define TSNET_FMT_BF16 1

/* tap-source builder modes (tsnet_build_taps) */
#define TSNET_TAPS_SAME 0     /* no padding: 1x1 convs, channel-concat targets                       */
#define TSNET_TAPS_REFLECT1 1 /* nn.ReflectionPad2d(1) for 3x3 stride-1 convs                       */
#define TSNET_TAPS_S2ZERO 2   /* zero pad 1 + parity split into 4 planes for 3x3 stride-2 convs      */
#define TSNET_TAPS_UP2REFLECT1 3 /* nn.Upsample(x2, bilinear, align_corners=False) + ReflectionPad2d(1) */
#define TSNET_TAPS_WINO 4     /* ReflectionPad2d(1) + Winograd F(2x2,3x3) input transform B^T d B: 16 planes of
                                 H/2 x W/2 tiles, [B, 16, H/2, W/2, Cp] (operand of tsnet_wino_gemm_fwd)       */

#define TSNET_MAX_TAPS 49

int tsnet_abi_version(void);
const char* tsnet_last_error(void);
/* 1 if the current device is compute capability 10.x, else 0 */
int tsnet_device_ok(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py: gpu_launches) */
long long tsnet_launch_count(void);

/* ---- weights --------------------------------------------------------------------------------
 * nn.Conv2d weight [Cout, Cin, KH, KW] fp32 (state_dict layout) -> packed K-major hi/lo.
 * fold_kw = 0: taps = KH*KW (row-major r,s), K per tap = Cp >= Cin, channel c at column tap*Cp + c.
 * fold_kw = 1: taps = KH, K per tap = Cp >= KW*Cin, column tap*Cp + s*Cin + c  (7x7 stems, see
 *              tsnet_stem_taps).  Rows >= Cout and unused columns are zero.
 * fold_kw = F > 1: as fold_kw = 1 with the channels of every horizontal tap padded to F >= Cin: column
 *              tap*Cp + s*F + c (tsnet_stem_conv_fwd: F = 8, one 16-byte chunk per tap).
 * scale: weights are multiplied by `scale` (power of two; FP16 mode range management) before the split. */
int tsnet_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int KH, int KW, int fold_kw, int Cp,
                           int Cout_pad, float scale, int fmt, uint16_t* w_hi, uint16_t* w_lo, void* stream);

/* ---- implicit-GEMM convolution on tcgen05 ---------------------------------------------------
 * Replaces nn.Conv2d (+ the preceding pad / upsample, which tsnet_build_taps folded into the tap
 * source): model/TSNet.py:27,42 (ResnetBlock 3x3), :70 (stride-2), :66 (7x7 stem, kw-folded),
 * :139 (map_conv 1x1), :147 (decoder up-convs), :193 (FuseNet 1x1).
 * y_raw [B,H,W,Cout] fp32 = conv + bias (pre-norm); stats_partial as above (may be NULL). */
typedef struct {
  int B, H, W;      /* output geometry */
  int Cout;         /* real output channels (multiple of 32) */
  int Cout_pad;     /* rows of the packed weight, multiple of block_n */
  int Cp;           /* K per tap = channels per pixel of the tap source (multiple of 64) */
  int Hp, Wp, planes; /* tap source geometry */
  int num_taps;
  int8_t tap_dy[TSNET_MAX_TAPS], tap_dx[TSNET_MAX_TAPS], tap_plane[TSNET_MAX_TAPS];
  int block_n;      /* 64, 128 or 256 */
  int split;        /* 1: hi*hi + hi*lo + lo*hi (fp32-faithful); 0: hi*hi only (fast, non-parity) */
  int fmt;          /* TSNET_FMT_* */
  float out_scale;  /* accumulators are multiplied by this before bias (undo operand pre-scaling) */
  const float* addend; /* optional fp32 [addend_rows, Cout] (device): y[m] += addend[m % addend_rows] before the
                          statistics.  Lets a conv over cat[a_i, t] be computed as conv(a_i) + conv(t) with the
                          source-independent half evaluated once per frame (FuseNet, model/TSNet.py:196-198) */
  int addend_rows;
  /* ---- fused InstanceNorm epilogue (optional; requires H*W == 1024, i.e. 8 tiles of 128 pixels per image) ----
   * fuse_in = 1: the 8 CTAs of a thread-block cluster compute the 8 pixel tiles of one image for one channel slab,
   * exchange the per-tile statistics over distributed shared memory and apply, in registers,
   *   v = (conv + bias [+ addend] - mean) * rstd ; if fuse_relu: v = max(v, 0) ; if fuse_residual: v += residual
   * then write fuse_act_out (fp32, optional) and the NEXT layer's tap source (hi / lo, mode SAME or REFLECT1,
   * optional) directly: y_raw / stats_partial are not written, and the separate tsnet_instnorm_reduce +
   * tsnet_build_taps launches of that layer disappear (model/TSNet.py:27-48: conv -> IN -> ReLU / + x -> pad). */
  int fuse_in;
  int fuse_relu;
  int fuse_mode;               /* TSNET_TAPS_SAME or TSNET_TAPS_REFLECT1 */
  const float* fuse_residual;  /* fp32 [B, H, W, Cout] or NULL */
  float* fuse_act_out;         /* fp32 [B, H, W, fuse_act_C_total] or NULL */
  int fuse_act_C_total, fuse_act_c_off;
  uint16_t* fuse_taps_hi;      /* [B, Hd, Wd, fuse_taps_Cp] or NULL */
  uint16_t* fuse_taps_lo;
  int fuse_taps_Cp, fuse_taps_c_off;
  float fuse_act_scale;        /* power-of-two scale of the written operands */
  float fuse_eps;              /* 1e-5 */
  int flags;                   /* TSNET_CONV_* launch-plan switches (0 = default plan); results never depend on them */
} tsnet_conv_desc;

/* launch-plan switches (tests compare the alternative plans bit for bit) */
#define TSNET_CONV_NO_VR 1         /* kw-folded stems: plain implicit-GEMM kernel instead of the vertical-reuse one  */
#define TSNET_CONV_NO_TAIL_SPLIT 2 /* no second launch with block_n / 2 for the last partial wave                    */
#define TSNET_CONV_ONE_CTA 4       /* block_n = 256: 1-CTA kernel instead of the default cta_group::2 pair kernel     */
#define TSNET_CONV_SMALL_FIRST 8   /* tsnet_wino_gemm_fwd only (changes rounding, not the function): per K block issue the
                                      hi*lo and lo*hi MMAs before the hi*hi ones (truncation-error experiment)          */

int tsnet_conv_gemm_fwd(const tsnet_conv_desc* d, const uint16_t* taps_hi, const uint16_t* taps_lo,
                        const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, float* y_raw,
                        float* stats_partial, void* stream);

/* ---- Winograd F(2x2, 3x3) path of the 32 x 32 ResnetBlock convolutions --------------------------------------------
 * model/TSNet.py:10-49 (ReflectionPad2d(1) + Conv2d(dim, dim, 3) of ResnetBlock: img_enc x 18, FuseNet x 2,
 * decoder x 2 n_blocks).  Same function as tsnet_conv_gemm_fwd on a REFLECT1 tap source, with 2.25 x fewer MACs:
 *   U = G g G^T   tsnet_wino_weight_transform (fp64 arithmetic -> fp32 [16, Cout, Cin]) then
 *                 tsnet_pack_conv_weight(U viewed as a 1x1 weight [16*Cout, Cin, 1, 1]) -> hi/lo [16*Cout, Cp]
 *   V = B^T d B   tsnet_build_taps(mode = TSNET_TAPS_WINO) on the producer's raw output (InstanceNorm / ReLU /
 *                 residual / reflect pad applied in the same pass) -> hi/lo [B, 16, H/2, W/2, C]
 *   M[p] = V[p] . U[p]^T   tsnet_wino_gemm_fwd: the 16 plane GEMMs as one batched tcgen05 launch (cta_group::2 pairs,
 *                 chunked accumulation with promotion to fp32 registers as in tsnet_conv_gemm_fwd)
 *                 -> fp32 [16, B * H/2 * W/2, Cout]
 *   y = A^T M A + bias (+ addend)   tsnet_wino_output -> y_raw [B, H, W, Cout] + stats_partial, exactly the outputs of
 *                 tsnet_conv_gemm_fwd, so everything downstream is unchanged. */
typedef struct {
  int B, TH, TW;    /* samples, tile rows / columns per sample (H/2, W/2); TH*TW % 128 == 0 */
  int C;            /* input channels = K of every plane GEMM (multiple of 64) */
  int Cout;         /* output channels (multiple of 256) */
  int split, fmt;
  float out_scale;  /* 1 / (weight scale * activation scale) */
  int chunk_kb;     /* 0 = default (2 K blocks per TMEM chunk in split mode) */
  int flags;        /* TSNET_CONV_ONE_CTA, TSNET_CONV_SMALL_FIRST */
} tsnet_wino_gemm_desc;

int tsnet_wino_weight_transform(const float* w_oihw, int Cout, int Cin, float* u_out, void* stream);
int tsnet_wino_gemm_fwd(const tsnet_wino_gemm_desc* d, const uint16_t* v_hi, const uint16_t* v_lo,
                        const uint16_t* u_hi, const uint16_t* u_lo, float* m_out, void* stream);
/* addend (optional): fp32 [addend_rows, C] added per output pixel (row index modulo) before the statistics -- the
 * source-independent half of FuseNet's first convolution (see tsnet_conv_desc.addend). W/2 must be a multiple of 8. */
int tsnet_wino_output(const float* m, int B, int H, int W, int C, const float* bias, const float* addend,
                      long long addend_rows, float* y_raw, float* stats_partial, void* stream);

/* Bridge between two consecutive Winograd layers (conv -> InstanceNorm -> [ReLU | + x] -> ReflectionPad2d -> conv,
 * model/TSNet.py:25-48): output transform of layer k (+ bias, + addend), InstanceNorm statistics (fp64, fixed order,
 * CTA-local: one CTA = one image x 32 channels held in shared memory), normalisation, ReLU / residual, optional fp32
 * act_out, and the input transform of layer k+1 in ONE pass: M is read once, the next V is written once; y_raw,
 * stats_partial, tsnet_instnorm_reduce and the separate tsnet_build_taps pass of that layer boundary disappear.
 * mean_rstd_out (optional, [B, C, 2]) receives the statistics. */
typedef struct {
  int B, H, W, C;
  int relu;
  int Cp_total, c_off;          /* destination operand planes [B, 16, H/2, W/2, Cp_total], channel window [c_off, +C) */
  int fmt;
  float scale, eps;             /* operand pre-scale (power of two); InstanceNorm eps (0 = 1e-5) */
  int act_C_total, act_c_off;   /* act_out channel window (0 = C) */
  long long addend_rows;
  /* optional (all four or none): the activations also leave as the un-normalised operand rows of the correlation --
   * corr_hi / corr_lo [B * H*W, C] at the row's sorted rank (corr_rank [B, H*W], from tsnet_corr_rank_table), values
   * scaled by corr_scale, and corr_ssq [B][C / 32][H*W] = per-slab partial sums of squares (-> tsnet_corr_norms).
   * Used by the last ResnetBlock of img_enc: the source features never make a separate trip through an operand pass. */
  uint16_t* corr_hi;
  uint16_t* corr_lo;
  const uint16_t* corr_rank;
  float* corr_ssq;
  float corr_scale;
  int variant;                  /* 0: 32-channel slabs, 1 CTA per SM (128 KB image buffer); 1: 16-channel slabs, 2 CTAs per SM
                                   (corr_ssq then has C / 16 slabs).  Same function, different launch geometry. */
} tsnet_wino_bridge_desc;
int tsnet_wino_bridge(const tsnet_wino_bridge_desc* d, const float* m, const float* bias, const float* addend,
                      const float* residual, float* act_out, float* mean_rstd_out, uint16_t* v_hi, uint16_t* v_lo,
                      void* stream);

/* ---- InstanceNorm statistics -----------------------------------------------------------------
 * nn.InstanceNorm2d(affine=False, eps=1e-5, biased variance): model/TSNet.py:28,43,66,71,149.
 * partial [B*HW/32, C, 2] -> mean_rstd [B, C, 2]; fixed-order (deterministic) Chan merge in fp64. */
int tsnet_instnorm_reduce(const float* stats_partial, int B, int HW, int C, float eps, float* mean_rstd,
                          void* stream);

/* ---- normalise / activate / residual / pad / split --------------------------------------------
 * One pass over a raw conv output that produces what the NEXT layer needs:
 *   v = raw; if mean_rstd: v = (v - mean) * rstd; if relu: v = max(v, 0); if residual: v += residual
 *   act_out (optional) <- v (fp32 NHWC)      taps (optional) <- pad/upsample(mode)(v) split into hi/lo
 * Replaces InstanceNorm + ReLU + residual add + ReflectionPad2d / Upsample / torch.cat of
 * model/TSNet.py:19-48, 145-150, 163, 196 (cat = two producers writing channel ranges c_off of a
 * wider tap source of Cp_total channels). */
typedef struct {
  int B, H, W, C;   /* source geometry */
  int mode;         /* TSNET_TAPS_* */
  int relu;
  int Cp_total;     /* channels per pixel of the destination tap source */
  int c_off;        /* first destination channel */
  int fmt;
  float scale;      /* activations are multiplied by `scale` before the split (power of two) */
  int act_C_total;  /* channels per pixel of act_out (0 = C) */
  int act_c_off;    /* first act_out channel */
  int avg_n;        /* > 1 (mode SAME only): v = mean over i < avg_n of f(raw[i*B + b]) -- torch.stack().mean(1)
                       of model/TSNet.py:400; raw / mean_rstd / residual then hold avg_n*B samples */
  int flags;        /* TSNET_TAPS_GENERIC_UP2: one-destination-pixel-per-thread kernel for mode UP2REFLECT1 (tests) */
  /* two-part residual (modes SAME / REFLECT1 / S2ZERO): the residual is torch.cat([residual, residual2], 1) without the
   * concatenation ever being written -- channels [0, res_split) from `residual` [B, H, W, res_split], the others from
   * residual2 [res2_batch, H, W, C - res_split], sample index modulo res2_batch (FuseNet: x = cat[src_fea_i, tar_fea],
   * model/TSNet.py:196, with ONE tar_fea shared by the n sources) */
  const float* residual2;
  int res_split, res2_batch;
} tsnet_taps_desc;
#define TSNET_TAPS_GENERIC_UP2 1

int tsnet_build_taps(const tsnet_taps_desc* d, const float* raw, const float* mean_rstd, const float* residual,
                     float* act_out, uint16_t* taps_hi, uint16_t* taps_lo, void* stream);

/* ---- encoder stem input ------------------------------------------------------------------------
 * Builds the kw-folded tap source of the 7x7 stem conv directly from the NCHW network inputs:
 * torch.cat([img, lbl]) (model/TSNet.py:312), Encoder.coord_conv (:107-125, channels x, y, r generated
 * analytically) and ReflectionPad2d(3) (:66).  Destination [B, H+6, W, Cp]: pixel (yp, x) holds, for
 * s = 0..6, the Cin = Cimg + Clbl + 3 channels of source pixel (reflect(yp-3), reflect(x+s-3)).
 * img may be NULL (label encoder).  The image is DIVIDED by img_div (the /255.0 of set_*_input, :268-286;
 * a true division so the rounding matches the reference).
 * img_kind 0: img = fp32 planes [B, Cimg, H, W] (mean-subtracted BGR in 0..255 units, what the reference's datasets
 *             emit); img_kind 1: img = uint8 BGR planes [B, 3, H, W] and img_mean3_host = the dataset mean (3 host
 *             floats): the dataset's `image -= mean` (dataset/dataset_video_face.py:329, :401) moves into the loader,
 *             (float(u8) - mean_c) / img_div -- 4 x fewer image bytes over PCIe and HBM, bit-identical values.
 * lbl_kind 0: lbl = fp32 one-hot planes [B, Clbl, H, W] (what the reference's callers pass);
 * lbl_kind 1: lbl = uint8 class-index map [B, H, W]; channel c is (lbl == c), i.e. utils/misc.py:50-67 `vl2ch`
 *             evaluated inside the loader -- 4 x Clbl fewer bytes over PCIe and HBM. */
int tsnet_stem_taps(const void* img_nchw, int Cimg, int img_kind, const float* img_mean3_host, float img_div,
                    const void* lbl, int Clbl, int lbl_kind, int B, int H, int W, int Cp, int fmt, float scale,
                    uint16_t* taps_hi, uint16_t* taps_lo, void* stream);

/* ---- encoder stem without a materialised operand -----------------------------------------------------------------
 * ReflectionPad2d(3) + Conv2d(Cin, 64, 7) (model/TSNet.py:66) of both encoders straight from the raw NCHW network inputs:
 * the kw-folded operand tile of every 8 x 16 output tile is generated in shared memory by producer warps of the
 * tcgen05 kernel -- torch.cat([img / 255, lbl]) (:312), the datasets' uint8 -> float / mean-subtract / one-hot staging
 * and Encoder.coord_conv (:107-125) never exist in HBM (tsnet_stem_taps materialises a 1.65 GB operand at bs = 32,
 * n_source = 3).  Needs Cimg + Clbl + 3 <= 8 (the face configuration: 3 + 2 + 3 and 0 + 2 + 3); the weight is packed
 * with tsnet_pack_conv_weight(fold_kw = 8).  Larger label sets (pose: 25 classes) use tsnet_stem_taps +
 * tsnet_conv_gemm_fwd.  img_kind / lbl_kind / img_mean / img_div as in tsnet_stem_taps.  Outputs as tsnet_conv_gemm_fwd. */
typedef struct {
  int B, H, W;
  int Cimg, Clbl, img_kind, lbl_kind;
  float img_mean[3], img_div;
  int Cout;          /* 64 */
  int split, fmt;
  float act_scale;   /* power-of-two pre-scale of the generated operands */
  float out_scale;   /* 1 / (weight scale * act_scale) */
} tsnet_stem_conv_desc;
int tsnet_stem_conv_fwd(const tsnet_stem_conv_desc* d, const void* img_nchw, const void* lbl, const uint16_t* w_hi,
                        const uint16_t* w_lo, const float* bias, float* y_raw, float* stats_partial, void* stream);

/* ---- correlation: masks -> class-sorted order, operands, tensor-core tiles, warp + mean -----------------------
 * model/TSNet.py:319-323 (normalise, target mask), :339-366 (per source: normalise, mask, two masked bmm,
 * softmax(100 x), expected coordinate, grid_sample), :392 (mean over sources) and the first half of
 * torch.cat([pg, sg]) feeding Decoder.map_conv (:163).  The HW x HW matrix is never materialised.
 *
 * Call order on one stream (SURVEY section 8b names this group `tsnet_corr_warp_fwd`):
 *   1. tsnet_corr_prepare     masks -> per-map class-sorted order, tile classes, work list, closed forms
 *   2. operands, rows written in the sorted order: tsnet_corr_operands (un-normalised rows + reciprocal norms; the
 *      default) for the target and -- unless the producing tsnet_wino_bridge already wrote them (corr_* fields +
 *      tsnet_corr_norms) -- the sources; or tsnet_l2norm_split x2 (rows normalised up front, rnorm_* = NULL)
 *   3. tsnet_corr_warp_fwd    = tsnet_corr_tiles (tcgen05 similarity tiles -> partial softmax states)
 *                             + tsnet_corr_finish (merge -> warp grid -> grid_sample -> source mean)
 *
 * Why sorted: the reference's similarity is (T.S) * (mt*ms + (1-mt)*(1-ms)) -- for {0,1} bbox masks every pair of
 * positions whose classes differ has logit exactly 0.  With the positions of each map stably sorted by class, a
 * 128-row target tile and a 256-column source chunk of different pure classes need no tensor work: their softmax
 * contribution (max 0, weight 256, coordinate sums) is written by tsnet_corr_prepare.  Soft (non-binary) masks are
 * handled exactly (they sort between the classes and are never skipped).
 *
 * bbox pointers are the FULL-RESOLUTION masks [B, bbox_h, bbox_w] (uint8 or fp32); nearest down-sampling to (h, w)
 * is integer indexing inside tsnet_corr_prepare.  coord_table = h + w floats: torch.linspace(-1,1,h) then
 * torch.linspace(-1,1,w) (:301-302).  The workspace (tsnet_corr_workspace_bytes, 256 B aligned, caller-owned) carries
 * everything between the calls; it may be reused by the next forward on the same stream. */
/* Limits of one call (checked; the message is in tsnet_last_error()): n_src <= 12 (the reference's callers use 1..8;
 * split a larger source set over several calls and average the means), B <= 1024, h*w a multiple of 256 and <= 1024,
 * C a multiple of 128 and <= 1024.  The reference geometry is h = w = 32, C = 512. */
typedef struct {
  int B, n_src, C, h, w;
  int bbox_h, bbox_w, bbox_dtype; /* 0 = uint8, 1 = fp32 */
  float temperature;              /* 100 (model/TSNet.py:359) */
  int split, fmt;
  float operand_scale;            /* product of the two operand pre-scales (accumulators are divided by it) */
  int sort;                       /* 1: class-sorted order + tile skipping (default); 0: raster order */
  int one_cta;                    /* 0 (default): 2-CTA tile kernel (cta_group::2, work items of 256 target rows);
                                     1: 1-CTA kernel (128-row items).  Read by prepare AND tiles from this descriptor,
                                     so the work-list granularity always matches the kernel that consumes it. */
  int chunk_kb;                   /* K blocks accumulated in TMEM before promotion to registers; 0 = whole K */
} tsnet_corr_desc;

size_t tsnet_corr_workspace_bytes(const tsnet_corr_desc* d);
int tsnet_corr_prepare(const tsnet_corr_desc* d, const void* tar_bbox, const void* const* src_bbox,
                       const float* coord_table, void* workspace, size_t workspace_bytes, void* stream);
/* position -> sorted-rank tables inside a prepared workspace, for tsnet_l2norm_split:
 * which = 0: target [B, hw]; which = 1: sources [n_src * B, hw] (source-major, like the source features) */
const uint16_t* tsnet_corr_rank_table(const tsnet_corr_desc* d, const void* workspace, int which);

/* F.normalize(fea, p=2, dim=1) (model/TSNet.py:319, :339; eps = 1e-12) of an fp32 NHWC feature map [B, HW, C],
 * written as 16-bit hi/lo K-major operands [B*HW, C] (scaled by `scale`).  rank (optional, [B, HW] uint16):
 * row p of sample b goes to row b*HW + rank[b*HW + p]. */
int tsnet_l2norm_split(const float* fea, int B, int HW, int C, int fmt, float scale, const uint16_t* rank,
                       uint16_t* out_hi, uint16_t* out_lo, void* stream);

/* Operands WITHOUT the normalisation (the default path since round 2): x * scale split into hi / lo at the row's
 * sorted rank plus rnorm = 1 / max(||x||_2, 1e-12) at the same rank; tsnet_corr_tiles applies rnorm_t[row] * rnorm_s[col]
 * to the similarity inside the softmax FMA (F.normalize of model/TSNet.py:319, :339 as a scale of the dot product: no
 * per-element division, and the operands of the SOURCES can be written by the pass that produces the features:
 * tsnet_wino_bridge's corr_* outputs + tsnet_corr_norms).  tsnet_corr_desc.operand_scale = scale * scale. */
int tsnet_corr_operands(const float* fea, int B, int HW, int C, int fmt, float scale, const uint16_t* rank,
                        uint16_t* out_hi, uint16_t* out_lo, float* rnorm, void* stream);
/* ssq_part [B][slabs][HW] (per-slab partial sums of squares from tsnet_wino_bridge) -> rnorm [B * HW] at the sorted rank */
int tsnet_corr_norms(const float* ssq_part, int B, int HW, int slabs, const uint16_t* rank, float* rnorm, void* stream);

/*   tar_hi/lo   [B*hw, C]          target operands, sorted rows (normalised if rnorm_* are NULL, else un-normalised)
 *   rnorm_t/s   NULL, or [B*hw] / [n_src*B*hw] reciprocal norms at the sorted ranks (see tsnet_corr_operands)
 *   src_hi/lo   [n_src, B*hw, C]   operands of all sources (ONE buffer, source-major), sorted rows
 *   src_fea     n_src pointers (host array) to fp32 NHWC un-normalised source features (sampled by grid_sample)
 *   out_mean    NULL or fp32 NHWC [B, hw, C] = mean_i grid_sample(src_fea_i, G_i)
 *   out_grids   NULL or [n_src, B, h, w, 2] (x, y) -- the reference's warp_grid2d_list (:369-370)
 *   taps_hi/lo  NULL or the hi/lo tap source [B, h, w, Cp_total] of the 1x1 conv that consumes the mean: channel
 *               window [c_off, c_off + C), values scaled by taps_scale ("grid_sample fused with the next conv's load") */
int tsnet_corr_warp_fwd(const tsnet_corr_desc* d, const uint16_t* tar_hi, const uint16_t* tar_lo,
                        const uint16_t* src_hi, const uint16_t* src_lo, const float* rnorm_t, const float* rnorm_s,
                        const float* const* src_fea, float* out_mean, float* out_grids, uint16_t* taps_hi,
                        uint16_t* taps_lo, int Cp_total, int c_off, float taps_scale, void* workspace,
                        size_t workspace_bytes, void* stream);
/* the two halves of tsnet_corr_warp_fwd, separately launchable (profiling, grids-only use).  tsnet_corr_tiles runs
 * the 2-CTA tile kernel (tcgen05.mma.cta_group::2, work items of 256 target rows) unless tsnet_corr_desc.one_cta
 * selects the 1-CTA kernel. */
int tsnet_corr_tiles(const tsnet_corr_desc* d, const uint16_t* tar_hi, const uint16_t* tar_lo, const uint16_t* src_hi,
                     const uint16_t* src_lo, const float* rnorm_t, const float* rnorm_s, void* workspace,
                     size_t workspace_bytes, void* stream);
int tsnet_corr_finish(const tsnet_corr_desc* d, const float* const* src_fea, float* out_mean, float* out_grids,
                      uint16_t* taps_hi, uint16_t* taps_lo, int Cp_total, int c_off, float taps_scale, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---- warp + source mean from given grids ---------------------------------------------------------------------
 * F.grid_sample(src_fea_i, G_i, bilinear, zeros, align_corners=False) (model/TSNet.py:366) + mean over sources
 * (:392) for caller-supplied grids [n_src, B, h, w, 2]; outputs as in tsnet_corr_warp_fwd. */
int tsnet_warp_mean_taps(const float* const* src_fea, int n_src, const float* grids, int B, int h, int w, int C,
                         float* out_mean, uint16_t* taps_hi, uint16_t* taps_lo, int Cp_total, int c_off, int fmt,
                         float scale, void* stream);

/* ---- output head ---------------------------------------------------------------------------------
 * ReflectionPad2d(3) + Conv2d(64 -> 3, 7x7) + Tanh (model/TSNet.py:151-152) on the fp32 NHWC
 * activation, written as NCHW [B, 3, H, W]; optional pose compositing
 * out = out * fore + fill_c * (1 - fore), fore = columns [fore_x0, fore_x1) (model/TSNet_pose.py:276-280,
 * :416-417; fill = -mean/255).  Pass fore_x1 <= fore_x0 to disable.  SIMT fp32 (Cout = 3 is not a
 * tensor-core shape). */
int tsnet_head_conv_tanh(const float* act_nhwc, const float* mean_rstd, int relu, int B, int H, int W, int Cin,
                         const float* w_oihw, const float* bias, int fore_x0, int fore_x1, const float* fill3,
                         float* out_nchw, void* stream);
/* mean_rstd (optional, [B, Cin, 2] from tsnet_instnorm_reduce) and relu: act_nhwc is then the RAW output of the last
 * up-conv and InstanceNorm + ReLU (model/TSNet.py:149-150) are applied while the tile is staged -- no separate
 * normalisation pass over the 256 x 256 x 64 activation. */

/* ---- demo post-processing -------------------------------------------------------
 * demo/demo_face.py:194-199 + sample_img :96-105 (demo_pose.py analogous): per image and channel
 *   y = (x - mean) / std * ref_std + ref_mean      (mean / unbiased std of the generated frame, per channel)
 *   y = clamp(y + img_mean, 0, 1) * 255 ; BGR -> RGB ; uint8 (truncation)
 * rec_nchw [B,3,H,W] fp32 (device); gen_mean_std [B*3, 2] = (mean, unbiased std) of every generated plane from
 * tsnet_plane_stats(rec_nchw, B*3, H*W, 1, ...) (device); ref_mean3 / ref_std3 device pointers to 3 floats (statistics
 * of the source frames, computed by the caller as in demo_face.py:180-182); img_mean3_host = IMG_MEAN/255 (host
 * pointer); out [B,H,W,3] uint8 RGB (device).  Given the same statistics the bytes equal the reference's fp32
 * arithmetic exactly (tests compare with torch.equal).  Removes the fp32 D2H + numpy passes per frame. */
int tsnet_postprocess_u8(const float* rec_nchw, int B, int H, int W, const float* gen_mean_std, const float* ref_mean3,
                         const float* ref_std3, const float* img_mean3_host, uint8_t* out_hwc_rgb, void* stream);

/* ---- train-mode branches inside forward() (forward only) ---------------------------------
 * model/TSNet.py:327-331 (target image statistics), :372-390 (image-space warp: F.unfold(src_img, 8, 8) ->
 * grid_sample with the warp grid -> F.fold; per-image re-normalisation to the target's mean / unbiased std; warp loss
 * 10 * L1), :402-405 (alignment loss 1 - mean cosine similarity of the two branch means; pass NULL means for the pose
 * variant, which has no such loss) and model/TSNet_pose.py:395-396 (foreground compositing, fore = columns
 * [fore_x0, fore_x1), fill = -mean/255; pass fore_x1 <= fore_x0 to disable).
 *   src_img    n_src host-array pointers to RAW NCHW images [B,3,H,W]; src_div[i] = 255 or 1 (use_prev), host array
 *   tar_img    raw NCHW target image, divided by tar_div
 *   grids      [n_src, B, h, w, 2] from tsnet_corr_warp_fwd
 *   pg_mean / sg_mean  fp32 NHWC [B, h*w, C] branch means (or both NULL)
 *   warp_out   [n_src, B, 3, H, W]  = the reference's warp_src_img_list
 *   losses2    device float[2] = (loss_warp, loss_align)
 * All reductions are fp64 with a fixed order (bit-reproducible). */
size_t tsnet_train_extras_workspace_bytes(int B, int n_src, int h, int w);
int tsnet_train_extras_fwd(const float* const* src_img, const float* src_div, int n_src, const float* tar_img,
                           float tar_div, const float* grids, int B, int H, int W, int h, int w, const float* pg_mean,
                           const float* sg_mean, int C, int fore_x0, int fore_x1, const float* fill3, float* warp_out,
                           float* losses2, void* workspace, size_t workspace_bytes, void* stream);
/* (mean, unbiased std) of x / div over each of `planes` contiguous runs of n floats -> mean_std [planes, 2] */
int tsnet_plane_stats(const float* x, int planes, int n, float div, float* mean_std, void* stream);

/* ---- reference-style direct convolution (validation kernel, fp32 SIMT) ----------------------------
 * Plain NHWC direct convolution with zero or reflect padding; used by the tests to cross-check the
 * tensor-core path on device and by nothing on the hot path. */
int tsnet_direct_conv_fp32(const float* x_nhwc, int B, int H, int W, int Cin, const float* w_oihw,
                           const float* bias, int Cout, int K, int stride, int pad, int reflect, float* y_nhwc,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TSNET_B200_H_ */
