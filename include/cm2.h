/*
 * cm2.h -- C ABI of libcm2.so: hand-written sm_100a kernels for the CenterMask2 inference path.
 *
 * The reference (Zeng-Yan/centermask2) is pure Python behind detectron2's registry; it has no FFI
 * of its own.  The drop-in boundary is therefore the registry surface; this
 * header is the layer *below* it: one entry point per kernel family, each replacing the torch /
 * torchvision / detectron2 call the reference makes at the cited place.  Paths are relative to
 * /root/reference/centermask2/centermask/ ; "[d2]" marks un-vendored detectron2 v0.5 code whose
 * semantics are described in SURVEY.md Appendix A.
 *
 * Conventions
 *   - Every function returns 0 (CM2_OK) or a negative CM2_ERR_* code; cm2_last_error() gives a
 *     thread-local message.  Nothing throws or aborts.
 *   - All pointers documented as device pointers are caller-allocated device memory; the library
 *     never allocates, frees or retains device memory.
 *   - `stream` is a cudaStream_t passed as void*.  All work is enqueued asynchronously on it; there
 *     are no hidden synchronisations, so every call is CUDA-graph capturable.
 *   - Activations are *pitched NHWC views* (cm2_act): channels are innermost and dense, the pixel,
 *     row and image strides are explicit (in elements).  This lets the caller keep every feature
 *     map inside a buffer with a one-pixel zero halo ([n, h+2, w+2, c], view = interior), which
 *     is what the tensor-core convolution engine requires.
 *   - There is no CPU fallback: on a device that is not sm_100 the tensor-core entry points return
 *     CM2_ERR_UNSUPPORTED.
 */
#ifndef CM2_H_
#define CM2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM2_VERSION 200

#define CM2_OK 0
#define CM2_ERR_BAD_SHAPE (-1)
#define CM2_ERR_UNSUPPORTED (-2)
#define CM2_ERR_CUDA (-3)

#define CM2_F32 0
#define CM2_BF16 1
#define CM2_U8 2
#define CM2_F16 3 /* IEEE half: operand type of the split-precision tensor-core convolution (see cm2_split_f16x2) */

#define CM2_ENGINE_SIMT 0 /* fp32-accumulate CUDA-core implicit GEMM; any shape; f32 or bf16 I/O   */
#define CM2_ENGINE_TC 1   /* tcgen05/TMEM implicit GEMM with TMA-staged tiles; bf16 or f16 in      */

#define CM2_MAX_SRC 8
#define CM2_MAX_SEG 8

/* Pitched NHWC view.  `data` addresses element (n=0, y=0, x=0, c=0); channel stride is 1. */
typedef struct cm2_act {
  void* data;
  int32_t n, h, w, c;
  int64_t sn, sh, sw; /* element strides of image, row, pixel */
} cm2_act;

/* One segment of a *segmented* halo tensor: several feature maps of different extent (the FPN levels the
 * shared-weight FCOS towers run on, fcos.py:227-238) stored back to back in one flat [rows, c] buffer.
 * Segment s occupies rows [row0, row0 + n*(h+2)*(w+2)) as a halo-1 block [n, h+2, w+2, c]; row0 must be a
 * multiple of 256; rows between segments are padding (ignored on input, untouched on output).
 * halo == 1 (tensor-core convolutions, cm2_groupnorm_apply_seg on bf16, cm2_groupnorm_apply_seg_split): the images of the segment SHARE their zero
 * frame -- line pitch w + 1, image pitch (h + 1)(w + 1) rows, n*(h+1)*(w+1) rows plus one trailing zero line + pixel
 * (w + 2 rows) that the caller keeps zero: 2.4 % instead of 9.7 % frame rows on the five FCOS levels at 800x1344. */
typedef struct cm2_seg {
  int64_t row0;
  int32_t n, h, w;
  int32_t halo; /* 0: own frame per image; 1: shared frame */
} cm2_seg;

int cm2_version(void);
const char* cm2_last_error(void);
/* Number of SMs / compute capability of the current device (0 if no device). */
int cm2_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Convolution (the dense contractions).  Replaces every F.conv2d / conv_transpose2d / F.linear on
 * the path: modeling/backbone/vovnet.py:205-236 (conv-FrozenBN-ReLU), :324-325 (torch.cat + 1x1,
 * here a *virtual* concat over `src[]`), [d2] FPN lateral/output convs incl. the top-down
 * nearest-2x upsample-add (res_mode 2), modeling/backbone/fpn.py:32-35, modeling/fcos/fcos.py:
 * 169-200, modeling/centermask/sam.py:58-83 (deconv: out_mode 1), maskiou_head.py:78-94.
 *
 *   out[p, co] = act( scale[co] * sum_{tap, s, c} src_s[p*stride + tap - pad][c] * W[co, tap, s, c]
 *                     + shift[co] + residual[p, co] )
 *
 * Weight layouts (built once by the caller; see centermask2_b200/packing.py):
 *   SIMT : [kh*kw*cin_total][cout]          row k = (ky*kw + kx)*cin_total + c,  dtype = `dtype`
 *   TC   : [cout_pad][k_tc] bf16 (f16 [2*cout_pad][k_tc] for split precision, see below), K-major;
 *          k = (tap, source, channel padded to 64 per source);
 *          k_tc = kh*kw * sum_s roundup(src_c[s], 64)  (cm2_conv_tc_klen); cout_pad =
 *          roundup(cout, 16); padding entries are zero.
 *
 * TC engine constraints (cm2_conv_tc_supported): bf16 or f16 sources, stride 1, kernel 1x1 (pad 0) or
 * 3x3 (pad 1); all sources and the output are interior views of one-pixel-halo buffers of identical
 * geometry (sh == (w+2)*sw, sn == (h+2)*sh; a source may be a channel slice, sw >= c), or, for 1x1
 * only, fully dense views (sh == w*sw, sn == h*sh).  With halo buffers the engine also (re)writes the
 * halo of the output with zeros, so a chain of convolutions keeps the invariant "halo == 0".
 *
 * Split precision (dtype CM2_F16; the fp32-accuracy variant of the tensor-core path): the caller keeps
 * activations in fp32 and converts a convolution's inputs with cm2_split_f16x2 into [.., 2c] half tensors
 * (channels [0, c) = hi = half(x), [c, 2c) = lo = half(x - hi)); src[i] is that split tensor (src[i].c = 2c,
 * 2c % 32 == 0).  The weights are f16 [2*cout_pad][k_tc] in the TC layout over the c LOGICAL channels of every
 * source: rows [0, cout_pad) = W_hi = half(s*W), rows [cout_pad, 2*cout_pad) = W_lo = half(s*W - W_hi), with a
 * per-output-channel power-of-two pre-scale s (undone through `scale`) that keeps W_lo out of the half
 * subnormals.  The kernel accumulates x_hi*W_hi + x_lo*W_hi + x_hi*W_lo -- 22 significant bits per operand, the
 * dropped x_lo*W_lo term is 2^-22 relative -- and, because tcgen05.mma adds into its accumulator with
 * truncation, sums the dominant x_hi*W_hi term in short chunks that are added in fp32 registers with
 * round-to-nearest (csrc/conv_tc.cu, conv_tc3_kernel).  Output and residual are CM2_F32; out_mode 3 is not
 * available.  out_dtype CM2_F16 (out_mode 0 / 2, no residual, no statistics): the result is stored directly as the
 * [hi | lo] pair of the NEXT convolution -- out.c = 2*cout, hi at channel co, lo at cout + co -- so a chain of
 * convolutions needs no fp32 copy and no cm2_split_f16x2 pass in between.
 * ------------------------------------------------------------------------------------------- */
typedef struct cm2_conv_desc {
  int32_t dtype;     /* CM2_F32 | CM2_BF16: sources, weights, residual; CM2_F16 (TC engine): f16 sources
                        and weights, f32 residual                                                 */
  int32_t out_dtype; /* CM2_F32 | CM2_BF16: output                                                */
  int32_t engine;    /* CM2_ENGINE_*                                                              */
  int32_t num_src;
  cm2_act src[CM2_MAX_SRC]; /* same n, h, w for all sources                                       */
  int32_t cout, kh, kw, stride, pad;
  const void* weight;
  const float* scale; /* device [cout] or NULL (1)                                                */
  const float* shift; /* device [cout] or NULL (0)                                                */
  int32_t relu;       /* ReLU after scale/shift/residual                                          */
  int32_t in_relu;    /* ReLU on the input while loading (fpn.py:34, P7 = conv(relu(P6))); SIMT   */
  cm2_act residual;   /* data == NULL: none.  dtype `dtype`                                       */
  int32_t res_mode;   /* 1: same extent as out   2: [n,ho/2,wo/2,cout] read with nearest-2x
                         upsample ([d2] FPN top-down path)                                        */
  int32_t out_mode;   /* 0: out is [n,ho,wo,cout]
                         1: 2x2/stride-2 transposed-conv scatter: GEMM column j = (dy*2+dx)*(cout/4)
                            + co goes to out[n, 2y+dy, 2x+dx, co]  (out is [n,2ho,2wo,cout/4])
                         2: phase-split store (TC engine): out is plane 0 of four phase planes,
                            [n, ceil(ho/2), ceil(wo/2), cout]; pixel (y,x) goes to plane
                            (y&1)*2+(x&1) at (y>>1, x>>1)
                         3: (TC engine) out_mode 1 fused with the class-gathered mask predictor
                            (sam.py:74-83, 96-97; mask_head.py:196-216): out is f32 [n, 2ho, 2wo, 1] =
                            sigmoid(pred_w[cls_n] . relu(deconv)[n, y, x, :] + pred_b[cls_n]); the deconv
                            output is rounded to bf16 in registers and never stored                       */
  cm2_act out;
  /* optional fused statistics of the stored (post-activation; bf16-rounded for a bf16 output) interior outputs,
   * fp64, zeroed by the call and accumulated by the epilogue (see stats_mode below).  NULL: off.  TC engine only,
   * out_mode 0, no residual.  Image index runs over all images of all segments for segmented tensors. */
  void* stats;
  /* 1: every source is stored as four stride-2 *phase planes* (see cm2_phase_split): src[i] is the
   * halo-1 interior view of plane 0, [n, ceil(H/2), ceil(W/2), c], plane q = (y&1)*2 + (x&1) starts
   * n*sn elements after plane q-1.  Requires a 3x3 / stride 2 / pad 1 convolution; TC engine only. */
  int32_t src_phase;
  /* > 0: sources and output are segmented halo tensors sharing the segment table `seg` (TC engine, stride 1,
   * out_mode 0, no residual): src[i].data / out.data address flat row 0, src[i].c / out.c are the channel counts
   * (= row pitch in elements); the other cm2_act fields are ignored.  One launch then covers all segments. */
  int32_t num_seg;
  cm2_seg seg[CM2_MAX_SEG];
  /* 1: stats = double [images][cout]          per-channel sums (eSE global pool, vovnet.py:254)
   * 2: stats = double [images][cout/8][2]     (sum, sum of squares) per 8-channel chunk (GroupNorm, fcos.py:182;
   *                                           consumed by cm2_groupnorm_apply_seg)                           */
  int32_t stats_mode;
  int32_t pred_ncls;        /* out_mode 3: number of predictor classes                                   */
  const float* pred_w;      /* out_mode 3: device [pred_ncls][cout/4]                                    */
  const float* pred_b;      /* out_mode 3: device [pred_ncls]                                            */
  const int64_t* pred_cls;  /* out_mode 3: device [n] class per image (ROI); clamped to [0, pred_ncls)   */
  /* Split-K (TC engine, bf16, out_mode 0, no residual / statistics / segments, cout % 16 == 0).  splitk >= 2: the K loop of
   * every output tile is cut into (at most) `splitk` slices that run as separate tiles -- for layers whose few output tiles
   * cannot fill the SMs (MaskIoU linear layers, P6 / P7, late stages at small batch).  fp32 partial sums go to `splitk_ws`
   * (device, 32-byte aligned, >= splitk * out.n * out.sn * 4 bytes) and are reduced in a fixed order by a second kernel that
   * applies scale / shift / ReLU: the result does not depend on scheduling.  0 / 1: off. */
  int32_t splitk;
  void* splitk_ws;
  int64_t splitk_ws_bytes;
} cm2_conv_desc;

int cm2_conv2d(const cm2_conv_desc* d, void* stream);
/* Split an fp32 tensor for the split-precision convolution: x is a flat array of `pixels` rows of `c` channels
 * (any NHWC buffer, halo included: zeros stay zeros); out is f16 [pixels][2c] with out[p][ch] = half(x[p][ch]) and
 * out[p][c + ch] = half(x[p][ch] - float(out[p][ch])).  c % 8 == 0, both pointers 16-byte aligned.  |x| must be
 * below the half range (65504); larger magnitudes saturate to +-inf and are reported by nobody. */
int cm2_split_f16x2(const float* x, void* out, int64_t pixels, int32_t c, void* stream);
/* K extent of the TC weight layout for this source list. */
int64_t cm2_conv_tc_klen(int32_t kh, int32_t kw, int32_t num_src, const int32_t* src_c);
/* 1 if the TC engine accepts this descriptor, 0 otherwise (message via cm2_last_error). */
int cm2_conv_tc_supported(const cm2_conv_desc* d);

/* ---------------------------------------------------------------------------------------------
 * Input side.  GeneralizedRCNN.preprocess_image [d2], restated in-tree at
 * /root/reference/deploy_utils.py:76-98: (x - mean) / std, zero pad right/bottom.
 * img: device CHW [3,h,w] (CM2_F32 or CM2_U8); out: view of image `out_index` is written for the
 * whole padded extent out.h x out.w (pixels beyond h, w are zero).
 * ------------------------------------------------------------------------------------------- */
int cm2_preprocess_image(const void* img, int32_t in_dtype, int32_t h, int32_t w, const float* mean3,
                         const float* std3, const cm2_act* out, int32_t out_dtype, int32_t out_index,
                         void* stream);

/* Fused input side for the tensor-core path: normalise + zero-pad (as above) + im2col of the first
 * stem convolution (vovnet.py:409: 3x3, stride 2, pad 1, Cin 3).  out: view [n, hp/2, wp/2, 32] bf16,
 * channel k = (ky*3 + kx)*3 + c for k < 27, zero for k >= 27; image `out_index` is written.  stem_1
 * then is a 1x1 convolution with K = 32 on the TC engine. */
int cm2_preprocess_im2col(const void* img, int32_t in_dtype, int32_t h, int32_t w, int32_t hp, int32_t wp,
                          const float* mean3, const float* std3, const cm2_act* out, int32_t out_index,
                          void* stream);
/* Whole batch in one launch (n <= CM2_MAX_BATCH_PTRS per launch; more are chunked): imgs / hs / ws are HOST arrays
 * of n device pointers ([3, hs[i], ws[i]] planar images) and extents; image i goes to out[out_index0 + i]. */
#define CM2_MAX_BATCH_PTRS 32
int cm2_preprocess_im2col_batch(const void* const* imgs, const int32_t* hs, const int32_t* ws, int32_t n, int32_t in_dtype,
                                int32_t hp, int32_t wp, const float* mean3, const float* std3, const cm2_act* out,
                                int32_t out_index0, void* stream);

/* stem_1 fused with the input side (csrc/stem.cu): normalise + zero-pad + Conv2d(3, 64, 3, stride 2, pad 1) + per-channel
 * scale / shift (folded FrozenBN) + optional ReLU in one pass over the raw planar images (deploy_utils.py:76-98,
 * vovnet.py:205-236, :392-400); replaces cm2_preprocess_im2col_batch + the K = 32 GEMM.  imgs / hs / ws as above (CM2_U8 or
 * CM2_F32).  w30: device bf16 [64][32], K-major, k = ky*10 + kx*3 + c (weights of k = 9, 19, 29, 30, 31 must be zero).
 * out: bf16 view [>= out_index0 + n, hp/2, wp/2, 64]; only interior pixels are written (the halo of a halo buffer is left
 * untouched).  scale / shift: device fp32 [64] or NULL. */
int cm2_stem1_fused_batch(const void* const* imgs, const int32_t* hs, const int32_t* ws, int32_t n, int32_t in_dtype,
                          int32_t hp, int32_t wp, const float* mean3, const float* std3, const void* w30,
                          const float* scale, const float* shift, int32_t relu, const cm2_act* out, int32_t out_index0,
                          void* stream);

/* The same for the fp32 engine ("Split precision" above): the normalised fp32 input is split into an f16 pair inside the kernel,
 * w30_hi_lo is device f16 [2][64][32] = W_hi, W_lo of s * W (s a per-channel power of two; scale = folded BN scale / s), and the
 * result is written as the [hi | lo] f16 operand pair of stem_2: out_split is an f16 view [>= out_index0 + n, hp/2, wp/2, 128]
 * (hi at channel co, lo at 64 + co). */
int cm2_stem1_fused_split_batch(const void* const* imgs, const int32_t* hs, const int32_t* ws, int32_t n, int32_t in_dtype,
                                int32_t hp, int32_t wp, const float* mean3, const float* std3, const void* w30_hi_lo,
                                const float* scale, const float* shift, int32_t relu, const cm2_act* out_split,
                                int32_t out_index0, void* stream);

/* Depthwise 3x3 convolution, padding 1, stride 1 or 2, no bias: the "dw_conv3x3" half of the depthwise bodies'
 * units (vovnet.py:110-130, Conv2d(c, c, 3, groups=c)); the pointwise 1x1 + FrozenBN + ReLU that follows is a
 * cm2_conv_nhwc call.  w: device fp32 [9][c] (tap-major: w[(ky*3+kx)*c + ch] = weight[ch][0][ky][kx]).
 * out extent = ((h-1)/stride+1, (w-1)/stride+1). */
int cm2_dwconv3x3(const cm2_act* in, const cm2_act* out, int32_t dtype, const float* w, int32_t stride, void* stream);

/* MaxPool2d(3, stride 2, ceil_mode=True), vovnet.py:349-350. */
int cm2_maxpool3x3s2_ceil(const cm2_act* in, const cm2_act* out, int32_t dtype, void* stream);

/* eSE (vovnet.py:238-260, applied at :327-330):
 *   pooled[n,c] = mean_hw x ;  gate = relu6(W.pooled + b + 3) / 6 ;  out = x*gate (+ identity).
 * cm2_ese_pool needs a float workspace of n*chunks*c floats, chunks = cm2_ese_pool_chunks(h*w).
 * cm2_ese_gate: `pooled` holds means when inv_count == 1, or raw sums with inv_count = 1/(h*w). */
int32_t cm2_ese_pool_chunks(int32_t hw);
int cm2_ese_pool(const cm2_act* x, int32_t dtype, float* workspace, float* pooled, void* stream);
int cm2_ese_gate(const float* pooled, float inv_count, const float* fc_w, const float* fc_b, float* gate,
                 int32_t n, int32_t c, void* stream);
int cm2_ese_apply(const cm2_act* x, const float* gate, const cm2_act* identity, const cm2_act* out,
                  int32_t dtype, void* stream);
/* Fused variants used with the tensor-core engine.
 * cm2_ese_gate_f64: as cm2_ese_gate, from the fp64 channel sums the producing convolution accumulated in its
 *   epilogue (cm2_conv_desc.stats, stats_mode 1); inv_count = 1/(h*w).
 * cm2_ese_apply_pool: y = x * gate (+ identity), stored to `full` (optional) and reduced by the
 *   MaxPool2d(3, 2, ceil_mode=True) that opens the next stage (vovnet.py:349-350) into `pool` (optional,
 *   [n, ho, wo, c]) in the same pass -- max over the values as stored.  Without `pool`, x / identity / full
 *   must be interior views of identically shaped one-pixel-halo buffers (flat streaming pass). */
int cm2_ese_gate_f64(const double* sums, double inv_count, const float* fc_w, const float* fc_b, float* gate,
                     int32_t n, int32_t c, void* stream);
int cm2_ese_apply_pool(const cm2_act* x, const float* gate, const cm2_act* identity, const cm2_act* full,
                       const cm2_act* pool, int32_t dtype, void* stream);

/* GroupNorm(groups, c) + optional ReLU in place; fcos.py:182-185 (eps 1e-5, biased variance over
 * (c/groups)*h*w per sample).  workspace: cm2_gn_workspace_floats(n, h*w, c, groups) floats. */
int64_t cm2_gn_workspace_floats(int32_t n, int32_t hw, int32_t c, int32_t groups);
int cm2_groupnorm_relu(const cm2_act* x, int32_t dtype, int32_t groups, const float* gamma,
                       const float* beta, float eps, int32_t relu, float* workspace, void* stream);

/* Phase-split copy (optionally with ReLU, fpn.py:34): in [n,h,w,c] -> four planes, out = plane 0 view
 * [n, ceil(h/2), ceil(w/2), c] (planes n*sn elements apart); pixel (y,x) -> plane (y&1)*2+(x&1) at
 * (y>>1, x>>1).  Positions of a plane that no input pixel maps to are left untouched (zero). */
int cm2_phase_split(const cm2_act* in, const cm2_act* out_plane0, int32_t dtype, int32_t relu, void* stream);

/* The same on a segmented halo tensor (see cm2_seg): one launch set for all segments; statistics are per
 * (segment, image, group).  x: flat [rows, c].  workspace: cm2_gn_seg_workspace_floats floats. */
int64_t cm2_gn_seg_workspace_floats(int32_t num_seg, const cm2_seg* seg, int32_t c, int32_t groups);
int cm2_groupnorm_relu_seg(void* x, int32_t dtype, int32_t c, int32_t num_seg, const cm2_seg* seg, int32_t groups,
                           const float* gamma, const float* beta, float eps, int32_t relu, float* workspace,
                           void* stream);
/* Normalise (+ReLU) in place from the statistics the producing convolution accumulated in its epilogue
 * (cm2_conv_desc.stats, stats_mode 2: double [images][c/8][2]); needs (c / groups) % 8 == 0.  One pass. */
int cm2_groupnorm_apply_seg(void* x, int32_t dtype, int32_t c, int32_t num_seg, const cm2_seg* seg, int32_t groups,
                            const float* gamma, const float* beta, float eps, int32_t relu, const double* stats,
                            void* stream);
/* fp32 engine: cm2_groupnorm_apply_seg reading the fp32 convolution output `x` and writing normalise + ReLU straight into
 * the [hi | lo] f16 operand tensor of the next convolution (out_split: f16 [rows][2c], same segment table, zero-initialised
 * by the caller: only interior pixels are written) -- see "Split precision" above. */
int cm2_groupnorm_apply_seg_split(const float* x, void* out_split, int32_t c, int32_t num_seg, const cm2_seg* seg,
                                  int32_t groups, const float* gamma, const float* beta, float eps, int32_t relu,
                                  const double* stats, void* stream);

/* Elementwise ReLU copy (P7 input when the conv engine cannot apply in_relu). */
int cm2_relu(const cm2_act* in, const cm2_act* out, int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FCOS post-process.  fcos/fcos_outputs.py:372-495.
 *
 * cm2_fcos_decode (forward_for_single_feature_map :396-466, one FPN level, all images):
 *   p = sigmoid(logit); c = sigmoid(ctr); candidate iff p > thresh (p*c > thresh when
 *   thresh_with_ctr); raw score = p*c; box = (x-l, y-t, x+r, y+b) with (l,t,r,b) = reg*stride,
 *   (x,y) = (col*stride + stride/2, row*stride + stride/2)  (fcos.py:132-144).
 *   logits: f32 view [n,h,w,ncls];  regctr: f32 view [n,h,w,>=5] with channels (l,t,r,b,ctr,...) where
 *   l..b are the raw bbox_pred outputs: the per-level Scale and the ReLU (fcos.py:233-238) are applied here,
 *   (l,t,r,b) = relu(reg * reg_scale) * stride.
 *   Candidates are appended (unordered) to segment (image, level) of the candidate arrays, each
 *   of capacity `cap`; cand_count[image*num_levels+level] counts *all* candidates, so a value > cap
 *   signals overflow to the caller.  cand_count must be zeroed by the caller.
 *
 * cm2_fcos_select (:444-449 upstream top-k, select_over_all_levels :468-495, ml_nms layers/ml_nms.py:
 *   93-96): per (image, level) keep the `pre_topk` best raw scores, then per image class-aware greedy
 *   NMS in descending sqrt(score) order (suppress when IoU > nms_thresh, torchvision arithmetic) and
 *   keep the first `post_topk` survivors.  Outputs are fixed-size [n][post_topk] with det_count[n].
 *   workspace bytes: cm2_fcos_select_workspace(n, num_levels, cap).
 * ------------------------------------------------------------------------------------------- */
typedef struct cm2_cand_buffers {
  float* boxes;    /* [n][levels][cap][4] */
  float* score;    /* [n][levels][cap]   raw p*c */
  int32_t* cls;    /* [n][levels][cap] */
  int32_t* flat;   /* [n][levels][cap]   loc*ncls + cls */
  int32_t* count;  /* [n][levels] */
} cm2_cand_buffers;

int cm2_fcos_decode(const cm2_act* logits, const cm2_act* regctr, int32_t stride, float reg_scale, float thresh,
                    int32_t thresh_with_ctr, int32_t level, int32_t num_levels, int32_t cap,
                    const cm2_cand_buffers* cand, void* stream);
/* All levels in one launch: logits / regctr / strides / reg_scales are HOST arrays of num_levels entries (<= 8);
 * level l writes candidate segment (image, l).  Same semantics as num_levels calls of cm2_fcos_decode. */
int cm2_fcos_decode_levels(const cm2_act* logits, const cm2_act* regctr, const int32_t* strides, const float* reg_scales,
                           int32_t num_levels, float thresh, int32_t thresh_with_ctr, int32_t cap,
                           const cm2_cand_buffers* cand, void* stream);

int64_t cm2_fcos_select_workspace(int32_t n, int32_t num_levels, int32_t cap);

typedef struct cm2_det_buffers {
  float* boxes;      /* [n][post_topk][4] */
  float* scores;     /* [n][post_topk]  sqrt(p*c) */
  int64_t* classes;  /* [n][post_topk] */
  float* locations;  /* [n][post_topk][2] */
  int32_t* count;    /* [n] */
} cm2_det_buffers;

int cm2_fcos_select(const cm2_cand_buffers* cand, int32_t n, int32_t num_levels, int32_t cap,
                    const int32_t* level_w, const int32_t* level_stride, int32_t ncls,
                    int32_t pre_topk, float nms_thresh, int32_t post_topk,
                    const cm2_det_buffers* det, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ROI stage.
 *
 * cm2_roialign_fpn: centermask/pooler.py:320-366 with assign_boxes_to_levels_by_ratio (:80-118,
 *   crit 0) or assign_boxes_to_levels (:121-152, crit 1) fused into [d2] ROIAlign (= torchvision
 *   roi_align, aligned=True, sampling_ratio as given; 0 = adaptive).  ROI slots are [n][r_cap] with
 *   det_count[n] valid ones per image; out is a view [n*r_cap, res, res, c] (invalid slots are
 *   written with zeros).  image_area: device float [n] = h*w of the unpadded image (pooler.py:70-77).
 *   level_out: device int32 [n*r_cap] (may be NULL).
 *   workspace: device scratch of n*r_cap int32 (contents overwritten: the ROI slots in launch order,
 *   largest boxes first, for the column-walk kernel); may be NULL, then the slower CTA-per-ROI
 *   kernel runs.  Results do not depend on it.
 * ------------------------------------------------------------------------------------------- */
int cm2_roialign_fpn(const cm2_act* feats, const int32_t* feat_stride, int32_t num_levels, int32_t dtype,
                     const float* boxes, const int32_t* det_count, int32_t n, int32_t r_cap,
                     const float* image_area, int32_t crit, int32_t sampling_ratio,
                     const cm2_act* out, int32_t* level_out, void* workspace, void* stream);

/* SpatialAttention, centermask/sam.py:23-28: x * sigmoid(conv3x3([mean_c x, max_c x])); views
 * [r,s,s,c]; w18 = conv weight [1][2][3][3] flattened (device). */
int cm2_spatial_attention(const cm2_act* x, const cm2_act* out, int32_t dtype, const float* w18,
                          void* stream);

/* predictor (sam.py:83,97) restricted to the predicted class + mask_rcnn_inference
 * (mask_head.py:196-216): probs[r, y, x] = sigmoid(x[r,y,x,:] . wp[cls_r,:] + bp[cls_r]).
 * x view [r, m, m, c]; wp f32 [ncls][c]; classes int64 [r]; probs f32 dense [r][m][m]. */
int cm2_mask_predict(const cm2_act* x, int32_t dtype, const float* wp, const float* bp,
                     const int64_t* classes, int32_t ncls, float* probs, void* stream);

/* MaskIoU input (maskiou_head.py:108-112): 2x2 max-pool of probs [r,2s,2s] written to channel 0 of
 * out view [r,s,s,cpad]; remaining channels zero. */
int cm2_maskiou_prep(const float* probs, const cm2_act* out, int32_t dtype, void* stream);

/* mask_iou_inference (maskiou_head.py:50-60): mask_scores[r] = scores[r] * iou[r, cls_r];
 * iou dense [r][ncls] of `dtype`. */
int cm2_maskiou_score(const void* iou, int32_t dtype, int32_t r, int32_t ncls, const int64_t* classes,
                      const float* scores, float* mask_scores, void* stream);

/* Keypoint branch: the tail of KRCNNConvDeconvUpsampleHead.layers -- interpolate(x, 2, "bilinear",
 * align_corners=False), keypoint_head.py:221 -- fused with keypoint_rcnn_inference (keypoint_head.py:95-120) ->
 * detectron2 heatmaps_to_keypoints [d2]: per ROI and keypoint, the 4res x 4res logit map is resized (bicubic,
 * align_corners=False) to ceil(h) x ceil(w) pixels of the box (h, w clamped to >= 1), and the first arg-max pixel gives
 * out[slot][kp] = (x, y, logit, score): x = (x_int + 0.5) * w / ceil(w) + x0 (same for y), score = 1 / sum over the
 * 4res x 4res map of exp(map - max).  pred_keypoints of the reference = columns (0, 1, 3).
 *   lowres: f32 dense [n*r_cap][res][res][4][k] = the ConvTranspose2d(k 4, s 2, p 1) output (keypoint_head.py:205-208)
 *           in phase layout: value at (y, x) of the 2res x 2res map sits at [y / 2][x / 2][(y & 1) * 2 + (x & 1)].
 *   boxes : f32 [n][r_cap][4] (16-byte aligned), det_count int32 [n]; slots beyond det_count are written with zeros.
 *   out   : f32 [n*r_cap][k][4].  res <= 24. */
int cm2_keypoints_decode(const float* lowres, const float* boxes, const int32_t* det_count, int32_t n, int32_t r_cap,
                         int32_t res, int32_t num_keypoints, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input side, upstream of preprocess: detectron2's ResizeShortestEdge / ResizeTransform for
 * uint8 HWC images (/root/reference/deploy_utils.py:60-73), i.e. PIL.Image.resize(BILINEAR): Pillow's two-pass
 * fixed-point resampler.  bounds_* [out][2] = (first input index, tap count), kk_* [out][ksize_*] = 22-bit
 * fixed-point taps, both built by centermask2_b200/transforms.py::pil_bilinear_coeffs (device memory).
 * src [h][w][c] -> tmp [h][ow][c] (horizontal pass) -> dst [oh][ow][c] (chw == 0) or [c][oh][ow] (chw == 1).
 * Bit-identical to Pillow 12.2.0.
 * ------------------------------------------------------------------------------------------- */
int cm2_resize_pil_u8(const uint8_t* src, uint8_t* tmp, uint8_t* dst, int32_t h, int32_t w, int32_t c, int32_t oh,
                      int32_t ow, const int32_t* bounds_x, const int32_t* kk_x, int32_t ksize_x,
                      const int32_t* bounds_y, const int32_t* kk_y, int32_t ksize_y, int32_t chw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Output side.  detector_postprocess + paste_masks_in_image [d2] (the fork's
 * restatement is /root/reference/deploy_utils.py:129-158).
 * cm2_scale_clip_boxes: boxes *= (sx, sy); clip to [0,out_w]x[0,out_h]; valid = w>0 && h>0.
 * cm2_paste_masks: bilinear grid_sample (zero padding, align_corners=False) of probs [r,m,m] at the
 *   pixel centres of out [r,out_h,out_w] (uint8 0/1), `>= threshold`, restricted to the window
 *   [floor(x0)-1, ceil(x1)+1) x [floor(y0)-1, ceil(y1)+1); rows of invalid ROIs are zero.
 * ------------------------------------------------------------------------------------------- */
int cm2_scale_clip_boxes(const float* boxes_in, float* boxes_out, uint8_t* valid, int32_t r, float sx,
                         float sy, float out_w, float out_h, void* stream);
/* Whole batch in one launch: boxes [n][r_cap][4]; params [n][4] = (sx, sy, out_w, out_h) in DEVICE memory.  det_count
 * (device int32 [n], may be NULL): slots >= det_count[image] hold no detection: box 0, valid 0 (so no mask is pasted). */
int cm2_scale_clip_boxes_batch(const float* boxes_in, float* boxes_out, uint8_t* valid, int32_t n, int32_t r_cap,
                               const float* params, const int32_t* det_count, void* stream);
int cm2_paste_masks(const float* probs, const float* boxes, const uint8_t* valid, uint8_t* out,
                    int32_t r, int32_t m, int32_t out_h, int32_t out_w, float threshold, void* stream);
/* Result record of every detection slot for the host / the final gather: records float32
 * [n][r_cap][11] = (x0, y0, x1, y1, score, class, mask score, location x, location y, valid, detections of the image);
 * fields 0..9 of slots >= det_count[image] are zero.  boxes [n][r_cap][4] (16-byte aligned), scores / mask_scores (may be
 * NULL) [n][r_cap], classes int64 [n][r_cap], locations (may be NULL) [n][r_cap][2], valid (may be NULL) uint8 [n][r_cap]. */
int cm2_pack_records(const float* boxes, const float* scores, const int64_t* classes, const float* mask_scores,
                     const float* locations, const uint8_t* valid, const int32_t* det_count, int32_t n, int32_t r_cap,
                     float* records, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Result encoding: COCO run-length encoding of the pasted masks, i.e. pycocotools' rleEncode as
 * called per mask by instances_to_coco_json (/root/reference/centermask2/centermask/evaluation/coco_evaluation.py:388-391).
 * masks: uint8 [r][h][w] (0 / non-zero).  Runs are counted in column-major order and start with a run of zeros.
 * cm2_rle_count: col_count / col_offset int32 [r][w] scratch; total[m] = number of value changes of mask m, so mask m
 *   has total[m] + 1 runs.  The caller then builds mask_offset (exclusive scan of total + 1) and sizes the outputs.
 * cm2_rle_write: positions (scratch) and runs: uint32 [sum(total + 1)]; runs[mask_offset[m] + k] = length of run k.
 * The LEB128-like string compression (rleToString) is done on the host (centermask2_b200/rle.py).
 * ------------------------------------------------------------------------------------------- */
int cm2_rle_count(const uint8_t* masks, int32_t r, int32_t h, int32_t w, int32_t* col_count, int32_t* col_offset,
                  int32_t* total, void* stream);
int cm2_rle_write(const uint8_t* masks, int32_t r, int32_t h, int32_t w, const int32_t* col_offset, const int32_t* total,
                  const int64_t* mask_offset, uint32_t* positions, uint32_t* runs, void* stream);
/* Both steps in one call with the offsets made on the device (no host round trip between them; for pipelined callers):
 * mask_offset int64 [r + 1] = exclusive scan of (total + 1), mask_offset[r] = number of runs of all masks.  positions / runs
 * have room for `capacity` entries; a mask whose runs would end beyond it is skipped, which the caller detects as
 * mask_offset[r] > capacity after the fact (and retries with larger buffers).
 * boxes / valid (may be NULL): the boxes [r][4] (16-byte aligned) and validity flags the masks were pasted with
 * (cm2_paste_masks): a pasted mask is zero outside the dilated window of its box, so only that window is scanned. */
int cm2_rle_encode(const uint8_t* masks, int32_t r, int32_t h, int32_t w, int32_t* col_count, int32_t* col_offset,
                   int32_t* total, int64_t* mask_offset, uint32_t* positions, uint32_t* runs, int64_t capacity,
                   const float* boxes, const uint8_t* valid, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CM2_H_ */
