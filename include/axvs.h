/* axvs.h -- C ABI of libaxvs.so: B200 (sm_100a) kernels for Axial-VS / MaXTron's axial-trajectory attention.
 *
 * Drop-in boundary.  The reference has no FFI of its own for this path: it is eager PyTorch
 * (nn.Linear / einsum / softmax / LayerNorm).  Each entry point below names the reference lines it
 * replaces ("Vk/" = MaXTron_Video-kMaX/, "WC/" = Vk/maxtron_deeplab/modeling/within_clip_tracking_module/,
 * "CC" = Vk/maxtron_deeplab/modeling/cross_clip_tracking_module/maxtron_cross_clip_tracking_module.py,
 * "TL/" = MaXTron_Tube-Link/).  The convention mirrors the reference's one native op
 * (WC/ops/src/cuda/ms_deform_attn_cuda.cu:25-85): contiguous device tensors in, work enqueued on the caller's
 * stream, nothing retained after the call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer in the current CUDA context unless it says "host";
 *   - `stream` is a cudaStream_t / CUstream passed as void* (0 = legacy default stream);
 *   - return 0 on success, a negative AXVS_E_* code otherwise; axvs_last_error() gives the message
 *     (thread-local).  No C++ exception crosses this boundary.  No host synchronisation, no allocation:
 *     the caller owns all buffers including the workspace -> every call is CUDA-graph capturable;
 *   - fp32 tensors are the reference's tensors (row-major, channels last, C = 256); bf16 buffers are internal;
 *   - the path is specialised for d_model C = 256, 8 heads of 32 (every shipped config: Vk config.py:21-34,
 *     CC:236-246, TL configs embed_dims=256 num_heads=8); other sizes return AXVS_E_UNSUPPORTED.
 */
#ifndef AXVS_H_
#define AXVS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AXVS_OK 0
#define AXVS_E_INVALID (-1)     /* bad argument (null pointer, non-positive size, misalignment) */
#define AXVS_E_UNSUPPORTED (-2) /* shape outside the specialisation (C != 256, K % 64, n_out % 256 ...) */
#define AXVS_E_WORKSPACE (-3)   /* workspace too small */
#define AXVS_E_CUDA (-4)        /* CUDA runtime error at launch */

typedef void* axvs_stream_t;

/* Row order of the sequences handed to a trajectory attention call. */
#define AXVS_AXIS_NONE 0   /* rows already in sequence order (cross-clip, non-axial "trajectory" layer)      */
#define AXVS_AXIS_H 1      /* height pass: '(B T)(H W) C -> (B W)(T H) C'   WC/temporal_attention.py:197     */
#define AXVS_AXIS_W 2      /* width pass:  '... -> (B H)(T W) C'            WC/temporal_attention.py:206     */

int axvs_version(void);
const char* axvs_last_error(void);
/* Hash of the sources this library was compiled from (-DAXVS_BUILD_ID, set by __graft_entry__.build); "unknown" for ad-hoc builds.
 * smoke() compares it with the hash of the tree it runs in, so a stale prebuilt binary is detected. */
const char* axvs_build_id(void);

/* Fusion level of the composite entry points (process-wide; default = 4, the fastest measured).
 *   0: one kernel per reference op group (tcgen05 GEMMs + attention + SIMT helpers; validation baseline)
 *   1: + proj_q / proj_kv / temporal softmax / proj / residual fused in one tcgen05 kernel
 *   2: + LayerNorm1 / FFN / residual / LayerNorm2 fused in one tcgen05 kernel
 *   3: + TMA-fed q|k|v projection with head-major output, one-shot per-frame attention writing UMMA tile images
 *   4: + the q|k|v projection reads the fp32 residual stream (+ pos) itself (permute, add and cast inside its A producers);
 *        every UMMA A operand lives in tensor memory
 * Returns the previous level; values outside the range are clamped. */
int axvs_set_fusion(int level);
/* CTA-pair kernels (tcgen05 cta_group::2: two SMs of a cluster share one M = 256 instruction stream, each staging half of every
 * weight unit, which halves the shared-memory traffic of the GEMM core: 512 instead of 670 clk per 32 KiB unit).  Bit mask:
 *   2 = fused temporal kernel (traj_pair_kernel), 4 = q|k|v projection (qkv_pair_kernel) and the fused front end of the MSDeformAttn
 *   layer (msda_front_pair_kernel), 8 = FFN (ffn_n256_pair_kernel); every pair kernel is bit-identical to its single-CTA
 *   counterpart (which `0` selects; the MSDeformAttn front end falls back to two generic GEMMs).
 *   16 = frame-major row order between the attention kernel and the temporal kernel: the per-frame attention outputs x_f
 *   (WC/temporal_attention.py:56) are stored with all tokens of frame 0 first, then frame 1, ... (each group padded to whole 128-row
 *   tiles), so the x_t tile of a tile of frame-t tokens IS its x_diag (WC/temporal_attention.py:61-63) and no x_diag image is written
 *   or read (0.5 KiB per token less each way); results are bit-identical to the pass-order layout.
 *   Default 30.  Returns the previous mask. */
int axvs_set_pair_mode(int on);
/* Attention core of the per-frame spatial attention (WC/temporal_attention.py:47-60) at fusion level >= 4.
 *   1 (default): tcgen05 kernel -- Q K^T and P V as UMMAs (scores / probabilities / outputs in tensor memory, operands by TMA),
 *                softmax thread-per-row in registers (csrc/attn_tc.cuh); frames of up to 224 tokens, any sequence length
 *   0: the mma.sync kernels (csrc/attn.cuh; validation baseline, and the fallback for longer frames).  Returns previous. */
int axvs_set_attn_core(int core);

/* ---- weights ------------------------------------------------------------------------------------------------
 * nn.Linear weights [n_out, k] fp32 are converted once to bf16 and laid out as the shared-memory image the
 * tensor-core kernels consume (K-blocks of 64, 8-row x 128-byte swizzle atoms), so that one TMA bulk copy
 * brings a weight tile in.  n_out % 8 == 0, k % 64 == 0. */
size_t axvs_packed_weight_bytes(int n_out, int k);
int axvs_pack_weight(const float* w, int n_out, int k, void* packed, axvs_stream_t stream);
/* "Unit" format of the fused kernels: 32 KiB units = [2 K-blocks][128 rows x 128 B swizzled] so one TMA bulk copy feeds
 * eight tensor-core instructions.  k_major = 0: units ordered (row tile, K group); 1: (K group, row tile);
 * 2: units of [256 rows x 128 B swizzled] of ONE K-block (B operand of an N = 256 instruction), ordered (row tile of 256, K-block).
 * n_out % 128 == 0 (256 for mode 2), k % 128 == 0.  Same size as axvs_packed_weight_bytes. */
int axvs_pack_weight_units(const float* w, int n_out, int k, int k_major, void* packed, axvs_stream_t stream);

/* One TrajectoryAttention's parameters (WC/temporal_attention.py:27-33; CC:85-89; TL .../msdeformattn_pixel_decoder.py:659-665).
 * w_qkv packs rows [Wq; Wk; Wv] (or the fused `qkv` Linear of the cross-clip variant): [768, 256]. */
typedef struct axvs_ta_weights {
  const void* w_qkv;  const float* b_qkv;    /* packed [768,256], bias [768] */
  const void* w_pq;   const float* b_pq;     /* proj_q   [256,256]           */
  const void* w_pkv;  const float* b_pkv;    /* proj_kv  [512,256]           */
  /* unit-format copies for the fused kernel (axvs_pack_weight_units, k_major = 0) */
  const void* w_qkv_u;                       /* [Wq; Wk; Wv] (768 x 256)                                     */
  const void* w_pq_u;                        /* proj_q                                                       */
  const void* w_pkv_u;                       /* proj_kv with rows re-ordered per head pair c = 0..3:
                                                [k2 rows 64c..64c+63 ; v2 rows 256+64c..256+64c+63]          */
  const void* w_proj_u;                      /* proj                                                         */
  const void* w_proj; const float* b_proj;   /* proj     [256,256]           */
} axvs_ta_weights;

/* One Temporal(Axial)TrajectoryAttentionLayer (WC/temporal_attention.py:159-175 / :104-119). */
typedef struct axvs_layer_weights {
  axvs_ta_weights attn_h;                    /* height_attn  (or temporal_attn of the non-axial layer) */
  axvs_ta_weights attn_w;                    /* width_attn   (unused for the non-axial layer)          */
  const float* ln1_g; const float* ln1_b;    /* norm1 */
  const void* w_ffn1; const float* b_ffn1;   /* linear1 packed [1024,256] */
  const void* w_ffn2; const float* b_ffn2;   /* linear2 packed [256,1024] */
  const void* w_ffn1_u;                      /* linear1, unit format, k_major = 0 (fused kernel) */
  const void* w_ffn2_u;                      /* linear2, unit format, k_major = 1 (fused kernel) */
  const float* ln2_g; const float* ln2_b;    /* norm2 */
  int d_ffn;                                 /* 1024 */
  const void* w_ffn1_n;                      /* linear1, unit format mode 2 (N = 256 units; may be NULL: falls back to w_ffn1_u) */
} axvs_layer_weights;

/* ---- building blocks (each is also a test surface) ------------------------------------------------------------ */

/* out[M, n_out] = act((A[M,K] @ W^T + bias) * scale) (+ resid), tcgen05 GEMM.  A bf16 row-major (lda elements).
 * out_bf16 != 0 -> bf16 output, else fp32 (+ optional fp32 residual with the same leading dimension).
 * `relu`: 0 = no activation, 1 = ReLU, 2 = GELU (erf form).
 * Replaces every nn.Linear on the path (WC/temporal_attention.py:42-44,64-65,75,182). */
int axvs_linear(const void* a_bf16, int lda, int M, int K, const void* w_packed, const float* bias, int n_out,
                float scale, int relu, void* out, int ldo, int out_bf16, const float* resid, axvs_stream_t stream);

/* Per-frame-softmax spatial attention (WC/temporal_attention.py:47-60; CC:101-110).
 * qkv bf16 [num_seq*N, 768] (q | k | v, heads of 32), N = F*n.  x bf16 [num_seq*N, F, 256]. */
int axvs_spatial_attention(const void* qkv_bf16, void* x_bf16, int num_seq, int F, int n, axvs_stream_t stream);

/* Full TrajectoryAttention forward + residual:
 *     out = resid + TA(query = q_in + pos, key = k_in + pos, value = v_in)
 * (WC/temporal_attention.py:35-76 and the residual add :204 / :213; CC:91-130,158-159 with pos = NULL and
 * q_in = k_in = v_in).  q_in/k_in/v_in/pos/resid/out fp32 canonical [(B T)(H W), 256]; the layer passes
 * q_in = k_in = v_in = src (:200-202).  `axis` selects how canonical tokens form sequences:
 *   AXVS_AXIS_H: B*W sequences of T*H tokens, AXVS_AXIS_W: B*H sequences of T*W tokens,
 *   AXVS_AXIS_NONE: B sequences of T*(H*W) tokens in storage order (frames of n = H*W tokens).
 * pos and resid may be NULL (resid NULL = the bare nn.Module `TrajectoryAttention.forward`).
 * pos_clips = number of clips the pos tensor covers: B, or 1 when one [T, H, W, 256] table is shared by all clips (the reference's
 * PositionEmbeddingSine3D output does not depend on the batch index, WC/pos_embeddings.py:86-130; the kernels then re-read the
 * one table, which stays L2-resident, instead of streaming B copies from HBM). */
size_t axvs_traj_attn_workspace_bytes(int B, int T, int H, int W);
int axvs_traj_attn_fwd(const float* q_in, const float* k_in, const float* v_in, const float* pos, int pos_clips, const float* resid,
                       float* out, const axvs_ta_weights* w, int B, int T, int H, int W, int axis,
                       void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* Attention-map side output of TrajectoryAttention (the `space_attn` return value, WC/temporal_attention.py:54,76; consumed
 * only by the visualiser, Vk/maxtron_deeplab/maxtron_wc_model.py:598-611).  Same inputs / axis convention as
 * axvs_traj_attn_fwd; maps fp32 [(num_seq * 8), N, F, n] with head the fast factor of the fused dim.  Slow path. */
int axvs_traj_attn_maps(const float* q_in, const float* k_in, const float* pos, float* maps, const axvs_ta_weights* w,
                        int B, int T, int H, int W, int axis, void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* y = LayerNorm(x) over C = 256 (nn.LayerNorm, biased variance).  y32 and/or y16 may be NULL. */
int axvs_layernorm(const float* x, const float* gamma, const float* beta, float* y32, void* y16_bf16, int rows,
                   float eps, axvs_stream_t stream);

/* out = LN2(s + W2 relu(W1 s + b1) + b2), s = LN1(x)   (WC/temporal_attention.py:181-185,217-218). */
size_t axvs_ffn_workspace_bytes(int rows, int d_ffn);
int axvs_ln_ffn_fwd(const float* x, float* out, const axvs_layer_weights* w, int rows,
                    void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* ---- layers --------------------------------------------------------------------------------------------------- */

/* TemporalAxialTrajectoryAttentionLayer.forward (WC/temporal_attention.py:187-220; TL copy :753-791):
 * src [(B T), (H W), 256], pos [pos_clips, T, H, W, 256] (pos_clips = B or 1, see axvs_traj_attn_fwd) -> out (same shape as src).  axial = 0 runs the non-axial
 * TemporalTrajectoryAttentionLayer (:131-155) using attn_h only. */
size_t axvs_layer_workspace_bytes(int B, int T, int H, int W, int d_ffn);
int axvs_axial_layer_fwd(const float* src, const float* pos, int pos_clips, float* out, const axvs_layer_weights* w,
                         int B, int T, int H, int W, int axial,
                         void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* ---- cross-clip module, after the trajectory attention (SURVEY.md section 8 rows A9 tail, A10) ------------------------------ */

/* fp32 [rows,256] -> bf16 [rows,256] (A operand of axvs_linear). */
int axvs_cast_bf16(const float* x, void* out_bf16, int rows, axvs_stream_t stream);

/* Temporal ASPP + residual + LayerNorm of the cross-clip layer (CC:176-201, 293-295; TL cc head :929-947):
 *   out = LN( GELU(LN_cf(proj(cat_d conv_d(z)))) + z ),  conv_d = Conv1d(256,256,k=3,dilation d,'same',replicate) over time.
 * x / out fp32 [b*T*Q, 256] with rows ordered (b, t, q) (the trajectory attention's order); out_bf16 optional copy.
 * Conv weights: packed with axvs_pack_weight as [256, 768] where K index = tap*256 + c_in. */
typedef struct axvs_aspp_weights {
  const void* w_conv[3];  const float* b_conv[3];  int dilation[3];
  const void* w_proj;                              /* 1x1 conv [256, 768], no bias */
  const float* lncf_g; const float* lncf_b;        /* channels-first LayerNorm, eps 1e-6 */
  const float* ln_g;   const float* ln_b;          /* conv_norms[i], eps 1e-5 */
  int split;                                       /* 1: w_conv / w_proj are split-precision images [W | W | W - bf16(W)] ([256, 3*768]) and the
                                                    *    GEMMs run on fp32 rows split into bf16 hi / lo halves: fp32-grade results (default of the
                                                    *    Python layer: the rows are few, and these embeddings decide the per-pixel labels) */
} axvs_aspp_weights;
size_t axvs_cc_aspp_workspace_bytes(int rows);
int axvs_cc_aspp_fwd(const float* x, float* out, void* out_bf16, const axvs_aspp_weights* w, int b, int T, int Q,
                     void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* MaXTronCCPredictor class branch (CC:47-50): logit = w_act . ce[t,q,:] + b_act, softmax over the T clips, pooled[q,:] =
 * sum_t a[t,q] ce[t,q,:].  ce bf16 [T*Q, 256] rows (t, q); pooled bf16 [Q, 256]. */
int axvs_cc_class_pool(const void* ce_bf16, const float* w_act, float b_act, void* pooled_bf16, int T, int Q, axvs_stream_t stream);

/* Mask logits (CC:62-69): out[q, t, p] = bn_scale * sum_c pixel[t, c, p] * mk[t*Q + q, c] + bn_shift, c < 128, q < Q <= 128.
 * pixel fp32 [T, 128, P] (P = V*H*W pixels of clip t), mk bf16 rows (t, q) with leading dimension ld_mk (elements),
 * out fp32 [Q, T, P] == the reference's final [1, Q, (T V), H, W]. */
int axvs_mask_einsum(const float* pixel, const void* mk_bf16, int ld_mk, float* out, int T, int Q, int P, float bn_scale,
                     float bn_shift, axvs_stream_t stream);
/* The same contraction with fp32 mask kernels and split-precision products (both operands split into bf16 hi / lo halves in shared
 * memory, three tensor-core products per pair): fp32-grade logits at the same HBM-bound cost.  pixel fp32 [T, channels, P] with
 * channels % 128 == 0 (128 for Video-kMaX, 256 for the Tube-Link mask features: TL cc head :776); mk fp32 rows (t, q), ld_mk % 4 == 0. */
int axvs_mask_einsum_f32(const float* pixel, const float* mk, int ld_mk, float* out, int T, int Q, int P, int channels, float bn_scale,
                         float bn_shift, axvs_stream_t stream);
/* out = act((a W^T + bias) * scale) on fp32 rows a [M, lda] (converted inside the GEMM's producers).  split = 1: w_packed is the
 * split-precision image of [n_out, 3*K] = [W | W | W - bf16(W)] (axvs_pack_weight) and a is split into bf16 hi / lo halves:
 * a_hi W_hi + a_lo W_hi + a_hi W_lo, the fp32 product to ~2^-17.  act: 0 none, 1 ReLU, 2 GELU (erf).  n_out % 256 == 0, K % 64 == 0.
 * Replaces the eval-mode ConvBN 1x1 heads of CC:53, 266-270, 300-301 (batch norm folded into W / bias by the caller). */
int axvs_linear_f32(const float* a, int lda, int M, int K, const void* w_packed, int split, const float* bias, int n_out, float scale, int act,
                    void* out, int ldo, int out_bf16, axvs_stream_t stream);

/* ---- clip-level kMaX decoder attention, query side ----------------------------------------
 * DEC = Vk/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py
 *
 * AttentionOperation.forward (DEC:49-71) in eval mode: logits = einsum('bhdl,bhdm->bhlm', q, k), per-head BatchNorm on the logits,
 * fp32 softmax over m, retrieved = einsum('bhlm,bhdm->bhdl'), per-channel BatchNorm, GELU.  All fp32.
 * q, k [N, heads*16, L]; v, out [N, heads*32, L] (channel = head*depth + d, DEC:215-217).  sim_affine [heads][2] and
 * val_affine [heads*32][2] are the folded eval-mode BN (scale = w / sqrt(var + eps), shift = b - mean*scale). */
int axvs_query_self_attn(const float* q, const float* k, const float* v, const float* sim_affine, const float* val_affine, float* out,
                         int N, int heads, int L, axvs_stream_t stream);

/* Layout / elementwise helpers of the kMaX transformer decoder layer's pixel side (Vk/maxtron_deeplab/modeling/transformer_decoder/
 * maxtron_transformer_decoder.py:75-124, 184-232; the 1x1 convolutions with their folded batch norms run on axvs_linear_f32 over token rows):
 *   axvs_cm_to_rows : x fp32 [N, C, M] (the reference's channel-major tensors) -> rows [N*M, C]; act = 2 applies GELU on the way (:186)
 *   axvs_rows_to_cm : rows [N*M, ld] -> out [N, C, M] (first C <= 256 channels); normalize != 0: F.normalize(p=2, dim=channel), eps 1e-12 (:102)
 *   axvs_dwconv5    : depthwise 5x5 convolution, padding 2, on channels-last rows [N, H, W, C] + folded batch norm [C][2] + act (:78-79);
 *                     w is the Conv2d weight [C, 1, 5, 5] flattened
 *   axvs_add_act    : y = act(a + b), act 0 / 1 (ReLU) / 2 (GELU erf) (:212-213, 219-220); b may be NULL */
int axvs_cm_to_rows(const float* x, float* rows, int N, int C, int M, int act, axvs_stream_t stream);
int axvs_rows_to_cm(const float* rows, int ld, float* out, int N, int C, int M, int normalize, axvs_stream_t stream);
int axvs_dwconv5(const float* x, const float* w, const float* affine, float* y, int N, int H, int W, int C, int act, axvs_stream_t stream);
int axvs_add_act(const float* a, const float* b, float* y, long long n, int act, axvs_stream_t stream);

/* Masked multi-head attention core of the Tube-Link mask decoder: mmcv `MultiheadAttention` (a wrapper of torch.nn.MultiheadAttention;
 * mmcv-full 1.6.1 is not vendored, its wrapper semantics are restated, see axial_vs_b200/tube_link.py) inside `DetrTransformerDecoderLayer`,
 * TL/mmdet/models/utils/transformer.py:408-451, called at TL/models/video/tube_link_vis/mask2former_video_cc_head.py:883-894.
 *   out[b, i, h*32 + :] = softmax_j( q[b, i, h] . k[b, j, h] + (mask[b*heads + h, i, j] ? -inf : 0) ) @ v[b, j, h]
 * q, k, v fp32, ALREADY projected (in_proj) with 32 channels per head; q must carry head_dim^-0.5 * log2(e) (the kernel works in the
 * exp2 domain).  Layout [B, N, heads*32], or [N, B, heads*32] when seq_first != 0 (mmcv batch_first = False).  mask: bytes
 * [B*heads, Nq, L], non-zero = blocked, or NULL.  out32 (fp32) and / or out16 (bf16, the A operand of the output projection), same layout
 * as q.  A row with every key blocked yields NaN like the reference op. */
size_t axvs_masked_mha_workspace_bytes(int B, int heads, int Nq, int L);
int axvs_masked_mha_fwd(const float* q, const float* k, const float* v, const unsigned char* mask, float* out32, void* out16_bf16, int B, int heads,
                        int Nq, int L, int seq_first, void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* Per-frame-softmax spatial attention of TrajectoryAttention in fp32 (split-precision cross-clip path; CC:104-110 = WC/temporal_attention.py:47-60):
 *   x[b, i, f, h*32 + :] = softmax_j( q[b, i, h] . k[b, f*n + j, h] ) @ v[b, f*n + j, h],  j over the n keys of frame f
 * q (pre-scaled by head_dim^-0.5 * log2 e), k, v fp32 [B, N, heads*32] with N = F * n; x fp32 [B, N, F, heads*32].
 * Runs on the masked-attention kernel with the key splits placed on the frames. */
size_t axvs_frame_attn_f32_workspace_bytes(int B, int heads, int N, int F);
int axvs_frame_attn_f32(const float* q, const float* k, const float* v, float* x, int B, int heads, int N, int F, int n, void* workspace,
                        size_t workspace_bytes, axvs_stream_t stream);

/* k-means cross-attention update (DEC:196-208): assign[n, m] = argmax_l mask_logits[n, l, m] (first maximum),
 * out[n, d, l] = sum over the pixels assigned to l of pixel_value[n, d, m]; divided by max(count, 1) when advanced != 0
 * (advanced_kmax, DEC:206-208).  mask_logits fp32 [N, L, M], pixel_value fp32 [N, 256, M], out fp32 [N, 256, L], L <= 128;
 * assign (int32 [N, M]) may be NULL.  Deterministic (no floating-point atomics). */
size_t axvs_kmeans_update_workspace_bytes(int N, int L, int M);
int axvs_kmeans_update(const float* mask_logits, const float* pixel_value, float* out, int* assign, int N, int L, int M, int advanced,
                       void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* ---- within-clip input / output projections ---------------------------------------------
 * input side  (WC/msdeformattn.py:355-358, 413-416): tokens[i, p, :] = GroupNorm(32, 256)(Conv2d(c_in, 256, 1)(x))[i, :, p]
 *   x fp32 NCHW [images, c_in, hw] (c_in % 64 == 0); w_packed = axvs_pack_weight of the conv weight [256, c_in];
 *   out fp32 token-major [images, hw, 256] -- the layout of the temporal layers (the reference's flatten + transpose is folded in);
 *   out_image_stride = floats between consecutive images of `out` (0 = hw * 256): with the stride of the multi-level token tensor
 *   [images, sum(H_l*W_l), 256] the level is written straight into its slice (the reference's torch.cat, WC/msdeformattn.py:106).
 * output side (WC/msdeformattn.py:359-362, 432-434): y[i, :, p] = GroupNorm(32, c_out)(Conv2d(256, c_out, 1))(tokens[i, p, :])
 *   tokens fp32 [images, hw, 256] (tokens_image_stride floats between images, 0 = hw * 256: a level slice of the multi-level token
 *   tensor is read in place, WC/msdeformattn.py:428-434); c_out % 32 == 0 (GroupNorm(32); e.g. 384 for res3 of the ConvNeXt-L configs); w_packed of the conv
 *   weight zero-padded to [ceil256(c_out), 256] rows and bias zero-padded to ceil256(c_out) entries (the GEMM works on 256-column
 *   chunks; padding columns are never stored); out fp32 NCHW [images, c_out, hw].
 * GroupNorm statistics are per image and per group (eps as given, nn.GroupNorm default 1e-5), reductions in a fixed order. */
size_t axvs_proj_workspace_bytes(int images);
int axvs_input_proj_fwd(const float* x_nchw, const void* w_packed, const float* bias, const float* gn_w, const float* gn_b, float* out_tokens,
                        long long out_image_stride, int images, int c_in, int hw, float eps, void* workspace, size_t workspace_bytes,
                        axvs_stream_t stream);
int axvs_output_proj_fwd(const float* tokens, long long tokens_image_stride, const void* w_packed, const float* bias, const float* gn_w,
                         const float* gn_b, float* out_nchw, int images, int c_out, int hw, float eps, axvs_stream_t stream);

/* ---- MSDeformAttn spatial encoder layer ------------------------------------------------------
 * MSDeformAttnTransformerEncoderLayer.forward in eval mode without padding (WC/msdeformattn.py:205-215 with
 * WC/ops/modules/ms_deform_attn.py:92-125; the module builds all-False masks, WC/msdeformattn.py:92):
 *   y = src + output_proj(sample(value_proj(src), softmax(attention_weights(src+pos)), ref + sampling_offsets(src+pos)/(W,H)))
 *   out = LN2(s + FFN(s)), s = LN1(y)
 * src, pos, out fp32 [images, len, 256] (levels concatenated along len, low resolution first); ref_points fp32
 * [images, len, n_levels, 2] normalised (x, y); shapes_hw = HOST array [n_levels][2] of (H, W).  8 heads of 32 channels.
 * pos_images / ref_images = number of images pos / ref_points cover: `images`, or 1 when the (input-independent) tensor is shared
 * by every image and broadcast inside the kernels.
 * w_oa packs [sampling_offsets.weight (8*L*P*2 rows, order head, level, point, xy) ; attention_weights.weight (8*L*P rows) ;
 * zero rows up to 512] with b_oa likewise (512 floats). */
typedef struct axvs_msda_weights {
  const void* w_value; const float* b_value;     /* value_proj  packed [256,256] */
  const void* w_oa;    const float* b_oa;        /* offsets | logits packed [512,256], bias [512] */
  const void* w_out;   const float* b_out;       /* output_proj packed [256,256] */
  const float* ln1_g; const float* ln1_b;
  const void* w_ffn1; const float* b_ffn1; const void* w_ffn2; const float* b_ffn2;   /* as in axvs_layer_weights */
  const void* w_ffn1_u; const void* w_ffn2_u; const void* w_ffn1_n;
  const float* ln2_g; const float* ln2_b;
  const void* w_front_u; const float* b_front;   /* optional (NULL: two generic GEMMs instead): axvs_pack_weight_units (k_major 0) of
                                                    [w_oa rows zero-padded to 384 ; value_proj.weight] = [640, 256] and the matching bias [640]:
                                                    value and offsets | logits projections in one pass over src / pos (4 points per level) */
  const void* w_out_u;                           /* optional (NULL: generic GEMM + LayerNorm kernel): axvs_pack_weight_units (k_major 0) of
                                                    output_proj.weight: output projection + residual + norm1 in one kernel */
  int d_ffn, n_levels, n_points;
} axvs_msda_weights;
size_t axvs_msda_layer_workspace_bytes(int rows, int d_ffn);
int axvs_msda_layer_fwd(const float* src, const float* pos, int pos_images, const float* ref_points, int ref_images, const int* shapes_hw,
                        float* out, const axvs_msda_weights* w, int images, int len, void* workspace, size_t workspace_bytes,
                        axvs_stream_t stream);

/* The sampling half of a multi-scale deformable attention on its own (value projection, offsets / weights projection of query + pos,
 * softmax, bilinear gather; WC/ops/modules/ms_deform_attn.py:98-123 == TL/mmdet/models/plugins/msdeformattn_pixel_decoder.py:573-614):
 * sampled bf16 [images*len, 256].  The Tube-Link attention module inserts its temporal branch between this and the output projection
 * (TL:616-632).  w_oa / b_oa: [sampling_offsets | attention_weights] stacked and zero-padded to 512 rows (as in axvs_msda_weights). */
size_t axvs_msda_sample_workspace_bytes(int rows);
int axvs_msda_sample_fwd(const float* value_in, const float* query_in, const float* pos, int pos_images, const float* ref_points, int ref_images,
                         const int* shapes_hw, const void* w_value, const float* b_value, const void* w_oa, const float* b_oa, int n_levels,
                         int n_points, void* sampled_bf16, int images, int len, void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* ---- clip-to-clip query matching ----------------------------------------------------------------------
 * MaXTronWCDeepLab.match_from_embds (Vk/maxtron_deeplab/maxtron_wc_model.py:391-400; copy in maxtron_cc_model.py) and its chains
 * (maxtron_wc_model.py:342-346, maxtron_cc_model.py:292-295).  The reference builds the cosine cost on the device, moves it to the host and
 * calls scipy.optimize.linear_sum_assignment once per adjacent clip pair; these entry points keep everything on the device.
 *
 * axvs_lsap: exact linear sum assignment of `batch` square cost matrices [n, n] fp32 (row = target, column = current), n <= 256;
 *   col4row[b, i] = column assigned to row i == linear_sum_assignment(cost[b])[1].  The algorithm, its float64 arithmetic and its tie
 *   rules are scipy's (rectangular_lsap.cpp), so the permutation equals scipy's on the same matrix, ties included; a matrix without a
 *   finite assignment yields -1 everywhere.
 * axvs_match_chain: emb [videos, clips, n, e] fp32 (mask embeddings of the clips of each video).  indices[v, 0] = identity;
 *   indices[v, i] = match_from_embds(aligned clip i-1, clip i): emb[v, i][indices[v, i]] is clip i aligned to the first clip.  One
 *   launch for the whole chain of every video. */
int axvs_lsap(const float* cost, int batch, int n, int* col4row, axvs_stream_t stream);
size_t axvs_match_chain_workspace_bytes(int videos, int n, int e);
int axvs_match_chain(const float* emb, int videos, int clips, int n, int e, int* indices, void* workspace, size_t workspace_bytes,
                     axvs_stream_t stream);

/* ---- kMaX pixel-decoder axial attention ------------------------------------------------------------- */

/* One AxialAttention pass (Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:105-157, eval mode) over an image batch:
 * qkv = BN(conv1x1(x)) as a tcgen05 GEMM (the batch norm folded into w_qkv / b_qkv by the caller), then per (sequence, head) in fp32:
 * logits = BN(q.k) + BN(q.rpe_q) + BN(k.rpe_k), softmax, BN(w v) + BN(w rpe_v).
 *   w_qkv   packed (axvs_pack_weight) split-precision image [2*heads*dk + heads*dv, 3*c_in] = [W | W | W - bf16(W)] of the conv weight with
 *           the qkv batch-norm scale folded into its rows (the GEMM multiplies hi / lo bf16 halves: fp32-grade logits); b_qkv the shift
 *   emb_*   fp32 relative-position embedding tables [2*255 - 1, depth] (_query_rpe / _key_rpe / _value_rpe ._embeddings.weight)
 *   sim_s/t fp32 [3*heads]      folded _batch_norm_similarity (y = x*s + t; channels: content, query-rpe, key-rpe heads)
 *   out_s/t fp32 [2*heads*dv]   folded _batch_norm_retrieved_output (content channels, then rpe channels) */
typedef struct axvs_kmax_axial_weights {
  const void* w_qkv; const float* b_qkv;
  const float* emb_q; const float* emb_k; const float* emb_v;
  const float* sim_s; const float* sim_t; const float* out_s; const float* out_t;
  int heads, dk, dv;
} axvs_kmax_axial_weights;

/* x: x_layout 0 = fp32 NCHW [images, c_in, H, W], 1 = fp32 token rows [images*H*W, c_in].  axis 1: one sequence per (image, w) running
 * along H (the `_height_axis` of AxialAttention2D, :179-182; also the 1-D module on [N, C, L] with H = L, W = 1); axis 2: one per
 * (image, h) along W (:184-187).  out: out_layout 0 = NCHW [images, heads*dv, H, W], 1 = token rows [images*H*W, heads*dv] (what the
 * next axis consumes: the reference's permutes at :183 and :188 are folded into these layouts).  Axis length <= 128 and within the shared-memory budget (112 at dk = 64, dv = 128). */
/* 1 (default): axes of 33..48 positions run their attention core on the tensor cores (split-bf16 mma.sync, same accuracy, 1.8x faster at
 * 41 positions); 0: always the fp32 SIMT kernel (the validation baseline; also used for <= 32 and 49..64 positions).  Returns the previous setting. */
int axvs_set_kmax_tensor_cores(int on);
size_t axvs_kmax_axial_workspace_bytes(int images, int c_in, int H, int W, int heads, int dk, int dv);
int axvs_kmax_axial_fwd(const float* x, int x_layout, int images, int c_in, int H, int W, int axis, const axvs_kmax_axial_weights* w,
                        float* out, int out_layout, void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* ---- post-path tail ------------------------------------------------------------------------------ */

/* MaXTronWCDeepLab.panoptic_mask_inference (Vk/maxtron_deeplab/maxtron_wc_model.py:439-553; identical copy in maxtron_cc_model.py:460-574),
 * the mask-wise panoptic merge, without the reference's per-slot host round trips:
 *   mask_cls  fp32 [N, num_classes + 1] class logits (void class last), mask_pred fp32 [N, P] mask logits (P = T*H*W pixels),
 *   cat_ids   int32 [num_classes]: label -> category id (the reference's id_cont_to_ids_dic, :469-473),
 *   is_thing  int32 [num_classes]: label in thing ids (:489),
 *   panoptic  int32 [P]: panoptic_seg_mask (category * label_divisor + instance for things, category for stuff, -1 unassigned),
 *   segments  int32 [1 + 4 N]: number of opened segments, then (slot, label, is_thing, id) per segment in acceptance order
 *             (segments_info / dic_tmp of the reference; the caller gathers the thing slots' embeddings from it).
 * All device pointers.  N <= 255, pixel_thr >= 0.2 (at most four slots can exceed it at one pixel). */
size_t axvs_panoptic_workspace_bytes(int N, long long P);
int axvs_panoptic_inference(const float* mask_cls, const float* mask_pred, int N, int num_classes, long long P,
                            const int* cat_ids, const int* is_thing, int label_divisor,
                            float pixel_thr, float thing_thr, float stuff_thr, float overlap_thr, float w_cls, float w_mask,
                            int* panoptic, int* segments, void* workspace, size_t workspace_bytes, axvs_stream_t stream);

/* PositionEmbeddingSine3D(num_pos_feats=128, normalize=True) + level_embed_3d[lvl], channels-last
 * (WC/pos_embeddings.py:86-130, WC/msdeformattn.py:112-115).  out fp32 [B,T,H,W,256]; level_embed may be NULL. */
int axvs_pos3d(float* out, const float* level_embed, int B, int T, int H, int W, axvs_stream_t stream);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------------------------
 * axvs_profile_enable(1) resets the counters and brackets every kernel launch made through this library with CUDA
 * events on the launching stream (up to 8192 launches); axvs_profile_enable(0) only counts launches.
 * axvs_profile_read synchronises the recorded events and returns, per kernel class (arrays of
 * axvs_profile_num_classes() entries): summed device ms, algorithmic FLOPs, algorithmic bytes, launch count since the
 * last enable, and how many of those launches were event-timed.  Host-synchronising; never call inside graph capture. */
int axvs_profile_enable(int on);
int axvs_profile_num_classes(void);
const char* axvs_profile_class_name(int cls);
int axvs_profile_read(double* ms, double* flops, double* bytes, long long* launches, long long* timed);

#ifdef __cplusplus
}
#endif
#endif /* AXVS_H_ */
