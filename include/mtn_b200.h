/*
 * mtn_b200.h -- C ABI of libmtn_b200.so: the B200 (sm_100a) implementation of the
 * MTN multimodal attention / feed-forward hot path.
 *
 * The reference (henryhungle/MTN @ 5105934) has no FFI layer: its boundary for this
 * path is the nn.Module call protocol of mtn.py.  Each entry point below names the
 * reference lines whose arithmetic it replaces; `mtn_b200/mtn.py` is the host-side
 * mirror that binds them (ctypes) behind the reference's own class names, and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.  The library
 *     borrows them for the duration of the call and allocates nothing: scratch
 *     comes from a caller-provided workspace (so every call is CUDA-graph
 *     capturable).  Launches are asynchronous on `stream` (a cudaStream_t passed as
 *     void*); no call synchronises.
 *   - "f32" = IEEE binary32, "f16" = IEEE binary16.  Tensor-core operands are f16
 *     (11-bit significand == TF32), accumulation / softmax / LayerNorm / residual
 *     stream are f32.  See DESIGN.md "Precision".
 *   - Matrices are row-major; `ld*` are leading dimensions in ELEMENTS.
 *   - Linear weights use the reference layout W[out, in] (nn.Linear, mtn.py:243).
 *   - Return value: 0 on success, a negative MTN_E_* code otherwise; the message is
 *     available from mtn_last_error() (thread-local).  Never aborts.
 *   - Re-entrant; no global mutable state except the lazily resolved driver entry
 *     point for cuTensorMapEncodeTiled.
 */
#ifndef MTN_B200_H_
#define MTN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTN_B200_ABI_VERSION 7   /* v7: mtn_decode_cluster_fwd; v6: mtn_ffn_fused_fwd; v5: batch strides in MtnAttnCoreArgs (KV-cached decoding); v4: `multimem` members in the backward structs */

enum {
  MTN_OK = 0,
  MTN_E_SHAPE = -1,     /* unsupported / inconsistent sizes                      */
  MTN_E_ALIGN = -2,     /* pointer or leading dimension not 16-byte aligned      */
  MTN_E_WORKSPACE = -3, /* workspace pointer NULL or too small                   */
  MTN_E_CUDA = -4,      /* CUDA runtime / driver error (message has the string)  */
  MTN_E_ARG = -5        /* NULL where a pointer is required, bad flag, ...       */
};

int mtn_abi_version(void);
const char *mtn_last_error(void);

/* ---- LayerNorm -----------------------------------------------------------------
 * Replaces LayerNorm.forward (mtn.py:111-114):
 *     y = a_2 * (x - mean) / (std_unbiased + eps) + b_2        (eps added to std!)
 * x: [rows, d] f32.  Writes y as f32 (y_f32, may be NULL) and/or f16 (y_f16, may be
 * NULL; this is the tensor-core operand for the following projection).            */
int mtn_layernorm_fwd(const float *x, const float *a_2, const float *b_2, float eps,
                      int rows, int d, float *y_f32, void *y_f16, void *stream);
/* Grouped variant: a_2 / b_2 are [groups, d]; row r is normalised with parameter set r / rows_per_group
 * (one launch for the same sublayer of several independent streams).                                 */
int mtn_layernorm_grouped_fwd(const float *x, const float *a_2, const float *b_2, float eps, int rows,
                              int d, int rows_per_group, float *y_f32, void *y_f16, void *stream);

/* ---- embeddings ----------------------------------------------
 * Replaces Embeddings.forward (mtn.py:288-289) + PositionalEncoding.forward (mtn.py:307-309)
 * and, when a_2/b_2 are given, the Encoder's per-stream LayerNorm (mtn.py:91, :96):
 *     y[r, :] = LN?( lut[ids[r], :] * scale + pe[r % L, :] )
 * ids: [rows] int64 (rows = B*L, position = r % L), lut: [vocab, d] f32, pe: [>=L, d] f32.   */
int mtn_embed_fwd(const int64_t *ids, const float *lut, const float *pe, int rows, int L, int d,
                  int vocab, float scale, const float *a_2, const float *b_2, float eps,
                  float *y_f32, void *y_f16, void *stream);

/* Training variant: PositionalEncoding's dropout (mtn.py:309) applied to lut*scale + pe BEFORE the stream
 * LayerNorm; element index = r * d + c (see MtnLinearArgs for the RNG contract).                              */
int mtn_embed_dropout_fwd(const int64_t *ids, const float *lut, const float *pe, int rows, int L, int d,
                          int vocab, float scale, const float *a_2, const float *b_2, float eps,
                          float *y_f32, void *y_f16, const void *drop_seed, uint32_t drop_site,
                          uint32_t drop_thresh, void *stream);

/* ---- video-feature preparation (boundary: Batch, data_utils.py:28-30) ----------------
 * ft: [frames, F] f32 raw features.  mask[frame] = any(ft[frame, :] != 1.0) (all-ones frames are
 * padding); padded frames are zeroed; output as f16 (video-encoder operand) and/or f32.        */
int mtn_feature_prep_fwd(const float *ft, int frames, int F, uint8_t *mask, void *out_f16,
                         float *out_f32, void *stream);
/* Features stored as f16 on the host (half the upload; the rounding is the one mtn_feature_prep_fwd applies on the
 * device, so everything downstream is bit-identical): same mask / zeroing, f16 in, f16 out.  F % 8 == 0.        */
int mtn_feature_prep_f16_fwd(const void *ft_f16, int frames, int F, uint8_t *mask, void *out_f16, void *stream);

/* ---- generator tail ---------------------------------------------
 * Replaces F.log_softmax(proj(x), -1) (mtn.py:68-69) after mtn_linear_fwd produced the logits,
 * and the arg-max of greedy decoding (data_utils.py:183).  x: [rows, ldx] f32, first V columns
 * are the vocabulary; y (may alias x, may be NULL) receives log-probabilities; argmax (may be
 * NULL) the first maximal index per row.                                                      */
int mtn_log_softmax_fwd(const float *x, int ldx, int rows, int V, float *y, int ldy,
                        int64_t *argmax, void *stream);

/* ---- label-smoothed loss (forward) -----------------------------------
 * Replaces LabelSmoothing.forward (label_smoothing.py:20-32: smoothed one-hot target, padding column
 * zeroed, padding rows zeroed with the reference's index-sum quirk) + nn.KLDivLoss(sum) applied to
 * log_softmax(logits), computed per row from the logits (or from log-probabilities -- same formula):
 *     loss[0] (+)= scale * sum_rows KL(true_dist_r || softmax(logits_r))
 * `accumulate` != 0 adds to loss[0] (main loss + lambda * auto-encoder losses, data_utils.py:133-151);
 * `scale` carries 1/norm.  Deterministic (fixed-order reductions).  target: [rows] int64.            */
size_t mtn_label_smoothing_workspace_bytes(int rows);
int mtn_label_smoothing_loss_fwd(const float *logits, int ld, int rows, int V, const int64_t *target,
                                 int64_t padding_idx, float smoothing, float scale, int accumulate,
                                 float *loss, void *workspace, size_t workspace_bytes, void *stream);

/* ---- casts / packing -----------------------------------------------------------
 * dst[r, c] = (f16) src[r, c]  (round-to-nearest-even, saturating to +-65504).
 * Used to pack nn.Linear weights once per parameter version and to convert module
 * inputs that arrive as f32 (e.g. raw I3D / VGGish features, mtn.py:35).           */
int mtn_cast_f32_to_f16(const float *src, int ld_src, void *dst, int ld_dst, int rows, int cols,
                        void *stream);

/* ---- mask packing --------------------------------------------------------------
 * The reference passes bool masks (B,1,Lk) [key padding, data_utils.py:34-38] or
 * (B,T,T) [causal & padding, data_utils.py:49-54] into attention() (mtn.py:226-227).
 * mask_u8: [B, rows_q, Lk] bytes (rows_q == 1 for key-padding masks), non-zero = keep.
 * bits: [B, rows_q, words] uint32, words = mtn_mask_words(Lk); bit k%32 of word k/32
 * is key k.  Packed once per forward and shared by every layer that uses the mask. */
int mtn_mask_words(int Lk);
int mtn_mask_pack(const uint8_t *mask_u8, int B, int rows_q, int Lk, uint32_t *bits, void *stream);

/* ---- fused linear --------------------------------------------------------------
 * Replaces nn.Linear call sites on the path (mtn.py:256-258, 267, 280, 35):
 *     C[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] ) + addend[m % add_period, n]
 * A: [M, K] f16 (lda), W: [N, K] f16 (ldw), bias: [N] f32 or NULL.
 * act: MTN_ACT_NONE | MTN_ACT_RELU (mtn.py:280, :378).
 * addend: f32 [*, ld_add] or NULL -- the residual x of SublayerConnection
 *         (mtn.py:127; add_period = 0 means row m) or the positional-encoding
 *         table (mtn.py:308; add_period = sequence length).  May alias out_f32.
 * Outputs: out_f32 [M, ld32] and/or out_f16 [M, ld16]; either may be NULL.
 * Constraints: K % 8 == 0, N % 8 == 0 (16-byte aligned rows); tails are zero-filled
 * by TMA (K) / predicated in the epilogue (M, N).
 * tcgen05 (UMMA 128xBNx16, f16 in / f32 accumulate in TMEM), TMA-staged operands. */
enum { MTN_ACT_NONE = 0, MTN_ACT_RELU = 1 };
typedef struct MtnLinearArgs {
  const void *A; int lda;
  const void *W; int ldw;
  const float *bias;
  int M, N, K;
  int act;
  const float *addend; int ld_add; int add_period;
  float *out_f32; int ld32;
  void *out_f16; int ld16;
  /* Strided batch (ABI v2): `batch` > 1 runs that many independent problems of identical shape in ONE
   * launch; problem b uses A + b*stride_A, W + b*stride_W, bias + b*stride_bias, addend + b*stride_add,
   * out_f32 + b*stride_out_f32, out_f16 + b*stride_out_f16 (strides in ELEMENTS).  Used to run the two
   * video modalities' Query-Aware Auto-Encoder projections (same shapes, different weights) together. */
  int batch;
  long long stride_A, stride_W, stride_bias, stride_add, stride_out_f32, stride_out_f16;
  /* ABI v3: out_f16 receives the value BEFORE the addend (video encoder in training: relu(.) without the
   * positional term is what the ReLU backward needs).                                                  */
  int out16_pre_add;
  /* ABI v3, training-mode dropout of the result (nn.Dropout of SublayerConnection mtn.py:127, of the FFN hidden
   * layer mtn.py:280, of PositionalEncoding mtn.py:309): element (m, n) is kept iff the 16 random bits that
   * Philox-4x32-10(counter = ((m*N + n) >> 3, drop_site), key = *drop_seed) assigns to it are >= drop_thresh
   * (= round(p * 65536)); kept elements are multiplied by 1/(1 - drop_thresh/65536).  Applied after the
   * activation and BEFORE the addend (drop_after_add = 0) or after it (1).  drop_seed: device pointer to a
   * 64-bit seed, NULL = no dropout.  The backward kernels regenerate the same decisions from (seed, site).  */
  const void *drop_seed; uint32_t drop_site, drop_thresh; int drop_after_add;
} MtnLinearArgs;
int mtn_linear_fwd(const MtnLinearArgs *args, void *stream);

/* ---- LayerNorm fused into the projection that consumes it (inference) ----------------
 * Replaces SublayerConnection's norm (mtn.py:126) TOGETHER with the first nn.Linear of the sublayer
 * (mtn.py:256 Q projection or packed Q|K|V, mtn.py:280 w_1) in one launch:
 *     out_f16[M, N] = act( LN(x)[M, d] W[N, d]^T + bias )
 * x: [M, d] f32 contiguous rows; LN as mtn_layernorm_fwd (unbiased std, eps added to std), rounded to
 * f16 exactly like its y_f16 output, so the result equals mtn_layernorm_fwd + mtn_linear_fwd
 * (measured: bit-identical at d = 256 / 512; at d = 128 single elements differ by one f16 rounding).  W: f16 [N, d] row-major (leading dimension ldw).  d in {128, 256, 512}
 * (mtn_ln_linear_supported); other sizes take the two-call form.                                  */
int mtn_ln_linear_supported(int d);
/* tools only: later launches make CTA (0,0) write 8 clock64() phase stamps to dev_buf8 (NULL: off). */
int mtn_ln_linear_debug_timestamps(void *dev_buf8);
int mtn_ln_linear_fwd(const float *x, const float *a_2, const float *b_2, float eps, int M, int d,
                      const void *W, int ldw, const float *bias, int N, int act, void *out_f16,
                      int ld16, void *stream);

/* ---- attention core ------------------------------------------------------------
 * Replaces attention() (mtn.py:221-231) plus the head split / concat views around
 * it (mtn.py:257, 265-266) for all B*h heads in one launch:
 *     S = Q_h K_h^T / sqrt(d_k);  S[mask == 0] = -1e9 (FINITE: a fully masked row
 *     is a uniform average);  P = softmax(S);  O_h = P V_h
 * q: f16, row (b*Lq + i), columns [h*d_k, (h+1)*d_k) of a matrix with leading
 *    dimension ldq (so a packed [Q|K|V] projection buffer can be addressed
 *    without copies);  k, v likewise over rows (b*Lk + j).
 * mask_bits: packed by mtn_mask_pack, or NULL for "no mask" (mtn.py:226);
 *    mask_rows_q == 1 broadcasts one key mask over all queries of a batch element.
 * out: f16 [B*Lq, ldo], head h written to columns [h*d_k, (h+1)*d_k)  (== the
 *    reference's transpose(1,2).contiguous().view(B, -1, h*d_k)).
 * d_k in {32, 64}.  S and O accumulate in TMEM; softmax is f32, one thread per row. */
typedef struct MtnAttnCoreArgs {
  const void *q; int ldq;
  const void *k; int ldk;
  const void *v; int ldv;
  const uint32_t *mask_bits; int mask_rows_q;
  int B, h, Lq, Lk, d_k;
  void *out; int ldo;
  /* ABI v3 (training): optional [B, h, Lq, 2] f32 -- per query row {row maximum of the masked scaled scores
   * in the log2 domain, 1 / softmax denominator}; saved for mtn_attn_core_bwd.  NULL: not written.       */
  float *stats;
  /* ABI v3, training-mode dropout of the probabilities (mtn.py:229-230): P[b, head, q, k] is kept iff the 16 random
   * bits Philox-4x32-10(counter = ((((b*h + head)*Lq + q) * Lk32 + k) >> 3, drop_site), key = *drop_seed), Lk32 =
   * Lk rounded up to 32, assigns to it are >= drop_thresh; kept probabilities are multiplied by
   * 1/(1 - drop_thresh/65536).  The softmax normalisation is unaffected.  drop_seed NULL = no dropout.        */
  const void *drop_seed; uint32_t drop_site, drop_thresh;
  /* ABI v5 (KV-cached decoding): element strides between consecutive batch elements of q / k / v / out.
   * 0 = dense (Lq*ldq, Lk*ldk, Lk*ldv, Lq*ldo).  A self-attention cache [B, T_max, 3d] is then addressed in place:
   * q = row t of every dialogue (Lq = 1, q_batch_stride = T_max*3d), k / v = its first t+1 rows.  Multiples of 8. */
  long long q_batch_stride, k_batch_stride, v_batch_stride, o_batch_stride;
} MtnAttnCoreArgs;
int mtn_attn_core_fwd(const MtnAttnCoreArgs *args, void *stream);

/* ---- few-row variants for KV-cached decoding (csrc/decode_rows.cu) -----------------
 * At one target position per dialogue every kernel of a decoding step works on M = B rows; the tcgen05 kernels pay
 * ~6 us of fixed cost per launch there.  Same contracts, same arithmetic (f16 operands, f32 accumulate, the same
 * softmax formulas), no tensor-memory / TMA setup:
 *   mtn_rows_linear_fwd  = mtn_linear_fwd for M <= 128 (N % 8 == 0, K % 32 == 0); `addend`, if given, must be the f32
 *                          output itself (in-place residual).  mma.sync, one CTA per 8 output columns.
 *   mtn_decode_attn_fwd  = mtn_attn_core_fwd for Lq <= 8 query rows per batch element, d_k = 64 (batch strides
 *                          honoured: the self-attention cache is addressed in place).  One warp per (batch, head).  */
int mtn_rows_linear_supported(int M, int N, int K);
int mtn_rows_linear_fwd(const MtnLinearArgs *args, void *stream);
/*   mtn_rows_ln_linear_fwd = mtn_layernorm_fwd + mtn_rows_linear_fwd in ONE launch (bit-identical): out_f16 = act(LN(x) W^T + b),
 *                          x f32 [M, d] with row pitch ldx, d in {128, 256, 512, 1024}, M <= 128.                          */
int mtn_rows_ln_linear_supported(int M, int N, int d);
int mtn_rows_ln_linear_fwd(const float *x, int ldx, const float *a_2, const float *b_2, float eps, int M, int d,
                           const void *W, int ldw, const float *bias, int N, int act, void *out_f16, int ld16,
                           void *stream);
int mtn_decode_attn_supported(int Lq, int d_k);
int mtn_decode_attn_fwd(const MtnAttnCoreArgs *args, void *stream);

/* ---- one KV-cached decoding step as ONE persistent kernel (ABI v6; csrc/decode_rows.cu) ----------------------------
 * Between mtn_prog_begin and mtn_prog_end, mtn_rows_linear_fwd, mtn_rows_ln_linear_fwd, mtn_decode_attn_fwd (Lq = 1) and
 * mtn_layernorm_fwd (d = 512) RECORD their launch as a stage instead of launching (every other entry point fails with
 * MTN_E_CUDA: it would run out of order).  mtn_prog_end copies the stage list (mtn_prog_stage_bytes() bytes per stage)
 * into host memory supplied by the caller (NULL: discard); after the caller has copied it to the device,
 * mtn_prog_launch runs the stages in order inside one cooperative kernel, a grid barrier between consecutive stages
 * (same device functions as the stand-alone kernels: bit-identical results).  `counter`: 4 bytes of device memory.
 * Graph-capturable; recording is thread-local and touches no CUDA state.                                              */
int mtn_prog_begin(void);
int mtn_prog_recording(void);
int mtn_prog_stage_bytes(void);
int mtn_prog_end(void *host_dst, size_t capacity, int *n_stages);
int mtn_prog_launch(const void *dev_prog, int n_stages, void *counter, void *stream);

/* ---- one KV-cached decoding step, ONE kernel, one thread-block cluster per dialogue group (ABI v7;
 * csrc/decode_cluster.cu) ---------------------------------------------------------------------------------------
 * The whole target path of one new position per dialogue -- reference call form data_utils.py:202-210 restricted to
 * the last position, i.e. for every DecoderLayer (mtn.py:181-218) its self-attention over the layer's cache, its
 * cross-attention sublayers over static (already projected) memories and its feed-forward sublayer, each a pre-norm
 * residual sublayer (mtn.py:125-127), then Decoder.norm (mtn.py:164) -- described as a flat list of sublayers
 * ("sites") in execution order.  A cluster of 8 CTAs owns ceil(B / #clusters) dialogues for the whole step; CTA r owns
 * head r.  d = 512, h = 8 (d_k = 64), d_ff = 2048, B <= 8 rows x the clusters the device keeps co-resident (13 on a B200:
 * 104 rows; mtn_decode_cluster_supported tells), n_sites <= mtn_decode_cluster_max_sites().
 *   kind 0  self-attention:  w_in = [Wq;Wk;Wv] f16 [3d, d], b_in [3d];  q_cache / k / v = column 0 / d / 2d of row 0,
 *           dialogue 0 of the layer's cache (f16 [B, T_max, 3d]: ld_kv = 3d, kv_batch_stride = T_max * 3d).  The new
 *           position's [Q|K|V] is written to cache row t; keys / values are cache rows 0..t (no mask: causal past).
 *   kind 1  cross-attention: w_in = Wq f16 [d, d], b_in [d];  k / v = head 0 of key 0 of dialogue 0 (f16, row pitch
 *           ld_kv, dialogue pitch kv_batch_stride, Lk keys);  mask_bits: bit-packed key mask (mtn_mask_pack) with ONE
 *           query row per dialogue, [D, mask_words] words (D = B / rows_per_dialogue), or NULL.  Masked scores take the reference's finite -1e9.
 *   kind 2  feed-forward:    w_in = w_1 f16 [d_ff, d], b_in [d_ff];  w_out = w_2 f16 [d, d_ff], b_out [d].
 * Attention sites: w_out = Wo f16 [d, d], b_out [d].  All weights contiguous (row pitch = their K).
 * `sites` is a HOST array (copied into the kernel's parameter space: graph-capturable, nothing to upload);
 * x_in: [B, d] f32 embedding (+ positional encoding) of the new position; out: [B, d] f32; taps: optional
 * [n_sites, B, d] f32 debug output (the residual rows after every sublayer) or NULL.                              */
typedef struct MtnDecodeSite {
  int kind; float ln_eps;
  const float *ln_a; const float *ln_b;
  const void *w_in;  const float *b_in;
  const void *w_out; const float *b_out;
  const void *k; const void *v; void *q_cache;
  int ld_kv, Lk;
  long long kv_batch_stride;
  const uint32_t *mask_bits; int mask_words;
} MtnDecodeSite;
typedef struct MtnDecodeClusterArgs {
  const MtnDecodeSite *sites; int n_sites;
  int B, d, h, d_ff;
  int t;
  const float *x_in; float *out;
  const float *norm_a; const float *norm_b; float norm_eps;
  float *taps;
  long long *stamps;   /* optional debug output [n_sites, 8]: clock64 at the phase boundaries of every sublayer (CTA 0) or NULL */
  /* optional last stage in the same kernel (greedy decoding, data_utils.py:180-184): tokens[r * tokens_stride] =
   * argmax_v (out[r] . gen_w[v] + gen_b[v]), v < gen_V (first maximal index) -- Generator.forward (mtn.py:68-69) followed by
   * the arg-max; gen_w: f16 [gen_V8, d] (rows >= gen_V zero, gen_V8 a multiple of 8), gen_b: [gen_V8].  gen_w NULL: off. */
  const void *gen_w; const float *gen_b; int gen_V, gen_V8;
  int64_t *tokens; long long tokens_stride;
  /* beam search: R = rows_per_dialogue consecutive target rows are the hypotheses of ONE dialogue (B = D * R): the
   * cross-attention sites' K / V and mask of row r are those of dialogue r / R (the memory stage is stored once per
   * dialogue); the self-attention caches stay per row.  0 or 1: one row per dialogue.                              */
  int rows_per_dialogue;
} MtnDecodeClusterArgs;
int mtn_decode_cluster_supported(int B, int d, int h, int d_ff);
int mtn_decode_cluster_max_sites(void);
int mtn_decode_cluster_fwd(const MtnDecodeClusterArgs *args, void *stream);

/* ---- one attention site --------------------------------------------------------
 * Replaces  SublayerConnection.forward(x, lambda x: attn(x, mem, mem, mask))
 * (mtn.py:125-127 around mtn.py:248-267):
 *     x_out = x + Wo . concat_h attention( LN(x) Wq^T + bq , mem Wk^T + bk , mem Wv^T + bv ) + bo
 * x: [B*Lq, d] f32.  For self-attention (mtn.py:183, :209) pass mem_f16 == NULL and
 * kv == NULL: keys/values are projected from LN(x) with w_qkv = [Wq;Wk;Wv] ([3d, d]).
 * For cross-attention either pass mem_f16 ([B*Lk, d] f16, the Encoder-normed memory)
 * with w_kv = [Wk;Wv] ([2d, d]) or, when the K/V projection was hoisted out of the
 * layer loop with mtn_linear_fwd, pass kv (f16, [B*Lk, ld_kv]; K at column kv_k_col,
 * V at column kv_v_col).
 * x_out may alias x.  Workspace: mtn_attn_site_workspace_bytes().                   */
typedef struct MtnAttnSiteArgs {
  int B, Lq, Lk, d, h;
  const float *x; float *x_out;
  const float *ln_a, *ln_b; float ln_eps;
  const void *w_q;   const float *b_q;      /* [d, d] f16 or, for self-attention, [3d, d] */
  const void *w_kv;  const float *b_kv;     /* [2d, d] f16 (cross-attention, not hoisted)  */
  const void *w_o;   const float *b_o;      /* [d, d] f16                                  */
  const void *mem_f16;                      /* [B*Lk, d] f16 or NULL                       */
  const void *kv; int ld_kv, kv_k_col, kv_v_col;
  const uint32_t *mask_bits; int mask_rows_q;
  void *workspace; size_t workspace_bytes;
} MtnAttnSiteArgs;
size_t mtn_attn_site_workspace_bytes(int B, int Lq, int Lk, int d);
int mtn_attn_site_fwd(const MtnAttnSiteArgs *args, void *stream);

/* ---- one attention site, ONE kernel (hoisted K/V) ----------------------------------
 * The same sublayer as mtn_attn_site_fwd for a site whose K / V projections already exist (static memories: `kv`),
 * taking xn = LayerNorm(x) in f16 (mtn_layernorm_fwd) and updating the f32 residual stream IN PLACE:
 *     x += Wo . concat_h softmax(mask((xn Wq_h^T + bq_h) K_h^T / sqrt(d_k))) V_h + bo
 * One launch: a cluster of two CTAs per (batch element, 128-query tile) keeps the Q projection, the attention and the
 * output projection on chip (csrc/site_fused.cu); the two CTAs exchange their heads' outputs through distributed
 * shared memory.  d_k = 64 with d in {256, 512} (mtn_attn_site_fused_supported); other shapes: mtn_attn_site_fwd.
 * ld_wq / ld_wo: row pitch of the weights in elements (0 = d), so w_q may be the first d rows of a [3d, d] pack. */
typedef struct MtnAttnSiteFusedArgs {
  int B, Lq, Lk, d, h;
  const void *xn_f16; int ld_xn;            /* [B*Lq, d] f16 */
  float *x; int ld_x;                       /* [B*Lq, d] f32, updated in place */
  const void *w_q; int ld_wq; const float *b_q;
  const void *w_o; int ld_wo; const float *b_o;
  const void *kv; int ld_kv, kv_k_col, kv_v_col;   /* f16 [B*Lk, ld_kv]; K at column kv_k_col, V at kv_v_col */
  const uint32_t *mask_bits; int mask_rows_q;
} MtnAttnSiteFusedArgs;
int mtn_attn_site_fused_supported(int d, int h);
int mtn_attn_site_fused_fwd(const MtnAttnSiteFusedArgs *args, void *stream);

/* ---- feed-forward sublayer -----------------------------------------------------
 * Replaces  SublayerConnection.forward(x, PositionwiseFeedForward)  (mtn.py:125-127
 * around mtn.py:279-280):   x_out = x + W2 relu(W1 LN(x) + b1) + b2
 * x: [rows, d] f32; w_1: [d_ff, d] f16; w_2: [d, d_ff] f16.  x_out may alias x.    */
typedef struct MtnFfnArgs {
  int rows, d, d_ff;
  const float *x; float *x_out;
  const float *ln_a, *ln_b; float ln_eps;
  const void *w_1; const float *b_1;
  const void *w_2; const float *b_2;
  void *workspace; size_t workspace_bytes;
} MtnFfnArgs;
size_t mtn_ffn_workspace_bytes(int rows, int d, int d_ff);
int mtn_ffn_fwd(const MtnFfnArgs *args, void *stream);

/* The same sublayer behind the LayerNorm in ONE kernel, the [rows, d_ff] hidden activation never written to HBM
 * (csrc/ffn_fused.cu; ABI v6):   x += W2 relu(W1 xn + b1) + b2,   xn = LayerNorm(x) as f16 (mtn_layernorm_fwd).
 * xn_f16: [rows, ld_xn] f16; x: [rows, ld_x] f32 updated in place; w_1: [d_ff, d] f16; w_2: [d, d_ff] f16 (mtn.py:276-277).
 * Supported: d = 512, d_ff a multiple of 128.  Results are bit-identical to the two mtn_linear_fwd launches of mtn_ffn_fwd. */
int mtn_ffn_fused_supported(int rows, int d, int d_ff);
int mtn_ffn_fused_fwd(const void *xn_f16, int ld_xn, float *x, int ld_x, int rows, int d, int d_ff,
                      const void *w_1, const float *b_1, const void *w_2, const float *b_2, void *stream);


/* =====================================================================================================
 * BACKWARD (training) entry points -- ABI v3.  The reference trains through torch.autograd (train.py:33-39,
 * data_utils.py:152-155 loss.backward()); these are the gradients of the entry points above, bound by
 * mtn_b200/train_engine.py behind one autograd.Function per reference module.
 *
 * Gradient scaling: tensor-core operands of the backward GEMMs are f16.  A backward pass multiplies the
 * incoming gradient by a power of two S picked on the device (mtn_grad_absmax + mtn_grad_scale -> {S, 1/S}),
 * keeps intermediates scaled, and un-scales every result in the kernel that produces it.  `scale` / `alpha`
 * below are DEVICE pointers to those scalars (NULL = 1), so a whole step stays CUDA-graph capturable.
 * ===================================================================================================== */

/* ---- general tensor-core GEMM ----------------------------------------------------------------------
 *   C[m, n] = act( alpha * sum_k A(m, k) B(n, k) + bias[n] ) [* (relu_mask[m, n] > 0)] + addend[.., n]
 * A(m, k) = A[m*lda + k] (a_mn = 0, "K-major") or A[k*lda + m] (a_mn = 1, "MN-major"); B likewise.
 * (0,0) is the forward linear, (0,1) its data gradient, (1,1) its weight gradient.  accumulate != 0:
 * out_f32 += alpha * A B^T with split-K over the CTAs (atomic f32 adds; bias / act / addend not allowed). */
typedef struct MtnGemmArgs {
  const void *A; int lda; int a_mn;
  const void *B; int ldb; int b_mn;
  int M, N, K;
  const float *alpha;
  const float *bias; int act;
  const void *relu_mask; int ld_mask;
  const float *addend; int ld_add; int add_period;
  int accumulate;
  float *out_f32; int ld32;
  void *out_f16; int ld16; int out16_pre_add;
  int batch;
  long long stride_A, stride_B, stride_bias, stride_add, stride_out_f32, stride_out_f16;
  float mask_scale;                   /* multiplies what passes relu_mask (0 = 1) */
  const void *drop_seed; uint32_t drop_site, drop_thresh; int drop_after_add;   /* see MtnLinearArgs */
  float *colsum_a;                    /* weight-gradient form only: colsum_a[m] += alpha * sum_k A(m, k) */
  int multimem;                       /* accumulate form: out_f32 / colsum_a are NVLS multicast addresses (see below) */
} MtnGemmArgs;
int mtn_gemm_f16(const MtnGemmArgs *args, void *stream);
int mtn_check_gemm_f16(const MtnGemmArgs *args, void *stream);   /* tests only */

/* ---- linear backward (autograd of nn.Linear, mtn.py:256-258, 267, 280, 35, 68) -------------------------
 * dgrad:  dX[M, K] = alpha * dY[M, N] W[N, K]  (* (relu_mask > 0): gradient through the ReLU that produced
 *         X, mtn.py:280) (+ addend).  W is read in its forward [out, in] layout.
 * wgrad:  dW[N, K] += alpha * dY[M, N]^T X[M, K]  (f32, atomically accumulated: the caller zeroes or keeps
 *         the running .grad).  dY and X are read in their forward row-major layouts.                      */
typedef struct MtnLinearDgradArgs {
  const void *dY; int lddy;            /* f16 [M, N] */
  const void *W; int ldw;              /* f16 [N, K] */
  int M, N, K;
  const float *alpha;
  const void *relu_mask; int ld_mask;  /* f16 [M, K] or NULL */
  float mask_scale;                    /* multiplies what passes the mask: 1/(1-p) of a dropout after the ReLU (0 = 1) */
  const float *addend; int ld_add;     /* f32 [M, K] or NULL */
  float *dX_f32; int ld32;
  void *dX_f16; int ld16;
  int batch;
  long long stride_dY, stride_W, stride_add, stride_dX_f32, stride_dX_f16;
} MtnLinearDgradArgs;
int mtn_linear_dgrad(const MtnLinearDgradArgs *args, void *stream);

typedef struct MtnLinearWgradArgs {
  const void *dY; int lddy;            /* f16 [M, N] */
  const void *X; int ldx;              /* f16 [M, K] */
  int M, N, K;
  const float *alpha;
  float *dW; int lddw;                 /* f32 [N, K] += */
  int batch;
  long long stride_dY, stride_X, stride_dW;
  float *dbias;                        /* optional f32 [N] += alpha * column sums of dY (the bias gradient), computed
                                        * by one extra MMA against a tile of ones inside the same kernel          */
  /* Data-parallel training on an NVSwitch box: pass dW / dbias as addresses inside the NVLS MULTICAST mapping of
   * the (symmetric) gradient buffer and set multimem = 1 -- the epilogue then issues multimem.red.add instead of
   * red.add, the switch adds the tile into EVERY rank's gradient buffer, and the gradient all-reduce disappears as
   * a separate step (one cross-GPU barrier before the optimizer replaces it).                                    */
  int multimem;
} MtnLinearWgradArgs;
int mtn_linear_wgrad(const MtnLinearWgradArgs *args, void *stream);

/* ---- operand cast + bias gradient ------------------------------------------------------------------
 * dst_f16[r, c] = f16( src[r, c] * scale  [0 where relu_mask[r, c] <= 0] )      (dst_f16 may be NULL)
 * colsum[c]    += alpha * sum_r (the same values in f32)                         (colsum may be NULL)
 * src is f32 (src_is_f16 = 0) or f16.  cols % 8 == 0.  The bias gradient of every nn.Linear on the path. */
int mtn_cast_colsum(const void *src, int src_is_f16, int ld_src, void *dst_f16, int ld_dst,
                    const void *relu_mask, int ld_mask, int rows, int cols, const float *scale,
                    const float *alpha, float *colsum, const void *drop_seed, uint32_t drop_site,
                    uint32_t drop_thresh, int multimem /* colsum is an NVLS multicast address */, void *stream);
/* (drop_*: the values are additionally thinned by the dropout of the linear layer's output they are the
 * gradient of -- see MtnLinearArgs; element index = r * cols + c.)                                        */

/* ---- gradient scale ---------------------------------------------------------------------------------
 * mtn_grad_absmax folds max|x| into *slot (a zero-initialised uint32 in device memory; may be called for
 * several tensors); mtn_grad_scale turns it into scale2 = {S, 1/S}, S = 2^k with absmax*S in [128, 256)
 * (S = 1 for an all-zero or non-finite gradient) and resets the slot.                                   */
int mtn_grad_absmax(const float *x, size_t n, uint32_t *slot, void *stream);
int mtn_grad_scale(uint32_t *slot, float *scale2, void *stream);
/* ---- fused Adam (optimizer of train.py:190-191 = torch.optim.Adam wrapped by NoamOpt) ------
 * state: 8 f32 in device memory {lr, beta1, beta2, eps, 1-beta1^t, 1-beta2^t, t, -}.
 * mtn_adam_advance: t += 1, refresh the bias corrections and, if noam_factor > 0, the learning rate of
 *   data_utils.py:112-117: lr = factor * model_size^-0.5 * min(t^-0.5, t * warmup^-1.5).
 * mtn_adam_step over a flat arena of n f32 parameters: torch.optim.Adam's update (no weight decay / amsgrad),
 *   optionally p_f16 = f16(p) (the tensor-core operand arena) and g = 0, in one pass.                        */
int mtn_adam_advance(float *state, float noam_factor, float model_size, float warmup, void *stream);
int mtn_adam_step(float *p, float *g, float *m, float *v, void *p_f16, size_t n, const float *state,
                  int zero_grad, void *stream);
/* *seed += 1 on the stream (the captured training step bumps the dropout seed once per replay). */
int mtn_seed_bump(uint64_t *seed, void *stream);
/* Stream-ordered zero fill (cudaMemsetAsync) of a gradient accumulation buffer. */
int mtn_zero(void *p, size_t bytes, void *stream);
/* y = (accumulate ? y : 0) + x * alpha[0]: un-scales an input gradient leaving the backward pass.  n % 4 == 0. */
int mtn_scale_f32(const float *x, const float *alpha, float *y, size_t n, int accumulate, void *stream);

/* ---- LayerNorm backward (autograd of mtn.py:111-114) --------------------------------------------------
 * dx = dres + dLN/dx(dy * dy_scale);  da_2 += param_alpha * sum_rows dy' (x - mean)/(std + eps);
 * db_2 += param_alpha * sum_rows dy'.  dres may be NULL or alias dx (residual-stream gradient in place). */
typedef struct MtnLayerNormBwdArgs {
  const float *x; const float *a_2; float eps;
  int rows, d;
  const float *dy; const float *dy_scale;
  const float *dres; float *dx;
  float *da_2; float *db_2; const float *param_alpha;
  /* optional fused outputs for the NEXT backward step (the sublayer whose output x is): an f16 copy of dx (its
   * GEMM operand) and dx_colsum[c] += param_alpha * sum_rows dx[:, c] (its output-bias gradient)            */
  void *dx_f16; float *dx_colsum;
  const void *drop_seed; uint32_t drop_site, drop_thresh;   /* dropout applied to dx_f16 / dx_colsum only */
  int multimem;                         /* da_2 / db_2 / dx_colsum are NVLS multicast addresses (see MtnLinearWgradArgs) */
} MtnLayerNormBwdArgs;
int mtn_layernorm_bwd(const MtnLayerNormBwdArgs *args, void *stream);

/* ---- embedding backward (autograd of mtn_embed_fwd) ---------------------------------------------------
 * dy: gradient of the (optionally stream-normalised) embedding output; the pre-norm value is recomputed.
 * dlut[ids[r], :] += param_alpha * scale * dpre[r, :] (atomic); da_2 / db_2 as in mtn_layernorm_bwd.      */
typedef struct MtnEmbedBwdArgs {
  const int64_t *ids; const float *lut; const float *pe;
  int rows, L, d, vocab; float scale;
  const float *a_2; float eps;
  const float *dy;
  float *dlut; float *da_2; float *db_2; const float *param_alpha;
  const void *drop_seed; uint32_t drop_site, drop_thresh;   /* the forward's dropout (mtn_embed_dropout_fwd) */
} MtnEmbedBwdArgs;
int mtn_embed_bwd(const MtnEmbedBwdArgs *args, void *stream);

/* ---- attention core backward (autograd of attention(), mtn.py:221-231) -----------------------------------
 * delta[b, h, q] = sum_c dO[row, h*d_k + c] * O[row, h*d_k + c]  (mtn_attn_delta), then
 *   dV = P^T dO,  dS = P (dO V^T - delta) / sqrt(d_k) [0 where masked],  dQ = dS K,  dK = dS^T Q
 * with P recomputed from q, k, mask and the forward's `stats`.  dq is f32 and ACCUMULATED atomically over the
 * key tiles (zero it first) -- or dq_f16 for Lk <= 128; dk / dv are f16, head h in columns [h*d_k, (h+1)*d_k). */
int mtn_attn_delta(const void *dO, int lddo, const void *O, int ldo, int B, int Lq, int h, int d_k,
                   float *delta, void *stream);
typedef struct MtnAttnCoreBwdArgs {
  const void *q; int ldq;
  const void *k; int ldk;
  const void *v; int ldv;
  const void *dO; int lddo;
  const float *stats; const float *delta;
  const uint32_t *mask_bits; int mask_rows_q;
  int B, h, Lq, Lk, d_k;
  float *dq; int lddq;
  void *dk; int lddk;
  void *dv; int lddv;
  const void *drop_seed; uint32_t drop_site, drop_thresh;   /* the forward's probability dropout (regenerated) */
  /* Lk <= 128 (one key tile): dQ is complete inside one CTA -- pass dq = NULL and dq_f16 (f16 [B*Lq, lddq16]) to
   * get it stored directly, without the f32 accumulation buffer.                                             */
  void *dq_f16; int lddq16;
} MtnAttnCoreBwdArgs;
int mtn_attn_core_bwd(const MtnAttnCoreBwdArgs *args, void *stream);

/* ---- generator tail / criterion backward ------------------------------------------------------------
 * log-softmax (mtn.py:68-69):  dz = dy - exp(y) * sum_v dy;  columns [V, lddz) of dz are zeroed.
 * label smoothing (label_smoothing.py:20-32 + KLDivLoss(sum)) from logits or log-probabilities z:
 *   dz_v = gscale * gout[0] * (T_r softmax(z)_v - t_v)   (t = smoothed target row, T_r its sum);
 * gout: device scalar (upstream gradient) or NULL.  workspace: >= 256 bytes.                               */
int mtn_log_softmax_bwd(const float *y, int ldy, const float *dy, int lddy, int rows, int V, float *dz,
                        int lddz, void *stream);
int mtn_label_smoothing_loss_bwd(const float *z, int ld, int rows, int V, const int64_t *target,
                                 int64_t padding_idx, float smoothing, float gscale, const float *gout,
                                 float *dz, int lddz, void *workspace, size_t workspace_bytes, void *stream);

/* ---- on-device self-check kernels (tests only) ---------------------------------
 * Plain one-thread-per-output CUDA kernels with the same f16-operand / f32-accumulate
 * arithmetic as the tensor-core kernels.  They exist so that tests can separate
 * "tcgen05/TMA layout bug" from "precision" on the GPU; the product path never
 * calls them.                                                                      */
int mtn_check_linear_fwd(const MtnLinearArgs *args, void *stream);
int mtn_check_attn_core_fwd(const MtnAttnCoreArgs *args, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MTN_B200_H_ */
