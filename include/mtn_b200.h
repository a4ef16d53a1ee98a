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

#define MTN_B200_ABI_VERSION 2

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

/* ---- video-feature preparation (boundary: Batch, data_utils.py:28-30) ----------------
 * ft: [frames, F] f32 raw features.  mask[frame] = any(ft[frame, :] != 1.0) (all-ones frames are
 * padding); padded frames are zeroed; output as f16 (video-encoder operand) and/or f32.        */
int mtn_feature_prep_fwd(const float *ft, int frames, int F, uint8_t *mask, void *out_f16,
                         float *out_f32, void *stream);

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
} MtnLinearArgs;
int mtn_linear_fwd(const MtnLinearArgs *args, void *stream);

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
} MtnAttnCoreArgs;
int mtn_attn_core_fwd(const MtnAttnCoreArgs *args, void *stream);

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
