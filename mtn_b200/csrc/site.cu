// Sublayer-level entry points: one pre-norm residual attention site and one pre-norm
// residual feed-forward block, composed from the LayerNorm / linear / attention-core
// kernels.  These are the units SublayerConnection.forward wraps in the reference
// (mtn.py:125-127 around :248-267 and :279-280).  All scratch lives in the caller's
// workspace; the launch sequence is stream-ordered and graph-capturable.
#include <stdlib.h>

#include "host.h"

namespace mtn {

static inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

struct Carver {
  uint8_t* p;
  size_t left;
  void* take(size_t bytes) {
    bytes = align_up(bytes);
    if (bytes > left) return nullptr;
    void* r = p;
    p += bytes;
    left -= bytes;
    return r;
  }
};

}  // namespace mtn

extern "C" size_t mtn_attn_site_workspace_bytes(int B, int Lq, int Lk, int d) {
  using mtn::align_up;
  const size_t rq = (size_t)B * Lq, rk = (size_t)B * Lk;
  return 256 + align_up(rq * d * 2)          // LN(x) as f16
         + align_up(rq * 3 * d * 2)          // Q or [Q|K|V]
         + align_up(rk * 2 * d * 2)          // [K|V] of the memory (cross, not hoisted)
         + align_up(rq * d * 2);             // concat-head attention output
}

extern "C" int mtn_attn_site_fwd(const MtnAttnSiteArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->x && a->x_out && a->ln_a && a->ln_b && a->w_q && a->w_o, MTN_E_ARG,
              "attn_site: NULL pointer");
  MTN_REQUIRE(a->B > 0 && a->Lq > 0 && a->d > 0 && a->h > 0 && a->d % a->h == 0, MTN_E_SHAPE,
              "attn_site: B=%d Lq=%d d=%d h=%d", a->B, a->Lq, a->d, a->h);
  const bool self = (a->mem_f16 == nullptr && a->kv == nullptr);
  const int Lk = self ? a->Lq : a->Lk;
  MTN_REQUIRE(Lk > 0, MTN_E_SHAPE, "attn_site: Lk=%d", Lk);
  MTN_REQUIRE(a->workspace != nullptr, MTN_E_WORKSPACE, "attn_site: workspace is NULL");
  uintptr_t w0 = reinterpret_cast<uintptr_t>(a->workspace);
  const size_t skew = (256 - (w0 & 255)) & 255;
  MTN_REQUIRE(a->workspace_bytes > skew, MTN_E_WORKSPACE, "attn_site: workspace too small");
  Carver cv{reinterpret_cast<uint8_t*>(a->workspace) + skew, a->workspace_bytes - skew};
  const int d = a->d, dk = d / a->h;
  const size_t rq = (size_t)a->B * a->Lq, rk = (size_t)a->B * Lk;

  void* xn = cv.take(rq * d * 2);
  void* qbuf = cv.take(rq * (self ? 3 : 1) * d * 2);
  void* kvbuf = (!self && a->kv == nullptr) ? cv.take(rk * 2 * d * 2) : nullptr;
  void* obuf = cv.take(rq * d * 2);
  MTN_REQUIRE(xn && qbuf && obuf && (self || a->kv || kvbuf), MTN_E_WORKSPACE,
              "attn_site: workspace of %zu bytes too small (need %zu)", a->workspace_bytes,
              mtn_attn_site_workspace_bytes(a->B, a->Lq, Lk, d));

  // Cross site with its own (not hoisted) memory: the K/V projection of the memory does not depend on x, so it runs on
  // an internal side stream CONCURRENTLY with LayerNorm + Q projection (fork / join by events: stream-ordered for the
  // caller and legal inside CUDA-graph capture) -- it is half of the site's FLOPs and used to sit in front of the core.
  const bool own_kv = !self && a->kv == nullptr;
  static cudaStream_t side = nullptr;
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t main_st = static_cast<cudaStream_t>(stream);
  if (own_kv) {
    MTN_REQUIRE(a->w_kv != nullptr && a->mem_f16 != nullptr, MTN_E_ARG, "attn_site: w_kv / mem_f16 is NULL for a non-hoisted cross site");
    if (side == nullptr) {
      MTN_CHECK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
      MTN_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      MTN_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    MTN_CHECK_CUDA(cudaEventRecord(ev_fork, main_st));
    MTN_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    MtnLinearArgs kvl = {};
    kvl.A = a->mem_f16; kvl.lda = d; kvl.W = a->w_kv; kvl.ldw = d; kvl.bias = a->b_kv;
    kvl.M = (int)rk; kvl.N = 2 * d; kvl.K = d; kvl.act = MTN_ACT_NONE;
    kvl.out_f16 = kvbuf; kvl.ld16 = 2 * d;
    int rck = mtn_linear_fwd(&kvl, side);
    MTN_CHECK_CUDA(cudaEventRecord(ev_join, side));     // always joined, also on error, so capture stays well-formed
    if (rck) {
      cudaStreamWaitEvent(main_st, ev_join, 0);
      return rck;
    }
  }

  // Cross sites with full query tiles take the ONE-kernel form (csrc/site_fused.cu: Q projection + attention + output
  // projection + residual; measured 1.23-1.33x the launch sequence at Lq = 256, profiles/r02c_site_bench.txt):
  // LayerNorm -> fused kernel, the residual stream updated in place (x_out is seeded with x when they differ).
  static int fused_mode = -1;   // MTN_B200_SITE_FUSED: 0 never, 1 whenever supported, default: Lq >= 64 or Lk <= 64
  if (fused_mode < 0) {
    const char* e = getenv("MTN_B200_SITE_FUSED");
    fused_mode = (e == nullptr || e[0] == 'a') ? 2 : (e[0] == '0' ? 0 : 1);
  }
  if (!self && fused_mode != 0 && mtn_attn_site_fused_supported(d, a->h) && (fused_mode == 1 || a->Lq >= 64 || Lk <= 64)) {
    int rcf = mtn_layernorm_fwd(a->x, a->ln_a, a->ln_b, a->ln_eps, (int)rq, d, nullptr, xn, stream);
    if (rcf == 0 && a->x_out != a->x)
      rcf = cudaMemcpyAsync(a->x_out, a->x, rq * d * sizeof(float), cudaMemcpyDeviceToDevice, main_st) == cudaSuccess
                ? 0 : set_error(MTN_E_CUDA, "attn_site: cudaMemcpyAsync failed");
    if (own_kv) MTN_CHECK_CUDA(cudaStreamWaitEvent(main_st, ev_join, 0));
    if (rcf) return rcf;
    MtnAttnSiteFusedArgs f = {};
    f.B = a->B; f.Lq = a->Lq; f.Lk = Lk; f.d = d; f.h = a->h;
    f.xn_f16 = xn; f.ld_xn = d; f.x = a->x_out; f.ld_x = d;
    f.w_q = a->w_q; f.ld_wq = d; f.b_q = a->b_q; f.w_o = a->w_o; f.ld_wo = d; f.b_o = a->b_o;
    if (own_kv) { f.kv = kvbuf; f.ld_kv = 2 * d; f.kv_k_col = 0; f.kv_v_col = d; }
    else { f.kv = a->kv; f.ld_kv = a->ld_kv; f.kv_k_col = a->kv_k_col; f.kv_v_col = a->kv_v_col; }
    f.mask_bits = a->mask_bits; f.mask_rows_q = a->mask_rows_q;
    return mtn_attn_site_fused_fwd(&f, stream);
  }

  int rc = mtn_layernorm_fwd(a->x, a->ln_a, a->ln_b, a->ln_eps, (int)rq, d, nullptr, xn, stream);
  if (rc == 0) {
    MtnLinearArgs lin0 = {};
    lin0.A = xn; lin0.lda = d; lin0.W = a->w_q; lin0.ldw = d; lin0.bias = a->b_q;
    lin0.M = (int)rq; lin0.N = self ? 3 * d : d; lin0.K = d; lin0.act = MTN_ACT_NONE;
    lin0.out_f16 = qbuf; lin0.ld16 = lin0.N;
    rc = mtn_linear_fwd(&lin0, stream);
  }
  if (own_kv) MTN_CHECK_CUDA(cudaStreamWaitEvent(main_st, ev_join, 0));
  if (rc) return rc;

  MtnAttnCoreArgs core = {};
  core.B = a->B; core.h = a->h; core.Lq = a->Lq; core.Lk = Lk; core.d_k = dk;
  core.mask_bits = a->mask_bits; core.mask_rows_q = a->mask_rows_q;
  core.q = qbuf; core.ldq = self ? 3 * d : d;
  core.out = obuf; core.ldo = d;
  if (self) {
    core.k = static_cast<const uint8_t*>(qbuf) + (size_t)d * 2; core.ldk = 3 * d;
    core.v = static_cast<const uint8_t*>(qbuf) + (size_t)2 * d * 2; core.ldv = 3 * d;
  } else if (a->kv != nullptr) {
    core.k = static_cast<const uint8_t*>(a->kv) + (size_t)a->kv_k_col * 2; core.ldk = a->ld_kv;
    core.v = static_cast<const uint8_t*>(a->kv) + (size_t)a->kv_v_col * 2; core.ldv = a->ld_kv;
  } else {
    core.k = kvbuf; core.ldk = 2 * d;
    core.v = static_cast<const uint8_t*>(kvbuf) + (size_t)d * 2; core.ldv = 2 * d;
  }
  rc = mtn_attn_core_fwd(&core, stream);
  if (rc) return rc;

  MtnLinearArgs ol = {};
  ol.A = obuf; ol.lda = d; ol.W = a->w_o; ol.ldw = d; ol.bias = a->b_o;
  ol.M = (int)rq; ol.N = d; ol.K = d; ol.act = MTN_ACT_NONE;
  ol.addend = a->x; ol.ld_add = d; ol.add_period = 0;   // residual (mtn.py:127)
  ol.out_f32 = a->x_out; ol.ld32 = d;
  return mtn_linear_fwd(&ol, stream);
}

extern "C" size_t mtn_ffn_workspace_bytes(int rows, int d, int d_ff) {
  using mtn::align_up;
  return 256 + align_up((size_t)rows * d * 2) + align_up((size_t)rows * d_ff * 2);
}

extern "C" int mtn_ffn_fwd(const MtnFfnArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->x && a->x_out && a->ln_a && a->ln_b && a->w_1 && a->w_2, MTN_E_ARG, "ffn: NULL pointer");
  MTN_REQUIRE(a->rows > 0 && a->d > 0 && a->d_ff > 0, MTN_E_SHAPE, "ffn: rows=%d d=%d d_ff=%d", a->rows, a->d,
              a->d_ff);
  MTN_REQUIRE(a->workspace != nullptr, MTN_E_WORKSPACE, "ffn: workspace is NULL");
  uintptr_t w0 = reinterpret_cast<uintptr_t>(a->workspace);
  const size_t skew = (256 - (w0 & 255)) & 255;
  MTN_REQUIRE(a->workspace_bytes > skew, MTN_E_WORKSPACE, "ffn: workspace too small");
  Carver cv{reinterpret_cast<uint8_t*>(a->workspace) + skew, a->workspace_bytes - skew};
  void* xn = cv.take((size_t)a->rows * a->d * 2);
  void* hid = cv.take((size_t)a->rows * a->d_ff * 2);
  MTN_REQUIRE(xn && hid, MTN_E_WORKSPACE, "ffn: workspace of %zu bytes too small (need %zu)",
              a->workspace_bytes, mtn_ffn_workspace_bytes(a->rows, a->d, a->d_ff));

  int rc = mtn_layernorm_fwd(a->x, a->ln_a, a->ln_b, a->ln_eps, a->rows, a->d, nullptr, xn, stream);
  if (rc) return rc;
  MtnLinearArgs l1 = {};
  l1.A = xn; l1.lda = a->d; l1.W = a->w_1; l1.ldw = a->d; l1.bias = a->b_1;
  l1.M = a->rows; l1.N = a->d_ff; l1.K = a->d; l1.act = MTN_ACT_RELU;     // mtn.py:280
  l1.out_f16 = hid; l1.ld16 = a->d_ff;
  rc = mtn_linear_fwd(&l1, stream);
  if (rc) return rc;
  MtnLinearArgs l2 = {};
  l2.A = hid; l2.lda = a->d_ff; l2.W = a->w_2; l2.ldw = a->d_ff; l2.bias = a->b_2;
  l2.M = a->rows; l2.N = a->d; l2.K = a->d_ff; l2.act = MTN_ACT_NONE;
  l2.addend = a->x; l2.ld_add = a->d; l2.add_period = 0;                  // residual (mtn.py:127)
  l2.out_f32 = a->x_out; l2.ld32 = a->d;
  return mtn_linear_fwd(&l2, stream);
}
