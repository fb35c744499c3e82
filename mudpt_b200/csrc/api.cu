// C ABI of libmudpt_b200.so (include/mudpt_b200.h): handle, frozen-weight conversion, tower
// forward / dgrad-only backward orchestration.  All launches go to the caller's stream; no
// device synchronisation happens here (except the one-time read of logit_scale in set_weight).
#include "../../include/mudpt_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "attention.h"
#include "augment.h"
#include "gemm.h"
#include "head.h"
#include "launch_count.h"
#include "prompt.h"
#include "peer.h"
#include "rowops.h"

namespace mudpt {
std::atomic<long long> g_launch_counter{0};
}

using namespace mudpt;
typedef __nv_bfloat16 bf16;

static thread_local std::string g_err;

namespace {

constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default (clip/model.py:164)

struct Layer {
  // bf16 GEMM operands: forward uses the [out,in] weight as stored, dgrad the transposed copy
  bf16 *w_in = nullptr, *w_in_t = nullptr, *w_out = nullptr, *w_out_t = nullptr;
  bf16 *w_fc = nullptr, *w_fc_t = nullptr, *w_pr = nullptr, *w_pr_t = nullptr;
  float *b_in = nullptr, *b_out = nullptr, *b_fc = nullptr, *b_pr = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  // LayerNorm folded into the Linear that follows it (rowops.cu fold_layernorm): fp32 sources kept for re-folding,
  // folded bf16 weight (+ transposed), folded bias, column sums, interleaved (colsum, bias')
  float *w_in_f32 = nullptr, *w_fc_f32 = nullptr;
  bf16 *w_in_ln = nullptr, *w_in_t_ln = nullptr, *w_fc_ln = nullptr, *w_fc_t_ln = nullptr;
  float *b_in_ln = nullptr, *b_fc_ln = nullptr, *cs_in = nullptr, *cs_fc = nullptr;
  float2 *sb_in = nullptr, *sb_fc = nullptr;
  int have = 0;  // bit per tensor
};

struct Tower {
  int d = 0, H = 0, layers = 0, n_ctx = 0, row0 = 0, depth = 0;
  bool causal = false;
  int L = 0;        // current tokens per sequence
  int S = 0;        // current sequences
  size_t cap_rows = 0;
  std::vector<Layer> lw;
  bool fold_dirty = true;  // a tensor that enters the LayerNorm folding changed since the last fold
  // saved activations, per layer
  std::vector<float*> x_in, x_mid;
  std::vector<bf16*> qkv, o, h;
  std::vector<float*> lse;
  // fused LayerNorm: bf16 copies of the LN inputs (GEMM A operands) + their per-64-column partial statistics
  std::vector<bf16*> xb_in, xb_mid;
  std::vector<float2*> st_in, st_mid;
  float2* dots = nullptr;  // row dots for the fused LayerNorm backward [rows, dot_cap]
  int dot_cap = 0;
  // transients
  bf16 *a_buf = nullptr, *g_buf = nullptr, *dh_buf = nullptr, *do_buf = nullptr, *dqkv_buf = nullptr, *dx_bf16 = nullptr;
  float *dx = nullptr, *dsum = nullptr, *splice_ws = nullptr, *head_ws = nullptr;  // head_ws: feature_head_workspace_floats(S, d, e)
  size_t head_cap = 0;  // floats in head_ws (sized by the sequence count, which may grow while rows = S * L shrinks)
  // exact work skipping: after the last block's attention only one row per sequence is consumed
  // (CLS, clip/model.py:548; EOT, trainers/mudpt.py:154), so its out-proj / MLP run on S gathered rows
  const int* sel_rows = nullptr;  // consumed row per sequence (null: row 0)
  size_t pr_cap = 0;              // sequences the buffers below hold
  std::vector<void*> pr_allocs;
  bf16 *pr_o = nullptr, *pr_xmb = nullptr, *pr_h = nullptr, *pr_g = nullptr, *pr_a = nullptr, *pr_dxb = nullptr, *pr_dh = nullptr,
       *pr_do = nullptr;
  float *pr_x = nullptr, *pr_xm = nullptr, *pr_xo = nullptr, *pr_dx = nullptr;
  float2 *pr_stm = nullptr, *pr_dots = nullptr;
  bool fwd_pruned = false;
  // ... and below block 0 only the spliced prompt rows carry a gradient anyone reads (everything else of the tower input
  // is frozen: patch / token embeddings, cls, positional embeddings), so block 0's d in-proj and LN1 backward run on
  // the S * n_ctx gathered prompt rows
  int* p0_rows = nullptr;  // absolute row of prompt token j of sequence s: s * L + row0 + j
  int p0_S = 0, p0_L = 0;  // shape p0_rows was built for
  size_t p0_cap = 0;       // rows the buffers below hold
  std::vector<void*> p0_allocs;
  bf16 *p0_dqkv = nullptr, *p0_a = nullptr, *p0_dxb = nullptr;
  float *p0_x = nullptr, *p0_dx = nullptr;
  std::vector<void*> ws_allocs;  // everything sized by cap_rows: released when the tower outgrows it
  GemmWorkspace gws;             // stream-K scratch of this tower's stream
  bool fwd_done = false;
  bool fwd_fused = false;  // the saved activations belong to the fused-LayerNorm formulation
  int first_splice = 0;  // 0: layer 0 splices prompts[0]; 1: layer-0 rows kept as given
};

}  // namespace

// Optional per-kernel-class timing with CUDA events on the launch stream (bench.py's roofline leg).
enum ProfCat { PC_GEMM = 0, PC_ATTN_FWD, PC_ATTN_BWD, PC_LN_FWD, PC_LN_BWD, PC_SPLICE, PC_HEAD, PC_STEM,
               // the eight GEMMs of a block, forward and dgrad (PC_GEMM: everything else -- patch embedding, pruned tail)
               PC_GEMM_QKV, PC_GEMM_OUT, PC_GEMM_FC, PC_GEMM_PROJ, PC_GEMM_DPROJ, PC_GEMM_DFC, PC_GEMM_DOUT, PC_GEMM_DQKV, PC_COUNT };
struct ProfRec { int cat; cudaEvent_t e0, e1; double flops, bytes; };
struct Prof {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};

struct mudpt_handle {
  mudpt_config cfg;
  std::string err;
  Prof prof;
  Tower vis, txt;
  // vision stem / heads
  bf16* w_conv = nullptr;  // [dv, Kp] bf16 (Kp = 3*p*p padded to a multiple of 8)
  int Kp = 0;
  float *cls = nullptr, *pos_v = nullptr, *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr,
        *ln_post_b = nullptr, *proj_v = nullptr;
  float *pos_t = nullptr, *ln_final_g = nullptr, *ln_final_b = nullptr, *proj_t = nullptr;
  float logit_scale_exp = 0.f;
  int stem_have = 0;
  bf16* patches = nullptr;
  size_t patches_cap = 0;
  int* eot = nullptr;
  size_t eot_cap = 0;
  float* head_ws = nullptr;
  size_t head_ws_cap = 0;
  std::vector<void*> allocs;
  long long launches_at_create = 0;
  // options (mudpt_set_option): LayerNorm folded into the GEMMs (default) or the stand-alone LN kernels -- the
  // fallback for checkpoints whose residual stream has a row mean far above its spread (bf16(x) instead of
  // bf16(LN(x)) as the GEMM operand would lose the signal there)
  int ln_fused = -1;  // -1 = by tower size (see tower_forward), 0 = never, 1 = always
  // LayerNorm dgrad in the dgrad GEMMs' epilogues (EPI_LN_BWD) instead of the stand-alone kernel.  Off by default:
  // measured on B200 the row dots it needs cost more at their producers (GELU' epilogue +66 us, attention backward
  // +82 us per block at the cfg-2 shapes) than the fused epilogue saves (60 us): profiles/r02_ln_fusion_ab.txt
  bool ln_bwd_fused = false;
  bool ln_bwd_bf16x = true;  // env MUDPT_LN_BWD_BF16X
  // exact work skipping: the last block's out-proj / MLP on the CLS / EOT rows only
  bool prune = true;
  // Gradient of the residual stream kept in bf16 between the LayerNorm backward kernels (the dgrad GEMMs read that copy
  // as their A operand anyway); fp32 rows are written for the deep-prompt window only, which is all the splice backward
  // reads.  8 instead of 14 B per element on the HBM-bound LayerNorm backward; each of the 2 x layers updates rounds the
  // stream to bf16 (2^-9 relative): measured effect on the prompt gradients in profiles/r02_grad_stream_bf16.txt.
  // Not used when the dense input gradient is requested (CoCoOp's d_x0) or with ln_bwd_fused.  env MUDPT_GRAD_BF16
  bool grad_bf16 = true;
};

namespace {

int fail(mudpt_handle* h, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  g_err = buf;
  return -1;
}

#define CK(h, expr)                                   \
  do {                                                \
    const char* _e = (expr);                          \
    if (_e) return fail((h), "%s (%s:%d)", _e, __FILE__, __LINE__); \
  } while (0)

#define CUDA_OK(h, expr)                                                                    \
  do {                                                                                      \
    cudaError_t _c = (expr);                                                                \
    if (_c != cudaSuccess) return fail((h), "%s: %s (%s:%d)", #expr, cudaGetErrorString(_c), __FILE__, __LINE__); \
  } while (0)

// CK with optional event bracketing: category, algorithmic FLOPs and bytes of the launch
#define CKP(h, st, cat, fl, by, expr)                                         \
  do {                                                                        \
    cudaEvent_t _e0 = nullptr, _e1 = nullptr;                                 \
    if ((h)->prof.on) { _e0 = (h)->prof.get(); _e1 = (h)->prof.get(); cudaEventRecord(_e0, (st)); } \
    const char* _e = (expr);                                                  \
    if (_e) return fail((h), "%s (%s:%d)", _e, __FILE__, __LINE__);           \
    if ((h)->prof.on) { cudaEventRecord(_e1, (st)); (h)->prof.recs.push_back(ProfRec{(cat), _e0, _e1, (double)(fl), (double)(by)}); } \
  } while (0)

template <typename T>
cudaError_t dev_alloc(mudpt_handle* h, T** p, size_t n, std::vector<void*>* group = nullptr) {
  void* q = nullptr;
  cudaError_t c = cudaMalloc(&q, n * sizeof(T) + 256);
  if (c != cudaSuccess) return c;
  (group ? *group : h->allocs).push_back(q);
  *p = reinterpret_cast<T*>(q);
  return cudaSuccess;
}

bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) != 0 : dflt;
}

// (re)allocate the activation workspace of a tower for S sequences of L tokens
int ensure_tower(mudpt_handle* h, Tower& t, int S, int L) {
  const size_t rows = static_cast<size_t>(S) * L;
  t.S = S;
  t.L = L;
  const size_t head_need = feature_head_workspace_floats(S, t.d, h->cfg.embed_dim);
  if (head_need > t.head_cap) {  // gathered CLS / EOT rows of the feature head + split-K partials
    CUDA_OK(h, dev_alloc(h, &t.head_ws, head_need));
    t.head_cap = head_need;
  }
  if (!t.gws.partials) {
    CUDA_OK(h, dev_alloc(h, &t.gws.partials, gemm_workspace_partial_bytes() / sizeof(float)));
    CUDA_OK(h, dev_alloc(h, &t.gws.flags, gemm_workspace_flag_bytes() / sizeof(unsigned)));
    CUDA_OK(h, cudaMemset(t.gws.flags, 0, gemm_workspace_flag_bytes()));
  }
  if (static_cast<size_t>(S) > t.pr_cap) {
    for (void* p : t.pr_allocs) cudaFree(p);
    t.pr_allocs.clear();
    gemm_clear_tensor_map_cache();
    std::vector<void*>* g = &t.pr_allocs;
    const size_t Ss = S, d = t.d;
    const int dots_mlp = (4 * t.d + gemm_dots_span(4 * t.d) - 1) / gemm_dots_span(4 * t.d);
    CUDA_OK(h, dev_alloc(h, &t.pr_o, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_xmb, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_h, Ss * 4 * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_g, Ss * 4 * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_a, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_dxb, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_dh, Ss * 4 * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_do, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_x, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_xm, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_xo, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_dx, Ss * d, g));
    CUDA_OK(h, dev_alloc(h, &t.pr_stm, Ss * (d / 64), g));
    CUDA_OK(h, dev_alloc(h, &t.pr_dots, Ss * dots_mlp, g));
    t.pr_cap = S;
    t.fwd_done = false;
  }
  const size_t R0 = static_cast<size_t>(S) * t.n_ctx;
  if (R0 > t.p0_cap) {
    for (void* p : t.p0_allocs) cudaFree(p);
    t.p0_allocs.clear();
    gemm_clear_tensor_map_cache();
    std::vector<void*>* g = &t.p0_allocs;
    CUDA_OK(h, dev_alloc(h, &t.p0_rows, R0, g));
    CUDA_OK(h, dev_alloc(h, &t.p0_dqkv, R0 * 3 * t.d, g));
    CUDA_OK(h, dev_alloc(h, &t.p0_a, R0 * t.d, g));
    CUDA_OK(h, dev_alloc(h, &t.p0_dxb, R0 * t.d, g));
    CUDA_OK(h, dev_alloc(h, &t.p0_x, R0 * t.d, g));
    CUDA_OK(h, dev_alloc(h, &t.p0_dx, R0 * t.d, g));
    t.p0_cap = R0;
    t.p0_S = 0;
  }
  if (R0 > 0 && (t.p0_S != S || t.p0_L != L)) {
    std::vector<int> r(R0);
    for (int q = 0; q < S; ++q)
      for (int j = 0; j < t.n_ctx; ++j) r[static_cast<size_t>(q) * t.n_ctx + j] = q * L + t.row0 + j;
    CUDA_OK(h, cudaMemcpy(t.p0_rows, r.data(), R0 * sizeof(int), cudaMemcpyHostToDevice));  // (set-up: shapes change rarely)
    t.p0_S = S;
    t.p0_L = L;
  }
  if (rows <= t.cap_rows) return 0;
  // grow: the outgrown buffers are released first (cudaFree waits for the device, so nothing in flight uses them);
  // a CoCoOp-style caller whose B x C changes would otherwise pile up GBs until mudpt_destroy
  for (void* p : t.ws_allocs) cudaFree(p);
  t.ws_allocs.clear();
  gemm_clear_tensor_map_cache();
  const size_t d = t.d;
  const size_t parts = d / 64;
  std::vector<void*>* g = &t.ws_allocs;
  t.x_in.assign(t.layers + 1, nullptr);
  t.x_mid.assign(t.layers, nullptr);
  t.qkv.assign(t.layers, nullptr);
  t.o.assign(t.layers, nullptr);
  t.h.assign(t.layers, nullptr);
  t.lse.assign(t.layers, nullptr);
  t.xb_in.assign(t.layers + 1, nullptr);
  t.xb_mid.assign(t.layers, nullptr);
  t.st_in.assign(t.layers + 1, nullptr);
  t.st_mid.assign(t.layers, nullptr);
  for (int i = 0; i <= t.layers; ++i) {
    CUDA_OK(h, dev_alloc(h, &t.x_in[i], rows * d, g));
    CUDA_OK(h, dev_alloc(h, &t.xb_in[i], rows * d, g));
    CUDA_OK(h, dev_alloc(h, &t.st_in[i], rows * parts, g));
  }
  for (int i = 0; i < t.layers; ++i) {
    CUDA_OK(h, dev_alloc(h, &t.x_mid[i], rows * d, g));
    CUDA_OK(h, dev_alloc(h, &t.xb_mid[i], rows * d, g));
    CUDA_OK(h, dev_alloc(h, &t.st_mid[i], rows * parts, g));
    CUDA_OK(h, dev_alloc(h, &t.qkv[i], rows * 3 * d, g));
    CUDA_OK(h, dev_alloc(h, &t.o[i], rows * d, g));
    CUDA_OK(h, dev_alloc(h, &t.h[i], rows * 4 * d, g));
    CUDA_OK(h, dev_alloc(h, &t.lse[i], rows * t.H, g));
  }
  const int dots_mlp = (4 * t.d + gemm_dots_span(4 * t.d) - 1) / gemm_dots_span(4 * t.d);
  t.dot_cap = dots_mlp > 3 * t.H ? dots_mlp : 3 * t.H;
  CUDA_OK(h, dev_alloc(h, &t.dots, rows * t.dot_cap, g));
  CUDA_OK(h, dev_alloc(h, &t.a_buf, rows * d, g));
  CUDA_OK(h, dev_alloc(h, &t.g_buf, rows * 4 * d, g));
  CUDA_OK(h, dev_alloc(h, &t.dh_buf, rows * 4 * d, g));
  CUDA_OK(h, dev_alloc(h, &t.do_buf, rows * d, g));
  CUDA_OK(h, dev_alloc(h, &t.dqkv_buf, rows * 3 * d, g));
  CUDA_OK(h, dev_alloc(h, &t.dx_bf16, rows * d, g));
  CUDA_OK(h, dev_alloc(h, &t.dx, rows * d, g));
  CUDA_OK(h, dev_alloc(h, &t.dsum, rows * t.H, g));
  if (!t.splice_ws) {
    const size_t nws = splice_bwd_workspace_floats(t.n_ctx > 0 ? t.n_ctx : 1, t.d);
    CUDA_OK(h, dev_alloc(h, &t.splice_ws, nws));
    CUDA_OK(h, cudaMemset(t.splice_ws, 0, nws * sizeof(float)));  // (the kernel's tickets start at zero and are left at zero)
  }
  t.cap_rows = rows;
  t.fwd_done = false;
  return 0;
}

// ---------------------------------------------------------------------------- weights
int store_f32(mudpt_handle* h, float** dst, const float* src, int64_t numel, int64_t expect, const char* name,
              cudaStream_t st) {
  if (numel != expect) return fail(h, "weight %s: expected %lld elements, got %lld", name, (long long)expect, (long long)numel);
  if (!*dst) CUDA_OK(h, dev_alloc(h, dst, static_cast<size_t>(numel)));
  CUDA_OK(h, cudaMemcpyAsync(*dst, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}
// [rows(out), cols(in)] fp32 -> bf16 as stored + bf16 transposed
int store_gemm_weight(mudpt_handle* h, bf16** w, bf16** wt, const float* src, int64_t numel, int rows, int cols,
                      const char* name, cudaStream_t st) {
  if (numel != static_cast<int64_t>(rows) * cols)
    return fail(h, "weight %s: expected %d x %d elements, got %lld", name, rows, cols, (long long)numel);
  if (!*w) CUDA_OK(h, dev_alloc(h, w, static_cast<size_t>(numel)));
  CK(h, cast_to_bf16(src, *w, static_cast<size_t>(numel), st));
  if (wt) {
    if (!*wt) CUDA_OK(h, dev_alloc(h, wt, static_cast<size_t>(numel)));
    CK(h, transpose_cast_bf16(src, *wt, rows, cols, st));
  }
  return 0;
}

enum LayerBits {
  LB_W_IN = 1 << 0, LB_B_IN = 1 << 1, LB_W_OUT = 1 << 2, LB_B_OUT = 1 << 3, LB_W_FC = 1 << 4, LB_B_FC = 1 << 5,
  LB_W_PR = 1 << 6, LB_B_PR = 1 << 7, LB_LN1_G = 1 << 8, LB_LN1_B = 1 << 9, LB_LN2_G = 1 << 10, LB_LN2_B = 1 << 11,
  LB_ALL = (1 << 12) - 1
};
enum StemBits {
  SB_CONV = 1 << 0, SB_CLS = 1 << 1, SB_POS_V = 1 << 2, SB_LNPRE_G = 1 << 3, SB_LNPRE_B = 1 << 4, SB_LNPOST_G = 1 << 5,
  SB_LNPOST_B = 1 << 6, SB_PROJ_V = 1 << 7, SB_POS_T = 1 << 8, SB_LNF_G = 1 << 9, SB_LNF_B = 1 << 10, SB_PROJ_T = 1 << 11,
  SB_SCALE = 1 << 12, SB_ALL = (1 << 13) - 1
};

int set_block_weight(mudpt_handle* h, Tower& t, int layer, const char* sub, const float* data, int64_t numel,
                     const char* full, cudaStream_t st) {
  if (layer < 0 || layer >= t.layers) return fail(h, "weight %s: layer out of range", full);
  Layer& l = t.lw[layer];
  const int d = t.d;
  int rc = 0;
  if (!strcmp(sub, "attn.in_proj_weight")) {
    rc = store_gemm_weight(h, &l.w_in, &l.w_in_t, data, numel, 3 * d, d, full, st);
    if (!rc) rc = store_f32(h, &l.w_in_f32, data, numel, static_cast<int64_t>(3) * d * d, full, st);
    l.have |= LB_W_IN;
  }
  else if (!strcmp(sub, "attn.in_proj_bias")) { rc = store_f32(h, &l.b_in, data, numel, 3 * d, full, st); l.have |= LB_B_IN; }
  else if (!strcmp(sub, "attn.out_proj.weight")) { rc = store_gemm_weight(h, &l.w_out, &l.w_out_t, data, numel, d, d, full, st); l.have |= LB_W_OUT; }
  else if (!strcmp(sub, "attn.out_proj.bias")) { rc = store_f32(h, &l.b_out, data, numel, d, full, st); l.have |= LB_B_OUT; }
  else if (!strcmp(sub, "mlp.c_fc.weight")) {
    rc = store_gemm_weight(h, &l.w_fc, &l.w_fc_t, data, numel, 4 * d, d, full, st);
    if (!rc) rc = store_f32(h, &l.w_fc_f32, data, numel, static_cast<int64_t>(4) * d * d, full, st);
    l.have |= LB_W_FC;
  }
  else if (!strcmp(sub, "mlp.c_fc.bias")) { rc = store_f32(h, &l.b_fc, data, numel, 4 * d, full, st); l.have |= LB_B_FC; }
  else if (!strcmp(sub, "mlp.c_proj.weight")) { rc = store_gemm_weight(h, &l.w_pr, &l.w_pr_t, data, numel, d, 4 * d, full, st); l.have |= LB_W_PR; }
  else if (!strcmp(sub, "mlp.c_proj.bias")) { rc = store_f32(h, &l.b_pr, data, numel, d, full, st); l.have |= LB_B_PR; }
  else if (!strcmp(sub, "ln_1.weight")) { rc = store_f32(h, &l.ln1_g, data, numel, d, full, st); l.have |= LB_LN1_G; }
  else if (!strcmp(sub, "ln_1.bias")) { rc = store_f32(h, &l.ln1_b, data, numel, d, full, st); l.have |= LB_LN1_B; }
  else if (!strcmp(sub, "ln_2.weight")) { rc = store_f32(h, &l.ln2_g, data, numel, d, full, st); l.have |= LB_LN2_G; }
  else if (!strcmp(sub, "ln_2.bias")) { rc = store_f32(h, &l.ln2_b, data, numel, d, full, st); l.have |= LB_LN2_B; }
  else return 1;
  t.fold_dirty = true;
  return rc;
}

// LayerNorm -> Linear folding of every block (once per weight load)
int ensure_folded(mudpt_handle* h, Tower& t, cudaStream_t st) {
  if (!t.fold_dirty) return 0;
  const int d = t.d;
  for (Layer& l : t.lw) {
    if (!l.w_in_ln) {
      CUDA_OK(h, dev_alloc(h, &l.w_in_ln, static_cast<size_t>(3) * d * d));
      CUDA_OK(h, dev_alloc(h, &l.w_in_t_ln, static_cast<size_t>(3) * d * d));
      CUDA_OK(h, dev_alloc(h, &l.w_fc_ln, static_cast<size_t>(4) * d * d));
      CUDA_OK(h, dev_alloc(h, &l.w_fc_t_ln, static_cast<size_t>(4) * d * d));
      CUDA_OK(h, dev_alloc(h, &l.b_in_ln, static_cast<size_t>(3) * d));
      CUDA_OK(h, dev_alloc(h, &l.b_fc_ln, static_cast<size_t>(4) * d));
      CUDA_OK(h, dev_alloc(h, &l.cs_in, static_cast<size_t>(3) * d));
      CUDA_OK(h, dev_alloc(h, &l.cs_fc, static_cast<size_t>(4) * d));
      CUDA_OK(h, dev_alloc(h, &l.sb_in, static_cast<size_t>(3) * d));
      CUDA_OK(h, dev_alloc(h, &l.sb_fc, static_cast<size_t>(4) * d));
    }
    CK(h, fold_layernorm(l.w_in_f32, l.ln1_g, l.ln1_b, l.b_in, l.w_in_ln, l.w_in_t_ln, l.b_in_ln, l.cs_in, l.sb_in, 3 * d, d, st));
    CK(h, fold_layernorm(l.w_fc_f32, l.ln2_g, l.ln2_b, l.b_fc, l.w_fc_ln, l.w_fc_t_ln, l.b_fc_ln, l.cs_fc, l.sb_fc, 4 * d, d, st));
  }
  t.fold_dirty = false;
  return 0;
}

// ---------------------------------------------------------------------------- tower passes
// Per block (clip/model.py:275-301): [splice] ; x += attn(ln_1(x)) ; x += c_proj(QuickGELU(c_fc(ln_2(x)))).
//
// Fused formulation (default): no LayerNorm kernel and no splice kernel inside the layer loop.  The residual GEMMs
// (out-proj, c_proj) write the new residual row as fp32 + bf16 + partial statistics (and the NEXT block's spliced
// prompt rows in place of the computed ones); the QKV / c_fc GEMMs read the bf16 row as their A operand and apply
// the normalisation in their epilogue (EPI_LN_*: gamma folded into the weight, mean as a rank-1 correction).
// Precondition of the fused path: x_in[0], xb_in[0], st_in[0] hold the tower input (layer-0 splice excepted).
int tower_forward(mudpt_handle* h, Tower& t, const float* prompts, int first_splice_layer, cudaStream_t st) {
  const int M = t.S * t.L, d = t.d;
  const double Md = static_cast<double>(M), dd = d;
  const double attn_fl = 4.0 * t.S * t.H * static_cast<double>(t.L) * t.L * 64.0;  // QK^T + PV, dense count
  // Measured on B200 (profiles/r02_ln_fusion_ab.txt): folding the LayerNorm into the GEMMs saves its 4 B/element read
  // but makes the four forward epilogues heavier (+2 B/element bf16 copy, statistics, rank-1 correction).  For the
  // large HBM-bound tower (text, 77,000 rows) that is a gain; for ~6,000-row towers, whose 25-40 us GEMMs are
  // launch / epilogue bound, the stand-alone LayerNorm kernels are faster: choose by row count unless told otherwise.
  const bool fused = h->ln_fused == 1 || (h->ln_fused < 0 && M >= 32768);
  const bool pruned = h->prune;
  const int parts = d / 64;
  auto spliced = [&](int i) { return i < t.depth && i >= first_splice_layer && t.n_ctx > 0 && i < t.layers; };
  if (fused && ensure_folded(h, t, st)) return -1;
  for (int i = 0; i < t.layers; ++i) {
    const Layer& w = t.lw[i];
    // the stand-alone LayerNorm does the block's splice in the same pass; the fused form splices in the previous block's
    // residual epilogue (layer 0: a kernel of its own, with the statistics of the spliced rows)
    if (spliced(i) && fused && i == 0) {
      const float* pr = prompts + static_cast<size_t>(i) * t.n_ctx * d;
      CKP(h, st, PC_SPLICE, 0, 2.0 * t.S * t.n_ctx * dd * 4, splice_fwd_stats(t.x_in[i], t.xb_in[i], t.st_in[i], pr, t.S, t.L, t.row0, t.n_ctx, d, st));
    }
    // x + attn(ln_1(x))   (clip/model.py:299)
    GemmEpilogue e1;
    e1.out0 = t.qkv[i]; e1.ldc = 3 * d;
    if (fused) {
      e1.mode = EPI_LN_BF16; e1.bias = w.b_in_ln; e1.colsum = w.cs_in; e1.ln_stats = t.st_in[i]; e1.ln_parts = parts; e1.ln_width = d; e1.ln_eps = kLnEps;
      CKP(h, st, PC_GEMM_QKV, 2.0 * Md * 3 * dd * dd, 2 * (Md * dd + 3 * dd * dd + Md * 3 * dd),
          gemm_bf16_tn(t.xb_in[i], d, w.w_in_ln, d, e1, M, 3 * d, d, st, &t.gws));
    } else {
      CKP(h, st, PC_LN_FWD, 0, Md * dd * 6,
          layernorm_fwd_splice(t.x_in[i], spliced(i) ? prompts + static_cast<size_t>(i) * t.n_ctx * d : nullptr, t.L, t.row0,
                               spliced(i) ? t.n_ctx : 0, w.ln1_g, w.ln1_b, t.a_buf, true, M, d, kLnEps, st));
      e1.mode = EPI_BF16; e1.bias = w.b_in;
      CKP(h, st, PC_GEMM_QKV, 2.0 * Md * 3 * dd * dd, 2 * (Md * dd + 3 * dd * dd + Md * 3 * dd),
          gemm_bf16_tn(t.a_buf, d, w.w_in, d, e1, M, 3 * d, d, st, &t.gws));
    }
    CKP(h, st, PC_ATTN_FWD, attn_fl, Md * dd * 2 * 4, attention_fwd(t.qkv[i], t.o[i], t.lse[i], t.S, t.L, t.H, d, t.causal, st));
    if (pruned && i == t.layers - 1) {
      // the block's output is consumed on one row per sequence: out-proj and the MLP on S gathered rows
      const int S = t.S;
      const double Sd = S;
      CKP(h, st, PC_SPLICE, 0, Sd * dd * 12, gather_rows(t.x_in[i], t.o[i], t.sel_rows, S, t.L, t.pr_x, t.pr_o, d, st));
      GemmEpilogue p2;
      p2.out0 = t.pr_xm; p2.bias = w.b_out; p2.resid = t.pr_x; p2.ldc = d;
      if (fused) { p2.mode = EPI_RESID_STATS; p2.out2 = t.pr_xmb; p2.stats_out = t.pr_stm; }
      else p2.mode = EPI_RESID_F32;
      CKP(h, st, PC_GEMM, 2.0 * Sd * dd * dd, 2 * (Sd * dd + dd * dd) + 10 * Sd * dd, gemm_bf16_tn(t.pr_o, d, w.w_out, d, p2, S, d, d, st, &t.gws));
      GemmEpilogue p3;
      p3.out0 = t.pr_h; p3.out1 = t.pr_g; p3.ldc = 4 * d;
      if (fused) {
        p3.mode = EPI_LN_GELU; p3.bias = w.b_fc_ln; p3.colsum = w.cs_fc; p3.ln_stats = t.pr_stm; p3.ln_parts = parts; p3.ln_width = d; p3.ln_eps = kLnEps;
        CKP(h, st, PC_GEMM, 2.0 * Sd * 4 * dd * dd, 2 * (Sd * dd + 4 * dd * dd + 2 * Sd * 4 * dd), gemm_bf16_tn(t.pr_xmb, d, w.w_fc_ln, d, p3, S, 4 * d, d, st, &t.gws));
      } else {
        CKP(h, st, PC_LN_FWD, 0, Sd * dd * 6, layernorm_fwd(t.pr_xm, w.ln2_g, w.ln2_b, t.pr_a, true, S, d, kLnEps, st));
        p3.mode = EPI_GELU; p3.bias = w.b_fc;
        CKP(h, st, PC_GEMM, 2.0 * Sd * 4 * dd * dd, 2 * (Sd * dd + 4 * dd * dd + 2 * Sd * 4 * dd), gemm_bf16_tn(t.pr_a, d, w.w_fc, d, p3, S, 4 * d, d, st, &t.gws));
      }
      GemmEpilogue p4;
      p4.mode = EPI_RESID_F32; p4.out0 = t.pr_xo; p4.bias = w.b_pr; p4.resid = t.pr_xm; p4.ldc = d;
      CKP(h, st, PC_GEMM, 2.0 * Sd * 4 * dd * dd, 2 * (Sd * 4 * dd + 4 * dd * dd) + 8 * Sd * dd, gemm_bf16_tn(t.pr_g, 4 * d, w.w_pr, 4 * d, p4, S, d, 4 * d, st, &t.gws));
      break;
    }
    GemmEpilogue e2;
    e2.out0 = t.x_mid[i]; e2.bias = w.b_out; e2.resid = t.x_in[i]; e2.ldc = d;
    if (fused) { e2.mode = EPI_RESID_STATS; e2.out2 = t.xb_mid[i]; e2.stats_out = t.st_mid[i]; }
    else e2.mode = EPI_RESID_F32;
    CKP(h, st, PC_GEMM_OUT, 2.0 * Md * dd * dd, 2 * (Md * dd + dd * dd) + (fused ? 10 : 8) * Md * dd,
        gemm_bf16_tn(t.o[i], d, w.w_out, d, e2, M, d, d, st, &t.gws));
    // x + c_proj(QuickGELU(c_fc(ln_2(x))))   (clip/model.py:300)
    GemmEpilogue e3;
    e3.out0 = t.h[i]; e3.out1 = t.g_buf; e3.ldc = 4 * d;
    if (fused) {
      e3.mode = EPI_LN_GELU; e3.bias = w.b_fc_ln; e3.colsum = w.cs_fc; e3.ln_stats = t.st_mid[i]; e3.ln_parts = parts; e3.ln_width = d; e3.ln_eps = kLnEps;
      CKP(h, st, PC_GEMM_FC, 2.0 * Md * 4 * dd * dd, 2 * (Md * dd + 4 * dd * dd + 2 * Md * 4 * dd),
          gemm_bf16_tn(t.xb_mid[i], d, w.w_fc_ln, d, e3, M, 4 * d, d, st, &t.gws));
    } else {
      CKP(h, st, PC_LN_FWD, 0, Md * dd * 6, layernorm_fwd(t.x_mid[i], w.ln2_g, w.ln2_b, t.a_buf, true, M, d, kLnEps, st));
      e3.mode = EPI_GELU; e3.bias = w.b_fc;
      CKP(h, st, PC_GEMM_FC, 2.0 * Md * 4 * dd * dd, 2 * (Md * dd + 4 * dd * dd + 2 * Md * 4 * dd),
          gemm_bf16_tn(t.a_buf, d, w.w_fc, d, e3, M, 4 * d, d, st, &t.gws));
    }
    GemmEpilogue e4;
    e4.out0 = t.x_in[i + 1]; e4.bias = w.b_pr; e4.resid = t.x_mid[i]; e4.ldc = d;
    if (fused) {
      e4.mode = EPI_RESID_STATS; e4.out2 = t.xb_in[i + 1]; e4.stats_out = t.st_in[i + 1];
      if (spliced(i + 1)) {  // the next block's prompt rows replace the computed ones (they get no gradient: splice_bwd)
        e4.splice_prompt = prompts + static_cast<size_t>(i + 1) * t.n_ctx * d;
        e4.splice_row0 = t.row0; e4.splice_n = t.n_ctx; e4.splice_L = t.L;
      }
    } else {
      e4.mode = EPI_RESID_F32;
    }
    CKP(h, st, PC_GEMM_PROJ, 2.0 * Md * 4 * dd * dd, 2 * (Md * 4 * dd + 4 * dd * dd) + (fused ? 10 : 8) * Md * dd,
        gemm_bf16_tn(t.g_buf, 4 * d, w.w_pr, 4 * d, e4, M, d, 4 * d, st, &t.gws));
  }
  t.fwd_done = true;
  t.fwd_fused = fused;
  t.fwd_pruned = pruned;
  return 0;
}

// x (and the row stride) the feature head reads its one row per sequence from
const float* tower_head_input(const Tower& t, const int** rows, int* L) {
  if (t.fwd_pruned) { *rows = nullptr; *L = 1; return t.pr_xo; }
  *rows = t.sel_rows; *L = t.L;
  return t.x_in[t.layers];
}

// Precondition: t.dx / t.dx_bf16 hold the gradient w.r.t. the tower output x_in[layers].
//
// Fused formulation: the LayerNorm dgrad  dx += rstd (g - mean(g) - xhat mean(g xhat)),  g = dy gamma, runs in the
// epilogue of the dgrad GEMM that produces g (through the gamma-folded weight).  Its two row means do not need g:
//   mean(g) = (1/d) dout . colsum,   mean(g xhat) = (1/d) dout . (y - b')      (dout = the GEMM's A operand, y = the
// saved output of the forward LN-GEMM), so the PRODUCER of dout (GELU' epilogue, attention backward) emits them.
int tower_backward(mudpt_handle* h, Tower& t, float* d_prompts, int first_splice_layer, cudaStream_t st, bool need_dx0 = false) {
  const int M = t.S * t.L, d = t.d;
  const double Md = static_cast<double>(M), dd = d;
  const double attn_fl = 2.5 * 4.0 * t.S * t.H * static_cast<double>(t.L) * t.L * 64.0;  // 2.5x forward
  const bool fused = t.fwd_fused && h->ln_bwd_fused;  // (the fused forward also keeps the fp32 LN inputs the kernel needs)
  // stand-alone LayerNorm backward after a fused forward: the bf16 row + its exact statistics instead of the fp32 row
  // (14 instead of 16 B/element on an HBM-bound kernel; xhat from bf16(x) -- the operand precision of every GEMM here)
  const bool xstats = t.fwd_fused && !fused && h->ln_bwd_bf16x;
  const int parts = d / 64;
  const int dots_mlp = (4 * d + gemm_dots_span(4 * d) - 1) / gemm_dots_span(4 * d);
  // bf16 gradient stream (see mudpt_handle::grad_bf16): the stand-alone LayerNorm backward reads the residual gradient
  // from t.dx_bf16 and writes it back there; t.dx (fp32) gets the rows of the deep-prompt window only.
  const bool gb = h->grad_bf16 && !fused && !need_dx0;
  const double ln_bwd_bytes = (gb ? 6.0 : 12.0) + (xstats ? 2.0 : 4.0);  // dy 2, x, residual in, gradient out (+ its bf16 copy)
  bool f32_whole = true;  // t.dx holds the whole stream in fp32 (head_backward / the pruned last block write both copies)
  auto ln_bwd_main = [&](const void* xrow, const float2* stats, const float* gamma, bool has_resid, bool whole_f32_out) -> const char* {
    if (!gb)
      return layernorm_bwd(t.a_buf, true, xrow, stats, gamma, has_resid ? t.dx : nullptr, t.dx, t.dx_bf16, M, d, kLnEps, st);
    const bool rb = has_resid && !f32_whole;
    const void* resid = !has_resid ? nullptr : (rb ? static_cast<const void*>(t.dx_bf16) : static_cast<const void*>(t.dx));
    const char* e = layernorm_bwd_stream(t.a_buf, true, xrow, stats, gamma, resid, rb, t.dx, t.dx_bf16, M, d, kLnEps, t.L, t.row0,
                                         whole_f32_out ? -1 : t.n_ctx, st);
    f32_whole = whole_f32_out;
    return e;
  };
  for (int i = t.layers - 1; i >= 0; --i) {
    const Layer& w = t.lw[i];
    const bool tail_pruned = t.fwd_pruned && i == t.layers - 1;
    if (tail_pruned) {
      // Precondition here: t.pr_dx / t.pr_dxb hold the gradient of the S consumed rows of the block output.
      const int S = t.S;
      const double Sd = S;
      GemmEpilogue p1;
      p1.out0 = t.pr_dh; p1.aux = t.pr_h; p1.ldc = 4 * d;
      if (fused) { p1.mode = EPI_GELU_BWD_DOTS; p1.sb = w.sb_fc; p1.dots_out = t.pr_dots; }
      else p1.mode = EPI_GELU_BWD;
      CKP(h, st, PC_GEMM, 2.0 * Sd * 4 * dd * dd, 2 * (Sd * dd + 4 * dd * dd + 2 * Sd * 4 * dd), gemm_bf16_tn(t.pr_dxb, d, w.w_pr_t, d, p1, S, 4 * d, d, st, &t.gws));
      GemmEpilogue p2;
      p2.ldc = d;
      if (fused) {
        p2.mode = EPI_LN_BWD; p2.out0 = t.pr_dx; p2.resid = t.pr_dx; p2.out2 = t.pr_dxb; p2.x2 = t.pr_xmb;
        p2.ln_stats = t.pr_stm; p2.ln_parts = parts; p2.ln_width = d; p2.ln_eps = kLnEps; p2.dots = t.pr_dots; p2.dot_parts = dots_mlp;
        CKP(h, st, PC_GEMM, 2.0 * Sd * 4 * dd * dd, 2 * (Sd * 4 * dd + 4 * dd * dd) + 12 * Sd * dd, gemm_bf16_tn(t.pr_dh, 4 * d, w.w_fc_t_ln, 4 * d, p2, S, d, 4 * d, st, &t.gws));
      } else {
        p2.mode = EPI_BF16; p2.out0 = t.pr_a;
        CKP(h, st, PC_GEMM, 2.0 * Sd * 4 * dd * dd, 2 * (Sd * 4 * dd + 4 * dd * dd) + 2 * Sd * dd, gemm_bf16_tn(t.pr_dh, 4 * d, w.w_fc_t, 4 * d, p2, S, d, 4 * d, st, &t.gws));
        CKP(h, st, PC_LN_BWD, 0, Sd * dd * 16, layernorm_bwd(t.pr_a, true, t.pr_xm, nullptr, w.ln2_g, t.pr_dx, t.pr_dx, t.pr_dxb, S, d, kLnEps, st));
      }
      GemmEpilogue p3;
      p3.mode = EPI_BF16; p3.out0 = t.pr_do; p3.ldc = d;
      CKP(h, st, PC_GEMM, 2.0 * Sd * dd * dd, 2 * (2 * Sd * dd + dd * dd), gemm_bf16_tn(t.pr_dxb, d, w.w_out_t, d, p3, S, d, d, st, &t.gws));
      // dO of the whole block: zero but for the S consumed rows
      CUDA_OK(h, cudaMemsetAsync(t.do_buf, 0, static_cast<size_t>(M) * d * sizeof(bf16), st));
      CKP(h, st, PC_SPLICE, 0, Sd * dd * 4, scatter_rows_bf16(t.pr_do, t.sel_rows, S, t.L, t.do_buf, d, st));
    } else {
    // MLP branch: dg = dx W_pr ; dh = dg * GELU'(h) ; dm = dh W_fc ; dx += LN2_bwd(dm)
    GemmEpilogue e1;
    e1.out0 = t.dh_buf; e1.aux = t.h[i]; e1.ldc = 4 * d;
    if (fused) { e1.mode = EPI_GELU_BWD_DOTS; e1.sb = w.sb_fc; e1.dots_out = t.dots; }
    else e1.mode = EPI_GELU_BWD;
    CKP(h, st, PC_GEMM_DPROJ, 2.0 * Md * 4 * dd * dd, 2 * (Md * dd + 4 * dd * dd + 2 * Md * 4 * dd),
        gemm_bf16_tn(t.dx_bf16, d, w.w_pr_t, d, e1, M, 4 * d, d, st, &t.gws));
    GemmEpilogue e2;
    e2.ldc = d;
    if (fused) {
      e2.mode = EPI_LN_BWD; e2.out0 = t.dx; e2.resid = t.dx; e2.out2 = t.dx_bf16; e2.x2 = t.xb_mid[i];
      e2.ln_stats = t.st_mid[i]; e2.ln_parts = parts; e2.ln_width = d; e2.ln_eps = kLnEps; e2.dots = t.dots; e2.dot_parts = dots_mlp;
      CKP(h, st, PC_GEMM_DFC, 2.0 * Md * 4 * dd * dd, 2 * (Md * 4 * dd + 4 * dd * dd) + 12 * Md * dd,
          gemm_bf16_tn(t.dh_buf, 4 * d, w.w_fc_t_ln, 4 * d, e2, M, d, 4 * d, st, &t.gws));
    } else {
      e2.mode = EPI_BF16; e2.out0 = t.a_buf;  // a_buf: forward transient, free during the backward
      CKP(h, st, PC_GEMM_DFC, 2.0 * Md * 4 * dd * dd, 2 * (Md * 4 * dd + 4 * dd * dd) + 2 * Md * dd,
          gemm_bf16_tn(t.dh_buf, 4 * d, w.w_fc_t, 4 * d, e2, M, d, 4 * d, st, &t.gws));
      CKP(h, st, PC_LN_BWD, 0, Md * dd * ln_bwd_bytes,
          ln_bwd_main(xstats ? static_cast<const void*>(t.xb_mid[i]) : t.x_mid[i], xstats ? t.st_mid[i] : nullptr, w.ln2_g, true, false));
    }
    // attention branch: dO = dx W_out ; (dQ,dK,dV) ; da = dQKV W_in ; dx += LN1_bwd(da)
    GemmEpilogue e3;
    e3.mode = EPI_BF16; e3.out0 = t.do_buf; e3.ldc = d;
    CKP(h, st, PC_GEMM_DOUT, 2.0 * Md * dd * dd, 2 * (2 * Md * dd + dd * dd), gemm_bf16_tn(t.dx_bf16, d, w.w_out_t, d, e3, M, d, d, st, &t.gws));
    }
    CKP(h, st, PC_ATTN_BWD, attn_fl, Md * dd * 2 * 8,
        attention_bwd(t.qkv[i], t.o[i], t.do_buf, t.lse[i], t.dsum, t.dqkv_buf, t.S, t.L, t.H, d, t.causal, st,
                      fused ? w.sb_in : nullptr, fused ? t.dots : nullptr));
    // exact work skipping below block 0: only the prompt rows of its input gradient are ever read (by the splice backward)
    const bool head_pruned = i == 0 && h->prune && !fused && !tail_pruned && !need_dx0 && first_splice_layer == 0 && t.depth > 0 &&
                             t.n_ctx > 0 && t.S * t.n_ctx < M;
    if (head_pruned) {
      const int R = t.S * t.n_ctx;
      const double Rd = R;
      CKP(h, st, PC_SPLICE, 0, Rd * dd * 12, gather_rows(nullptr, t.dqkv_buf, t.p0_rows, R, 0, nullptr, t.p0_dqkv, 3 * d, st));
      GemmEpilogue q4;
      q4.mode = EPI_BF16; q4.out0 = t.p0_a; q4.ldc = d;
      CKP(h, st, PC_GEMM, 2.0 * Rd * 3 * dd * dd, 2 * (Rd * 3 * dd + 3 * dd * dd) + 2 * Rd * dd,
          gemm_bf16_tn(t.p0_dqkv, 3 * d, w.w_in_t, 3 * d, q4, R, d, 3 * d, st, &t.gws));
      CKP(h, st, PC_SPLICE, 0, Rd * dd * 8, gather_rows(t.x_in[0], nullptr, t.p0_rows, R, 0, t.p0_x, nullptr, d, st));
      CKP(h, st, PC_SPLICE, 0, Rd * dd * 8, gather_rows(t.dx, nullptr, t.p0_rows, R, 0, t.p0_dx, nullptr, d, st));
      CKP(h, st, PC_LN_BWD, 0, Rd * dd * 16, layernorm_bwd(t.p0_a, true, t.p0_x, nullptr, w.ln1_g, t.p0_dx, t.p0_dx, t.p0_dxb, R, d, kLnEps, st));
      CKP(h, st, PC_SPLICE, 0, Rd * dd * 8, scatter_rows(t.p0_dx, t.p0_rows, R, 0, t.dx, nullptr, d, false, st));
      CKP(h, st, PC_SPLICE, 0, t.S * t.n_ctx * dd * 4, splice_bwd(t.dx, t.dx_bf16, d_prompts, t.splice_ws, t.S, t.L, t.row0, t.n_ctx, d, false, st));
      continue;
    }
    GemmEpilogue e4;
    e4.ldc = d;
    if (fused) {
      e4.mode = EPI_LN_BWD; e4.out0 = t.dx; e4.resid = tail_pruned ? nullptr : t.dx; e4.out2 = t.dx_bf16; e4.x2 = t.xb_in[i];
      e4.ln_stats = t.st_in[i]; e4.ln_parts = parts; e4.ln_width = d; e4.ln_eps = kLnEps; e4.dots = t.dots; e4.dot_parts = 3 * t.H;
      CKP(h, st, PC_GEMM_DQKV, 2.0 * Md * 3 * dd * dd, 2 * (Md * 3 * dd + 3 * dd * dd) + 12 * Md * dd,
          gemm_bf16_tn(t.dqkv_buf, 3 * d, w.w_in_t_ln, 3 * d, e4, M, d, 3 * d, st, &t.gws));
    } else {
      e4.mode = EPI_BF16; e4.out0 = t.a_buf;
      CKP(h, st, PC_GEMM_DQKV, 2.0 * Md * 3 * dd * dd, 2 * (Md * 3 * dd + 3 * dd * dd) + 2 * Md * dd,
          gemm_bf16_tn(t.dqkv_buf, 3 * d, w.w_in_t, 3 * d, e4, M, d, 3 * d, st, &t.gws));
      // (the pruned last block adds its S consumed rows into the fp32 stream right below: every fp32 row is written there)
      CKP(h, st, PC_LN_BWD, 0, Md * dd * (tail_pruned ? 8.0 + (xstats ? 2.0 : 4.0) : ln_bwd_bytes),
          ln_bwd_main(xstats ? static_cast<const void*>(t.xb_in[i]) : t.x_in[i], xstats ? t.st_in[i] : nullptr, w.ln1_g, !tail_pruned,
                      tail_pruned));
    }
    if (tail_pruned)  // the residual path of the S consumed rows (everything else of it is zero)
      CKP(h, st, PC_SPLICE, 0, t.S * dd * 14, scatter_rows(t.pr_dx, t.sel_rows, t.S, t.L, t.dx, t.dx_bf16, d, true, st));
    // splice backward: the inserted prompt rows collect the batch-summed gradient; the rows
    // they overwrote get none (clip/model.py:281-297)
    if (i < t.depth && i >= first_splice_layer && t.n_ctx > 0)
      CKP(h, st, PC_SPLICE, 0, t.S * t.n_ctx * dd * 4,
          splice_bwd(t.dx, t.dx_bf16, d_prompts + static_cast<size_t>(i) * t.n_ctx * d, t.splice_ws, t.S, t.L, t.row0, t.n_ctx, d, i > 0, st));
  }
  return 0;
}

// Gradient of the feature head into the tower output: the S consumed rows only (pruned: a dense [S, d] buffer;
// otherwise scattered into the zeroed full gradient).
int head_backward(mudpt_handle* h, Tower& t, const float* d_f, const float* gamma, const float* proj, cudaStream_t st) {
  const int e = h->cfg.embed_dim;
  if (t.fwd_pruned) {
    CK(h, feature_head_bwd(d_f, t.pr_xo, nullptr, gamma, proj, t.pr_dx, t.pr_dxb, t.head_ws, t.S, 1, t.d, e, kLnEps, st));
    return 0;
  }
  const size_t n = static_cast<size_t>(t.S) * t.L * t.d;
  CUDA_OK(h, cudaMemsetAsync(t.dx, 0, n * sizeof(float), st));
  CUDA_OK(h, cudaMemsetAsync(t.dx_bf16, 0, n * sizeof(bf16), st));
  CK(h, feature_head_bwd(d_f, t.x_in[t.layers], t.sel_rows, gamma, proj, t.dx, t.dx_bf16, t.head_ws, t.S, t.L, t.d, e, kLnEps, st));
  return 0;
}

bool tower_complete(const Tower& t) {
  for (const Layer& l : t.lw)
    if (l.have != LB_ALL) return false;
  return true;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int mudpt_abi_version(void) { return MUDPT_ABI_VERSION; }
const char* mudpt_global_last_error(void) { return g_err.c_str(); }
const char* mudpt_last_error(mudpt_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int mudpt_create(const mudpt_config* cfg, mudpt_handle** out) {
  if (!cfg || !out) return fail(nullptr, "mudpt_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, "mudpt_create: no CUDA device (there is no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "mudpt_create: bad device ordinal %d", cfg->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(nullptr, "mudpt_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  if (cfg->vision_width % 64 || cfg->transformer_width % 64 || cfg->transformer_width != cfg->transformer_heads * 64)
    return fail(nullptr, "mudpt_create: tower widths must be heads * 64");
  if (cfg->image_resolution % cfg->vision_patch_size) return fail(nullptr, "mudpt_create: resolution %% patch != 0");
  if (cfg->prompt_depth < 1 || cfg->n_ctx < 0) return fail(nullptr, "mudpt_create: PROMPT_DEPTH should be > 0");
  if (cfg->embed_dim != cfg->transformer_width)
    return fail(nullptr, "mudpt_create: embed_dim must equal transformer_width (clip/model.py:519, trainers/mudpt.py:175)");
  cudaSetDevice(cfg->device);
  mudpt_handle* h = new mudpt_handle();
  h->cfg = *cfg;
  const int np = (cfg->image_resolution / cfg->vision_patch_size) * (cfg->image_resolution / cfg->vision_patch_size);
  Tower& v = h->vis;
  v.d = cfg->vision_width; v.H = cfg->vision_width / 64; v.layers = cfg->vision_layers; v.n_ctx = cfg->n_ctx;
  v.L = np + 1 + cfg->n_ctx; v.row0 = np + 1; v.depth = cfg->prompt_depth; v.causal = false;
  v.lw.resize(v.layers);
  Tower& t = h->txt;
  t.d = cfg->transformer_width; t.H = cfg->transformer_heads; t.layers = cfg->transformer_layers; t.n_ctx = cfg->n_ctx;
  t.L = cfg->context_length; t.row0 = 1; t.depth = cfg->prompt_depth; t.causal = true;
  t.lw.resize(t.layers);
  h->Kp = (3 * cfg->vision_patch_size * cfg->vision_patch_size + 7) & ~7;
  h->launches_at_create = g_launch_counter.load();
  h->ln_fused = getenv("MUDPT_LN_FUSED") ? (atoi(getenv("MUDPT_LN_FUSED")) != 0 ? 1 : 0) : -1;
  h->ln_bwd_fused = env_flag("MUDPT_LN_BWD_FUSED", false);
  h->ln_bwd_bf16x = env_flag("MUDPT_LN_BWD_BF16X", true);
  h->prune = env_flag("MUDPT_PRUNE", true);
  h->grad_bf16 = env_flag("MUDPT_GRAD_BF16", true);
  *out = h;
  return 0;
}

void mudpt_destroy(mudpt_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (void* p : h->allocs) cudaFree(p);
  for (Tower* t : {&h->vis, &h->txt}) {
    for (void* p : t->ws_allocs) cudaFree(p);
    for (void* p : t->pr_allocs) cudaFree(p);
    for (void* p : t->p0_allocs) cudaFree(p);
  }
  for (ProfRec& r : h->prof.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (cudaEvent_t e : h->prof.pool) cudaEventDestroy(e);
  gemm_clear_tensor_map_cache();
  delete h;
}

int mudpt_set_option(mudpt_handle* h, const char* name, int32_t value) {
  if (!h || !name) return fail(h, "mudpt_set_option: null argument");
  if (!strcmp(name, "ln_fused")) h->ln_fused = value < 0 ? -1 : (value != 0 ? 1 : 0);
  else if (!strcmp(name, "ln_bwd_fused")) h->ln_bwd_fused = value != 0;
  else if (!strcmp(name, "prune")) h->prune = value != 0;
  else if (!strcmp(name, "grad_stream_bf16")) h->grad_bf16 = value != 0;
  else return fail(h, "mudpt_set_option: unknown option %s", name);
  h->vis.fwd_done = h->txt.fwd_done = false;  // saved activations belong to the previous formulation
  return 0;
}

int mudpt_set_weight(mudpt_handle* h, const char* name, const float* data, int64_t numel, void* stream) {
  if (!h || !name || !data) return fail(h, "mudpt_set_weight: null argument");
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const mudpt_config& c = h->cfg;
  const int dv = c.vision_width, dt = c.transformer_width, e = c.embed_dim;
  const int np = (c.image_resolution / c.vision_patch_size) * (c.image_resolution / c.vision_patch_size);
  int layer = -1, off = 0;
  if (sscanf(name, "visual.transformer.resblocks.%d.%n", &layer, &off) == 1 && off > 0)
    return set_block_weight(h, h->vis, layer, name + off, data, numel, name, st);
  if (sscanf(name, "transformer.resblocks.%d.%n", &layer, &off) == 1 && off > 0)
    return set_block_weight(h, h->txt, layer, name + off, data, numel, name, st);
  int rc = 1;
  if (!strcmp(name, "visual.conv1.weight")) {
    const int k = 3 * c.vision_patch_size * c.vision_patch_size;
    if (numel != static_cast<int64_t>(dv) * k) return fail(h, "weight %s: bad size", name);
    if (!h->w_conv) CUDA_OK(h, dev_alloc(h, &h->w_conv, static_cast<size_t>(dv) * h->Kp));
    CUDA_OK(h, cudaMemsetAsync(h->w_conv, 0, static_cast<size_t>(dv) * h->Kp * sizeof(bf16), st));
    if (h->Kp == k) {
      CK(h, cast_to_bf16(data, h->w_conv, static_cast<size_t>(numel), st));
    } else {  // padded rows (ViT-L/14: 588 -> 592)
      for (int r = 0; r < dv; ++r) CK(h, cast_to_bf16(data + static_cast<size_t>(r) * k, h->w_conv + static_cast<size_t>(r) * h->Kp, k, st));
    }
    h->stem_have |= SB_CONV; rc = 0;
  } else if (!strcmp(name, "visual.class_embedding")) { rc = store_f32(h, &h->cls, data, numel, dv, name, st); h->stem_have |= SB_CLS; }
  else if (!strcmp(name, "visual.positional_embedding")) { rc = store_f32(h, &h->pos_v, data, numel, static_cast<int64_t>(np + 1) * dv, name, st); h->stem_have |= SB_POS_V; }
  else if (!strcmp(name, "visual.ln_pre.weight")) { rc = store_f32(h, &h->ln_pre_g, data, numel, dv, name, st); h->stem_have |= SB_LNPRE_G; }
  else if (!strcmp(name, "visual.ln_pre.bias")) { rc = store_f32(h, &h->ln_pre_b, data, numel, dv, name, st); h->stem_have |= SB_LNPRE_B; }
  else if (!strcmp(name, "visual.ln_post.weight")) { rc = store_f32(h, &h->ln_post_g, data, numel, dv, name, st); h->stem_have |= SB_LNPOST_G; }
  else if (!strcmp(name, "visual.ln_post.bias")) { rc = store_f32(h, &h->ln_post_b, data, numel, dv, name, st); h->stem_have |= SB_LNPOST_B; }
  else if (!strcmp(name, "visual.proj")) { rc = store_f32(h, &h->proj_v, data, numel, static_cast<int64_t>(dv) * e, name, st); h->stem_have |= SB_PROJ_V; }
  else if (!strcmp(name, "positional_embedding")) { rc = store_f32(h, &h->pos_t, data, numel, static_cast<int64_t>(c.context_length) * dt, name, st); h->stem_have |= SB_POS_T; }
  else if (!strcmp(name, "ln_final.weight")) { rc = store_f32(h, &h->ln_final_g, data, numel, dt, name, st); h->stem_have |= SB_LNF_G; }
  else if (!strcmp(name, "ln_final.bias")) { rc = store_f32(h, &h->ln_final_b, data, numel, dt, name, st); h->stem_have |= SB_LNF_B; }
  else if (!strcmp(name, "text_projection")) { rc = store_f32(h, &h->proj_t, data, numel, static_cast<int64_t>(dt) * e, name, st); h->stem_have |= SB_PROJ_T; }
  else if (!strcmp(name, "logit_scale")) {
    if (numel != 1) return fail(h, "weight logit_scale: expected 1 element");
    float v = 0.f;
    CUDA_OK(h, cudaMemcpyAsync(&v, data, sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_OK(h, cudaStreamSynchronize(st));
    h->logit_scale_exp = expf(v);  // logit_scale.exp(), trainers/mudpt.py:181 (frozen)
    h->stem_have |= SB_SCALE; rc = 0;
  }
  return rc;
}

int mudpt_weights_complete(mudpt_handle* h) {
  if (!h) return fail(h, "null handle");
  if (h->stem_have != SB_ALL) return fail(h, "stem/head weights missing (mask 0x%x of 0x%x)", h->stem_have, SB_ALL);
  if (!tower_complete(h->vis)) return fail(h, "vision tower weights incomplete");
  if (!tower_complete(h->txt)) return fail(h, "text tower weights incomplete");
  return 0;
}

int mudpt_vision_forward(mudpt_handle* h, const float* images, int32_t B, const float* prompts, float* f_img, void* stream) {
  if (!h || !images || !f_img || (!prompts && h->cfg.n_ctx > 0)) return fail(h, "mudpt_vision_forward: null argument");
  if (B <= 0) return fail(h, "mudpt_vision_forward: empty batch");
  if (mudpt_weights_complete(h)) return -1;
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const mudpt_config& c = h->cfg;
  Tower& t = h->vis;
  const int np = t.row0 - 1;
  if (ensure_tower(h, t, B, t.row0 + t.n_ctx)) return -1;
  const size_t need = static_cast<size_t>(B) * np * h->Kp;
  if (need > h->patches_cap) {
    CUDA_OK(h, dev_alloc(h, &h->patches, need));
    CUDA_OK(h, cudaMemsetAsync(h->patches, 0, need * sizeof(bf16), st));
    h->patches_cap = need;
  }
  // conv1 as a GEMM over extracted patches, scattered to token rows 1..np with +pos (clip/model.py:527-531)
  CK(h, im2col_bf16(images, h->patches, B, c.image_resolution, c.vision_patch_size, h->Kp, st));
  GemmEpilogue ep;
  ep.mode = EPI_PATCH; ep.out0 = t.x_in[0]; ep.resid = h->pos_v; ep.ldc = t.d; ep.patch_np = np; ep.patch_L = t.L;
  CK(h, gemm_bf16_tn(h->patches, h->Kp, h->w_conv, h->Kp, ep, B * np, t.d, h->Kp, st));
  CK(h, write_cls_rows(t.x_in[0], h->cls, h->pos_v, B, t.L, t.d, st));
  // ln_pre in place (:541); the prompt rows are overwritten by the layer-0 splice with
  // prompts[0] = ln_pre(visual_ctx + shared_ctx), which is identical for every image
  CK(h, layernorm_fwd(t.x_in[0], h->ln_pre_g, h->ln_pre_b, t.x_in[0], false, B * t.L, t.d, kLnEps, st));
  if (h->ln_fused == 1 || (h->ln_fused < 0 && B * t.L >= 32768))  // (the rule of tower_forward)
    CK(h, rowstats(t.x_in[0], t.xb_in[0], t.st_in[0], B * t.L, t.d, st));  // tower input as fp32 + bf16 + statistics
  t.sel_rows = nullptr;  // CLS row (clip/model.py:548)
  if (tower_forward(h, t, prompts, 0, st)) return -1;
  const int* rows; int Lh;
  const float* xh = tower_head_input(t, &rows, &Lh);
  CK(h, feature_head_fwd(xh, rows, h->ln_post_g, h->ln_post_b, h->proj_v, f_img, t.head_ws, B, Lh, t.d, c.embed_dim, kLnEps, st));
  return 0;
}

int mudpt_vision_backward(mudpt_handle* h, const float* d_f_img, float* d_prompts, void* stream) {
  if (!h || !d_f_img || !d_prompts) return fail(h, "mudpt_vision_backward: null argument");
  Tower& t = h->vis;
  if (!t.fwd_done) return fail(h, "mudpt_vision_backward: no forward pass to differentiate");
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (head_backward(h, t, d_f_img, h->ln_post_g, h->proj_v, st)) return -1;
  return tower_backward(h, t, d_prompts, 0, st);
}

int mudpt_text_set_classes(mudpt_handle* h, const float* embeddings, int32_t C, int32_t src_len, int32_t seq_len,
                           const int32_t* eot_host, void* stream) {
  if (!h || !embeddings || !eot_host) return fail(h, "mudpt_text_set_classes: null argument");
  const mudpt_config& c = h->cfg;
  if (C <= 0 || seq_len <= 0 || seq_len > src_len || src_len > c.context_length)
    return fail(h, "mudpt_text_set_classes: bad lengths (C=%d src_len=%d seq_len=%d)", C, src_len, seq_len);
  if (seq_len < 1 + c.n_ctx) return fail(h, "mudpt_text_set_classes: seq_len shorter than SOT + n_ctx");
  for (int i = 0; i < C; ++i)
    if (eot_host[i] < 0 || eot_host[i] >= seq_len) return fail(h, "mudpt_text_set_classes: eot[%d]=%d outside seq_len %d", i, eot_host[i], seq_len);
  if (!(h->stem_have & SB_POS_T)) return fail(h, "mudpt_text_set_classes: positional_embedding not set");
  cudaSetDevice(c.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Tower& t = h->txt;
  if (ensure_tower(h, t, C, seq_len)) return -1;
  if (static_cast<size_t>(C) > h->eot_cap) {
    CUDA_OK(h, dev_alloc(h, &h->eot, static_cast<size_t>(C)));
    h->eot_cap = C;
  }
  CUDA_OK(h, cudaMemcpyAsync(h->eot, eot_host, C * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_OK(h, cudaStreamSynchronize(st));  // eot_host may be freed by the caller after return
  CK(h, add_positional(t.x_in[0], embeddings, h->pos_t, C, seq_len, src_len, t.d, st));
  CK(h, rowstats(t.x_in[0], t.xb_in[0], t.st_in[0], C * seq_len, t.d, st));  // (the prompt rows are rewritten every step)
  t.fwd_done = false;
  return 0;
}

int mudpt_text_forward(mudpt_handle* h, const float* prompts, int32_t splice_layer0, float* f_txt, void* stream) {
  if (!h || !f_txt || (!prompts && h->cfg.n_ctx > 0)) return fail(h, "mudpt_text_forward: null argument");
  if (mudpt_weights_complete(h)) return -1;
  Tower& t = h->txt;
  if (t.cap_rows == 0 || t.S <= 0) return fail(h, "mudpt_text_forward: call mudpt_text_set_classes first");
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  t.sel_rows = h->eot;  // EOT row (trainers/mudpt.py:154)
  if (tower_forward(h, t, prompts, splice_layer0 ? 0 : 1, st)) return -1;
  const int* rows; int Lh;
  const float* xh = tower_head_input(t, &rows, &Lh);
  CK(h, feature_head_fwd(xh, rows, h->ln_final_g, h->ln_final_b, h->proj_t, f_txt, t.head_ws, t.S, Lh, t.d, h->cfg.embed_dim, kLnEps, st));
  t.first_splice = splice_layer0 ? 0 : 1;  // the backward must mirror the forward's splice set
  return 0;
}

int mudpt_text_backward(mudpt_handle* h, const float* d_f_txt, float* d_prompts, float* d_x0, void* stream) {
  if (!h || !d_f_txt || (!d_prompts && h->cfg.n_ctx > 0)) return fail(h, "mudpt_text_backward: null argument");
  Tower& t = h->txt;
  if (!t.fwd_done) return fail(h, "mudpt_text_backward: no forward pass to differentiate");
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(t.S) * t.L * t.d;
  if (head_backward(h, t, d_f_txt, h->ln_final_g, h->proj_t, st)) return -1;
  if (tower_backward(h, t, d_prompts, t.first_splice, st, d_x0 != nullptr)) return -1;
  if (d_x0) CUDA_OK(h, cudaMemcpyAsync(d_x0, t.dx, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int mudpt_logits_head(mudpt_handle* h, const float* f_img, const float* f_txt, const int64_t* labels, int32_t B,
                      int32_t C, float inv_global_batch, float* logits, float* loss, float* d_f_img, float* d_f_txt,
                      void* stream) {
  if (!h || !f_img || !f_txt || !logits) return fail(h, "mudpt_logits_head: null argument");
  if (labels && !loss) return fail(h, "mudpt_logits_head: loss output required with labels");
  if (!(h->stem_have & SB_SCALE)) return fail(h, "mudpt_logits_head: logit_scale not set");
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t need = logits_head_workspace_floats(B, C, h->cfg.embed_dim);
  if (need > h->head_ws_cap) {
    CUDA_OK(h, dev_alloc(h, &h->head_ws, need));
    h->head_ws_cap = need;
  }
  CK(h, logits_head(f_img, f_txt, reinterpret_cast<const long long*>(labels), h->logit_scale_exp, B, C, h->cfg.embed_dim,
                    inv_global_batch, h->head_ws, logits, loss, d_f_img, d_f_txt, st));
  return 0;
}

int mudpt_logits_backward(mudpt_handle* h, const float* f_img, const float* f_txt, const float* dlogits, int32_t B,
                          int32_t C, float* d_f_img, float* d_f_txt, void* stream) {
  if (!h || !f_img || !f_txt || !dlogits || !d_f_img || !d_f_txt) return fail(h, "mudpt_logits_backward: null argument");
  cudaSetDevice(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t need = logits_head_workspace_floats(B, C, h->cfg.embed_dim);
  if (need > h->head_ws_cap) {
    CUDA_OK(h, dev_alloc(h, &h->head_ws, need));
    h->head_ws_cap = need;
  }
  CK(h, logits_head_bwd(f_img, f_txt, dlogits, h->logit_scale_exp, B, C, h->cfg.embed_dim, h->head_ws, d_f_img, d_f_txt, st));
  return 0;
}

// ---------------------------------------------------------------------------- unit pieces
#define CKG(expr)                                       \
  do {                                                  \
    const char* _e = (expr);                            \
    if (_e) return fail(nullptr, "%s", _e);             \
  } while (0)

int mudpt_layernorm_forward(const float* x, const float* gamma, const float* beta, void* out, int32_t out_bf16,
                            int32_t rows, int32_t width, void* stream) {
  CKG(layernorm_fwd(x, gamma, beta, out, out_bf16 != 0, rows, width, kLnEps, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_layernorm_backward(const float* dy, const float* x, const float* gamma, const float* resid, float* dx,
                             uint16_t* dx_bf16, int32_t rows, int32_t width, void* stream) {
  CKG(layernorm_bwd(dy, false, x, nullptr, gamma, resid, dx, reinterpret_cast<bf16*>(dx_bf16), rows, width, kLnEps, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_layernorm_backward_stream(const uint16_t* dy, const void* x, const float* x_stats, const float* gamma, const void* resid,
                                    int32_t resid_bf16, float* dx, uint16_t* dx_bf16, int32_t rows, int32_t width, int32_t win_L,
                                    int32_t win_row0, int32_t win_n, void* stream) {
  CKG(layernorm_bwd_stream(dy, true, x, reinterpret_cast<const float2*>(x_stats), gamma, resid, resid_bf16 != 0, dx,
                           reinterpret_cast<bf16*>(dx_bf16), rows, width, kLnEps, win_L, win_row0, win_n, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_splice_forward(float* x, const float* prompt, int32_t S, int32_t L, int32_t row0, int32_t n, int32_t width, void* stream) {
  CKG(splice_fwd(x, prompt, S, L, row0, n, width, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_splice_backward(float* dx, uint16_t* dx_bf16, float* d_prompt, int32_t S, int32_t L, int32_t row0, int32_t n,
                          int32_t width, int32_t zero_rows, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = nullptr;  // unit entry point: stream-ordered scratch (the towers own a persistent one)
  if (cudaMallocAsync(reinterpret_cast<void**>(&ws), splice_bwd_workspace_floats(n, width) * sizeof(float), st) != cudaSuccess)
    return fail(nullptr, "mudpt_splice_backward: scratch allocation failed");
  cudaMemsetAsync(ws, 0, splice_bwd_workspace_floats(n, width) * sizeof(float), st);
  const char* e = splice_bwd(dx, reinterpret_cast<bf16*>(dx_bf16), d_prompt, ws, S, L, row0, n, width, zero_rows != 0, st);
  cudaFreeAsync(ws, st);
  if (e) return fail(nullptr, "%s", e);
  return 0;
}
int mudpt_attention_forward(const uint16_t* qkv, uint16_t* o, float* lse2, int32_t S, int32_t L, int32_t H, int32_t causal, void* stream) {
  CKG(attention_fwd(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(o), lse2, S, L, H, H * 64, causal != 0, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_attention_backward(const uint16_t* qkv, const uint16_t* o, const uint16_t* d_o, const float* lse2, float* dsum,
                             uint16_t* dqkv, int32_t S, int32_t L, int32_t H, int32_t causal, void* stream) {
  CKG(attention_bwd(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o),
                    lse2, dsum, reinterpret_cast<bf16*>(dqkv), S, L, H, H * 64, causal != 0, static_cast<cudaStream_t>(stream)));
  return 0;
}
// stream-K scratch of the handle-less GEMM entry points (tests, one stream at a time)
static GemmWorkspace* unit_workspace() {
  static GemmWorkspace ws;
  if (!ws.partials) {
    void *p = nullptr, *f = nullptr;
    if (cudaMalloc(&p, gemm_workspace_partial_bytes()) != cudaSuccess || cudaMalloc(&f, gemm_workspace_flag_bytes()) != cudaSuccess)
      return nullptr;
    cudaMemset(f, 0, gemm_workspace_flag_bytes());
    ws.partials = static_cast<float*>(p);
    ws.flags = static_cast<unsigned*>(f);
  }
  return &ws;
}

int mudpt_gemm_fused(const uint16_t* A, const uint16_t* B, int32_t M, int32_t N, int32_t K, const mudpt_gemm_epilogue* e,
                     void* stream) {
  if (!A || !B || !e) return fail(nullptr, "mudpt_gemm_fused: null argument");
  GemmEpilogue ep;
  ep.mode = e->mode; ep.ldc = e->ldc; ep.out0 = e->out0; ep.out1 = e->out1; ep.out2 = e->out2; ep.bias = e->bias;
  ep.resid = e->resid; ep.aux = e->aux; ep.ln_stats = reinterpret_cast<const float2*>(e->ln_stats); ep.ln_parts = e->ln_parts;
  ep.ln_width = e->ln_width; ep.ln_eps = e->ln_eps; ep.dot_parts = e->dot_parts; ep.colsum = e->colsum;
  ep.stats_out = reinterpret_cast<float2*>(e->stats_out); ep.splice_prompt = e->splice_prompt; ep.splice_row0 = e->splice_row0;
  ep.splice_n = e->splice_n; ep.splice_L = e->splice_L > 0 ? e->splice_L : 1; ep.x2 = e->x2;
  ep.dots = reinterpret_cast<const float2*>(e->dots); ep.sb = reinterpret_cast<const float2*>(e->sb);
  ep.dots_out = reinterpret_cast<float2*>(e->dots_out);
  gemm_set_stream_k(e->stream_k);
  const char* err = gemm_bf16_tn(reinterpret_cast<const bf16*>(A), K, reinterpret_cast<const bf16*>(B), K, ep, M, N, K,
                                 static_cast<cudaStream_t>(stream), e->stream_k != 0 ? unit_workspace() : nullptr);
  gemm_set_stream_k(-2);  // back to the process default
  if (err) return fail(nullptr, "%s", err);
  return 0;
}
int32_t mudpt_gemm_dots_span(int32_t N) { return gemm_dots_span(N); }
int mudpt_rowstats(const float* x, uint16_t* xb, float* stats, int32_t rows, int32_t width, void* stream) {
  CKG(rowstats(x, reinterpret_cast<bf16*>(xb), reinterpret_cast<float2*>(stats), rows, width, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_fold_layernorm(const float* W, const float* gamma, const float* beta, const float* bias, uint16_t* Wl,
                         uint16_t* Wlt, float* bias_l, float* colsum, float* sb, int32_t N, int32_t K, void* stream) {
  CKG(fold_layernorm(W, gamma, beta, bias, reinterpret_cast<bf16*>(Wl), reinterpret_cast<bf16*>(Wlt), bias_l, colsum,
                     reinterpret_cast<float2*>(sb), N, K, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_attention_backward_dots(const uint16_t* qkv, const uint16_t* o, const uint16_t* d_o, const float* lse2,
                                  float* dsum, uint16_t* dqkv, int32_t S, int32_t L, int32_t H, int32_t causal,
                                  const float* ln_sb, float* ln_dots, void* stream) {
  CKG(attention_bwd(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o),
                    lse2, dsum, reinterpret_cast<bf16*>(dqkv), S, L, H, H * 64, causal != 0, static_cast<cudaStream_t>(stream),
                    reinterpret_cast<const float2*>(ln_sb), reinterpret_cast<float2*>(ln_dots)));
  return 0;
}
int mudpt_gemm_bf16(const uint16_t* A, const uint16_t* B, int32_t M, int32_t N, int32_t K, int32_t mode, void* out0,
                    void* out1, const float* bias, const float* resid, const void* aux, int32_t ldc, int32_t patch_np,
                    int32_t patch_L, void* stream) {
  GemmEpilogue ep;
  ep.mode = mode; ep.out0 = out0; ep.out1 = out1; ep.bias = bias; ep.resid = resid; ep.aux = aux; ep.ldc = ldc;
  ep.patch_np = patch_np > 0 ? patch_np : 1; ep.patch_L = patch_L > 0 ? patch_L : 1;
  CKG(gemm_bf16_tn(reinterpret_cast<const bf16*>(A), K, reinterpret_cast<const bf16*>(B), K, ep, M, N, K, static_cast<cudaStream_t>(stream),
                   unit_workspace()));
  return 0;
}
int mudpt_set_attention_tc(int32_t mode) {
  attention_tc_set_mode(mode);
  return 0;
}
int mudpt_im2col(const float* images, uint16_t* patches, int32_t B, int32_t R, int32_t patch, int32_t ld, void* stream) {
  CKG(im2col_bf16(images, reinterpret_cast<bf16*>(patches), B, R, patch, ld, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_peer_all_gather_rows(const void* peers_dev, int32_t world, int32_t n_total, int32_t width, float* out, void* stream) {
  if (!peers_dev || !out) return fail(nullptr, "mudpt_peer_all_gather_rows: null argument");
  CKG(peer_gather_rows(reinterpret_cast<const float* const*>(peers_dev), world, n_total, width, out, static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_peer_reduce_scatter_rows(const void* peers_dev, int32_t world, int32_t rank, int32_t n_total, int32_t width, float* out,
                                   void* stream) {
  if (!peers_dev || !out) return fail(nullptr, "mudpt_peer_reduce_scatter_rows: null argument");
  CKG(peer_reduce_scatter_rows(reinterpret_cast<const float* const*>(peers_dev), world, rank, n_total, width, out,
                               static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_cast_bf16(const float* in, uint16_t* out, int64_t numel, void* stream) {
  CKG(cast_to_bf16(in, reinterpret_cast<bf16*>(out), static_cast<size_t>(numel), static_cast<cudaStream_t>(stream)));
  return 0;
}

static PromptArgs to_prompt_args(const mudpt_prompt_args* p) {
  PromptArgs a;
  a.n = p->n; a.depth = p->depth; a.dt = p->dt; a.dv = p->dv; a.eps = p->eps;
  a.ctx = p->ctx; a.deep = p->deep; a.We = p->We; a.be = p->be; a.Wd = p->Wd; a.bd = p->bd; a.vctx = p->vctx; a.vdeep = p->vdeep;
  a.Wv = p->Wv; a.bv = p->bv; a.ln_g = p->ln_g; a.ln_b = p->ln_b; a.pos = p->pos; a.P_v = p->P_v; a.P_t = p->P_t; a.ln_in = p->ln_in;
  a.dP_v = p->dP_v; a.dP_t = p->dP_t; a.u = p->u; a.d_ctx = p->d_ctx; a.d_deep = p->d_deep; a.d_We = p->d_We; a.d_be = p->d_be;
  a.d_Wd = p->d_Wd; a.d_bd = p->d_bd; a.d_vctx = p->d_vctx; a.d_vdeep = p->d_vdeep; a.d_Wv = p->d_Wv; a.d_bv = p->d_bv;
  return a;
}
int mudpt_prompt_forward(const mudpt_prompt_args* args, void* stream) {
  if (!args) return fail(nullptr, "mudpt_prompt_forward: null argument");
  CKG(prompt_forward(to_prompt_args(args), static_cast<cudaStream_t>(stream)));
  return 0;
}
int mudpt_prompt_backward(const mudpt_prompt_args* args, void* stream) {
  if (!args) return fail(nullptr, "mudpt_prompt_backward: null argument");
  CKG(prompt_backward(to_prompt_args(args), static_cast<cudaStream_t>(stream)));
  return 0;
}

int mudpt_sgd_step(void* const* params, const void* const* grads, void* const* bufs, const int64_t* numel, int32_t n,
                   float lr, float momentum, float dampening, float weight_decay, int32_t nesterov, int32_t first_step,
                   void* stream) {
  if (!params || !grads || !numel || (momentum != 0.f && !bufs)) return fail(nullptr, "mudpt_sgd_step: null argument");
  CKG(sgd_step(params, grads, bufs, reinterpret_cast<const long long*>(numel), n, lr, momentum, dampening, weight_decay,
               nesterov != 0, first_step != 0, static_cast<cudaStream_t>(stream)));
  return 0;
}

int64_t mudpt_augment_workspace_bytes(const mudpt_image_desc* descs_host, int32_t n, int32_t out_h, int32_t out_w) {
  const char* e = nullptr;
  const long long b = augment_workspace_bytes(descs_host, n, out_h, out_w, &e);
  if (e) return fail(nullptr, "%s", e);
  return b;
}
int mudpt_augment_images(const mudpt_image_desc* descs, const mudpt_image_desc* descs_host, int32_t n, int32_t out_h,
                         int32_t out_w, const float* mean_host, const float* std_host, void* workspace,
                         int64_t workspace_bytes, float* out, void* stream) {
  CKG(augment_images(descs, descs_host, n, out_h, out_w, mean_host, std_host, workspace, workspace_bytes, out,
                     static_cast<cudaStream_t>(stream)));
  return 0;
}

int mudpt_debug_buffer(mudpt_handle* h, int32_t tower, const char* name, int32_t layer, void** ptr, int64_t* numel) {
  if (!h || !name || !ptr || !numel) return fail(h, "mudpt_debug_buffer: null argument");
  Tower& t = tower == MUDPT_TOWER_VISION ? h->vis : h->txt;
  if (t.cap_rows == 0) return fail(h, "mudpt_debug_buffer: tower has no workspace yet");
  const int64_t rows = static_cast<int64_t>(t.S) * t.L;
  if (!strcmp(name, "dx")) { *ptr = t.dx; *numel = rows * t.d; return 0; }
  if (!strcmp(name, "x_in")) {
    if (layer < 0 || layer > t.layers) return fail(h, "mudpt_debug_buffer: layer out of range");
    *ptr = t.x_in[layer]; *numel = rows * t.d; return 0;
  }
  if (layer < 0 || layer >= t.layers) return fail(h, "mudpt_debug_buffer: layer out of range");
  if (!strcmp(name, "x_mid")) { *ptr = t.x_mid[layer]; *numel = rows * t.d; return 0; }
  if (!strcmp(name, "qkv")) { *ptr = t.qkv[layer]; *numel = rows * 3 * t.d; return 0; }
  if (!strcmp(name, "o")) { *ptr = t.o[layer]; *numel = rows * t.d; return 0; }
  if (!strcmp(name, "h")) { *ptr = t.h[layer]; *numel = rows * 4 * t.d; return 0; }
  if (!strcmp(name, "lse")) { *ptr = t.lse[layer]; *numel = rows * t.H; return 0; }
  return fail(h, "mudpt_debug_buffer: unknown buffer %s", name);
}

int mudpt_profile_begin(mudpt_handle* h) {
  if (!h) return fail(h, "null handle");
  for (ProfRec& r : h->prof.recs) { h->prof.pool.push_back(r.e0); h->prof.pool.push_back(r.e1); }
  h->prof.recs.clear();
  h->prof.on = true;
  return 0;
}

// out[cat*4 + {0,1,2,3}] = {total ms, launches, algorithmic FLOPs, algorithmic bytes}; cat order:
// gemm (other), attn_fwd, attn_bwd, ln_fwd, ln_bwd, splice, head, stem, gemm_qkv, gemm_out, gemm_fc, gemm_proj,
// gemm_dproj, gemm_dfc, gemm_dout, gemm_dqkv.  Blocks until the recorded work is done.
// With peaks (TFLOP/s, GB/s) > 0 a fifth value per class: the sum over its launches of the time the launch would take at
// the bound that applies to it, max(FLOPs / peak, algorithmic bytes / hbm) -- per launch, so a class that mixes
// tensor-bound and HBM-bound launches is measured against the right one each time.
static int profile_collect(mudpt_handle* h, double* out_host, int stride, double peak_tflops, double hbm_gbs) {
  h->prof.on = false;
  for (int i = 0; i < PC_COUNT * stride; ++i) out_host[i] = 0.0;
  for (ProfRec& r : h->prof.recs) {
    CUDA_OK(h, cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CUDA_OK(h, cudaEventElapsedTime(&ms, r.e0, r.e1));
    out_host[r.cat * stride + 0] += ms;
    out_host[r.cat * stride + 1] += 1.0;
    out_host[r.cat * stride + 2] += r.flops;
    out_host[r.cat * stride + 3] += r.bytes;
    if (stride > 4 && peak_tflops > 0 && hbm_gbs > 0) {
      const double t_fl = r.flops / (peak_tflops * 1e12), t_by = r.bytes / (hbm_gbs * 1e9);
      out_host[r.cat * stride + 4] += 1e3 * (t_fl > t_by ? t_fl : t_by);
    }
    h->prof.pool.push_back(r.e0);
    h->prof.pool.push_back(r.e1);
  }
  h->prof.recs.clear();
  return 0;
}

int mudpt_profile_end(mudpt_handle* h, double* out_host, int32_t n_out) {
  if (!h || !out_host) return fail(h, "mudpt_profile_end: null argument");
  if (n_out < PC_COUNT * 4) return fail(h, "mudpt_profile_end: output too small");
  return profile_collect(h, out_host, 4, 0.0, 0.0);
}

int mudpt_profile_end_bound(mudpt_handle* h, double* out_host, int32_t n_out, double peak_tflops, double hbm_gbs) {
  if (!h || !out_host) return fail(h, "mudpt_profile_end_bound: null argument");
  if (n_out < PC_COUNT * 5) return fail(h, "mudpt_profile_end_bound: output too small");
  return profile_collect(h, out_host, 5, peak_tflops, hbm_gbs);
}

int64_t mudpt_launch_count(mudpt_handle* h) { return h ? g_launch_counter.load() - h->launches_at_create : 0; }

#ifdef MUDPT_BRINGUP
int mudpt_bringup_simt_gemm(int on) { gemm_set_bringup_simt(on != 0); return 0; }
#endif

}  // extern "C"
