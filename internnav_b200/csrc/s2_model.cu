// System-2 executor (Qwen2.5-VL ViT + decoder prefill).  See s2_model.h.
#include "s2_model.h"

#include <math.h>

#include <algorithm>
#include <memory>

#include "s2_kernels.h"

namespace n1 {

namespace {

template <typename T>
T* upload(const std::vector<T>& v, cudaStream_t s) {
  T* d = nullptr;
  N1_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) N1_CUDA(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return d;
}

void linear(const Lin& L, const bf16* A, int lda, void* out, int ldo, int M, GemmEpilogue e, cudaStream_t s) {
  e.bias = L.b;
  gemm_bf16(A, lda, L.w, L.ldw, out, ldo, M, L.N, L.K, e, s);
}

// [rows_a; rows_b; ...] stacked into one K-major bf16 matrix (+ optional stacked fp32 bias)
Lin stack_lin(Arena& a, const WeightSource& ws, const std::vector<std::string>& names, bool bias, long cols,
              cudaStream_t s) {
  Lin L;
  long rows = 0;
  for (const std::string& n : names) rows += ws.get(n + ".weight").numel() / cols;
  L.N = (int)rows, L.K = (int)((cols + 7) & ~7L), L.ldw = L.K;
  N1_CHECK(L.N % 8 == 0, "stacked linear: N % 8");
  L.w = a.alloc_n<bf16>((size_t)rows * L.ldw);
  if (bias) L.b = a.alloc_n<float>(rows);
  long r0 = 0;
  for (const std::string& n : names) {
    const SrcTensor& w = ws.get(n + ".weight");
    const long r = w.numel() / cols;
    pack2d(w.data, w.dtype, cols, 0, r, cols, L.w + (size_t)r0 * L.ldw, 1, L.ldw, s);
    if (bias) {
      const SrcTensor& b = ws.get(n + ".bias");
      pack2d(b.data, b.dtype, 1, 0, r, 1, L.b + r0, 0, 1, s);
    }
    r0 += r;
  }
  return L;
}

// SwiGLU pair: rows interleaved (gate_j, up_j), padded to inter_pad zero rows so the GEMM N is a multiple of 16
Lin gateup_lin(Arena& a, const WeightSource& ws, const std::string& p, bool bias, long inter, long inter_pad, long cols,
               cudaStream_t s) {
  Lin L;
  L.N = (int)(2 * inter_pad), L.K = (int)((cols + 7) & ~7L), L.ldw = L.K;
  L.w = a.alloc_n<bf16>((size_t)L.N * L.ldw);
  N1_CUDA(cudaMemsetAsync(L.w, 0, (size_t)L.N * L.ldw * sizeof(bf16), s));
  const SrcTensor& g = ws.get(p + "gate_proj.weight");
  const SrcTensor& u = ws.get(p + "up_proj.weight");
  N1_CHECK(g.numel() == inter * cols && u.numel() == inter * cols, p + "gate/up shape mismatch");
  N1_CHECK(g.dtype == u.dtype, p + "gate/up dtype mismatch");
  interleave_rows(g.data, u.data, g.dtype, inter, cols, L.w, 1, L.ldw, s);
  if (bias) {
    L.b = a.alloc_n<float>(L.N);
    N1_CUDA(cudaMemsetAsync(L.b, 0, L.N * sizeof(float), s));
    const SrcTensor& gb = ws.get(p + "gate_proj.bias");
    const SrcTensor& ub = ws.get(p + "up_proj.bias");
    interleave_rows(gb.data, ub.data, gb.dtype, inter, 1, L.b, 0, 1, s);
  }
  return L;
}

Lin plain_lin(Arena& a, const WeightSource& ws, const std::string& name, bool bias, long rows, long cols,
              cudaStream_t s) {
  Lin L;
  L.N = (int)rows, L.K = (int)((cols + 7) & ~7L);
  L.w = ws.mat(a, name + ".weight", 0, rows, cols, &L.ldw, s);
  if (bias) L.b = ws.f32(a, name + ".bias", s);
  N1_CHECK(L.N % 8 == 0, name + ": N % 8");
  return L;
}

}  // namespace

VitPlan::~VitPlan() {
  cudaFree(window_index), cudaFree(reverse_index), cudaFree(cu_window), cudaFree(cu_full), cudaFree(rope);
}
LlmPlan::~LlmPlan() { cudaFree(cu), cudaFree(kind), cudaFree(src), cudaFree(out_rows), cudaFree(rope); }

// ------------------------------------------------------------------------------------------------ load
void S2Model::load(const WeightSource& ws, const S2Dims& d, cudaStream_t s) {
  dims = d;
  N1_CHECK(d.v_hidden % d.v_heads == 0 && d.v_hidden / d.v_heads == 80, "vision head_dim must be 80");
  N1_CHECK(d.head_dim == 128 && d.hidden == d.heads * d.head_dim, "decoder head_dim must be 128, hidden = heads * 128");
  N1_CHECK(d.heads % d.kv_heads == 0, "heads % kv_heads");
  N1_CHECK(d.mrope[0] + d.mrope[1] + d.mrope[2] == d.head_dim / 2, "mrope_section must sum to head_dim / 2");
  N1_CHECK(d.v_hidden % 8 == 0 && d.hidden % 8 == 0 && d.v_out == d.hidden, "hidden sizes");
  const WeightSource v = ws.sub("visual.");
  const int Hv = d.v_hidden;
  patch_k_ = 3 * d.v_tpatch * d.v_patch * d.v_patch;
  v_patch_ = plain_lin(arena_, v, "patch_embed.proj", false, Hv, patch_k_, s);
  v_inter_pad_ = (d.v_inter + 7) & ~7;
  vblk_.resize(d.v_depth);
  for (int i = 0; i < d.v_depth; ++i) {
    const std::string p = "blocks." + std::to_string(i) + ".";
    VBlock& b = vblk_[i];
    b.n1 = v.f32(arena_, p + "norm1.weight", s);
    b.n2 = v.f32(arena_, p + "norm2.weight", s);
    b.qkv = plain_lin(arena_, v, p + "attn.qkv", true, 3 * Hv, Hv, s);
    b.proj = plain_lin(arena_, v, p + "attn.proj", true, Hv, Hv, s);
    b.gateup = gateup_lin(arena_, v, p + "mlp.", true, d.v_inter, v_inter_pad_, Hv, s);
    b.down = plain_lin(arena_, v, p + "mlp.down_proj", true, Hv, d.v_inter, s);
    N1_CHECK(b.down.K == v_inter_pad_, "vision down_proj K padding");
  }
  merger_ln_ = v.f32(arena_, "merger.ln_q.weight", s);
  const int unit = d.v_merge * d.v_merge;
  merger0_ = plain_lin(arena_, v, "merger.mlp.0", true, Hv * unit, Hv * unit, s);
  merger2_ = plain_lin(arena_, v, "merger.mlp.2", true, d.v_out, Hv * unit, s);

  const WeightSource m = ws.sub("model.");
  const int H = d.hidden;
  {
    int ld;
    embed_ = m.mat(arena_, "embed_tokens.weight", 0, d.vocab, H, &ld, s);
    latentq_ = m.mat(arena_, "latent_queries", 0, d.n_query, H, &ld, s);
  }
  inter_pad_ = (d.inter + 7) & ~7;
  lblk_.resize(d.layers);
  for (int i = 0; i < d.layers; ++i) {
    const std::string p = "layers." + std::to_string(i) + ".";
    LBlock& b = lblk_[i];
    b.n1 = m.f32(arena_, p + "input_layernorm.weight", s);
    b.n2 = m.f32(arena_, p + "post_attention_layernorm.weight", s);
    b.qkv = stack_lin(arena_, m, {p + "self_attn.q_proj", p + "self_attn.k_proj", p + "self_attn.v_proj"}, true, H, s);
    N1_CHECK(b.qkv.N == (d.heads + 2 * d.kv_heads) * d.head_dim, "qkv rows");
    b.o = plain_lin(arena_, m, p + "self_attn.o_proj", false, H, H, s);
    b.gateup = gateup_lin(arena_, m, p + "mlp.", false, d.inter, inter_pad_, H, s);
    b.down = plain_lin(arena_, m, p + "mlp.down_proj", false, H, d.inter, s);
  }
  final_norm_ = m.f32(arena_, "norm.weight", s);
  N1_CUDA(cudaStreamSynchronize(s));
  loaded_ = true;
}

// ------------------------------------------------------------------------------------------------ plans
VitPlan* S2Model::make_vit_plan(const int32_t* grid, int n_img, cudaStream_t s) const {
  N1_CHECK(loaded_, "System-2 weights not loaded");
  std::unique_ptr<VitPlan> p(new VitPlan());
  vit_index(grid, n_img, dims.v_merge, dims.v_window / dims.v_merge / dims.v_patch, p->host);
  p->window_index = upload(p->host.window_index, s);
  p->reverse_index = upload(p->host.reverse_index, s);
  p->cu_window = upload(p->host.cu_window, s);
  p->cu_full = upload(p->host.cu_full, s);
  p->n_window = (int)p->host.cu_window.size() - 1;
  p->n_full = (int)p->host.cu_full.size() - 1;
  int* pos = upload(p->host.pos_hw, s);
  const int half = (dims.v_hidden / dims.v_heads) / 2;
  N1_CUDA(cudaMalloc(&p->rope, (size_t)p->host.n_patches * half * sizeof(float2)));
  vit_rope_table(pos, p->rope, p->host.n_patches, half, 10000.0f, s);
  N1_CUDA(cudaStreamSynchronize(s));
  cudaFree(pos);
  return p.release();
}

LlmPlan* S2Model::make_llm_plan(const int32_t* ids, const int32_t* lens, int B, const int32_t* grid, int n_img,
                                cudaStream_t s) const {
  N1_CHECK(loaded_, "System-2 weights not loaded");
  std::unique_ptr<LlmPlan> p(new LlmPlan());
  const int nq = dims.n_query, unit = dims.v_merge * dims.v_merge;
  p->B = B, p->n_query = nq;
  std::vector<int> kind, src, out_rows, pos_all;
  p->h_cu.push_back(0);
  int cursor = 0;
  long img_tok = 0, off = 0;
  std::vector<std::vector<int>> pos_seq(B);
  for (int b = 0; b < B; ++b) {
    const int len = lens[b];
    std::vector<int> seq(ids + off, ids + off + len);
    off += len;
    for (int q = 0; q < nq; ++q) seq.push_back(kTrajTokenId);  // internvla_n1.py L327
    const int L = (int)seq.size();
    int delta = 0;
    rope_index_one(seq.data(), L, grid, n_img, dims.v_merge, cursor, pos_seq[b], delta);
    p->h_delta.push_back(delta);
    for (int i = 0; i < L; ++i) {
      if (i >= len) {
        kind.push_back(2), src.push_back(i - len);  // latent_queries[q]
        out_rows.push_back((int)kind.size() - 1);
      } else if (seq[i] == kImageTokenId) {
        kind.push_back(1), src.push_back((int)img_tok++);  // image features in order (L332: text_embeds[image_idx] = ...)
      } else {
        N1_CHECK(seq[i] >= 0 && seq[i] < dims.vocab, "token id out of vocabulary");
        kind.push_back(0), src.push_back(seq[i]);
      }
    }
    p->h_cu.push_back(p->h_cu.back() + L);
    p->max_len = std::max(p->max_len, L);
  }
  p->tokens = p->h_cu.back();
  p->n_image_tokens = img_tok;
  long expect = 0;
  for (int i = 0; i < cursor; ++i) expect += (long)grid[i * 3] * grid[i * 3 + 1] * grid[i * 3 + 2] / unit;
  N1_CHECK(expect == img_tok, "Image features and image tokens do not match: tokens " + std::to_string(img_tok) +
                                  ", features " + std::to_string(expect));  // same check as internvla_n1.py L135-138
  // [3, tokens] layout over the packed batch
  p->h_pos3.assign((size_t)3 * p->tokens, 0);
  for (int b = 0; b < B; ++b) {
    const int L = p->h_cu[b + 1] - p->h_cu[b];
    for (int st = 0; st < 3; ++st)
      for (int i = 0; i < L; ++i) p->h_pos3[(size_t)st * p->tokens + p->h_cu[b] + i] = pos_seq[b][(size_t)st * L + i];
  }
  p->cu = upload(p->h_cu, s);
  p->kind = upload(kind, s);
  p->src = upload(src, s);
  p->out_rows = upload(out_rows, s);
  int* pos = upload(p->h_pos3, s);
  const int half = dims.head_dim / 2;
  N1_CUDA(cudaMalloc(&p->rope, (size_t)p->tokens * half * sizeof(float2)));
  mrope_table(pos, p->rope, p->tokens, half, dims.mrope[0], dims.mrope[1], dims.rope_theta, s);
  N1_CUDA(cudaStreamSynchronize(s));
  cudaFree(pos);
  return p.release();
}

// ------------------------------------------------------------------------------------------------ vision tower
size_t S2Model::vit_impl(Carver c, const VitPlan& p, const bf16* pixels, bf16* out, cudaStream_t s) const {
  const int Hv = dims.v_hidden, unit = dims.v_merge * dims.v_merge;
  const long N = p.host.n_patches, Nm = N / unit;
  bf16* xin = c.take<bf16>(N * patch_k_);
  bf16* x = c.take<bf16>(N * Hv);
  bf16* ln = c.take<bf16>(N * Hv);
  bf16* qkv = c.take<bf16>(N * 3 * Hv);
  bf16* att = c.take<bf16>(N * Hv);
  bf16* hid = c.take<bf16>(N * (long)std::max(v_inter_pad_, Hv));
  bf16* m2 = c.take<bf16>(Nm * dims.v_out);
  if (c.dry()) return c.used();

  gather_rows(pixels, p.window_index, xin, N, unit, patch_k_, s);  // hidden_states[window_index] on merge groups
  linear(v_patch_, xin, patch_k_, x, Hv, (int)N, GemmEpilogue(), s);
  for (int l = 0; l < dims.v_depth; ++l) {
    const VBlock& b = vblk_[l];
    layernorm(x, Hv, ln, Hv, b.n1, nullptr, (int)N, Hv, 1e-6f, 1, s);
    linear(b.qkv, ln, Hv, qkv, 3 * Hv, (int)N, GemmEpilogue(), s);
    apply_rope(qkv, 3 * Hv, p.rope, N, 2 * dims.v_heads, 80, s);  // q and k are adjacent column blocks
    bool full = false;
    for (int i = 0; i < dims.n_fullatt; ++i) full |= dims.fullatt[i] == l;
    AttnParams a = {};
    a.q = qkv, a.k = qkv + Hv, a.v = qkv + 2 * Hv, a.o = att;
    a.ldq = a.ldk = a.ldv = 3 * Hv, a.ldo = Hv;
    a.heads_q = a.heads_kv = dims.v_heads, a.hd = 80;
    a.batch = full ? p.n_full : p.n_window;
    a.cu_q = a.cu_k = full ? p.cu_full : p.cu_window;
    a.max_seq_q = full ? p.host.max_full : p.host.max_window;
    a.kv_div = 1, a.scale = 1.0f / sqrtf(80.f);
    attention(a, s);
    GemmEpilogue res;
    res.residual = x, res.ldr = Hv;
    linear(b.proj, att, Hv, x, Hv, (int)N, res, s);
    layernorm(x, Hv, ln, Hv, b.n2, nullptr, (int)N, Hv, 1e-6f, 1, s);
    GemmEpilogue sw;
    sw.act = ACT_SWIGLU;
    linear(b.gateup, ln, Hv, hid, v_inter_pad_, (int)N, sw, s);
    linear(b.down, hid, v_inter_pad_, x, Hv, (int)N, res, s);
  }
  // merger: RMSNorm over Hv, then groups of merge^2 patches concatenated (a pure view), MLP with exact GELU
  layernorm(x, Hv, ln, Hv, merger_ln_, nullptr, (int)N, Hv, 1e-6f, 1, s);
  GemmEpilogue ge;
  ge.act = ACT_GELU;
  linear(merger0_, ln, Hv * unit, hid, Hv * unit, (int)Nm, ge, s);
  linear(merger2_, hid, Hv * unit, m2, dims.v_out, (int)Nm, GemmEpilogue(), s);
  gather_rows(m2, p.reverse_index, out, Nm, 1, dims.v_out, s);  // merged[argsort(window_index)]
  return c.used();
}

size_t S2Model::ws_vit(const VitPlan& p) const { return vit_impl(Carver(nullptr, 0), p, nullptr, nullptr, nullptr); }
void S2Model::vit_forward(const VitPlan& p, void* ws, size_t ws_bytes, const bf16* pixels, bf16* out,
                          cudaStream_t s) const {
  N1_CHECK(loaded_ && ws, "vit_forward: not loaded / null workspace");
  if (ws_bytes < ws_vit(p)) throw Error(-7, "vit_forward: workspace too small");
  vit_impl(Carver(ws, ws_bytes), p, pixels, out, s);
}

// ------------------------------------------------------------------------------------------------ decoder prefill
size_t S2Model::llm_impl(Carver c, const LlmPlan& p, const bf16* image_feats, bf16* out, cudaStream_t s) const {
  const int H = dims.hidden, hd = dims.head_dim;
  const long T = p.tokens;
  const int qkv_n = (dims.heads + 2 * dims.kv_heads) * hd;
  bf16* x = c.take<bf16>(T * H);
  const long sel_rows = 3L * p.B * p.n_query;  // the last layer reuses `ln` for three [B * n_query, H] buffers
  bf16* ln = c.take<bf16>((T > sel_rows ? T : sel_rows) * H);
  bf16* qkv = c.take<bf16>(T * qkv_n);
  bf16* att = c.take<bf16>(T * H);
  bf16* hid = c.take<bf16>(T * (long)inter_pad_);
  if (c.dry()) return c.used();

  build_embeds(p.kind, p.src, embed_, image_feats, latentq_, x, T, H, s);
  for (int l = 0; l < dims.layers; ++l) {
    const LBlock& b = lblk_[l];
    layernorm(x, H, ln, H, b.n1, nullptr, (int)T, H, dims.rms_eps, 1, s);
    linear(b.qkv, ln, H, qkv, qkv_n, (int)T, GemmEpilogue(), s);
    apply_rope(qkv, qkv_n, p.rope, T, dims.heads + dims.kv_heads, hd, s);
    AttnParams a = {};
    a.q = qkv, a.k = qkv + (long)dims.heads * hd, a.v = qkv + (long)(dims.heads + dims.kv_heads) * hd, a.o = att;
    a.ldq = a.ldk = a.ldv = qkv_n, a.ldo = H;
    a.heads_q = dims.heads, a.heads_kv = dims.kv_heads, a.hd = hd;
    a.batch = p.B, a.cu_q = a.cu_k = p.cu, a.max_seq_q = p.max_len;
    a.kv_div = 1, a.causal = 1, a.scale = 1.0f / sqrtf((float)hd);
    attention(a, s);
    GemmEpilogue res;
    res.residual = x, res.ldr = H;
    if (l == dims.layers - 1) {
      // Only the n_query TRAJ rows of each sequence are read after the last layer (internvla_n1.py L345): every token
      // still contributes K/V to the attention above, but o_proj and the MLP run on those B * n_query rows alone.
      const int R = p.B * p.n_query;
      bf16* att_sel = ln;                      // [R, H] scratch (ln is free here)
      bf16* x_sel = ln + (long)R * H;          // [R, H]
      bf16* ln_sel = ln + 2L * R * H;          // [R, H]
      gather_rows(att, p.out_rows, att_sel, R, 1, H, s);
      gather_rows(x, p.out_rows, x_sel, R, 1, H, s);
      GemmEpilogue rs;
      rs.residual = x_sel, rs.ldr = H;
      linear(b.o, att_sel, H, x_sel, H, R, rs, s);
      layernorm(x_sel, H, ln_sel, H, b.n2, nullptr, R, H, dims.rms_eps, 1, s);
      GemmEpilogue sw;
      sw.act = ACT_SWIGLU;
      linear(b.gateup, ln_sel, H, hid, inter_pad_, R, sw, s);
      linear(b.down, hid, inter_pad_, x_sel, H, R, rs, s);
      // outputs.hidden_states[-1][:, -N_QUERY:, :] -- the last entry is post final-norm
      layernorm(x_sel, H, out, H, final_norm_, nullptr, R, H, dims.rms_eps, 1, s);
      break;
    }
    linear(b.o, att, H, x, H, (int)T, res, s);
    layernorm(x, H, ln, H, b.n2, nullptr, (int)T, H, dims.rms_eps, 1, s);
    GemmEpilogue sw;
    sw.act = ACT_SWIGLU;
    linear(b.gateup, ln, H, hid, inter_pad_, (int)T, sw, s);
    linear(b.down, hid, inter_pad_, x, H, (int)T, res, s);
  }
  return c.used();
}

size_t S2Model::ws_llm(const LlmPlan& p) const { return llm_impl(Carver(nullptr, 0), p, nullptr, nullptr, nullptr); }
void S2Model::llm_prefill(const LlmPlan& p, void* ws, size_t ws_bytes, const bf16* image_feats, bf16* out,
                          cudaStream_t s) const {
  N1_CHECK(loaded_ && ws, "llm_prefill: not loaded / null workspace");
  if (ws_bytes < ws_llm(p)) throw Error(-7, "llm_prefill: workspace too small");
  llm_impl(Carver(ws, ws_bytes), p, image_feats, out, s);
}

}  // namespace n1
