// System-2 executor (Qwen2.5-VL ViT + decoder prefill).  See s2_model.h.
#include "s2_model.h"

#include <math.h>
#include <stdlib.h>

#include <string.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <vector>

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
// ---- pooled plan storage: device blocks (and their pinned staging twins) are recycled across plans
namespace {
struct PlanBlock {
  void* dev = nullptr;
  void* host = nullptr;  // pinned staging of the integer arrays
  size_t bytes = 0;
  cudaEvent_t last_use = nullptr;  // recorded by the previous owner's consumers
};
std::mutex g_pool_mu;
std::vector<PlanBlock> g_pool;

PlanBlock pool_take(size_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); ++i)
      if (g_pool[i].bytes >= bytes) {
        PlanBlock b = g_pool[i];
        g_pool.erase(g_pool.begin() + i);
        return b;
      }
  }
  PlanBlock b;
  b.bytes = (bytes + (1 << 20) - 1) & ~size_t((1 << 20) - 1);  // 1 MiB granularity: similar prompts share a size class
  N1_CUDA(cudaMalloc(&b.dev, b.bytes));
  N1_CUDA(cudaMallocHost(&b.host, b.bytes));
  N1_CUDA(cudaEventCreateWithFlags(&b.last_use, cudaEventDisableTiming));
  return b;
}
void pool_give(const PlanBlock& b) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (g_pool.size() < 128) {
    g_pool.push_back(b);
    return;
  }
  cudaFree(b.dev), cudaFreeHost(b.host), cudaEventDestroy(b.last_use);
}
struct PlanHostTwin {  // keeps the pinned twin reachable from the plan without widening the public struct
  std::mutex mu;
  std::vector<std::pair<const void*, void*>> map;
} g_twin;
}  // namespace

void LlmPlan::wait_ready(cudaStream_t s) const {
  if (ready) N1_CUDA(cudaStreamWaitEvent(s, ready, 0));
}
void LlmPlan::mark_used(cudaStream_t s) const {
  if (last_use) N1_CUDA(cudaEventRecord(last_use, s));
}
LlmPlan::~LlmPlan() {
  if (ready) cudaEventDestroy(ready);
  if (block) {
    PlanBlock b;
    b.dev = block, b.bytes = block_bytes, b.last_use = last_use;
    {
      std::lock_guard<std::mutex> lk(g_twin.mu);
      for (size_t i = 0; i < g_twin.map.size(); ++i)
        if (g_twin.map[i].first == block) {
          b.host = g_twin.map[i].second;
          g_twin.map.erase(g_twin.map.begin() + i);
          break;
        }
    }
    pool_give(b);
  }
}

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
  // Qwen2.5-VL-7B does not tie the output projection to the embedding; generate_latents never reads it, generate() does.
  if (ws.has("lm_head.weight")) lm_head_ = plain_lin(arena_, ws, "lm_head", false, d.vocab, H, s);
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
                                cudaStream_t s, int max_new_tokens) const {
  N1_CHECK(loaded_, "System-2 weights not loaded");
  N1_CHECK(max_new_tokens != 0, "generation plan: max_new_tokens must be >= 1");
  std::unique_ptr<LlmPlan> p(new LlmPlan());
  const bool gen = max_new_tokens > 0;
  const int nq = gen ? 0 : dims.n_query, unit = dims.v_merge * dims.v_merge;
  p->B = B, p->n_query = dims.n_query;
  std::vector<int> kind, src, out_rows, pos_all;
  p->h_cu.push_back(0);
  int cursor = 0;
  long img_tok = 0, off = 0;
  std::vector<std::vector<int>> pos_seq(B);
  for (int b = 0; b < B; ++b) {
    const int len = lens[b];
    N1_CHECK(len > 0, "empty prompt");
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
    if (gen) out_rows.push_back((int)kind.size() - 1);  // logits of the last prompt token start the decode
    p->h_cu.push_back(p->h_cu.back() + L);
    p->max_len = std::max(p->max_len, L);
  }
  p->tokens = p->h_cu.back();
  p->n_image_tokens = img_tok;
  p->n_out = (int)out_rows.size();
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
  const int half = dims.head_dim / 2;
  std::vector<int> dest, len_v;
  if (gen) {
    p->max_new = max_new_tokens;
    p->slot = p->max_len + max_new_tokens + dims.n_query;
    dest.resize(p->tokens), len_v.resize(B);
    for (int b = 0; b < B; ++b) {
      len_v[b] = p->h_cu[b + 1] - p->h_cu[b];
      for (int i = 0; i < len_v[b]; ++i) dest[p->h_cu[b] + i] = b * p->slot + i;
    }
  }
  // one pooled block: [ints: cu | kind | src | out_rows | pos3 | dest | len | delta] [rope table]; one H2D copy from the
  // block's pinned twin, one kernel, no allocation / free / synchronisation once the pool holds a block of this size
  const std::vector<int>* parts[8] = {&p->h_cu, &kind, &src, &out_rows, &p->h_pos3, &dest, &len_v, &p->h_delta};
  size_t off_i[9] = {0};
  for (int i = 0; i < 8; ++i) off_i[i + 1] = off_i[i] + ((parts[i]->size() + 3) & ~size_t(3));  // 16-byte aligned parts
  const size_t int_bytes = (off_i[8] * sizeof(int) + 255) & ~size_t(255);
  const size_t need = int_bytes + (size_t)p->tokens * half * sizeof(float2);
  PlanBlock blk = pool_take(need);
  p->block = blk.dev, p->block_bytes = blk.bytes, p->last_use = blk.last_use;
  {
    std::lock_guard<std::mutex> lk(g_twin.mu);
    g_twin.map.emplace_back(blk.dev, blk.host);
  }
  N1_CUDA(cudaStreamWaitEvent(s, blk.last_use, 0));  // the previous owner's consumers (any stream) are done before we overwrite
  N1_CUDA(cudaEventSynchronize(blk.last_use));       // ... and before the pinned twin is rewritten (no-op unless just recycled)
  int* hp = static_cast<int*>(blk.host);
  for (int i = 0; i < 8; ++i)
    if (!parts[i]->empty()) memcpy(hp + off_i[i], parts[i]->data(), parts[i]->size() * sizeof(int));
  N1_CUDA(cudaMemcpyAsync(blk.dev, blk.host, off_i[8] * sizeof(int), cudaMemcpyHostToDevice, s));
  int* dp = static_cast<int*>(blk.dev);
  p->cu = dp + off_i[0], p->kind = dp + off_i[1], p->src = dp + off_i[2], p->out_rows = dp + off_i[3];
  int* pos = dp + off_i[4];
  if (gen) p->dest_rows = dp + off_i[5], p->d_len = dp + off_i[6], p->d_delta = dp + off_i[7];
  p->rope = reinterpret_cast<float2*>(static_cast<uint8_t*>(blk.dev) + int_bytes);
  mrope_table(pos, p->rope, p->tokens, half, dims.mrope[0], dims.mrope[1], dims.rope_theta, s);
  N1_CUDA(cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming));
  N1_CUDA(cudaEventRecord(p->ready, s));
  N1_CUDA(cudaEventRecord(p->last_use, s));  // a plan that is never consumed still leaves its block in a defined state
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
size_t S2Model::llm_impl(Carver c, const LlmPlan& p, const bf16* image_feats, bf16* out, cudaStream_t s,
                         const KvCache* kv) const {
  const int H = dims.hidden, hd = dims.head_dim;
  const long T = p.tokens;
  const int qkv_n = (dims.heads + 2 * dims.kv_heads) * hd;
  bf16* x = c.take<bf16>(T * H);
  const long sel_rows = 3L * p.n_out;  // the last layer reuses `ln` for three [n_out, H] buffers
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
    if (kv)  // keep the rotated keys and the values of every prompt token for the decode passes
      kv_append(qkv, qkv_n, dims.heads * hd, (dims.heads + dims.kv_heads) * hd, dims.kv_heads * hd, p.dest_rows, T,
                kv->k + l * kv->layer_stride, kv->v + l * kv->layer_stride, s);
    AttnParams a = {};
    a.q = qkv, a.k = qkv + (long)dims.heads * hd, a.v = qkv + (long)(dims.heads + dims.kv_heads) * hd, a.o = att;
    a.ldq = a.ldk = a.ldv = qkv_n, a.ldo = H;
    a.heads_q = dims.heads, a.heads_kv = dims.kv_heads, a.hd = hd;
    a.batch = p.B, a.cu_q = a.cu_k = p.cu, a.max_seq_q = p.max_len, a.total_rows = T;
    a.kv_div = 1, a.causal = 1, a.scale = 1.0f / sqrtf((float)hd);
    attention(a, s);
    GemmEpilogue res;
    res.residual = x, res.ldr = H;
    if (l == dims.layers - 1) {
      // Only the n_query TRAJ rows of each sequence are read after the last layer (internvla_n1.py L345): every token
      // still contributes K/V to the attention above, but o_proj and the MLP run on those B * n_query rows alone.
      // (Generation plans read the last prompt row of each sequence instead.)
      const int R = p.n_out;
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
  N1_CHECK(p.max_new == 0, "llm_prefill: this is a generation plan (use llm_generate)");
  if (ws_bytes < ws_llm(p)) throw Error(-7, "llm_prefill: workspace too small");
  p.wait_ready(s);
  llm_impl(Carver(ws, ws_bytes), p, image_feats, out, s);
  p.mark_used(s);
}

// ------------------------------------------------------------------------------------------------ greedy decode
struct S2Model::GenBufs {
  int *cur_tok, *gen, *finished, *next, *k_len, *n_active, *out_tokens, *dest, *pos3, *kind, *src;
  float2* rope;
  bf16 *x, *ln, *qkv, *att, *hid, *normed, *logits;
};

namespace {
void linear_rows(const Lin& L, const bf16* A, int lda, bf16* out, int ldo, int M, GemmEpilogue e, cudaStream_t s) {
  e.bias = L.b;
  gemm_bf16(A, lda, L.w, L.ldw, out, ldo, M, L.N, L.K, e, s);
}
}  // namespace

// One pass of B * per_seq new tokens (rows of g.x) through all layers against the cache: per_seq = 1 is a decode step,
// per_seq = 1 + n_query the latent pass.  g.dest / g.rope / g.k_len describe the chunk (gen_rows).  Result: final-norm
// states of every row in g.normed.
void S2Model::chunk_pass(const GenBufs& g, const LlmPlan& p, const KvCache& kv, int per_seq, cudaStream_t s) const {
  const int H = dims.hidden, hd = dims.head_dim, R = p.B * per_seq;
  const int qkv_n = (dims.heads + 2 * dims.kv_heads) * hd, kvd = dims.kv_heads * hd;
  for (int l = 0; l < dims.layers; ++l) {
    const LBlock& b = lblk_[l];
    layernorm(g.x, H, g.ln, H, b.n1, nullptr, R, H, dims.rms_eps, 1, s);
    linear_rows(b.qkv, g.ln, H, g.qkv, qkv_n, R, GemmEpilogue(), s);
    apply_rope(g.qkv, qkv_n, g.rope, R, dims.heads + dims.kv_heads, hd, s);
    bf16* ck = kv.k + l * kv.layer_stride;
    bf16* cv = kv.v + l * kv.layer_stride;
    kv_append(g.qkv, qkv_n, dims.heads * hd, (dims.heads + dims.kv_heads) * hd, kvd, g.dest, R, ck, cv, s);
    AttnParams a = {};
    a.q = g.qkv, a.k = ck, a.v = cv, a.o = g.att;
    a.ldq = qkv_n, a.ldk = a.ldv = kvd, a.ldo = H;
    a.heads_q = dims.heads, a.heads_kv = dims.kv_heads, a.hd = hd;
    a.batch = p.B, a.seq_q = per_seq, a.k_len = g.k_len, a.k_slot = p.slot;
    a.kv_div = 1, a.causal = 1, a.scale = 1.0f / sqrtf((float)hd);
    attention(a, s);
    GemmEpilogue res;
    res.residual = g.x, res.ldr = H;
    linear_rows(b.o, g.att, H, g.x, H, R, res, s);
    layernorm(g.x, H, g.ln, H, b.n2, nullptr, R, H, dims.rms_eps, 1, s);
    GemmEpilogue sw;
    sw.act = ACT_SWIGLU;
    linear_rows(b.gateup, g.ln, H, g.hid, inter_pad_, R, sw, s);
    linear_rows(b.down, g.hid, inter_pad_, g.x, H, R, res, s);
  }
  layernorm(g.x, H, g.normed, H, final_norm_, nullptr, R, H, dims.rms_eps, 1, s);
}

size_t S2Model::gen_impl(Carver c, const LlmPlan& p, const bf16* image_feats, const int32_t* eos, int n_eos,
                         int32_t pad, GenResult* out, bf16* latents, cudaStream_t s) const {
  const int H = dims.hidden, hd = dims.head_dim, B = p.B, nq = dims.n_query;
  const int qkv_n = (dims.heads + 2 * dims.kv_heads) * hd, kvd = dims.kv_heads * hd, half = hd / 2;
  const int R5 = B * (nq + 1);
  KvCache kv;
  kv.layer_stride = (long)B * p.slot * kvd;
  kv.k = c.take<bf16>((size_t)dims.layers * kv.layer_stride);
  kv.v = c.take<bf16>((size_t)dims.layers * kv.layer_stride);
  GenBufs g;
  g.cur_tok = c.take<int>(B), g.gen = c.take<int>(B), g.finished = c.take<int>(B), g.next = c.take<int>(B);
  g.k_len = c.take<int>(B), g.n_active = c.take<int>(1), g.out_tokens = c.take<int>((size_t)B * p.max_new);
  g.dest = c.take<int>(R5), g.pos3 = c.take<int>(3 * R5), g.kind = c.take<int>(R5), g.src = c.take<int>(R5);
  g.rope = c.take<float2>((size_t)R5 * half);
  g.x = c.take<bf16>((size_t)R5 * H), g.ln = c.take<bf16>((size_t)R5 * H), g.qkv = c.take<bf16>((size_t)R5 * qkv_n);
  g.att = c.take<bf16>((size_t)R5 * H), g.hid = c.take<bf16>((size_t)R5 * inter_pad_);
  g.normed = c.take<bf16>((size_t)R5 * H), g.logits = c.take<bf16>((size_t)B * dims.vocab);
  if (c.dry()) return llm_impl(c, p, nullptr, nullptr, nullptr, &kv);  // prefill scratch follows the decode state

  // 1. prompt prefill, K/V kept; final-norm state of the last prompt token of each sequence -> g.normed [B, H]
  llm_impl(c, p, image_feats, g.normed, s, &kv);
  std::vector<int32_t> fill((size_t)B * p.max_new, pad);
  N1_CUDA(cudaMemcpyAsync(g.out_tokens, fill.data(), fill.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  N1_CUDA(cudaMemsetAsync(g.gen, 0, B * sizeof(int), s));
  N1_CUDA(cudaMemsetAsync(g.finished, 0, B * sizeof(int), s));
  int steps = 0;
  for (int it = 0; it < p.max_new; ++it) {
    // logits = lm_head(hidden[:, -1]) in bf16, next = argmax (GenerationMixin greedy search)
    linear_rows(lm_head_, g.normed, H, g.logits, dims.vocab, B, GemmEpilogue(), s);
    argmax_rows(g.logits, dims.vocab, dims.vocab, B, g.next, s);
    gen_update(g.next, g.cur_tok, g.gen, g.finished, g.out_tokens, p.max_new, eos, n_eos, B, g.n_active, s);
    int active = 0;
    N1_CUDA(cudaMemcpyAsync(&active, g.n_active, sizeof(int), cudaMemcpyDeviceToHost, s));
    N1_CUDA(cudaStreamSynchronize(s));
    if (active == 0) break;
    // 2. one decode pass: the token just sampled is row (len + gen - 1) of its sequence
    gen_rows(p.d_len, p.d_delta, g.gen, 1, 1, B, p.slot, g.dest, g.pos3, g.k_len, s);
    mrope_table(g.pos3, g.rope, B, half, dims.mrope[0], dims.mrope[1], dims.rope_theta, s);
    gather_rows(embed_, g.cur_tok, g.x, B, 1, H, s);
    chunk_pass(g, p, kv, 1, s);
    ++steps;
  }
  if (latents) {
    // 3. generate_latents(output_ids, ...) on the cache: the last sampled token has no K/V yet, so the pass covers
    //    [last token, TRAJ x nq]; rows 1..nq of each sequence are the latent plan (internvla_n1.py L327, L345).
    gen_rows(p.d_len, p.d_delta, g.gen, 1, nq + 1, B, p.slot, g.dest, g.pos3, g.k_len, s);
    mrope_table(g.pos3, g.rope, R5, half, dims.mrope[0], dims.mrope[1], dims.rope_theta, s);
    latent_src(g.cur_tok, B, nq, g.kind, g.src, s);
    build_embeds(g.kind, g.src, embed_, nullptr, latentq_, g.x, R5, H, s);
    chunk_pass(g, p, kv, nq + 1, s);
    N1_CUDA(cudaMemcpy2DAsync(latents, (size_t)nq * H * sizeof(bf16), g.normed + H, (size_t)(nq + 1) * H * sizeof(bf16),
                              (size_t)nq * H * sizeof(bf16), B, cudaMemcpyDeviceToDevice, s));
  }
  N1_CUDA(cudaMemcpyAsync(out->tokens, g.out_tokens, (size_t)B * p.max_new * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  N1_CUDA(cudaMemcpyAsync(out->lens, g.gen, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  N1_CUDA(cudaStreamSynchronize(s));
  out->steps = steps;
  return c.used();
}

size_t S2Model::ws_generate(const LlmPlan& p) const {
  N1_CHECK(p.max_new > 0, "ws_generate: not a generation plan");
  return gen_impl(Carver(nullptr, 0), p, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr);
}
void S2Model::llm_generate(const LlmPlan& p, void* ws, size_t ws_bytes, const bf16* image_feats, const int32_t* eos,
                           int n_eos, int32_t pad, GenResult& out, bf16* latents, cudaStream_t s) const {
  N1_CHECK(loaded_ && ws, "llm_generate: not loaded / null workspace");
  N1_CHECK(p.max_new > 0, "llm_generate: the plan was not created for generation");
  if (!has_lm_head()) throw Error(-6, "llm_generate: lm_head.weight was not part of the loaded state_dict");
  N1_CHECK(out.tokens && out.lens, "llm_generate: null output buffers");
  if (ws_bytes < ws_generate(p)) throw Error(-7, "llm_generate: workspace too small");
  p.wait_ready(s);
  gen_impl(Carver(ws, ws_bytes), p, image_feats, eos, n_eos, pad, &out, latents, s);
  p.mark_used(s);
}

}  // namespace n1
