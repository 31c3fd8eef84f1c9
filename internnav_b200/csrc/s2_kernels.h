// Launchers for the System-2 (Qwen2.5-VL) glue kernels (s2_kernels.cu).
#pragma once
#include "n1_ops.h"

namespace n1 {

// dst[r, :] = src[idx[r / group] * group + r % group, :]   (rows of `cols` bf16; cols % 8 == 0)
void gather_rows(const bf16* src, const int* idx, bf16* dst, long rows, int group, int cols, cudaStream_t s);
// ViT 2-D rotary tables: cs[tok, j] = (cos, sin)(pos[tok, j < half/2 ? 0 : 1] * theta^(-2 (j % (half/2)) / half)),
// j in [0, half), half = head_dim / 2.
void vit_rope_table(const int* pos_hw, float2* cs, long tokens, int half, float theta, cudaStream_t s);
// LLM multimodal rotary tables: cs[tok, j] = (cos, sin)(pos[stream(j)][tok] * theta^(-2 j / hd)), j in [0, hd/2);
// stream(j) from mrope_section (t, h, w widths summing to hd/2); pos is [3, tokens] int32.
void mrope_table(const int* pos3, float2* cs, long tokens, int half, int sec_t, int sec_h, float theta, cudaStream_t s);
// In-place rotate-half rotary on the first `heads` heads (each `hd` wide) of every row of x (row stride ld):
// out[j] = x[j] c_j - x[j + hd/2] s_j ; out[j + hd/2] = x[j + hd/2] c_j + x[j] s_j, tables shared by all heads.
void apply_rope(bf16* x, int ld, const float2* cs, long tokens, int heads, int hd, cudaStream_t s);
// LLM input embeddings: kind[tok] 0 -> embed_tokens[src[tok]], 1 -> image_feats[src[tok]], 2 -> latent_queries[src[tok]]
void build_embeds(const int* kind, const int* src, const bf16* embed_tokens, const bf16* image_feats,
                  const bf16* latent_queries, bf16* out, long tokens, int H, cudaStream_t s);

}  // namespace n1
