// Training branch, System-1 side: what the training schedule (internnav_b200/train_s1.py) needs from the inference
// executor -- the frozen RGB branch of the RGB-D encoder (navdp_backbone.py L151-171: image_token is detached).
// The tokens are delivered WITHOUT former_pe: that table is trainable (navdp_backbone.py L183-192 adds one embedding to
// RGB and depth tokens alike), so the training schedule adds the current master copy itself.
#include "s1_model.h"

namespace n1 {

size_t S1Model::ws_rgb_tokens(int B) const {
  return vit_forward(rgb_, Carver(nullptr, 0), nullptr, false, B * dims.frames, nullptr, 0, nullptr);
}

void S1Model::rgb_tokens(void* ws, size_t ws_bytes, const float* rgb, bf16* mem, int B, cudaStream_t s) const {
  N1_CHECK(loaded_ && ws && rgb && mem && B > 0, "rgb_tokens: not loaded / null buffers");
  if (ws_bytes < ws_rgb_tokens(B)) throw Error(-7, "rgb_tokens: workspace too small");
  vit_forward(rgb_, Carver(ws, ws_bytes), rgb, false, B * dims.frames, mem, 0, s, /*with_pe=*/false);
}

}  // namespace n1
