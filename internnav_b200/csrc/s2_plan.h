// Host-side integer planning for the System-2 path: everything the reference derives from `input_ids` and
// `image_grid_thw` with Python loops, restated in C++ (bit-exact integer work):
//   * get_rope_index  (images only)     -- internnav/dataset/rope2d.py L6-181 == HF Qwen2_5_VLModel.get_rope_index
//   * rot_pos_emb position ids, get_window_index, cu_seqlens -- transformers modeling_qwen2_5_vl.py
//     (Qwen2_5_VisionTransformerPretrainedModel.rot_pos_emb / get_window_index / forward)
#pragma once
#include <stdint.h>

#include <vector>

namespace n1 {

constexpr int kImageTokenId = 151655;   // IMAGE_TOKEN_INDEX, internvla_n1.py L19
constexpr int kVideoTokenId = 151656;
constexpr int kVisionStartId = 151652;
constexpr int kTrajTokenId = 151667;    // TRAJ_TOKEN_INDEX, internvla_n1.py L18

struct VitIndex {
  long n_patches = 0;                 // sum t*h*w
  std::vector<int> window_index;      // [n_patches / merge^2] merged-token order -> source merged token
  std::vector<int> reverse_index;     // argsort(window_index)
  std::vector<int> cu_window;         // unique_consecutive(cu_window_seqlens), in patches
  std::vector<int> cu_full;           // per (image, frame) sequence boundaries, in patches
  std::vector<int> pos_hw;            // [n_patches, 2] (h, w) ids in WINDOW order
  int max_window = 0, max_full = 0;
};
// grid_thw: [n_img, 3] (t, h, w) in patches.  merge = spatial_merge_size, window = window_size / merge / patch_size.
void vit_index(const int32_t* grid_thw, int n_img, int merge, int vit_merger_window, VitIndex& out);

// One sequence: position ids [3, len] and the mrope delta.  `cursor` = index of the next unused image in grid_thw
// (advanced), matching the reference's running image_index across the batch.
void rope_index_one(const int32_t* ids, int len, const int32_t* grid_thw, int n_img, int merge, int& cursor,
                    std::vector<int>& pos3, int& delta);

}  // namespace n1
