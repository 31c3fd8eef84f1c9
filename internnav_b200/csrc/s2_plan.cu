#include "s2_plan.h"

#include <algorithm>
#include <numeric>

#include "n1_ops.h"

namespace n1 {

void vit_index(const int32_t* grid_thw, int n_img, int merge, int win, VitIndex& o) {
  o = VitIndex();
  const int unit = merge * merge;
  std::vector<int> cu_window_raw = {0};
  int window_index_id = 0;
  std::vector<int> pos_src;  // (h, w) per patch in the processor's patch order (merge-block major)
  for (int im = 0; im < n_img; ++im) {
    const int t = grid_thw[im * 3], h = grid_thw[im * 3 + 1], w = grid_thw[im * 3 + 2];
    N1_CHECK(t > 0 && h > 0 && w > 0 && h % merge == 0 && w % merge == 0, "bad image_grid_thw");
    // ---- rot_pos_emb ids: hpos/wpos reshaped (h/m, m, w/m, m) -> permute(0, 2, 1, 3) -> flatten, repeated t times
    for (int rep = 0; rep < t; ++rep)
      for (int bh = 0; bh < h / merge; ++bh)
        for (int bw = 0; bw < w / merge; ++bw)
          for (int ih = 0; ih < merge; ++ih)
            for (int iw = 0; iw < merge; ++iw) {
              pos_src.push_back(bh * merge + ih);
              pos_src.push_back(bw * merge + iw);
            }
    // ---- get_window_index
    const int gh = h / merge, gw = w / merge;
    const int pad_h = win - gh % win, pad_w = win - gw % win;  // NB: a full extra window when already divisible
    const int nwh = (gh + pad_h) / win, nww = (gw + pad_w) / win;
    for (int tt = 0; tt < t; ++tt)
      for (int wh = 0; wh < nwh; ++wh)
        for (int ww = 0; ww < nww; ++ww) {
          int count = 0;
          for (int ih = 0; ih < win; ++ih)
            for (int iw = 0; iw < win; ++iw) {
              const int y = wh * win + ih, x = ww * win + iw;
              if (y < gh && x < gw) {
                o.window_index.push_back(window_index_id + (tt * gh + y) * gw + x);
                ++count;
              }
            }
          cu_window_raw.push_back(cu_window_raw.back() + count * unit);
        }
    window_index_id += t * gh * gw;
    for (int tt = 0; tt < t; ++tt) o.cu_full.push_back(h * w);
    o.n_patches += (long)t * h * w;
  }
  // unique_consecutive
  for (int v : cu_window_raw)
    if (o.cu_window.empty() || o.cu_window.back() != v) o.cu_window.push_back(v);
  // cu_full: cumsum with leading zero
  {
    std::vector<int> c = {0};
    for (int v : o.cu_full) c.push_back(c.back() + v);
    o.cu_full = c;
  }
  for (size_t i = 1; i < o.cu_window.size(); ++i) o.max_window = std::max(o.max_window, o.cu_window[i] - o.cu_window[i - 1]);
  for (size_t i = 1; i < o.cu_full.size(); ++i) o.max_full = std::max(o.max_full, o.cu_full[i] - o.cu_full[i - 1]);
  const int n_merged = (int)o.window_index.size();
  N1_CHECK((long)n_merged * unit == o.n_patches, "window index does not cover all patches");
  o.reverse_index.resize(n_merged);
  for (int i = 0; i < n_merged; ++i) o.reverse_index[o.window_index[i]] = i;  // argsort of a permutation
  // position ids reordered like hidden_states[window_index] on groups of `unit` patches
  o.pos_hw.resize(o.n_patches * 2);
  for (int i = 0; i < n_merged; ++i)
    for (int u = 0; u < unit; ++u) {
      const long dst = (long)i * unit + u, src = (long)o.window_index[i] * unit + u;
      o.pos_hw[dst * 2] = pos_src[src * 2];
      o.pos_hw[dst * 2 + 1] = pos_src[src * 2 + 1];
    }
}

void rope_index_one(const int32_t* ids, int len, const int32_t* grid_thw, int n_img, int merge, int& cursor,
                    std::vector<int>& pos3, int& delta) {
  pos3.assign((size_t)3 * len, 0);
  int image_nums = 0;
  for (int i = 0; i + 1 < len; ++i)
    if (ids[i] == kVisionStartId && ids[i + 1] == kImageTokenId) ++image_nums;
  int st = 0;
  int out = 0;          // number of positions written
  int last_max = -1;    // max of the last appended block
  bool any = false;
  auto put = [&](int a, int b, int c) {
    N1_CHECK(out < len, "rope index overflow (image tokens do not match image_grid_thw)");
    pos3[out] = a, pos3[len + out] = b, pos3[2 * len + out] = c;
    ++out;
  };
  for (int n = 0; n < image_nums; ++n) {
    int ed = -1;
    for (int i = st; i < len; ++i)
      if (ids[i] == kImageTokenId) {
        ed = i;
        break;
      }
    N1_CHECK(ed >= 0, "image token not found");
    N1_CHECK(cursor < n_img, "more image placeholders than rows in image_grid_thw");
    const int t = grid_thw[cursor * 3], h = grid_thw[cursor * 3 + 1] / merge, w = grid_thw[cursor * 3 + 2] / merge;
    ++cursor;
    const int text_len = ed - st;
    const int st_idx = any ? last_max + 1 : 0;
    for (int i = 0; i < text_len; ++i) put(i + st_idx, i + st_idx, i + st_idx);
    // images: second_per_grid_t = 0 -> temporal index 0 for every frame
    for (int tt = 0; tt < t; ++tt)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) put(0 + text_len + st_idx, y + text_len + st_idx, x + text_len + st_idx);
    last_max = std::max(h, w) - 1 + text_len + st_idx;
    if (text_len > 0 && t * h * w == 0) last_max = text_len - 1 + st_idx;
    any = true;
    st = ed + t * h * w;
  }
  if (st < len) {
    const int st_idx = any ? last_max + 1 : 0;
    const int text_len = len - st;
    for (int i = 0; i < text_len; ++i) put(i + st_idx, i + st_idx, i + st_idx);
    last_max = text_len - 1 + st_idx;
    any = true;
  }
  N1_CHECK(out == len, "rope index: sequence length mismatch");
  int mx = 0;
  for (int v : pos3) mx = std::max(mx, v);
  delta = mx + 1 - len;
}

}  // namespace n1
