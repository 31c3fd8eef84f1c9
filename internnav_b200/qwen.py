"""Host-side driver of the System-2 path (Qwen2.5-VL vision tower + decoder prefill) over libn1b200.so.

`System2` owns the packed weights (inside the library handle) and caches integer plans; its methods are the batched
equivalents of the calls InternVLAN1ForCausalLM.generate_latents makes (internvla_n1.py L320-347):
`visual(pixel_values, grid_thw)` and `model(inputs_embeds, position_ids)` + the last-n_query slice.
"""
import ctypes

import torch

from . import _lib
from ._lib import TensorDesc, c_void_p, check

QWEN25VL_7B = dict(
    v_depth=32, v_hidden=1280, v_heads=16, v_inter=3420, v_patch=14, v_tpatch=2, v_merge=2, v_window=112, v_out=3584,
    fullatt=[7, 15, 23, 31],
    layers=28, hidden=3584, heads=28, kv_heads=4, head_dim=128, inter=18944, vocab=152064,
    rms_eps=1e-6, rope_theta=1000000.0, mrope=[16, 24, 24], n_query=4)


class S2Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("v_depth", "v_hidden", "v_heads", "v_inter", "v_patch", "v_tpatch",
                                              "v_merge", "v_window", "v_out", "n_fullatt")] + \
               [("fullatt", ctypes.c_int32 * 16)] + \
               [(n, ctypes.c_int32) for n in ("layers", "hidden", "heads", "kv_heads", "head_dim", "inter", "vocab")] + \
               [("rms_eps", ctypes.c_float), ("rope_theta", ctypes.c_float), ("mrope", ctypes.c_int32 * 3),
                ("n_query", ctypes.c_int32)]


def _dims_struct(cfg):
    d = S2Dims()
    for k in ("v_depth", "v_hidden", "v_heads", "v_inter", "v_patch", "v_tpatch", "v_merge", "v_window", "v_out",
              "layers", "hidden", "heads", "kv_heads", "head_dim", "inter", "vocab", "n_query"):
        setattr(d, k, int(cfg[k]))
    d.n_fullatt = len(cfg["fullatt"])
    for i, v in enumerate(cfg["fullatt"]):
        d.fullatt[i] = int(v)
    d.rms_eps, d.rope_theta = float(cfg["rms_eps"]), float(cfg["rope_theta"])
    for i in range(3):
        d.mrope[i] = int(cfg["mrope"][i])
    return d


def _bind(L):
    if getattr(L, "_s2_bound", False):
        return
    vp = c_void_p
    L.n1_s2_load.restype = ctypes.c_int
    L.n1_s2_load.argtypes = [vp, ctypes.POINTER(S2Dims), ctypes.POINTER(TensorDesc), ctypes.c_int, vp]
    L.n1_vit_plan_create.restype = ctypes.c_int
    L.n1_vit_plan_create.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.POINTER(vp), vp]
    L.n1_vit_plan_destroy.restype = None
    L.n1_vit_plan_destroy.argtypes = [vp]
    L.n1_vit_plan_patches.restype = ctypes.c_int64
    L.n1_vit_plan_patches.argtypes = [vp]
    L.n1_llm_plan_create.restype = ctypes.c_int
    L.n1_llm_plan_create.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.POINTER(vp), vp]
    L.n1_llm_plan_destroy.restype = None
    L.n1_llm_plan_destroy.argtypes = [vp]
    L.n1_llm_plan_tokens.restype = ctypes.c_int64
    L.n1_llm_plan_tokens.argtypes = [vp]
    L.n1_llm_plan_image_tokens.restype = ctypes.c_int64
    L.n1_llm_plan_image_tokens.argtypes = [vp]
    L.n1_llm_plan_positions.restype = ctypes.c_int
    L.n1_llm_plan_positions.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    L.n1_vit_workspace_bytes.restype = ctypes.c_size_t
    L.n1_vit_workspace_bytes.argtypes = [vp, vp]
    L.n1_llm_workspace_bytes.restype = ctypes.c_size_t
    L.n1_llm_workspace_bytes.argtypes = [vp, vp]
    L.n1_qwen_vit.restype = ctypes.c_int
    L.n1_qwen_vit.argtypes = [vp, vp, vp, ctypes.c_size_t, vp, vp, vp]
    L.n1_llm_prefill.restype = ctypes.c_int
    L.n1_llm_prefill.argtypes = [vp, vp, vp, ctypes.c_size_t, vp, vp, vp]
    L.n1_rope_index.restype = ctypes.c_int
    L.n1_vit_window_index.restype = ctypes.c_int
    L.n1_gen_plan_create.restype = ctypes.c_int
    L.n1_gen_plan_create.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp), vp]
    L.n1_generate_workspace_bytes.restype = ctypes.c_size_t
    L.n1_generate_workspace_bytes.argtypes = [vp, vp]
    L.n1_s2_has_lm_head.restype = ctypes.c_int
    L.n1_s2_has_lm_head.argtypes = [vp]
    L.n1_llm_generate.restype = ctypes.c_int
    L.n1_llm_generate.argtypes = [vp, vp, vp, ctypes.c_size_t, vp, ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                  ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), vp,
                                  ctypes.POINTER(ctypes.c_int32), vp]
    L.n1_s2_set_latent_queries.restype = ctypes.c_int
    L.n1_s2_set_latent_queries.argtypes = [vp, vp, vp]
    L.n1_s2_train_workspace_bytes.restype = ctypes.c_size_t
    L.n1_s2_train_workspace_bytes.argtypes = [vp, vp]
    L.n1_s2_train_forward.restype = ctypes.c_int
    L.n1_s2_train_forward.argtypes = [vp, vp, vp, ctypes.c_size_t, vp, vp, vp]
    L.n1_s2_train_backward.restype = ctypes.c_int
    L.n1_s2_train_backward.argtypes = [vp, vp, vp, ctypes.c_size_t, vp, vp, vp]
    L._s2_bound = True


S2_SYMBOLS = ["n1_s2_load", "n1_vit_plan_create", "n1_vit_plan_destroy", "n1_vit_plan_patches", "n1_llm_plan_create",
              "n1_llm_plan_destroy", "n1_llm_plan_tokens", "n1_llm_plan_image_tokens", "n1_llm_plan_positions",
              "n1_vit_workspace_bytes", "n1_llm_workspace_bytes", "n1_qwen_vit", "n1_llm_prefill", "n1_rope_index",
              "n1_vit_window_index", "n1_gen_plan_create", "n1_generate_workspace_bytes", "n1_s2_has_lm_head",
              "n1_llm_generate", "n1_s2_train_workspace_bytes", "n1_s2_train_forward", "n1_s2_train_backward", "n1_s2_set_latent_queries"]

EOS_TOKEN_IDS = (151645, 151643)  # Qwen2.5-VL generation_config.json: <|im_end|>, <|endoftext|>
PAD_TOKEN_ID = 151643


def normalise_keys(state_dict):
    """Accept the transformers 4.51 layout (`visual.*`, `model.*`) and the 5.x one (`model.visual.*`,
    `model.language_model.*`); System-1 keys (`model.navdp.*`) are dropped, `lm_head.weight` is kept for generate()."""
    out = {}
    for k, v in state_dict.items():
        if k.startswith("model.visual."):
            k = k[len("model."):]
        elif k.startswith("model.language_model."):
            k = "model." + k[len("model.language_model."):]
        if k.startswith("model.navdp."):
            continue
        if k.startswith("visual.") or k.startswith("model.") or k == "lm_head.weight":
            out[k] = v
    return out


class System2:
    def __init__(self, cfg=None, device="cuda:0"):
        self.cfg = dict(QWEN25VL_7B if cfg is None else cfg)
        self.device = torch.device(device)
        self._handle = None
        self._vit_plans, self._llm_plans = {}, {}
        self._ws = {}

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict):
        L = _lib.lib()
        _bind(L)
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("n1b200 has no CPU path: System2 needs device='cuda:N'")
        sd = normalise_keys(state_dict)
        h = c_void_p()
        check(L.n1_create(ctypes.byref(h), dev.index or 0))
        keep, descs = [], []
        for name, t in sd.items():
            if t.dtype not in (torch.float32, torch.bfloat16):
                t = t.float()
            t = t.detach().to(dev).contiguous()
            keep.append(t)
            d = TensorDesc()
            d.name, d.data, d.dtype = name.encode(), t.data_ptr(), _lib.dtype_code(t)
            d.ndim = 1
            d.shape[0] = t.numel()
            descs.append(d)
        arr = (TensorDesc * len(descs))(*descs)
        dims = _dims_struct(self.cfg)
        with torch.cuda.device(dev):
            check(L.n1_s2_load(h, ctypes.byref(dims), arr, len(descs), _lib.stream_ptr()))
            torch.cuda.synchronize()
        self._handle = h
        del keep

    def __del__(self):
        try:
            L = _lib.lib()
            for p in self._vit_plans.values():
                L.n1_vit_plan_destroy(p)
            for p in self._llm_plans.values():
                L.n1_llm_plan_destroy(p[0])
            if self._handle is not None:
                L.n1_destroy(self._handle)
        except Exception:
            pass

    def _h(self):
        if self._handle is None:
            raise RuntimeError("System-2 weights not loaded")
        return self._handle

    def _scratch(self, key, nbytes):
        k = (key, torch.cuda.current_stream().cuda_stream)
        buf = self._ws.get(k)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
            self._ws[k] = buf
        return buf

    # ------------------------------------------------------------------ plans (integer host work, cached by shape)
    def vit_plan(self, grid_thw):
        key = tuple(int(v) for g in grid_thw for v in g)
        p = self._vit_plans.get(key)
        if p is None:
            L = _lib.lib()
            arr = (ctypes.c_int32 * len(key))(*key)
            p = c_void_p()
            with torch.cuda.device(self.device):
                check(L.n1_vit_plan_create(self._h(), arr, len(key) // 3, ctypes.byref(p), _lib.stream_ptr()))
            self._vit_plans[key] = p
        return p

    def llm_plan(self, prompts, grid_thw):
        """prompts: list of token-id lists (no TRAJ tokens); grid_thw: image grids in prompt order across the batch."""
        gkey = tuple(int(v) for g in grid_thw for v in g)
        key = (tuple(tuple(p) for p in prompts), gkey)
        hit = self._llm_plans.get(key)
        if hit is None:
            L = _lib.lib()
            flat = [int(t) for p in prompts for t in p]
            ids = (ctypes.c_int32 * len(flat))(*flat)
            lens = (ctypes.c_int32 * len(prompts))(*[len(p) for p in prompts])
            garr = (ctypes.c_int32 * max(1, len(gkey)))(*gkey)
            p = c_void_p()
            with torch.cuda.device(self.device):
                check(L.n1_llm_plan_create(self._h(), ids, lens, len(prompts), garr, len(gkey) // 3, ctypes.byref(p),
                                           _lib.stream_ptr()))
            self._evict_plans(L)
            hit = (p, len(prompts))
            self._llm_plans[key] = hit
        return hit[0]

    def _evict_plans(self, L, keep=8):
        """Prompts change every step in deployment, so the cache only serves repeated calls within a step (generate +
        latents).  Oldest first: the library recycles a destroyed plan's device block, and the oldest plan's consumers
        finished long ago, so the recycling never has to wait."""
        while len(self._llm_plans) > keep:
            k = next(iter(self._llm_plans))
            old, _ = self._llm_plans.pop(k)
            L.n1_llm_plan_destroy(old)

    def positions(self, plan, B):
        L = _lib.lib()
        n = L.n1_llm_plan_tokens(plan)
        pos = (ctypes.c_int32 * (3 * n))()
        dl = (ctypes.c_int32 * B)()
        check(L.n1_llm_plan_positions(plan, pos, dl))
        return torch.tensor(list(pos), dtype=torch.int64).view(3, n), torch.tensor(list(dl), dtype=torch.int64)

    # ------------------------------------------------------------------ hot calls
    def visual(self, pixel_values, grid_thw):
        """self.visual(pixel_values, grid_thw=image_grid_thw): [N, 1176] -> [N/4, hidden] bf16."""
        L = _lib.lib()
        plan = self.vit_plan(grid_thw)
        px = pixel_values.to(self.device, torch.bfloat16).contiguous()
        n = L.n1_vit_plan_patches(plan)
        assert px.shape[0] == n, "pixel_values rows (%d) do not match image_grid_thw (%d patches)" % (px.shape[0], n)
        out = torch.empty(n // (self.cfg["v_merge"] ** 2), self.cfg["v_out"], device=self.device, dtype=torch.bfloat16)
        nb = L.n1_vit_workspace_bytes(self._h(), plan)
        ws = self._scratch("vit", nb)
        check(L.n1_qwen_vit(self._h(), plan, _lib.ptr(ws), nb, _lib.ptr(px), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def prefill_latents(self, prompts, image_feats, grid_thw):
        """Embedding splice + decoder prefill + last-n_query slice for B prompts -> [B, n_query, hidden] bf16."""
        L = _lib.lib()
        plan = self.llm_plan(prompts, grid_thw)
        B = len(prompts)
        feats = image_feats.to(self.device, torch.bfloat16).contiguous()
        assert feats.shape[0] == L.n1_llm_plan_image_tokens(plan), "image features and image tokens do not match"
        out = torch.empty(B, self.cfg["n_query"], self.cfg["hidden"], device=self.device, dtype=torch.bfloat16)
        nb = L.n1_llm_workspace_bytes(self._h(), plan)
        ws = self._scratch("llm", nb)
        check(L.n1_llm_prefill(self._h(), plan, _lib.ptr(ws), nb, _lib.ptr(feats), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def generate_latents(self, prompts, pixel_values, grid_thw):
        """Batched InternVLAN1ForCausalLM.generate_latents: one prompt per environment, images in prompt order."""
        return self.prefill_latents(prompts, self.visual(pixel_values, grid_thw), grid_thw)

    # ------------------------------------------------------------------ training branch (UNVALIDATED on GPU, see s2_train.cu)
    def train_forward(self, prompts, pixel_values, grid_thw, image_feats=None):
        """Hidden states at the TRAJ positions for a training batch (prompts WITHOUT the TRAJ tokens): [B, n_query, H].
        Keeps the K/V cache and the per-layer TRAJ-row tensors for `train_backward` (same prompts, next call)."""
        L = _lib.lib()
        plan = self.gen_plan(prompts, grid_thw, 1)
        feats = self.visual(pixel_values, grid_thw) if image_feats is None else image_feats
        feats = feats.to(self.device, torch.bfloat16).contiguous()
        out = torch.empty(len(prompts), self.cfg["n_query"], self.cfg["hidden"], device=self.device, dtype=torch.bfloat16)
        nb = L.n1_s2_train_workspace_bytes(self._h(), plan)
        ws = self._scratch("train", nb)
        check(L.n1_s2_train_forward(self._h(), plan, _lib.ptr(ws), nb, _lib.ptr(feats), _lib.ptr(out), _lib.stream_ptr()))
        self._train_state = (plan, ws, nb)
        return out

    def set_latent_queries(self, latent_queries):
        """Push updated `latent_queries` ([1, n_query, H] or [n_query, H]) into the library after an optimizer step."""
        t = latent_queries.reshape(self.cfg["n_query"], self.cfg["hidden"]).to(self.device, torch.bfloat16).contiguous()
        check(_lib.lib().n1_s2_set_latent_queries(self._h(), _lib.ptr(t), _lib.stream_ptr()))
        torch.cuda.current_stream().synchronize()

    def train_backward(self, grad_states):
        """d loss / d traj states [B, n_query, H] -> d loss / d latent_queries fp32 [1, n_query, H]."""
        plan, ws, nb = self._train_state
        g = grad_states.to(self.device, torch.bfloat16).contiguous()
        out = torch.empty(1, self.cfg["n_query"], self.cfg["hidden"], device=self.device, dtype=torch.float32)
        check(_lib.lib().n1_s2_train_backward(self._h(), plan, _lib.ptr(ws), nb, _lib.ptr(g), _lib.ptr(out),
                                              _lib.stream_ptr()))
        return out

    def gen_plan(self, prompts, grid_thw, max_new_tokens):
        gkey = tuple(int(v) for g in grid_thw for v in g)
        key = ("gen", int(max_new_tokens), tuple(tuple(p) for p in prompts), gkey)
        hit = self._llm_plans.get(key)
        if hit is None:
            L = _lib.lib()
            flat = [int(t) for p in prompts for t in p]
            ids = (ctypes.c_int32 * len(flat))(*flat)
            lens = (ctypes.c_int32 * len(prompts))(*[len(p) for p in prompts])
            garr = (ctypes.c_int32 * max(1, len(gkey)))(*gkey)
            p = c_void_p()
            with torch.cuda.device(self.device):
                check(L.n1_gen_plan_create(self._h(), ids, lens, len(prompts), garr, len(gkey) // 3, int(max_new_tokens),
                                           ctypes.byref(p), _lib.stream_ptr()))
            self._evict_plans(L)
            hit = (p, len(prompts))
            self._llm_plans[key] = hit
        return hit[0]

    def generate(self, prompts, pixel_values, grid_thw, max_new_tokens=128, eos_token_ids=EOS_TOKEN_IDS,
                 pad_token_id=PAD_TOKEN_ID, with_latents=False, image_feats=None):
        """Greedy decode for B prompts (`model.generate(do_sample=False, max_new_tokens=...)`, internvla_n1_policy.py
        L169-176).  Returns (list of B generated-token lists, each ending with its eos id unless the budget ran out,
        latents [B, n_query, hidden] or None, decode passes run).  With `with_latents` the K/V cache of the decode is
        extended by the TRAJ tokens, which equals `generate_latents(output_ids, ...)` without a second prefill."""
        L = _lib.lib()
        if not L.n1_s2_has_lm_head(self._h()):
            raise RuntimeError("generate() needs lm_head.weight in the loaded state_dict; there is no fallback")
        plan = self.gen_plan(prompts, grid_thw, max_new_tokens)
        B = len(prompts)
        feats = self.visual(pixel_values, grid_thw) if image_feats is None else image_feats
        feats = feats.to(self.device, torch.bfloat16).contiguous()
        assert feats.shape[0] == L.n1_llm_plan_image_tokens(plan), "image features and image tokens do not match"
        lat = torch.empty(B, self.cfg["n_query"], self.cfg["hidden"], device=self.device, dtype=torch.bfloat16) \
            if with_latents else None
        nb = L.n1_generate_workspace_bytes(self._h(), plan)
        ws = self._scratch("gen", nb)
        eos = (ctypes.c_int32 * max(1, len(eos_token_ids)))(*[int(e) for e in eos_token_ids])
        toks = (ctypes.c_int32 * (B * int(max_new_tokens)))()
        lens = (ctypes.c_int32 * B)()
        passes = ctypes.c_int32(0)
        check(L.n1_llm_generate(self._h(), plan, _lib.ptr(ws), nb, _lib.ptr(feats), eos, len(eos_token_ids),
                                int(pad_token_id), toks, lens, _lib.ptr(lat), ctypes.byref(passes), _lib.stream_ptr()))
        out = [list(toks[b * max_new_tokens: b * max_new_tokens + lens[b]]) for b in range(B)]
        return out, lat, passes.value
