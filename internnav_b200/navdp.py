"""Host-side mirror of the reference System-1 head, dispatching to libn1b200.so.

`NavDP_Policy_DPT_CriticSum_DAT` keeps the constructor arguments, attribute names and method signatures of the
reference class (internnav/model/basemodel/internvla_n1/navdp.py L16-312) so `InternVLAN1MetaModel.build_navdp`
(internvla_n1_arch.py L141-143) and `generate_traj` (internvla_n1.py L434-441) work unchanged, but no arithmetic runs
in PyTorch: weights are handed to `n1_s1_load` once, and `predict_noise` / `predict_pointgoal_action_async` are C-ABI
calls.  Extensions over the reference (which is batch = 1): every method accepts B environments and
treats each exactly like the reference treats its single one.
"""
import ctypes

import torch

from . import _lib
from ._lib import OP_DENOISE, OP_GOAL, OP_RGBD, S1Dims, TensorDesc, c_void_p, check


class _Workspace:
    """Grow-only device scratch, one per (object, stream); the C ABI never allocates."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        return self.buf


class NavDP_Policy_DPT_CriticSum_DAT(torch.nn.Module):
    def __init__(self, image_size=224, memory_size=2, predict_size=32, temporal_depth=16, heads=8, token_dim=384,
                 vlm_token_dim=3584, channels=3, dropout=0.1, scratch=False, finetune=False, use_critic=False,
                 input_dtype="bf16", navdp_pretrained=None, navdp_version=0.0, device="cuda:0", n_query=4):
        super().__init__()
        assert image_size == 224 and token_dim == 384 and heads == 8, "n1b200 System-1 kernels: 224 px, D=384, 8 heads"
        self.image_size, self.memory_size, self.predict_size = image_size, memory_size, predict_size
        self.temporal_depth, self.attention_heads, self.token_dim = temporal_depth, heads, token_dim
        self.vlm_token_dim, self.input_channels, self.dropout = vlm_token_dim, channels, dropout
        self.use_critic, self.n_query = use_critic, n_query
        self.input_dtype = torch.bfloat16 if input_dtype == "bf16" else torch.float32
        self.num_train_timesteps = 20  # DDPMScheduler(num_train_timesteps=20, ...) navdp.py L74-76
        self.navdp_pretrained = navdp_pretrained
        self.model_name = "NavDP_Policy_DPT_CriticSum_DAT"
        self._handle = None
        self._device = torch.device(device)
        self._ws = {}
        self._keepalive = None

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Takes a reference-format state_dict (keys as in SURVEY.md Appendix A; an optional `model.navdp.` prefix is
        stripped), moves it to the device and packs it inside the library."""
        sd = {}
        for k, v in state_dict.items():
            for pre in ("model.navdp.", "navdp."):
                if k.startswith(pre):
                    k = k[len(pre):]
            sd[k] = v
        self._load(sd)
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def _load(self, sd):
        L = _lib.lib()
        dev = self._device
        if dev.type != "cuda":
            raise RuntimeError("n1b200 has no CPU path: construct NavDP with device='cuda:N'")
        if not L.n1_device_ok(dev.index or 0):
            raise _lib.N1Error(-4, L.n1_last_error().decode())
        h = c_void_p()
        check(L.n1_create(ctypes.byref(h), dev.index or 0))
        keep, descs = [], []
        for name, t in sd.items():
            if not torch.is_tensor(t) or t.dtype not in (torch.float32, torch.bfloat16, torch.float16, torch.float64):
                continue
            if t.dtype in (torch.float16, torch.float64):
                t = t.float()
            t = t.detach().to(dev).contiguous()
            keep.append(t)
            d = TensorDesc()
            d.name = name.encode()
            d.data = t.data_ptr()
            d.dtype = _lib.dtype_code(t)
            d.ndim = min(t.dim(), 4)
            shape = list(t.shape) if t.dim() <= 4 else [t.numel()]
            for i, s in enumerate(shape[:4]):
                d.shape[i] = s
            if t.dim() > 4:
                d.ndim = 1
            descs.append(d)
        arr = (TensorDesc * len(descs))(*descs)
        dims = S1Dims(self.token_dim, self.attention_heads, self.temporal_depth, self.predict_size, self.memory_size,
                      self.vlm_token_dim, self.n_query)
        with torch.cuda.device(dev):
            check(L.n1_s1_load(h, ctypes.byref(dims), arr, len(descs), _lib.stream_ptr()))
            torch.cuda.synchronize()
        if self._handle is not None:
            L.n1_destroy(self._handle)
        self._handle = h
        del keep  # packed copies live inside the handle

    def load_model(self):
        if self.navdp_pretrained is None:
            return
        sd = torch.load(self.navdp_pretrained, map_location="cpu")
        self.load_state_dict(sd.get("state_dict", sd))

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().n1_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _h(self):
        if self._handle is None:
            raise RuntimeError("NavDP weights not loaded (call load_state_dict first)")
        return self._handle

    def _scratch(self, op, B, Ns=0, T=0):
        n = _lib.lib().n1_workspace_bytes(self._h(), op, B, Ns, T)
        if n == 0:
            raise _lib.N1Error(-7, _lib.lib().n1_last_error().decode())
        key = (op, torch.cuda.current_stream().cuda_stream)
        ws = self._ws.setdefault(key, _Workspace())
        return ws.get(n, self._device), n

    # ------------------------------------------------------------------ reference API
    def rgbd_encoder(self, images, depths):
        """DAT_RGBD_Patch_Backbone.forward (navdp_backbone.py L151-202): [B,F,224,224,3], [B,F,224,224,1] -> [B,16F,384]."""
        B = images.shape[0]
        want = (self.memory_size, self.image_size, self.image_size)
        if tuple(images.shape[1:]) != want + (3,) or tuple(depths.shape[1:]) != want + (1,) or depths.shape[0] != B:
            raise ValueError("rgbd_encoder takes images [B, %d, %d, %d, 3] and depths [B, %d, %d, %d, 1]; got %s and %s"
                             % (want + want + (tuple(images.shape), tuple(depths.shape))))
        rgb = images.to(self._device, torch.float32).contiguous()
        dep = depths.to(self._device, torch.float32).contiguous()
        out = torch.empty(B, 16 * self.memory_size, self.token_dim, device=self._device, dtype=torch.bfloat16)
        ws, n = self._scratch(OP_RGBD, B)
        check(_lib.lib().n1_rgbd_encode(self._h(), _lib.ptr(ws), n, _lib.ptr(rgb), _lib.ptr(dep), _lib.ptr(out), B,
                                       _lib.stream_ptr()))
        return out

    def rgb_memory_tokens(self, images):
        """Training branch: tokens of the frozen RGB ViT (final norm, cls dropped; WITHOUT former_pe, which is trainable
        and added by the training schedule from its master copy): images [B, frames, 224, 224, 3] -> bf16
        [B, frames * 256, D]."""
        B = images.shape[0]
        want = (self.memory_size, self.image_size, self.image_size, 3)
        if tuple(images.shape[1:]) != want:
            raise ValueError("rgb_memory_tokens takes images [B, %d, %d, %d, 3]; got %s" % (want[:3] + (tuple(images.shape),)))
        rgb = images.to(self._device, torch.float32).contiguous()
        mem = torch.zeros(B, 2 * self.memory_size * 256, self.token_dim, device=self._device, dtype=torch.bfloat16)
        L = _lib.lib()
        nb = L.n1_rgb_tokens_workspace_bytes(self._h(), B)
        ws = torch.empty(int(nb) + 256, dtype=torch.uint8, device=self._device)
        check(L.n1_rgb_tokens(self._h(), _lib.ptr(ws), nb, _lib.ptr(rgb), _lib.ptr(mem), B, _lib.stream_ptr()))
        return mem[:, : self.memory_size * 256]

    def goal_embed(self, vlm_tokens):
        """vlm_embed_mlp + goal_compressor (navdp.py L237-238): [B, n_query, 3584] -> [B, 1, 384]."""
        B = vlm_tokens.shape[0]
        lat = vlm_tokens.to(self._device, torch.bfloat16).contiguous()
        assert lat.shape[1] == self.n_query and lat.shape[2] == self.vlm_token_dim
        out = torch.empty(B, 1, self.token_dim, device=self._device, dtype=torch.bfloat16)
        ws, n = self._scratch(OP_GOAL, B)
        check(_lib.lib().n1_goal_compress(self._h(), _lib.ptr(ws), n, _lib.ptr(lat), _lib.ptr(out), B,
                                         _lib.stream_ptr()))
        return out

    def predict_noise(self, last_actions, timestep, goal_embed, rgbd_embed=None):
        """navdp.py L177-195.  last_actions [B*Ns, T, 3]; timestep tensor [1] (shared) or [B]; goal [B,1,384];
        rgbd [B,16F,384].  Returns eps [B*Ns, T, 3] in last_actions.dtype."""
        assert rgbd_embed is not None, "the async N1 path always conditions on RGB-D memory"
        B = goal_embed.shape[0]
        R, T, _ = last_actions.shape
        Ns = R // B
        x = last_actions.to(self._device, torch.float32).contiguous()
        goal = goal_embed.to(self._device, torch.bfloat16).contiguous()
        rgbd = rgbd_embed.to(self._device, torch.bfloat16).contiguous()
        eps = torch.empty_like(x)
        ts = timestep.reshape(-1)
        if ts.numel() == 1:
            tdev, tsc = None, int(ts.item())
        else:
            tdev, tsc = ts.to(self._device, torch.int32).contiguous(), 0
        ws, n = self._scratch(OP_DENOISE, B, Ns, T)
        check(_lib.lib().n1_navdp_eps(self._h(), _lib.ptr(ws), n, _lib.ptr(x), _lib.ptr(tdev), tsc, _lib.ptr(goal),
                                     _lib.ptr(rgbd), _lib.ptr(eps), B, Ns, T, _lib.stream_ptr()))
        return eps.to(last_actions.dtype)

    GRAPH_MAX_ROWS = 16384  # below this many decoder rows the K-step loop is launch-bound: replay it as a CUDA graph

    def sample(self, goal_embed, rgbd_embed, x_init, step_noise, num_steps=None, graph=None):
        """The DDPM loop of navdp.py L242-253 as one C-ABI call.  x_init [B*Ns,T,3], step_noise [K-1,B*Ns,T,3].
        `graph`: replay the call's launch sequence (thousands of short kernels) from a cached CUDA graph; default = only
        in the launch-bound regime (small batches)."""
        K = num_steps or self.num_train_timesteps
        if graph is None:
            graph = x_init.shape[0] * x_init.shape[1] <= self.GRAPH_MAX_ROWS and not torch.cuda.is_current_stream_capturing()
        if graph:
            return self._sample_graphed(goal_embed, rgbd_embed, x_init, step_noise, K)
        return self._sample_eager(goal_embed, rgbd_embed, x_init, step_noise, K)

    def _sample_graphed(self, goal_embed, rgbd_embed, x_init, step_noise, K):
        dev = self._device
        key = (tuple(goal_embed.shape), tuple(x_init.shape), K, step_noise is None)
        ent = getattr(self, "_graphs", None)
        if ent is None:
            ent = self._graphs = {}
        hit = ent.get(key)
        if hit is None:
            st = dict(goal=torch.empty(goal_embed.shape, device=dev, dtype=torch.bfloat16),
                      rgbd=torch.empty(rgbd_embed.shape, device=dev, dtype=torch.bfloat16),
                      x0=torch.empty(x_init.shape, device=dev, dtype=torch.float32),
                      nz=None if step_noise is None else torch.empty(step_noise.shape, device=dev, dtype=torch.float32))
            for k, v in (("goal", goal_embed), ("rgbd", rgbd_embed), ("x0", x_init), ("nz", step_noise)):
                if v is not None:
                    st[k].copy_(v)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._sample_eager(st["goal"], st["rgbd"], st["x0"], st["nz"], K)  # warm-up: attributes, scratch of `side`
                side.synchronize()
                before = _lib.prof_read()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):  # the S2 thread may allocate meanwhile
                    out = self._sample_eager(st["goal"], st["rgbd"], st["x0"], st["nz"], K)
                nodes = _lib.prof_read()  # kernels captured into the graph = launches of every replay
                _lib.lib().n1_prof_add(before["gemm_launches"], before["total_launches"])  # restore the running counters
            torch.cuda.current_stream(dev).wait_stream(side)
            hit = ent[key] = (g, st, out, side, nodes)
        g, st, out, _, nodes = hit
        for k, v in (("goal", goal_embed), ("rgbd", rgbd_embed), ("x0", x_init), ("nz", step_noise)):
            if v is not None:
                st[k].copy_(v, non_blocking=True)
        g.replay()
        _lib.lib().n1_prof_add(nodes["gemm_launches"], nodes["total_launches"])
        return out.clone()

    def _sample_eager(self, goal_embed, rgbd_embed, x_init, step_noise, K):
        B = goal_embed.shape[0]
        R, T, _ = x_init.shape
        Ns = R // B
        x0 = x_init.to(self._device, torch.float32).contiguous()
        nz = step_noise.to(self._device, torch.float32).contiguous() if step_noise is not None else None
        if nz is not None:
            assert nz.shape[0] == K - 1 and tuple(nz.shape[1:]) == tuple(x0.shape)
        goal = goal_embed.to(self._device, torch.bfloat16).contiguous()
        rgbd = rgbd_embed.to(self._device, torch.bfloat16).contiguous()
        out = torch.empty_like(x0)
        ws, n = self._scratch(OP_DENOISE, B, Ns, T)
        check(_lib.lib().n1_navdp_sample(self._h(), _lib.ptr(ws), n, _lib.ptr(goal), _lib.ptr(rgbd), _lib.ptr(x0),
                                        _lib.ptr(nz), _lib.ptr(out), B, Ns, T, K, _lib.stream_ptr()))
        return out

    def draw_noise(self, B, sample_num, dtype):
        """Consumes the torch RNG streams exactly like the reference: the initial sample is drawn on the CPU generator
        in the model dtype and moved (navdp.py L242-244); each of the K-1 variance-noise tensors is drawn on the device
        generator in that dtype (diffusers randn_tensor inside DDPMScheduler.step)."""
        K = self.num_train_timesteps
        x_init = torch.randn((sample_num * B, self.predict_size, 3), dtype=dtype).to(self._device)
        with torch.cuda.device(self._device):
            noise = torch.stack([torch.randn((sample_num * B, self.predict_size, 3), device=self._device, dtype=dtype)
                                 for _ in range(K - 1)]) if K > 1 else None
        return x_init, noise

    # ------------------------------------------------------------------ training-branch forward (no autograd yet)
    def add_noise(self, original_samples, noise, timesteps):
        """DDPMScheduler.add_noise (navdp.py L173): sqrt(acp_t) x0 + sqrt(1 - acp_t) eps, tables from the library."""
        K = self.num_train_timesteps
        buf = (ctypes.c_float * (K * 5))()
        check(_lib.lib().n1_ddpm_tables(K, buf))
        tab = torch.tensor(list(buf), dtype=torch.float32, device=original_samples.device).view(K, 5)
        t = timesteps.to(original_samples.device).long()
        s2 = tab[t, 0].view(-1, 1, 1)                 # sqrt(1 - acp_t)
        s1 = (1.0 / tab[t, 1]).view(-1, 1, 1)         # sqrt(acp_t)
        return (s1 * original_samples.float() + s2 * noise.float()).to(original_samples.dtype)

    def forward_vlm_traj(self, vlm_tokens, input_images, input_depths, tensor_label_actions,
                         tensor_augment_actions=None, noise=None, timesteps=None):
        """navdp.py L291-312 (the System-1 half of the training forward): one noise prediction per
        (episode, frame) sample with its own timestep and its own condition.  The returned tensors carry no autograd graph
        (the backward of this branch is scheduled by train_s1.S1TrainStep on the library's backward kernels); `noise` / `timesteps` may be passed in, else
        they are drawn like `sample_noise` (navdp.py L165-175).  Returns (noise_pred, noise)."""
        label = tensor_label_actions.flatten(0, 1).to(self._device)
        n = label.shape[0]
        if noise is None:
            noise = torch.randn(label.shape, dtype=label.dtype).to(self._device)
        if timesteps is None:
            timesteps = torch.randint(0, self.num_train_timesteps, (n,)).long()
        noisy = self.add_noise(label, noise.to(self._device), timesteps)
        goal = self.goal_embed(vlm_tokens)
        rgbd = self.rgbd_encoder(input_images, input_depths)
        pred = self.predict_noise(noisy.float(), timesteps, goal, rgbd)
        return pred.to(label.dtype), noise.to(self._device)

    def predict_pointgoal_action_async(self, vlm_tokens, input_images=None, input_depths=None, vlm_mask=None,
                                       sample_num=32, x_init=None, step_noise=None):
        """navdp.py L197-253.  Reference semantics are bs = 1 (extra rows are dropped at L227-231); here every one of
        the B rows is an independent environment.  Returns [sample_num * B, predict_size, 3] in the latent dtype."""
        with torch.no_grad():
            B = vlm_tokens.shape[0]
            dtype = vlm_tokens.dtype if vlm_tokens.dtype in (torch.bfloat16, torch.float32) else torch.float32
            goal = self.goal_embed(vlm_tokens)
            rgbd = self.rgbd_encoder(input_images, input_depths)
            if x_init is None:
                x_init, step_noise = self.draw_noise(B, sample_num, dtype)
            traj = self.sample(goal, rgbd, x_init, step_noise)
            return traj.to(dtype)
