"""Plain fp32 PyTorch implementation of the `ops` contract of internnav_b200/train_s1.py -- TEST INFRASTRUCTURE.

Lets the CPU suite run the training SCHEDULE (which kernel is called with which operand, what is saved, in which order
gradients are accumulated) against the oracle's gradients.  It enforces the operand constraints of the real kernels
(K and N multiples of 8, unit inner stride, row strides multiples of 8 elements) so that a schedule that passes here is
callable on the GPU backend.  It says nothing about the kernels themselves (tests/test_bwd_ops_gpu.py does)."""
import math

import torch
import torch.nn.functional as F


class TorchOps:
    dtype = torch.float32

    def __init__(self):
        self.calls = {}

    def _count(self, k):
        self.calls[k] = self.calls.get(k, 0) + 1

    def cast(self, t):
        return t.float().contiguous()

    @staticmethod
    def _operand(t):
        assert t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0, ("GEMM operand layout", t.shape, t.stride())

    def mm_nt(self, a, w, bias=None, out_fp32=False):
        self._count("mm_nt")
        self._operand(a), self._operand(w)
        assert a.shape[1] == w.shape[1] and a.shape[1] % 8 == 0 and w.shape[0] % 8 == 0, (a.shape, w.shape)
        y = a @ w.t()
        return y + bias if bias is not None else y

    def wgrad(self, dy, x):
        """dy^T @ x with the kernel's operand constraints (rows of 8-element multiples) or the transposing fallback."""
        self._count("wgrad")
        assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0] and dy.stride(1) == 1 and x.stride(1) == 1
        if dy.stride(0) % 8 == 0 and x.stride(0) % 8 == 0 and dy.shape[1] % 8 == 0 and x.shape[1] % 8 == 0:
            return dy.t() @ x
        return self.mm_nt(self.transpose(dy), self.transpose(x), out_fp32=True)

    def transpose(self, x):
        self._count("transpose")
        assert x.dim() == 2 and x.stride(1) == 1
        r = x.shape[0]
        out = torch.zeros(x.shape[1], (r + 7) // 8 * 8)
        out[:, :r] = x.t()
        return out

    def colsum(self, a, b=None):
        self._count("colsum")
        assert a.stride(1) == 1 and (b is None or b.stride(1) == 1)
        return (a if b is None else a * b).sum(0)

    def layernorm(self, x, w, b, eps):
        self._count("layernorm")
        assert x.dim() == 2 and x.stride(1) == 1
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)

    def norm_bwd(self, dy, x, w, eps):
        self._count("norm_bwd")
        assert dy.is_contiguous() and x.is_contiguous()
        mu = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
        xh = (x - mu) * rstd
        g = dy * w
        dx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        return dx, (dy * xh).sum(0), dy.sum(0)

    @staticmethod
    def _heads(t, batch, s, heads, hd):
        assert t.stride(1) == 1 and t.shape == (batch * s, heads * hd)
        return t.reshape(batch, s, heads, hd).transpose(1, 2)

    def _probs(self, q, k, heads, hd, batch, sq, sk, causal):
        s = self._heads(q, batch, sq, heads, hd) @ self._heads(k, batch, sk, heads, hd).transpose(-1, -2) / math.sqrt(hd)
        if causal:
            i, j = torch.arange(sq)[:, None], torch.arange(sk)[None, :]
            s = s.masked_fill(j > i + (sk - sq), float("-inf"))
        return s.softmax(-1)

    def attention(self, q, k, v, heads, hd, batch, sq, sk, causal):
        self._count("attention")
        p = self._probs(q, k, heads, hd, batch, sq, sk, causal)
        return (p @ self._heads(v, batch, sk, heads, hd)).transpose(1, 2).reshape(batch * sq, heads * hd)

    def attention_bwd(self, q, k, v, o, do, heads, hd, batch, sq, sk, causal):
        self._count("attention_bwd")
        assert o.is_contiguous() and do.is_contiguous()
        p = self._probs(q, k, heads, hd, batch, sq, sk, causal)
        qh, kh, vh = (self._heads(t, batch, n, heads, hd) for t, n in ((q, sq), (k, sk), (v, sk)))
        doh = self._heads(do, batch, sq, heads, hd)
        dv = p.transpose(-1, -2) @ doh
        dp = doh @ vh.transpose(-1, -2)
        D = (doh * self._heads(o, batch, sq, heads, hd)).sum(-1, keepdim=True)      # the kernel's row term do . o
        ds = p * (dp - D) / math.sqrt(hd)
        dq, dk = ds @ kh, ds.transpose(-1, -2) @ qh
        back = lambda t, n: t.transpose(1, 2).reshape(batch * n, heads * hd)          # noqa: E731
        return back(dq, sq), back(dk, sk), back(dv, sk)

    def act_fwd(self, pre, kind):
        self._count("act_fwd")
        return F.gelu(pre) if kind == 1 else F.relu(pre)

    def act_bwd(self, pre, dy, kind):
        self._count("act_bwd")
        assert pre.is_contiguous() and dy.is_contiguous()
        if kind == 1:
            return dy * (0.5 * (1 + torch.erf(pre / math.sqrt(2))) + pre * torch.exp(-0.5 * pre * pre) / math.sqrt(2 * math.pi))
        return dy * (pre > 0)

    def sgemm(self, a, b, trans_a=False, trans_b=False):
        self._count("sgemm")
        assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.stride(1) == 1 and b.stride(1) == 1
        return (a.t() if trans_a else a) @ (b.t() if trans_b else b)

    def patchify_depth(self, frames):
        self._count("patchify_depth")
        assert frames.dtype == torch.float32 and frames.is_contiguous() and tuple(frames.shape[1:]) == (224, 224)
        cols = F.unfold(frames.unsqueeze(1), kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 196)
        return F.pad(cols, (0, 4))

    def scale_cols(self, x, gamma, add=None):
        self._count("scale_cols")
        assert x.dim() == 2 and x.stride(1) == 1 and (add is None or (add.shape == x.shape and add.stride(1) == 1))
        y = x * gamma
        return y + add if add is not None else y
