"""ctypes bindings of the backward primitives (include/n1b200.h, "training: backward primitives").

STATUS: first version, compiled but not yet validated on a B200 -- used only by tests/test_bwd_ops_gpu.py (skipped until
a parity run is on record) and, later, by the training step.  Nothing on the inference path imports this module.
"""
import ctypes
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

_bound = False
BWD_SYMBOLS = ["n1_op_act_fwd", "n1_op_transpose", "n1_op_colsum", "n1_op_norm_bwd", "n1_op_act_bwd", "n1_op_swiglu_bwd",
               "n1_op_rope_transposed", "n1_op_attention_bwd", "n1_op_adamw"]


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        vp = c_void_p
        L.n1_op_transpose.argtypes = [vp, c_int, c_int, c_int, vp, c_int, c_int, vp]
        L.n1_op_colsum.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, c_int, vp]
        L.n1_op_norm_bwd.argtypes = [vp, c_int, vp, c_int, vp, vp, c_int, vp, c_int, vp, vp, c_int, c_int, c_float, c_int,
                                     c_int, vp]
        L.n1_op_act_fwd.argtypes = [vp, vp, c_int64, c_int, vp]
        L.n1_op_act_bwd.argtypes = [vp, vp, vp, c_int64, c_int, vp]
        L.n1_op_swiglu_bwd.argtypes = [vp, vp, vp, c_int64, c_int, vp]
        L.n1_op_rope_transposed.argtypes = [vp, c_int, vp, c_int64, c_int, c_int, vp]
        L.n1_op_attention_bwd.argtypes = [vp] * 8 + [c_int] * 12 + [vp, vp, c_int, c_int, c_int, c_float, vp, c_int, vp]
        L.n1_op_adamw.argtypes = [vp, vp, vp, vp, vp, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, vp]
        for n in BWD_SYMBOLS:
            getattr(L, n).restype = c_int
        _bound = True
    return L


def transpose(x, rows_pad=None):
    """x bf16 [rows, cols] -> [cols, rows_pad] with zero padding (rows_pad defaults to rows rounded up to 8)."""
    rows, cols = x.shape
    rows_pad = rows_pad or (rows + 7) // 8 * 8
    out = torch.empty(cols, rows_pad, dtype=torch.bfloat16, device=x.device)
    check(_L().n1_op_transpose(ptr(x), rows, cols, x.stride(0), ptr(out), rows_pad, rows_pad, stream_ptr()))
    return out


def colsum(a, b=None, out=None, accumulate=False):
    rows, cols = a.shape
    if out is None:
        out = torch.zeros(cols, dtype=torch.float32, device=a.device)
    check(_L().n1_op_colsum(ptr(a), ptr(b), rows, cols, a.stride(0), b.stride(0) if b is not None else 0, ptr(out),
                            1 if accumulate else 0, stream_ptr()))
    return out


def norm_bwd(dy, x, w, eps, rms=False, residual_grad=None, need_param_grads=True):
    rows, D = x.shape
    dx = torch.empty_like(x)
    dw = torch.zeros(D, dtype=torch.float32, device=x.device) if need_param_grads else None
    db = torch.zeros(D, dtype=torch.float32, device=x.device) if need_param_grads and not rms else None
    check(_L().n1_op_norm_bwd(ptr(dy), dy.stride(0), ptr(x), x.stride(0), ptr(w), ptr(residual_grad),
                              residual_grad.stride(0) if residual_grad is not None else 0, ptr(dx), dx.stride(0), ptr(dw),
                              ptr(db), rows, D, eps, 1 if rms else 0, 0, stream_ptr()))
    return dx, dw, db


def act_fwd(pre, act):
    out = torch.empty_like(pre)
    check(_L().n1_op_act_fwd(ptr(pre), ptr(out), pre.numel(), act, stream_ptr()))
    return out


def act_bwd(pre, dy, act):
    out = torch.empty_like(pre)
    check(_L().n1_op_act_bwd(ptr(pre), ptr(dy), ptr(out), pre.numel(), act, stream_ptr()))
    return out


def swiglu_bwd(pre, dact):
    out = torch.empty_like(pre)
    check(_L().n1_op_swiglu_bwd(ptr(pre), ptr(dact), ptr(out), dact.shape[0], dact.shape[1], stream_ptr()))
    return out


def rope_transposed(x, cos_sin, heads, head_dim):
    """in place on the first `heads` heads of every row of x [rows, >= heads * head_dim]; cos_sin fp32 [rows, hd/2, 2]"""
    check(_L().n1_op_rope_transposed(ptr(x), x.stride(0), ptr(cos_sin), x.shape[0], heads, head_dim, stream_ptr()))
    return x


def attention_bwd(q, k, v, o, dout, heads_q, heads_kv, head_dim, batch, seq_q, seq_k, causal=False, kv_div=1, scale=None,
                  cu_q=None, cu_k=None, max_seq_q=0, k_len=None, k_slot=0):
    dq = torch.zeros(q.shape[0], heads_q * head_dim, dtype=torch.bfloat16, device=q.device)
    dk = torch.zeros(k.shape[0], heads_kv * head_dim, dtype=torch.float32, device=q.device)
    dv = torch.zeros_like(dk)
    scale = head_dim ** -0.5 if scale is None else scale
    for t in (q, k, v, o, dout):   # column slices of packed projections are legal operands: row stride + unit inner stride
        assert t.is_cuda and t.dtype == torch.bfloat16 and t.stride(1) == 1
    vp = lambda t: c_void_p(t.data_ptr())
    check(_L().n1_op_attention_bwd(vp(q), vp(k), vp(v), vp(o), vp(dout), ptr(dq), ptr(dk), ptr(dv), q.stride(0),
                                   k.stride(0), v.stride(0), o.stride(0), dout.stride(0), dq.stride(0), heads_q, heads_kv,
                                   head_dim, batch, seq_q, seq_k, ptr(cu_q), ptr(cu_k), max_seq_q, kv_div,
                                   1 if causal else 0, scale, ptr(k_len), k_slot, stream_ptr()))
    return dq, dk, dv


def adamw(master, working, grad, m, v, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, step=1):
    check(_L().n1_op_adamw(ptr(master), ptr(working), ptr(grad), ptr(m), ptr(v), master.numel(), lr, betas[0], betas[1], eps,
                           weight_decay, step, stream_ptr()))
