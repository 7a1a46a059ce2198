"""TEST INFRASTRUCTURE ONLY -- the fused per-layer attention written out with the C oracle.

One oracle call per (query frame t1, neighbour frame t2), exactly as the reference module loops
(models/ops/modules/ms_deform_attn.py:130-225), with the glue (bias add, loc = ref + off / (W,H), softmax over
levels x points / k) in fp32 torch on the CPU *in the kernel's operation order*, so the sampling locations are
bit-identical to the ones the CUDA kernels form and floor() picks the same cells.  Differentiable: the C
oracle's analytic backward is wrapped in an autograd.Function.
"""
import torch

from oracle import c_oracle


def neighbour_frames(t1, n_frame, T2):
    if t1 < n_frame:
        return [t for t in (t1 - 1, t1, t1 + 1) if 0 <= t < n_frame]
    return list(range(T2))


class _OracleOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, shapes, lsi, loc, attn):
        ctx.save_for_backward(value, shapes, lsi, loc, attn)
        return c_oracle.forward(value, shapes, lsi, loc, attn)

    @staticmethod
    def backward(ctx, grad_out):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        gv, gl, ga = c_oracle.backward(value, shapes, lsi, loc, attn, grad_out.contiguous())
        return gv, None, None, gl, ga


def snippet_attention_oracle(value, mask, shapes, lsi, proj, off_bias, logit_bias, ref, n_frame):
    """value (N,T2,S,M,D) fp32 CPU; mask (N,T2,S) bool or None; proj (N,T1,Lq,3*M*L*P); ref (N,T1,Lq,L,2)
    -> out (N,T1,Lq,M*D).  All arguments may require grad."""
    N, T2, S, M, D = value.shape
    _, T1, Lq, W = proj.shape
    L = shapes.shape[0]
    P = W // (3 * M * L)
    mlp = M * L * P
    if mask is not None:
        value = value.masked_fill(mask[..., None, None], 0.0)
    off = proj[..., :2 * mlp]
    logits = proj[..., 2 * mlp:]
    if off_bias is not None:
        off = off + off_bias
    if logit_bias is not None:
        logits = logits + logit_bias
    off = off.view(N, T1, Lq, M, L, P, 2)
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(off.dtype)
    loc = ref[:, :, :, None, :, None, :] + off / wh[None, None, None, None, :, None, :]
    att = torch.softmax(logits.view(N, T1, Lq, M, L * P), -1).view(N, T1, Lq, M, L, P)
    outs = []
    for t1 in range(T1):
        nb = neighbour_frames(t1, n_frame, T2)
        a = (att[:, t1] / len(nb)).contiguous()
        l1 = loc[:, t1].contiguous()
        acc = 0
        for t2 in nb:
            acc = acc + _OracleOp.apply(value[:, t2].contiguous(), shapes, lsi, l1, a)
        outs.append(acc)
    return torch.stack(outs, 1)
