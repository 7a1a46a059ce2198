"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REAL reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

It imports the reference's own ``ms_deform_attn_core_pytorch``
(models/ops/functions/ms_deform_attn_func.py:45-65) and ``MSDeformAttn`` module
(models/ops/modules/ms_deform_attn.py:37-243, use_pytroch_deform=True) on CPU and stores
their inputs, outputs and autograd gradients.  Nothing from the reference is copied into the
repo; only numeric vectors are committed.  The GPU box has no /root/reference, so tests read
these files instead.

Cases
  testpy_seed3   : the fixture of the reference's only test, models/ops/test.py:21-85
                   (seed 3, shapes (6,4),(3,2), N=1 M=2 Lq=2 L=2 P=2; fwd fp64, fwd fp32,
                   gradients for D in 30,32,64,71)
  snipper_small  : Snipper-like geometry (M=8, D=48, L=3, P=4) on small levels, encoder-style
                   local sampling including out-of-range points, N=1
  frames_levels  : k neighbour frames presented as k*L levels
  module_encoder / module_decoder : the reference nn.Module, fp64, perturbed weights,
                   padding mask, T=4 (+2 future frames for the decoder, attention_vis on)
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SNIPPER_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _ref():
    sys.path.insert(0, REF)
    from models.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
    from models.ops.modules import MSDeformAttn
    return ms_deform_attn_core_pytorch, MSDeformAttn


def _lsi(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def _np(t):
    return t.detach().cpu().numpy()


def _grads(core, value, shapes, loc, attn, grad_out):
    v = value.clone().requires_grad_(True)
    s = loc.clone().requires_grad_(True)
    a = attn.clone().requires_grad_(True)
    out = core(v, shapes, s, a)
    out.backward(grad_out)
    return out.detach(), v.grad, s.grad, a.grad


def case_testpy(core):
    # reference models/ops/test.py:21-36 -- same seed, same draw order, CPU generator
    N, M, Dd = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    d = {"shapes": _np(shapes), "lsi": _np(_lsi(shapes))}

    def draw(D):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
        return value, loc, attn

    value, loc, attn = draw(Dd)  # check_forward_equal_with_pytorch_double
    d.update(dbl_value=_np(value), dbl_loc=_np(loc), dbl_attn=_np(attn),
             dbl_out=_np(core(value.double(), shapes, loc.double(), attn.double())))
    value, loc, attn = draw(Dd)  # check_forward_equal_with_pytorch_float
    d.update(flt_value=_np(value), flt_loc=_np(loc), flt_attn=_np(attn),
             flt_out=_np(core(value, shapes, loc, attn)))
    gen = torch.Generator().manual_seed(1234)
    for D in (30, 32, 64, 71):  # check_gradient_numerical channel list (the stored subset)
        value, loc, attn = draw(D)
        go = torch.randn(N, Lq, M * D, generator=gen, dtype=torch.float64)
        out, gv, gl, ga = _grads(core, value.double(), shapes, loc.double(), attn.double(), go)
        k = "g%d_" % D
        d.update({k + "value": _np(value), k + "loc": _np(loc), k + "attn": _np(attn),
                  k + "grad_out": _np(go), k + "out": _np(out), k + "grad_value": _np(gv),
                  k + "grad_loc": _np(gl), k + "grad_attn": _np(ga)})
    return d


def _encoder_like(N, M, D, shapes, P, sigma_px, seed):
    """Queries = every pixel of every level; samples = pixel centre + N(0, sigma) px offset,
    some pushed outside [0,1] to exercise zero padding."""
    g = torch.Generator().manual_seed(seed)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    refs = []
    for H, W in shapes.tolist():
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5,
                                torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
    ref = torch.cat(refs, 0)  # (S,2) valid for every level (normalised coords)
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
    off = torch.randn(N, S, M, L, P, 2, generator=g) * sigma_px
    loc = ref[None, :, None, None, None, :] + off / wh[None, None, None, :, None, :]
    value = torch.randn(N, S, M, D, generator=g)
    attn = torch.softmax(torch.randn(N, S, M, L * P, generator=g), -1).view(N, S, M, L, P)
    go = torch.randn(N, S, M * D, generator=g)
    return value, loc, attn, go


def case_snipper_small(core):
    shapes = torch.as_tensor([(5, 7), (3, 4), (2, 2)], dtype=torch.long)
    value, loc, attn, go = _encoder_like(1, 8, 48, shapes, 4, 2.0, seed=0)
    out64, gv, gl, ga = _grads(core, value.double(), shapes, loc.double(), attn.double(), go.double())
    out32 = core(value, shapes, loc, attn)
    return dict(shapes=_np(shapes), lsi=_np(_lsi(shapes)), value=_np(value), loc=_np(loc),
                attn=_np(attn), grad_out=_np(go), out_f64=_np(out64), out_f32=_np(out32),
                grad_value=_np(gv), grad_loc=_np(gl), grad_attn=_np(ga))


def case_frames_levels(core):
    """k=3 neighbour frames as 3*L levels: the oracle gets a concatenated copy of the frames
    (it splits ``value`` by level sizes, ms_deform_attn_func.py:50)."""
    base = torch.as_tensor([(4, 5), (2, 3)], dtype=torch.long)
    k = 3
    shapes = base.repeat(k, 1)
    value, loc, attn, go = _encoder_like(1, 4, 16, shapes, 2, 2.0, seed=5)
    out64, gv, gl, ga = _grads(core, value.double(), shapes, loc.double(), attn.double(), go.double())
    return dict(shapes=_np(shapes), lsi=_np(_lsi(shapes)), value=_np(value), loc=_np(loc),
                attn=_np(attn), grad_out=_np(go), out_f64=_np(out64),
                grad_value=_np(gv), grad_loc=_np(gl), grad_attn=_np(ga))


def case_module(MSDeformAttn, mode):
    torch.manual_seed(11 if mode == "encoder" else 12)
    d_model, M, L, P, n_frame = 48, 4, 3, 4, 4
    shapes = torch.as_tensor([(4, 6), (2, 3), (1, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    N = 2 if mode == "decoder" else 1
    T2 = n_frame
    if mode == "encoder":
        T1, Lq, vis = n_frame, S, False
    else:
        T1, Lq, vis = n_frame + 2, 5, True
    mod = MSDeformAttn(d_model, L, M, P, n_frame, mode, True, vis).double()
    with torch.no_grad():  # move away from the degenerate init (zero weights)
        mod.sampling_offsets[0].weight.normal_(0, 0.5)
        mod.attention_weights[0].weight.normal_(0, 0.5)
        mod.attention_weights[0].bias.normal_(0, 0.5)
        for p in mod.parameters():
            p.copy_(p.float().double())  # keep every number fp32-representable
    query = torch.randn(N, T1, Lq, d_model).double().requires_grad_(True)
    ref = torch.rand(N, T1, Lq, L, 2).double().requires_grad_(True)
    src = torch.randn(N, T2, S, d_model).double().requires_grad_(True)
    pix_mask = torch.rand(N, 1, S, 1) < 0.15
    mask = pix_mask.expand(N, T2, S, d_model).contiguous()
    res = mod(query, ref, src, shapes, _lsi(shapes), mask)
    out, visd = (res if vis else (res, None))
    go = torch.randn(out.shape).double()
    out.backward(go)
    d = dict(shapes=_np(shapes), lsi=_np(_lsi(shapes)), query=_np(query), ref=_np(ref),
             src=_np(src), mask=_np(mask[..., 0]), grad_out=_np(go), out=_np(out),
             grad_query=_np(query.grad), grad_ref=_np(ref.grad), grad_src=_np(src.grad),
             cfg=np.array([d_model, L, M, P, n_frame, T1, T2, Lq, N]))
    for k, v in mod.state_dict().items():
        d["sd." + k] = _np(v)
    for k, p in mod.named_parameters():  # unique params only (aliases share storage)
        d["pg." + k] = _np(p.grad)
    if vis:
        for t1, (vl, va) in enumerate(zip(*visd)):
            d["vis_loc.%d" % t1] = _np(vl)
            d["vis_att.%d" % t1] = _np(va)
    return d


def main():
    core, MSDeformAttn = _ref()
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "testpy_seed3": case_testpy(core),
        "snipper_small": case_snipper_small(core),
        "frames_levels": case_frames_levels(core),
        "module_encoder": case_module(MSDeformAttn, "encoder"),
        "module_decoder": case_module(MSDeformAttn, "decoder"),
    }
    for name, d in cases.items():
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-16s %4d arrays %8.1f KB" % (name, len(d), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
