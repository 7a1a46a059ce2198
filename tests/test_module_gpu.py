"""GPU parity tests of the ``MSDeformAttn`` module (fused snippet kernels and per-call loop)
against (a) golden vectors from the reference module and (b) the CPU oracle restatement."""
import copy

import numpy as np
import pytest
import torch

from conftest import level_start_index, load_golden, rel_err
from oracle import torch_ref

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _golden_module(g, mode, dtype):
    from snipper_b200 import MSDeformAttn
    d_model, L, M, P, n_frame, T1, T2, Lq, N = [int(x) for x in g["cfg"]]
    mod = MSDeformAttn(d_model, L, M, P, n_frame, mode, False, mode == "decoder")
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd.")}
    mod.load_state_dict(sd, strict=True)  # reference checkpoint keys load as-is
    return mod.to(DEV, dtype)


@pytest.mark.parametrize("mode", ["encoder", "decoder"])
@pytest.mark.parametrize("dtype,ftol,btol", [(torch.float64, 1e-10, 1e-9), (torch.float32, 1e-5, 1e-4)])
def test_module_golden(mode, dtype, ftol, btol):
    """Reference MSDeformAttn (use_pytroch_deform=True, fp64) vs ours; D=12 -> per-call path."""
    g = load_golden("module_" + mode)
    mod = _golden_module(g, mode, dtype)
    d_model = int(g["cfg"][0])
    query = T(g["query"]).to(DEV, dtype).requires_grad_(True)
    ref = T(g["ref"]).to(DEV, dtype).requires_grad_(True)
    src = T(g["src"]).to(DEV, dtype).requires_grad_(True)
    mask = T(g["mask"])[..., None].expand(-1, -1, -1, d_model).to(DEV)
    res = mod(query, ref, src, T(g["shapes"]).to(DEV), T(g["lsi"]).to(DEV), mask)
    out, vis = res if mode == "decoder" else (res, None)
    assert rel_err(out, g["out"]) < ftol
    out.backward(T(g["grad_out"]).to(DEV, dtype))
    assert rel_err(query.grad, g["grad_query"]) < btol
    assert rel_err(ref.grad, g["grad_ref"]) < btol
    assert rel_err(src.grad, g["grad_src"]) < btol
    for k, p in mod.named_parameters():
        assert rel_err(p.grad, g["pg." + k]) < btol, k
    if vis is not None:
        for t1, (vl, va) in enumerate(zip(*vis)):
            assert rel_err(vl, g["vis_loc.%d" % t1]) < ftol
            assert rel_err(va, g["vis_att.%d" % t1]) < ftol


def _pair(d_model, M, L, P, n_frame, mode, seed):
    """(ours on GPU fp32, oracle restatement on CPU fp64) with identical perturbed weights."""
    from snipper_b200 import MSDeformAttn
    torch.manual_seed(seed)
    vis = mode == "decoder"
    ours = MSDeformAttn(d_model, L, M, P, n_frame, mode, False, vis)
    with torch.no_grad():
        ours.sampling_offsets[0].weight.normal_(0, 0.2)
        ours.attention_weights[0].weight.normal_(0, 0.3)
        ours.attention_weights[0].bias.normal_(0, 0.3)
    oracle = torch_ref.SnippetMSDeformAttnRef(d_model, L, M, P, n_frame, mode, True, vis).double()
    oracle.load_state_dict({k: v.double() for k, v in ours.state_dict().items()})
    return ours.to(DEV), oracle


def _inputs(N, T1, T2, Lq, shapes, d_model, seed, encoder_ref=False):
    g = torch.Generator().manual_seed(seed)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    query = torch.randn(N, T1, Lq, d_model, generator=g)
    src = torch.randn(N, T2, S, d_model, generator=g)
    if encoder_ref:  # one frame of reference points expanded over T1 (stride 0), as the encoder does
        ref = torch.rand(N, 1, Lq, L, 2, generator=g).expand(N, T1, Lq, L, 2)
    else:
        ref = torch.rand(N, T1, Lq, L, 2, generator=g)
    mask = (torch.rand(N, 1, S, 1, generator=g) < 0.1).expand(N, T2, S, d_model)
    grad_out = torch.randn(N, T1, Lq, d_model, generator=g)
    return query, ref, src, mask, grad_out


def _run(mod, dev, dtype, query, ref, src, mask, grad_out, shapes):
    q = query.to(dev, dtype).requires_grad_(True)
    r = ref.to(dev, dtype).requires_grad_(True)
    s = src.to(dev, dtype).requires_grad_(True)
    for p in mod.parameters():
        p.grad = None
    res = mod(q, r, s, shapes.to(dev), level_start_index(shapes).to(dev), mask.to(dev))
    out, vis = res if isinstance(res, tuple) else (res, None)
    out.backward(grad_out.to(dev, dtype))
    pg = {k: p.grad.detach().cpu() for k, p in mod.named_parameters()}
    return out.detach().cpu(), q.grad.cpu(), r.grad.cpu(), s.grad.cpu(), pg, vis


CONFIGS = [
    # d_model, M, L, P, n_frame, mode, T1 extra (future frames), Lq (None = S)
    (128, 8, 3, 4, 4, "encoder", 0, None),
    (128, 8, 3, 4, 4, "decoder", 2, 7),
    (384, 8, 3, 4, 4, "encoder", 0, None),   # Snipper: D = 48
    (384, 8, 3, 4, 4, "decoder", 2, 60),
    (384, 8, 3, 4, 1, "encoder", 0, None),   # T = 1 (BASELINE config 1)
    (256, 4, 2, 8, 3, "decoder", 1, 5),      # D = 64, P = 8
]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_fused_module_vs_oracle(cfg):
    d_model, M, L, P, n_frame, mode, fut, Lq = cfg
    shapes = torch.as_tensor([(9, 12), (5, 6), (3, 3)][:L], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    ours, oracle = _pair(d_model, M, L, P, n_frame, mode, seed=7)
    assert ours._can_fuse(torch.empty(1, device=DEV))
    N, T1, T2 = 2, n_frame + fut, n_frame
    inp = _inputs(N, T1, T2, Lq or S, shapes, d_model, seed=8, encoder_ref=(mode == "encoder"))
    got = _run(ours, DEV, torch.float32, *inp, shapes)
    want = _run(oracle, "cpu", torch.float64, *inp, shapes)
    assert rel_err(got[0], want[0]) < 1e-5
    for i in (1, 2, 3):
        assert rel_err(got[i], want[i]) < 1e-4, i
    for k in want[4]:
        assert rel_err(got[4][k], want[4][k]) < 1e-4, k
    if mode == "decoder":
        for a, b in zip(got[5][0], want[5][0]):
            assert a.shape == b.shape and rel_err(a, b) < 1e-5
        for a, b in zip(got[5][1], want[5][1]):
            assert a.shape == b.shape and rel_err(a, b) < 1e-5


@pytest.mark.parametrize("mode,fut,Lq", [("encoder", 0, None), ("decoder", 2, 11)])
def test_fused_equals_per_call_loop(mode, fut, Lq):
    """The one-launch fused path and the reference-style (t1,t2) loop agree on the GPU."""
    shapes = torch.as_tensor([(9, 12), (5, 6), (3, 3)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    ours, _ = _pair(384, 8, 3, 4, 4, mode, seed=3)
    loop = copy.deepcopy(ours)
    loop.fused = False
    inp = _inputs(2, 4 + fut, 4, Lq or S, shapes, 384, seed=4)
    a = _run(ours, DEV, torch.float32, *inp, shapes)
    b = _run(loop, DEV, torch.float32, *inp, shapes)
    assert rel_err(a[0], b[0]) < 1e-5
    for i in (1, 2, 3):
        assert rel_err(a[i], b[i]) < 1e-4
    for k in b[4]:
        assert rel_err(a[4][k], b[4][k]) < 1e-4, k


def test_unaliased_slots_fall_back_to_loop():
    """If a user un-aliases the frame slots, the fused shortcut is invalid and must not be taken."""
    from snipper_b200 import MSDeformAttn
    mod = MSDeformAttn(128, 3, 8, 4, 4, "encoder").to(DEV)
    assert mod._can_fuse(torch.empty(1, device=DEV))
    mod.sampling_offsets[1] = copy.deepcopy(mod.sampling_offsets[0])
    assert not mod._can_fuse(torch.empty(1, device=DEV))


def test_fused_op_empty_and_single_frame_edge_cases():
    """Degenerate shapes of the fused per-layer op: no queries, no batch, one frame, one query."""
    import snipper_b200  # noqa: F401
    shapes = torch.as_tensor([(5, 7), (3, 4)], dtype=torch.long, device=DEV)
    lsi = torch.as_tensor([0, 35], dtype=torch.long, device=DEV)
    S, M, D, L, P = 47, 8, 48, 2, 4

    def call(N, T2, T1, Lq, n_frame):
        value = torch.randn(N, T2, S, M, D, device=DEV)
        off = torch.randn(N, T1, Lq, M, L, P, 2, device=DEV)
        logits = torch.randn(N, T1, Lq, M, L, P, device=DEV)
        ref = torch.rand(N, T1, Lq, L, 2, device=DEV)
        return torch.ops.snipper_b200.snippet_forward(value, shapes, lsi, off, logits, ref, n_frame)

    assert call(1, 4, 4, 0, 4).shape == (1, 4, 0, M * D)
    assert call(0, 4, 4, 9, 4).shape == (0, 4, 9, M * D)
    out = call(2, 1, 1, 1, 1)                      # T = 1: every query frame has exactly one neighbour
    assert out.shape == (2, 1, 1, M * D) and torch.isfinite(out).all()
    out = call(1, 3, 5, 33, 3)                     # two future query frames attend to all three source frames
    assert out.shape == (1, 5, 33, M * D) and torch.isfinite(out).all()
    with pytest.raises(RuntimeError):
        call(1, 2, 2, 4, 3)                        # n_frame > T2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [0, 5, 16, 1000, 65536 + 7])
def test_masked_zero_inplace_equals_masked_fill(dtype, n):
    g = torch.Generator().manual_seed(n)
    data = torch.randn(n, generator=g).to(dtype).to(DEV)
    mask = (torch.rand(n, generator=g) < 0.3).to(DEV)
    want = data.masked_fill(mask, 0.0)
    torch.ops.snipper_b200.masked_zero_(data, mask)
    assert torch.equal(data, want)


def test_fused_module_with_padding_mask_matches_masked_fill_path():
    """The in-place masked value production + masked grad_value of the fused path against the per-call loop,
    which keeps the reference's out-of-place masked_fill (forward and every gradient)."""
    shapes = torch.as_tensor([(9, 12), (5, 6), (3, 3)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    ours, _ = _pair(384, 8, 3, 4, 4, "encoder", seed=21)
    loop = copy.deepcopy(ours)
    loop.fused = False
    g = torch.Generator().manual_seed(5)
    N, T = 2, 4
    q = torch.randn(N, T, S, 384, generator=g)
    src = torch.randn(N, T, S, 384, generator=g)
    refp = torch.rand(N, T, S, 3, 2, generator=g)
    pix = torch.rand(N, T, S, 1, generator=g) < 0.25
    mask = pix.expand(N, T, S, 384).contiguous()

    def run(mod):
        a = q.to(DEV).requires_grad_(True)
        b = src.to(DEV).requires_grad_(True)
        out = mod(a, refp.to(DEV), b, shapes.to(DEV), torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1])).to(DEV),
                  mask.to(DEV))
        out.square().sum().backward()
        return out.detach(), a.grad, b.grad, [p.grad.clone() for p in mod.parameters()]

    for m in (ours, loop):
        m.zero_grad(set_to_none=True)
    x = run(ours)
    y = run(loop)
    assert rel_err(x[0], y[0]) < 1e-5
    assert rel_err(x[1], y[1]) < 1e-4 and rel_err(x[2], y[2]) < 1e-4
    for gx, gy in zip(x[3], y[3]):
        assert rel_err(gx, gy) < 1e-4
    # the source rows under the padding mask feed nothing but the (zeroed) value: their gradient is exactly zero
    assert float(x[2][pix.expand_as(x[2]).to(DEV)].abs().max()) == 0.0
