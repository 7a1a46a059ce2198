"""Seeded shape fuzzing of the CUDA path against the C oracle: random level pyramids, head counts, channel
widths (fast and generic kernels), point counts, ragged query counts, batch strides, out-of-range samples.
Forward <= 1e-5, backward <= 1e-4 relative (fp32), as everywhere."""
import os
import random

import pytest
import torch

from conftest import level_start_index, rel_err
from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# MSDA_FUZZ_SCALE=k multiplies the number of seeds of every fuzz test (soak runs; default 1)
_SCALE = max(int(os.environ.get("MSDA_FUZZ_SCALE", "1")), 1)


def seeds(n):
    return range(n * _SCALE)


def _random_case(rng):
    L = rng.choice([1, 2, 3, 4, 6])
    shapes = torch.as_tensor([(rng.randint(1, 14), rng.randint(1, 14)) for _ in range(L)], dtype=torch.long)
    M = rng.choice([1, 2, 3, 8])
    D = rng.choice([4, 8, 16, 20, 32, 48, 48, 64, 100, 128])
    P = rng.choice([1, 2, 4, 5, 8])
    N = rng.choice([1, 2, 3])
    Lq = rng.choice([1, 7, 16, 17, 60, 131])
    S = int(shapes.prod(1).sum())
    g = torch.Generator().manual_seed(rng.randint(0, 1 << 30))
    value = torch.randn(N, S, M, D, generator=g)
    spread = rng.choice([0.2, 1.0, 3.0])            # > 1: a good share of the samples falls outside [0,1]
    loc = 0.5 + (torch.rand(N, Lq, M, L, P, 2, generator=g) - 0.5) * (1.0 + spread)
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    go = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value, shapes=shapes, lsi=level_start_index(shapes), loc=loc, attn=attn, grad_out=go,
                strided=rng.random() < 0.3)


@pytest.mark.parametrize("seed", seeds(24))
def test_random_shapes_per_call(seed):
    from snipper_b200 import MSDeformAttnFunction
    rng = random.Random(1000 + seed)
    c = _random_case(rng)
    v = c["value"].to(DEV)
    if c["strided"] and v.shape[0] > 1:             # value[:, t] of a frame-stacked tensor: batch stride != S*M*D
        big = torch.zeros(v.shape[0], 2, *v.shape[1:], device=DEV)
        big[:, 1] = v
        v = big[:, 1]
    v = v.detach().requires_grad_(True)
    s = c["loc"].to(DEV).requires_grad_(True)
    a = c["attn"].to(DEV).requires_grad_(True)
    out = MSDeformAttnFunction.apply(v, c["shapes"].to(DEV), c["lsi"].to(DEV), s, a, 64)
    out.backward(c["grad_out"].to(DEV))
    args = (c["value"].double(), c["shapes"], c["lsi"], c["loc"].double(), c["attn"].double())
    ref_out = c_oracle.forward(*args)
    ref = c_oracle.backward(*args, c["grad_out"].double())
    assert rel_err(out, ref_out) < 1e-5
    for got, want in zip((v.grad, s.grad, a.grad), ref):
        assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("seed", seeds(8))
def test_random_shapes_fused_vs_per_call(seed):
    """Fused per-layer op on random geometry against T1 x |neighbours| per-call launches (oracle-checked above)."""
    rng = random.Random(2000 + seed)
    L = rng.choice([1, 2, 3])
    shapes = torch.as_tensor([(rng.randint(2, 10), rng.randint(2, 10)) for _ in range(L)], dtype=torch.long, device=DEV)
    lsi = level_start_index(shapes.cpu()).to(DEV)
    S = int(shapes.prod(1).sum())
    M, D, P = rng.choice([(8, 48, 4), (4, 32, 2), (2, 64, 8), (8, 16, 4)])
    N, n_frame, fut = rng.choice([1, 2]), rng.choice([1, 2, 4]), rng.choice([0, 1, 2])
    T2, T1, Lq = n_frame, n_frame + fut, rng.choice([1, 9, 33])
    g = torch.Generator().manual_seed(seed)
    value = torch.randn(N, T2, S, M, D, generator=g).to(DEV)
    off = (torch.randn(N, T1, Lq, M, L, P, 2, generator=g) * 2.0).to(DEV)
    logits = torch.randn(N, T1, Lq, M, L, P, generator=g).to(DEV)
    ref = (torch.rand(N, T1, Lq, L, 2, generator=g) * 1.2 - 0.1).to(DEV)
    got = torch.ops.snipper_b200.snippet_forward(value, shapes, lsi, off, logits, ref, n_frame)
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
    loc = ref[:, :, :, None, :, None, :] + off / wh[None, None, None, None, :, None, :]
    att = torch.softmax(logits.flatten(-2), -1).view(N, T1, Lq, M, L, P)
    want = torch.zeros_like(got)
    for t1 in range(T1):
        nb = [t for t in (t1 - 1, t1, t1 + 1) if 0 <= t < n_frame] if t1 < n_frame else list(range(T2))
        for t2 in nb:
            want[:, t1] += torch.ops.snipper_b200.msda_forward(value[:, t2], shapes, lsi, loc[:, t1].contiguous(),
                                                               (att[:, t1] / len(nb)).contiguous(), 64)
    assert rel_err(got, want) < 1e-5


def _composition(value, shapes, lsi, off, logits, ref, n_frame):
    """The fused op written out with per-call launches and torch glue (differentiable)."""
    from snipper_b200 import MSDeformAttnFunction
    N, T2 = value.shape[:2]
    _, T1, Lq, M, L, P, _ = off.shape
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(off.dtype)
    loc = ref[:, :, :, None, :, None, :] + off / wh[None, None, None, None, :, None, :]
    att = torch.softmax(logits.flatten(-2), -1).view(N, T1, Lq, M, L, P)
    outs = []
    for t1 in range(T1):
        nb = [t for t in (t1 - 1, t1, t1 + 1) if 0 <= t < n_frame] if t1 < n_frame else list(range(T2))
        acc = 0
        for t2 in nb:
            acc = acc + MSDeformAttnFunction.apply(value[:, t2], shapes, lsi, loc[:, t1].contiguous(),
                                                   (att[:, t1] / len(nb)).contiguous(), 64)
        outs.append(acc)
    return torch.stack(outs, 1)


@pytest.mark.parametrize("seed", seeds(8))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_random_shapes_fused_backward_vs_composition(seed, dtype):
    rng = random.Random(3000 + seed)
    L = rng.choice([1, 2, 3])
    shapes = torch.as_tensor([(rng.randint(2, 9), rng.randint(2, 9)) for _ in range(L)], dtype=torch.long, device=DEV)
    lsi = level_start_index(shapes.cpu()).to(DEV)
    S = int(shapes.prod(1).sum())
    M, D, P = rng.choice([(8, 48, 4), (4, 32, 2), (2, 64, 8), (8, 16, 4), (2, 128, 8)])
    N, n_frame, fut = rng.choice([1, 2]), rng.choice([1, 2, 4]), rng.choice([0, 1, 2])
    T2, T1, Lq = n_frame, n_frame + fut, rng.choice([1, 9, 33])
    g = torch.Generator().manual_seed(100 + seed)
    value = torch.randn(N, T2, S, M, D, generator=g).to(dtype)
    off = torch.randn(N, T1, Lq, M, L, P, 2, generator=g) * 2.0
    logits = torch.randn(N, T1, Lq, M, L, P, generator=g)
    ref = torch.rand(N, T1, Lq, L, 2, generator=g) * 1.2 - 0.1
    go = torch.randn(N, T1, Lq, M * D, generator=g).to(dtype)

    def run(fn):
        leaves = [t.to(DEV).requires_grad_(True) for t in (value, off, logits, ref)]
        out = fn(leaves[0], shapes, lsi, leaves[1], leaves[2], leaves[3], n_frame)
        out.backward(go.to(DEV))
        return [out.detach()] + [t.grad for t in leaves]

    a = run(lambda *x: torch.ops.snipper_b200.snippet_forward(*x))
    b = run(_composition)
    ftol, vtol = (1e-5, 1e-4) if dtype == torch.float32 else (1e-2, 1e-2)
    assert rel_err(a[0], b[0]) < ftol
    assert rel_err(a[1], b[1]) < vtol            # grad_value
    for i in (2, 3, 4):                          # grad_offsets, grad_logits, grad_reference_points (fp32 in both modes)
        assert rel_err(a[i], b[i]) < (1e-4 if dtype == torch.float32 else 2e-2), i


@pytest.mark.parametrize("seed", seeds(10))
def test_random_shapes_deterministic_backward(seed):
    """Deterministic mode on random geometry: equals the oracle and is bit-identical run to run."""
    rng = random.Random(4000 + seed)
    c = _random_case(rng)
    dev_args = [c[k].to(DEV) for k in ("value", "shapes", "lsi", "loc", "attn", "grad_out")]
    a = torch.ops.snipper_b200.msda_backward(*dev_args, 64, True)
    b = torch.ops.snipper_b200.msda_backward(*dev_args, 64, True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    args = (c["value"].double(), c["shapes"], c["lsi"], c["loc"].double(), c["attn"].double())
    for got, want in zip(a, c_oracle.backward(*args, c["grad_out"].double())):
        assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("seed", seeds(10))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_random_shapes_packed_layer_strategies_and_masks_vs_oracle(seed, dtype):
    """The packed per-layer op on random geometry -- pre-summed neighbour frames and direct gather, per-channel
    and per-pixel padding masks, biases in-kernel -- against the C oracle composed per (t1,t2)."""
    from snipper_b200 import ops
    from snippet_oracle import snippet_attention_oracle
    rng = random.Random(5000 + seed)
    L = rng.choice([1, 2, 3])
    shapes = torch.as_tensor([(rng.randint(2, 9), rng.randint(2, 9)) for _ in range(L)], dtype=torch.long)
    lsi = level_start_index(shapes)
    S = int(shapes.prod(1).sum())
    M, D, P = rng.choice([(8, 48, 4), (4, 32, 2), (2, 64, 8), (8, 16, 4), (2, 128, 8)])
    N, n_frame, fut = rng.choice([1, 2]), rng.choice([1, 2, 3, 4]), rng.choice([0, 0, 1, 2])
    T2 = n_frame + rng.choice([0, 0, 1])                 # source frames beyond n_frame only feed future queries
    T1, Lq = n_frame + fut, rng.choice([1, 9, 33, S])
    mlp = M * L * P
    g = torch.Generator().manual_seed(200 + seed)
    value = torch.randn(N, T2, S, M, D, generator=g).to(dtype)
    proj = torch.cat((torch.randn(N, T1, Lq, 2 * mlp, generator=g) * 2.0, torch.randn(N, T1, Lq, mlp, generator=g)), -1)
    ob, lb = torch.randn(2 * mlp, generator=g), torch.randn(mlp, generator=g)
    ref = torch.rand(N, T1, Lq, L, 2, generator=g) * 1.2 - 0.1
    pix = torch.rand(N, T2, S, generator=g) < 0.2
    go = torch.randn(N, T1, Lq, M * D, generator=g).to(dtype)

    leaves = [t.clone().requires_grad_(True) for t in (value.float(), proj, ob, lb, ref)]
    want_out = snippet_attention_oracle(leaves[0], pix, shapes, lsi, leaves[1], leaves[2], leaves[3], leaves[4], n_frame)
    want_out.backward(go.float())
    want = [want_out.detach()] + [t.grad for t in leaves]

    ftol, vtol, stol = (1e-5, 1e-4, 1e-4) if dtype == torch.float32 else (1e-2, 1e-2, 1e-4)
    # planar slots exist for fp32 D = 48 only; elsewhere the switch changes nothing and the variant is skipped
    variants = [(True, False), (False, False)] + ([(True, True)] if (D == 48 and dtype == torch.float32) else [])
    for presum, planar in variants:
        ops.set_planar_slots(planar)
        for per_pixel in (True, False):
            lv = [t.to(DEV).requires_grad_(True) for t in (value, proj, ob, lb, ref)]
            mask = pix.to(DEV) if per_pixel else pix[..., None].expand(N, T2, S, M * D).contiguous().to(DEV)
            out = ops.snippet_attention(lv[0], mask, shapes.to(DEV), lsi.to(DEV), lv[1], lv[2], lv[3], lv[4], n_frame,
                                        presum=presum)
            out.backward(go.to(DEV))
            got = [out.detach()] + [t.grad for t in lv]
            tag = (presum, planar, per_pixel)
            assert rel_err(got[0], want[0]) < ftol, tag
            assert rel_err(got[1], want[1]) < vtol, tag
            # bf16 + pre-summed: the slot sums are themselves stored in bf16, so the fp32 gradients carry bf16-level error
            st = 1e-2 if (presum and dtype == torch.bfloat16) else stol
            for i in (2, 3, 4, 5):
                assert rel_err(got[i], want[i]) < st, (tag, i)
            assert not pix.any() or float(got[1].float().cpu()[pix].abs().max()) == 0.0
    ops.set_planar_slots(False)


@pytest.mark.parametrize("seed", seeds(12))
def test_random_shapes_planar_slots_vs_oracle(seed):
    """The planar-slot kernels (csrc/msda_planar.cu; fp32, D = 48) on random geometry: odd widths (pairs that start on
    odd cells read the shifted plane-B copy), one-pixel-wide levels (every sample takes the border path), far offsets,
    the all-frames slot of future query frames, in-kernel encoder reference points -- against the C oracle composed per
    (t1,t2), and bit-identical forward between two runs."""
    from snipper_b200 import ops
    from snipper_b200.modules import EncoderGrid
    from snippet_oracle import snippet_attention_oracle
    rng = random.Random(7000 + seed)
    L = rng.choice([1, 2, 3, 4])
    sizes = [(rng.randint(1, 11), rng.randint(1, 11)) for _ in range(L)]
    shapes = torch.as_tensor(sizes, dtype=torch.long)
    lsi = level_start_index(shapes)
    S = int(shapes.prod(1).sum())
    M, D, P = rng.choice([1, 2, 3, 8]), 48, rng.choice([1, 2, 4, 8])
    if (M * L * P) % 2:          # the packed projection row is read as float2: M*L*P must be even (ops.snippet_supported)
        M += 1
    if L * P == 1:               # a softmax over one logit has an identically zero gradient: nothing to compare against
        P = 2
    N, n_frame, fut = rng.choice([1, 2]), rng.choice([1, 2, 3, 4]), rng.choice([0, 0, 1, 2])
    T2 = n_frame + rng.choice([0, 0, 1])
    T1 = n_frame + fut
    encoder = seed % 3 == 0                       # queries = pixels, reference points computed in-kernel
    Lq = S if encoder else rng.choice([1, 7, 40, 65])
    mlp = M * L * P
    g = torch.Generator().manual_seed(300 + seed)
    value = torch.randn(N, T2, S, M, D, generator=g)
    proj = torch.cat((torch.randn(N, T1, Lq, 2 * mlp, generator=g) * rng.choice([0.5, 2.0, 6.0]),
                      torch.randn(N, T1, Lq, mlp, generator=g)), -1)
    ob, lb = torch.randn(2 * mlp, generator=g), torch.randn(mlp, generator=g)
    pix = torch.rand(N, T2, S, generator=g) < 0.2
    go = torch.randn(N, T1, Lq, M * D, generator=g)
    vr = None
    if encoder:
        vr = torch.rand(N, L, 2, generator=g) * 0.4 + 0.6
        ref = EncoderGrid(vr, sizes, T1).tensor().contiguous()
    else:
        ref = torch.rand(N, T1, Lq, L, 2, generator=g) * 1.2 - 0.1

    leaves = [t.clone().requires_grad_(True) for t in (value, proj, ob, lb, ref)]
    want_out = snippet_attention_oracle(leaves[0], pix, shapes, lsi, leaves[1], leaves[2], leaves[3], leaves[4], n_frame)
    want_out.backward(go)
    want = [want_out.detach()] + [t.grad for t in leaves]

    outs = []
    ops.set_planar_slots(True)
    for rep in range(2):
        ops.STATS.reset()
        ops.STATS.timing = True
        lv = [t.to(DEV).requires_grad_(True) for t in (value, proj, ob, lb, ref)]
        out = ops.snippet_attention(lv[0], pix.to(DEV), shapes.to(DEV), lsi.to(DEV), lv[1], lv[2], lv[3],
                                    None if encoder else lv[4], n_frame, presum=True,
                                    valid_ratios=vr.to(DEV) if encoder else None)
        out.backward(go.to(DEV))
        ops.STATS.timing = False
        assert sorted(e[0] for e in ops.STATS.events) == ["frame_sum_planar", "frame_unsum_planar",
                                                          "snippet_backward_planar", "snippet_forward_planar"]
        got = [out.detach()] + [t.grad for t in lv]
        outs.append(got[0])
        assert rel_err(got[0], want[0]) < 1e-5
        assert rel_err(got[1], want[1]) < 1e-4
        for i in (2, 3, 4) + (() if encoder else (5,)):
            assert rel_err(got[i], want[i]) < 1e-4, i
        assert not pix.any() or float(got[1].cpu()[pix].abs().max()) == 0.0
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("planar", [True, False])
def test_packed_layer_with_32_samples_per_query(planar):
    """L*P = 32, the most the fused kernels take: the planar backward then needs exactly 48 KB of dynamic shared memory
    on top of its static level table -- a launch that failed until the opt-in threshold counted the static part
    (found by the MSDA_FUZZ_SCALE=8 soak run)."""
    from snipper_b200 import ops
    from snippet_oracle import snippet_attention_oracle
    g = torch.Generator().manual_seed(77)
    sizes = [(6, 7), (5, 4), (3, 3), (2, 5)]
    shapes = torch.as_tensor(sizes, dtype=torch.long)
    lsi = level_start_index(shapes)
    S = int(shapes.prod(1).sum())
    N, T, M, D, L, P, Lq = 1, 3, 2, 48, 4, 8, 37
    mlp = M * L * P
    value = torch.randn(N, T, S, M, D, generator=g)
    proj = torch.cat((torch.randn(N, T, Lq, 2 * mlp, generator=g) * 2.0, torch.randn(N, T, Lq, mlp, generator=g)), -1)
    ob, lb = torch.randn(2 * mlp, generator=g), torch.randn(mlp, generator=g)
    ref = torch.rand(N, T, Lq, L, 2, generator=g)
    pix = torch.rand(N, T, S, generator=g) < 0.2
    go = torch.randn(N, T, Lq, M * D, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (value, proj, ob, lb, ref)]
    want_out = snippet_attention_oracle(leaves[0], pix, shapes, lsi, leaves[1], leaves[2], leaves[3], leaves[4], T)
    want_out.backward(go)
    want = [want_out.detach()] + [t.grad for t in leaves]
    ops.set_planar_slots(planar)
    lv = [t.to(DEV).requires_grad_(True) for t in (value, proj, ob, lb, ref)]
    out = ops.snippet_attention(lv[0], pix.to(DEV), shapes.to(DEV), lsi.to(DEV), lv[1], lv[2], lv[3], lv[4], T, presum=True)
    out.backward(go.to(DEV))
    got = [out.detach()] + [t.grad for t in lv]
    assert rel_err(got[0], want[0]) < 1e-5
    for i in range(1, 6):
        assert rel_err(got[i], want[i]) < 1e-4, i
