"""Full-size parity of the FUSED layer kernels at the exact launches bench.py times (BASELINE configs 2-4):

  encoder   N = 1 and 2, T = 4, levels (75,100),(38,50),(19,25), S = Lq = 9875, M = 8, D = 48, P = 4, padding mask
  decoder   T1 = 4 + 2 future frames, Lq = 60 (config 3)

against the C oracle composed per (t1,t2) as the reference module loops
(models/ops/modules/ms_deform_attn.py:130-225), fp32: forward <= 1e-5, gradients <= 1e-4 (max-abs relative,
BASELINE.json); bf16 I/O: 1e-2.  Both neighbour-frame strategies (pre-summed slots / direct gather) and both
mask layouts (the reference's materialised (N,T,S,C) tensor / one byte per pixel) are covered.
"""
import pytest
import torch

from conftest import level_start_index, rel_err
from snippet_oracle import snippet_attention_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LEVELS = [(75, 100), (38, 50), (19, 25)]
M, D, P = 8, 48, 4


def _encoder_reference_points(levels):
    refs = []
    for H, W in levels:
        ys, xs = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
    return torch.cat(refs, 0)  # (S, 2)


def _case(N, T1, T2, Lq, seed, encoder, sigma_px=3.0):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(LEVELS, dtype=torch.long)
    L = len(LEVELS)
    S = int(shapes.prod(1).sum())
    mlp = M * L * P
    value = torch.randn(N, T2, S, M, D, generator=g)
    proj = torch.cat((torch.randn(N, T1, Lq, 2 * mlp, generator=g) * sigma_px,     # offsets in pixels
                      torch.randn(N, T1, Lq, mlp, generator=g)), -1).contiguous()  # logits
    off_bias = torch.randn(2 * mlp, generator=g)
    logit_bias = torch.randn(mlp, generator=g) * 0.3
    if encoder:
        ref = _encoder_reference_points(LEVELS)[None, None, :, None, :].expand(N, T1, Lq, L, 2).contiguous()
    else:
        ref = torch.rand(N, T1, Lq, L, 2, generator=g) * 1.1 - 0.05
    pix = torch.rand(N, T2, S, generator=g) < 0.15                                   # padding mask, per pixel
    grad_out = torch.randn(N, T1, Lq, M * D, generator=g)
    return dict(seed=(seed, N, T1, T2, Lq), value=value, proj=proj, off_bias=off_bias, logit_bias=logit_bias, ref=ref, pix=pix,
                grad_out=grad_out, shapes=shapes, lsi=level_start_index(shapes))


_ORACLE_CACHE = {}


def _oracle(c, n_frame, dtype):
    """Oracle results are shared by the strategy / mask-layout variants of one case (seconds of CPU each)."""
    key = (c["seed"], n_frame, dtype)
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE.clear()                                       # keep one case resident
        _ORACLE_CACHE[key] = _oracle_run(c, n_frame, dtype)
    return _ORACLE_CACHE[key]


def _oracle_run(c, n_frame, dtype):
    value = c["value"].to(dtype).float().clone().requires_grad_(True)   # bf16 mode: the oracle sees the rounded inputs
    proj = c["proj"].clone().requires_grad_(True)
    ref = c["ref"].clone().requires_grad_(True)
    ob = c["off_bias"].clone().requires_grad_(True)
    lb = c["logit_bias"].clone().requires_grad_(True)
    out = snippet_attention_oracle(value, c["pix"], c["shapes"], c["lsi"], proj, ob, lb, ref, n_frame)
    out.backward(c["grad_out"].to(dtype).float())
    return [out.detach(), value.grad, proj.grad, ref.grad, ob.grad, lb.grad]


def _ours(c, n_frame, dtype, presum, per_pixel_mask, planar=False):
    from snipper_b200 import ops
    ops.set_planar_slots(planar)
    try:
        return _ours_run(c, n_frame, dtype, presum, per_pixel_mask)
    finally:
        ops.set_planar_slots(False)


def _ours_run(c, n_frame, dtype, presum, per_pixel_mask):
    from snipper_b200 import ops
    value = c["value"].detach().to(DEV, dtype).requires_grad_(True)
    proj = c["proj"].to(DEV).requires_grad_(True)
    ref = c["ref"].to(DEV).requires_grad_(True)
    ob = c["off_bias"].to(DEV).requires_grad_(True)
    lb = c["logit_bias"].to(DEV).requires_grad_(True)
    N, T2, S = c["pix"].shape
    pix = c["pix"].to(DEV)
    mask = pix if per_pixel_mask else pix[..., None].expand(N, T2, S, M * D).contiguous()
    out = ops.snippet_attention(value, mask, c["shapes"].to(DEV), c["lsi"].to(DEV), proj, ob, lb, ref, n_frame,
                                presum=presum)
    out.backward(c["grad_out"].to(DEV, dtype))
    torch.cuda.synchronize()
    return [out.detach(), value.grad, proj.grad, ref.grad, ob.grad, lb.grad]


def _compare(got, want, dtype, pix, presum=False):
    names = ["out", "grad_value", "grad_proj", "grad_ref", "grad_off_bias", "grad_logit_bias"]
    if dtype == torch.float32:
        tol = [1e-5, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4]
    elif presum:  # the neighbour-frame sums are stored in bf16 too: every output carries bf16-level error
        tol = [1e-2] * 6
    else:   # bf16 value / out / grad_out (1e-2); the fp32 tensors see bf16-rounded inputs on both sides
        tol = [1e-2, 1e-2, 1e-4, 1e-4, 1e-4, 1e-4]
    errs = {n: rel_err(g, w) for n, g, w in zip(names, got, want)}
    for n, t in zip(names, tol):
        assert errs[n] < t, (n, errs)
    # no gradient reaches padded value elements: exactly zero, not just small
    gv = got[1].float().cpu()
    assert float(gv[pix].abs().max()) == 0.0


@pytest.mark.parametrize("N", [1, 2])
@pytest.mark.parametrize("presum,per_pixel_mask,planar", [(True, False, False), (True, True, False), (True, True, True),
                                                          (False, True, False)])
def test_encoder_layer_full_size_fp32(N, presum, per_pixel_mask, planar):
    """The launch `roofline.kernel` names in bench.py (N=1) and the training launch (N=2)."""
    S = sum(h * w for h, w in LEVELS)
    c = _case(N, 4, 4, S, seed=10 + N, encoder=True)
    want = _oracle(c, 4, torch.float32)
    got = _ours(c, 4, torch.float32, presum, per_pixel_mask, planar)
    _compare(got, want, torch.float32, c["pix"])


@pytest.mark.parametrize("presum", [True, False])
def test_encoder_layer_full_size_bf16(presum):
    S = sum(h * w for h, w in LEVELS)
    c = _case(1, 4, 4, S, seed=21, encoder=True)
    want = _oracle(c, 4, torch.bfloat16)
    got = _ours(c, 4, torch.bfloat16, presum, True)
    _compare(got, want, torch.bfloat16, c["pix"], presum)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("presum,per_pixel_mask,planar", [(False, False, False), (False, True, False), (True, True, False),
                                                          (True, True, True)])
def test_decoder_layer_config3_forecasting(dtype, presum, per_pixel_mask, planar):
    """BASELINE config 3: T = 4 observed + 2 future query frames, 60 queries, all of `memory` as value.
    (presum + planar: the all-frames slot of the future query frames in the planar layout.)"""
    c = _case(1, 6, 4, 60, seed=33, encoder=False, sigma_px=6.0)
    want = _oracle(c, 4, dtype)
    got = _ours(c, 4, dtype, presum, per_pixel_mask, planar)
    _compare(got, want, dtype, c["pix"], presum)


@pytest.mark.parametrize("per_pixel_mask", [True, False])
def test_decoder_layer_bench_launch(per_pixel_mask):
    """The decoder launch bench.py times (BASELINE config 2: T1 = T2 = 4, 60 queries): the few-queries split kernel with
    the padding mask applied in the gather, forward and every gradient against the oracle."""
    from snipper_b200 import ops
    c = _case(1, 4, 4, 60, seed=34, encoder=False, sigma_px=6.0)
    want = _oracle(c, 4, torch.float32)
    ops.STATS.reset()
    ops.STATS.timing = True
    got = _ours(c, 4, torch.float32, None, per_pixel_mask)          # presum=None: the shapes pick the direct gather
    ops.STATS.timing = False
    assert sorted({e[0] for e in ops.STATS.events}) == ["snippet_backward", "snippet_forward"]
    _compare(got, want, torch.float32, c["pix"])


def test_auto_strategy_matches_the_shapes():
    from snipper_b200 import ops
    S = sum(h * w for h, w in LEVELS)
    assert ops.prefers_presum(4, 4, 4, S, 3, S, 4)            # encoder: sum the neighbour frames first
    assert not ops.prefers_presum(4, 6, 4, S, 3, 60, 4)       # decoder: gather the few samples directly
    assert not ops.prefers_presum(1, 1, 1, S, 3, S, 4)        # T = 1: nothing to sum


def test_presummed_value_is_the_masked_neighbour_sum():
    """msda_frame_sum / msda_frame_unsum against their torch definitions, incl. the all-frames slot."""
    from snipper_b200 import capi
    g = torch.Generator().manual_seed(5)
    N, T2, T1, n_frame, S, C = 2, 4, 6, 4, 37, 64
    value = torch.randn(N, T2, S, C, generator=g).to(DEV)
    pix = (torch.rand(N, T2, S, generator=g) < 0.3).to(DEV)
    full = pix[..., None].expand(N, T2, S, C).contiguous()
    masked = value.masked_fill(full, 0.0)
    slots = [masked[:, max(j - 1, 0):min(j + 1, n_frame - 1) + 1].sum(1) for j in range(n_frame)] + [masked.sum(1)]
    want = torch.stack(slots, 1)
    L_ = capi.lib()
    st = torch.cuda.current_stream().cuda_stream
    for mask, mrs, mcs in ((full, C, 1), (pix, 1, 0)):
        for dtype, code, tol in ((torch.float32, capi.MSDA_DTYPE_F32, 1e-6), (torch.bfloat16, capi.MSDA_DTYPE_BF16, 1e-2)):
            v = value.to(dtype)
            vsum = torch.empty(N, 5, S, C, device=DEV, dtype=dtype)
            assert L_.msda_frame_sum(v.data_ptr(), mask.data_ptr(), vsum.data_ptr(), N, T2, T1, n_frame, S, C, 0, 0,
                                     mrs, mcs, code, st) == 0
            assert rel_err(vsum, want) < tol
            gsum = torch.randn(N, 5, S, C, generator=g).to(DEV)
            gv = torch.empty(N, T2, S, C, device=DEV, dtype=dtype)
            assert L_.msda_frame_unsum(gsum.data_ptr(), mask.data_ptr(), gv.data_ptr(), N, T2, T1, n_frame, S, C,
                                       mrs, mcs, code, st) == 0
            want_gv = torch.stack([sum(gsum[:, j] for j in range(n_frame) if abs(j - t) <= 1) + gsum[:, 4]
                                   for t in range(T2)], 1).masked_fill(full, 0.0)
            assert rel_err(gv, want_gv) < tol
            assert float(gv.float()[full].abs().max()) == 0.0
    # fewer query frames than n_frame, more source frames than n_frame, no mask
    N, T2, T1, n_frame = 1, 5, 2, 3
    value = torch.randn(N, T2, S, C, generator=g).to(DEV)
    vsum = torch.empty(N, 2, S, C, device=DEV)
    assert L_.msda_frame_sum(value.data_ptr(), None, vsum.data_ptr(), N, T2, T1, n_frame, S, C, 0, 0, 0, 0,
                             capi.MSDA_DTYPE_F32, st) == 0
    assert rel_err(vsum[:, 0], value[:, 0:2].sum(1)) < 1e-6 and rel_err(vsum[:, 1], value[:, 0:3].sum(1)) < 1e-6
    gsum = torch.randn(N, 2, S, C, generator=g).to(DEV)
    gv = torch.full((N, T2, S, C), 7.0, device=DEV)
    assert L_.msda_frame_unsum(gsum.data_ptr(), None, gv.data_ptr(), N, T2, T1, n_frame, S, C, 0, 0,
                               capi.MSDA_DTYPE_F32, st) == 0
    want_gv = torch.stack([gsum[:, 0] + gsum[:, 1], gsum[:, 0] + gsum[:, 1], gsum[:, 1],
                           torch.zeros_like(gsum[:, 0]), torch.zeros_like(gsum[:, 0])], 1)
    assert rel_err(gv, want_gv) < 1e-6


def _planar_decode(buf, N, NS, S, M):
    """planar slots (N, NS, elems) -> (N, NS, S, M, 48) via plane A + the EVEN copy of plane B, and the odd copy
    separately (csrc/msda_planar.cu: A [m][s][32], Be [m][s][16], Bo [m][s+1][16], SB = (S + 3) & ~1 cells per head)."""
    SB = (S + 3) & ~1
    a_el, b_el = M * S * 32, M * SB * 16
    A = buf[..., :a_el].view(N, NS, M, S, 32)
    Be = buf[..., a_el:a_el + b_el].view(N, NS, M, SB, 16)[:, :, :, :S]
    Bo = buf[..., a_el + b_el:a_el + 2 * b_el].view(N, NS, M, SB, 16)[:, :, :, 1:S + 1]
    even = torch.cat((A, Be), -1).permute(0, 1, 3, 2, 4)
    odd = torch.cat((A, Bo), -1).permute(0, 1, 3, 2, 4)
    return even, odd


def test_planar_slots_are_the_masked_neighbour_sums_relaid():
    """msda_frame_sum_planar / msda_frame_unsum_planar against the torch definitions of the cell-major passes."""
    from snipper_b200 import capi
    g = torch.Generator().manual_seed(6)
    L_ = capi.lib()
    st = torch.cuda.current_stream().cuda_stream
    for (N, T2, T1, n_frame, S, M) in ((2, 4, 6, 4, 37, 8), (1, 5, 2, 3, 10, 2), (1, 1, 1, 1, 1, 1)):
        D, C = 48, M * 48
        NS = min(T1, n_frame) + (1 if T1 > n_frame else 0)
        elems = L_.msda_planar_slot_bytes(S, M, D, capi.MSDA_DTYPE_F32) // 4
        assert elems == M * (S * 32 + ((S + 3) & ~1) * 32)
        value = torch.randn(N, T2, S, C, generator=g).to(DEV)
        pix = (torch.rand(N, T2, S, generator=g) < 0.3).to(DEV)
        full = pix[..., None].expand(N, T2, S, C).contiguous()
        masked = value.masked_fill(full, 0.0)
        lo_hi = [(max(j - 1, 0), min(j + 1, n_frame - 1)) for j in range(min(T1, n_frame))] + ([(0, T2 - 1)] if T1 > n_frame else [])
        want = torch.stack([masked[:, lo:hi + 1].sum(1) for lo, hi in lo_hi], 1).view(N, NS, S, M, D)
        for mask, mrs, mcs in ((full, C, 1), (pix, 1, 0), (None, 0, 0)):
            w = want if mask is not None else torch.stack([value[:, lo:hi + 1].sum(1) for lo, hi in lo_hi], 1).view(N, NS, S, M, D)
            buf = torch.full((N, NS, elems), float("nan"), device=DEV)
            assert L_.msda_frame_sum_planar(value.data_ptr(), None if mask is None else mask.data_ptr(), buf.data_ptr(),
                                            N, T2, T1, n_frame, S, M, D, 0, 0, mrs, mcs, capi.MSDA_DTYPE_F32, st) == 0
            even, odd = _planar_decode(buf, N, NS, S, M)
            assert rel_err(even, w) < 1e-6 and torch.equal(even, odd)
            # gradient slots: the two plane-B copies are SUMMED back (the scatter writes either one)
            gbuf = torch.randn(N, NS, elems, generator=g).to(DEV)
            ge, go = _planar_decode(gbuf, N, NS, S, M)
            gslots = torch.cat((ge[..., :32], ge[..., 32:] + go[..., 32:]), -1)                     # (N,NS,S,M,48)
            gv = torch.full((N, T2, S, C), 7.0, device=DEV)
            assert L_.msda_frame_unsum_planar(gbuf.data_ptr(), None if mask is None else mask.data_ptr(), gv.data_ptr(),
                                              N, T2, T1, n_frame, S, M, D, mrs, mcs, capi.MSDA_DTYPE_F32, st) == 0
            want_gv = torch.stack([sum((gslots[:, j] for j, (lo, hi) in enumerate(lo_hi) if lo <= t <= hi),
                                       torch.zeros_like(gslots[:, 0])) for t in range(T2)], 1).reshape(N, T2, S, C)
            if mask is not None:
                want_gv = want_gv.masked_fill(full, 0.0)
            assert rel_err(gv, want_gv) < 1e-6
    # the layout only exists for fp32 heads of 48 channels
    assert L_.msda_planar_slot_bytes(100, 8, 32, capi.MSDA_DTYPE_F32) == 0
    assert L_.msda_planar_slot_bytes(100, 8, 48, capi.MSDA_DTYPE_BF16) == 0


# ------------------------------------------------------------------- deterministic mode of the fused layer
def _ours_deterministic(c, n_frame):
    import snipper_b200
    snipper_b200.set_deterministic(True)
    try:
        return _ours(c, n_frame, torch.float32, None, True)      # presum=None: the deterministic mode picks the slots
    finally:
        snipper_b200.set_deterministic(False)


def test_encoder_layer_full_size_deterministic_backward():
    """north_star: 'offers a deterministic two-pass mode' -- for the kernel training actually runs: bit-identical
    run to run at the full bench size, and equal to the oracle."""
    S = sum(h * w for h, w in LEVELS)
    c = _case(1, 4, 4, S, seed=11, encoder=True)
    a = _ours_deterministic(c, 4)
    b = _ours_deterministic(c, 4)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    _compare(a, _oracle(c, 4, torch.float32), torch.float32, c["pix"])


def test_decoder_layer_deterministic_backward_shares_the_all_frames_slot():
    """T1 = 4 + 2: both future query frames scatter into ONE slot; the ordered reduction merges them canonically."""
    c = _case(2, 6, 4, 60, seed=34, encoder=False, sigma_px=6.0)
    a = _ours_deterministic(c, 4)
    b = _ours_deterministic(c, 4)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    _compare(a, _oracle(c, 4, torch.float32), torch.float32, c["pix"])


def test_module_deterministic_mode_stays_on_the_fused_path():
    import snipper_b200
    from snipper_b200 import ops
    torch.manual_seed(2)
    shapes = torch.as_tensor([(19, 25), (10, 13), (5, 7)], dtype=torch.long, device=DEV)
    lsi = level_start_index(shapes.cpu()).to(DEV)
    S = int(shapes.prod(1).sum())
    mod = snipper_b200.MSDeformAttn(384, 3, 8, 4, 4, "encoder").to(DEV)
    with torch.no_grad():
        mod.sampling_offsets[0].weight.normal_(0, 0.05)
        mod.attention_weights[0].weight.normal_(0, 0.2)
    q = torch.randn(2, 4, S, 384, device=DEV)
    src = torch.randn(2, 4, S, 384, device=DEV)
    refp = torch.rand(2, 4, S, 3, 2, device=DEV)
    pad = torch.rand(2, 4, S, 1, device=DEV) < 0.1

    def run():
        for p in mod.parameters():
            p.grad = None
        a, b = q.clone().requires_grad_(True), src.clone().requires_grad_(True)
        ops.STATS.reset()
        ops.STATS.timing = True
        mod(a, refp, b, shapes, lsi, pad.expand(2, 4, S, 384)).square().sum().backward()
        ops.STATS.timing = False
        return [a.grad, b.grad] + [p.grad.clone() for p in mod.parameters()], sorted({e[0] for e in ops.STATS.events})

    want, tags0 = run()
    snipper_b200.set_deterministic(True)
    try:
        g1, tags = run()
        g2, _ = run()
    finally:
        snipper_b200.set_deterministic(False)
    assert tags == ["frame_sum", "frame_unsum", "snippet_backward_deterministic", "snippet_forward_presummed"], tags
    assert "snippet_backward_presummed" in tags0
    for x, y in zip(g1, g2):
        assert torch.equal(x, y)                       # bit-identical run to run
    for x, y in zip(g1, want):
        assert rel_err(x, y) < 1e-4                    # and the same gradients as the atomic path


def test_encoder_reference_points_in_kernel_are_bit_identical():
    """the encoder's reference points as a function of the query index (valid ratios in, no
    (N,T,S,L,2) tensor) reproduce get_reference_points (deformable_transformer.py:219-232) bit for bit -- forward,
    every gradient, atomic-free deterministic mode included."""
    import snipper_b200
    from snipper_b200 import ops
    from snipper_b200.modules import EncoderGrid
    S = sum(h * w for h, w in LEVELS)
    c = _case(2, 4, 4, S, seed=41, encoder=True)
    g = torch.Generator().manual_seed(9)
    vr = (torch.rand(2, 3, 2, generator=g) * 0.4 + 0.6).to(DEV)           # padded frames: valid ratios < 1
    grid = EncoderGrid(vr, LEVELS, 4)
    ref = grid.tensor().contiguous()
    value = c["value"].to(DEV)
    proj, ob, lb = c["proj"].to(DEV), c["off_bias"].to(DEV), c["logit_bias"].to(DEV)
    shapes, lsi, pix, go = c["shapes"].to(DEV), c["lsi"].to(DEV), c["pix"].to(DEV), c["grad_out"].to(DEV)

    def run(reference_points, valid_ratios, deterministic):
        snipper_b200.set_deterministic(deterministic)
        try:
            v, p = value.clone().requires_grad_(True), proj.clone().requires_grad_(True)
            out = ops.snippet_attention(v, pix, shapes, lsi, p, ob, lb, reference_points, 4, valid_ratios=valid_ratios)
            out.backward(go)
            return out.detach(), v.grad, p.grad
        finally:
            snipper_b200.set_deterministic(False)

    a = run(ref, None, True)
    b = run(None, vr, True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    a = run(ref, None, False)
    b = run(None, vr, False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])           # grad_value differs only by atomic ordering
    assert rel_err(b[1], a[1]) < 1e-5
