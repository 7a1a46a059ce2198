"""CPU tests of the host-side logic: C-ABI surface, error behaviour without a GPU, module
contract (state-dict keys, init), custom-op fake kernels, and the N>1 path on gloo (world 2)."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import snipper_b200
from snipper_b200 import capi, sharding
from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    declared = set(re.findall(r"MSDA_API\s+[\w\s\*]*?\b(msda_\w+)\s*\(", header))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert capi.lib().msda_abi_version() == capi.MSDA_ABI_VERSION
    assert b"im2col_step" in capi.lib().msda_error_string(capi.MSDA_ERR_IM2COL_STEP)


def test_argument_validation_needs_no_gpu():
    """Validation happens before any launch, so these calls are safe on a CPU-only box."""
    L = capi.lib()
    # null pointers
    assert L.msda_forward(0, 0, 0, 0, 0, 0, 1, 4, 2, 16, 1, 3, 2, 0, 64, 0, 0) == capi.MSDA_ERR_INVALID_ARGUMENT
    # im2col_step: batch 3, step 2 -> 3 % 2 != 0 (reference ms_deform_attn_cuda.cu:50-52)
    assert L.msda_forward(8, 8, 8, 8, 8, 8, 3, 4, 2, 16, 1, 3, 2, 0, 2, 0, 0) == capi.MSDA_ERR_IM2COL_STEP
    # bad dtype tag
    assert L.msda_forward(8, 8, 8, 8, 8, 8, 1, 4, 2, 16, 1, 3, 2, 0, 64, 7, 0) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    # empty problems are a no-op
    assert L.msda_forward(0, 0, 0, 0, 0, 0, 0, 4, 2, 16, 1, 3, 2, 0, 64, 0, 0) == capi.MSDA_OK
    assert not hasattr(L, "msda_set_tuning")   # stateless ABI: no entry point writes process state
    assert L.msda_backward_workspace_bytes(1, 100, 8, 48, 3, 100, 4, 0, 0) == 0
    assert L.msda_backward_workspace_bytes(1, 100, 8, 48, 3, 100, 4, 0, capi.MSDA_FLAG_DETERMINISTIC) > 0


def test_fused_and_mask_entry_points_validate_without_a_gpu():
    L = capi.lib()
    F32, BF16, F64 = capi.MSDA_DTYPE_F32, capi.MSDA_DTYPE_BF16, capi.MSDA_DTYPE_F64
    # msda_masked_zero: empty is a no-op, null pointers / bad dtype / misaligned mask are rejected
    assert L.msda_masked_zero(0, 0, 0, F32, 0) == capi.MSDA_OK
    assert L.msda_masked_zero(0, 0, 16, F32, 0) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert L.msda_masked_zero(256, 256, 16, F64, 0) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert L.msda_masked_zero(256, 257, 16, F32, 0) == capi.MSDA_ERR_INVALID_ARGUMENT
    # msda_snippet_forward(value, shapes, lsi, offsets, logits, ref, out, N,T2,T1,n_frame,S,M,D,L,Lq,P, strides x6,
    #                      biases x2, encoder valid ratios, mask, mask row / col stride, dtype, flags, stream)
    def fwd(N=1, T2=4, T1=4, n_frame=4, S=100, M=8, D=48, Lv=3, Lq=10, P=4, ors=0, lrs=0, dtype=F32, ptr=256,
            mask=None, mrs=0, mcs=0, flags=0, vr=None):
        return L.msda_snippet_forward(ptr, ptr, ptr, ptr, ptr, ptr, ptr, N, T2, T1, n_frame, S, M, D, Lv, Lq, P,
                                      0, 0, 0, 0, ors, lrs, None, None, vr, mask, mrs, mcs, dtype, flags, 0)
    assert fwd(N=0) == capi.MSDA_OK and fwd(Lq=0) == capi.MSDA_OK          # empty problems: nothing is launched
    assert fwd(n_frame=5) == capi.MSDA_ERR_INVALID_ARGUMENT                 # n_frame > T2
    assert fwd(D=40) == capi.MSDA_ERR_INVALID_ARGUMENT                      # D % 16 != 0
    assert fwd(Lv=9, P=4) == capi.MSDA_ERR_INVALID_ARGUMENT                 # L*P > 32
    assert fwd(dtype=F64) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert fwd(ors=8 * 3 * 4 * 2 - 2) == capi.MSDA_ERR_INVALID_ARGUMENT     # rows would overlap
    assert fwd(ors=8 * 3 * 4 * 3 + 1) == capi.MSDA_ERR_INVALID_ARGUMENT     # odd row stride breaks the float2 loads
    assert fwd(ptr=0) == capi.MSDA_ERR_INVALID_ARGUMENT                     # null pointers with work to do
    # padding mask: per-channel masks are read as 32-bit words; no mask together with presummed value
    assert fwd(N=0, mask=257, mrs=384, mcs=1) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert fwd(N=0, mask=256, mrs=386, mcs=1) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert fwd(N=0, mask=256, mrs=384, mcs=2) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert fwd(N=0, mask=257, mrs=1, mcs=0) == capi.MSDA_OK
    assert fwd(N=0, mask=256, mrs=384, mcs=1, flags=capi.MSDA_FLAG_PRESUMMED) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert fwd(N=0, flags=capi.MSDA_FLAG_PRESUMMED) == capi.MSDA_OK
    assert fwd(N=0, flags=64) == capi.MSDA_ERR_INVALID_ARGUMENT
    # in-kernel encoder reference points: the queries must be the pixels of the pyramid (Lq == S)
    assert fwd(N=0, vr=256, Lq=100) == capi.MSDA_OK and fwd(N=0, vr=256, Lq=10) == capi.MSDA_ERR_INVALID_ARGUMENT
    # deterministic mode of the fused layer: pre-summed float32 only, needs its workspace
    def bwd(flags, dtype=F32, ws=None, ws_bytes=0):
        return L.msda_snippet_backward(256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 1, 4, 4, 4, 100, 8, 48, 3, 10, 4,
                                       0, 0, 0, 0, 0, 0, None, None, None, None, 0, 0, dtype, flags, ws, ws_bytes, 0)
    DET, PRE = capi.MSDA_FLAG_DETERMINISTIC, capi.MSDA_FLAG_PRESUMMED
    assert bwd(DET) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert bwd(DET | PRE, dtype=BF16) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert bwd(DET | PRE) == capi.MSDA_ERR_WORKSPACE
    need = L.msda_snippet_backward_workspace_bytes(1, 4, 4, 100, 8, 48, 3, 10, 4, F32, DET | PRE)
    assert need > 0 and L.msda_snippet_backward_workspace_bytes(1, 4, 4, 100, 8, 48, 3, 10, 4, F32, PRE) == 0
    assert bwd(DET | PRE, ws=256, ws_bytes=need - 1) == capi.MSDA_ERR_WORKSPACE
    assert bwd(DET | PRE, ws=257, ws_bytes=need) == capi.MSDA_ERR_WORKSPACE
    # neighbour-frame pre-summation: slot structure, strategy choice, validation
    assert L.msda_snippet_num_slots(4, 4) == 4 and L.msda_snippet_num_slots(6, 4) == 5 and L.msda_snippet_num_slots(2, 4) == 2
    assert L.msda_snippet_prefers_presum(4, 4, 4, 9875, 3, 9875, 4) == 1     # encoder: 10 frame pairs -> 4 gathers
    assert L.msda_snippet_prefers_presum(4, 6, 4, 9875, 3, 60, 4) == 0       # decoder: 60 queries
    assert L.msda_snippet_prefers_presum(1, 1, 1, 9875, 3, 9875, 4) == 0     # T = 1: nothing to sum
    def fsum(N=1, T2=4, T1=4, n_frame=4, S=100, C=384, dtype=F32, ptr=256, mask=None, mrs=0, mcs=0):
        return L.msda_frame_sum(ptr, mask, ptr, N, T2, T1, n_frame, S, C, 0, 0, mrs, mcs, dtype, 0)
    assert fsum(N=0) == capi.MSDA_OK
    assert fsum(N=0, C=6) == capi.MSDA_ERR_INVALID_ARGUMENT                   # rows are walked in 16-byte chunks
    assert fsum(N=0, C=12, dtype=BF16) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert fsum(N=0, n_frame=5) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert fsum(N=0, dtype=F64) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert fsum(ptr=0) == capi.MSDA_ERR_INVALID_ARGUMENT
    # planar slots (opt-in): the layout exists for float32 heads of 48 channels; it rides on the pre-summed path only
    PLN = capi.MSDA_FLAG_PLANAR
    assert L.msda_planar_slot_bytes(9875, 8, 48, F32) == 8 * (9875 * 128 + 9878 * 128)   # plane A + two copies of plane B
    assert L.msda_planar_slot_bytes(9875, 8, 32, F32) == 0 and L.msda_planar_slot_bytes(9875, 8, 48, BF16) == 0
    assert L.msda_planar_slot_bytes(0, 8, 48, F32) == 0 and L.msda_planar_slot_bytes(1 << 22, 8, 48, F32) == 0   # 32-bit offsets
    assert fwd(N=0, flags=PRE | PLN) == capi.MSDA_OK
    assert fwd(N=0, flags=PLN) == capi.MSDA_ERR_INVALID_ARGUMENT              # planar without pre-summed slots
    assert fwd(N=0, flags=PRE | PLN, dtype=BF16) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert fwd(N=0, flags=PRE | PLN, D=32) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert fwd(flags=PRE | PLN, ptr=320) == capi.MSDA_ERR_INVALID_ARGUMENT     # slots must be 128-byte aligned
    assert bwd(PRE | PLN | DET) == capi.MSDA_ERR_INVALID_ARGUMENT              # the deterministic mode walks cell-major slots
    def fsum_pl(N=1, T2=4, T1=4, n_frame=4, S=100, M=8, D=48, dtype=F32, ptr=256, mask=None, mrs=0, mcs=0):
        return L.msda_frame_sum_planar(ptr, mask, ptr, N, T2, T1, n_frame, S, M, D, 0, 0, mrs, mcs, dtype, 0)
    assert fsum_pl(N=0) == capi.MSDA_OK
    assert fsum_pl(N=0, D=64) == capi.MSDA_ERR_UNSUPPORTED_DTYPE and fsum_pl(N=0, dtype=BF16) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert fsum_pl(N=0, n_frame=5) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert fsum_pl(ptr=0) == capi.MSDA_ERR_INVALID_ARGUMENT and fsum_pl(ptr=336) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert L.msda_frame_unsum_planar(256, None, 256, 0, 4, 4, 4, 100, 8, 48, 0, 0, F32, 0) == capi.MSDA_OK
    assert L.msda_frame_unsum_planar(336, None, 256, 1, 4, 4, 4, 100, 8, 48, 0, 0, F32, 0) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert L.msda_frame_unsum(0, None, 0, 0, 4, 4, 4, 100, 384, 0, 0, F32, 0) == capi.MSDA_OK
    assert L.msda_frame_unsum(0, None, 0, 1, 4, 4, 4, 100, 384, 0, 0, F32, 0) == capi.MSDA_ERR_INVALID_ARGUMENT
    # layer tail: float32 rows of 128*k <= 1024 channels; pos and its output come together
    tail = lambda rows=4, cols=384, dtype=F32, pos=None, out2=None, ptr=256: L.msda_layer_tail(
        ptr, ptr, ptr, ptr, ptr, pos, ptr, out2, rows, cols, 1e-5, dtype, 0)
    assert tail(rows=0) == capi.MSDA_OK
    assert tail(cols=100) == capi.MSDA_ERR_INVALID_ARGUMENT and tail(cols=2048) == capi.MSDA_ERR_INVALID_ARGUMENT
    assert tail(dtype=BF16) == capi.MSDA_ERR_UNSUPPORTED_DTYPE
    assert tail(pos=256) == capi.MSDA_ERR_INVALID_ARGUMENT and tail(ptr=0) == capi.MSDA_ERR_INVALID_ARGUMENT
    # deterministic mode: 32-bit corner ids -> a clear error instead of a wrong answer
    assert L.msda_backward(256, 256, 256, 256, 256, 256, 256, 256, 256, 64, 9875, 8, 48, 12, 9875 * 4, 8, 0, 64, F32,
                           capi.MSDA_FLAG_DETERMINISTIC, 256, 1 << 40, 0) == capi.MSDA_ERR_TOO_LARGE
    # bf16 needs D % 16 == 0 in the per-call entry points too
    assert L.msda_forward(256, 256, 256, 256, 256, 256, 1, 4, 2, 24, 1, 3, 2, 0, 64, BF16, 0) == capi.MSDA_ERR_UNSUPPORTED_DTYPE


def test_cpu_tensors_raise_like_the_reference():
    shim = snipper_b200.install_extension_shim()
    args = (torch.zeros(1, 4, 2, 16), torch.tensor([[2, 2]]), torch.tensor([0]),
            torch.zeros(1, 3, 2, 1, 2, 2), torch.zeros(1, 3, 2, 1, 2))
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):  # ms_deform_attn.h:38
        shim.ms_deform_attn_forward(*args, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        shim.ms_deform_attn_backward(*args, torch.zeros(1, 3, 32), 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        snipper_b200.MSDeformAttnFunction.apply(*args, 64)
    import MultiScaleDeformableAttention as by_name  # importable under the reference's module name
    assert by_name is shim


def test_fake_kernels_give_shapes_without_a_device():
    v = torch.empty(2, 50, 8, 48, device="meta")
    sh = torch.empty(3, 2, dtype=torch.long, device="meta")
    lsi = torch.empty(3, dtype=torch.long, device="meta")
    loc = torch.empty(2, 7, 8, 3, 4, 2, device="meta")
    att = torch.empty(2, 7, 8, 3, 4, device="meta")
    out = torch.ops.snipper_b200.msda_forward(v, sh, lsi, loc, att, 64)
    assert out.shape == (2, 7, 384)
    gv, gl, ga = torch.ops.snipper_b200.msda_backward(v, sh, lsi, loc, att, out, 64, False)
    assert gv.shape == v.shape and gl.shape == loc.shape and ga.shape == att.shape
    v5 = torch.empty(2, 4, 50, 8, 48, device="meta")
    off = torch.empty(2, 6, 7, 8, 3, 4, 2, device="meta")
    lg = torch.empty(2, 6, 7, 8, 3, 4, device="meta")
    ref = torch.empty(2, 6, 7, 3, 2, device="meta")
    assert torch.ops.snipper_b200.snippet_forward(v5, sh, lsi, off, lg, ref, 4).shape == (2, 6, 7, 384)
    proj = torch.empty(2, 6, 7, 3 * 8 * 3 * 4, device="meta")
    out, carry = torch.ops.snipper_b200.snippet_attn(v5, None, sh, lsi, proj, None, None, ref, None, 4, True, True)
    # fp32, D = 48: planar slots (4 frame slots + the all-frames slot), plane A 32 + planes Be / Bo 16 + 16 channels
    assert out.shape == (2, 6, 7, 384) and carry.shape == (2, 5, 8 * (50 * 32 + 52 * 32))
    gv, gp = torch.ops.snipper_b200.snippet_attn_backward(carry, None, sh, lsi, proj, None, None, ref, None, out, 4, True, 4,
                                                          False, [50, 8, 48])
    assert gv.shape == v5.shape and gp.shape == proj.shape
    out, carry = torch.ops.snipper_b200.snippet_attn(v5, None, sh, lsi, proj, None, None, ref, None, 4, True)
    assert out.shape == (2, 6, 7, 384) and carry.shape == (2, 5, 50, 8, 48)      # cell-major slots
    out, carry = torch.ops.snipper_b200.snippet_attn(v5, None, sh, lsi, proj, None, None, ref, None, 4, False)
    assert out.shape == (2, 6, 7, 384) and carry.numel() == 0
    gv, gp = torch.ops.snipper_b200.snippet_attn_backward(v5, None, sh, lsi, proj, None, None, ref, None, out, 4, False, 4,
                                                          False)
    assert gv.shape == v5.shape and gp.shape == proj.shape


def test_mask_layout_accepts_reference_and_per_pixel_masks():
    from snipper_b200.ops import mask_layout
    N, T2, S, C = 2, 3, 5, 8
    full = torch.zeros(N, T2, S, C, dtype=torch.bool)
    m, rs, cs = mask_layout(full, N, T2, S, C)
    assert (rs, cs) == (C, 1) and m.data_ptr() == full.data_ptr()               # the reference's materialised mask
    pix = torch.zeros(N, T2, S, 1, dtype=torch.bool)
    m, rs, cs = mask_layout(pix.expand(N, T2, S, C), N, T2, S, C)
    assert (rs, cs) == (1, 0)                                                   # one byte per pixel
    assert mask_layout(pix, N, T2, S, C)[1:] == (1, 0) and mask_layout(pix[..., 0], N, T2, S, C)[1:] == (1, 0)
    odd = torch.zeros(N, T2, S, C + 1, dtype=torch.bool)[..., :C]               # row stride 9: not word-aligned
    m, rs, cs = mask_layout(odd, N, T2, S, C)
    assert (rs, cs) == (C, 1) and m.is_contiguous()
    assert mask_layout(None, N, T2, S, C) == (None, 0, 0)
    with pytest.raises(RuntimeError):
        mask_layout(torch.zeros(N, T2, S, C), N, T2, S, C)                      # not bool
    with pytest.raises(RuntimeError):
        mask_layout(torch.zeros(N, T2, S + 1, C, dtype=torch.bool), N, T2, S, C)


def test_module_contract_matches_reference_checkpoint_layout():
    g = load_golden("module_decoder")
    want = sorted(k[3:] for k in g if k.startswith("sd."))
    mod = snipper_b200.MSDeformAttn(48, 3, 4, 4, 4, "decoder", False, True)
    assert sorted(mod.state_dict().keys()) == want
    assert all(m is mod.sampling_offsets[0] for m in mod.sampling_offsets)  # aliased slots (:68-71)
    # initialisation (reference :78-97): zero weights, per-head direction grid scaled by (p+1)
    assert mod.sampling_offsets[0].weight.abs().max() == 0 and mod.attention_weights[0].bias.abs().max() == 0
    b = mod.sampling_offsets[0].bias.view(4, 3, 4, 2)
    assert torch.allclose(b[0, :, :, 0], torch.tensor([1., 2., 3., 4.]).expand(3, 4))
    assert torch.allclose(b[1, 0, 2], torch.tensor([0., 3.]), atol=1e-6)
    assert snipper_b200.modules.neighbour_frames(0, 4, 4) == [0, 1]
    assert snipper_b200.modules.neighbour_frames(2, 4, 4) == [1, 2, 3]
    assert snipper_b200.modules.neighbour_frames(5, 4, 4) == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        snipper_b200.MSDeformAttn(50, 3, 8, 4)


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in range(*sharding.shard_range(n, r, world))]
            assert got == list(range(n))
            sizes = [b - a for a, b in (sharding.shard_range(n, r, world) for r in range(world))]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(9, rank, world)
    elapsed = 1.0 + rank  # rank 1 is the straggler
    thr = sharding.aggregate_throughput(hi - lo, elapsed)
    mx = sharding.max_over_ranks(elapsed)
    ranges = [None] * world
    dist.all_gather_object(ranges, (lo, hi))
    q.put((rank, thr, mx, ranges))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, thr, mx, ranges in res:
        assert mx == 2.0                      # slowest rank
        assert abs(thr - 9 / 2.0) < 1e-12     # all items / slowest time
        assert ranges == [(0, 5), (5, 9)]
