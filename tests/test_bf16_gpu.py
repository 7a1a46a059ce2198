"""GPU parity tests of the bf16 I/O mode (MSDA_DTYPE_BF16): value / output / grad_output are bf16,
locations, weights, every gradient accumulation and all arithmetic are fp32.

Tolerance (BASELINE.json north_star): forward within 1e-2 relative in bf16.  The oracle is
evaluated in fp64 on the SAME bf16-rounded inputs, so the only error sources are fp32 accumulation
and the final rounding of the bf16 outputs (2^-9 relative per element); the fp32 outputs
(grad_sampling_loc, grad_attn_weight) are held to the fp32 backward tolerance 1e-4.
"""
import pytest
import torch

from conftest import make_case, rel_err
from oracle import c_oracle
from test_msda_gpu import cuda_fwd_bwd

pytestmark = pytest.mark.gpu

BF16_TOL, F32_BWD_TOL = 1e-2, 1e-4
LEVELS = [(19, 25), (10, 13), (5, 7)]


def _bf16_case(N, M, D, P, Lq, regime, seed):
    c = make_case(N, M, D, LEVELS, P, Lq=Lq, regime=regime, seed=seed)
    c["value"] = c["value"].bfloat16()
    c["grad_out"] = c["grad_out"].bfloat16()
    return c


@pytest.mark.parametrize("D", [16, 32, 48, 64, 96, 128])
@pytest.mark.parametrize("regime", ["uniform", "local"])
def test_bf16_percall_vs_oracle(D, regime):
    c = _bf16_case(2, 8 if D == 48 else 3, D, 4, 157, regime, seed=D)
    out, (gv, gl, ga) = cuda_fwd_bwd(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"], c["grad_out"])
    assert out.dtype == torch.bfloat16 and gv.dtype == torch.bfloat16
    assert gl.dtype == torch.float32 and ga.dtype == torch.float32
    args = (c["value"].double(), c["shapes"], c["lsi"], c["loc"].double(), c["attn"].double())
    ref_out = c_oracle.forward(*args)
    ref_gv, ref_gl, ref_ga = c_oracle.backward(*args, c["grad_out"].double())
    assert rel_err(out, ref_out) < BF16_TOL
    assert rel_err(gv, ref_gv) < BF16_TOL
    assert rel_err(gl, ref_gl) < F32_BWD_TOL
    assert rel_err(ga, ref_ga) < F32_BWD_TOL


def test_bf16_wide_heads_many_samples():
    """D = 128 with L*P = 32 samples per query: the backward needs more than 48 KB of dynamic shared memory."""
    c = make_case(1, 2, 128, [(9, 12), (5, 6), (3, 3), (2, 2)], 8, Lq=37, regime="uniform", seed=9)
    c["value"], c["grad_out"] = c["value"].bfloat16(), c["grad_out"].bfloat16()
    out, (gv, gl, ga) = cuda_fwd_bwd(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"], c["grad_out"])
    args = (c["value"].double(), c["shapes"], c["lsi"], c["loc"].double(), c["attn"].double())
    ref_gv, ref_gl, ref_ga = c_oracle.backward(*args, c["grad_out"].double())
    assert rel_err(out, c_oracle.forward(*args)) < BF16_TOL
    assert rel_err(gv, ref_gv) < BF16_TOL and rel_err(gl, ref_gl) < F32_BWD_TOL and rel_err(ga, ref_ga) < F32_BWD_TOL


def test_bf16_output_is_the_rounded_fp32_result():
    """Same inputs through the fp32 kernels: the bf16 kernel's output is that result rounded once."""
    c = _bf16_case(1, 8, 48, 4, None, "local", seed=5)
    out16, _ = cuda_fwd_bwd(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    out32, _ = cuda_fwd_bwd(c["value"].float(), c["shapes"], c["lsi"], c["loc"], c["attn"])
    # identical up to fp32 summation order (8 vs 4 channels per lane do not interact) -> at most 1 bf16 ulp
    assert rel_err(out16.float(), out32.bfloat16().float()) < 2.0 ** -7


def test_bf16_rejects_unsupported_shapes():
    c = _bf16_case(1, 2, 24, 2, 9, "uniform", seed=1)   # D % 16 != 0
    with pytest.raises(RuntimeError):
        cuda_fwd_bwd(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    c = _bf16_case(1, 2, 32, 2, 9, "uniform", seed=1)
    with pytest.raises(RuntimeError):                      # locations must be fp32
        cuda_fwd_bwd(c["value"], c["shapes"], c["lsi"], c["loc"].bfloat16(), c["attn"])


@pytest.mark.parametrize("mode,fut,Lq", [("encoder", 0, None), ("decoder", 2, 23)])
def test_bf16_fused_snippet_vs_fp32_kernels(mode, fut, Lq):
    """Fused per-layer op with a bf16 value tensor against the (oracle-verified) fp32 fused op run
    on the same, up-cast inputs."""
    dev = "cuda:0"
    g = torch.Generator().manual_seed(11)
    shapes = torch.as_tensor(LEVELS, dtype=torch.long)
    S = int(shapes.prod(1).sum())
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    N, T2, M, D, L, P = 2, 4, 8, 48, 3, 4
    T1, Lq = T2 + fut, (Lq or S)
    value = torch.randn(N, T2, S, M, D, generator=g).bfloat16()
    offsets = torch.randn(N, T1, Lq, M, L, P, 2, generator=g) * 2.5
    logits = torch.randn(N, T1, Lq, M, L, P, generator=g)
    ref = torch.rand(N, T1, Lq, L, 2, generator=g) * 1.1 - 0.05
    go = torch.randn(N, T1, Lq, M * D, generator=g).bfloat16()

    def run(v, gout):
        v = v.to(dev).requires_grad_(True)
        o = offsets.to(dev).requires_grad_(True)
        z = logits.to(dev).requires_grad_(True)
        out = torch.ops.snipper_b200.snippet_forward(v, shapes.to(dev), lsi.to(dev), o, z, ref.to(dev), T2)
        out.backward(gout.to(dev))
        torch.cuda.synchronize()
        return out.detach(), v.grad, o.grad, z.grad

    a = run(value, go)
    b = run(value.float(), go.float())
    assert a[0].dtype == torch.bfloat16 and a[1].dtype == torch.bfloat16
    assert rel_err(a[0], b[0]) < BF16_TOL
    assert rel_err(a[1], b[1]) < BF16_TOL
    assert rel_err(a[2], b[2]) < F32_BWD_TOL
    assert rel_err(a[3], b[3]) < F32_BWD_TOL


def test_module_under_autocast_uses_bf16_kernels():
    """torch.autocast(bfloat16): the Linear layers emit bf16, the fused kernel gathers bf16."""
    from snipper_b200 import MSDeformAttn, ops
    dev = "cuda:0"
    torch.manual_seed(3)
    shapes = torch.as_tensor(LEVELS, dtype=torch.long, device=dev)
    S = int(shapes.prod(1).sum())
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    mod = MSDeformAttn(384, 3, 8, 4, 4, "encoder").to(dev)
    with torch.no_grad():
        mod.sampling_offsets[0].weight.normal_(0, 0.02)
        mod.attention_weights[0].weight.normal_(0, 0.05)
    q = torch.randn(1, 4, S, 384, device=dev)
    src = torch.randn(1, 4, S, 384, device=dev)
    refp = torch.rand(1, 4, S, 3, 2, device=dev)
    want = mod(q, refp, src, shapes, lsi)
    ops.STATS.reset()
    ops.STATS.timing = True
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = mod(q, refp, src, shapes, lsi)
    ops.STATS.timing = False
    # encoder-sized query set: one streaming pass that sums the neighbour frames + one gather launch, nothing else
    assert ops.STATS.launches == 2
    assert sorted(tag for tag, _, _, _ in ops.STATS.events) == ["frame_sum", "snippet_forward_presummed"]
    assert rel_err(got.float(), want) < 3e-2   # bf16 GEMMs on both sides of the op dominate this error
