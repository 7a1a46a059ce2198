"""Fused layer tails and the graphed runner (rank 4) on the GPU."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("rows,C", [(1, 128), (7, 384), (9875, 384), (240, 256), (33, 1024)])
@pytest.mark.parametrize("with_pos,with_bias", [(True, True), (False, True), (True, False)])
def test_layer_tail_kernel_vs_torch(rows, C, with_pos, with_bias):
    """out = LayerNorm(residual + (y + bias)), out_pos = out + pos  (deformable_transformer.py:204-205, :188-190)."""
    import snipper_b200  # noqa: F401
    g = torch.Generator().manual_seed(rows + C)
    y = torch.randn(rows, C, generator=g).to(DEV)
    res = (torch.randn(rows, C, generator=g) * 3 + 1).to(DEV)
    bias = torch.randn(C, generator=g).to(DEV) if with_bias else None
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    pos = torch.randn(rows, C, generator=g).to(DEV) if with_pos else None
    out, out_pos = torch.ops.snipper_b200.layer_tail(y, bias, res, gamma, beta, pos, 1e-5)
    x = res.double() + (y.double() + (bias.double() if with_bias else 0))
    want = F.layer_norm(x, (C,), gamma.double(), beta.double(), 1e-5)
    assert rel_err(out, want) < 1e-5
    if with_pos:
        assert rel_err(out_pos, want + pos.double()) < 1e-5
    else:
        assert out_pos.numel() == 0


def test_layer_tail_rejects_what_it_does_not_cover():
    import snipper_b200  # noqa: F401
    y = torch.randn(4, 100, device=DEV)
    v = torch.randn(100, device=DEV)
    with pytest.raises(RuntimeError):
        torch.ops.snipper_b200.layer_tail(y, None, y, v, v, None, 1e-5)          # 100 channels: not 128*k
    from snipper_b200 import ops
    a = torch.randn(4, 128, device=DEV, requires_grad=True)
    assert not ops.layer_tail_supported(a, a)                                     # autograd is recording: stock ops
    with torch.no_grad():
        assert ops.layer_tail_supported(a, a)


def _network(**kw):
    import snipper_b200
    from snipper_b200.harness.snipper_net import build_snipper
    torch.manual_seed(5)
    net = build_snipper(snipper_b200.MSDeformAttn, **kw)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "sampling_offsets" in n and n.endswith("weight"):
                p.normal_(0, 0.02)
            if "attention_weights" in n:
                p.normal_(0, 0.2)
    return net.to(DEV).eval()


@pytest.mark.parametrize("future", [0, 2])
def test_fused_layer_tails_reproduce_the_stock_layers(future):
    import snipper_b200
    from snipper_b200 import ops
    net = _network(num_frames=4, num_future_frames=future, enc_layers=3, dec_layers=2, num_queries=11)
    x = torch.rand(2, 12, 192, 256, device=DEV)
    with torch.no_grad():
        want, (_, w_refs, w_att) = net(x)
        assert snipper_b200.enable_fused_layer_tails(net) == 5
        assert net.transformer.encoder.analytic_reference_points        # reference points computed in-kernel from now on
        ops.STATS.reset()
        ops.STATS.timing = True
        got, (_, g_refs, g_att) = net(x)
        ops.STATS.timing = False
    tails = [e for e in ops.STATS.events if e[0] == "layer_tail"]
    assert len(tails) == 3 * 2 + 2 * 3                                            # 2 per encoder layer, 3 per decoder layer
    assert sum(e[1][2] for e in tails) == 2 + 2                                   # x + pos emitted for encoder layers 2, 3 and the decoder's cross-attention queries
    for k in ("pred_logits", "pred_kpts2d", "pred_depth"):
        assert rel_err(got[k], want[k]) < 1e-4, k                                 # whole network, two op orders
    for a, b in zip(got["heatmaps"], want["heatmaps"]):
        assert rel_err(a, b) < 1e-4
    assert rel_err(g_refs, w_refs) < 1e-4
    for (gl, ga), (wl, wa) in zip(g_att, w_att):                                  # the decoder's attention_vis payload
        for a, b in zip(gl, wl):
            assert rel_err(a, b) < 1e-4
    # training keeps the stock layers: gradients flow and match
    net.train()
    for p in net.parameters():
        p.grad = None
    ops.STATS.reset()
    ops.STATS.timing = True
    out, _ = net(x[:1])
    ops.STATS.timing = False
    assert not [e for e in ops.STATS.events if e[0] == "layer_tail"]
    assert snipper_b200.disable_fused_layer_tails(net) == 5


def test_graph_runner_replays_bit_identically_and_counts_launches():
    import snipper_b200
    net = _network(num_frames=4, num_future_frames=0, enc_layers=2, dec_layers=2, num_queries=7)
    snipper_b200.enable_fused_layer_tails(net)

    def fn(x):
        out, _ = net(x)
        return {k: out[k] for k in ("pred_logits", "pred_kpts2d", "pred_depth")}

    runner = snipper_b200.GraphRunner(fn)
    xs = [torch.rand(1, 12, 192, 256, device=DEV) for _ in range(3)]
    with torch.no_grad():
        eager = [{k: v.clone() for k, v in fn(x).items()} for x in xs]
    for _ in range(2):
        for x, want in zip(xs, eager):
            got = runner(x)
            for k in want:
                assert torch.equal(got[k], want[k]), k                            # same kernels, same order: same bits
    # 2 encoder layers: frame_sum + gather + 2 tails; 2 decoder layers: gather + 3 tails
    assert runner.launches_per_replay(xs[0]) == 2 * 4 + 2 * 4
    host = xs[0].cpu().pin_memory()
    got = runner(host)                                                            # pinned host input: H2D inside the call
    torch.cuda.synchronize()
    assert torch.equal(got["pred_kpts2d"], eager[0]["pred_kpts2d"])
    # input pipelining: the next snippet's H2D copy is staged on a side stream while the current one computes
    hosts = [x.cpu().pin_memory() for x in xs]
    runner.prefetch(hosts[0])
    for i in range(6):
        got = runner(hosts[i % 3])                                                # picks the staged copy up
        runner.prefetch(hosts[(i + 1) % 3])
        torch.cuda.synchronize()
        assert torch.equal(got["pred_kpts2d"], eager[i % 3]["pred_kpts2d"]), i
    got = runner(hosts[2])                                                        # not the staged tensor: direct copy
    torch.cuda.synchronize()
    assert torch.equal(got["pred_kpts2d"], eager[2]["pred_kpts2d"])
    other = torch.rand(2, 12, 128, 160, device=DEV)                               # a second resolution gets its own graph
    with torch.no_grad():
        want = fn(other)["pred_logits"].clone()
    assert torch.equal(runner(other)["pred_logits"], want)
