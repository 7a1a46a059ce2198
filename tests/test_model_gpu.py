"""End-to-end parity on the GPU: the Snipper-shaped harness network with the B200 attention module
vs the same network with the CPU oracle attention (identical weights), forward and backward."""
import pytest
import torch

from conftest import rel_err
from oracle import torch_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def strict_fp32():
    """cuDNN runs convolutions in TF32 by default; the CPU oracle network is strict fp32."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _models(**kw):
    import snipper_b200
    from snipper_b200.harness.snipper_net import build_snipper
    torch.manual_seed(5)
    ours = build_snipper(snipper_b200.MSDeformAttn, **kw)
    with torch.no_grad():  # leave the degenerate init (floor() sensitivity)
        for n, p in ours.named_parameters():
            if "sampling_offsets" in n and n.endswith("weight"):
                p.normal_(0, 0.02)
            if "attention_weights" in n:
                p.normal_(0, 0.2)
    oracle = build_snipper(torch_ref.SnippetMSDeformAttnRef, **kw)
    oracle.load_state_dict(ours.state_dict(), strict=True)
    return ours.to(DEV), oracle


@pytest.mark.parametrize("future", [0, 2])
def test_inference_matches_cpu_oracle_network(future):
    kw = dict(num_frames=2, num_future_frames=future, enc_layers=2, dec_layers=2, num_queries=7, dropout=0.0)
    ours, oracle = _models(**kw)
    ours.eval(), oracle.eval()
    x = torch.rand(2, 6, 96, 128)
    with torch.no_grad():
        got, _ = ours(x.to(DEV))
        want, _ = oracle(x)
    for k in ("pred_logits", "pred_kpts2d", "pred_depth"):
        assert rel_err(got[k], want[k]) < 2e-4, k   # 4 attention layers + ResNet-50 in fp32 on two devices
    for a, b in zip(got["heatmaps"], want["heatmaps"]):
        assert rel_err(a, b) < 2e-4


def test_training_step_gradients_match_cpu_oracle_network():
    kw = dict(num_frames=2, num_future_frames=0, enc_layers=1, dec_layers=1, num_queries=5, dropout=0.0)
    ours, oracle = _models(**kw)
    ours.train(), oracle.train()
    x = torch.rand(1, 6, 96, 128)

    def loss_of(model, inp):
        out, _ = model(inp)
        return (out["pred_kpts2d"] ** 2).mean() + out["pred_logits"].mean() + sum((h ** 2).mean() for h in out["heatmaps"])

    loss_of(ours, x.to(DEV)).backward()
    loss_of(oracle, x).backward()
    checked = 0
    for (n, p), (_, q) in zip(ours.named_parameters(), oracle.named_parameters()):
        if q.grad is None:
            assert p.grad is None, n
            continue
        assert p.grad is not None, n   # every trainable parameter gets a gradient (DDP needs no unused-param search)
        if q.grad.abs().max() > 1e-8:
            # backbone conv gradients go through cuDNN's fp32 algorithms (vs CPU direct conv): looser
            tol = 3e-2 if n.startswith("backbone.") else 5e-3
            assert rel_err(p.grad, q.grad) < tol, n
            checked += 1
    assert checked > 50


def test_cuda_graph_replay_is_bitwise_stable():
    import snipper_b200
    from snipper_b200.harness.snipper_net import build_snipper
    torch.manual_seed(1)
    model = build_snipper(snipper_b200.MSDeformAttn, num_frames=2, enc_layers=1, dec_layers=1, num_queries=5).to(DEV).eval()
    x = torch.rand(1, 6, 96, 128, device=DEV)
    with torch.no_grad():
        eager, _ = model(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        static = x.clone()
        with torch.cuda.graph(g):
            out, _ = model(static)
        static.copy_(x)
        g.replay()
        torch.cuda.synchronize()
    assert torch.equal(out["pred_kpts2d"], eager["pred_kpts2d"])
