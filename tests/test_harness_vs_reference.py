"""The bench harness network (snipper_b200/harness) is a faithful restatement of the reference
model: load the REAL reference's state_dict into it and compare outputs on CPU.  Runs only in
the build container (skipped where /root/reference is absent)."""
import pytest
import torch

import ref_loader
from conftest import rel_err
from oracle import torch_ref

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("num_future", [0, 2])
def test_harness_matches_reference_model(num_future):
    from snipper_b200.harness.snipper_net import build_snipper
    torch.manual_seed(0)
    kw = dict(hidden_dim=96, num_frames=2, num_future_frames=num_future, enc_layers=1, dec_layers=2,
              num_queries=5, dim_feedforward=64, dropout=0.0)
    ref_model, args = ref_loader.build_reference_model(use_pytorch_deform=1, **kw)
    ref_model.eval()
    with torch.no_grad():  # leave the degenerate init so the attention actually depends on the queries
        for n, p in ref_model.named_parameters():
            if "sampling_offsets" in n and n.endswith("weight"):
                p.normal_(0, 0.05)
            if "attention_weights" in n:
                p.normal_(0, 0.2)
    mine = build_snipper(torch_ref.SnippetMSDeformAttnRef, **kw).eval()
    missing, unexpected = mine.load_state_dict(ref_model.state_dict(), strict=True)
    assert not missing and not unexpected
    x = torch.rand(1, 3 * 2, 96, 128)
    with torch.no_grad():
        want, (w_ref0, w_refs, w_att) = ref_model(x)
        got, (g_ref0, g_refs, g_att) = mine(x)
    for k in ("pred_logits", "pred_kpts2d", "pred_depth"):
        assert rel_err(got[k], want[k]) < 1e-5, k
    for a, b in zip(got["heatmaps"], want["heatmaps"]):
        assert rel_err(a, b) < 1e-5
    for a, b in zip(got["aux_outputs"], want["aux_outputs"]):
        for k in a:
            assert rel_err(a[k], b[k]) < 1e-5
    assert rel_err(g_refs, w_refs) < 1e-5
    assert len(g_att) == len(w_att)
