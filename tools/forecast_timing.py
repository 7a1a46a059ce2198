"""BASELINE config 3: Snipper T=4+2 forecasting (num_future_frames=2), enc6/dec6, 60 queries x 15 keypoints,
batch 1, synthetic 600x800 snippet -- forward time per snippet on one GPU (CUDA events, eager and CUDA graph).
Sharding over GPUs is the same no-collective scheme as bench.py.   python tools/forecast_timing.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import snipper_b200
    from snipper_b200 import ops
    from snipper_b200.harness.snipper_net import build_snipper
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    out = {}
    for fut in (0, 2):
        model = build_snipper(snipper_b200.MSDeformAttn, num_future_frames=fut).to(dev).eval()
        snipper_b200.enable_fused_layer_tails(model)     # as bench.py runs the network
        x = torch.rand(1, 12, 600, 800, device=dev)
        with torch.no_grad():
            for _ in range(3):
                model(x)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y, _ = model(x)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                g.replay()
            s.record()
            for _ in range(10):
                g.replay()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            ops.STATS.reset()
            ops.STATS.timing = True
            model(x)
            torch.cuda.synchronize()
            ops.STATS.timing = False
            msda = sum(sum(v) for v in ops.STATS.kernel_ms().values())
        out["T4+%d" % fut] = {"ms_per_snippet_graph": round(ms, 3), "snippets_per_s": round(1e3 / ms, 2),
                             "msda_ms": round(msda, 3), "msda_launches": ops.STATS.launches,
                             "pred_kpts2d": list(y["pred_kpts2d"].shape)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
