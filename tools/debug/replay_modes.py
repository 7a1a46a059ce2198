"""How should back-to-back CUDA-graph replays of the bench network be issued?  Times one snippet per step with inputs
resident: (a) replays enqueued back to back, (b) a stream synchronize after every replay, (c) an event wait every 2nd
replay.  python tools/debug/replay_modes.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import snipper_b200
    from snipper_b200.harness.snipper_net import build_snipper
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = build_snipper(snipper_b200.MSDeformAttn).to(dev).eval()
    snipper_b200.enable_fused_layer_tails(model)
    runner = snipper_b200.GraphRunner(lambda x: model(x)[0]["pred_kpts2d"])
    xs = [torch.rand(1, 12, 600, 800, device=dev) for _ in range(4)]
    for x in xs:
        runner(x)
    torch.cuda.synchronize()

    def timed(mode, steps=40):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for i in range(steps):
            runner(xs[i % 4])
            if mode == "sync_every_step":
                torch.cuda.current_stream().synchronize()
            elif mode == "sync_every_2" and i % 2 == 1:
                torch.cuda.current_stream().synchronize()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / steps

    out = {}
    for rep in range(2):
        for mode in ("back_to_back", "sync_every_step", "sync_every_2"):
            out.setdefault(mode, []).append(round(timed(mode), 3))
    print(json.dumps({"ms_per_step": out}))


if __name__ == "__main__":
    main()
