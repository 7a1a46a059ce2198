"""Probe the planar kernels on small problems, one subprocess per case with a hard timeout (a hung kernel must not
take the GPU job down).  usage: python tools/debug/planar_probe.py [case ...]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

CASES = {
    "full_tiles_fwd": dict(levels=[(8, 8)], M=1, N=1, T=1, Lq=64, bwd=False),
    "partial_tile_fwd": dict(levels=[(8, 8)], M=1, N=1, T=1, Lq=19, bwd=False),
    "multi_level_fwd": dict(levels=[(9, 7), (5, 4), (3, 2)], M=2, N=1, T=4, Lq=85, bwd=False),
    "multi_level_bwd": dict(levels=[(9, 7), (5, 4), (3, 2)], M=2, N=1, T=4, Lq=85, bwd=True),
    "encoder_fwd": dict(levels=[(75, 100), (38, 50), (19, 25)], M=8, N=1, T=4, Lq=9875, bwd=False),
    "encoder_bwd": dict(levels=[(75, 100), (38, 50), (19, 25)], M=8, N=1, T=4, Lq=9875, bwd=True),
}


def run(name):
    import torch
    from snipper_b200 import ops
    c = CASES[name]
    g = torch.Generator().manual_seed(1)
    shapes = torch.as_tensor(c["levels"], dtype=torch.long)
    L, P, D, M, N, T, Lq = len(c["levels"]), 4, 48, c["M"], c["N"], c["T"], c["Lq"]
    S = int(shapes.prod(1).sum())
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    mlp = M * L * P
    value = torch.randn(N, T, S, M, D, generator=g).cuda()
    proj = torch.cat((torch.randn(N, T, Lq, 2 * mlp, generator=g) * 2.0, torch.randn(N, T, Lq, mlp, generator=g)), -1).cuda()
    ref = torch.rand(N, T, Lq, L, 2, generator=g).cuda()
    go = torch.randn(N, T, Lq, M * D, generator=g).cuda()
    outs = []
    for planar in (False, True):
        ops.set_planar_slots(planar)
        v, p = value.clone().requires_grad_(c["bwd"]), proj.clone().requires_grad_(c["bwd"])
        out = ops.snippet_attention(v, None, shapes.cuda(), lsi.cuda(), p, None, None, ref, T, presum=True)
        if c["bwd"]:
            out.backward(go)
        torch.cuda.synchronize()
        outs.append([out.detach()] + ([v.grad, p.grad] if c["bwd"] else []))
        print(name, "planar" if planar else "cell-major", "done", flush=True)
    for a, b in zip(*outs):
        print(name, "max abs diff %.3e (max %.3e)" % (float((a - b).abs().max()), float(a.abs().max())), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run(sys.argv[2])
    else:
        for name in (sys.argv[1:] or list(CASES)):
            try:
                r = subprocess.run([sys.executable, __file__, "--one", name], timeout=60, capture_output=True, text=True)
                print(r.stdout.strip(), ("\nSTDERR: " + r.stderr.strip()[-800:]) if r.returncode else "", flush=True)
            except subprocess.TimeoutExpired as e:
                print(name, "TIMEOUT (hang); partial output:", (e.stdout or b"").decode()[-300:] if isinstance(e.stdout, bytes) else e.stdout, flush=True)
