"""Op-level timing: B200 kernels vs the reference's vendored CUDA op (oracle/_ref, recompiled
for sm_100a) on Snipper shapes.  CUDA events over back-to-back launches; L2 is flushed between
launches when --flush is given.  Prints one JSON line per (shape, pass).

    python tools/opbench.py [--iters 50] [--flush] [--ref]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LEVELS = [(75, 100), (38, 50), (19, 25)]
PEAK = 6457.1
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def algorithmic_bytes(N, S, M, D, L, P, Lq, e=4):
    """SURVEY.md section 8d (value terms capped at the bytes actually touched for small Lq)."""
    v = min(N * S * M * D, 4 * N * Lq * M * L * P * D)
    fwd = e * (v + 3 * N * Lq * M * L * P + N * Lq * M * D)
    bwd = e * (2 * v + N * Lq * M * D + 6 * N * Lq * M * L * P)
    return fwd, bwd


def make(N, Lq, M=8, D=48, P=4, levels=LEVELS, regime="local", device="cuda", seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(levels, dtype=torch.long)
    L = len(levels)
    S = int(shapes.prod(1).sum())
    value = torch.randn(N, S, M, D, generator=g)
    if regime == "uniform" or Lq != S:
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    elif regime == "init":
        # what a randomly initialised Snipper produces: pixel centres + a fixed per-head direction
        # times (p+1) px (reference ms_deform_attn.py:82-87) -- best-case locality
        import math
        refs = []
        for H, W in levels:
            ys, xs = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
            refs.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
        ref = torch.cat(refs, 0)
        th = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        d = torch.stack([th.cos(), th.sin()], -1)
        d = d / d.abs().max(-1, keepdim=True)[0]
        off = d.view(1, 1, M, 1, 1, 2) * torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, 1, 1, P, 1)
        wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
        loc = (ref[None, :, None, None, None, :] + off / wh[None, None, None, :, None, :]).expand(N, Lq, M, L, P, 2).contiguous()
    else:  # encoder-like: pixel centres + N(0, 3px)
        refs = []
        for H, W in levels:
            ys, xs = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
            refs.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
        ref = torch.cat(refs, 0)
        wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
        loc = ref[None, :, None, None, None, :] + torch.randn(N, Lq, M, L, P, 2, generator=g) * 3.0 / wh[None, None, None, :, None, :]
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    go = torch.randn(N, Lq, M * D, generator=g)
    return [t.to(device) for t in (value, shapes, lsi, loc, attn, go)]


WARMUP = 5


INNER = 10


def time_fn(fn, iters, flush, inner=None):
    """Per-launch device time in us.  Without --flush: `inner` back-to-back launches between two
    events (host launch overhead hidden behind the queue).  With --flush: a 256 MiB memset evicts
    L2 before every single launch; the memset also keeps the queue busy, hiding host overhead."""
    inner = inner or INNER
    buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(WARMUP):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        if buf is not None:
            buf.zero_()
            s.record()
            fn()
            e.record()
            n = 1
        else:
            s.record()
            for _ in range(inner):
                fn()
            e.record()
            n = inner
        e.synchronize()
        times.append(s.elapsed_time(e) * 1e3 / n)
    times.sort()
    return times[len(times) // 2], times[0]


def direct_calls(value, shapes, lsi, loc, attn, go, bf16=False):
    """Forward/backward closures that call the C ABI directly (ctypes), bypassing the torch
    custom-op dispatcher (~30 us of host time per call, which would hide a 20 us kernel)."""
    from snipper_b200 import capi
    L = capi.lib()
    N, S, M, D = value.shape
    _, Lq, _, Lv, P, _ = loc.shape
    dt = 2 if bf16 else 0
    if bf16:
        value, go = value.bfloat16(), go.bfloat16()
    out = torch.empty(N, Lq, M * D, device=value.device, dtype=value.dtype)
    gv, gl, ga = torch.empty(value.shape, device=value.device), torch.empty_like(loc), torch.empty_like(attn)
    st = torch.cuda.current_stream().cuda_stream
    keep = (out, gv, gl, ga, value, go)

    def fwd():
        r = L.msda_forward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                           out.data_ptr(), N, S, M, D, Lv, Lq, P, 0, 64, dt, st)
        assert r == 0, r

    def bwd():
        r = L.msda_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                            go.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), N, S, M, D, Lv, Lq, P,
                            0, 64, dt, 0, 0, 0, st)
        assert r == 0, r

    def bwd_nomemset():
        r = L.msda_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                            go.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), N, S, M, D, Lv, Lq, P,
                            0, 64, dt, capi.MSDA_FLAG_ACCUMULATE_VALUE, 0, 0, st)
        assert r == 0, r

    if bf16:
        return fwd, bwd, bwd_nomemset, None, keep

    ws_bytes = L.msda_backward_workspace_bytes(N, S, M, D, Lv, Lq, P, 0, capi.MSDA_FLAG_DETERMINISTIC)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=value.device)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256

    def bwd_det():
        r = L.msda_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                            go.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), N, S, M, D, Lv, Lq, P,
                            0, 64, 0, capi.MSDA_FLAG_DETERMINISTIC, ws_ptr, ws_bytes, st)
        assert r == 0, r

    return fwd, bwd, bwd_nomemset, bwd_det, keep + (ws,)


def snippet_calls(N, T1, T2, Lq, n_frame=4, M=8, D=48, P=4, levels=LEVELS, encoder=True, seed=0, regime="local",
                  bf16=False):
    """Closures timing the fused per-layer entry points directly through the C ABI."""
    from snipper_b200 import capi
    L = capi.lib()
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(levels, dtype=torch.long)
    Lv = len(levels)
    S = int(shapes.prod(1).sum())
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, T2, S, M, D, generator=g).cuda()
    if regime == "init":
        import math
        th = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        dd = torch.stack([th.cos(), th.sin()], -1)
        dd = dd / dd.abs().max(-1, keepdim=True)[0]
        offsets = (dd.view(1, 1, 1, M, 1, 1, 2) * torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, 1, 1, 1, P, 1)
                   ).expand(N, T1, Lq, M, Lv, P, 2).contiguous().cuda()
    else:
        offsets = (torch.randn(N, T1, Lq, M, Lv, P, 2, generator=g) * 3.0).cuda()
    logits = torch.randn(N, T1, Lq, M, Lv, P, generator=g).cuda()
    if encoder:
        refs = []
        for H, W in levels:
            ys, xs = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
            refs.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
        ref = torch.cat(refs, 0)[None, None, :, None, :].expand(N, T1, Lq, Lv, 2).contiguous().cuda()
    else:
        ref = torch.rand(N, T1, Lq, Lv, 2, generator=g).cuda()
    go = torch.randn(N, T1, Lq, M * D, generator=g).cuda()
    dt = 2 if bf16 else 0
    if bf16:
        value, go = value.bfloat16(), go.bfloat16()
    out = torch.empty_like(go)
    gv, goff, glog = torch.empty(value.shape, device="cuda"), torch.empty_like(offsets), torch.empty_like(logits)
    shapes, lsi = shapes.cuda(), lsi.cuda()
    st = torch.cuda.current_stream().cuda_stream
    keep = (value, offsets, logits, ref, go, out, gv, goff, glog, shapes, lsi)
    rs = (ref.stride(0), ref.stride(1))

    def fwd():
        r = L.msda_snippet_forward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                   logits.data_ptr(), ref.data_ptr(), out.data_ptr(), N, T2, T1, n_frame, S, M, D,
                                   Lv, Lq, P, 0, 0, rs[0], rs[1], 0, 0, None, None, None, None, 0, 0, dt, 0, st)
        assert r == 0, r

    def bwd():
        r = L.msda_snippet_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                    logits.data_ptr(), ref.data_ptr(), go.data_ptr(), gv.data_ptr(),
                                    goff.data_ptr(), glog.data_ptr(), N, T2, T1, n_frame, S, M, D, Lv, Lq, P,
                                    0, 0, rs[0], rs[1], 0, 0, None, None, None, None, 0, 0, dt, 0, None, 0, st)
        assert r == 0, r

    # neighbour-frame pre-summation: streaming pass + one gather per query frame (and the mirror in the backward)
    slots = L.msda_snippet_num_slots(T1, n_frame)
    vsum = torch.empty(N, slots, S, M, D, device="cuda", dtype=value.dtype)
    gsum = torch.empty(N, slots, S, M, D, device="cuda")
    gv2 = torch.empty_like(value)
    pix = torch.zeros(N, T2, S, dtype=torch.bool, device="cuda")   # per-pixel padding mask (nothing padded)
    PRE = capi.MSDA_FLAG_PRESUMMED
    keep = keep + (vsum, gsum, gv2, pix)

    def fsum():
        r = L.msda_frame_sum(value.data_ptr(), pix.data_ptr(), vsum.data_ptr(), N, T2, T1, n_frame, S, M * D, 0, 0,
                             1, 0, dt, st)
        assert r == 0, r

    def fwd_pre():
        r = L.msda_snippet_forward(vsum.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                   logits.data_ptr(), ref.data_ptr(), out.data_ptr(), N, T2, T1, n_frame, S, M, D,
                                   Lv, Lq, P, 0, 0, rs[0], rs[1], 0, 0, None, None, None, None, 0, 0, dt, PRE, st)
        assert r == 0, r

    def bwd_pre():
        r = L.msda_snippet_backward(vsum.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                    logits.data_ptr(), ref.data_ptr(), go.data_ptr(), gsum.data_ptr(),
                                    goff.data_ptr(), glog.data_ptr(), N, T2, T1, n_frame, S, M, D, Lv, Lq, P,
                                    0, 0, rs[0], rs[1], 0, 0, None, None, None, None, 0, 0, dt, PRE, None, 0, st)
        assert r == 0, r

    def funsum():
        r = L.msda_frame_unsum(gsum.data_ptr(), pix.data_ptr(), gv2.data_ptr(), N, T2, T1, n_frame, S, M * D, 1, 0, dt, st)
        assert r == 0, r

    DET = PRE | capi.MSDA_FLAG_DETERMINISTIC
    ws_bytes = 0 if bf16 else L.msda_snippet_backward_workspace_bytes(N, T1, n_frame, S, M, D, Lv, Lq, P, dt, DET)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    keep = keep + (ws,)

    def bwd_pre_det():
        r = L.msda_snippet_backward(vsum.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                    logits.data_ptr(), ref.data_ptr(), go.data_ptr(), gsum.data_ptr(),
                                    goff.data_ptr(), glog.data_ptr(), N, T2, T1, n_frame, S, M, D, Lv, Lq, P,
                                    0, 0, rs[0], rs[1], 0, 0, None, None, None, None, 0, 0, dt, DET, ws_ptr, ws_bytes, st)
        assert r == 0, r

    fullmask = pix[..., None].expand(N, T2, S, M * D).contiguous()   # the reference's materialised (N,T,S,C) mask
    keep = keep + (fullmask,)

    def masked(fn_name, mask, mrs, mcs):
        def run():
            if fn_name == "fwd":
                r = L.msda_snippet_forward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                           logits.data_ptr(), ref.data_ptr(), out.data_ptr(), N, T2, T1, n_frame, S, M, D,
                                           Lv, Lq, P, 0, 0, rs[0], rs[1], 0, 0, None, None, None, mask.data_ptr(), mrs, mcs,
                                           dt, 0, st)
            else:
                r = L.msda_snippet_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                            logits.data_ptr(), ref.data_ptr(), go.data_ptr(), gv.data_ptr(),
                                            goff.data_ptr(), glog.data_ptr(), N, T2, T1, n_frame, S, M, D, Lv, Lq, P,
                                            0, 0, rs[0], rs[1], 0, 0, None, None, None, mask.data_ptr(), mrs, mcs,
                                            dt, 0, None, 0, st)
            assert r == 0, r
        return run

    # planar slots (csrc/msda_planar.cu): the same four passes on the library's own slot layout
    pl_elems = 0 if bf16 else L.msda_planar_slot_bytes(S, M, D, 0) // 4
    PLN = PRE | capi.MSDA_FLAG_PLANAR
    if pl_elems:
        pvsum = torch.empty(N, slots, pl_elems, device="cuda")
        pgsum = torch.empty(N, slots, pl_elems, device="cuda")
        keep = keep + (pvsum, pgsum)

    def fsum_pl():
        r = L.msda_frame_sum_planar(value.data_ptr(), pix.data_ptr(), pvsum.data_ptr(), N, T2, T1, n_frame, S, M, D, 0, 0,
                                    1, 0, dt, st)
        assert r == 0, r

    def fwd_pl():
        r = L.msda_snippet_forward(pvsum.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                   logits.data_ptr(), ref.data_ptr(), out.data_ptr(), N, T2, T1, n_frame, S, M, D,
                                   Lv, Lq, P, 0, 0, rs[0], rs[1], 0, 0, None, None, None, None, 0, 0, dt, PLN, st)
        assert r == 0, r

    def bwd_pl():
        r = L.msda_snippet_backward(pvsum.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), offsets.data_ptr(),
                                    logits.data_ptr(), ref.data_ptr(), go.data_ptr(), pgsum.data_ptr(),
                                    goff.data_ptr(), glog.data_ptr(), N, T2, T1, n_frame, S, M, D, Lv, Lq, P,
                                    0, 0, rs[0], rs[1], 0, 0, None, None, None, None, 0, 0, dt, PLN, None, 0, st)
        assert r == 0, r

    def funsum_pl():
        r = L.msda_frame_unsum_planar(pgsum.data_ptr(), pix.data_ptr(), gv2.data_ptr(), N, T2, T1, n_frame, S, M, D,
                                      1, 0, dt, st)
        assert r == 0, r

    def layer_fwd_pl():
        fsum_pl()
        fwd_pl()

    def layer_bwd_pl():
        bwd_pl()
        funsum_pl()

    def layer_fwd_pre():
        fsum()
        fwd_pre()

    def layer_bwd_pre():
        bwd_pre()
        funsum()

    e = 2 if bf16 else 4   # value / output / grad_output element size; everything else is fp32
    samples = N * T1 * Lq * M * Lv * P
    vbytes = min(N * T2 * S * M * D, 4 * samples * D * 3)
    fwd_b = e * (vbytes + N * T1 * Lq * M * D) + 4 * 3 * samples
    bwd_b = e * (vbytes + N * T1 * Lq * M * D) + 4 * (vbytes + 6 * samples)
    full = N * S * M * D
    extra = {"frame_sum": (fsum, e * full * (T2 + slots)), "fwd_presummed": (fwd_pre, fwd_b),
             "layer_fwd_presummed(sum+gather)": (layer_fwd_pre, fwd_b),
             "bwd_presummed": (bwd_pre, bwd_b), "frame_unsum": (funsum, full * (4 * slots + e * T2)),
             "layer_bwd_presummed(scatter+unsum)": (layer_bwd_pre, bwd_b)}
    if pl_elems:
        extra.update({"frame_sum_planar": (fsum_pl, e * full * T2 + 4 * N * slots * pl_elems), "fwd_planar": (fwd_pl, fwd_b),
                      "layer_fwd_planar(sum+gather)": (layer_fwd_pl, fwd_b), "bwd_planar": (bwd_pl, bwd_b),
                      "frame_unsum_planar": (funsum_pl, 4 * N * slots * pl_elems + e * full * T2),
                      "layer_bwd_planar(scatter+unsum)": (layer_bwd_pl, bwd_b)})
    if not bf16:
        extra["bwd_presummed_deterministic"] = (bwd_pre_det, bwd_b)
    if Lq <= 1024:   # the few-queries (decoder) launches are the ones that take the mask inside the gather
        extra["fwd_direct_pixel_mask"] = (masked("fwd", pix, 1, 0), fwd_b)
        extra["fwd_direct_channel_mask"] = (masked("fwd", fullmask, M * D, 1), fwd_b)
        extra["bwd_direct_pixel_mask"] = (masked("bwd", pix, 1, 0), bwd_b)
        extra["bwd_direct_channel_mask"] = (masked("bwd", fullmask, M * D, 1), bwd_b)
    return fwd, bwd, fwd_b, bwd_b, keep, extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--ref", action="store_true", help="also time the vendored reference op")
    ap.add_argument("--regime", default="local")
    ap.add_argument("--pairs", type=int, default=0, help="tile length knob for D=48 (8/16/32)")
    ap.add_argument("--snip-pairs", type=int, default=0)
    ap.add_argument("--bf16", action="store_true", help="also time the bf16 I/O mode")
    ap.add_argument("--head-major", action="store_true",
                    help="also time the per-call kernels on a head-major copy (value (N*M,S,1,D)): what a packed value layout would give")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--inner", type=int, default=10, help="back-to-back launches per timed interval")
    ap.add_argument("--only", default="", help="comma-separated substrings: time only the fused-layer passes whose name contains one")
    ap.add_argument("--cases", default="snip_enc_N1,snip_dec_N1,enc_N1,enc_N2,enc_N8,dec_N1,dec_N2")
    args = ap.parse_args()
    global WARMUP, INNER
    WARMUP = args.warmup
    INNER = args.inner
    if args.pairs:       # benchmark knobs are read from the environment once, at the first launch
        os.environ["MSDA_PAIRS_D48"] = str(args.pairs)
    if args.snip_pairs:
        os.environ["MSDA_SNIP_PAIRS_D48"] = str(args.snip_pairs)
    import snipper_b200  # noqa: F401
    from snipper_b200 import capi
    ref = None
    if args.ref:
        from oracle.build_ref import load_ref
        ref = load_ref()
    S = sum(h * w for h, w in LEVELS)
    cases = [("enc_N1", 1, S), ("enc_N2", 2, S), ("enc_N8", 8, S), ("dec_N1", 1, 60), ("dec_N2", 2, 60)]
    for name, N, T1, Lq, enc in (("snip_enc_N1", 1, 4, S, True), ("snip_dec_N1", 1, 6, 60, False)):
        if name not in args.cases.split(","):
            continue
        fwd, bwd, fb, bb, keep, extra = snippet_calls(N, T1, 4, Lq, encoder=enc, regime=args.regime)
        rows = [("ours_fused_layer", "fwd_direct", fwd, fb), ("ours_fused_layer", "bwd_direct", bwd, bb)]
        rows += [("ours_fused_layer", k, fn, nb) for k, (fn, nb) in extra.items()]
        if args.bf16:
            hf, hb, hfb, hbb, hkeep, hextra = snippet_calls(N, T1, 4, Lq, encoder=enc, regime=args.regime, bf16=True)
            rows += [("ours_fused_layer_bf16", "fwd_direct", hf, hfb), ("ours_fused_layer_bf16", "bwd_direct", hb, hbb)]
            rows += [("ours_fused_layer_bf16", k, fn, nb) for k, (fn, nb) in hextra.items()]
        if args.only:
            # the streaming passes still run once so that the gathers they feed see real data
            for impl, which, fn, nbytes in rows:
                if which.startswith("frame_sum"):
                    fn()
            rows = [r for r in rows if any(k in r[1] for k in args.only.split(","))]
        for impl, which, fn, nbytes in rows:
            med, best = time_fn(fn, args.iters, args.flush)
            print(json.dumps({"case": name, "impl": impl, "pass": which, "us_median": round(med, 2),
                              "us_best": round(best, 2), "alg_MB": round(nbytes / 1e6, 2),
                              "GBps": round(nbytes / med / 1e3, 1),
                              "frac_of_measured_hbm": round(nbytes / med / 1e3 / PEAK, 4),
                              "l2_flush": args.flush, "regime": args.regime, "pairs": args.snip_pairs or 16}))
        del keep
    for name, N, Lq in cases:
        if name not in args.cases.split(","):
            continue
        value, shapes, lsi, loc, attn, go = make(N, Lq, regime=args.regime)
        fb, bb = algorithmic_bytes(N, S, 8, 48, 3, 4, Lq)
        fwd, bwd, bwd_nm, bwd_det, keep = direct_calls(value, shapes, lsi, loc, attn, go)
        rows = [("ours", "fwd", fwd, fb), ("ours", "bwd", bwd, bb), ("ours", "bwd_nomemset", bwd_nm, bb),
                ("ours", "bwd_deterministic", bwd_det, bb)]
        if ref is not None:
            rows += [("vendored", "fwd", lambda: ref.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64), fb),
                     ("vendored", "bwd", lambda: ref.ms_deform_attn_backward(value, shapes, lsi, loc, attn, go, 64), bb)]
        if args.bf16:
            v16 = min(N * S * 8 * 48, 4 * N * Lq * 8 * 3 * 4 * 48)
            fb16 = 2 * (v16 + N * Lq * 8 * 48) + 4 * 3 * N * Lq * 8 * 12
            bb16 = 2 * (v16 + N * Lq * 8 * 48) + 4 * (v16 + 6 * N * Lq * 8 * 12)
            bfwd, bbwd, _, _, bkeep = direct_calls(value, shapes, lsi, loc, attn, go, bf16=True)
            rows += [("ours_bf16", "fwd", bfwd, fb16), ("ours_bf16", "bwd", bbwd, bb16)]
        if args.head_major:
            N_, S_, M_, D_ = value.shape
            hv = value.permute(0, 2, 1, 3).reshape(N_ * M_, S_, 1, D_).contiguous()
            hl = loc.permute(0, 2, 1, 3, 4, 5).reshape(N_ * M_, Lq, 1, 3, 4, 2).contiguous()
            ha = attn.permute(0, 2, 1, 3, 4).reshape(N_ * M_, Lq, 1, 3, 4).contiguous()
            hg = go.view(N_, Lq, M_, D_).permute(0, 2, 1, 3).reshape(N_ * M_, Lq, D_).contiguous()
            hfwd, hbwd, hbwd_nm, _, hkeep = direct_calls(hv, shapes, lsi, hl, ha, hg)
            rows += [("ours_head_major", "fwd", hfwd, fb), ("ours_head_major", "bwd", hbwd, bb)]
        for impl, which, fn, nbytes in rows:
            med, best = time_fn(fn, args.iters, args.flush)
            print(json.dumps({"case": name, "impl": impl, "pass": which, "us_median": round(med, 2),
                              "us_best": round(best, 2), "alg_MB": round(nbytes / 1e6, 2),
                              "GBps": round(nbytes / med / 1e3, 1), "frac_of_measured_hbm": round(nbytes / med / 1e3 / PEAK, 4),
                              "l2_flush": args.flush, "regime": args.regime, "pairs": args.pairs or 16}))


if __name__ == "__main__":
    main()
