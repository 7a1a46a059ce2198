"""TEST INFRASTRUCTURE ONLY -- torch/CPU restatement of the reference's PyTorch path.

* :func:`msda_core_torch` restates ``ms_deform_attn_core_pytorch``
  (reference models/ops/functions/ms_deform_attn_func.py:45-65): per level, reshape the
  level's slab of ``value`` to an image batch (N*M, D, H, W) and bilinearly sample it with
  ``F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False)`` on the grid
  ``2*loc - 1``; weight by the attention weights and sum over levels x points.
* :class:`SnippetMSDeformAttnRef` restates Snipper's per-frame wrapper
  (reference models/ops/modules/ms_deform_attn.py:37-243): same constructor arguments,
  same state-dict keys, same neighbour-frame rule, softmax over (levels, points, frames).

Both are differentiable through autograd, so they double as the gradient oracle for the
module-level tests and as the CPU baseline ("port") timed by bench.py.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def msda_core_torch(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D), spatial_shapes (L,2) [(H,W)], sampling_locations (N,Lq,M,L,P,2) in
    (x,y) normalised coords, attention_weights (N,Lq,M,L,P)  ->  (N, Lq, M*D)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    sizes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    # (N*M, 1, Lq, L, P): one weight per sampled vector
    w = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L, P)
    per_level = []
    start = 0
    for l, (H, W) in enumerate(sizes):
        feat = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(N * M, D, H, W)
        grid = (2.0 * sampling_locations[:, :, :, l] - 1.0).permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        per_level.append(F.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))  # (N*M, D, Lq, P)
        start += H * W
    sampled = torch.stack(per_level, dim=3)  # (N*M, D, Lq, L, P)
    out = (sampled * w).flatten(3).sum(-1)   # (N*M, D, Lq)
    return out.view(N, M * D, Lq).transpose(1, 2).contiguous()


def neighbour_frames(t1, n_frame, T2):
    """Frames a query of frame t1 samples from (reference ms_deform_attn.py:137-140, 189, 201):
    observed frames look at t1-1, t1, t1+1 clipped to [0, n_frame); future frames at all T2."""
    if t1 < n_frame:
        return [t for t in (t1 - 1, t1, t1 + 1) if 0 <= t < n_frame]
    return list(range(T2))


class SnippetMSDeformAttnRef(nn.Module):
    """CPU restatement of the reference ``MSDeformAttn`` (same ctor, same state-dict keys).

    ``core`` selects the inner op: the grid_sample restatement (default) or any callable
    ``core(value, shapes, level_start_index, loc, attn) -> (N,Lq,M*D)``.
    """

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, n_frame=4,
                 mode="encoder", use_pytroch_deform=True, attention_vis=False, core=None):
        super().__init__()
        assert d_model % n_heads == 0 and mode in ("encoder", "decoder")
        self.d_model, self.n_levels, self.n_heads = d_model, n_levels, n_heads
        self.n_points, self.n_frame, self.mode = n_points, n_frame, mode
        self.attention_vis = attention_vis
        self.core = core
        # one Linear shared by every frame slot (reference ms_deform_attn.py:68-71)
        off = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        att = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.sampling_offsets = nn.ModuleList([off] * n_frame)
        self.attention_weights = nn.ModuleList([att] * n_frame)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # reference ms_deform_attn.py:78-97
        M, L, P = self.n_heads, self.n_levels, self.n_points
        ang = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        d = torch.stack([ang.cos(), ang.sin()], -1)
        d = d / d.abs().max(-1, keepdim=True)[0]
        bias = d.view(M, 1, 1, 2).repeat(1, L, P, 1) * torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, P, 1)
        with torch.no_grad():
            self.sampling_offsets[0].weight.zero_()
            self.sampling_offsets[0].bias.copy_(bias.reshape(-1))
            self.attention_weights[0].weight.zero_()
            self.attention_weights[0].bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    def _core(self, value, shapes, lsi, loc, attn):
        if self.core is not None:
            return self.core(value, shapes, lsi, loc, attn)
        return msda_core_torch(value, shapes, loc, attn)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                input_level_start_index, input_padding_mask=None):
        N, T1, Lq, _ = query.shape
        _, T2, S, _ = input_flatten.shape
        M, L, P = self.n_heads, self.n_levels, self.n_points
        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask, 0.0)
        value = value.view(N, T2, S, M, self.d_model // M)
        wh = torch.stack([input_spatial_shapes[:, 1], input_spatial_shapes[:, 0]], -1)  # (L,2) = (W,H)
        outs, vis_loc, vis_att = [], [], []
        for t1 in range(T1):
            nb = neighbour_frames(t1, self.n_frame, T2)
            q = query[:, t1]
            logits = torch.stack([self.attention_weights[t2](q).view(N, Lq, M, L, P) for t2 in nb], -1)
            A = F.softmax(logits.flatten(-3), -1).view(N, Lq, M, L, P, len(nb))
            acc, locs = 0, []
            for j, t2 in enumerate(nb):
                off = self.sampling_offsets[t2](q).view(N, Lq, M, L, P, 2) / wh[None, None, None, :, None, :]
                loc = reference_points[:, t1, :, None, :, None, :] + off
                locs.append(loc.detach())
                acc = acc + self._core(value[:, t2], input_spatial_shapes, input_level_start_index,
                                       loc, A[..., j])
            outs.append(acc)
            if self.attention_vis:
                vis_loc.append(torch.stack(locs, -2))
                vis_att.append(A.detach())
        out = self.output_proj(torch.stack(outs, 1))
        if self.attention_vis:
            return out, (vis_loc, vis_att)
        return out
