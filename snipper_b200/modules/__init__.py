from .ms_deform_attn import MSDeformAttn, neighbour_frames

__all__ = ["MSDeformAttn", "neighbour_frames"]
