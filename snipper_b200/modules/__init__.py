from .ms_deform_attn import EncoderGrid, MSDeformAttn, neighbour_frames

__all__ = ["MSDeformAttn", "EncoderGrid", "neighbour_frames"]
