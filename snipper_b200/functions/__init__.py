from .ms_deform_attn_func import MSDeformAttnFunction, ms_deform_attn

__all__ = ["MSDeformAttnFunction", "ms_deform_attn"]
