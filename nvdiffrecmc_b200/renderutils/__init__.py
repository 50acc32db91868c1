"""Drop-in mirror of the reference's `render.renderutils` package: the same twelve public names (render/renderutils/__init__.py:9-10),
every one backed by a libmcshade kernel (`use_python=True` selects the PyTorch twin in bsdf.py / loss.py)."""
from . import ops as _ops

_PUBLIC = ("pbr_bsdf", "pbr_specular", "lambert", "frostbite_diffuse", "prepare_shading_normal",      # shading
           "_fresnel_shlick", "_ndf_ggx", "_lambda_ggx", "_masking_smith",                            # GGX building blocks
           "image_loss", "xfm_points", "xfm_vectors")                                                 # loss + transforms
globals().update({name: getattr(_ops, name) for name in _PUBLIC})
__all__ = list(_PUBLIC)
