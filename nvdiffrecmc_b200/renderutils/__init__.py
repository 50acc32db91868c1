# Same public surface as the reference package (render/renderutils/__init__.py:9-10), minus
# xfm_points / xfm_vectors / image_loss which SURVEY.md section 8 marks "next" (row f3).
from .ops import prepare_shading_normal, lambert, frostbite_diffuse, pbr_specular, pbr_bsdf, _fresnel_shlick, _ndf_ggx, _lambda_ggx, _masking_smith
__all__ = ["prepare_shading_normal", "lambert", "frostbite_diffuse", "pbr_specular", "pbr_bsdf", "_fresnel_shlick", "_ndf_ggx", "_lambda_ggx",
           "_masking_smith"]
