# Same public surface as the reference package (render/optixutils/__init__.py:9-10).
from .ops import (OptiXContext, optix_build_bvh, optix_env_shade, bilateral_denoiser, bilateral_denoiser2, trace_visibility, trace_closest,
                  shade_combine, denoise_and_combine)
__all__ = ["OptiXContext", "optix_build_bvh", "optix_env_shade", "bilateral_denoiser"]
