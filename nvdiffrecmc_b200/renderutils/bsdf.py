"""PyTorch (autograd) twins of the CUDA ops, selected with use_python=True exactly like the
reference's render/renderutils/bsdf.py.  They exist for validation only (any device, any dtype)
and are written from the formulas, not from the CUDA kernels: Lambert, Frostbite/Disney diffuse,
Schlick Fresnel, GGX NDF, Smith-correlated masking, and the tangent-space shading normal with
two-sided flip and view-dependent bending."""
import math

import torch

EPS = 1e-4            # SPECULAR_EPSILON (renderutils/c_src/bsdf.cu:12)
BEND_THRESHOLD = 0.1  # NORMAL_THRESHOLD (renderutils/c_src/normal.cu:12)


def dot3(a, b):
    return (a * b).sum(dim=-1, keepdim=True)


def unit(v):
    return torch.nn.functional.normalize(v, dim=-1)


def _cclamp(c):
    return c.clamp(EPS, 1.0 - EPS)


def lambert(nrm, wi):
    return dot3(nrm, wi).clamp(min=0.0) / math.pi


def fresnel_schlick(f0, f90, cos_theta):
    s = (1.0 - _cclamp(cos_theta)) ** 5
    return f0 * (1.0 - s) + f90 * s


def frostbite_diffuse(nrm, wi, wo, lin_rough):
    wi_n, wo_n = dot3(wi, nrm), dot3(wo, nrm)
    wi_h = dot3(wi, unit(wo + wi))
    f90 = 0.5 * lin_rough + 2.0 * wi_h * wi_h * lin_rough
    e = 1.0 - (0.51 / 1.51) * lin_rough
    val = fresnel_schlick(1.0, f90, wi_n) * fresnel_schlick(1.0, f90, wo_n) * e
    return torch.where((wi_n > 0) & (wo_n > 0), val, torch.zeros_like(val))


def ndf_ggx(alpha_sqr, cos_theta):
    c = _cclamp(cos_theta)
    d = (c * alpha_sqr - c) * c + 1.0
    return alpha_sqr / (math.pi * d * d)


def lambda_ggx(alpha_sqr, cos_theta):
    c2 = _cclamp(cos_theta) ** 2
    return 0.5 * (torch.sqrt(1.0 + alpha_sqr * (1.0 - c2) / c2) - 1.0)


def masking_smith(alpha_sqr, cos_i, cos_o):
    return 1.0 / (1.0 + lambda_ggx(alpha_sqr, cos_i) + lambda_ggx(alpha_sqr, cos_o))


def pbr_specular(col, nrm, wo, wi, alpha, min_roughness=0.08):
    a2 = alpha.clamp(min_roughness * min_roughness, 1.0) ** 2
    h = unit(wo + wi)
    wo_n, wi_n = dot3(wo, nrm), dot3(wi, nrm)
    w = fresnel_schlick(col, 1.0, dot3(wo, h)) * ndf_ggx(a2, dot3(nrm, h)) * masking_smith(a2, wo_n, wi_n) * 0.25 / wo_n.clamp(min=EPS)
    return torch.where((wo_n > EPS) & (wi_n > EPS), w, torch.zeros_like(w))


def pbr_bsdf(kd, arm, pos, nrm, view_pos, light_pos, min_roughness, BSDF):
    wo, wi = unit(view_pos - pos), unit(light_pos - pos)
    occl, rough, metal = arm[..., 0:1], arm[..., 1:2], arm[..., 2:3]
    spec_col = (0.04 * (1.0 - metal) + kd * metal) * (1.0 - occl)
    diff_col = kd * (1.0 - metal)
    d = lambert(nrm, wi) if BSDF == 0 else frostbite_diffuse(nrm, wi, wo, rough)
    return diff_col * d + pbr_specular(spec_col, nrm, wo, wi, rough * rough, min_roughness)


def prepare_shading_normal(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading, opengl):
    n, t, v = unit(smooth_nrm), unit(smooth_tng), unit(view_pos - pos)
    b = unit(torch.cross(t, n, dim=-1))
    sgn = -1.0 if opengl else 1.0
    sh = unit(t * perturbed_nrm[..., 0:1] + sgn * b * perturbed_nrm[..., 1:2] + n * perturbed_nrm[..., 2:3].clamp(min=0.0))
    g = geom_nrm
    if two_sided_shading:
        front = dot3(g, v) > 0
        sh = torch.where(front, sh, -sh)
        g = torch.where(front, g, -g)
    w = (dot3(v, sh) / BEND_THRESHOLD).clamp(0.0, 1.0)
    return torch.lerp(g, sh, w)
