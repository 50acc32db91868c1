"""Run the REFERENCE's own `render.shade()` (render/render.py:30-164, imported unmodified from /root/reference) with its two plugin
packages swapped for replacements -- TEST INFRASTRUCTURE, only usable where /root/reference exists (the build container).

`render/render.py` does `from . import renderutils as ru` / `from . import optixutils as ou` and `import nvdiffrast.torch as dr`;
`render/util.py` also imports `imageio`.  None of nvdiffrast / imageio / a GPU exist in the build container, so:
  * `nvdiffrast.torch` is a stub whose only implemented entry point is `texture(..., filter_mode='linear', boundary_mode='clamp')`
    (bilinear look-up with clamped borders == torch grid_sample(align_corners=False, padding_mode='border')); shade() uses it for the
    jittered regulariser taps only, never for the shaded colour;
  * `imageio` is an empty stub (image IO is not on this path);
  * `render.optixutils` / `render.renderutils` are whatever backend the caller passes: on CPU the ORACLE-backed stand-in below (golden
    generation), on a GPU box this repository's packages (the drop-in itself, INTEGRATION.md option A);
  * the reference hard-codes device="cuda" in a few factory calls (render.py:50,63; util.py:62-66); without a GPU those calls are
    redirected to the CPU for the duration of the run.
"""
import contextlib
import importlib
import sys
import types

import numpy as np
import torch

REF_ROOT = "/root/reference"


def _texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode='auto', boundary_mode='wrap', max_mip_level=None):
    assert filter_mode == 'linear' and boundary_mode == 'clamp', "stub implements the one mode render.shade() uses"
    grid = uv * 2.0 - 1.0
    out = torch.nn.functional.grid_sample(tex.permute(0, 3, 1, 2), grid, mode='bilinear', padding_mode='border', align_corners=False)
    return out.permute(0, 2, 3, 1)


def _unavailable(name):
    def f(*a, **k):
        raise RuntimeError("nvdiffrast.torch.%s is not available in this environment (stub)" % name)
    return f


@contextlib.contextmanager
def reference_render(ou_backend, ru_backend):
    """Context manager yielding the reference's `render.render` module with the given optixutils / renderutils replacements."""
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "render" or k.startswith("render.") or k.startswith("nvdiffrast") or k in ("imageio", "denoiser", "denoiser.denoiser")}
    for k in saved:
        sys.modules.pop(k, None)
    dr = types.ModuleType("nvdiffrast.torch")
    dr.texture = _texture
    for n in ("interpolate", "rasterize", "antialias", "RasterizeCudaContext", "RasterizeGLContext", "DepthPeeler"):
        setattr(dr, n, _unavailable(n))
    nv = types.ModuleType("nvdiffrast"); nv.torch = dr
    sys.modules.update({"nvdiffrast": nv, "nvdiffrast.torch": dr, "imageio": types.ModuleType("imageio"),
                        "render.optixutils": ou_backend, "render.renderutils": ru_backend})
    sys.path.insert(0, REF_ROOT)
    patched = []
    if not torch.cuda.is_available():
        def cpu(fn):
            def g(*a, **k):
                if str(k.get("device", "")).startswith("cuda"):
                    k["device"] = "cpu"
                return fn(*a, **k)
            return g
        for n in ("normal", "arange", "tensor", "zeros", "ones", "full"):
            patched.append((n, getattr(torch, n)))
            setattr(torch, n, cpu(getattr(torch, n)))
    try:
        pkg = importlib.import_module("render")          # namespace package rooted at /root/reference/render
        pkg.optixutils, pkg.renderutils = ou_backend, ru_backend
        mod = importlib.import_module("render.render")
        den = importlib.import_module("denoiser.denoiser")
        yield mod, importlib.import_module("render.light"), den
    finally:
        for n, fn in patched:
            setattr(torch, n, fn)
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "render" or k.startswith("render.") or k.startswith("nvdiffrast") or k in ("imageio", "denoiser", "denoiser.denoiser")]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


# ------------------------------------------------------------------------------------------------
# CPU stand-ins with the reference packages' surface, backed by the oracle (forward only)
# ------------------------------------------------------------------------------------------------
def oracle_backends(orc, perms):
    ou = types.ModuleType("oracle_optixutils")

    class OptiXContext:
        def __init__(self):
            self.scene = None
    ou.OptiXContext = OptiXContext

    def optix_build_bvh(ctx, verts, tris, rebuild):
        ctx.scene = orc.scene(verts.detach().cpu().numpy().astype(np.float32), tris.detach().cpu().numpy().astype(np.int32))
    ou.optix_build_bvh = optix_build_bvh

    def optix_env_shade(ctx, mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols, BSDF='pbr', n_samples_x=8, rnd_seed=None,
                        shadow_scale=1.0):
        n = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), np.float32)
        B, H, W = ro.shape[:3]
        view = np.broadcast_to(n(gb_view_pos), (B, H, W, 3))
        d, s = orc.env_shade(ctx.scene, n(mask), n(ro), n(gb_pos), n(gb_normal), view, n(gb_kd), n(gb_ks), n(light), n(pdf), n(rows), n(cols), perms,
                             BSDF=BSDF, n_samples_x=n_samples_x, rnd_seed=rnd_seed, shadow_scale=shadow_scale)
        return torch.tensor(d), torch.tensor(s)
    ou.optix_env_shade = optix_env_shade

    def bilateral_denoiser(col, nrm, zdz, sigma):
        n = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), np.float32)
        return torch.tensor(orc.bilateral_denoiser(n(col), n(nrm), n(zdz), sigma))
    ou.bilateral_denoiser = bilateral_denoiser

    # renderutils: the REFERENCE's own PyTorch implementation (use_python=True; its CUDA plugin cannot be built here)
    sys.path.insert(0, REF_ROOT + "/render")
    try:
        for k in [k for k in sys.modules if k == "renderutils" or k.startswith("renderutils.")]:
            sys.modules.pop(k)
        ref_ru = importlib.import_module("renderutils")
    finally:
        sys.path.remove(REF_ROOT + "/render")
    ru = types.ModuleType("reference_renderutils_python_path")

    def prepare_shading_normal(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading=True, opengl=True, use_python=False):
        if perturbed_nrm is None:      # ops.py:217-218 allocates this default on 'cuda'
            perturbed_nrm = torch.tensor([0, 0, 1], dtype=torch.float32)[None, None, None, ...]
        return ref_ru.prepare_shading_normal(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading, opengl, use_python=True)
    ru.prepare_shading_normal = prepare_shading_normal
    return ou, ru


class _Tex:
    def __init__(self, img):
        self.img = img

    def sample(self, texc, texc_deriv, filter_mode='linear-mipmap-linear'):
        return self.img


class _Flags:
    n_samples = 4
    decorrelated = False
    denoiser_demodulate = True


def make_inputs(res=40, B=2, N=4, seed=21):
    """Synthetic shade() inputs (numpy): G-buffer of a blob+ring scene, per-pixel material images, a 32 x 64 HDR-ish probe."""
    from common import make_case
    c = make_case(res=res, B=B, N=N, light="hdr", light_hw=(32, 64), seed=seed, level=2)
    rast = np.zeros((B, res, res, 4), np.float32); rast[..., 3] = c["mask"]
    gb_depth = np.stack([c["depth"], np.full_like(c["depth"], 0.01)], -1).astype(np.float32)
    ks = c["ks"].copy()
    return dict(verts=c["verts"], tris=c["tris"], rast=rast, gb_depth=gb_depth, gb_pos=c["pos"], gb_geometric_normal=c["geom_nrm"], gb_normal=c["smooth_nrm"],
                gb_tangent=c["tangent"], view_pos=c["view"], kd=c["kd"], ks=ks, light=c["light"], perms=c["perms"], n_samples_x=N, sigma_influence=1.0,
                rnd_seed=5, shadow_scale=1.0)


def run_shade(inp, ou_backend, ru_backend, device="cpu", bsdf="pbr"):
    """Calls the reference's shade() on `inp`; returns the buffers it produces as numpy arrays."""
    t = lambda k: torch.tensor(inp[k], device=device)
    with reference_render(ou_backend, ru_backend) as (render, light, den):
        render.rnd_seed = int(inp["rnd_seed"])
        ctx = ou_backend.OptiXContext()
        ou_backend.optix_build_bvh(ctx, t("verts"), t("tris"), rebuild=1)
        lgt = light.EnvironmentLight(t("light"))                                   # render/light.py:21-59 (update_pdf in the constructor)
        flags = _Flags(); flags.n_samples = int(inp["n_samples_x"])
        material = {"bsdf": bsdf, "kd": _Tex(t("kd")), "ks": _Tex(t("ks"))}
        denoiser = den.BilateralDenoiser(influence=float(inp["sigma_influence"]))
        B, H, W = inp["rast"].shape[:3]
        texc = torch.zeros(B, H, W, 2, device=device)
        with torch.no_grad():
            buffers = render.shade(flags, t("rast"), t("gb_depth"), t("gb_pos"), t("gb_geometric_normal"), t("gb_normal"), t("gb_tangent"), texc, texc,
                                   t("view_pos"), lgt, material, ctx, None, None, denoiser, float(inp["shadow_scale"]))
        return {k: v.detach().cpu().numpy() for k, v in buffers.items()}, {"pdf": lgt._pdf.cpu().numpy(), "rows": lgt.rows.cpu().numpy(), "cols": lgt.cols.cpu().numpy()}
