"""Primary visibility and attribute interpolation without a rasteriser.

Stand-ins for the two nvdiffrast calls of the reference's G-buffer pass (render/render.py:208-234): `rasterize` traces one
primary ray per pixel through the LBVH that `optix_build_bvh` already built for the shadow rays and returns nvdiffrast's
`rast` tensor `(u, v, z/w, triangle_id + 1)`; `interpolate` evaluates vertex attributes at those barycentrics and is
differentiable with respect to the attributes (float atomics in the backward pass, like dr.interpolate).  Screen-space
derivatives (`rast_db`, `diff_attrs`) and antialiasing are not provided."""
import torch
from . import _lib as L


def rasterize(optix_ctx, mtx, resolution):
    """mtx: [B,4,4] clip-space transform (clip = mtx @ (p, 1), the `mtx_in` of render_mesh, render.py:289-293);
    resolution: (H, W).  Returns rast [B,H,W,4] fp32; a pixel whose ray hits nothing is all zeros."""
    L.require_cuda(mtx)
    if mtx.dim() != 3 or mtx.shape[1:] != (4, 4):
        raise ValueError("rasterize: mtx must be [B,4,4]")
    m = mtx.detach().to(torch.float32).contiguous()
    B, (H, W) = m.shape[0], resolution
    rast = torch.empty(B, H, W, 4, dtype=torch.float32, device=m.device)
    L.check(L.lib().mcs_rasterize(optix_ctx.cpp_wrapper, m.data_ptr(), B, H, W, rast.data_ptr(), L.stream_ptr()), "rasterize")
    return rast


class _interpolate_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        L.require_cuda(attr, rast, tri)
        if tri.dtype != torch.int32:
            raise TypeError("interpolate: tri must be int32 [T,3]")
        a = attr.to(torch.float32).contiguous(); r = rast.to(torch.float32).contiguous(); t = tri.contiguous()
        batched = a.dim() == 3
        V, Cn = a.shape[-2], a.shape[-1]
        B, H, W = r.shape[0], r.shape[1], r.shape[2]
        if batched and a.shape[0] != B:
            raise ValueError("interpolate: attribute batch %d does not match rast batch %d" % (a.shape[0], B))
        out = torch.empty(B, H, W, Cn, dtype=torch.float32, device=a.device)
        L.check(L.lib().mcs_interpolate_fwd(a.data_ptr(), V * Cn if batched else 0, V, Cn, t.data_ptr(), t.shape[0], r.data_ptr(), B, H, W,
                                            out.data_ptr(), L.stream_ptr()), "interpolate (forward)")
        ctx.save_for_backward(a, r, t)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, r, t = ctx.saved_tensors
        batched = a.dim() == 3
        V, Cn = a.shape[-2], a.shape[-1]
        B, H, W = r.shape[0], r.shape[1], r.shape[2]
        d_attr = torch.zeros_like(a)
        g = dout.to(torch.float32).contiguous()
        L.check(L.lib().mcs_interpolate_bwd(a.data_ptr(), V * Cn if batched else 0, V, Cn, t.data_ptr(), t.shape[0], r.data_ptr(), B, H, W,
                                            g.data_ptr(), d_attr.data_ptr(), L.stream_ptr()), "interpolate (backward)")
        return d_attr, None, None


def interpolate(attr, rast, tri):
    """attr [V,C] or [B,V,C], rast from `rasterize`, tri int32 [T,3].  Returns (out [B,H,W,C], None) like dr.interpolate."""
    return _interpolate_func.apply(attr, rast, tri), None


class _texel_fetch_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, idx):
        L.require_cuda(tex, idx)
        if idx.dtype != torch.int64:
            raise TypeError("texel_fetch: idx must be int64")
        t = tex.to(torch.float32).contiguous(); ix = idx.contiguous()
        T, Cn = t.shape[0], t.shape[1]
        out = torch.empty(*ix.shape, Cn, dtype=torch.float32, device=t.device)
        L.check(L.lib().mcs_texel_fetch_fwd(t.data_ptr(), T, Cn, ix.data_ptr(), ix.numel(), out.data_ptr(), L.stream_ptr()), "texel_fetch (forward)")
        ctx.save_for_backward(ix)
        ctx.shape = (T, Cn)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ix,) = ctx.saved_tensors
        T, Cn = ctx.shape
        d_tex = torch.zeros(T, Cn, dtype=torch.float32, device=ix.device)
        g = dout.to(torch.float32).contiguous()
        L.check(L.lib().mcs_texel_fetch_bwd(T, Cn, ix.data_ptr(), ix.numel(), g.data_ptr(), d_tex.data_ptr(), L.stream_ptr()), "texel_fetch (backward)")
        return d_tex, None


def texel_fetch(tex, idx):
    """tex [T,C] fp32, idx int64 [...] -> [..., C]; `tex[idx]` with an atomic scatter-add backward (nearest-filter material look-up)."""
    return _texel_fetch_func.apply(tex, idx)
