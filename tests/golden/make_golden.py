"""Generates tests/golden/*.npz by running the UNMODIFIED reference Python code from /root/reference in this
container (CPU tensors):

  * render/renderutils ops with use_python=True (render/renderutils/bsdf.py) -- forward values and every input
    gradient through torch autograd, on the reference tests' input distribution (torch.rand, tests/test_bsdf.py);
  * the PyTorch bilateral filter of render/optixutils/tests/filter_test.py:31-74: that script cannot be imported
    (it runs at import time and needs CUDA), so its BilateralDenoiser class is extracted from the file's AST at
    generation time, with device="cuda" replaced by "cpu" and its stale 11-channel input layout fed accordingly
    (col, nrm, kd [ignored by the filter], zdz).  Nothing from the reference is copied into this repository.

Run once (committed outputs travel to the GPU box, /root/reference does not):
    python tests/golden/make_golden.py
"""
import ast
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "render"))
import renderutils as ru  # noqa: E402  (the reference package; plugin compilation is lazy and never triggered here)

RES = 8


def rnd(gen, *shape):
    return torch.rand(*shape, generator=gen, dtype=torch.float32)


def run(fn, ins, name):
    ins = [i.clone().requires_grad_(True) for i in ins]
    out = fn(*ins)
    g = torch.Generator().manual_seed(999)
    dout = torch.rand(out.shape, generator=g)
    out.backward(dout)
    d = {"out": out.detach().numpy(), "dout": dout.numpy()}
    for k, i in enumerate(ins):
        d["in%d" % k] = i.detach().numpy()
        d["grad%d" % k] = i.grad.numpy() if i.grad is not None else np.zeros_like(i.detach().numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, out.shape)


def main():
    gen = torch.Generator().manual_seed(0)
    v3 = lambda: rnd(gen, 1, RES, RES, 3)
    v1 = lambda: rnd(gen, 1, RES, RES, 1)
    for bsdf in ("lambert", "frostbite"):
        run(lambda *a: ru.pbr_bsdf(*a, bsdf=bsdf, use_python=True), [v3() for _ in range(6)], "ref_pbr_bsdf_" + bsdf)
    run(lambda *a: ru.pbr_specular(*a, use_python=True), [v3(), v3(), v3(), v3(), v1()], "ref_pbr_specular")
    run(lambda *a: ru.lambert(*a, use_python=True), [v3(), v3()], "ref_lambert")
    run(lambda *a: ru.frostbite_diffuse(*a, use_python=True), [v3(), v3(), v3(), v1()], "ref_frostbite")
    run(lambda *a: ru._fresnel_shlick(*a, use_python=True), [v3(), v3(), v1()], "ref_fresnel_shlick")
    run(lambda *a: ru._ndf_ggx(*a, use_python=True), [v1(), v1()], "ref_ndf_ggx")
    run(lambda *a: ru._lambda_ggx(*a, use_python=True), [v1(), v1()], "ref_lambda_ggx")
    run(lambda *a: ru._masking_smith(*a, use_python=True), [v1(), v1(), v1()], "ref_masking_smith")
    for ts, gl in ((True, True), (False, False)):
        run(lambda *a: ru.prepare_shading_normal(*a, two_sided_shading=ts, opengl=gl, use_python=True), [v3() for _ in range(6)],
            "ref_prepare_shading_normal_%d%d" % (ts, gl))

    # ---- bilateral filter: class pulled out of the reference test script at run time ----
    src = open(os.path.join(REF, "render/optixutils/tests/filter_test.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in ("BilateralDenoiser", "dot")]
    code = "\n".join(ast.get_source_segment(src, n) for n in keep).replace('device="cuda"', 'device="cpu"')
    ns = {"torch": torch, "np": np, "math": __import__("math")}
    exec(compile(code, "filter_test_extract", "exec"), ns)
    H, W = 20, 27
    col = rnd(gen, 1, H, W, 3) * 2
    nrm = torch.nn.functional.normalize(rnd(gen, 1, H, W, 3) + torch.tensor([0.0, 0.0, 2.0]), dim=-1)
    z = torch.cumsum(rnd(gen, 1, H, W, 1) * 0.05, dim=2)
    dz = rnd(gen, 1, H, W, 1) * 0.05 + 0.01
    kd = torch.ones(1, H, W, 3)
    for sigma in (1.0, 2.0):
        den = ns["BilateralDenoiser"](sigma=sigma)
        out = den.forward(torch.cat([col, nrm, kd, z, dz], dim=-1))
        np.savez_compressed(os.path.join(OUT, "ref_bilateral_sigma%d.npz" % int(sigma)), col=col.numpy(), nrm=nrm.numpy(),
                            zdz=torch.cat([z, dz], -1).numpy(), sigma=np.float32(sigma), out=out.numpy())
        print("wrote bilateral", sigma)


if __name__ == "__main__":
    main()


def make_loss_xfm():
    """Row f3: ru.image_loss / ru.xfm_points / ru.xfm_vectors with use_python=True (render/renderutils/loss.py, ops.py:515,535)."""
    g = torch.Generator().manual_seed(5)
    out = {}
    img = (torch.rand(2, 9, 7, 3, generator=g) * 4).requires_grad_(True); tgt = (torch.rand(2, 9, 7, 3, generator=g) * 4).requires_grad_(True)
    out['img'] = img.detach().numpy(); out['tgt'] = tgt.detach().numpy()
    for loss in ['l1', 'mse', 'smape', 'relmse']:
        for tm in ['none', 'log_srgb']:
            img.grad = None; tgt.grad = None
            v = ru.image_loss(img, tgt, loss=loss, tonemapper=tm, use_python=True); v.backward()
            out['%s_%s' % (loss, tm)] = np.float64(v.item()); out['%s_%s_gi' % (loss, tm)] = img.grad.numpy().copy(); out['%s_%s_gt' % (loss, tm)] = tgt.grad.numpy().copy()
    pts = torch.rand(1, 11, 3, generator=g).requires_grad_(True); mtx = torch.rand(3, 4, 4, generator=g)
    o = ru.xfm_points(pts, mtx, use_python=True); d = torch.rand(o.shape, generator=g); o.backward(d)
    out.update(pts=pts.detach().numpy(), mtx=mtx.numpy(), xfmp=o.detach().numpy(), xfmp_d=d.numpy(), xfmp_g=pts.grad.numpy().copy())
    pts.grad = None
    o = ru.xfm_vectors(pts, mtx, use_python=True); d = torch.rand(o.shape, generator=g); o.backward(d)
    out.update(xfmv=o.detach().numpy(), xfmv_d=d.numpy(), xfmv_g=pts.grad.numpy().copy())
    np.savez_compressed(os.path.join(OUT, 'ref_loss_xfm.npz'), **out)
    print("wrote ref_loss_xfm")


if __name__ == "__main__":
    make_loss_xfm()


def make_update_pdf():
    """Row a17: EnvironmentLight.update_pdf (render/light.py:46-59) and util.pixel_grid (render/util.py:62-66).  Neither module can be
    imported here (nvdiffrast / imageio at import time, device="cuda" literals), so the two function bodies are taken from the files'
    ASTs at generation time, `device="cuda"` is rewritten to "cpu", and they run unmodified otherwise.  Nothing is copied into the repo."""
    import types

    def extract(path, want):
        tree = ast.parse(open(path).read())
        found = {}
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef) and node.name in want:
                found[node.name] = node
        return found

    class Cpu(ast.NodeTransformer):
        def visit_Constant(self, n):
            return ast.copy_location(ast.Constant("cpu"), n) if n.value == "cuda" else n

    pg = Cpu().visit(extract(os.path.join(REF, "render", "util.py"), {"pixel_grid"})["pixel_grid"])
    up = Cpu().visit(extract(os.path.join(REF, "render", "light.py"), {"update_pdf"})["update_pdf"])
    mod = ast.Module(body=[pg, up], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "np": np}
    exec(compile(mod, "<reference update_pdf>", "exec"), ns)
    ns["util"] = types.SimpleNamespace(pixel_grid=ns["pixel_grid"])
    out = {}
    g = torch.Generator().manual_seed(21)
    for name, hw in (("a", (16, 32)), ("b", (64, 64))):
        base = torch.rand(hw[0], hw[1], 3, generator=g) ** 3 * 10
        base[hw[0] // 4] = 0                                   # an all-black row exercises the `> 0` guards (light.py:58-59)
        self = types.SimpleNamespace(base=base)
        ns["update_pdf"](self)
        out.update({name + "_base": base.numpy(), name + "_pdf": self._pdf.numpy(), name + "_cols": self.cols.numpy(), name + "_rows": self.rows.numpy()})
    np.savez_compressed(os.path.join(OUT, "ref_update_pdf.npz"), **out)
    print("wrote ref_update_pdf")


if __name__ == "__main__":
    make_update_pdf()


def make_env_shade():
    """Rows a1-a8: outputs of the reference's OWN raygen program (kernel.cu, compiled unmodified for the host by oracle/ref_shim ->
    oracle/_ref) on a small scene, frozen with their inputs, so that the pin also holds where oracle/_ref cannot be loaded.  Shadow
    rays are answered by the oracle's brute-force predicate (the OptiX runtime is closed source)."""
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), ""))          # tests/
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))           # repo root
    from common import make_case, oracle
    from oracle import Reference
    o = oracle()
    ref = Reference(o)
    for tag, bsdf, N, shadow in (("pbr", "pbr", 4, 1.0), ("diffuse", "diffuse", 3, 0.5)):
        c = make_case(res=14, B=2, N=N, level=1, light_hw=(16, 32), seed=7, perm_rows=64)
        args = (c["scene"], c["mask"], c["ro"], c["pos"], c["nrm"], c["view"], c["kd"], c["ks"], c["light"], c["pdf"], c["rows"], c["cols"], c["perms"])
        kw = dict(BSDF=bsdf, n_samples_x=N, rnd_seed=5, shadow_scale=shadow)
        d, s = ref.env_shade(*args, **kw)
        g = np.random.default_rng(2)
        gd = g.uniform(size=d.shape).astype(np.float32); gs = g.uniform(size=d.shape).astype(np.float32)
        grads = ref.env_shade(*args, grads=(gd, gs), **kw)
        out = {k: c[k] for k in ("verts", "tris", "mask", "ro", "pos", "nrm", "view", "kd", "ks", "light", "pdf", "rows", "cols", "perms")}
        out.update(diff=d, spec=s, diff_grad=gd, spec_grad=gs, pos_grad=grads[0], nrm_grad=grads[1], kd_grad=grads[2], ks_grad=grads[3],
                   light_grad=grads[4], n_samples_x=np.int32(N), rnd_seed=np.int32(5), shadow_scale=np.float32(shadow))
        np.savez_compressed(os.path.join(OUT, "ref_env_shade_%s.npz" % tag), **out)
        print("wrote ref_env_shade_" + tag, d.shape, float(d.mean()))


if __name__ == "__main__":
    make_env_shade()
