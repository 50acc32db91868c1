"""Freeze the output of the REFERENCE's own render.shade() (render/render.py:30-164, imported unmodified) on a synthetic G-buffer.
Runs only where /root/reference exists (build container, CPU): plugins = oracle-backed optixutils stand-in + the reference's own
PyTorch renderutils path (tests/refshade.py).  usage: python tests/golden/make_shade_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def generate():
    from common import oracle
    import refshade
    inp = refshade.make_inputs()
    ou, ru = refshade.oracle_backends(oracle(), inp["perms"])
    buf, lgt = refshade.run_shade(inp, ou, ru)
    out = dict(inp)
    for k in ("shaded", "diffuse_light", "specular_light", "normal", "kd", "ks"):
        out["out_" + k] = buf[k]
    out.update({"lgt_pdf": lgt["pdf"], "lgt_rows": lgt["rows"], "lgt_cols": lgt["cols"]})
    return out


if __name__ == "__main__":
    d = generate()
    np.savez_compressed(os.path.join(HERE, "ref_shade_pbr.npz"), **d)
    print("wrote ref_shade_pbr.npz:", {k: getattr(v, "shape", v) for k, v in d.items() if k.startswith("out_")})
