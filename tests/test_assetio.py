"""CPU: on-disk formats (row f4): Radiance HDR (flat + run-length encoded) and Wavefront OBJ, round trips and hand-made files."""
import os

import numpy as np
import pytest

from nvdiffrecmc_b200 import assetio, synth


def test_hdr_roundtrip_and_precision(tmp_path):
    img = synth.hdr_light(24, 40)
    p = str(tmp_path / "a.hdr")
    assetio.save_hdr(p, img)
    back = assetio.load_hdr(p)
    assert back.shape == img.shape
    # RGBE: 8-bit mantissa shared exponent => relative error < 2^-7 of the pixel's max channel
    tol = img.max(-1, keepdims=True) / 128.0
    assert (np.abs(back - img) <= tol + 1e-30).all()
    assert assetio.load_hdr(p).dtype == np.float32


def test_hdr_reads_run_length_encoded_scanlines(tmp_path):
    """Hand-encode a 2 x 16 image with new-style RLE (what HDR tools and the reference's probes use)."""
    H, W = 2, 16
    rgbe = np.zeros((H, W, 4), np.uint8)
    rgbe[0, :, :] = [128, 64, 32, 129]                        # constant row -> pure run
    rgbe[1, :, 0] = np.arange(W) * 3; rgbe[1, :, 1] = 7; rgbe[1, :, 2] = np.arange(W)[::-1]; rgbe[1, :, 3] = 130
    body = bytearray()
    for y in range(H):
        body += bytes([2, 2, W >> 8, W & 255])
        for c in range(4):
            row = rgbe[y, :, c]
            if (row == row[0]).all():
                body += bytes([128 + W, int(row[0])])
            else:
                body += bytes([8]) + row[:8].tobytes() + bytes([8]) + row[8:].tobytes()      # two literal dumps
    p = str(tmp_path / "rle.hdr")
    with open(p, "wb") as f:
        f.write(b"#?RADIANCE\n# hand made\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1\n\n-Y 2 +X 16\n" + bytes(body))
    img = assetio.load_hdr(p)
    assert np.allclose(img[0, 3], np.array([128, 64, 32]) * 2.0 ** (129 - 136))
    assert np.allclose(img[1, 5], np.array([15, 7, 10]) * 2.0 ** (130 - 136))
    with open(p, "wb") as f:
        f.write(b"P6\n")
    with pytest.raises(ValueError):
        assetio.load_hdr(p)


def test_obj_roundtrip_and_polygons(tmp_path):
    v, f = synth.scene_mesh("blob", level=1)
    vn = synth.vertex_normals(v, f)
    p = str(tmp_path / "m.obj")
    assetio.save_obj(p, v, f, v_nrm=vn, t_nrm_idx=f, mtllib="m.mtl", usemtl="mat0")
    m = assetio.load_obj(p)
    assert np.array_equal(m["v_pos"], v) and np.array_equal(m["t_pos_idx"], f) and np.array_equal(m["t_nrm_idx"], f)
    assert np.allclose(m["v_nrm"], vn) and m["mtllib"] == "m.mtl" and m["usemtl"][0][0] == "mat0" and m["v_tex"] is None
    # quads, negative indices, texcoords with the OpenGL v flip (obj.py:77)
    q = str(tmp_path / "q.obj")
    with open(q, "w") as fh:
        fh.write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 0.25\nf 1/1 2/2 3/3 4/4\nf -4/-4 -3/-3 -2/-2\n")
    m = assetio.load_obj(q)
    assert m["t_pos_idx"].tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    assert m["t_tex_idx"].tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]] and np.allclose(m["v_tex"][3], [0, 0.75])


@pytest.mark.skipif(not os.path.exists("/root/reference/data/irrmaps/aerodynamics_workshop_2k.hdr"), reason="reference assets only exist in the build container")
def test_reads_the_reference_assets_when_present():
    img = assetio.load_hdr("/root/reference/data/irrmaps/aerodynamics_workshop_2k.hdr")
    assert img.shape == (1024, 2048, 3) and np.isfinite(img).all() and img.min() >= 0 and img.max() > 10
    m = assetio.load_obj("/root/reference/data/spot/spot.obj")
    assert m["t_pos_idx"].shape == (5856, 3) and m["v_pos"].shape[0] == 2930       # spot = 2 930 verts / 5 856 tris
    b = assetio.load_obj("/root/reference/data/bob/bob_tri.obj")
    assert b["t_pos_idx"].shape == (10688, 3)
