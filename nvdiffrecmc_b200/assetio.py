"""On-disk formats used by the hot path's fixtures, without imageio / freeimage:

  * Radiance RGBE `.hdr` reader / writer  -- the reference goes through imageio (render/util.py:355-384; light.py:71-79 load_env,
    light.py:89-93 save_env_map);
  * Wavefront `.obj` reader / writer (positions, normals, texcoords, triangulated faces) -- render/obj.py:31-176.

numpy only; tensors are created by the caller.  The HDR reader handles flat and new-style run-length-encoded scanlines, the writer
emits flat scanlines (every reader accepts them).
"""
import numpy as np


# ------------------------------------------------------------------------------------------------ Radiance HDR
def _rgbe_to_float(rgbe):
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - (128 + 8)), 0.0).astype(np.float32)
    return rgbe[..., :3].astype(np.float32) * scale[..., None]


def _float_to_rgbe(img):
    img = np.maximum(np.asarray(img, np.float32), 0.0)
    m = img.max(axis=-1)
    mant, ex = np.frexp(m)                                   # m = mant * 2^ex, mant in [0.5, 1)
    scale = np.where(m > 1e-32, mant * 256.0 / np.maximum(m, 1e-38), 0.0)
    out = np.zeros(img.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(img * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, ex + 128, 0).astype(np.uint8)
    return out


def load_hdr(path):
    """-> float32 [H, W, 3] (linear radiance).  Supports '-Y H +X W' orientation (the only one the reference's probes use)."""
    with open(path, "rb") as f:
        data = f.read()
    if not (data.startswith(b"#?RADIANCE") or data.startswith(b"#?RGBE")):
        raise ValueError("%s: not a Radiance HDR file" % path)
    pos = 0
    fmt_ok = False
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if line.startswith(b"FORMAT="):
            fmt_ok = line.strip() == b"FORMAT=32-bit_rle_rgbe"
        if line == b"":
            break
    if not fmt_ok:
        raise ValueError("%s: unsupported FORMAT (need 32-bit_rle_rgbe)" % path)
    end = data.index(b"\n", pos)
    res = data[pos:end].split()
    pos = end + 1
    if len(res) != 4 or res[0] != b"-Y" or res[2] != b"+X":
        raise ValueError("%s: unsupported resolution line %r" % (path, data[pos:end]))
    H, W = int(res[1]), int(res[3])
    buf = np.frombuffer(data, np.uint8, offset=pos)
    img = np.zeros((H, W, 4), np.uint8)
    p = 0
    for y in range(H):
        if W < 8 or W > 0x7FFF or buf[p] != 2 or buf[p + 1] != 2 or (buf[p + 2] & 0x80):
            img[y] = buf[p:p + 4 * W].reshape(W, 4)           # flat scanline
            p += 4 * W
            continue
        if ((int(buf[p + 2]) << 8) | int(buf[p + 3])) != W:
            raise ValueError("%s: scanline width mismatch" % path)
        p += 4
        for c in range(4):                                    # new-style RLE: each channel separately
            x = 0
            while x < W:
                n = int(buf[p]); p += 1
                if n > 128:
                    n -= 128
                    img[y, x:x + n, c] = buf[p]; p += 1
                else:
                    img[y, x:x + n, c] = buf[p:p + n]; p += n
                x += n
    return _rgbe_to_float(img)


def save_hdr(path, img):
    """float [H, W, 3] -> Radiance RGBE file with flat scanlines."""
    img = np.asarray(img, np.float32)
    assert img.ndim == 3 and img.shape[2] == 3
    H, W = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(("-Y %d +X %d\n" % (H, W)).encode())
        f.write(_float_to_rgbe(img).tobytes())


# ------------------------------------------------------------------------------------------------ Wavefront OBJ
def load_obj(path):
    """-> dict(v_pos [V,3] f32, v_nrm [N,3] | None, v_tex [T,2] | None, t_pos_idx [F,3] i32, t_nrm_idx | None, t_tex_idx | None,
    mtllib, usemtl list).  Polygons are fan-triangulated, negative indices resolved, v flipped to OpenGL (1 - v) like obj.py:77."""
    v, vn, vt = [], [], []
    fp, fn, ft = [], [], []
    mtllib, mats = None, []
    with open(path, "r") as f:
        for line in f:
            s = line.split()
            if not s:
                continue
            k = s[0].lower()
            if k == "v":
                v.append([float(x) for x in s[1:4]])
            elif k == "vn":
                vn.append([float(x) for x in s[1:4]])
            elif k == "vt":
                vt.append([float(s[1]), 1.0 - float(s[2])])
            elif k == "mtllib":
                mtllib = s[1]
            elif k == "usemtl":
                mats.append((s[1], len(fp)))
            elif k == "f":
                idx = []
                for tok in s[1:]:
                    parts = (tok.split("/") + ["", ""])[:3]
                    def res(t, n):
                        if t == "":
                            return -1
                        i = int(t)
                        return i - 1 if i > 0 else n + i
                    idx.append((res(parts[0], len(v)), res(parts[1], len(vt)), res(parts[2], len(vn))))
                for i in range(1, len(idx) - 1):
                    tri = (idx[0], idx[i], idx[i + 1])
                    fp.append([t[0] for t in tri]); ft.append([t[1] for t in tri]); fn.append([t[2] for t in tri])
    a = lambda x, dt, w: np.asarray(x, dt).reshape(-1, w)
    has_t = len(vt) > 0 and all(i >= 0 for t in ft for i in t)
    has_n = len(vn) > 0 and all(i >= 0 for t in fn for i in t)
    return dict(v_pos=a(v, np.float32, 3), v_nrm=a(vn, np.float32, 3) if len(vn) else None, v_tex=a(vt, np.float32, 2) if len(vt) else None,
                t_pos_idx=a(fp, np.int32, 3), t_nrm_idx=a(fn, np.int32, 3) if has_n else None, t_tex_idx=a(ft, np.int32, 3) if has_t else None,
                mtllib=mtllib, usemtl=mats)


def save_obj(path, v_pos, t_pos_idx, v_nrm=None, t_nrm_idx=None, v_tex=None, t_tex_idx=None, mtllib=None, usemtl=None):
    """render/obj.py:127-176 layout: v / vt (v flipped back) / vn / f with 1-based p/t/n triples."""
    with open(path, "w") as f:
        if mtllib:
            f.write("mtllib %s\n" % mtllib)
        f.write("g default\n")
        for p in np.asarray(v_pos):
            f.write("v %s %s %s\n" % (repr(float(p[0])), repr(float(p[1])), repr(float(p[2]))))
        if v_tex is not None:
            for t in np.asarray(v_tex):
                f.write("vt %s %s\n" % (repr(float(t[0])), repr(float(1.0 - t[1]))))
        if v_nrm is not None:
            for n in np.asarray(v_nrm):
                f.write("vn %s %s %s\n" % (repr(float(n[0])), repr(float(n[1])), repr(float(n[2]))))
        if usemtl:
            f.write("usemtl %s\n" % usemtl)
        tp = np.asarray(t_pos_idx)
        for i in range(tp.shape[0]):
            toks = []
            for c in range(3):
                s = str(int(tp[i, c]) + 1)
                if v_tex is not None and t_tex_idx is not None:
                    s += "/" + str(int(t_tex_idx[i][c]) + 1)
                elif v_nrm is not None and t_nrm_idx is not None:
                    s += "/"
                if v_nrm is not None and t_nrm_idx is not None:
                    s += "/" + str(int(t_nrm_idx[i][c]) + 1)
                toks.append(s)
            f.write("f " + " ".join(toks) + "\n")
