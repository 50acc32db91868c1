"""On-disk formats used by the hot path's fixtures, without imageio / freeimage:

  * Radiance RGBE `.hdr` reader / writer  -- the reference goes through imageio (render/util.py:355-384; light.py:71-79 load_env,
    light.py:89-93 save_env_map);
  * Wavefront `.obj` reader / writer (positions, normals, texcoords, triangulated faces) -- render/obj.py:31-176;
  * Wavefront `.mtl` reader / writer (statement level) -- render/material.py:21-95;
  * PNG reader / writer (8 / 16 bit, grey / grey+alpha / RGB / RGBA, all five scanline filters) -- texture maps, render/util.py:355-376.

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


# ------------------------------------------------------------------------------------------------ Wavefront MTL
_MTL_PATHS = ("bsdf", "map_kd", "map_ks", "bump", "map_bump", "map_d", "map_ka", "map_ke")


def load_mtl(path):
    """-> list of dicts, one per `newmtl` block, keys lower-cased: 'name', path-valued statements (`bsdf`, `map_kd`, `map_ks`, `bump`,
    kept as strings relative to the .mtl) and numeric statements (`kd`, `ks`, `ka`, `ns`, ...) as float32 arrays.  'bsdf' defaults to
    'pbr' (render/material.py:21-48; the conversion of constants / image maps into mip-mapped textures, :50-69, stays with the caller)."""
    materials = []
    with open(path, "r") as f:
        for raw in f:
            line = raw.split("#", 1)[0].split()
            if not line:
                continue
            key, data = line[0].lower(), line[1:]
            if key == "newmtl":
                materials.append({"name": data[0] if data else ""})
            elif materials and data:
                m = materials[-1]
                if key in _MTL_PATHS:
                    m[key] = data[-1]                       # options such as `-bm 1.0` precede the file name
                else:
                    try:
                        m[key] = np.asarray([float(d) for d in data], np.float32)
                    except ValueError:
                        m[key] = " ".join(data)
    for m in materials:
        m.setdefault("bsdf", "pbr")
    return materials


def save_mtl(path, materials):
    """Inverse of load_mtl for the statements render/material.py:75-95 writes (name, bsdf, map_* / bump paths, numeric constants)."""
    with open(path, "w") as f:
        for m in materials:
            f.write("newmtl %s\n" % m.get("name", "defaultMat"))
            for k, v in m.items():
                if k == "name":
                    continue
                if isinstance(v, str):
                    f.write("%s %s\n" % (k if k in _MTL_PATHS else k, v))
                else:
                    f.write("%s %s\n" % (k, " ".join("%.9g" % float(x) for x in np.ravel(v))))
            f.write("\n")


# ------------------------------------------------------------------------------------------------ PNG (zlib only)
_PNG_SIG = b"\x89PNG\r\n\x1a\n"
_PNG_CHANNELS = {0: 1, 2: 3, 4: 2, 6: 4}


def _paeth(a, b, c):
    p = a.astype(np.int32) + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c)).astype(np.uint8)


def load_png(path):
    """-> uint8 or uint16 array [H, W] / [H, W, C] (grey, grey+alpha, RGB, RGBA; 8 or 16 bit; non-interlaced).  The reference reads
    texture maps through imageio (render/util.py:355-367: `load_image_raw`); callers divide by 255 / 65535 as it does."""
    import struct, zlib
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != _PNG_SIG:
        raise ValueError("%s: not a PNG file" % path)
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if zlib.crc32(typ + body) & 0xFFFFFFFF != struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]:
            raise ValueError("%s: corrupt %s chunk" % (path, typ.decode("latin1")))
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        pos += 12 + n
    if hdr is None:
        raise ValueError("%s: missing IHDR" % path)
    W, H, depth, ctype, comp, filt, interlace = hdr
    if ctype not in _PNG_CHANNELS or depth not in (8, 16) or interlace != 0:
        raise ValueError("%s: unsupported PNG variant (colour type %d, depth %d, interlace %d)" % (path, ctype, depth, interlace))
    C = _PNG_CHANNELS[ctype]
    bpp = C * depth // 8
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8)
    stride = W * bpp
    if raw.size != H * (stride + 1):
        raise ValueError("%s: truncated image data" % path)
    rows = raw.reshape(H, stride + 1)
    out = np.zeros((H, stride), np.uint8)
    prev = np.zeros(stride, np.uint8)
    for y in range(H):
        ft, line = int(rows[y, 0]), rows[y, 1:]
        if ft == 0:
            cur = line.copy()
        elif ft == 2:
            cur = line + prev
        elif ft in (1, 3, 4):
            # left-dependent filters: sequential over pixels, vectorised over the bytes of a pixel
            cur = np.zeros(stride, np.uint8)
            lp = line.reshape(W, bpp); pp = prev.reshape(W, bpp); cp = cur.reshape(W, bpp)
            left = np.zeros(bpp, np.uint8); upleft = np.zeros(bpp, np.uint8)
            for x in range(W):
                if ft == 1:
                    pred = left
                elif ft == 3:
                    pred = ((left.astype(np.int32) + pp[x]) >> 1).astype(np.uint8)
                else:
                    pred = _paeth(left, pp[x], upleft)
                cp[x] = lp[x] + pred
                left, upleft = cp[x], pp[x]
        else:
            raise ValueError("%s: bad filter type %d" % (path, ft))
        out[y] = cur
        prev = cur
    if depth == 16:
        img = out.reshape(H, W, C, 2)
        img = (img[..., 0].astype(np.uint16) << 8) | img[..., 1]
    else:
        img = out.reshape(H, W, C)
    return img[..., 0] if C == 1 else img


def save_png(path, img):
    """uint8 / uint16 [H,W] or [H,W,{1,2,3,4}] -> PNG (filter 0, zlib level 6).  Floats are taken as [0,1] and written as 8 bit
    (render/util.py:369-376 `save_image`: clip, *255, round)."""
    import struct, zlib
    a = np.asarray(img)
    if a.dtype.kind == "f":
        a = np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    H, W, C = a.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[C]
    if a.dtype == np.uint16:
        depth = 16
        b = np.stack([(a >> 8).astype(np.uint8), (a & 255).astype(np.uint8)], -1).reshape(H, W * C * 2)
    elif a.dtype == np.uint8:
        depth = 8
        b = a.reshape(H, W * C)
    else:
        raise TypeError("save_png: unsupported dtype %s" % a.dtype)
    raw = np.concatenate([np.zeros((H, 1), np.uint8), b], 1).tobytes()

    def chunk(typ, body):
        return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(_PNG_SIG + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
