"""Synthetic scenes for parity tests and benchmarks (SURVEY.md section 8d, rows f2/f4 stand-ins).

The reference's assets (data/spot, data/bob, HDR probes) and nvdiffrast are not available on the GPU
box, so everything is procedural and seeded:
  * meshes: displaced icosphere ("blob") + optional torus and ground plane so that shadow rays see
    real occlusion (self-shadowing concavities, inter-object shadows);
  * cameras: the reference's validation orbit (dataset_mesh.py:62-71: perspective 45 deg,
    translate(0,0,-RADIUS) @ rotate_x(-0.4) @ rotate_y(ang), RADIUS = 3, train.py:42);
  * G-buffer: primary rays -> closest hit (own BVH on the GPU, oracle brute force on the CPU) ->
    rast-compatible mask, interpolated position / smooth normal / tangent / face normal;
  * materials: kd = U[0,1)^3, ks = (0, U[0.1,1), U[0,1)) per pixel (configs/bob.json:10-11 ranges);
  * light: U[0.25,0.75) trainable-style probe (light.py:98-101) or a synthetic high-dynamic-range
    "sun + sky + bright windows" probe for the importance-sampling stress case.
Pure numpy: usable from CPU tests, results are uploaded by the caller.
"""
import numpy as np


# ------------------------------------------------------------------------------------------ meshes
def icosphere(level):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(level):
        edge = {}
        verts = list(v)

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in edge:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                edge[key] = len(verts) - 1
            return edge[key]
        nf = []
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v = np.array(verts)
        f = np.array(nf, np.int64)
    return v, f


def _noise_dir(p, rng, octaves=4):
    """Smooth pseudo-noise on the sphere: sum of random low-frequency sinusoids."""
    out = np.zeros(p.shape[0])
    amp, freq = 1.0, 1.5
    for _ in range(octaves):
        k = rng.normal(size=(3, 3)) * freq
        ph = rng.uniform(0, 2 * np.pi, size=3)
        out += amp * np.prod(np.sin(p @ k + ph), axis=1)
        amp *= 0.5
        freq *= 2.0
    return out


def blob_mesh(level=4, seed=5, displacement=0.25, radius=0.8):
    """Closed genus-0 mesh with bumps and concavities: 20*4^level triangles (level 4 -> 5120)."""
    v, f = icosphere(level)
    rng = np.random.default_rng(seed)
    r = radius * (1.0 + displacement * _noise_dir(v, rng))
    return (v * r[:, None]).astype(np.float32), f.astype(np.int32)


def torus_mesh(R=1.15, r=0.12, nu=64, nv=16, tilt=0.5):
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    w = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    U, W = np.meshgrid(u, w, indexing='ij')
    x = (R + r * np.cos(W)) * np.cos(U); y = r * np.sin(W); z = (R + r * np.cos(W)) * np.sin(U)
    p = np.stack([x, y, z], -1).reshape(-1, 3)
    c, s = np.cos(tilt), np.sin(tilt)
    rot = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    p = p @ rot.T
    idx = np.arange(nu * nv).reshape(nu, nv)
    a = idx; b = np.roll(idx, -1, 0); c_ = np.roll(np.roll(idx, -1, 0), -1, 1); d = np.roll(idx, -1, 1)
    f = np.concatenate([np.stack([a, b, c_], -1).reshape(-1, 3), np.stack([a, c_, d], -1).reshape(-1, 3)])
    return p.astype(np.float32), f.astype(np.int32)


def plane_mesh(y=-1.0, half=2.0, n=8):
    g = np.linspace(-half, half, n + 1)
    X, Z = np.meshgrid(g, g, indexing='ij')
    p = np.stack([X, np.full_like(X, y), Z], -1).reshape(-1, 3)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    f = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([a, d, c], -1).reshape(-1, 3)])
    return p.astype(np.float32), f.astype(np.int32)


def merge(*meshes):
    vs, fs, off = [], [], 0
    for v, f in meshes:
        vs.append(v); fs.append(f + off); off += v.shape[0]
    return np.concatenate(vs).astype(np.float32), np.concatenate(fs).astype(np.int32)


def scene_mesh(kind="blob", level=4, seed=5):
    """'blob': displaced icosphere; 'blob+torus': plus a tilted ring (inter-object shadows);
    'full': plus a ground plane."""
    if kind == "grid1m":
        return grid1m_mesh()
    if kind == "bob-like":        # ~11 k triangles, the size of data/bob/bob_tri.obj (10 688): blob + a finely tessellated ring
        return merge(blob_mesh(4, seed), torus_mesh(nu=128, nv=24))
    m = [blob_mesh(level, seed)]
    if kind in ("blob+torus", "full"):
        m.append(torus_mesh())
    if kind == "full":
        m.append(plane_mesh())
    return merge(*m)


def grid1m_mesh(n=724):
    """BASELINE configs[4]-like ~1.08 M triangles: a displaced n x n height field (2 n^2 triangles) with an overhanging ring, so
    that shadow rays see both a huge flat-ish occluder and inter-object shadows."""
    g = np.linspace(-1, 1, n + 1, dtype=np.float32)
    X, Z = np.meshgrid(g, g, indexing="ij")
    Y = (0.25 * np.sin(7 * X) * np.cos(5 * Z) - 0.2).astype(np.float32)
    v = np.stack([X, Y, Z], -1).reshape(-1, 3)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    f = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([a, d, c], -1).reshape(-1, 3)]).astype(np.int32)
    return merge((v, f), torus_mesh(R=0.6, r=0.08, nu=256, nv=64, tilt=0.3))


def vertex_normals(v, f):
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    vn = np.zeros_like(v, dtype=np.float64)
    for k in range(3):
        np.add.at(vn, f[:, k], fn)
    vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-20)
    return vn.astype(np.float32)


# ------------------------------------------------------------------------------------------ cameras
def perspective(fovy=0.7854, aspect=1.0, n=0.1, f=1000.0):
    y = np.tan(fovy / 2)
    return np.array([[1 / (y * aspect), 0, 0, 0], [0, 1 / -y, 0, 0], [0, 0, -(f + n) / (f - n), -(2 * f * n) / (f - n)], [0, 0, -1, 0]], np.float32)


def orbit_view(ang, radius=3.0, tilt=-0.4):
    """Model-view of the reference's validation orbit (dataset_mesh.py:62-71)."""
    cx, sx = np.cos(tilt), np.sin(tilt)
    cy, sy = np.cos(ang), np.sin(ang)
    rx = np.array([[1, 0, 0, 0], [0, cx, sx, 0], [0, -sx, cx, 0], [0, 0, 0, 1]], np.float64)
    ry = np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]], np.float64)
    tr = np.eye(4); tr[2, 3] = -radius
    return tr @ rx @ ry


def primary_rays(mv, res, fovy=0.7854):
    """World-space pinhole rays through pixel centres for model-view `mv`; returns (campos[3], ro[H,W,3], rd[H,W,3])."""
    H = W = res
    inv = np.linalg.inv(mv)
    campos = inv[:3, 3]
    t = np.tan(fovy / 2)
    ys = (1 - 2 * (np.arange(H) + 0.5) / H) * t
    xs = (2 * (np.arange(W) + 0.5) / W - 1) * t
    X, Y = np.meshgrid(xs, ys)
    d_cam = np.stack([X, Y, -np.ones_like(X)], -1)
    d = d_cam @ inv[:3, :3].T
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    ro = np.broadcast_to(campos, d.shape)
    return campos.astype(np.float32), np.ascontiguousarray(ro, np.float32), d.astype(np.float32)


# ------------------------------------------------------------------------------------------ G-buffer
def assemble_gbuffer(v, f, vn, tri_id, tuv, campos, seed=1, ks_mode="random"):
    """tri_id [H,W] int (-1 miss), tuv [H,W,3] = (t,u,v) -> dict of [H,W,*] fp32 arrays in the layout
    render.shade() hands to optix_env_shade (render/render.py:99-115)."""
    H, W = tri_id.shape
    hit = tri_id >= 0
    tid = np.where(hit, tri_id, 0)
    u, w = tuv[..., 1:2], tuv[..., 2:3]
    b0 = 1 - u - w
    i0, i1, i2 = f[tid, 0], f[tid, 1], f[tid, 2]
    pos = b0 * v[i0] + u * v[i1] + w * v[i2]
    nrm = b0 * vn[i0] + u * vn[i1] + w * vn[i2]
    gn = np.cross(v[i1] - v[i0], v[i2] - v[i0])
    gn /= np.maximum(np.linalg.norm(gn, axis=-1, keepdims=True), 1e-20)
    tng = np.cross(np.array([0.0, 1.0, 0.0], np.float32), nrm)
    tng = np.where(np.linalg.norm(tng, axis=-1, keepdims=True) > 1e-6, tng, np.array([1.0, 0, 0], np.float32))
    rng = np.random.default_rng(seed)
    kd = rng.uniform(0, 1, size=(H, W, 3))
    if ks_mode == "metal":        # data/spot/metal.mtl: ks = (0, 0.2, 1)
        ks = np.broadcast_to(np.array([0.0, 0.2, 1.0]), (H, W, 3)).copy()
    else:
        ks = np.stack([np.zeros((H, W)), rng.uniform(0.1, 1.0, size=(H, W)), rng.uniform(0, 1, size=(H, W))], -1)
    m = hit[..., None].astype(np.float32)
    depth = np.where(hit, tuv[..., 0], 0.0)
    out = dict(mask=(tri_id + 1).astype(np.float32) * hit,       # rast[...,-1] = triangle id + 1, 0 = background
               pos=pos * m, smooth_nrm=nrm * m, tangent=tng * m, geom_nrm=gn * m, kd=kd * m, ks=ks * m,
               view_pos=np.asarray(campos, np.float32), depth=depth)
    return {k: np.ascontiguousarray(a, np.float32) for k, a in out.items()}


# ------------------------------------------------------------------------------------------ lights
def random_light(res=256, seed=2, scale=0.5, bias=0.25):
    """create_trainable_env_rnd (light.py:98-101): U[bias, bias+scale)."""
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, 1, size=(res, res, 3)) * scale + bias).astype(np.float32)


def hdr_light(H=256, W=512, seed=7):
    """High-dynamic-range lat-long probe: smooth sky gradient + a small very bright sun + a few
    bright 'windows' -- the high-frequency importance-sampling case (stand-in for
    data/irrmaps/aerodynamics_workshop_2k.hdr)."""
    rng = np.random.default_rng(seed)
    ys = (np.arange(H) + 0.5) / H
    xs = (np.arange(W) + 0.5) / W
    X, Y = np.meshgrid(xs, ys)
    sky = np.stack([0.3 + 0.4 * (1 - Y), 0.4 + 0.4 * (1 - Y), 0.6 + 0.5 * (1 - Y)], -1) * (Y < 0.5)[..., None]
    ground = np.stack([0.15 * np.ones_like(Y)] * 3, -1) * (Y >= 0.5)[..., None]
    img = sky + ground
    sun = np.exp(-(((X - 0.3) * 2) ** 2 + (Y - 0.22) ** 2) / (2 * 0.008 ** 2))
    img += sun[..., None] * np.array([900.0, 800.0, 600.0])
    for _ in range(6):
        cx, cy, sx, sy = rng.uniform(0, 1), rng.uniform(0.25, 0.6), rng.uniform(0.01, 0.04), rng.uniform(0.01, 0.03)
        win = ((np.abs(X - cx) < sx) & (np.abs(Y - cy) < sy)).astype(np.float64)
        img += win[..., None] * rng.uniform(5, 40, size=3)
    return np.maximum(img, 1e-4).astype(np.float32)


def make_perms(n_samples_x, seed=3, rows=32768):
    """Permutation table with the reference's shape/dtype (ops.py:84-86) from a seeded generator."""
    rng = np.random.default_rng(seed)
    S = n_samples_x * n_samples_x
    return np.argsort(rng.random((rows, S)), axis=-1).astype(np.int32)
