// bvh8.cuh -- 8-wide compressed BVH (CWBVH-style, Ylitie/Karras/Laine 2017).  EXPERIMENT, off by default (MCS_BVH8=0):
// bit-exact and parity-green, but on B200 the fused kernel was 15 % slower with it than with binary nodes + deferred leaf
// tests (profiles/r01_bvh8_envshade_summary.json: long-scoreboard stalls fell from 35 % to 13 % as intended, but the 600-
// instruction node test made 45 % of the stall samples instruction-fetch misses).  Kept for round 2 (rolled child loop).
//
// Why (profiles/r01_v3_*): with 64-byte binary nodes the node array of even a 7k-triangle mesh (460 KB)
// does not fit in L1; the trace loop's top stall was the L2 round trip of every node fetch (long
// scoreboard ~46 % of its samples).  An 8-wide node with 8-bit child boxes quantised on a per-node grid
// is 80 bytes per EIGHT children: the whole hierarchy of that mesh is ~45 KB (L1 resident) and a ray needs
// ~3x fewer dependent node fetches.
//
// Node layout (5 x 16 bytes):
//   q0: p.x, p.y, p.z (grid origin = node box min), {ex, ey, ez, imask} (biased exponents of the per-axis cell size,
//       bit i of imask = child slot i is an internal node)
//   q1: child_base (index of the first internal child; internal children are contiguous in slot order),
//       tri_base (first triangle of this node's leaf children), meta[0..3], meta[4..7]
//       meta: 0 = empty slot; internal: 0b001_00000 | (24 + slot); leaf: (unary triangle count: 1->001, 2->011, 3->111) << 5 | offset
//   q2: qlo_x[0..7], qlo_y[0..7]      q3: qlo_z[0..7], qhi_x[0..7]      q4: qhi_y[0..7], qhi_z[0..7]
// A node visit yields a 32-bit hit mask: bits 24..31 = internal children hit (by slot), bits 0..23 = triangles of the
// leaf children hit (relative to tri_base).  Conservative by construction: child boxes are rounded outward onto the
// grid, slab comparison relaxed by 4 ulp, and the (already padded) binary boxes are what gets quantised.
#pragma once
#include "bvh_traverse.cuh"

struct Bvh8View {
    const float4 *nodes;     // 5 float4 per node
    const float4 *tris;      // 3 float4 per triangle, wide-leaf order
};

struct Ray8 {
    float ix, iy, iz;        // 1/d (zero components nudged)
    float ox, oy, oz;        // o * (1/d)
    bool nx, ny, nz;         // direction component negative
};
__device__ __forceinline__ Ray8 ray8_pre(f3 o, f3 d)
{
    const RayPre r = ray_pre(o, d);
    Ray8 q;
    q.ix = r.ix; q.iy = r.iy; q.iz = r.iz; q.ox = r.ox; q.oy = r.oy; q.oz = r.oz;
    q.nx = r.ix < 0.0f; q.ny = r.iy < 0.0f; q.nz = r.iz < 0.0f;
    return q;
}

// byte k of `w` -> float(32768 + byte) with ONE byte-permute (0x47000000 | byte << 8); the 32768 is folded into the FMA addend
template <int K>
__device__ __forceinline__ float q2f(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504u | (K << 4))); }

template <int C>
__device__ __forceinline__ uint32_t child_hit(uint32_t lx, uint32_t hx, uint32_t ly, uint32_t hy, uint32_t lz, uint32_t hz, uint32_t meta,
                                              float ax, float bx, float ay, float by, float az, float bz)
{
    const float t0x = fmaf(q2f<C & 3>(lx), ax, bx), t1x = fmaf(q2f<C & 3>(hx), ax, bx);
    const float t0y = fmaf(q2f<C & 3>(ly), ay, by), t1y = fmaf(q2f<C & 3>(hy), ay, by);
    const float t0z = fmaf(q2f<C & 3>(lz), az, bz), t1z = fmaf(q2f<C & 3>(hz), az, bz);
    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, MCS_TMAX));
    const uint32_t mb = (meta >> (8 * (C & 3))) & 0xFFu;
    const uint32_t bits = (mb >> 5) << (mb & 31u);
    return tn <= tf * 1.0000004f ? bits : 0u;
}

// Intersect the 8 children of node `idx`.  Returns the hit mask; child_base / tri_base / imask via references.
__device__ __forceinline__ uint32_t bvh8_visit(const float4 *__restrict__ nodes, int idx, const Ray8 &r, uint32_t &child_base, uint32_t &tri_base, uint32_t &imask)
{
    const float4 *n = nodes + 5 * (size_t)idx;
    const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3), q4 = __ldg(n + 4);
    const uint32_t e = __float_as_uint(q0.w);
    imask = e >> 24;
    child_base = __float_as_uint(q1.x); tri_base = __float_as_uint(q1.y);
    const uint32_t m0 = __float_as_uint(q1.z), m1 = __float_as_uint(q1.w);
    // t = (p + q*cell - o) * idir = q * (cell*idir) + (p*idir - o*idir); q enters as (32768 + q)
    const float ax = __uint_as_float((e & 0xFFu) << 23) * r.ix, ay = __uint_as_float(((e >> 8) & 0xFFu) << 23) * r.iy,
                az = __uint_as_float(((e >> 16) & 0xFFu) << 23) * r.iz;
    const float bx = fmaf(-32768.0f, ax, fmaf(q0.x, r.ix, -r.ox)), by = fmaf(-32768.0f, ay, fmaf(q0.y, r.iy, -r.oy)),
                bz = fmaf(-32768.0f, az, fmaf(q0.z, r.iz, -r.oz));
    // entry / exit planes per axis depend on the ray direction sign: swap the lo/hi byte words once per node
    const uint32_t lx0 = __float_as_uint(q2.x), lx1 = __float_as_uint(q2.y), ly0 = __float_as_uint(q2.z), ly1 = __float_as_uint(q2.w);
    const uint32_t lz0 = __float_as_uint(q3.x), lz1 = __float_as_uint(q3.y), hx0 = __float_as_uint(q3.z), hx1 = __float_as_uint(q3.w);
    const uint32_t hy0 = __float_as_uint(q4.x), hy1 = __float_as_uint(q4.y), hz0 = __float_as_uint(q4.z), hz1 = __float_as_uint(q4.w);
    const uint32_t nx0 = r.nx ? hx0 : lx0, nx1 = r.nx ? hx1 : lx1, fx0 = r.nx ? lx0 : hx0, fx1 = r.nx ? lx1 : hx1;
    const uint32_t ny0 = r.ny ? hy0 : ly0, ny1 = r.ny ? hy1 : ly1, fy0 = r.ny ? ly0 : hy0, fy1 = r.ny ? ly1 : hy1;
    const uint32_t nz0 = r.nz ? hz0 : lz0, nz1 = r.nz ? hz1 : lz1, fz0 = r.nz ? lz0 : hz0, fz1 = r.nz ? lz1 : hz1;
    uint32_t hits = 0;
    hits |= child_hit<0>(nx0, fx0, ny0, fy0, nz0, fz0, m0, ax, bx, ay, by, az, bz);
    hits |= child_hit<1>(nx0, fx0, ny0, fy0, nz0, fz0, m0, ax, bx, ay, by, az, bz);
    hits |= child_hit<2>(nx0, fx0, ny0, fy0, nz0, fz0, m0, ax, bx, ay, by, az, bz);
    hits |= child_hit<3>(nx0, fx0, ny0, fy0, nz0, fz0, m0, ax, bx, ay, by, az, bz);
    hits |= child_hit<4>(nx1, fx1, ny1, fy1, nz1, fz1, m1, ax, bx, ay, by, az, bz);
    hits |= child_hit<5>(nx1, fx1, ny1, fy1, nz1, fz1, m1, ax, bx, ay, by, az, bz);
    hits |= child_hit<6>(nx1, fx1, ny1, fy1, nz1, fz1, m1, ax, bx, ay, by, az, bz);
    hits |= child_hit<7>(nx1, fx1, ny1, fy1, nz1, fz1, m1, ax, bx, ay, by, az, bz);
    return hits;
}

#define MCS_STACK8 24

// Traversal state of one ray: current node group (base index, [hit bits 31..24 | imask 7..0]) + stack of groups.
struct Trav8 {
    uint2 ng;
    int sp;
    __device__ __forceinline__ void start() { ng = make_uint2(0u, 0x01000001u); sp = 0; }     // "slot 0 of a virtual parent" = root node 0
};

// Pops / selects the next internal child to visit.  Returns false when the walk is finished.
__device__ __forceinline__ bool trav8_next(Trav8 &t, uint2 *stack, int &node_idx)
{
    if (!(t.ng.y & 0xFF000000u)) {
        if (t.sp == 0) return false;
        t.ng = stack[--t.sp];
    }
    const int b = __ffs((int)(t.ng.y >> 24)) - 1;            // child slot
    t.ng.y &= ~(1u << (b + 24));
    node_idx = (int)(t.ng.x + __popc(t.ng.y & 0xFFu & ((1u << b) - 1u)));
    if (t.ng.y & 0xFF000000u) stack[t.sp++] = t.ng;
    return true;
}

// Any-hit query (immediate triangle tests): used by the stand-alone visibility kernel.
__device__ __forceinline__ bool bvh8_occluded(const Bvh8View &b, f3 o, f3 d)
{
    const Ray8 r = ray8_pre(o, d);
    uint2 stack[MCS_STACK8];
    Trav8 t;
    t.start();
    int idx;
    while (trav8_next(t, stack, idx)) {
        uint32_t cb, tb, im;
        const uint32_t hits = bvh8_visit(b.nodes, idx, r, cb, tb, im);
        t.ng = make_uint2(cb, (hits & 0xFF000000u) | im);
        uint32_t tm = hits & 0x00FFFFFFu;
        while (tm) {
            const int k = __ffs((int)tm) - 1;
            tm &= tm - 1;
            const float4 *tp = b.tris + 3 * (size_t)(tb + k);
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            float tt, uu, vv;
            if (mt_hit(o, d, F3(t0.x, t0.y, t0.z), F3(t1.x, t1.y, t1.z), F3(t2.x, t2.y, t2.z), MCS_TMAX, tt, uu, vv)) return true;
        }
    }
    return false;
}
