// sr_device.cuh -- device functions shared by the fused render kernel (sr_render.cu) and the stage kernels
// (sr_wave.cu): the reference-arithmetic layer, the FP32 candidate search and the filtered predicates.
// (Moved verbatim out of sr_render.cu; DESIGN.md sections 1-3 describe the numerics contract.)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "sr_types.h"

namespace sr {

// ---------------------------------------------------------------------------------------------
// exact FP64 vector helpers (no contraction, left-to-right like Engine3D/Vector.cs)
// ---------------------------------------------------------------------------------------------
// Phase synchronisation (profiles/r01s_instruction_fetch_findings.md).  The SM's instruction cache holds ~2 K
// instructions and the camera-ray path is several times that, so a frame whose time goes into camera rays is bound
// by instruction fetch.  While `sync` is on, the warps of a block fetch their tiles together and meet at barriers
// between the stages of a camera ray (sphere search | exact spheres | mesh search | root-box clip + exact triangles |
// shading), so they run the same few hundred instructions at about the same time and share the fetched lines.
// Measured on one B200 (ms per frame, free-running -> synchronised, 256-thread blocks): config2 0.767 -> 0.700;
// but config3 48.3 -> 51.4 and config5 17.1 -> 19.2 (long walks of very different lengths: the warps wait at the
// barriers longer than the shared fetches save), so the host turns it on for small scenes only (sr_api.cu).
// The flag is uniform over the launch; every thread of a block reaches every SR_SYNC_POINT while it is on.
#ifndef SR_THREADS
#define SR_THREADS 256
#endif
// closest_hit used to be out of line (code size): its Hit result and counters then travel through local memory.
// Inlined: config2 0.638 -> 0.590, config3 47.0 -> 46.0, config4 222 -> 181, config5 16.5 -> 15.1 ms.
#ifndef SR_CH_INLINE
#define SR_CH_INLINE __forceinline__
#endif
// ... and so did the exact evaluators' results (BestPrim, clipped start, offset), once they had a single call
// site each: config2 0.582 -> 0.537, config4 173 -> 166 ms.
#ifndef SR_EX_INLINE
#define SR_EX_INLINE __forceinline__
#endif
#ifndef SR_PREFETCH
#define SR_PREFETCH 1     // prefetch the far child when both are hit: config3 47.6 -> 46.6, config4 223 -> 218 ms
#endif
#define SR_SYNC_POINT(on) do { if (on) __syncthreads(); } while (0)

struct d3 { double x, y, z; };

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ d3 mk(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 vadd(d3 a, d3 b) { return mk(dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)); }
__device__ __forceinline__ d3 vsub(d3 a, d3 b) { return mk(dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)); }
__device__ __forceinline__ d3 vscale(d3 a, double s) { return mk(dmul(a.x, s), dmul(a.y, s), dmul(a.z, s)); }
__device__ __forceinline__ d3 vneg(d3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ double vdot(d3 a, d3 b)
{
    return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z));   // Vector.cs:99-102
}
__device__ __forceinline__ double vlen(d3 a) { return __dsqrt_rn(vdot(a, a)); }   // Vector.cs:121-128
__device__ __forceinline__ d3 vnormalise(d3 a)                                    // Vector.cs:177-185
{
    double inv = ddiv(1.0, vlen(a));
    return vscale(a, inv);
}
// Matrix.Multiply3X3 / TransformDirection[Reverse] on a 3x4 row-major block (Matrix.cs:34-41)
__device__ __forceinline__ d3 mul3x3(const double* m, d3 v)
{
    return mk(dadd(dadd(dmul(v.x, m[0]), dmul(v.y, m[1])), dmul(v.z, m[2])),
              dadd(dadd(dmul(v.x, m[4]), dmul(v.y, m[5])), dmul(v.z, m[6])),
              dadd(dadd(dmul(v.x, m[8]), dmul(v.y, m[9])), dmul(v.z, m[10])));
}
// Matrix.Multiply3X4 (Matrix.cs:50-57)
__device__ __forceinline__ d3 mul3x4(const double* m, d3 v)
{
    return mk(dadd(dadd(dadd(dmul(v.x, m[0]), dmul(v.y, m[1])), dmul(v.z, m[2])), m[3]),
              dadd(dadd(dadd(dmul(v.x, m[4]), dmul(v.y, m[5])), dmul(v.z, m[6])), m[7]),
              dadd(dadd(dadd(dmul(v.x, m[8]), dmul(v.y, m[9])), dmul(v.z, m[10])), m[11]));
}

// Color.ModulatePackedColor (Color.cs:124-133)
__device__ __forceinline__ uint32_t modulate(uint32_t c, uint32_t amount)
{
    uint32_t r = (((c >> 16) & 0xff) * amount) >> 8;
    uint32_t g = (((c >> 8) & 0xff) * amount) >> 8;
    uint32_t b = ((c & 0xff) * amount) >> 8;
    return 0xff000000u | (r << 16) | (g << 8) | b;
}
// C# (byte)double for an in-range value: truncation
__device__ __forceinline__ uint32_t to_byte(double v) { return (uint32_t)__double2int_rz(v) & 0xffu; }

// Texture3D extension, id 1: Texture3DCache index quantisation (Texture3DCache.cs:98-100)
// with N = 128, clamped, then an integer pattern.
__device__ __forceinline__ uint32_t texture3d_sample(int id, d3 p)
{
    if (id != 1) return 255u;
    const double n1 = 127.0;
    int qx = __double2int_rz(dmul(dadd(p.x, 0.5), n1));
    int qy = __double2int_rz(dmul(dadd(p.y, 0.5), n1));
    int qz = __double2int_rz(dmul(dadd(p.z, 0.5), n1));
    qx = min(max(qx, 0), 127); qy = min(max(qy, 0), 127); qz = min(max(qz, 0), 127);
    int cell = ((qx >> 3) ^ (qy >> 3) ^ (qz >> 3)) & 1;
    int grain = (qx * 3 + qy * 5 + qz * 7) & 31;
    return (uint32_t)(255 - cell * 80 - grain);
}
__device__ __forceinline__ uint32_t mirror_blend(uint32_t local, uint32_t refl)
{
    uint32_t r = (3u * ((local >> 16) & 0xff) + ((refl >> 16) & 0xff)) >> 2;
    uint32_t g = (3u * ((local >> 8) & 0xff) + ((refl >> 8) & 0xff)) >> 2;
    uint32_t b = (3u * (local & 0xff) + (refl & 0xff)) >> 2;
    return 0xff000000u | (r << 16) | (g << 8) | b;
}

struct Counters {
    unsigned int node_visits, prim_tests, sphere_tests, shaded, filter_tests, filter_unsure, filter_mismatch, bundled;
    int bundle_skip;     // shading points this thread has seen (its periodic cone-walk probe, shade_and_shadow)
    int* stack;          // the thread's one traversal stack (kStackEntries), shared by every BVH walk
};
// Counters of the out-of-line (exact / per-camera-ray) functions.  These are reached through a pointer,
// so they live in local memory; each function accumulates in registers and adds once on return.  Keeping
// them apart leaves the hot shadow-ray counters above in registers.
struct XCounters {
    unsigned int node_visits, prim_tests, sphere_tests, filter_tests, filter_unsure, filter_mismatch;
    int* stack;          // same stack as Counters::stack
};

// ---------------------------------------------------------------------------------------------
// exact primitive tests
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ldg2(const void* p, int i)
{
    return __ldg(reinterpret_cast<const double2*>(p) + i);
}

// Triangle.IntersectRay + Plane.IntersectRay (Triangle.cs:83-104, Plane.cs:67-103).
// `limit`: hits with rayFrac > limit are of no use to the caller (they lose the strict `<`).
__device__ __forceinline__ bool tri_intersect(const TriRec* __restrict__ t, d3 s, d3 dir, double limit, double* rf_out)
{
    const double2 a0 = ldg2(t, 0), a1 = ldg2(t, 1);          // n.x n.y | n.z d
    const d3 n = mk(a0.x, a0.y, a1.x);
    const double start_dist = vdot(s, n);
    const double dir_dist = vdot(dir, n);
    if (dir_dist >= 0.0) return false;                        // one-sided
    double rf = dsub(a1.y, start_dist);
    if (!(rf <= 0.0)) return false;
    rf = ddiv(rf, dir_dist);
    if (rf > limit) return false;
    const d3 pos = vadd(s, vscale(dir, rf));
    const double2 a2 = ldg2(t, 2), a3 = ldg2(t, 3);          // v1.x v1.y | v1.z den1
    const d3 w = vsub(pos, mk(a2.x, a2.y, a3.x));
    const double2 a4 = ldg2(t, 4), a5 = ldg2(t, 5);          // e2p.x e2p.y | e2p.z den2
    const double sN = ddiv(vdot(w, mk(a4.x, a4.y, a5.x)), a3.y);
    if (sN < 0.0 || sN > 1.0) return false;
    const double2 a6 = ldg2(t, 6), a7 = ldg2(t, 7);          // e1p.x e1p.y | e1p.z (color,index)
    const double u = ddiv(vdot(w, mk(a6.x, a6.y, a7.x)), a5.y);
    if (sN >= 0.0 && u >= 0.0 && dadd(sN, u) <= 1.0) { *rf_out = rf; return true; }
    return false;
}

// Sphere.IntersectRay up to rayFrac (Sphere.cs:152-192); dirn = dir after Vector.Normalise.
__device__ __forceinline__ bool sphere_intersect(const SphereRec* __restrict__ sp, d3 s, d3 dirn, double* rf_out)
{
    const double2 a0 = ldg2(sp, 0), a1 = ldg2(sp, 1), a2 = ldg2(sp, 2);   // c.x c.y | c.z r | r2 (color,index)
    const d3 o = vsub(s, mk(a0.x, a0.y, a1.x));
    const double proj = vdot(o, dirn);
    if (proj > a1.y) return false;
    const double dist_sqr = vdot(o, o);
    const double term = dadd(dsub(dmul(proj, proj), dist_sqr), a2.x);
    if (term < 1e-10) return false;
    const double root = __dsqrt_rn(term);
    const double f1 = dsub(-proj, root);
    const double f2 = dadd(-proj, root);
    const double rf = (f1 >= 0.0) ? f1 : f2;
    if (rf < 0.0) return false;
    *rf_out = rf;
    return true;
}

// ---------------------------------------------------------------------------------------------
// AxisAlignedBox.ContainsPoint / ClipLineSegment (AxisAlignedBox.cs:143-149,175-216) as used by
// SpatialSubdivision.IntersectRay (SpatialSubdivision.cs:389-401).  The six Plane objects have
// exact axis unit normals, so Plane.IntersectLineSegment (Plane.cs:111-138) reduces, bit for bit,
// to the per-axis quotients below (min planes first, then max planes: AxisAlignedBox.cs:22-27).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool box_contains(const double* mn, const double* mx, d3 p)
{
    const double e = 1e-10;
    return dsub(mn[0], e) < p.x && p.x < dadd(mx[0], e) && dsub(mn[1], e) < p.y && p.y < dadd(mx[1], e) &&
           dsub(mn[2], e) < p.z && p.z < dadd(mx[2], e);
}

__device__ __forceinline__ bool box_first_crossing(const double* mn, const double* mx, d3 s, d3 e, d3* pos)
{
    double closest = 1.7976931348623157e308;
    const d3 span = vsub(e, s);
    const double sv[3] = {s.x, s.y, s.z}, ev[3] = {e.x, e.y, e.z};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int ax = k % 3;
        // min plane k<3: (s - min) / (s - e);  max plane: (max - s) / (e - s)
        const double num = (k < 3) ? dsub(sv[ax], mn[ax]) : dsub(mx[ax], sv[ax]);
        const double den = (k < 3) ? dsub(sv[ax], ev[ax]) : dsub(ev[ax], sv[ax]);
        const double lf = ddiv(num, den);
        if (0.0 <= lf && lf <= 1.0 && lf < closest) {
            const d3 p = vadd(s, vscale(span, lf));
            if (box_contains(mn, mx, p)) { closest = lf; *pos = p; }
        }
    }
    return closest != 1.7976931348623157e308;
}

// Returns false when the ray misses the root box (ClippedRayCount++).  On success *start is the
// clipped start and *offset the rayFrac offset of :401.
__device__ __forceinline__ bool reference_clip(const double* mn, const double* mx, d3* start, d3 dir, double* offset)
{
    const d3 s = *start;
    const d3 end = vadd(s, vscale(dir, 10000.0));
    const bool start_inside = box_contains(mn, mx, s);
    const bool end_inside = box_contains(mn, mx, end);
    *offset = 0.0;
    if (start_inside && end_inside) return true;
    d3 p;
    if (!box_first_crossing(mn, mx, s, end, &p)) return false;
    if (start_inside) return true;
    *start = p;
    *offset = ddiv(vlen(vsub(s, p)), vlen(dir));   // originalStart.Distance(start) / dir.Length
    return true;
}

// reference_clip when the answer is obvious: the start lies outside the box and the ray enters through the
// interior of ONE face.  ClipLineSegment's loop (AxisAlignedBox.cs:175-216) then keeps exactly that face's crossing:
// of the other five candidates, the far faces lie further along the segment and the other near faces are crossed
// where the ray is still outside the box on the entry axis -- all by `margin` in space, >> the 1e-10 of
// ContainsPoint (AxisAlignedBox.cs:143-149) and >> the error of the FP32 estimate that decides this (~1e-6 of the
// operands).  The crossing itself is computed with the reference's operations for that one plane, so *start and
// *offset are bit-identical to reference_clip's (1 FP64 division instead of 7, no 6 x ContainsPoint).  false: not
// that clear -- run reference_clip.  (SR_CLIP_CHECK builds run both and count differences as filter mismatches.)
__device__ __forceinline__ bool reference_clip_face(const double* mn, const double* mx, d3* start, d3 dir, double* offset)
{
    const d3 s = *start;
    const float sx = (float)s.x, sy = (float)s.y, sz = (float)s.z;
    const float gx = (float)dir.x * 10000.0f, gy = (float)dir.y * 10000.0f, gz = (float)dir.z * 10000.0f;   // the segment's span
    const float agx = fabsf(gx), agy = fabsf(gy), agz = fabsf(gz);
    if (!(fminf(agx, fminf(agy, agz)) > 1e-20f) || !(fmaxf(agx, fmaxf(agy, agz)) < 1e30f)) return false;
    const float lox = (float)mn[0], loy = (float)mn[1], loz = (float)mn[2], hix = (float)mx[0], hiy = (float)mx[1], hiz = (float)mx[2];
    const float ix = 1.0f / gx, iy = 1.0f / gy, iz = 1.0f / gz;
    const float ax = (lox - sx) * ix, bx = (hix - sx) * ix, ay = (loy - sy) * iy, by = (hiy - sy) * iy,
                az = (loz - sz) * iz, bz = (hiz - sz) * iz;
    const float nx = fminf(ax, bx), fx = fmaxf(ax, bx), ny = fminf(ay, by), fy = fmaxf(ay, by), nz = fminf(az, bz), fz = fmaxf(az, bz);
    const float t_in = fmaxf(nx, fmaxf(ny, nz));
    const float big = fmaxf(fmaxf(fabsf(sx), fabsf(sy)), fmaxf(fabsf(sz), fmaxf(fmaxf(fabsf(lox), fabsf(hix)),
                            fmaxf(fmaxf(fabsf(loy), fabsf(hiy)), fmaxf(fabsf(loz), fabsf(hiz))))));
    const float margin = 1e-4f * big;
    if (!(t_in < 0.999f)) return false;
    int axis;                  // entry axis; its near plane is the min plane when the span is positive
    float ga;
    if (t_in == nx) { axis = 0; ga = agx; } else if (t_in == ny) { axis = 1; ga = agy; } else { axis = 2; ga = agz; }
    // the start is outside by `margin` on the entry axis, the box is not flat there
    if (!(t_in * ga > margin)) return false;
    if (!(((axis == 0 ? fx : axis == 1 ? fy : fz) - t_in) * ga > margin)) return false;
    // on the other two axes: the entry point lies inside the slab by `margin`, and the slab's near plane is crossed
    // while the ray is still `margin` outside the box on the entry axis
    if (axis != 0 && (!((t_in - nx) * agx > margin) || !((fx - t_in) * agx > margin) || !((t_in - nx) * ga > margin))) return false;
    if (axis != 1 && (!((t_in - ny) * agy > margin) || !((fy - t_in) * agy > margin) || !((t_in - ny) * ga > margin))) return false;
    if (axis != 2 && (!((t_in - nz) * agz > margin) || !((fz - t_in) * agz > margin) || !((t_in - nz) * ga > margin))) return false;
    // the reference's arithmetic for that plane (box_first_crossing, k = axis or axis + 3)
    const d3 end = vadd(s, vscale(dir, 10000.0));
    const d3 span = vsub(end, s);
    const double sa = axis == 0 ? s.x : axis == 1 ? s.y : s.z;
    const double ea = axis == 0 ? end.x : axis == 1 ? end.y : end.z;
    const double da = axis == 0 ? dir.x : axis == 1 ? dir.y : dir.z;
    const bool min_plane = da > 0.0;
    const double num = min_plane ? dsub(sa, mn[axis]) : dsub(mx[axis], sa);
    const double den = min_plane ? dsub(sa, ea) : dsub(ea, sa);
    const double lf = ddiv(num, den);
    if (!(0.0 <= lf && lf <= 1.0)) return false;              // (cannot happen inside the margins; stay safe)
    const d3 p = vadd(s, vscale(span, lf));
    *start = p;
    *offset = ddiv(vlen(vsub(s, p)), vlen(dir));               // originalStart.Distance(start) / dir.Length
    return true;
}

// ---------------------------------------------------------------------------------------------
// FP32 BVH traversal (candidate search only)
// ---------------------------------------------------------------------------------------------
struct TravRay {
    float ox, oy, oz;      // origin near the root box
    float ix, iy, iz;      // 1/dir
    float nox, noy, noz;   // -o * (1/dir)
    double t_off;          // exact-parameter value at the traversal origin
};

// Conservative entry into [mn - pad, mx + pad] along s + t*dir, t >= 0 (plain FP64, not part of
// the reference arithmetic).  Returns false if the ray cannot touch the box.
__device__ __forceinline__ bool entry_clip(const double* mn, const double* mx, double pad, d3 s, d3 dir, double* t_enter)
{
    double t0 = 0.0, t1 = 1.7976931348623157e308;
    const double sv[3] = {s.x, s.y, s.z}, dv[3] = {dir.x, dir.y, dir.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double lo = mn[k] - pad, hi = mx[k] + pad;
        if (dv[k] == 0.0) {
            if (sv[k] < lo || sv[k] > hi) return false;
        } else {
            const double inv = 1.0 / dv[k];
            double a = (lo - sv[k]) * inv, b = (hi - sv[k]) * inv;
            if (a > b) { const double tmp = a; a = b; b = tmp; }
            t0 = fmax(t0, a); t1 = fmin(t1, b);
        }
    }
    if (t0 > t1 * (1.0 + 1e-12) + 1e-12) return false;
    *t_enter = t0 > 0.0 ? t0 * (1.0 - 1e-9) : 0.0;
    return true;
}

__device__ __forceinline__ TravRay make_trav(d3 s, d3 dir, double t_enter)
{
    TravRay r;
    r.ox = (float)(s.x + dir.x * t_enter); r.oy = (float)(s.y + dir.y * t_enter); r.oz = (float)(s.z + dir.z * t_enter);
    const float dx = (float)dir.x, dy = (float)dir.y, dz = (float)dir.z;
    r.ix = 1.0f / dx; r.iy = 1.0f / dy; r.iz = 1.0f / dz;      // +-inf for zero components
    r.nox = -r.ox * r.ix; r.noy = -r.oy * r.iy; r.noz = -r.oz * r.iz;
    // 0 * inf = NaN: fminf/fmaxf drop NaNs, so such an axis never constrains the slab
    r.t_off = t_enter;
    return r;
}

__device__ __forceinline__ bool slab(const TravRay& r, float lox, float loy, float loz, float hix, float hiy, float hiz,
                                     float tcull, float* t_entry)
{
    const float ax = __fmaf_rn(lox, r.ix, r.nox), bx = __fmaf_rn(hix, r.ix, r.nox);
    const float ay = __fmaf_rn(loy, r.iy, r.noy), by = __fmaf_rn(hiy, r.iy, r.noy);
    const float az = __fmaf_rn(loz, r.iz, r.noz), bz = __fmaf_rn(hiz, r.iz, r.noz);
    float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tcull));
    *t_entry = tmin;
    // no slack here: the boxes are padded in space by 64u x scale (sr_bvh.cpp), > 4x what the roundings
    // of o, 1/d, the planes and the six FMAs can move a crossing (DESIGN.md "FP32 candidate search")
    return tmin <= tmax;
}

// limit in exact-parameter units -> conservative FP32 cull distance measured from the traversal origin
__device__ __forceinline__ float cull_from(double limit, double t_off)
{
    if (limit >= 1e300) return CUDART_INF_F;
    const float v = __double2float_ru(limit - t_off);
    return fmaxf(v, 0.0f) * 1.00002f + 1e-6f;
}

// ---------------------------------------------------------------------------------------------
// BVH2 traversal skeleton, shared by every walk (exact, filtered any-hit / closest-hit, cone).
//   BOX(lox,loy,loz,hix,hiy,hiz,&t) -> bool : conservative FP32 slab test of one child
//   LEAF(first, count) -> bool              : true = stop the walk (any-hit found / give up)
// One node or one leaf per iteration ("if-if"): measured faster here than the "while-while" form
// (46 vs 58 ms on the 1M-triangle config) -- the rays of a warp are coherent and leaves hold 1-2
// triangles, so waiting for every lane to reach a leaf costs more than it saves.
// The thread's single stack is passed in (Counters::stack); returns the number of nodes visited.
// ---------------------------------------------------------------------------------------------
#ifndef SR_WALK_MODE
#define SR_WALK_MODE 0     // 0: one node or one leaf per iteration ("if-if"); 1: "while-while" (the stage kernels)
#endif
// GLOBAL = false: `nodes` is a block's shared-memory copy of a small tree (plain loads, nothing to prefetch)
template <bool GLOBAL = true, class BOX, class LEAF>
__device__ __forceinline__ unsigned int walk_bvh(const BvhNode* __restrict__ nodes, int n_prims, int* __restrict__ stack, BOX box,
                                                 LEAF leaf)
{
    // a handful of primitives (config 2's 12-triangle room): testing them all costs less than the ~10 box
    // pairs of their tree.  Leaf order == storage order, so this is the whole array.
    if (n_prims <= kTinyMesh) { leaf(0, n_prims); return 0; }
    int sp = 0;
    int cur = 0;
    unsigned int nv = 0;
#if SR_WALK_MODE == 1
    // "while-while": every lane descends until it holds a leaf (or has nothing left), then the lanes of the warp
    // test their leaves TOGETHER.  In the if-if form the leaf code runs whenever some lane happens to reach a leaf:
    // with 4 of 32 lanes on average (profiles/r02a_*), a quarter of all issue slots at 12 % lane utilisation.
    constexpr int kDone = (int)0x80000000;
    for (;;) {
        while (cur >= 0) {
            const float4* p = reinterpret_cast<const float4*>(nodes + cur);
            const float4 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
            const int2 d = __ldg(reinterpret_cast<const int2*>(p + 3));
            nv++;
            float t0, t1;
            const bool h0 = box(a.x, a.y, a.z, a.w, b.x, b.y, &t0);
            const bool h1 = box(b.z, b.w, cc.x, cc.y, cc.z, cc.w, &t1);
            if (h0 && h1) {
                const bool first0 = t0 <= t1;
                const int far = first0 ? d.y : d.x;
                stack[sp++] = far;
#if SR_PREFETCH >= 1
                if (far >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + far));
#endif
                cur = first0 ? d.x : d.y;
            } else if (h0) {
                cur = d.x;
            } else if (h1) {
                cur = d.y;
            } else {
                cur = sp > 0 ? stack[--sp] : kDone;
            }
        }
        if (cur == kDone) break;
        const int code = -1 - cur;
        if (leaf(code >> 4, code & 15)) break;
        cur = sp > 0 ? stack[--sp] : kDone;
        if (cur == kDone) break;
    }
#else
    for (;;) {
        if (cur >= 0) {
            const float4* p = reinterpret_cast<const float4*>(nodes + cur);
            const float4 a = GLOBAL ? __ldg(p) : p[0], b = GLOBAL ? __ldg(p + 1) : p[1], cc = GLOBAL ? __ldg(p + 2) : p[2];
            const int2 d = GLOBAL ? __ldg(reinterpret_cast<const int2*>(p + 3)) : *reinterpret_cast<const int2*>(p + 3);
            nv++;
            float t0, t1;
            const bool h0 = box(a.x, a.y, a.z, a.w, b.x, b.y, &t0);
            const bool h1 = box(b.z, b.w, cc.x, cc.y, cc.z, cc.w, &t1);
            if (h0 && h1) {
                const bool first0 = t0 <= t1;
                const int far = first0 ? d.y : d.x;
                stack[sp++] = far;
#if SR_PREFETCH >= 1
                // the far child is needed after the whole near subtree: start its fetch now
                if (GLOBAL && far >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + far));
#endif
                cur = first0 ? d.x : d.y;
                continue;
            }
            if (h0) { cur = d.x; continue; }
            if (h1) { cur = d.y; continue; }
        } else {
            const int code = -1 - cur;
            if (leaf(code >> 4, code & 15)) break;
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
#endif
    return nv;
}

struct BestPrim { double rf; int k; int index; };

// Generic BVH walk.  PRIM = 0 triangles, 1 spheres.  ANY: stop at the first primitive whose
// exact rayFrac is <= limit (shadow rays); otherwise find the minimum rayFrac, ties to the
// lowest list index (GeometryCollection.cs:53, SpatialSubdivision.cs:644).
template <int PRIM, bool ANY>
__device__ __forceinline__ bool walk(const BvhNode* __restrict__ nodes, const void* __restrict__ prims, const TravRay& tr,
                                     d3 s, d3 dir, double limit, double any_offset, BestPrim* best, XCounters* c)
{
    unsigned int np = 0;
    bool found = false;
    float tcull = cull_from(ANY ? limit : best->rf, tr.t_off);
    const unsigned int nv = walk_bvh(
        nodes, kTinyMesh + 1, c->stack,
        [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float* t) {
            return slab(tr, lox, loy, loz, hix, hiy, hiz, tcull, t);
        },
        [&](int first, int count) {
            for (int i = 0; i < count; i++) {
                const int k = first + i;
                double rf;
                np++;
                if (PRIM == 0) {
                    const TriRec* t = reinterpret_cast<const TriRec*>(prims) + k;
                    if (!tri_intersect(t, s, dir, ANY ? limit : best->rf, &rf)) continue;
                    if (ANY) {
                        if (dadd(rf, any_offset) <= 1.0) { found = true; return true; }
                        continue;
                    }
                    const int index = __ldg(reinterpret_cast<const int*>(t) + 31);
                    if (rf < best->rf || (rf == best->rf && index < best->index)) {
                        best->rf = rf; best->k = k; best->index = index;
                        tcull = cull_from(rf, tr.t_off);
                    }
                } else {
                    const SphereRec* q = reinterpret_cast<const SphereRec*>(prims) + k;
                    if (!sphere_intersect(q, s, dir, &rf)) continue;
                    if (ANY) {
                        if (rf <= limit) { found = true; return true; }
                        continue;
                    }
                    const int index = __ldg(reinterpret_cast<const int*>(q) + 11);
                    if (rf < best->rf || (rf == best->rf && index < best->index)) {
                        best->rf = rf; best->k = k; best->index = index;
                        tcull = cull_from(rf, tr.t_off);
                    }
                }
            }
            return false;
        });
    c->node_visits += nv; c->prim_tests += np;
    if (PRIM == 1) c->sphere_tests += np;
    return found;
}

// Linear scan (SOFTRAY_ACCEL_BRUTE): GeometryCollection.IntersectRay (GeometryCollection.cs:44-69).
template <int PRIM, bool ANY>
__device__ __forceinline__ bool scan(const void* __restrict__ prims, int n, d3 s, d3 dir, double limit, double any_offset,
                                     BestPrim* best, XCounters* c)
{
    unsigned int np = 0;
    bool found = false;
    for (int k = 0; k < n; k++) {
        double rf;
        np++;
        if (PRIM == 0) {
            const TriRec* t = reinterpret_cast<const TriRec*>(prims) + k;
            if (!tri_intersect(t, s, dir, ANY ? limit : best->rf, &rf)) continue;
            if (ANY) { if (dadd(rf, any_offset) <= 1.0) { found = true; break; } continue; }
            if (rf < best->rf) { best->rf = rf; best->k = k; best->index = k; }
        } else {
            const SphereRec* q = reinterpret_cast<const SphereRec*>(prims) + k;
            if (!sphere_intersect(q, s, dir, &rf)) continue;
            if (ANY) { if (rf <= limit) { found = true; break; } continue; }
            if (rf < best->rf) { best->rf = rf; best->k = k; best->index = k; }
        }
    }
    c->prim_tests += np;
    if (PRIM == 1) c->sphere_tests += np;
    return found;
}

// ---------------------------------------------------------------------------------------------
// FP32 filtered predicate for shadow rays (DESIGN.md "Filtered predicates")
//
// A shadow ray only needs a boolean: does ANY triangle report rayFrac (+ clip offset) <= 1.0
// (ShadowMethod.cs:171).  The filter evaluates the same plane / barycentric formulation in FP32
// (FMA allowed) together with a running bound on how far each FP32 quantity can be from the value
// the FP64 reference arithmetic produces, and answers only when every comparison clears its bound:
//     0 = surely no triangle occludes, 1 = surely one does, 2 = cannot tell (-> exact FP64 path).
// The ray is parametrised backwards from its far end:  P(tau) = anchor + g * tau,  g = -dir,
// anchor = start + dir (the shadow receiver `end`, which lies on the geometry, so every FP32 operand
// stays of the order of the scene size whatever the distance of the light).  rayFrac_total = 1 - tau:
//     rayFrac + offset <= 1.0          <=>  tau >= 0
//     hit not before the clipped start <=>  tau <= min(1, tau_out)   (tau_out: where P leaves the root
//                                            box; SpatialSubdivision.cs:389-401 moves the start there)
// u = 2^-24.  First-order bounds (each constant below is >= 1.5x the derived one):
//     |g.n  - exact| <= 5u |g|_1            (|n_k| <= 1)                     used: 8u
//     |num  - exact| <= 4u |d| + 5u |o|_1   (num = d - o.n)                  used: 8u
//     |tau  - exact| <= (E_num + |tau| E_gn) / gn * 1.07 + 3u |tau|          used: 1.1, 4u
//     |w_k  - exact| <= |g|_inf E_tau + 3u (|o|_inf + |g|_inf |tau|) + 2u V  (w = P - v1, V = max |coord|)
//     |s    - exact| <= a1 (E_w + 4u W)  <= a1 (|g|_inf E_tau + 7u (|o|_inf + |g|_inf |tau| + V))   used: 12u
// ---------------------------------------------------------------------------------------------
constexpr float kU = 5.9604644775390625e-8f;   // 2^-24

struct FRay {
    float ox, oy, oz;         // anchor
    float gx, gy, gz;         // direction of increasing tau
    float ix, iy, iz;         // 1 / g
    float nox, noy, noz;      // -o / g
    float g1, ginf, o1, oinf; // |g|_1, |g|_inf, |o|_1, |o|_inf
    float tmin_hi;            // upper bound of the smallest admissible tau (0 for shadow rays)
    float tmax_lo, tmax_hi;   // bracket of the largest admissible tau
    float tcull;              // traversal cull distance (>= tmax_hi, with slack)
};

// 0: no triangle of the mesh can be hit, 1: traverse, 2: cannot tell
__device__ __forceinline__ int fray_setup(const DevMesh& m, int subdivision, d3 anchor, d3 dir, FRay* r)
{
    r->ox = __double2float_rn(anchor.x); r->oy = __double2float_rn(anchor.y); r->oz = __double2float_rn(anchor.z);
    r->gx = -__double2float_rn(dir.x); r->gy = -__double2float_rn(dir.y); r->gz = -__double2float_rn(dir.z);
    const float agx = fabsf(r->gx), agy = fabsf(r->gy), agz = fabsf(r->gz);
    const float aox = fabsf(r->ox), aoy = fabsf(r->oy), aoz = fabsf(r->oz);
    r->g1 = agx + agy + agz; r->ginf = fmaxf(agx, fmaxf(agy, agz));
    r->o1 = aox + aoy + aoz; r->oinf = fmaxf(aox, fmaxf(aoy, aoz));
    // an exactly axis-parallel or degenerate direction or a non-finite operand: exact path
    if (!(fminf(agx, fminf(agy, agz)) > 1e-30f) || !(r->ginf < 1e30f) || !(r->oinf < 1e30f)) return 2;
    r->ix = __fdiv_rn(1.0f, r->gx); r->iy = __fdiv_rn(1.0f, r->gy); r->iz = __fdiv_rn(1.0f, r->gz);
    r->nox = -r->ox * r->ix; r->noy = -r->oy * r->iy; r->noz = -r->oz * r->iz;
    // the root box along tau, with a per-axis bound on every crossing:
    //   c_k = (plane_k - o_k) / g_k;  |c_k - exact| <= |1/g_k| u (|plane_k| + |o_k|) + 3u |c_k|        used: 2u, 4u
    // (+ 4e-10: the 1e-10 tolerance of AxisAlignedBox.ContainsPoint, AxisAlignedBox.cs:9,143-149)
    const float fxa = (m.fmin[0] - r->ox) * r->ix, fxb = (m.fmax[0] - r->ox) * r->ix;
    const float fya = (m.fmin[1] - r->oy) * r->iy, fyb = (m.fmax[1] - r->oy) * r->iy;
    const float fza = (m.fmin[2] - r->oz) * r->iz, fzb = (m.fmax[2] - r->oz) * r->iz;
    const float farx = fmaxf(fxa, fxb), fary = fmaxf(fya, fyb), farz = fmaxf(fza, fzb);
    const float nearx = fminf(fxa, fxb), neary = fminf(fya, fyb), nearz = fminf(fza, fzb);
    const float px = fabsf(r->ix) * ((2.0f * kU) * (m.scale + aox) + 4e-10f);
    const float py = fabsf(r->iy) * ((2.0f * kU) * (m.scale + aoy) + 4e-10f);
    const float pz = fabsf(r->iz) * ((2.0f * kU) * (m.scale + aoz) + 4e-10f);
    const float out_lo = fminf(farx - (px + (4.0f * kU) * fabsf(farx)),
                               fminf(fary - (py + (4.0f * kU) * fabsf(fary)), farz - (pz + (4.0f * kU) * fabsf(farz))));
    const float out_hi = fminf(farx + (px + (4.0f * kU) * fabsf(farx)),
                               fminf(fary + (py + (4.0f * kU) * fabsf(fary)), farz + (pz + (4.0f * kU) * fabsf(farz))));
    const float in_lo = fmaxf(nearx - (px + (4.0f * kU) * fabsf(nearx)),
                              fmaxf(neary - (py + (4.0f * kU) * fabsf(neary)), nearz - (pz + (4.0f * kU) * fabsf(nearz))));
    if (!(out_hi >= 0.0f)) return out_hi < 0.0f ? 0 : 2;      // the box lies wholly behind the anchor (NaN: cannot tell)
    r->tmax_hi = fminf(1.0f, out_hi);
    if (in_lo > r->tmax_hi) return 0;                         // the box is missed, or lies beyond the ray's start
    // an anchor far from the mesh: the BVH pad (sr_bvh.cpp) assumes an origin within ~2x its scale
    if (!(r->oinf <= 2.0f * m.scale) || !(in_lo == in_lo)) return 2;
    r->tmax_lo = subdivision ? fminf(1.0f, out_lo) : 1.0f;
    r->tmin_hi = 0.0f;
    r->tcull = r->tmax_hi * 1.00002f + 1e-6f;
    return 1;
}

// Forward (camera / reflection) ray  start + dir * t,  t >= 0.  The FP32 anchor is the point at
// t0 <= (exact entry into the root box), computed in FP64 and then rounded, so that every FP32 operand
// is of the order of the mesh whatever the distance of the camera; tau = t - t0.
// 0: the ray surely misses the root box (no hit), 1: traverse, 2: cannot tell.
__device__ __forceinline__ int fray_setup_box(const float* __restrict__ bmin, const float* __restrict__ bmax, float scale,
                                              int subdivision, d3 s, d3 dir, FRay* r, double* t0_out)
{
    struct { const float* fmin; const float* fmax; float scale; } m = {bmin, bmax, scale};
    const float sx = __double2float_rn(s.x), sy = __double2float_rn(s.y), sz = __double2float_rn(s.z);
    r->gx = __double2float_rn(dir.x); r->gy = __double2float_rn(dir.y); r->gz = __double2float_rn(dir.z);
    const float agx = fabsf(r->gx), agy = fabsf(r->gy), agz = fabsf(r->gz);
    const float asx = fabsf(sx), asy = fabsf(sy), asz = fabsf(sz);
    r->g1 = agx + agy + agz; r->ginf = fmaxf(agx, fmaxf(agy, agz));
    if (!(fminf(agx, fminf(agy, agz)) > 1e-30f) || !(r->ginf < 1e30f) || !(fmaxf(asx, fmaxf(asy, asz)) < 1e30f)) return 2;
    r->ix = __fdiv_rn(1.0f, r->gx); r->iy = __fdiv_rn(1.0f, r->gy); r->iz = __fdiv_rn(1.0f, r->gz);
    // crossings of the root box from the (possibly far) start, per-axis bounds as in fray_setup
    const float fxa = (m.fmin[0] - sx) * r->ix, fxb = (m.fmax[0] - sx) * r->ix;
    const float fya = (m.fmin[1] - sy) * r->iy, fyb = (m.fmax[1] - sy) * r->iy;
    const float fza = (m.fmin[2] - sz) * r->iz, fzb = (m.fmax[2] - sz) * r->iz;
    const float farx = fmaxf(fxa, fxb), fary = fmaxf(fya, fyb), farz = fmaxf(fza, fzb);
    const float nearx = fminf(fxa, fxb), neary = fminf(fya, fyb), nearz = fminf(fza, fzb);
    const float px = fabsf(r->ix) * ((2.0f * kU) * (m.scale + asx) + 4e-10f);
    const float py = fabsf(r->iy) * ((2.0f * kU) * (m.scale + asy) + 4e-10f);
    const float pz = fabsf(r->iz) * ((2.0f * kU) * (m.scale + asz) + 4e-10f);
    const float out_hi = fminf(farx + (px + (4.0f * kU) * fabsf(farx)),
                               fminf(fary + (py + (4.0f * kU) * fabsf(fary)), farz + (pz + (4.0f * kU) * fabsf(farz))));
    const float in_lo = fmaxf(nearx - (px + (4.0f * kU) * fabsf(nearx)),
                              fmaxf(neary - (py + (4.0f * kU) * fabsf(neary)), nearz - (pz + (4.0f * kU) * fabsf(nearz))));
    const float in_hi = fmaxf(nearx + (px + (4.0f * kU) * fabsf(nearx)),
                              fmaxf(neary + (py + (4.0f * kU) * fabsf(neary)), nearz + (pz + (4.0f * kU) * fabsf(nearz))));
    if (!(out_hi >= 0.0f)) return out_hi < 0.0f ? 0 : 2;      // the box lies behind the start
    if (in_lo > out_hi) return 0;                             // the box is missed
    if (!(in_lo == in_lo) || !(in_hi == in_hi) || !(out_hi < 9000.0f)) return 2;   // (the reference's ray ends at t = 10000)
    const float t0 = fmaxf(in_lo, 0.0f);
    *t0_out = (double)t0;
    const double ax = s.x + dir.x * (double)t0, ay = s.y + dir.y * (double)t0, az = s.z + dir.z * (double)t0;
    r->ox = __double2float_rn(ax); r->oy = __double2float_rn(ay); r->oz = __double2float_rn(az);
    const float aox = fabsf(r->ox), aoy = fabsf(r->oy), aoz = fabsf(r->oz);
    r->o1 = aox + aoy + aoz; r->oinf = fmaxf(aox, fmaxf(aoy, aoz));
    if (!(r->oinf <= 2.0f * m.scale)) return 2;
    r->nox = -r->ox * r->ix; r->noy = -r->oy * r->iy; r->noz = -r->oz * r->iz;
    // the clipped start of SpatialSubdivision.IntersectRay lies at tau in [0, gap]
    r->tmin_hi = subdivision ? (fmaxf(in_hi, 0.0f) - t0) * (1.0f + 4.0f * kU) + (4.0f * kU) * fabsf(in_hi) : 0.0f;
    r->tmax_hi = (out_hi - t0) * (1.0f + 4.0f * kU) + (4.0f * kU) * fabsf(out_hi);
    r->tmax_lo = 1e30f;
    r->tcull = r->tmax_hi * 1.00002f + 1e-6f;
    return 1;
}

__device__ __forceinline__ int fray_setup_fwd(const DevMesh& m, int subdivision, d3 s, d3 dir, FRay* r, double* t0_out)
{
    return fray_setup_box(m.fmin, m.fmax, m.scale, subdivision, s, dir, r, t0_out);
}

__device__ __forceinline__ bool fslab(const FRay& r, float lox, float loy, float loz, float hix, float hiy, float hiz,
                                      float* t_entry)
{
    const float ax = __fmaf_rn(lox, r.ix, r.nox), bx = __fmaf_rn(hix, r.ix, r.nox);
    const float ay = __fmaf_rn(loy, r.iy, r.noy), by = __fmaf_rn(hiy, r.iy, r.noy);
    const float az = __fmaf_rn(loz, r.iz, r.noz), bz = __fmaf_rn(hiz, r.iz, r.noz);
    const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), r.tcull));
    *t_entry = tmin;
    return tmin <= tmax;                         // like slab(): the boxes are padded in space
}

// 0 surely missed (or outside the admissible tau range), 1 surely hit inside it, 2 cannot tell.
// FWD = false: tau runs against the ray (shadow rays, g = -dir); FWD = true: along it (g = dir), which
// flips the sign of the plane.  thi: candidates surely beyond it are of no interest (-> 0).
// On 1 and 2, [*tau_o - *etau_o, *tau_o + *etau_o] contains the exact tau (0 +- 0 when not even that is known).
#ifndef SR_FILTER_PRELOAD
#define SR_FILTER_PRELOAD 1     // config5 search 8.01 -> 7.79 ms, config3 shadow 4.83 -> 4.62, config4 unchanged
#endif
template <bool FWD>
__device__ __forceinline__ int tri_filter(const TriFilt* __restrict__ t, const FRay& r, float V, float thi, float* tau_o,
                                          float* etau_o)
{
    const float4* p = reinterpret_cast<const float4*>(t);
    const float4 q0 = __ldg(p);                                          // n, d
#if SR_FILTER_PRELOAD
    // all four 16-byte words of the record at once: the later ones are needed only past the early outs, but a load
    // issued there is a load waited for (the leaf test runs with few lanes and nothing else to hide the latency)
    const float4 q3 = __ldg(p + 3), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
#endif
    float gn = __fmaf_rn(r.gz, q0.z, __fmaf_rn(r.gy, q0.y, r.gx * q0.x));
    if (FWD) gn = -gn;
    const float e_gn = (8.0f * kU) * r.g1;
    *tau_o = 0.0f; *etau_o = 0.0f;
    if (!(gn > e_gn)) return gn < -e_gn ? 0 : 2;                         // dir.n >= 0: one-sided (Plane.cs:75)
    float num = __fmaf_rn(-r.oz, q0.z, __fmaf_rn(-r.oy, q0.y, __fmaf_rn(-r.ox, q0.x, q0.w)));
    if (FWD) num = -num;
    const float e_num = (8.0f * kU) * (fabsf(q0.w) + r.o1);
    if (num < -e_num) return 0;                                          // tau < 0
    if (!(gn > 16.0f * e_gn)) return 2;                                  // grazing: tau not trustworthy
    const float rg = __fdividef(1.0f, gn);
    const float tau = num * rg;
    const float e_tau = (e_num + fabsf(tau) * e_gn) * rg * 1.1f + (4.0f * kU) * fabsf(tau);
    if (tau - e_tau > thi) return 0;
#if !SR_FILTER_PRELOAD
    const float4 q3 = __ldg(p + 3);                                      // v1, -
#endif
    const float wx = __fmaf_rn(r.gx, tau, r.ox) - q3.x, wy = __fmaf_rn(r.gy, tau, r.oy) - q3.y,
                wz = __fmaf_rn(r.gz, tau, r.oz) - q3.z;
#if !SR_FILTER_PRELOAD
    const float4 q1 = __ldg(p + 1);                                      // a, a1
#endif
    if (q1.w < 0.0f) return 0;                                           // zero-area triangle: never hit
    const float k = r.ginf * e_tau + (12.0f * kU) * (r.oinf + r.ginf * fabsf(tau) + V);
    const float sN = __fmaf_rn(wz, q1.z, __fmaf_rn(wy, q1.y, wx * q1.x));
    const float e_s = q1.w * k;
    if (sN < -e_s || sN > 1.0f + e_s) return 0;
#if !SR_FILTER_PRELOAD
    const float4 q2 = __ldg(p + 2);                                      // b, b1
#endif
    const float uu = __fmaf_rn(wz, q2.z, __fmaf_rn(wy, q2.y, wx * q2.x));
    const float e_u = q2.w * k;
    if (uu < -e_u) return 0;
    const float sum = sN + uu, e_sum = e_s + e_u + 4.0f * kU;
    if (sum > 1.0f + e_sum) return 0;
    *tau_o = tau; *etau_o = e_tau;
    if (num > e_num && tau - e_tau > r.tmin_hi && tau + e_tau < r.tmax_lo && sN > e_s && uu > e_u && sum < 1.0f - e_sum)
        return 1;
    if (!(e_tau == e_tau) || !(tau == tau)) { *tau_o = 0.0f; *etau_o = 0.0f; }
    return 2;                                                            // also every NaN / inf case
}

// BVH walk with the filter at the leaves.  0 surely clear, 1 surely occluded, 2 cannot tell: then
// unsure[0..*n_unsure) lists the triangles the exact arithmetic has to look at (every other triangle is
// surely missed); *n_unsure > kMaxCand means too many to list.
__device__ __forceinline__ int walk_filter_any(const BvhNode* __restrict__ nodes, const TriFilt* __restrict__ filt, int n_tris,
                                               const FRay& r, float V, int* unsure, int* n_unsure, Counters* c)
{
    bool hit = false;
    int nu = 0;
    int u0 = -1, u1 = -1, u2 = -1, u3 = -1;
    unsigned int nf = 0;
    c->node_visits += walk_bvh(
        nodes, n_tris, c->stack,
        [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float* t) {
            return fslab(r, lox, loy, loz, hix, hiy, hiz, t);
        },
        [&](int first, int count) {
            for (int i = 0; i < count; i++) {
                nf++;
                float tau, etau;
                const int res = tri_filter<false>(filt + first + i, r, V, r.tmax_hi, &tau, &etau);
                if (res == 1) { hit = true; return true; }
                if (res == 2) {
                    if (nu == 0) u0 = first + i; else if (nu == 1) u1 = first + i; else if (nu == 2) u2 = first + i;
                    else if (nu == 3) u3 = first + i;
                    nu++;
                }
            }
            return false;
        });
    c->filter_tests += nf;
    if (hit) return 1;
    if (nu == 0) return 0;
    unsure[0] = u0; unsure[1] = u1; unsure[2] = u2; unsure[3] = u3;
    *n_unsure = nu;
    return 2;
}

// Nearest hit with the filter at the leaves.  Tracks the sure hit with the smallest upper bound
// (best) and the smallest lower bound of every OTHER candidate, sure or not (other_lo): the winner of
// the exact arithmetic is known iff best_hi < other_lo.
constexpr int kMaxCand = 4;
// best_k < 0: no sure hit.  cand[0..n_cand): every triangle that is not surely missed (incl. the best);
// n_cand > kMaxCand: too many to list.  The exact winner is among the candidates whose lower bound does
// not exceed best_hi (everything else is surely missed or surely behind the best sure hit).
struct FClosest { float best_hi; int best_k; int n_cand; int cand[kMaxCand]; float cand_lo[kMaxCand]; };

__device__ __forceinline__ void walk_filter_closest(const BvhNode* __restrict__ nodes, const TriFilt* __restrict__ filt, int n_tris,
                                                    FRay& r, float V, FClosest* out, XCounters* c, float limit_tau = 1e30f)
{
    // limit_tau: hits beyond it are of no use to the caller (a composite frame already has a nearer hit in
    // another instance): they are culled like hits behind a sure hit
    float best_hi = limit_tau;
    r.tcull = fminf(r.tcull, limit_tau * 1.00002f + 1e-6f);
    int best_k = -1;
    int n_cand = 0;
    int c0 = -1, c1 = -1, c2 = -1, c3 = -1;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    unsigned int nf = 0;
    c->node_visits += walk_bvh(
        nodes, n_tris, c->stack,
        [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float* t) {
            return fslab(r, lox, loy, loz, hix, hiy, hiz, t);
        },
        [&](int first, int count) {
            for (int i = 0; i < count; i++) {
                nf++;
                float tau, etau;
                const int res = tri_filter<true>(filt + first + i, r, V, best_hi, &tau, &etau);
                if (res == 0) continue;
                const float lo = tau - etau, hi = tau + etau;
                if (n_cand == 0) { c0 = first + i; l0 = lo; }
                else if (n_cand == 1) { c1 = first + i; l1 = lo; }
                else if (n_cand == 2) { c2 = first + i; l2 = lo; }
                else if (n_cand == 3) { c3 = first + i; l3 = lo; }
                n_cand++;
                if (res == 1 && hi < best_hi) {
                    best_hi = hi; best_k = first + i;
                    r.tcull = hi * 1.00002f + 1e-6f;
                }
            }
            return false;
        });
    c->filter_tests += nf;
    out->best_hi = best_hi; out->best_k = best_k; out->n_cand = n_cand;
    out->cand[0] = c0; out->cand[1] = c1; out->cand[2] = c2; out->cand[3] = c3;
    out->cand_lo[0] = l0; out->cand_lo[1] = l1; out->cand_lo[2] = l2; out->cand_lo[3] = l3;
}

// ---------------------------------------------------------------------------------------------
// FP32 filter for spheres (Sphere.IntersectRay, Sphere.cs:152-219).  The ray is P(tau) = o + g tau with
// g = dir rounded (not normalised); L = |g|.  In DISTANCE units (the unit of a sphere's rayFrac, SURVEY
// App. A #3):  proj = (o - c).g / L,  term = proj^2 - |o - c|^2 + r^2,  d = -proj -+ sqrt(term),
// rayFrac = T0 + d  with T0 = (distance from the ray start to the anchor o).
// Bounds (u = 2^-24):  |o_k - c_k| error e_o = 3u (|o|_inf + V);  E_p = (e_o |g|_1 + 5u |o-c|_1 |g|_inf) / L + 6u |proj|;
// E_t = 2|proj| E_p + E_p^2 + 2 |o-c|_1 e_o + 3 e_o^2 + 8u (proj^2 + |o-c|^2 + r^2);
// E_root = E_t / (2 sqrt(term - E_t)) + 2u root;  E_d = E_p + E_root + 2u (|proj| + root) + 4u T0.
// 0: surely missed; 1: surely hit from outside (rayFrac = T0 + d1 within [*lo, *hi]); 2: cannot tell.
// ---------------------------------------------------------------------------------------------
struct SRay { float inv_len, T0; };

template <bool GLOBAL = true>
__device__ __forceinline__ int sphere_filter(const float4* __restrict__ rec, const FRay& r, const SRay& sr, float V, float* lo,
                                             float* hi)
{
    const float4 q = GLOBAL ? __ldg(rec) : *rec;
    const float ox = r.ox - q.x, oy = r.oy - q.y, oz = r.oz - q.z;
    const float e_o = (3.0f * kU) * (r.oinf + V);
    const float o1 = fabsf(ox) + fabsf(oy) + fabsf(oz);
    const float proj = __fmaf_rn(oz, r.gz, __fmaf_rn(oy, r.gy, ox * r.gx)) * sr.inv_len;
    const float e_p = (e_o * r.g1 + (5.0f * kU) * o1 * r.ginf) * sr.inv_len * 1.01f + (6.0f * kU) * fabsf(proj);
    const float oo = __fmaf_rn(oz, oz, __fmaf_rn(oy, oy, ox * ox));
    const float rr = q.w * q.w;
    const float term = __fmaf_rn(proj, proj, rr - oo);
    const float e_t = 2.0f * fabsf(proj) * e_p + e_p * e_p + 2.0f * o1 * e_o + 3.0f * e_o * e_o +
                      (8.0f * kU) * (proj * proj + oo + rr);
    *lo = 0.0f; *hi = 0.0f;
    if (term + e_t < 1e-10f) return 0;                                   // `term < EPSILON` (Sphere.cs:176)
    const float term_lo = term - e_t;
    if (!(term_lo > 1.0001e-10f)) return 2;
    const float root = sqrtf(term);
    const float e_root = __fdividef(e_t, 2.0f * sqrtf(term_lo)) * 1.01f + (2.0f * kU) * root;
    const float e_d = e_p + e_root + (2.0f * kU) * (fabsf(proj) + root) + (4.0f * kU) * sr.T0;
    const float d1 = -proj - root, d2 = -proj + root;
    if (sr.T0 + d2 + e_d < 0.0f) return 0;                               // the whole sphere lies behind the start
    const float rf = sr.T0 + d1;
    if (rf - e_d > 0.0f) { *lo = rf - e_d; *hi = rf + e_d; return 1; }  // entering from outside: rayFrac = f1
    return 2;                                                            // start inside / on the sphere, NaN, ...
}

// The sphere part of rootGeometry for one ray.  ANY: is there a sphere with rayFrac <= 1.0 (shadow rays)?
// returns 0 no, 1 yes, 2 cannot tell.  !ANY: nearest sphere -> candidates, as walk_filter_closest.
// In both cases list[0..*n_list) names the spheres the exact arithmetic has to look at (> kMaxCand: all).
// a block's shared-memory copy of a small sphere set's tree and filter records (sr_render.cu stages it)
struct SphereStage { const BvhNode* nodes; const float4* filt; };

template <bool ANY, bool GLOBAL>
__device__ __forceinline__ int spheres_filter_impl(const DevScene& sc, const BvhNode* __restrict__ nodes, const float4* __restrict__ filt, d3 s,
                                                   d3 dir, int* list, int* n_list, unsigned int* nv_out, unsigned int* nf_out, int* stack)
{
    FRay r; double t0;
    *n_list = 0;
    int st = fray_setup_box(sc.sph_fmin, sc.sph_fmax, sc.sph_scale, 0, s, dir, &r, &t0);
    if (st != 1) return st == 0 ? 0 : 2;
    const float len = sqrtf(r.gx * r.gx + r.gy * r.gy + r.gz * r.gz);
    SRay sr; sr.inv_len = __fdiv_rn(1.0f, len); sr.T0 = (float)t0 * len;
    if (ANY) {
        // a sphere's rayFrac is at least the distance to its (padded) box: beyond 1.0 nothing can occlude
        if (sr.T0 * (1.0f - 8.0f * kU) > 1.0f) return 0;
        r.tcull = fminf(r.tcull, (1.0f - sr.T0) * sr.inv_len * 1.0001f + 1e-6f);
    }
    float best_hi = 1e30f;
    bool occluded = false;
    int n = 0, c0 = -1, c1 = -1, c2 = -1, c3 = -1;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    unsigned int nf = 0;
    *nv_out += walk_bvh<GLOBAL>(
        nodes, sc.n_spheres, stack,
        [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float* t) {
            return fslab(r, lox, loy, loz, hix, hiy, hiz, t);
        },
        [&](int first, int count) {
            for (int i = 0; i < count; i++) {
                nf++;
                float lo, hi;
                const int res = sphere_filter<GLOBAL>(filt + first + i, r, sr, sc.sph_scale, &lo, &hi);
                if (res == 0) continue;
                if (ANY) {
                    if (res == 1 && hi <= 1.0f) { occluded = true; return true; }
                    if (res == 1 && lo > 1.0f) continue;                  // hit, but further than 1.0 from the start
                } else if (res == 1 && lo > best_hi) {
                    continue;                                             // surely behind a sure hit
                }
                if (n == 0) { c0 = first + i; l0 = lo; } else if (n == 1) { c1 = first + i; l1 = lo; }
                else if (n == 2) { c2 = first + i; l2 = lo; } else if (n == 3) { c3 = first + i; l3 = lo; }
                n++;
                if (!ANY && res == 1 && hi < best_hi) {
                    best_hi = hi;
                    r.tcull = fminf(r.tcull, (hi - sr.T0) * sr.inv_len * 1.0001f + 1e-6f);
                }
            }
            return false;
        });
    *nf_out += nf;
    if (ANY && occluded) return 1;
    if (n == 0) return 0;
    if (n > kMaxCand) { *n_list = n; return 2; }
    // candidates that can still beat (or tie) the best sure hit; an undecided candidate has lo = 0
    const int cs[4] = {c0, c1, c2, c3};
    const float ls[4] = {l0, l1, l2, l3};
    int m = 0;
    for (int j = 0; j < n; j++)
        if (ANY || ls[j] <= best_hi) list[m++] = cs[j];
    *n_list = m;
    return 2;
}

// STAGED is a compile-time property of the kernel (sr_render.cu instantiates it both ways): two copies of the walk in
// one kernel cost the fused kernel 4.5 % on config2 through its instruction cache
template <bool ANY, bool STAGED = false>
__device__ __forceinline__ int spheres_filter(const DevScene& sc, d3 s, d3 dir, int* list, int* n_list, unsigned int* nv_out,
                                              unsigned int* nf_out, int* stack, const SphereStage* stage = nullptr)
{
    if (STAGED) return spheres_filter_impl<ANY, false>(sc, stage->nodes, stage->filt, s, dir, list, n_list, nv_out, nf_out, stack);
    return spheres_filter_impl<ANY, true>(sc, sc.sphere_nodes, sc.sph_filt, s, dir, list, n_list, nv_out, nf_out, stack);
}

// ---------------------------------------------------------------------------------------------
// Shadow bundles (DESIGN.md "Shadow bundles").  The softShadowQuality rays of one shading point all
// end in the same point `end` and start inside the ball of radius rho around the light: they lie in
// the cone  { end + tau * g : 0 <= tau <= 1, |g - g_c| <= rho },  g_c = light - end.  One conservative
// walk of that cone through the BVH tries to PROVE that every triangle is missed by every ray of the
// cone in the reference arithmetic; if it succeeds all rays escape and none has to be traced.
// Per triangle, any one of these suffices (bounds as in tri_filter; spread = rho |n|):
//   R1  g_c.n + spread < -E          every ray is back-facing (Plane.cs:75)
//   R2  d - end.n < -E               the plane lies beyond `end`: rayFrac > 1 for every ray
//   R4  tau1 > 1                     every ray meets the plane before its start
//   R3  every plane hit P(g) lies within Rp = rho tau2 + |g_c| (tau2 - tau1) of the axis hit P_c, and
//       s(P_c), u(P_c) are outside the triangle by more than a1 Rp, b1 Rp   (s, u are linear in P)
//   where [tau1, tau2] bounds num / (g.n) over the cone.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cone_slab(const FRay& r, float rho, float lox, float loy, float loz, float hix, float hiy,
                                          float hiz)
{
    // pass 1: a point of the cone inside the box has tau <= 1, so its axis point lies in the box grown
    // by rho (L-inf ball contains the L2 ball); pass 2: it then has tau <= tmax1, so grow by rho*tmax1 only
    float e = rho * r.tcull;
    float tmax1;
    {
        const float ax = __fmaf_rn(lox - e, r.ix, r.nox), bx = __fmaf_rn(hix + e, r.ix, r.nox);
        const float ay = __fmaf_rn(loy - e, r.iy, r.noy), by = __fmaf_rn(hiy + e, r.iy, r.noy);
        const float az = __fmaf_rn(loz - e, r.iz, r.noz), bz = __fmaf_rn(hiz + e, r.iz, r.noz);
        const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
        tmax1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), r.tcull));
        if (!(tmin <= tmax1)) return false;
    }
    e = rho * tmax1 * 1.00001f;
    const float ax = __fmaf_rn(lox - e, r.ix, r.nox), bx = __fmaf_rn(hix + e, r.ix, r.nox);
    const float ay = __fmaf_rn(loy - e, r.iy, r.noy), by = __fmaf_rn(hiy + e, r.iy, r.noy);
    const float az = __fmaf_rn(loz - e, r.iz, r.noz), bz = __fmaf_rn(hiz + e, r.iz, r.noz);
    const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), r.tcull));
    return tmin <= tmax;
}

// true: every ray of the cone surely misses this triangle (in the reference arithmetic)
__device__ __forceinline__ bool tri_cone_reject(const TriFilt* __restrict__ t, const FRay& r, float rho, float glen, float V)
{
    const float4* p = reinterpret_cast<const float4*>(t);
    const float4 q0 = __ldg(p);
    const float gn = __fmaf_rn(r.gz, q0.z, __fmaf_rn(r.gy, q0.y, r.gx * q0.x));
    const float e_gn = (8.0f * kU) * r.g1;
    const float spread = rho * (1.0f + 4.0f * kU);                     // |n| <= 1 + u
    if (gn + spread < -e_gn) return true;                              // R1
    const float num = __fmaf_rn(-r.oz, q0.z, __fmaf_rn(-r.oy, q0.y, __fmaf_rn(-r.ox, q0.x, q0.w)));
    const float e_num = (8.0f * kU) * (fabsf(q0.w) + r.o1);
    if (num < -e_num) return true;                                     // R2
    const float4 q1 = __ldg(p + 1);
    if (q1.w < 0.0f) return true;                                      // zero-area triangle: never hit
    const float gn_min = gn - spread - e_gn, gn_max = gn + spread + e_gn;
    if (!(gn_min > 16.0f * e_gn) || !(gn_min > 0.05f * gn_max)) return false;   // some rays graze the plane
    const float num_lo = fmaxf(num - e_num, 0.0f), num_hi = num + e_num;
    const float tau1 = __fdividef(num_lo, gn_max) * (1.0f - 8.0f * kU);
    const float tau2 = __fdividef(num_hi, gn_min) * (1.0f + 8.0f * kU);
    if (tau1 > r.tmax_hi) return true;                                 // R4
    // R3 around the axis hit
    const float rg = __fdividef(1.0f, gn);
    const float tau = num * rg;
    const float e_tau = (e_num + fabsf(tau) * e_gn) * rg * 1.1f + (4.0f * kU) * fabsf(tau);
    const float4 q3 = __ldg(p + 3);
    const float wx = __fmaf_rn(r.gx, tau, r.ox) - q3.x, wy = __fmaf_rn(r.gy, tau, r.oy) - q3.y,
                wz = __fmaf_rn(r.gz, tau, r.oz) - q3.z;
    const float k = r.ginf * e_tau + (12.0f * kU) * (r.oinf + r.ginf * fabsf(tau) + V);
    const float rp = (rho * tau2 + glen * (fmaxf(tau2, tau) - fminf(tau1, tau))) * 1.0001f + k;
    const float sN = __fmaf_rn(wz, q1.z, __fmaf_rn(wy, q1.y, wx * q1.x));
    const float m_s = q1.w * rp;
    if (sN < -m_s || sN > 1.0f + m_s) return true;
    const float4 q2 = __ldg(p + 2);
    const float uu = __fmaf_rn(wz, q2.z, __fmaf_rn(wy, q2.y, wx * q2.x));
    const float m_u = q2.w * rp;
    if (uu < -m_u) return true;
    if (sN + uu > 1.0f + m_s + m_u + 4.0f * kU) return true;
    return false;
}

// true: proven that no ray from the ball (light, rho) to `end` hits any triangle of the mesh
__device__ __forceinline__ bool bundle_clear(const DevMesh& m, d3 end, d3 light, float rho, int budget, Counters* c)
{
    FRay r;
    r.ox = __double2float_rn(end.x); r.oy = __double2float_rn(end.y); r.oz = __double2float_rn(end.z);
    r.gx = __double2float_rn(light.x - end.x); r.gy = __double2float_rn(light.y - end.y); r.gz = __double2float_rn(light.z - end.z);
    const float agx = fabsf(r.gx), agy = fabsf(r.gy), agz = fabsf(r.gz);
    const float aox = fabsf(r.ox), aoy = fabsf(r.oy), aoz = fabsf(r.oz);
    r.g1 = agx + agy + agz; r.ginf = fmaxf(agx, fmaxf(agy, agz));
    r.o1 = aox + aoy + aoz; r.oinf = fmaxf(aox, fmaxf(aoy, aoz));
    if (!(fminf(agx, fminf(agy, agz)) > 1e-30f) || !(r.ginf < 1e30f) || !(r.oinf <= 2.0f * m.scale)) return false;
    r.ix = __fdiv_rn(1.0f, r.gx); r.iy = __fdiv_rn(1.0f, r.gy); r.iz = __fdiv_rn(1.0f, r.gz);
    r.nox = -r.ox * r.ix; r.noy = -r.oy * r.iy; r.noz = -r.oz * r.iz;
    r.tmax_hi = 1.0f; r.tmax_lo = 1.0f; r.tcull = 1.00002f;
    // the centre g_c itself is rounded: widen the ball by that much
    const float rho_w = rho + (4.0f * kU) * r.ginf;
    const float glen = sqrtf(r.gx * r.gx + r.gy * r.gy + r.gz * r.gz) * (1.0f + 8.0f * kU);
    bool clear = true;
    unsigned int nf = 0;
    c->node_visits += walk_bvh(
        m.nodes, m.n_tris, c->stack,
        [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float* t) {
            *t = 0.0f;                                      // (no front-to-back order needed: every leaf must pass)
            if (--budget < 0) clear = false;                // too much geometry near the cone: trace the rays
            return clear && cone_slab(r, rho_w, lox, loy, loz, hix, hiy, hiz);
        },
        [&](int first, int count) {
            for (int i = 0; i < count; i++) {
                nf++;
                if (!tri_cone_reject(m.filt + first + i, r, rho_w, glen, m.scale)) { clear = false; return true; }
            }
            return false;
        });
    c->filter_tests += nf;
    return clear;
}

// ---------------------------------------------------------------------------------------------
// Shadow bundles with SUSPECTS (stage kernels).  bundle_clear above is all-or-nothing and its two-pass box test
// grows every box by rho * (exit parameter): wide enough that on a dense mesh (config3: triangles of 0.001 under
// a light of radius 0.2) the walk drowns in leaf boxes near the receiver and never succeeds.  Here:
//  * the cone is tested against a box EXACTLY for its L-infinity hull: a point of a ray of the bundle at parameter
//    tau is  o + (g_c + delta) tau  with |delta_k| <= rho, so on axis k it lies between o_k + (g_k - rho) tau and
//    o_k + (g_k + rho) tau.  "Some point of the hull is inside [lo, hi] for some tau in [0, tcull]" is six linear
//    inequalities in tau -- an interval intersection like the slab test, with slopes g_k -+ rho instead of g_k.
//    A slope whose sign FP32 cannot vouch for drops its inequality (conservative).
//  * the triangles no rule R1-R4 rejects are RETURNED (<= kMaxSuspects): the rays of the bundle then test only
//    those, with the per-ray filter, and never walk.  Every other triangle is proven to be missed by every ray of
//    the bundle in the reference arithmetic, so the answers are those of the per-ray walks.
// Returns the number of suspects (0: every ray escapes), or -1: too many / out of budget (walk the rays).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxSuspects = 8;

struct ConeAxes {
    float ia[3], ib[3];       // 1 / (g_k + rho), 1 / (g_k - rho)
    int sa[3], sb[3];         // sign of g_k + rho / g_k - rho: +1, -1, 0 = unknown (the inequality is dropped)
};

__device__ __forceinline__ bool cone_hull_slab(const FRay& r, const ConeAxes& ca, float lox, float loy, float loz, float hix, float hiy,
                                               float hiz)
{
    const float lo[3] = {lox, loy, loz}, hi[3] = {hix, hiy, hiz}, o[3] = {r.ox, r.oy, r.oz};
    float tmin = 0.0f, tmax = r.tcull;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float qa = (lo[k] - o[k]) * ca.ia[k];        // (g_k + rho) tau >= lo_k - o_k
        const float qb = (hi[k] - o[k]) * ca.ib[k];        // (g_k - rho) tau <= hi_k - o_k
        if (ca.sa[k] > 0) tmin = fmaxf(tmin, qa); else if (ca.sa[k] < 0) tmax = fminf(tmax, qa);
        if (ca.sb[k] > 0) tmax = fminf(tmax, qb); else if (ca.sb[k] < 0) tmin = fmaxf(tmin, qb);
    }
    // the boxes are padded in space (sr_bvh.cpp), which covers the roundings above as it does for fslab
    return tmin <= tmax;
}

__device__ __forceinline__ int bundle_suspects(const DevMesh& m, d3 end, d3 light, float rho, int budget, int* __restrict__ suspects,
                                               Counters* c)
{
    FRay r;
    r.ox = __double2float_rn(end.x); r.oy = __double2float_rn(end.y); r.oz = __double2float_rn(end.z);
    r.gx = __double2float_rn(light.x - end.x); r.gy = __double2float_rn(light.y - end.y); r.gz = __double2float_rn(light.z - end.z);
    const float agx = fabsf(r.gx), agy = fabsf(r.gy), agz = fabsf(r.gz);
    const float aox = fabsf(r.ox), aoy = fabsf(r.oy), aoz = fabsf(r.oz);
    r.g1 = agx + agy + agz; r.ginf = fmaxf(agx, fmaxf(agy, agz));
    r.o1 = aox + aoy + aoz; r.oinf = fmaxf(aox, fmaxf(aoy, aoz));
    if (!(r.ginf < 1e30f) || !(r.ginf > 1e-30f) || !(r.oinf <= 2.0f * m.scale)) return -1;
    r.ix = r.iy = r.iz = 0.0f; r.nox = r.noy = r.noz = 0.0f;       // (the hull test has its own reciprocals)
    r.tmin_hi = 0.0f; r.tmax_hi = 1.0f; r.tmax_lo = 1.0f; r.tcull = 1.00002f;
    // the centre g_c itself is rounded: widen the ball by that much
    const float rho_w = rho * (1.0f + 1e-5f) + (4.0f * kU) * r.ginf;
    const float glen = sqrtf(r.gx * r.gx + r.gy * r.gy + r.gz * r.gz) * (1.0f + 8.0f * kU);
    ConeAxes ca;
    {
        const float g[3] = {r.gx, r.gy, r.gz};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float a = g[k] + rho_w, b = g[k] - rho_w;
            const float eps = (16.0f * kU) * (fabsf(g[k]) + rho_w) + 1e-30f;
            ca.sa[k] = a > eps ? 1 : (a < -eps ? -1 : 0);
            ca.sb[k] = b > eps ? 1 : (b < -eps ? -1 : 0);
            ca.ia[k] = ca.sa[k] ? __fdiv_rn(1.0f, a) : 0.0f;
            ca.ib[k] = ca.sb[k] ? __fdiv_rn(1.0f, b) : 0.0f;
        }
    }
    int n = 0;
    int s0 = -1, s1 = -1, s2 = -1, s3 = -1, s4 = -1, s5 = -1, s6 = -1, s7 = -1;
    bool give_up = false;
    unsigned int nf = 0;
    c->node_visits += walk_bvh(
        m.nodes, m.n_tris, c->stack,
        [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float* t) {
            *t = 0.0f;                                      // (no front-to-back order needed: every leaf is looked at)
            if (--budget < 0) give_up = true;               // much geometry near the cone: trace the rays
            return !give_up && cone_hull_slab(r, ca, lox, loy, loz, hix, hiy, hiz);
        },
        [&](int first, int count) {
            for (int i = 0; i < count; i++) {
                nf++;
                if (tri_cone_reject(m.filt + first + i, r, rho_w, glen, m.scale)) continue;
                const int k = first + i;
                switch (n) {
                case 0: s0 = k; break; case 1: s1 = k; break; case 2: s2 = k; break; case 3: s3 = k; break;
                case 4: s4 = k; break; case 5: s5 = k; break; case 6: s6 = k; break; case 7: s7 = k; break;
                default: give_up = true; return true;
                }
                n++;
            }
            return false;
        });
    c->filter_tests += nf;
    if (give_up) return -1;
    suspects[0] = s0; suspects[1] = s1; suspects[2] = s2; suspects[3] = s3; suspects[4] = s4; suspects[5] = s5; suspects[6] = s6; suspects[7] = s7;
    return n;
}

// ---------------------------------------------------------------------------------------------
// rootGeometry: [ExtraGeometry spheres..., mesh through SpatialSubdivision | GeometryCollection]
// ---------------------------------------------------------------------------------------------
struct Hit {
    double rf;
    d3 pos, normal;
    uint32_t color;
    int32_t id;        // >= 0 triangle index, <= -2 sphere -(i+2)
};

constexpr double kNoHit = 1.7976931348623157e308;   // double.MaxValue (GeometryCollection.cs:48)

// Exact (reference arithmetic) nearest sphere / nearest triangle.
__device__ SR_EX_INLINE void spheres_closest_exact(const DevScene& sc, d3 s, d3 dirn, const int* list, int n_list, BestPrim* bs,
                                                   XCounters* c)
{
    if (n_list > 0) {       // only the spheres the filter could not rule out
        for (int j = 0; j < n_list; j++) {
            const int k = list[j];
            double rf;
            c->prim_tests++; c->sphere_tests++;
            if (!sphere_intersect(sc.spheres + k, s, dirn, &rf)) continue;
            const int index = __ldg(reinterpret_cast<const int*>(sc.spheres + k) + 11);
            if (rf < bs->rf || (rf == bs->rf && index < bs->index)) { bs->rf = rf; bs->k = k; bs->index = index; }
        }
        return;
    }
    if (sc.sphere_nodes) {
        double te;
        if (entry_clip(sc.sph_bmin, sc.sph_bmax, 0.0, s, dirn, &te)) {
            const TravRay tr = make_trav(s, dirn, te);
            walk<1, false>(sc.sphere_nodes, sc.spheres, tr, s, dirn, kNoHit, 0.0, bs, c);
        }
    } else {
        scan<1, false>(sc.spheres, sc.n_spheres, s, dirn, kNoHit, 0.0, bs, c);
    }
}

// n_list > 0: only the listed triangles (the filter has proven every other one missed or beaten)
__device__ SR_EX_INLINE void mesh_closest_exact(const DevMesh& m, int subdivision, d3 s, d3 dir, const int* list, int n_list,
                                                BestPrim* bt, d3* ts_out, double* offset_out, XCounters* c)
{
    d3 ts = s; double offset = 0.0;
    *ts_out = s; *offset_out = 0.0;
    if (subdivision) {
        // (a candidate list means the filter is on: SOFTRAY_FILTER_OFF / _VERIFY always take reference_clip)
        bool clipped = n_list > 0 && reference_clip_face(m.bmin, m.bmax, &ts, dir, &offset);
#ifdef SR_CLIP_CHECK
        if (clipped) {
            d3 ts2 = s; double offset2 = 0.0;
            const bool ok2 = reference_clip(m.bmin, m.bmax, &ts2, dir, &offset2);
            if (!ok2 || ts2.x != ts.x || ts2.y != ts.y || ts2.z != ts.z || offset2 != offset) c->filter_mismatch++;
        }
#endif
        if (!clipped && !reference_clip(m.bmin, m.bmax, &ts, dir, &offset)) return;
    }
    *ts_out = ts; *offset_out = offset;
    if (n_list > 0) {
        for (int j = 0; j < n_list; j++) {
            const int k = list[j];
            double rf;
            c->prim_tests++;
            if (!tri_intersect(m.tris + k, ts, dir, bt->rf, &rf)) continue;
            const int index = __ldg(reinterpret_cast<const int*>(m.tris + k) + 31);
            if (rf < bt->rf || (rf == bt->rf && index < bt->index)) { bt->rf = rf; bt->k = k; bt->index = index; }
        }
        return;
    }
    if (m.nodes) {
        double te = 0.0;
        // the clipped start already lies on/in the root box; otherwise enter it first
        if (subdivision || entry_clip(m.bmin, m.bmax, 0.0, ts, dir, &te)) {
            const TravRay tr = make_trav(ts, dir, te);
            walk<0, false>(m.nodes, m.tris, tr, ts, dir, kNoHit, 0.0, bt, c);
        }
    } else {
        scan<0, false>(m.tris, m.n_tris, ts, dir, kNoHit, 0.0, bt, c);
    }
}

// IRayIntersectable.IntersectRay of rootGeometry for a camera / reflection ray: the nearest hit.
// limit: mesh hits with a rayFrac beyond it cannot matter to the caller (kNoHit: none).  Only prunes the search.
template <bool STAGED = false>
__device__ SR_CH_INLINE bool closest_hit(const DevScene& sc, const DevMesh& m, int subdivision, int filter_mode, d3 s, d3 dir,
                                            Hit* h, XCounters* c, int sync, double limit = 1.7976931348623157e308,
                                            const SphereStage* stage = nullptr)
{
    // --- spheres (tested first in list order) ---
    BestPrim bs; bs.rf = kNoHit; bs.k = -1; bs.index = 0x7fffffff;
    d3 dirn = dir;
    if (sc.n_spheres > 0) {
        dirn = vnormalise(dir);                          // Sphere.cs:160
        int known = 2;                                   // 0: surely no sphere; 2: look at list (or at all of them)
        int list[kMaxCand]; int n_list = 0;
        if (filter_mode != 1 && sc.sphere_nodes != nullptr) {
            unsigned int nv = 0, nf = 0;
            known = spheres_filter<false, STAGED>(sc, s, dir, list, &n_list, &nv, &nf, c->stack, stage);
            c->node_visits += nv; c->filter_tests += nf;
        }
        SR_SYNC_POINT(sync & 2);
        const bool listed = known == 2 && n_list >= 1 && n_list <= kMaxCand;
        const bool verify = filter_mode == 2;
        const bool use_list = listed && !verify;
        // one call site (the function is inlined): the listed spheres, or all of them (VERIFY; an undecided filter)
        if (verify || known == 2) spheres_closest_exact(sc, s, dirn, use_list ? list : nullptr, use_list ? n_list : 0, &bs, c);
        if (verify) {
            bool in_list = bs.k < 0 || !listed;
            for (int j = 0; listed && j < n_list; j++) in_list = in_list || list[j] == bs.k;
            if ((known == 0 && bs.k >= 0) || !in_list) c->filter_mismatch++;
            if (known == 2 && n_list != 1) c->filter_unsure++;
        } else if (listed) {
            if (n_list > 1) c->filter_unsure++;
        } else if (known == 2) {
            if (filter_mode != 1 && sc.sphere_nodes != nullptr) c->filter_unsure++;
        }
    }
    SR_SYNC_POINT(sync & 4);
    // --- mesh ---
    BestPrim bt; bt.rf = kNoHit; bt.k = -1; bt.index = 0x7fffffff;
    d3 ts = s; double offset = 0.0;
    if (m.n_tris > 0) {
        // 0: surely no hit; 1: the winner is among list[0..n_list) (one entry: it IS the winner); 2: full exact walk
        int known = 2;
        int list[kMaxCand]; int n_list = 0;
        bool sure_hit = false;
        if (filter_mode != 1 && m.nodes != nullptr) {
            FRay r; double t0;
            known = fray_setup_fwd(m, subdivision, s, dir, &r, &t0);
            if (known == 1) {
                FClosest fc;
                // rayFrac = t0 + tau, so tau <= limit - t0 (rounded up) keeps every hit that can still win or tie
                // (a sphere already hit limits it too: a triangle wins only with a SMALLER rayFrac, GeometryCollection.cs:53)
                const double lim = (filter_mode != 2 && bs.k >= 0 && bs.rf < limit) ? bs.rf : limit;
                const float limit_tau = lim < 1e300 ? __double2float_ru(lim - t0) * (1.0f + 4.0f * kU) + 1e-6f : 1e30f;
                walk_filter_closest(m.nodes, m.filt, m.n_tris, r, m.scale, &fc, c, limit_tau);
                sure_hit = fc.best_k >= 0;
                if (fc.n_cand == 0) known = 0;
                else if (fc.n_cand > kMaxCand) known = 2;
                else {
                    for (int j = 0; j < fc.n_cand; j++)
                        if (fc.cand[j] == fc.best_k || fc.cand_lo[j] <= fc.best_hi) list[n_list++] = fc.cand[j];
                }
            }
        }
        SR_SYNC_POINT(sync & 8);
        const bool verify = filter_mode == 2;
        const bool use_list = known == 1 && !verify;
        // one call site (the function is inlined): the candidates, or the full exact walk (VERIFY; an undecided filter)
        if (verify || known != 0)
            mesh_closest_exact(m, subdivision, s, dir, use_list ? list : nullptr, use_list ? n_list : 0, &bt, &ts, &offset, c);
        if (verify) {
            bool in_list = bt.k < 0;
            for (int j = 0; j < n_list; j++) in_list = in_list || list[j] == bt.k;
            // contradictions: "surely nothing" but the exact walk hits; the exact winner is not a candidate;
            // a sure hit exists but the exact walk finds nothing
            if ((known == 0 && bt.k >= 0) || (known == 1 && !in_list) || (known == 1 && sure_hit && bt.k < 0))
                c->filter_mismatch++;
            if (known == 2 || n_list > 1) c->filter_unsure++;
        } else if (known == 1) {
            if (n_list > 1) c->filter_unsure++;
        } else if (known == 2) {
            if (filter_mode != 1 && m.nodes != nullptr) c->filter_unsure++;
        }
    }
    SR_SYNC_POINT(sync & 16);
    const double rf_tri = bt.k >= 0 ? dadd(bt.rf, offset) : kNoHit;   // SpatialSubdivision.cs:416
    if (bt.k >= 0 && rf_tri < bs.rf) {
        const TriRec* t = m.tris + bt.k;
        const double2 a0 = ldg2(t, 0), a1 = ldg2(t, 1);
        h->rf = rf_tri;
        h->pos = vadd(ts, vscale(dir, bt.rf));             // Plane.cs:86 from the clipped start
        h->normal = mk(a0.x, a0.y, a1.x);
        h->color = __ldg(reinterpret_cast<const uint32_t*>(t) + 30);
        h->id = bt.index;
        return true;
    }
    if (bs.k >= 0) {
        const SphereRec* q = sc.spheres + bs.k;
        const double2 a0 = ldg2(q, 0), a1 = ldg2(q, 1);
        h->rf = bs.rf;
        h->pos = vadd(s, vscale(dirn, bs.rf));             // Sphere.cs:198
        h->normal = vnormalise(vsub(h->pos, mk(a0.x, a0.y, a1.x)));
        h->color = __ldg(reinterpret_cast<const uint32_t*>(q) + 10);
        h->id = -(bs.index + 2);
        return true;
    }
    return false;
}

// "shadowInfo != null && shadowInfo.rayFrac <= 1.0" (ShadowMethod.cs:171): true iff ANY primitive
// reports a rayFrac <= 1.0, because the minimum of the reported rayFracs is what the chain returns.
// Exact (reference arithmetic) versions, one per primitive kind.
static __device__ __noinline__ bool occluded_mesh(const DevMesh& m, int subdivision, d3 s, d3 dir, const int* list, int n_list,
                                          XCounters* c)
{
    BestPrim dummy; dummy.rf = kNoHit; dummy.k = -1; dummy.index = 0;
    if (m.n_tris <= 0) return false;
    d3 ts = s; double offset = 0.0;
    if (subdivision && !reference_clip(m.bmin, m.bmax, &ts, dir, &offset)) return false;
    // rf + offset <= 1.0 needs rf <= 1.0 - offset (+ an ulp of slack for the walk's cull)
    const double limit = (1.0 - offset) * (1.0 + 1e-12) + 1e-300;
    if (n_list > 0) {       // only the triangles the filter could not decide; every other one is surely missed
        for (int j = 0; j < n_list; j++) {
            double rf;
            c->prim_tests++;
            if (tri_intersect(m.tris + list[j], ts, dir, limit, &rf) && dadd(rf, offset) <= 1.0) return true;
        }
        return false;
    }
    if (m.nodes) {
        double te = 0.0;
        // the clipped start already lies on/in the root box; otherwise enter it first
        if (subdivision || entry_clip(m.bmin, m.bmax, 0.0, ts, dir, &te)) {
            const TravRay tr = make_trav(ts, dir, te);
            return walk<0, true>(m.nodes, m.tris, tr, ts, dir, limit, offset, &dummy, c);
        }
        return false;
    }
    return scan<0, true>(m.tris, m.n_tris, ts, dir, limit, offset, &dummy, c);
}

static __device__ __noinline__ bool occluded_spheres(const DevScene& sc, d3 s, d3 dir, const int* list, int n_list, XCounters* c)
{
    BestPrim dummy; dummy.rf = kNoHit; dummy.k = -1; dummy.index = 0;
    if (sc.n_spheres <= 0) return false;
    const d3 dirn = vnormalise(dir);
    if (n_list > 0) {       // only the spheres the filter could not decide
        for (int j = 0; j < n_list; j++) {
            double rf;
            c->prim_tests++; c->sphere_tests++;
            if (sphere_intersect(sc.spheres + list[j], s, dirn, &rf) && rf <= 1.0) return true;
        }
        return false;
    }
    if (sc.sphere_nodes) {
        double te;
        if (entry_clip(sc.sph_bmin, sc.sph_bmax, 0.0, s, dirn, &te) && te <= 1.0) {
            const TravRay tr = make_trav(s, dirn, te);
            return walk<1, true>(sc.sphere_nodes, sc.spheres, tr, s, dirn, 1.0, 0.0, &dummy, c);
        }
        return false;
    }
    return scan<1, true>(sc.spheres, sc.n_spheres, s, dirn, 1.0, 0.0, &dummy, c);
}

// ---------------------------------------------------------------------------------------------
// ShadingMethod (ShadingMethod.cs:36-68,110-177) + Instance.TransformPosToView (Instance.cs:168-184)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t shade(const DevFrame& f, const DevInstance& in, d3 pos, d3 normal, uint32_t color)
{
    d3 v = mul3x4(in.M, pos);
    const double vz = v.z;
    v.x = dmul(ddiv(v.x, vz), f.fov_depth);
    v.y = dmul(ddiv(v.y, vz), f.fov_depth);
    v.z = dmul(dadd(dsub(vz, in.pos_z), 1.0), 0.5);
    const d3 nv = mul3x3(in.M, normal);
    d3 to_light;
    if (f.point_lighting) to_light = vnormalise(vsub(mk(f.light_pos_view[0], f.light_pos_view[1], f.light_pos_view[2]), v));
    else to_light = vneg(mk(f.light_dir_view[0], f.light_dir_view[1], f.light_dir_view[2]));
    const double ldn = vdot(to_light, nv);
    const double diffuse = fmax(0.0, ldn);
    double specular = 0.0;
    const double amb_diff = dadd(f.ambient, diffuse);
    // ambient + diffuse >= 1: min(1, . + specular) is 1 whatever the (non-negative) specular term is
    if (f.specular_lighting && amb_diff < 1.0) {
        const d3 to_cam = vnormalise(vneg(v));
        const d3 refl = vsub(vscale(nv, dmul(2.0, ldn)), to_light);
        const double cos_a = vdot(refl, to_cam);
        // Math.Pow then Math.Max (:153-154).  |pow(c, s)| <= |c|^s (or the result is NaN / negative and Max makes it 0):
        // below spec_skip = 2^(-82/s) it is < 2^-82, less than half an ulp of any ambient + diffuse >= 2^-20, so adding
        // it returns ambient + diffuse bit for bit -- the ~300-instruction FP64 pow is only run where it can matter
        if (!(fabs(cos_a) < f.spec_skip && amb_diff >= 9.5367431640625e-07))
            specular = fmax(0.0, pow(cos_a, f.shininess));
    }
    double ch = dadd(amb_diff, specular);                   // white material, (a + d) + s
    ch = fmin(ch, 1.0);
    return modulate(color, to_byte(dmul(255.0, ch)));
}

// One shadow ray of ShadowMethod.TraceRaysForSoftShadows (ShadowMethod.cs:147-177): sample i of the
// area light towards `end`.  Reference arithmetic for start / dir; exact versions on demand.
__device__ __forceinline__ void shadow_ray(const DevFrame& f, const DevInstance& in, const double* __restrict__ offsets, d3 end,
                                           int i, d3* start, d3* dir)
{
    const d3 off = mk(offsets[3 * i], offsets[3 * i + 1], offsets[3 * i + 2]);
    if (f.point_lighting) {
        *start = vadd(mk(in.light_pos_model[0], in.light_pos_model[1], in.light_pos_model[2]), off);
        *dir = vsub(end, *start);
    } else {
        *dir = mk(in.light_dir_model[0], in.light_dir_model[1], in.light_dir_model[2]);
        *start = vadd(vadd(end, vscale(*dir, 1000.0)), off);
    }
}

}  // namespace sr
