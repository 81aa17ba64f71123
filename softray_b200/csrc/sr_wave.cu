// sr_wave.cu -- the raytrace hot path as STAGE KERNELS (DESIGN.md "Stage kernels"): the same per-ray arithmetic
// as the fused kernel (sr_render.cu; both include sr_device.cuh), cut where the fused kernel's own profile says it
// hurts -- a 38.8 K-instruction program against a 32 KB instruction cache, 80 registers per thread (37 % occupancy)
// under latency-bound walks, half the lanes idle because the rays of a warp need different stages at different
// times, and a frame whose slowest tile holds an SM while the others idle.
//
// One frame = chunks of 8x4-pixel tiles; per chunk (records cross HBM between the stages):
//   search<CAMERA>   ray generation (exact FP64) + FP32 filtered closest-hit walk (instance hierarchy for composite
//                    frames)                                   -> <= 4 candidate triangles per sample (Cand, 20 B)
//   hit<CAMERA>      the candidates through the reference arithmetic (root-box clip + Triangle.IntersectRay) ->
//                    winner; Texture3D + ShadingMethod        -> slot colour (4 B), shadow work item (32 B),
//                                                                 reflection ray (56 B); undecided rays -> a list
//   fallback<CAMERA> the full exact walk for the few rays the search could not bracket (compacted: whole warps of them)
//   search/hit/fallback<LIST>   the same for the reflection rays, once per bounce
//   shadow           ShadowMethod.TraceRaysForSoftShadows over the compacted shading points: cone test, then the
//                    filtered any-hit rays; undecided rays -> a list      -> escaped count per slot
//   shadow_fallback  those rays through the reference arithmetic
//   compose          ModulatePackedColor by the shadow byte, mirror blend chain, sub-pixel sums, packed-ARGB store
// Every stage keeps the fused kernel's arithmetic, so frames are identical bit for bit (tests/test_cuda_wave.py
// requires wave == fused == oracle).  Replaces RaytraceBlock -> TraceRayComplex -> IRayIntersectable.IntersectRay ->
// Surface.DrawPixel (Engine3D/Renderer.cs:1690-1925 and the Raytrace/*Method.cs decorators), like sr_render.cu.
#ifndef SR_WAVE_WALK_MODE
#define SR_WAVE_WALK_MODE 0      // traversal loop of the stage kernels: 0 if-if, 1 while-while (measured: config3 41.2 -> 48.4 ms, config5 14.5 -> 15.0)
#endif
#define SR_WALK_MODE SR_WAVE_WALK_MODE
#include "sr_device.cuh"
#include "sr_wave.h"
#include "../../include/softray_cuda.h"

#include <cstdlib>

namespace sr {

namespace {

constexpr int kWaveThreads = 256;

enum : uint32_t { kStateInvalid = 0, kStateMiss = 1, kStateListed = 2, kStateUndecided = 3 };

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Append one record per lane with `want` to a device queue: one atomic per warp.  Must be reached by all 32 lanes.
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool want)
{
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    uint32_t base = 0;
    if (mask != 0u && lane_id() == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, mask ? __ffs(mask) - 1 : 0);
    return base + (uint32_t)__popc(mask & ((1u << lane_id()) - 1u));
}

// One atomic per warp per non-zero counter.  The per-thread counts are 32-bit (a thread sees a few thousand rays per
// launch at most), so the warp sum is ONE redux instruction instead of ten shuffles.
__device__ __forceinline__ void flush_counters(DevCounters* counters, const unsigned long long (&v)[14])
{
#pragma unroll
    for (int k = 0; k < 14; k++) {
        const unsigned int x = __reduce_add_sync(0xffffffffu, (unsigned int)v[k]);
        if (lane_id() == 0 && x) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + k, (unsigned long long)x);
    }
}

// ---------------------------------------------------------------------------------------------
// rays of a batch
// ---------------------------------------------------------------------------------------------
// RaytraceBlock's ray generation (Renderer.cs:1717-1797) for work item w of a tile: the same operations as the
// fused kernel's loop body.  Returns false for items outside the image / band.
__device__ __forceinline__ bool camera_ray(const DevFrame& f, const DevInstance& in0, int tile, int w, d3* start, d3* dir, bool* is_view)
{
    const int n = f.sub_pixel_res, nn = n * n;
    const int W = f.width, H = f.height;
    const int ty = (int)f.fd_tiles_x.div((uint32_t)tile), tx = tile - ty * f.tiles_x;
    const int band_j = (int)f.fd_tiles_per_band.div((uint32_t)ty);
    const int row0 = f.start_row + (f.band_index + band_j * f.band_count) * f.band_height;
    const int band_r0 = (ty - band_j * f.tiles_per_band) * 4;
    const int px = (int)f.fd_nn.div((uint32_t)w), si = w - px * nn;
    const int col = tx * 8 + (px & 7);
    const int band_r = band_r0 + (px >> 3);
    const int row = row0 + band_r;
    if (col >= W || band_r >= f.band_height || row > f.end_row) return false;
    const int sx = (int)f.fd_n.div((uint32_t)si), sy = si - sx * n;               // subX outer, subY inner (:1762-1764)
    double fx = 0.0, fy = 0.0;
    if (n > 1) {
        fx = dsub(ddiv((double)sx, (double)(n - 1)), 0.5);                        // :1767-1768
        fy = dsub(ddiv((double)sy, (double)(n - 1)), 0.5);
    }
    if (f.focal_blur && n > 1) {                                                  // App. A #11
        const d3 dir_view = mk(-dsub(ddiv((double)col, (double)W), 0.5), dmul(-dsub(ddiv((double)row, (double)H), 0.5), f.aspect),
                               f.fov_depth);
        const d3 dw = mul3x3(in0.Minv, dir_view);
        const d3 focal_pt = vadd(vscale(dw, f.focal_depth), mk(in0.start[0], in0.start[1], in0.start[2]));   // :1759
        const d3 sv = mk(dmul(ddiv(fx, (double)W), f.focal_strength), dmul(ddiv(fy, (double)H), f.focal_strength), -in0.pos_z);
        *start = mul3x3(in0.Minv, sv);                                            // :1776-1778
        *dir = vsub(focal_pt, *start);                                            // :1790
        *is_view = false;
    } else {
        *start = mk(in0.start[0], in0.start[1], in0.start[2]);
        *dir = mk(-dsub(ddiv(dadd((double)col, fx), (double)W), 0.5), dmul(-dsub(ddiv(dadd((double)row, fy), (double)H), 0.5), f.aspect),
                  f.fov_depth);                                                   // :1728, :1794-1796
        *is_view = true;
    }
    return true;
}

// SRC 0: camera rays of the chunk (ray r = sample r); SRC 1: the reflection rays of list `a.ref_in`.
// For a single-instance frame *dir is in model space; for a composite frame it stays in view space (every instance
// rotates it for itself).
template <int SRC>
__device__ __forceinline__ bool batch_ray(const WaveArgs& a, const DevInstance* __restrict__ insts, uint32_t r, d3* start, d3* dir,
                                          uint32_t* sample, uint32_t* depth)
{
    const DevFrame& f = a.f;
    if (SRC == 0) {
        const int nn = f.sub_pixel_res * f.sub_pixel_res;
        const int per_tile = 32 * nn;
        const int t = (int)f.fd_per_tile.div(r), w = (int)(r - (uint32_t)t * (uint32_t)per_tile);
        bool is_view;
        *sample = r; *depth = 0;
        if (!camera_ray(f, insts[0], a.tile0 + t, w, start, dir, &is_view)) return false;
        if (f.n_instances == 1 && is_view) *dir = mul3x3(insts[0].Minv, *dir);
        return true;
    } else {
        const RefRay* q = a.b.ref[a.ref_in] + r;
        const double2 v0 = ldg2(q, 0), v1 = ldg2(q, 1), v2 = ldg2(q, 2);
        *start = mk(v0.x, v0.y, v1.x);
        *dir = mk(v1.y, v2.x, v2.y);
        const uint2 sd = __ldg(reinterpret_cast<const uint2*>(q) + 6);
        *sample = sd.x; *depth = sd.y;
        return true;
    }
}

// ---------------------------------------------------------------------------------------------
// search: candidates of one ray
// ---------------------------------------------------------------------------------------------
struct Search {
    float best_hi;                 // upper bound (ray-parameter units) of the nearest SURE hit so far
    int n;                         // candidates listed (> kMaxCand: too many)
    int k[kMaxCand];
    int inst[kMaxCand];
    float lo[kMaxCand];            // lower bound of each candidate's rayFrac
    bool undecided;
};

__device__ __forceinline__ void search_add(Search& st, int inst, int k, float lo)
{
    if (st.n >= kMaxCand) {
        // drop what a nearer sure hit has already beaten, then try again
        int m = 0;
#pragma unroll
        for (int j = 0; j < kMaxCand; j++)
            if (st.lo[j] <= st.best_hi) { st.k[m] = st.k[j]; st.inst[m] = st.inst[j]; st.lo[m] = st.lo[j]; m++; }
        st.n = m;
        if (m >= kMaxCand) { st.n = kMaxCand + 1; return; }
    }
#pragma unroll
    for (int j = 0; j < kMaxCand; j++)
        if (j == st.n) { st.k[j] = k; st.inst[j] = inst; st.lo[j] = lo; }
    st.n++;
}

// The mesh of one instance against one ray (model space): merges its candidates into st.  Same walk as the
// fused kernel's closest_hit (fray_setup_fwd + walk_filter_closest); rayFrac = t0 + tau.
__device__ __forceinline__ void search_mesh(const DevMesh& m, int subdivision, d3 s, d3 dir, int inst, Search& st, XCounters* c)
{
    if (m.n_tris <= 0) return;
    FRay r; double t0;
    const int known = fray_setup_fwd(m, subdivision, s, dir, &r, &t0);
    if (known == 0) return;
    if (known == 2) { st.undecided = true; return; }
    const float limit_tau = st.best_hi < 1e29f ? __double2float_ru((double)st.best_hi - t0) * (1.0f + 4.0f * kU) + 1e-6f : 1e30f;
    if (limit_tau < 0.0f) return;                       // the whole mesh lies behind a sure hit of another instance
    FClosest fc;
    walk_filter_closest(m.nodes, m.filt, m.n_tris, r, m.scale, &fc, c, limit_tau);
    if (fc.n_cand == 0) return;
    if (fc.n_cand > kMaxCand) { st.undecided = true; return; }
    if (fc.best_k >= 0) {
        const float hi = __double2float_ru(t0 + (double)fc.best_hi) * (1.0f + 2.0f * kU);
        if (hi < st.best_hi) st.best_hi = hi;
    }
#pragma unroll
    for (int j = 0; j < kMaxCand; j++) {
        if (j < fc.n_cand && (fc.cand[j] == fc.best_k || fc.cand_lo[j] <= fc.best_hi)) {
            const float lo = __double2float_rd(t0 + (double)fc.cand_lo[j]) * (1.0f - 2.0f * kU);
            if (st.n <= kMaxCand) search_add(st, inst, fc.cand[j], lo);
        }
    }
}

template <int SRC, int MINB>
__global__ void __launch_bounds__(kWaveThreads, MINB) k_search(const __grid_constant__ WaveArgs a)
{
    const DevFrame& f = a.f;
    const DevInstance* __restrict__ insts = a.insts;
    int walk_stack[kStackEntries];
    XCounters xc; xc.stack = walk_stack; xc.node_visits = 0; xc.prim_tests = 0; xc.sphere_tests = 0; xc.filter_tests = 0;
    xc.filter_unsure = 0; xc.filter_mismatch = 0;
    const uint32_t n_rays = SRC == 0 ? a.n_rays : __ldg(&a.b.counts->n_ref[a.ref_in]);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_rays; base += stride) {
        const uint32_t r = base + lane_id();
        if (r >= n_rays) continue;
        d3 s, dir; uint32_t sample, depth;
        if (!batch_ray<SRC>(a, insts, r, &s, &dir, &sample, &depth)) { a.b.meta[r] = kStateInvalid; continue; }
        Search st; st.best_hi = 1e30f; st.n = 0; st.undecided = false;
#pragma unroll
        for (int j = 0; j < kMaxCand; j++) { st.k[j] = -1; st.inst[j] = 0; st.lo[j] = 0.0f; }
        if (f.n_instances == 1) {
            search_mesh(a.sc.meshes[insts[0].mesh], f.subdivision, s, dir, 0, st, &xc);
        } else {
            // composite frame: the hierarchy over the view-space boxes of the instances, front to
            // back; rigid transforms keep |dir|, so every instance's rayFrac is the parameter along the view ray
            FRay tr;
            tr.ox = tr.oy = tr.oz = 0.0f;
            tr.gx = __double2float_rn(dir.x); tr.gy = __double2float_rn(dir.y); tr.gz = __double2float_rn(dir.z);
            tr.ix = __fdiv_rn(1.0f, tr.gx); tr.iy = __fdiv_rn(1.0f, tr.gy); tr.iz = __fdiv_rn(1.0f, tr.gz);
            tr.nox = tr.noy = tr.noz = 0.0f;
            tr.tcull = CUDART_INF_F;
            int stack[kTlasStackEntries];
            int sp = 0, cur = 0;
            for (;;) {
                if (cur >= 0) {
                    const float4* p = reinterpret_cast<const float4*>(f.tlas_nodes + cur);
                    const float4 A = __ldg(p), B = __ldg(p + 1), CC = __ldg(p + 2);
                    const int4 d = __ldg(reinterpret_cast<const int4*>(p + 3));
                    xc.node_visits++;
                    float t0, t1;
                    const bool h0 = fslab(tr, A.x, A.y, A.z, A.w, B.x, B.y, &t0);
                    const bool h1 = fslab(tr, B.z, B.w, CC.x, CC.y, CC.z, CC.w, &t1);
                    if (h0 && h1) {
                        const bool first0 = t0 <= t1;
                        stack[sp++] = first0 ? d.y : d.x;
                        cur = first0 ? d.x : d.y;
                        continue;
                    }
                    if (h0) { cur = d.x; continue; }
                    if (h1) { cur = d.y; continue; }
                } else {
                    const int code = -1 - cur;
                    const int first = code >> 4, count = code & 15;
                    for (int j = 0; j < count; j++) {
                        const int i = __ldg(f.tlas_order + first + j);
                        const DevInstance& in = insts[i];
                        {   // the view ray (from the view origin) against the instance's bounding sphere: the hierarchy's boxes are
                            // the AABBs of rotated boxes, three times the cross-section of a round mesh, and an instance that
                            // is entered costs a ray set-up and a partial walk (config4: 3.6 of the 4.6 instances a ray enters)
                            const d3 cv = mk(in.bs_center_view[0], in.bs_center_view[1], in.bs_center_view[2]);
                            const double b = cv.x * dir.x + cv.y * dir.y + cv.z * dir.z;
                            const double cc = cv.x * cv.x + cv.y * cv.y + cv.z * cv.z;
                            const double dd = dir.x * dir.x + dir.y * dir.y + dir.z * dir.z;
                            const double out2 = cc - in.bs_radius2;                     // > 0: the origin is outside the sphere
                            if (out2 > 0.0 && (b <= 0.0 || b * b < out2 * dd * (1.0 - 1e-9))) continue;
                        }
                        search_mesh(a.sc.meshes[in.mesh], f.subdivision, mk(in.start[0], in.start[1], in.start[2]), mul3x3(in.Minv, dir), i,
                                    st, &xc);
                        if (st.best_hi < 1e29f) tr.tcull = st.best_hi * 1.00002f + 1e-6f;
                    }
                }
                if (sp == 0 || st.undecided || st.n > kMaxCand) break;
                cur = stack[--sp];
            }
        }
        uint32_t state;
        int4 cand = make_int4(-1, -1, -1, -1);
        uint32_t insts_of = 0;
        if (st.undecided || st.n > kMaxCand) state = kStateUndecided;
        else {
            int m = 0;
            int ck[kMaxCand] = {-1, -1, -1, -1};
#pragma unroll
            for (int j = 0; j < kMaxCand; j++)
                if (j < st.n && st.lo[j] <= st.best_hi) {
#pragma unroll
                    for (int q = 0; q < kMaxCand; q++)
                        if (q == m) { ck[q] = st.k[j]; insts_of |= (uint32_t)st.inst[j] << (4 + 7 * q); }
                    m++;
                }
            cand = make_int4(ck[0], ck[1], ck[2], ck[3]);
            state = m == 0 ? kStateMiss : kStateListed;
        }
        a.b.cand[r] = cand;
        a.b.meta[r] = state | insts_of;
    }
    const unsigned long long v[14] = {0, 0, 0, xc.node_visits, 0, 0, 0, 0, xc.filter_tests, 0, 0, 0, 0, 0};
    flush_counters(a.counters, v);
}

// ---------------------------------------------------------------------------------------------
// hit: exact winner among the candidates, then Texture3D + ShadingMethod, then the work the hit spawns
// ---------------------------------------------------------------------------------------------
struct ExactHit { bool hit; int inst; int k; int index; double rf_from_clip; double rf; d3 clipped_start; d3 dir; };

// What the fused kernel's trace_camera_ray does after closest_hit returned: record the shading point.
// Returns through the flags what this lane appends to the queues.  All lanes of the warp call it.
template <int SRC>
__device__ __forceinline__ void spawn_from_hit(const WaveArgs& a, const DevInstance* __restrict__ insts, bool active, bool hit,
                                               const ExactHit& eh, uint32_t sample, uint32_t depth, unsigned int* n_shaded,
                                               unsigned int* n_shadow, unsigned int* n_secondary, unsigned int* n_hits)
{
    const DevFrame& f = a.f;
    const uint32_t S = a.n_samples;
    bool want_shadow = false, want_ref = false;
    ShadowItem item; RefRay ray;
    if (active) {
        if (!hit) {
            if (SRC == 0) a.b.sample_state[sample] = 0x80u;                      // valid, no shading point: background
            else a.b.sample_state[sample] = (uint8_t)(0x80u | 0x08u | depth);    // `depth` shading points, then the background (tail)
        } else {
            const DevInstance& in = insts[eh.inst];
            const DevMesh& m = a.sc.meshes[in.mesh];
            const TriRec* t = m.tris + eh.k;
            const double2 a0 = ldg2(t, 0), a1 = ldg2(t, 1);
            const d3 pos = vadd(eh.clipped_start, vscale(eh.dir, eh.rf_from_clip));   // Plane.cs:86 from the clipped start
            const d3 normal = mk(a0.x, a0.y, a1.x);
            uint32_t color = __ldg(reinterpret_cast<const uint32_t*>(t) + 30);
            if (SRC == 0) { a.b.sample_id[sample] = in.tri_base + eh.index; (*n_hits)++; }
            (*n_shaded)++;
            if (f.texture3d_id) color = modulate(color, texture3d_sample(f.texture3d_id, pos));
            if (f.shading) color = shade(f, in, pos, normal, color);
            const uint32_t slot = depth * S + sample;
            a.b.slot_color[slot] = color;
            if (f.shadows) a.b.slot_escaped[slot] = 0u;          // (the shadow kernels add to it)
            a.b.sample_state[sample] = (uint8_t)(0x80u | (depth + 1u));
            if (f.shadows) {
                const d3 end = vadd(pos, vscale(normal, 0.001));                  // ShadowMethod.cs:150
                item.end[0] = end.x; item.end[1] = end.y; item.end[2] = end.z; item.slot = slot; item.inst = (uint32_t)eh.inst;
                want_shadow = true;
                *n_shadow += (unsigned int)f.shadow_samples;
            }
            const int bounces = (f.reflection_depth > 0 && f.n_instances == 1) ? f.reflection_depth : 0;
            if ((int)depth < bounces) {
                // r = d - n * (2 (d.n)), from pos + n*0.001 (PathTracingMethod.cs:10,52)
                const d3 rd = vsub(eh.dir, vscale(normal, dmul(2.0, vdot(eh.dir, normal))));
                const d3 rs = vadd(pos, vscale(normal, 0.001));
                ray.o[0] = rs.x; ray.o[1] = rs.y; ray.o[2] = rs.z; ray.d[0] = rd.x; ray.d[1] = rd.y; ray.d[2] = rd.z;
                ray.sample = sample; ray.depth = depth + 1u;
                want_ref = true;
                (*n_secondary)++;
            }
        }
    }
    if (f.shadows) {
        const uint32_t at = warp_append(&a.b.counts->n_shadow, want_shadow);
        if (want_shadow) a.b.shadow[at] = item;
    }
    if (f.reflection_depth > 0 && f.n_instances == 1) {
        const uint32_t at = warp_append(&a.b.counts->n_ref[a.ref_out], want_ref);
        if (want_ref) a.b.ref[a.ref_out][at] = ray;
    }
}

template <int SRC, int MINB>
__global__ void __launch_bounds__(kWaveThreads, MINB) k_hit(const __grid_constant__ WaveArgs a)
{
    const DevFrame& f = a.f;
    const DevInstance* __restrict__ insts = a.insts;
    int walk_stack[4];                  // (the listed-candidate evaluators never walk)
    XCounters xc; xc.stack = walk_stack; xc.node_visits = 0; xc.prim_tests = 0; xc.sphere_tests = 0; xc.filter_tests = 0;
    xc.filter_unsure = 0; xc.filter_mismatch = 0;
    unsigned int n_primary = 0, n_shadow = 0, n_secondary = 0, n_hits = 0, n_shaded = 0, n_undecided = 0;
    const uint32_t n_rays = SRC == 0 ? a.n_rays : __ldg(&a.b.counts->n_ref[a.ref_in]);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_rays; base += stride) {
        const uint32_t r = base + lane_id();
        bool active = r < n_rays;
        d3 s = mk(0, 0, 0), dir = mk(0, 0, 0); uint32_t sample = 0, depth = 0;
        uint32_t meta = kStateInvalid;
        if (active) {
            meta = a.b.meta[r];
            active = batch_ray<SRC>(a, insts, r, &s, &dir, &sample, &depth);
            if (SRC == 0 && !active) a.b.sample_state[r] = 0u;                    // outside the image / band (r < n_rays here)
        }
        const uint32_t state = meta & 3u;
        ExactHit eh; eh.hit = false; eh.inst = 0; eh.k = -1; eh.index = 0x7fffffff; eh.rf = kNoHit; eh.rf_from_clip = 0.0;
        eh.clipped_start = s; eh.dir = dir;
        bool undecided = false;
        if (active) {
            if (SRC == 0) n_primary++;
            if (state == kStateUndecided) { undecided = true; xc.filter_unsure++; n_undecided++; }
            else if (state == kStateListed) {
                const int4 cv = a.b.cand[r];
                const int ck[kMaxCand] = {cv.x, cv.y, cv.z, cv.w};
                if (f.n_instances == 1) {
                    int list[kMaxCand]; int n_list = 0;
#pragma unroll
                    for (int j = 0; j < kMaxCand; j++) if (ck[j] >= 0) list[n_list++] = ck[j];
                    if (n_list > 1) xc.filter_unsure++;
                    BestPrim bt; bt.rf = kNoHit; bt.k = -1; bt.index = 0x7fffffff;
                    d3 ts; double offset;
                    if (n_list > 0)     // (always: lets the compiler drop the evaluator's full-walk branch from this kernel)
                        mesh_closest_exact(a.sc.meshes[insts[0].mesh], f.subdivision, s, dir, list, n_list, &bt, &ts, &offset, &xc);
                    if (bt.k >= 0) {
                        eh.hit = true; eh.k = bt.k; eh.index = bt.index; eh.rf_from_clip = bt.rf; eh.rf = dadd(bt.rf, offset);
                        eh.clipped_start = ts;
                    }
                } else {
                    int n_list = 0;
#pragma unroll
                    for (int j = 0; j < kMaxCand; j++) {
                        if (ck[j] < 0) continue;
                        n_list++;
                        const int i = (int)((meta >> (4 + 7 * j)) & 127u);
                        const DevInstance& in = insts[i];
                        const d3 si = mk(in.start[0], in.start[1], in.start[2]);
                        const d3 di = mul3x3(in.Minv, dir);
                        BestPrim bt; bt.rf = kNoHit; bt.k = -1; bt.index = 0x7fffffff;
                        d3 ts; double offset;
                        const int one = ck[j];
                        mesh_closest_exact(a.sc.meshes[in.mesh], f.subdivision, si, di, &one, 1, &bt, &ts, &offset, &xc);
                        if (bt.k < 0) continue;
                        const double rf = dadd(bt.rf, offset);                    // SpatialSubdivision.cs:416
                        // nearest over instances, ties to the lowest instance, then the lowest triangle index
                        if (rf < eh.rf || (rf == eh.rf && (i < eh.inst || (i == eh.inst && bt.index < eh.index)))) {
                            eh.hit = true; eh.inst = i; eh.k = bt.k; eh.index = bt.index; eh.rf_from_clip = bt.rf; eh.rf = rf;
                            eh.clipped_start = ts; eh.dir = di;
                        }
                    }
                    if (n_list > 1) xc.filter_unsure++;
                }
            }
        }
        // the rays the search could not bracket go to the fallback kernel, compacted
        {
            const uint32_t at = warp_append(&a.b.counts->n_fallback[a.fb_slot], undecided);
            if (undecided) a.b.fallback[at] = r;
        }
        spawn_from_hit<SRC>(a, insts, active && !undecided, eh.hit, eh, sample, depth, &n_shaded, &n_shadow, &n_secondary, &n_hits);
    }
    const unsigned long long v[14] = {n_primary, n_shadow, n_secondary, xc.node_visits, xc.prim_tests, 0, n_hits, n_shaded, 0, xc.filter_unsure, 0, 0, n_undecided, 0};
    flush_counters(a.counters, v);
}

// The rays of the fallback list through the full reference-arithmetic search: SpatialSubdivision.IntersectRay's
// root-box clip (reference_clip), then every triangle whose (FP32, padded) leaf box the ray touches through
// Triangle.IntersectRay, nearest rayFrac wins, ties to the lowest triangle index / lowest instance -- the result of
// mesh_closest_exact's full walk.  These rays are few (thousands per frame) but their walks are long (a ray through
// the pole of a lat-long sphere meets hundreds of sliver triangles), so ONE WARP takes one ray: the walk is uniform
// over the warp (every node fetch is a broadcast), leaf triangles are collected 32 at a time and tested by the 32
// lanes side by side, a lexicographic warp reduction keeps the winner.  (Per-lane rays left the kernel at the mercy
// of its slowest thread: 15 ms per config4 frame for 1 178 rays.)
struct WarpBest { double rf; int index; int k; };

__device__ __forceinline__ void fallback_flush(const TriRec* __restrict__ tris, const int* list, int n, d3 ts, d3 dir, WarpBest& best,
                                               unsigned int* n_tests)
{
    const int lane = (int)lane_id();
    double rf = kNoHit; int index = 0x7fffffff, k = -1;
    __syncwarp();
    if (lane < n) {
        const int kk = list[lane];
        double r;
        (*n_tests)++;
        if (tri_intersect(tris + kk, ts, dir, best.rf, &r)) { rf = r; k = kk; index = __ldg(reinterpret_cast<const int*>(tris + kk) + 31); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double orf = __shfl_xor_sync(0xffffffffu, rf, o);
        const int oidx = __shfl_xor_sync(0xffffffffu, index, o);
        const int ok = __shfl_xor_sync(0xffffffffu, k, o);
        if (orf < rf || (orf == rf && oidx < index)) { rf = orf; index = oidx; k = ok; }
    }
    if (k >= 0 && (rf < best.rf || (rf == best.rf && index < best.index))) { best.rf = rf; best.index = index; best.k = k; }
    __syncwarp();
}

template <int SRC>
__global__ void __launch_bounds__(kWaveThreads) k_fallback(const __grid_constant__ WaveArgs a)
{
    const DevFrame& f = a.f;
    const DevInstance* __restrict__ insts = a.insts;
    __shared__ int s_list[kWaveThreads / 32][32];
    __shared__ int s_inst[kWaveThreads / 32][SOFTRAY_MAX_INSTANCES];
    int* list = s_list[threadIdx.x >> 5];
    int walk_stack[kStackEntries];
    unsigned int n_nodes = 0, n_tests = 0, n_shadow = 0, n_secondary = 0, n_hits = 0, n_shaded = 0;
    const uint32_t n = __ldg(&a.b.counts->n_fallback[a.fb_slot]);
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < n; q += n_warps) {
        d3 s, dir; uint32_t sample, depth;
        const uint32_t r = a.b.fallback[q];
        batch_ray<SRC>(a, insts, r, &s, &dir, &sample, &depth);
        ExactHit eh; eh.hit = false; eh.inst = 0; eh.k = -1; eh.index = 0x7fffffff; eh.rf = kNoHit; eh.rf_from_clip = 0.0;
        eh.clipped_start = s; eh.dir = dir;
        // composite frames: only the instances whose view-space box the ray touches (the frame's instance hierarchy,
        // walked without a cull distance: its boxes are padded like every other box) -- not all of them
        int n_inst = f.n_instances;
        bool listed = false;
        if (f.n_instances > 1 && f.tlas_nodes != nullptr) {
            FRay tr;
            tr.ox = tr.oy = tr.oz = 0.0f;
            tr.gx = __double2float_rn(dir.x); tr.gy = __double2float_rn(dir.y); tr.gz = __double2float_rn(dir.z);
            tr.ix = __fdiv_rn(1.0f, tr.gx); tr.iy = __fdiv_rn(1.0f, tr.gy); tr.iz = __fdiv_rn(1.0f, tr.gz);
            tr.nox = tr.noy = tr.noz = 0.0f;
            tr.tcull = CUDART_INF_F;
            int* ilist = s_inst[threadIdx.x >> 5];
            int tstack[kTlasStackEntries];
            int sp = 0, cur = 0, nl = 0;
            for (;;) {
                if (cur >= 0) {
                    const float4* p = reinterpret_cast<const float4*>(f.tlas_nodes + cur);
                    const float4 A = __ldg(p), B = __ldg(p + 1), CC = __ldg(p + 2);
                    const int2 d = __ldg(reinterpret_cast<const int2*>(p + 3));
                    float t0, t1;
                    const bool h0 = fslab(tr, A.x, A.y, A.z, A.w, B.x, B.y, &t0);
                    const bool h1 = fslab(tr, B.z, B.w, CC.x, CC.y, CC.z, CC.w, &t1);
                    if (h0 && h1) { tstack[sp++] = d.y; cur = d.x; continue; }
                    if (h0) { cur = d.x; continue; }
                    if (h1) { cur = d.y; continue; }
                } else {
                    const int code = -1 - cur;
                    const int first = code >> 4, count = code & 15;
                    for (int j = 0; j < count; j++) { if (nl < SOFTRAY_MAX_INSTANCES) ilist[nl] = __ldg(f.tlas_order + first + j); nl++; }
                }
                if (sp == 0) break;
                cur = tstack[--sp];
            }
            __syncwarp();
            if (nl <= SOFTRAY_MAX_INSTANCES) { n_inst = nl; listed = true; }
        }
        for (int ii = 0; ii < n_inst; ii++) {
            const int i = listed ? s_inst[threadIdx.x >> 5][ii] : ii;
            const DevInstance& in = insts[i];
            const DevMesh& m = a.sc.meshes[in.mesh];
            if (m.n_tris <= 0) continue;
            const d3 si = f.n_instances == 1 ? s : mk(in.start[0], in.start[1], in.start[2]);
            const d3 di = f.n_instances == 1 ? dir : mul3x3(in.Minv, dir);
            d3 ts = si; double offset = 0.0;
            if (f.subdivision && !reference_clip(m.bmin, m.bmax, &ts, di, &offset)) continue;      // SpatialSubdivision.cs:389-398
            double te = 0.0;
            if (!f.subdivision && !entry_clip(m.bmin, m.bmax, 0.0, ts, di, &te)) continue;
            const TravRay tr = make_trav(ts, di, te);
            // rayFrac = rf + offset must beat the best of the earlier instances (strictly: the lowest instance keeps a tie)
            WarpBest best; best.rf = eh.hit ? eh.rf - offset : kNoHit; best.index = 0x7fffffff; best.k = -1;
            if (eh.hit) best.rf = best.rf * (1.0 + 1e-15) + 1e-300;                                 // (only a cull: the comparison below decides)
            float tcull = cull_from(best.rf, tr.t_off);
            int n_list = 0;
            const unsigned int nv = walk_bvh(
                m.nodes, m.n_tris, walk_stack,
                [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float* t) {
                    return slab(tr, lox, loy, loz, hix, hiy, hiz, tcull, t);
                },
                [&](int first, int count) {
                    for (int j = 0; j < count; j++) {
                        list[n_list++] = first + j;                                                 // (every lane stores the same value)
                        if (n_list == 32) {
                            fallback_flush(m.tris, list, n_list, ts, di, best, &n_tests);
                            n_list = 0;
                            tcull = cull_from(best.rf, tr.t_off);
                        }
                    }
                    return false;
                });
            if (n_list > 0) fallback_flush(m.tris, list, n_list, ts, di, best, &n_tests);
            if (lane_id() == 0) n_nodes += nv;
            if (best.k < 0) continue;
            const double rf = dadd(best.rf, offset);                                                // SpatialSubdivision.cs:416
            if (rf < eh.rf || (rf == eh.rf && i < eh.inst)) {                                        // ties: the lowest instance
                eh.hit = true; eh.inst = i; eh.k = best.k; eh.index = best.index; eh.rf_from_clip = best.rf; eh.rf = rf;
                eh.clipped_start = ts; eh.dir = di;
            }
        }
        spawn_from_hit<SRC>(a, insts, lane_id() == 0, eh.hit, eh, sample, depth, &n_shaded, &n_shadow, &n_secondary, &n_hits);
    }
    const unsigned long long v[14] = {0, n_shadow, n_secondary, n_nodes, n_tests, 0, n_hits, n_shaded, 0, 0, 0, 0, 0, 0};
    flush_counters(a.counters, v);
}

// ---------------------------------------------------------------------------------------------
// shadow: ShadowMethod.TraceRaysForSoftShadows (ShadowMethod.cs:144-180) over the compacted shading points
// ---------------------------------------------------------------------------------------------
// One ShadowMethod ray (sample i of shading point `end`): the filtered any-hit answer.  With suspects (n_sus > 0) the
// ray tests those triangles only -- walk_filter_any's leaf test; every other triangle is proven missed by the cone
// walk -- otherwise it walks.  Returns 0 escaped, 1 occluded, 2 cannot tell (then *fb names what the reference
// arithmetic has to look at).
__device__ __forceinline__ int shadow_sample(const DevFrame& f, const DevInstance& in, const DevMesh& m, const double* offsets, d3 end, int i,
                                             const int* suspects, int n_sus, ShadowFallback* fb, Counters* c)
{
    d3 start, dir;
    shadow_ray(f, in, offsets, end, i, &start, &dir);
    const d3 anchor = f.point_lighting ? end : vadd(start, dir);   // the ray's far end
    FRay r;
    int list[kMaxCand] = {-1, -1, -1, -1}; int n_list = 0;
    int res = fray_setup(m, f.subdivision, anchor, dir, &r);
    if (res == 1) {
        if (n_sus > 0) {
            bool hit = false;
            unsigned int nf = 0;
#pragma unroll
            for (int j = 0; j < kMaxSuspects; j++) {
                if (j < n_sus && !hit) {
                    float tau, etau;
                    nf++;
                    const int t = tri_filter<false>(m.filt + suspects[j], r, m.scale, r.tmax_hi, &tau, &etau);
                    if (t == 1) hit = true;
                    else if (t == 2) { if (n_list < kMaxCand) list[n_list] = suspects[j]; n_list++; }
                }
            }
            c->filter_tests += nf;
            res = hit ? 1 : (n_list ? 2 : 0);
        } else {
            res = walk_filter_any(m.nodes, m.filt, m.n_tris, r, m.scale, list, &n_list, c);
        }
    }
    if (res == 2) {
        c->filter_unsure++;
        const bool listed = n_list >= 1 && n_list <= kMaxCand;
        fb->sample = (uint32_t)i;
        fb->list[0] = listed ? list[0] : -2; fb->list[1] = listed && n_list > 1 ? list[1] : -1;
        fb->list[2] = listed && n_list > 2 ? list[2] : -1; fb->list[3] = listed && n_list > 3 ? list[3] : -1;
    }
    return res;
}

// a ray the filter could not decide: onto the fallback list, or (list full) through the reference arithmetic here
__device__ __forceinline__ void shadow_unsure(const WaveArgs& a, const DevInstance& in, const DevMesh& m, const double* offsets, d3 end,
                                              bool unsure, const ShadowFallback& fb, int* escaped, Counters* c)
{
    if (!__any_sync(0xffffffffu, unsure)) return;
    const uint32_t at = warp_append(&a.b.counts->n_shadow_fallback, unsure);
    if (!unsure) return;
    if (at < a.cap_shadow_fallback) { a.b.shadow_fallback[at] = fb; return; }
    d3 start, dir;
    shadow_ray(a.f, in, offsets, end, (int)fb.sample, &start, &dir);
    int list[kMaxCand]; int n_list = 0;
    for (int j = 0; j < kMaxCand; j++) if (fb.list[j] >= 0) list[n_list++] = fb.list[j];
    XCounters xc; xc.stack = c->stack; xc.node_visits = 0; xc.prim_tests = 0; xc.sphere_tests = 0; xc.filter_tests = 0;
    xc.filter_unsure = 0; xc.filter_mismatch = 0;
    if (!occluded_mesh(m, a.f.subdivision, start, dir, n_list ? list : nullptr, n_list, &xc)) (*escaped)++;
    c->node_visits += xc.node_visits; c->prim_tests += xc.prim_tests;
}

__device__ __forceinline__ void load_shadow_item(const WaveArgs& a, uint32_t q, d3* end, uint32_t* slot, uint32_t* inst)
{
    const ShadowItem* it = a.b.shadow + q;
    const double2 v0 = ldg2(it, 0);
    const double v1 = __ldg(reinterpret_cast<const double*>(it) + 2);
    const uint2 si = __ldg(reinterpret_cast<const uint2*>(it) + 3);
    *end = mk(v0.x, v0.y, v1); *slot = si.x; *inst = si.y;
}

// shadow, first pass: one conservative walk of the cone of each shading point's rays (bundle_suspects).  0 suspects:
// every ray escapes.  1..8: the rays test those triangles, here.  Too many / out of budget: the point goes onto the
// walk list, whose rays k_shadow_walk traces in small independent pieces.  (One kernel doing both left the frame
// waiting for single warps that walked 100 long rays one after the other: 10 ms of tail on config3 at any GPU count.)
template <int MINB>
__global__ void __launch_bounds__(kWaveThreads, MINB) k_shadow(const __grid_constant__ WaveArgs a)
{
    const DevFrame& f = a.f;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_offsets = reinterpret_cast<double*>(smem_raw);
    for (int i = threadIdx.x; i < 3 * f.shadow_samples; i += blockDim.x) s_offsets[i] = a.offsets[i];
    __syncthreads();
    const DevInstance* __restrict__ insts = a.insts;
    int walk_stack[kStackEntries];
    Counters c; c.node_visits = 0; c.prim_tests = 0; c.sphere_tests = 0; c.shaded = 0; c.filter_tests = 0; c.filter_unsure = 0;
    c.filter_mismatch = 0; c.bundled = 0; c.bundle_skip = 0; c.stack = walk_stack;
    const uint32_t n_items = __ldg(&a.b.counts->n_shadow);
    const int n = f.shadow_samples;
    unsigned int n_listed = 0;
    // a warp takes 32 consecutive shading points (neighbouring pixels: their cones run side by side); warps fetch
    // their groups from a queue
    for (;;) {
        uint32_t group = 0;
        if (lane_id() == 0) group = atomicAdd(&a.b.counts->shadow_head, 1u);
        group = __shfl_sync(0xffffffffu, group, 0);
        if ((unsigned long long)group * 32ull >= n_items) break;
        const uint32_t q = group * 32u + lane_id();
        const bool active = q < n_items;
        d3 end = mk(0, 0, 0); uint32_t slot = 0, inst = 0;
        if (active) load_shadow_item(a, q, &end, &slot, &inst);
        const DevInstance& in = insts[inst];
        const DevMesh& m = a.sc.meshes[in.mesh];
        int escaped = 0;
        int suspects[kMaxSuspects];
        int n_sus = -1;
        if (active) {
            const d3 light = mk(in.light_pos_model[0], in.light_pos_model[1], in.light_pos_model[2]);
            n_sus = bundle_suspects(m, end, light, f.light_radius, f.bundle_budget, suspects, &c);
            if (n_sus == 0) { c.bundled += (unsigned int)n; escaped = n; }
            if (n_sus > 0) n_listed += (unsigned int)n;
        }
        if (__any_sync(0xffffffffu, n_sus > 0)) {
            for (int i = 0; i < n; i++) {
                bool unsure = false;
                ShadowFallback fb; fb.item = q;
                if (n_sus > 0) {
                    const int res = shadow_sample(f, in, m, s_offsets, end, i, suspects, n_sus, &fb, &c);
                    unsure = res == 2;
                    if (res == 0) escaped++;
                }
                shadow_unsure(a, in, m, s_offsets, end, unsure, fb, &escaped, &c);
            }
        }
        if (active) a.b.slot_escaped[slot] = (uint32_t)escaped;
        {
            const bool walk = active && n_sus < 0;
            const uint32_t at = warp_append(&a.b.counts->n_walk, walk);
            if (walk) a.b.walk_list[at] = q;
        }
    }
    const unsigned long long v[14] = {0, 0, 0, c.node_visits, c.prim_tests, 0, 0, 0, c.filter_tests, c.filter_unsure, 0, c.bundled, 0, n_listed};
    flush_counters(a.counters, v);
}

// shadow, second pass: the rays of the shading points on the walk list (all of them when the frame has no bundles:
// fewer than 16 samples per point), kWalkPiece samples of one point per thread.  Consecutive lanes hold consecutive
// points (neighbouring pixels) and the same samples, so their walks run side by side; the pieces of one point add
// their escaped rays with one atomic each.
constexpr int kWalkPiece = 2;        // (8: config3 on 8 GPUs waited 2.6 ms for the longest pieces)

template <int MINB>
__global__ void __launch_bounds__(kWaveThreads, MINB) k_shadow_walk(const __grid_constant__ WaveArgs a)
{
    const DevFrame& f = a.f;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_offsets = reinterpret_cast<double*>(smem_raw);
    for (int i = threadIdx.x; i < 3 * f.shadow_samples; i += blockDim.x) s_offsets[i] = a.offsets[i];
    __syncthreads();
    const DevInstance* __restrict__ insts = a.insts;
    int walk_stack[kStackEntries];
    Counters c; c.node_visits = 0; c.prim_tests = 0; c.sphere_tests = 0; c.shaded = 0; c.filter_tests = 0; c.filter_unsure = 0;
    c.filter_mismatch = 0; c.bundled = 0; c.bundle_skip = 0; c.stack = walk_stack;
    const bool all_points = f.bundle_budget <= 0 || !f.point_lighting;
    const uint32_t n_points = all_points ? __ldg(&a.b.counts->n_shadow) : __ldg(&a.b.counts->n_walk);
    const int n = f.shadow_samples;
    const uint32_t pieces = (uint32_t)((n + kWalkPiece - 1) / kWalkPiece);
    const uint32_t groups_per_piece = (n_points + 31u) / 32u;
    const unsigned long long n_groups = (unsigned long long)groups_per_piece * pieces;
    for (;;) {
        uint32_t group = 0;
        if (lane_id() == 0) group = atomicAdd(&a.b.counts->walk_head, 1u);
        group = __shfl_sync(0xffffffffu, group, 0);
        if (group >= n_groups) break;
        const uint32_t piece = group / groups_per_piece, pg = group - piece * groups_per_piece;
        const uint32_t w = pg * 32u + lane_id();
        const bool active = w < n_points;
        uint32_t q = 0;
        if (active) q = all_points ? w : a.b.walk_list[w];
        d3 end = mk(0, 0, 0); uint32_t slot = 0, inst = 0;
        if (active) load_shadow_item(a, q, &end, &slot, &inst);
        const DevInstance& in = insts[inst];
        const DevMesh& m = a.sc.meshes[in.mesh];
        int escaped = 0;
        const int i0 = (int)piece * kWalkPiece, i1 = min(n, i0 + kWalkPiece);
        for (int i = i0; i < i1; i++) {
            bool unsure = false;
            ShadowFallback fb; fb.item = q;
            if (active) {
                const int res = shadow_sample(f, in, m, s_offsets, end, i, nullptr, 0, &fb, &c);
                unsure = res == 2;
                if (res == 0) escaped++;
            }
            shadow_unsure(a, in, m, s_offsets, end, unsure, fb, &escaped, &c);
        }
        if (active && escaped) atomicAdd(&a.b.slot_escaped[slot], (uint32_t)escaped);
    }
    const unsigned long long v[14] = {0, 0, 0, c.node_visits, c.prim_tests, 0, 0, 0, c.filter_tests, c.filter_unsure, 0, 0, 0, 0};
    flush_counters(a.counters, v);
}

__global__ void __launch_bounds__(kWaveThreads) k_shadow_fallback(const __grid_constant__ WaveArgs a)
{
    const DevFrame& f = a.f;
    const DevInstance* __restrict__ insts = a.insts;
    int walk_stack[kStackEntries];
    XCounters xc; xc.stack = walk_stack; xc.node_visits = 0; xc.prim_tests = 0; xc.sphere_tests = 0; xc.filter_tests = 0;
    xc.filter_unsure = 0; xc.filter_mismatch = 0;
    const uint32_t n = min(__ldg(&a.b.counts->n_shadow_fallback), a.cap_shadow_fallback);
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const ShadowFallback fb = a.b.shadow_fallback[q];
        const ShadowItem it = a.b.shadow[fb.item];
        const DevInstance& in = insts[it.inst];
        const DevMesh& m = a.sc.meshes[in.mesh];
        d3 start, dir;
        shadow_ray(f, in, a.offsets, mk(it.end[0], it.end[1], it.end[2]), (int)fb.sample, &start, &dir);
        int list[kMaxCand]; int n_list = 0;
        for (int j = 0; j < kMaxCand; j++) if (fb.list[j] >= 0) list[n_list++] = fb.list[j];
        const bool occ = occluded_mesh(m, f.subdivision, start, dir, n_list ? list : nullptr, n_list, &xc);
        if (!occ) atomicAdd(&a.b.slot_escaped[it.slot], 1u);
    }
    const unsigned long long v[14] = {0, 0, 0, xc.node_visits, xc.prim_tests, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    flush_counters(a.counters, v);
}

// ---------------------------------------------------------------------------------------------
// compose: shadow byte, mirror blend chain, sub-pixel sums, Surface.DrawPixel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWaveThreads) k_compose(const __grid_constant__ WaveArgs a, uint32_t* __restrict__ pixels,
                                                         int32_t* __restrict__ hit_ids)
{
    const DevFrame& f = a.f;
    const int n = f.sub_pixel_res, nn = n * n;
    const uint32_t S = a.n_samples;
    const uint32_t n_px = a.n_tiles * 32u;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_px) return;
    const int t = (int)(p >> 5), px = (int)(p & 31u);
    const int tile = a.tile0 + t;
    const int ty = (int)f.fd_tiles_x.div((uint32_t)tile), tx = tile - ty * f.tiles_x;
    const int band_j = (int)f.fd_tiles_per_band.div((uint32_t)ty);
    const int row0 = f.start_row + (f.band_index + band_j * f.band_count) * f.band_height;
    const int band_r = (ty - band_j * f.tiles_per_band) * 4 + (px >> 3);
    const int col = tx * 8 + (px & 7), row = row0 + band_r;
    if (col >= f.width || band_r >= f.band_height || row > f.end_row) return;
    int sum_r = 0, sum_g = 0, sum_b = 0, id = -1;
    for (int si = 0; si < nn; si++) {
        const uint32_t s = (uint32_t)t * 32u * (uint32_t)nn + (uint32_t)(px * nn + si);
        const uint32_t st = a.b.sample_state[s];
        const int n_hits = (int)(st & 7u);
        uint32_t acc = f.background;
        if (n_hits > 0) {
            uint32_t local[5];
            for (int k = 0; k < n_hits; k++) {
                uint32_t color = a.b.slot_color[(uint32_t)k * S + s];
                if (f.shadows) {
                    const uint32_t escaped = a.b.slot_escaped[(uint32_t)k * S + s];
                    const double frac = ddiv((double)escaped, (double)f.shadow_samples);   // ShadowMethod.cs:178
                    color = modulate(color, to_byte(dmul(frac, 255.0)));
                }
                local[k] = color;
            }
            acc = (st & 8u) ? mirror_blend(local[n_hits - 1], f.background) : local[n_hits - 1];
            for (int k = n_hits - 2; k >= 0; k--) acc = mirror_blend(local[k], acc);
            if (si == nn - 1) id = a.b.sample_id[s];
        }
        sum_r += (int)((acc >> 16) & 0xff); sum_g += (int)((acc >> 8) & 0xff); sum_b += (int)(acc & 0xff);
    }
    sum_r /= nn; sum_g /= nn; sum_b /= nn;                                       // Renderer.cs:1820-1822
    const size_t idx = (size_t)row * (size_t)f.width + (size_t)col;
    pixels[idx] = 0xff000000u | ((uint32_t)(sum_r & 0xff) << 16) | ((uint32_t)(sum_g & 0xff) << 8) | (uint32_t)(sum_b & 0xff);
    if (hit_ids) hit_ids[idx] = id;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side: buffers and the launch sequence of one frame
// ---------------------------------------------------------------------------------------------
size_t wave_buffer_bytes(uint32_t cap_samples, int max_depth_slots, WaveLayout* lay)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
    const size_t S = cap_samples;
    lay->counts = take(sizeof(WaveCounts));
    lay->cand = take(S * sizeof(int4));
    lay->meta = take(S * sizeof(uint32_t));
    lay->slot_color = take(S * (size_t)max_depth_slots * sizeof(uint32_t));
    lay->slot_escaped = take(S * (size_t)max_depth_slots * sizeof(uint32_t));
    lay->sample_state = take(S);
    lay->sample_id = take(S * sizeof(int32_t));
    const size_t n_ref = max_depth_slots > 1 ? S : 1;         // (no reflection: no reflection-ray lists)
    lay->ref0 = take(n_ref * sizeof(RefRay));
    lay->ref1 = take(n_ref * sizeof(RefRay));
    lay->shadow = take(S * (size_t)max_depth_slots * sizeof(ShadowItem));
    lay->fallback = take(S * sizeof(uint32_t));
    lay->walk_list = take(S * (size_t)max_depth_slots * sizeof(uint32_t));
    lay->shadow_fallback = take(S * sizeof(ShadowFallback));      // (capacity checked by the host against the count: see wave_render)
    return off;
}

void wave_bind(void* base, const WaveLayout& lay, WaveBufs* b)
{
    unsigned char* p = static_cast<unsigned char*>(base);
    b->counts = reinterpret_cast<WaveCounts*>(p + lay.counts);
    b->cand = reinterpret_cast<int4*>(p + lay.cand);
    b->meta = reinterpret_cast<uint32_t*>(p + lay.meta);
    b->slot_color = reinterpret_cast<uint32_t*>(p + lay.slot_color);
    b->slot_escaped = reinterpret_cast<uint32_t*>(p + lay.slot_escaped);
    b->sample_state = reinterpret_cast<uint8_t*>(p + lay.sample_state);
    b->sample_id = reinterpret_cast<int32_t*>(p + lay.sample_id);
    b->ref[0] = reinterpret_cast<RefRay*>(p + lay.ref0);
    b->ref[1] = reinterpret_cast<RefRay*>(p + lay.ref1);
    b->shadow = reinterpret_cast<ShadowItem*>(p + lay.shadow);
    b->fallback = reinterpret_cast<uint32_t*>(p + lay.fallback);
    b->walk_list = reinterpret_cast<uint32_t*>(p + lay.walk_list);
    b->shadow_fallback = reinterpret_cast<ShadowFallback*>(p + lay.shadow_fallback);
}

int wave_max_depth_slots() { return 5; }

namespace {
int env_occ(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    const int v = (e && *e) ? std::atoi(e) : dflt;
    return v < 1 ? 1 : (v > 6 ? 6 : v);
}
template <int SRC>
void launch_search(int occ, int grid, cudaStream_t st, const WaveArgs& a)
{
    switch (occ) {
    case 2: k_search<SRC, 2><<<grid, kWaveThreads, 0, st>>>(a); break;
    case 3: k_search<SRC, 3><<<grid, kWaveThreads, 0, st>>>(a); break;
    case 4: k_search<SRC, 4><<<grid, kWaveThreads, 0, st>>>(a); break;
    case 5: k_search<SRC, 5><<<grid, kWaveThreads, 0, st>>>(a); break;
    default: k_search<SRC, 6><<<grid, kWaveThreads, 0, st>>>(a); break;
    }
}
template <int SRC>
void launch_hit(int occ, int grid, cudaStream_t st, const WaveArgs& a)
{
    switch (occ) {
    case 2: k_hit<SRC, 2><<<grid, kWaveThreads, 0, st>>>(a); break;
    case 3: k_hit<SRC, 3><<<grid, kWaveThreads, 0, st>>>(a); break;
    default: k_hit<SRC, 4><<<grid, kWaveThreads, 0, st>>>(a); break;
    }
}
cudaError_t launch_shadow_walk(int occ, int grid, size_t smem, cudaStream_t st, const WaveArgs& a)
{
#define SR_WALK_CASE(N)                                                                                                     \
    case N:                                                                                                                 \
        if (smem > 48 * 1024) {                                                                                             \
            cudaError_t e = cudaFuncSetAttribute(k_shadow_walk<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                                 \
        }                                                                                                                   \
        k_shadow_walk<N><<<grid, kWaveThreads, smem, st>>>(a);                                                              \
        break;
    switch (occ) {
        SR_WALK_CASE(3) SR_WALK_CASE(4)
    default:
        SR_WALK_CASE(5)
    }
#undef SR_WALK_CASE
    return cudaSuccess;
}

cudaError_t launch_shadow(int occ, int grid, size_t smem, cudaStream_t st, const WaveArgs& a)
{
#define SR_SHADOW_CASE(N)                                                                                                   \
    case N:                                                                                                                 \
        if (smem > 48 * 1024) {                                                                                             \
            cudaError_t e = cudaFuncSetAttribute(k_shadow<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
            if (e != cudaSuccess) return e;                                                                                 \
        }                                                                                                                   \
        k_shadow<N><<<grid, kWaveThreads, smem, st>>>(a);                                                                   \
        break;
    switch (occ) {
        SR_SHADOW_CASE(2) SR_SHADOW_CASE(3) SR_SHADOW_CASE(4) SR_SHADOW_CASE(5)
    default:
        SR_SHADOW_CASE(6)
    }
#undef SR_SHADOW_CASE
    return cudaSuccess;
}
}  // namespace

// softray_frame.profile_stages: an event after every stage kernel; wave_stage_times() turns them into ms per stage
// once the stream has been synchronised.
void StageTimer::mark(int which)
{
    if (!on) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
    ev.push_back(e); stage.push_back(which);
}
void StageTimer::collect(double* ms_stage, int n)
{
    for (int k = 0; k < n; k++) ms_stage[k] = 0.0;
    for (size_t i = 1; i < ev.size(); i++) {
        float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
        if (stage[i] >= 0 && stage[i] < n) ms_stage[stage[i]] += ms;
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear(); stage.clear();
}

// One frame: every chunk of tiles through the stage kernels.  Everything is stream-ordered; nothing synchronises
// the host.  Consecutive chunks alternate between two streams (each with its own set of buffers): while one chunk
// sits in a stage with little parallelism (the fallback kernels: a few thousand rays with long serial walks; the
// tail of any kernel), the other chunk's kernels fill the machine.
cudaError_t wave_render(const DevFrame& f, const DevScene& sc, const DevInstance* d_insts, const double* d_offsets, const WaveBufs bufs[2],
                        uint32_t cap_samples, uint32_t* d_pixels, int32_t* d_ids, DevCounters* d_counters, int sm_count,
                        cudaStream_t stream, cudaStream_t side_stream, cudaEvent_t ev_fork, cudaEvent_t ev_join, int* launches,
                        StageTimer* prof, const HostCopy* host)
{
    const int nn = f.sub_pixel_res * f.sub_pixel_res;
    const uint32_t per_tile = 32u * (uint32_t)nn;
    const long long n_tiles = (long long)f.tiles_x * f.tiles_y;
    if (cap_samples / per_tile == 0) return cudaErrorInvalidValue;
    long long n_chunks = (n_tiles * per_tile + cap_samples - 1) / cap_samples;
    const bool profiling = prof != nullptr && prof->on;
    const bool two = side_stream != nullptr && !profiling && env_occ("SOFTRAY_WAVE_STREAMS", 2) >= 2 && n_tiles * per_tile >= (1u << 18);
    if (two && n_chunks < 2) n_chunks = 2;
    if (host) {     // the last chunk's copy is exposed: more, smaller chunks -- but not below 4 M samples (config3 as 8 chunks: 7.2 -> 8.5 ms)
        long long want = n_tiles * per_tile / (1ll << 22);
        if (want > 8) want = 8;
        if (n_chunks < want) n_chunks = want;
    }
    if (two && (n_chunks & 1)) n_chunks++;
    long long tiles_per_chunk = (n_tiles + n_chunks - 1) / n_chunks;
    if (host) {     // whole tile rows per chunk: a chunk is then a few runs of pixel rows
        tiles_per_chunk = (tiles_per_chunk / f.tiles_x) * f.tiles_x;
        if (tiles_per_chunk < f.tiles_x) tiles_per_chunk = f.tiles_x;
        if (tiles_per_chunk * per_tile > cap_samples) return cudaErrorInvalidValue;
    }
    size_t ev_used = 0;
    const int bounces = (f.reflection_depth > 0 && f.n_instances == 1) ? f.reflection_depth : 0;
    const size_t smem_shadow = sizeof(double) * 3 * (size_t)(f.shadows ? f.shadow_samples : 0);
    const int occ_search = env_occ("SOFTRAY_WAVE_SEARCH_OCC", 4), occ_hit = env_occ("SOFTRAY_WAVE_HIT_OCC", 3),
              occ_shadow = env_occ("SOFTRAY_WAVE_SHADOW_OCC", 4), occ_walk = env_occ("SOFTRAY_WAVE_WALK_OCC", 4);
    const int persistent = sm_count * 8;          // grid of the list kernels (grid-stride over a device-side count)
    int n_launch = 0;
    StageTimer none;
    StageTimer& tm = profiling ? *prof : none;
    tm.st = stream;
    cudaError_t e;
    if (two) {
        if ((e = cudaEventRecord(ev_fork, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(side_stream, ev_fork, 0)) != cudaSuccess) return e;
    }
    int ci = 0;
    for (long long tile0 = 0; tile0 < n_tiles; tile0 += tiles_per_chunk, ci++) {
        cudaStream_t st = (two && (ci & 1)) ? side_stream : stream;
        const WaveBufs& wb = bufs[(two && (ci & 1)) ? 1 : 0];
        WaveArgs a;
        a.f = f; a.sc = sc; a.insts = d_insts; a.offsets = d_offsets; a.b = wb; a.counters = d_counters;
        a.tile0 = (int)tile0;
        a.n_tiles = (uint32_t)((n_tiles - tile0) < tiles_per_chunk ? (n_tiles - tile0) : tiles_per_chunk);
        a.n_samples = a.n_tiles * per_tile;
        a.n_rays = a.n_samples;
        a.ref_in = 0; a.ref_out = 0; a.fb_slot = 0;
        a.cap_shadow_fallback = cap_samples;
        if ((e = cudaMemsetAsync(wb.counts, 0, sizeof(WaveCounts), st)) != cudaSuccess) return e;
        // (grid-stride loops: a thread takes several rays, the per-thread prologue / counter flush is paid once)
        const int grid_need = (int)((a.n_rays + kWaveThreads - 1) / kWaveThreads);
        // every thread the same number of rays: the walks are long, a thread with one ray more is the kernel's tail
        const int per_thread = (grid_need + sm_count * 32 - 1) / (sm_count * 32);
        const int grid_cam = (grid_need + per_thread - 1) / per_thread;
        tm.mark(-1);
        launch_search<0>(occ_search, grid_cam, st, a); tm.mark(0);
        launch_hit<0>(occ_hit, grid_cam, st, a); tm.mark(1);
        k_fallback<0><<<persistent, kWaveThreads, 0, st>>>(a); tm.mark(2);
        n_launch += 3;
        for (int depth = 1; depth <= bounces; depth++) {
            a.ref_in = (depth - 1) & 1; a.ref_out = depth & 1; a.fb_slot = depth;
            // (the list written two bounces ago is consumed: its counter restarts)
            if ((e = cudaMemsetAsync(&wb.counts->n_ref[a.ref_out], 0, sizeof(uint32_t), st)) != cudaSuccess) return e;
            launch_search<1>(occ_search, persistent, st, a); tm.mark(3);
            launch_hit<1>(occ_hit, persistent, st, a); tm.mark(4);
            k_fallback<1><<<persistent, kWaveThreads, 0, st>>>(a); tm.mark(5);
            n_launch += 3;
        }
        if (f.shadows) {
            // cone walks first (frames with bundles), then the rays of the points they could not settle
            if (f.bundle_budget > 0 && f.point_lighting)
                if ((e = launch_shadow(occ_shadow, sm_count * occ_shadow, smem_shadow, st, a)) != cudaSuccess) return e;
            if ((e = launch_shadow_walk(occ_walk, sm_count * occ_walk, smem_shadow, st, a)) != cudaSuccess) return e;
            n_launch += 1;
            tm.mark(6);
            k_shadow_fallback<<<persistent, kWaveThreads, 0, st>>>(a); tm.mark(7);
            n_launch += 2;
        }
        k_compose<<<(int)((a.n_tiles * 32u + kWaveThreads - 1) / kWaveThreads), kWaveThreads, 0, st>>>(a, d_pixels, d_ids);
        n_launch += 1;
        tm.mark(8);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (host) {
            // this chunk's rows -> the host surface by DMA, behind an event: the SMs go on with the next chunk
            if (ev_used >= host->events->size()) {
                cudaEvent_t ev;
                if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
                host->events->push_back(ev);
            }
            cudaEvent_t ev = (*host->events)[ev_used++];
            if ((e = cudaEventRecord(ev, st)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(host->copy_stream, ev, 0)) != cudaSuccess) return e;
            const long long ty0 = tile0 / f.tiles_x, ty1 = (tile0 + a.n_tiles) / f.tiles_x;
            long long run_first = -1, run_last = -1;                 // a run of consecutive pixel rows
            auto flush = [&]() -> cudaError_t {
                if (run_first < 0) return cudaSuccess;
                const size_t off = (size_t)run_first * (size_t)f.width, cnt = (size_t)(run_last - run_first + 1) * (size_t)f.width;
                cudaError_t ce = cudaMemcpyAsync(host->h_pixels + off, host->d_pixels + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, host->copy_stream);
                if (ce == cudaSuccess && host->h_ids)
                    ce = cudaMemcpyAsync(host->h_ids + off, host->d_ids + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, host->copy_stream);
                run_first = -1;
                return ce;
            };
            for (long long ty = ty0; ty < ty1; ty++) {
                const long long band_j = ty / f.tiles_per_band;
                const long long band_top = f.start_row + (f.band_index + band_j * f.band_count) * (long long)f.band_height;
                const long long r0 = band_top + (ty - band_j * f.tiles_per_band) * 4;
                long long r1 = r0 + 3;
                if (r1 > band_top + f.band_height - 1) r1 = band_top + f.band_height - 1;
                if (r1 > f.end_row) r1 = f.end_row;
                if (r1 < r0) continue;
                if (run_first >= 0 && r0 == run_last + 1) { run_last = r1; continue; }
                if ((e = flush()) != cudaSuccess) return e;
                run_first = r0; run_last = r1;
            }
            if ((e = flush()) != cudaSuccess) return e;
        }
    }
    if (host) {     // the frame's stream ends when the last copy has landed
        if (ev_used >= host->events->size()) {
            cudaEvent_t ev;
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
            host->events->push_back(ev);
        }
        cudaEvent_t ev = (*host->events)[ev_used++];
        if ((e = cudaEventRecord(ev, host->copy_stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream, ev, 0)) != cudaSuccess) return e;
    }
    if (two) {
        if ((e = cudaEventRecord(ev_join, side_stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream, ev_join, 0)) != cudaSuccess) return e;
    }
    if (launches) *launches = n_launch;
    return cudaSuccess;
}

}  // namespace sr
