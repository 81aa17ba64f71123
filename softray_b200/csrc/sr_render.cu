// sr_render.cu -- the fused raytrace kernel for sm_100a: ray generation, BVH traversal,
// ray/sphere + ray/triangle intersection, Lambert/Phong shading, soft shadow rays, bounded
// reflection, Texture3D modulation, sub-pixel accumulation and the packed-ARGB store, in one
// persistent launch.  Replaces RaytraceBlock -> TraceRayComplex -> IRayIntersectable.IntersectRay
// -> Surface.DrawPixel (Engine3D/Renderer.cs:1690-1925 and the Raytrace/*Method.cs decorators).
//
// Numerics.  Every decision the reference takes in FP64 (which primitive wins, rayFrac, position,
// normal, lighting, truncation to bytes) is taken here with the SAME operations in the SAME order,
// written with __dmul_rn/__dadd_rn/... so nvcc can never contract them into FMAs (the reference
// never fuses).  Only the search for candidates is approximate: BVH boxes are
// FP32, rounded outward and padded, so the slab test can only produce false positives.


#include "sr_device.cuh"

namespace sr {

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct PixelOut { uint32_t color; int32_t id; };

__device__ __forceinline__ bool occluded_exact(const DevScene& sc, const DevInstance& in, const DevMesh& m, int subdivision,
                                               d3 start, d3 dir, const int* list, int n_list, XCounters* c)
{
    if (occluded_mesh(m, subdivision, start, dir, list, n_list, c)) return true;
    return in.sph_can_shadow && occluded_spheres(sc, start, dir, nullptr, 0, c);
}

__device__ __forceinline__ uint32_t shade_and_shadow(const DevFrame& f, const DevScene& sc, const DevInstance& in,
                                                     const DevMesh& m, const double* __restrict__ offsets, const Hit& h,
                                                     Counters* c, XCounters* xc, unsigned int* n_shadow, unsigned int* bundle_score)
{
    uint32_t color = h.color;
    c->shaded++;
    if (f.texture3d_id) color = modulate(color, texture3d_sample(f.texture3d_id, h.pos));
    if (f.shading) color = shade(f, in, h.pos, h.normal, color);
    if (f.shadows) {
        // ShadowMethod.TraceRaysForSoftShadows (ShadowMethod.cs:144-180)
        const d3 end = vadd(h.pos, vscale(h.normal, 0.001));
        int escaped = 0;
        const int n = f.shadow_samples;
        const bool use_filter = f.filter_mode != 1 && m.nodes != nullptr && m.n_tris > 0;
        *n_shadow += (unsigned int)n;
        const d3 light = mk(in.light_pos_model[0], in.light_pos_model[1], in.light_pos_model[2]);
        bool all_clear = false;
        if (use_filter && f.point_lighting && !in.sph_can_shadow && f.bundle_budget > 0) {
            // Whether cone walks pay is a property of the frame (config2: every one succeeds; config3: none does), and
            // a thread shades too few points to find out by itself (40 per thread on a 4K frame: a per-thread
            // exponential back-off still tried at 14 % of config3's shading points, 45.5 vs 43.0 ms without).  The
            // block keeps the score: its first 64 attempts explore, then a point tries only while at least one attempt
            // in 32 succeeds (a success saves ~100 rays, a failure costs two or three), plus one probe per 256 points
            // of a thread so that a block can change its mind.
            const unsigned int ok = bundle_score[0], bad = bundle_score[1];
            c->bundle_skip++;                                      // (shading points of this thread)
            if (bad < 64u || ok * 32u >= bad || (c->bundle_skip & 255) == 0) {
                all_clear = bundle_clear(m, end, light, f.light_radius, f.bundle_budget, c);
                if (all_clear) c->bundled += (unsigned int)n;
                atomicAdd(&bundle_score[all_clear ? 0 : 1], 1u);
            }
        }
        if (all_clear && f.filter_mode != 2) escaped = n;
        else
        for (int i = 0; i < n; i++) {
            d3 start, dir;
            shadow_ray(f, in, offsets, end, i, &start, &dir);
            bool occ;
            if (use_filter) {
                // anchor = the ray's far end (start + dir): `end` itself for a point light
                const d3 anchor = f.point_lighting ? end : vadd(start, dir);
                FRay r;
                int list[kMaxCand]; int n_list = 0;
                int res = fray_setup(m, f.subdivision, anchor, dir, &r);
                if (res == 1) res = walk_filter_any(m.nodes, m.filt, m.n_tris, r, m.scale, list, &n_list, c);
                if (n_list > kMaxCand) n_list = 0;                  // too many undecided triangles: full exact walk
                if (f.filter_mode == 2) {
                    occ = occluded_mesh(m, f.subdivision, start, dir, nullptr, 0, xc);
                    if ((res == 0 && occ) || (res == 1 && !occ) || (all_clear && occ)) c->filter_mismatch++;
                    if (res == 2) {
                        c->filter_unsure++;
                        if (n_list > 0 && occluded_mesh(m, f.subdivision, start, dir, list, n_list, xc) != occ) c->filter_mismatch++;
                    }
                } else if (res == 2) {
                    // the exact arithmetic looks only at the triangles the filter could not decide
                    c->filter_unsure++;
                    occ = occluded_mesh(m, f.subdivision, start, dir, list, n_list, xc);
                } else {
                    occ = res == 1;
                }
                if (!occ && in.sph_can_shadow) {
                    int sres = 2;
                    int sl[kMaxCand]; int nsl = 0;
                    if (sc.sphere_nodes != nullptr) {
                        unsigned int nv = 0, nf = 0;
                        sres = spheres_filter<true>(sc, start, dir, sl, &nsl, &nv, &nf, c->stack);
                        c->node_visits += nv; c->filter_tests += nf;
                    }
                    if (nsl > kMaxCand) nsl = 0;
                    if (f.filter_mode == 2) {
                        occ = occluded_spheres(sc, start, dir, nullptr, 0, xc);
                        if ((sres == 0 && occ) || (sres == 1 && !occ)) c->filter_mismatch++;
                        if (sres == 2 && nsl > 0 && occluded_spheres(sc, start, dir, sl, nsl, xc) != occ) c->filter_mismatch++;
                    } else if (sres == 2) {
                        occ = occluded_spheres(sc, start, dir, sl, nsl, xc);
                    } else {
                        occ = sres == 1;
                    }
                }
            } else {
                occ = occluded_exact(sc, in, m, f.subdivision, start, dir, nullptr, 0, xc);
            }
            if (!occ) escaped++;
        }
        const double frac = ddiv((double)escaped, (double)n);
        color = modulate(color, to_byte(dmul(frac, 255.0)));
    }
    return color;
}

// TraceRayComplex (Renderer.cs:1850-1879) for one camera ray (+ mirror bounces, + composite instances)
template <bool STAGED>
__device__ __forceinline__ PixelOut trace_camera_ray(const DevFrame& f, const DevScene& sc, const DevInstance* __restrict__ insts,
                                                     const double* __restrict__ offsets, const d3* starts, const d3* dirs_view_or_world,
                                                     bool dirs_are_view, Counters* c, XCounters* xc, unsigned int* n_shadow,
                                                     unsigned int* n_secondary, bool* hit_out, int sync, bool valid, unsigned int* bundle_score,
                                                     const SphereStage* stage)
{
    PixelOut out; out.color = f.background; out.id = -1;
    Hit h; int which = 0; bool hit = false;
    d3 dir0 = mk(0, 0, 0);
    if (f.n_instances == 1) {
        const DevInstance& in = insts[0];
        dir0 = dirs_are_view ? mul3x3(in.Minv, dirs_view_or_world[0]) : dirs_view_or_world[0];
        hit = closest_hit<STAGED>(sc, sc.meshes[in.mesh], f.subdivision, f.filter_mode, starts[0], dir0, &h, xc, sync, 1.7976931348623157e308, stage);
    } else {
        // extension: nearest hit across instances, ties to the lowest instance.  Rigid
        // transforms keep |dir|, so every instance's rayFrac is the parameter along dirs_view_or_world[0]
        // from the view origin and the instance hierarchy can cull with it.
        double best = kNoHit;
        const d3 dv = dirs_view_or_world[0];
        if (f.tlas_nodes != nullptr) {
            FRay tr;
            tr.ox = tr.oy = tr.oz = 0.0f;
            tr.gx = __double2float_rn(dv.x); tr.gy = __double2float_rn(dv.y); tr.gz = __double2float_rn(dv.z);
            tr.ix = __fdiv_rn(1.0f, tr.gx); tr.iy = __fdiv_rn(1.0f, tr.gy); tr.iz = __fdiv_rn(1.0f, tr.gz);
            tr.nox = tr.noy = tr.noz = 0.0f;         // the view origin itself: the slab test is (plane * 1/g)
            tr.tcull = CUDART_INF_F;
            int stack[kTlasStackEntries];            // own stack: closest_hit below uses the shared one
            int sp = 0, cur = 0;
            for (;;) {
                if (cur >= 0) {
                    const float4* p = reinterpret_cast<const float4*>(f.tlas_nodes + cur);
                    const float4 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
                    const int4 d = __ldg(reinterpret_cast<const int4*>(p + 3));
                    c->node_visits++;
                    float t0, t1;
                    const bool h0 = fslab(tr, a.x, a.y, a.z, a.w, b.x, b.y, &t0);
                    const bool h1 = fslab(tr, b.z, b.w, cc.x, cc.y, cc.z, cc.w, &t1);
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
                        const d3 d = mul3x3(in.Minv, dv);
                        Hit hi;
                        if (closest_hit(sc, sc.meshes[in.mesh], f.subdivision, f.filter_mode, mk(in.start[0], in.start[1], in.start[2]),
                                        d, &hi, xc, false, f.filter_mode == 2 ? kNoHit : best) &&
                            (hi.rf < best || (hi.rf == best && i < which))) {
                            best = hi.rf; h = hi; which = i; hit = true; dir0 = d;
                            tr.tcull = __double2float_ru(best) * 1.00002f + 1e-6f;
                        }
                    }
                }
                if (sp == 0) break;
                cur = stack[--sp];
            }
        } else {
            for (int i = 0; i < f.n_instances; i++) {
                const DevInstance& in = insts[i];
                const d3 d = mul3x3(in.Minv, dv);
                Hit hi;
                if (closest_hit(sc, sc.meshes[in.mesh], f.subdivision, f.filter_mode, mk(in.start[0], in.start[1], in.start[2]), d, &hi,
                                xc, false, f.filter_mode == 2 ? kNoHit : best) &&
                    hi.rf < best) {
                    best = hi.rf; h = hi; which = i; hit = true; dir0 = d;
                }
            }
        }
    }
    *hit_out = hit;
    if (!hit || !valid) return out;
    const DevInstance& in = insts[which];
    const DevMesh& m = sc.meshes[in.mesh];
    out.id = h.id >= 0 ? in.tri_base + h.id : h.id;
    // local colour of the camera hit, then of up to reflection_depth mirror hits (one call site: the
    // shading + shadow code is the bulk of the kernel, it must not be instantiated twice)
    uint32_t local[5];
    int depth = 0;
    uint32_t tail = 0; bool have_tail = false;
    const int bounces = (f.reflection_depth > 0 && f.n_instances == 1) ? f.reflection_depth : 0;
    d3 d = dir0;
    for (int b = 0;; b++) {
        local[depth] = shade_and_shadow(f, sc, in, m, offsets, h, c, xc, n_shadow, bundle_score);
        if (b >= bounces) break;
        // r = d - n * (2 (d.n)), from pos + n*0.001 (PathTracingMethod.cs:10,52)
        const d3 r = vsub(d, vscale(h.normal, dmul(2.0, vdot(d, h.normal))));
        const d3 rs = vadd(h.pos, vscale(h.normal, 0.001));
        (*n_secondary)++;
        Hit h2;
        if (!closest_hit<STAGED>(sc, m, f.subdivision, f.filter_mode, rs, r, &h2, xc, false, 1.7976931348623157e308, stage)) { tail = f.background; have_tail = true; break; }
        h = h2; d = r;
        depth++;
    }
    uint32_t acc;
    if (have_tail) acc = mirror_blend(local[depth], tail); else acc = local[depth];
    for (int k = depth - 1; k >= 0; k--) acc = mirror_blend(local[k], acc);
    out.color = acc;
    return out;
}

#ifndef SR_MIN_BLOCKS
#define SR_MIN_BLOCKS (768 / SR_THREADS)     // 768 threads x 85 registers per SM: measured best over the four big configs (profiles/)
#endif
template <bool STAGED>
__global__ void __launch_bounds__(SR_THREADS, SR_MIN_BLOCKS)
render_kernel(const __grid_constant__ DevFrame f, const __grid_constant__ DevScene sc, const DevInstance* __restrict__ g_insts, const double* __restrict__ g_offsets,
              uint32_t* __restrict__ pixels, int32_t* __restrict__ hit_ids, unsigned int* __restrict__ tile_counter,
              DevCounters* __restrict__ counters)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevInstance* s_insts = reinterpret_cast<DevInstance*>(smem_raw);
    double* s_offsets = reinterpret_cast<double*>(smem_raw + sizeof(DevInstance) * f.n_instances);
    {
        const int n_words = (int)(sizeof(DevInstance) * f.n_instances / 8);
        const double* src = reinterpret_cast<const double*>(g_insts);
        double* dst = reinterpret_cast<double*>(s_insts);
        for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
        const int n_off = f.shadows ? 3 * f.shadow_samples : 0;
        for (int i = threadIdx.x; i < n_off; i += blockDim.x) s_offsets[i] = g_offsets[i];
    }
    // a small sphere set (config2: 1000 spheres = 16 KB of filter records + their tree) is staged in shared memory:
    // every camera ray of the block walks it (BASELINE.json north_star: "shared-memory staging of small sphere sets")
    SphereStage stage_v; const SphereStage* stage = nullptr;
    if (STAGED) {
        const size_t at = (sizeof(DevInstance) * (size_t)f.n_instances + sizeof(double) * 3 * (size_t)(f.shadows ? f.shadow_samples : 0) + 63) & ~(size_t)63;
        float4* s_nodes = reinterpret_cast<float4*>(smem_raw + at);
        float4* s_filt = s_nodes + 4 * (size_t)sc.n_sphere_nodes;
        const float4* gn = reinterpret_cast<const float4*>(sc.sphere_nodes);
        for (int i = threadIdx.x; i < 4 * sc.n_sphere_nodes; i += blockDim.x) s_nodes[i] = __ldg(gn + i);
        for (int i = threadIdx.x; i < sc.n_spheres; i += blockDim.x) s_filt[i] = __ldg(sc.sph_filt + i);
        stage_v.nodes = reinterpret_cast<const BvhNode*>(s_nodes); stage_v.filt = s_filt;
        stage = &stage_v;
    }
    __syncthreads();

    __shared__ int s_acc[SR_THREADS / 32][32][4];           // per warp, per pixel of its tile: sums of R, G, B and the hit id
    const int lane = threadIdx.x & 31;
    const int W = f.width, H = f.height, n = f.sub_pixel_res;
    const int n_tiles = f.tiles_x * f.tiles_y;
    Counters c; c.node_visits = 0; c.prim_tests = 0; c.sphere_tests = 0; c.shaded = 0;
    c.filter_tests = 0; c.filter_unsure = 0; c.filter_mismatch = 0; c.bundled = 0;
    c.bundle_skip = 0;
    int walk_stack[kStackEntries];
    c.stack = walk_stack;
    XCounters xc; xc.stack = walk_stack; xc.node_visits = 0; xc.prim_tests = 0; xc.sphere_tests = 0; xc.filter_tests = 0; xc.filter_unsure = 0;
    xc.filter_mismatch = 0;
    unsigned int n_primary = 0, n_shadow = 0, n_secondary = 0, n_hits = 0;

    // phase synchronisation (uniform over the launch): the warps of a block take consecutive tiles and start them
    // together.  Every warp takes part in every barrier: one without a tile repeats the last tile and discards the
    // result (into xc_void).
    __shared__ int s_tile_base;
    __shared__ unsigned int s_bundle_score[2];      // cone walks of this block that succeeded / failed (shade_and_shadow)
    if (threadIdx.x == 0) { s_bundle_score[0] = 0u; s_bundle_score[1] = 0u; }
    __syncthreads();
    const int sync = f.phase_sync;             // bit 0: tile fetch + start of a camera ray; bits 1-4: the stages of closest_hit
    XCounters xc_void = xc;
    for (;;) {
        int tile = 0;
        bool tile_valid = true;
        if (sync) {
            __syncthreads();
            if (threadIdx.x == 0) s_tile_base = (int)atomicAdd(tile_counter, (unsigned int)(SR_THREADS / 32));
            __syncthreads();
            if (s_tile_base >= n_tiles) break;
            tile = s_tile_base + (int)(threadIdx.x >> 5);
            if (tile >= n_tiles) { tile = n_tiles - 1; tile_valid = false; }
        } else {
            if (lane == 0) tile = (int)atomicAdd(tile_counter, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= n_tiles) break;
        }
        const int ty = tile / f.tiles_x, tx = tile - ty * f.tiles_x;
        const int band_j = ty / f.tiles_per_band;
        const int row0 = f.start_row + (f.band_index + band_j * f.band_count) * f.band_height;
        const int band_r0 = (ty - band_j * f.tiles_per_band) * 4;

        // The work items of a tile are its (pixel, sub-pixel) pairs, pixel-major, 32 at a time across the
        // lanes: with n x n supersampling the lanes of a warp hold sub-rays of the SAME pixel (or of a few
        // adjacent ones), which traverse almost identically, instead of 32 different pixels' rays looping
        // over their sub-rays out of step.  Sums are integer (Renderer.cs:1811-1822), so the order in which
        // the sub-rays arrive is irrelevant.  1 spp: one item per pixel, lane == pixel.
        int* acc = s_acc[threadIdx.x >> 5][lane];
        acc[0] = 0; acc[1] = 0; acc[2] = 0; acc[3] = -1;
        __syncwarp();
        const int nn = n * n;
        for (int w = lane; w < 32 * nn; w += 32) {
            const int px = w / nn, si = w - px * nn;
            const int col = tx * 8 + (px & 7);
            const int band_r = band_r0 + (px >> 3);
            const int row = row0 + band_r;
            const bool valid = tile_valid && !(col >= W || band_r >= f.band_height || row > f.end_row);
            if (!sync && !valid) continue;
            SR_SYNC_POINT(sync & 1);
            const int sx = si / n, sy = si - sx * n;                                      // subX outer, subY inner (:1762-1764)
            double fx = 0.0, fy = 0.0;                                                    // n == 1: (col + 0.0) / W == col / W
            if (n > 1) {
                fx = dsub(ddiv((double)sx, (double)(n - 1)), 0.5);                        // :1767-1768
                fy = dsub(ddiv((double)sy, (double)(n - 1)), 0.5);
            }
            const DevInstance& in0 = s_insts[0];
            d3 start, dir; bool is_view;
            if (f.focal_blur && n > 1) {                                                  // App. A #11
                const d3 dir_view = mk(-dsub(ddiv((double)col, (double)W), 0.5),
                                       dmul(-dsub(ddiv((double)row, (double)H), 0.5), f.aspect), f.fov_depth);
                const d3 dw = mul3x3(in0.Minv, dir_view);
                const d3 focal_pt = vadd(vscale(dw, f.focal_depth), mk(in0.start[0], in0.start[1], in0.start[2]));   // :1759
                const d3 sv = mk(dmul(ddiv(fx, (double)W), f.focal_strength), dmul(ddiv(fy, (double)H), f.focal_strength),
                                 -in0.pos_z);                                             // :1776-1778
                start = mul3x3(in0.Minv, sv);
                dir = vsub(focal_pt, start);                                              // :1790
                is_view = false;
            } else {
                start = mk(in0.start[0], in0.start[1], in0.start[2]);
                dir = mk(-dsub(ddiv(dadd((double)col, fx), (double)W), 0.5),
                         dmul(-dsub(ddiv(dadd((double)row, fy), (double)H), 0.5), f.aspect), f.fov_depth);   // :1728, :1794-1796
                is_view = true;
            }
            if (valid) n_primary++;
            bool hit = false;
            const PixelOut s1 = trace_camera_ray<STAGED>(f, sc, s_insts, s_offsets, &start, &dir, is_view, &c, valid ? &xc : &xc_void, &n_shadow,
                                                 &n_secondary, &hit, sync, valid, s_bundle_score, stage);
            if (!valid) continue;
            if (hit) n_hits++;
            int* pa = s_acc[threadIdx.x >> 5][px];
            if (nn == 1) {
                pa[0] = (int)((s1.color >> 16) & 0xff); pa[1] = (int)((s1.color >> 8) & 0xff); pa[2] = (int)(s1.color & 0xff);
            } else {
                atomicAdd(pa + 0, (int)((s1.color >> 16) & 0xff));
                atomicAdd(pa + 1, (int)((s1.color >> 8) & 0xff));
                atomicAdd(pa + 2, (int)(s1.color & 0xff));
            }
            if (si == nn - 1) pa[3] = s1.id;                                              // hit id of the last sub-ray
        }
        __syncwarp();
        {
            const int col = tx * 8 + (lane & 7);
            const int band_r = band_r0 + (lane >> 3);
            const int row = row0 + band_r;
            if (tile_valid && col < W && band_r < f.band_height && row <= f.end_row) {
                const int sum_r = acc[0] / nn, sum_g = acc[1] / nn, sum_b = acc[2] / nn;  // :1820-1822
                const size_t idx = (size_t)row * (size_t)W + (size_t)col;
                pixels[idx] = 0xff000000u | ((uint32_t)(sum_r & 0xff) << 16) | ((uint32_t)(sum_g & 0xff) << 8) |
                              (uint32_t)(sum_b & 0xff);                                   // Surface.DrawPixel
                if (hit_ids) hit_ids[idx] = acc[3];
            }
        }
        __syncwarp();
    }

    // one atomic per warp per counter
    unsigned long long v[12] = {n_primary, n_shadow, n_secondary, (unsigned long long)c.node_visits + xc.node_visits,
                                (unsigned long long)c.prim_tests + xc.prim_tests, (unsigned long long)c.sphere_tests + xc.sphere_tests,
                                n_hits, c.shaded, (unsigned long long)c.filter_tests + xc.filter_tests,
                                (unsigned long long)c.filter_unsure + xc.filter_unsure,
                                (unsigned long long)c.filter_mismatch + xc.filter_mismatch, c.bundled};
#pragma unroll
    for (int k = 0; k < 12; k++) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + k, x);
    }
}

// host-callable launcher (sr_api.cu)
cudaError_t launch_render(const DevFrame& f, const DevScene& sc, const DevInstance* d_insts, const double* d_offsets,
                          uint32_t* d_pixels, int32_t* d_ids, unsigned int* d_tile_counter, DevCounters* d_counters,
                          int grid_blocks, cudaStream_t stream)
{
    size_t smem = sizeof(DevInstance) * (size_t)f.n_instances + (f.shadows ? sizeof(double) * 3 * (size_t)f.shadow_samples : 0);
    if (f.stage_spheres) smem = ((smem + 63) & ~(size_t)63) + sizeof(BvhNode) * (size_t)sc.n_sphere_nodes + sizeof(float4) * (size_t)sc.n_spheres;
    if (smem > 48 * 1024) {
        cudaError_t e = f.stage_spheres ? cudaFuncSetAttribute(render_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                        : cudaFuncSetAttribute(render_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (f.stage_spheres) render_kernel<true><<<grid_blocks, SR_THREADS, smem, stream>>>(f, sc, d_insts, d_offsets, d_pixels, d_ids, d_tile_counter, d_counters);
    else render_kernel<false><<<grid_blocks, SR_THREADS, smem, stream>>>(f, sc, d_insts, d_offsets, d_pixels, d_ids, d_tile_counter, d_counters);
    return cudaGetLastError();
}

int render_kernel_block_threads() { return SR_THREADS; }

int render_kernel_occupancy(int smem_bytes, bool staged)
{
    int nb = 0;
    if (smem_bytes > 48 * 1024) {
        const cudaError_t e = staged ? cudaFuncSetAttribute(render_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
                                     : cudaFuncSetAttribute(render_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return 0;
    }
    const cudaError_t e = staged ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, render_kernel<true>, SR_THREADS, (size_t)smem_bytes)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, render_kernel<false>, SR_THREADS, (size_t)smem_bytes);
    if (e != cudaSuccess) return 0;
    return nb;
}

}  // namespace sr
