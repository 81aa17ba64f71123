/*
 * softray_cuda.h -- C ABI of libsoftray_cuda.so, the B200 (sm_100a) drop-in for SoftRay's
 * per-pixel raytrace hot path.
 *
 * The reference (voidstar69/softray, C#) has no FFI today.  The seam this library plugs into is
 * the private frame driver  Renderer.RaytraceGeometry(Instance)  (Engine3D/Renderer.cs:1501-1687):
 * everything after PreCalculate() (:1531) -- decorator chain, row blocks, RaytraceBlock,
 * TraceRayComplex/TraceRaySimple, IRayIntersectable.IntersectRay, Surface.DrawPixel -- becomes one
 * P/Invoke to softray_render().  See INTEGRATION.md for the C# binding.
 *
 * Conventions
 *   - plain C, POD structs, natural alignment (every struct below is free of padding surprises:
 *     8-byte members first or explicitly padded), no callbacks, no C++ types, no torch types.
 *   - every entry point returns SOFTRAY_OK (0) or a negative SOFTRAY_E_* code; nothing throws or
 *     aborts across the boundary.  softray_last_error() gives a UTF-8 message for the last failure
 *     on that context (or the process-wide last error when ctx is NULL).
 *   - the caller owns every pointer it passes; scene_create copies; render writes only
 *     pixels/hit_ids/stats.
 *   - one in-flight call per softray_ctx ("Not multithread safe!", Renderer.cs:1498); distinct
 *     contexts are independent (one context per GPU / per process rank).
 *   - NO CPU FALLBACK: without a CUDA device softray_create fails with SOFTRAY_E_NO_DEVICE.
 *   - pixel format: 0xAARRGGBB in a little-endian uint32 (Surface.cs:98-101), row-major,
 *     index y*width+x (Surface.cs:179).
 */
#ifndef SOFTRAY_CUDA_H
#define SOFTRAY_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOFTRAY_ABI_VERSION 4
#define SOFTRAY_MAX_INSTANCES 128   /* instances per frame (composite extension) */
#define SOFTRAY_MAX_SHADOW_SAMPLES 1024

/* ---- error codes (map back to the .NET exceptions the reference throws) ------------------- */
#define SOFTRAY_OK                     0
#define SOFTRAY_E_INVALID_ARG        (-1)  /* ArgumentNullException / ArgumentOutOfRangeException   */
#define SOFTRAY_E_VERTEX_OUTSIDE_BBOX (-2) /* "A triangle vertex is outside the bounding box",
                                              SpatialSubdivision.cs:285-295                          */
#define SOFTRAY_E_NO_DEVICE          (-3)  /* no CUDA device: there is no CPU fallback               */
#define SOFTRAY_E_CUDA               (-4)  /* a CUDA runtime call failed                             */
#define SOFTRAY_E_OOM                (-5)  /* host or device allocation failed                       */
#define SOFTRAY_E_UNSUPPORTED        (-6)  /* a flag combination this path does not implement
                                              (voxels, light fields, AO, path tracing, static
                                              shadow cache: they stay on the managed path)          */
#define SOFTRAY_E_FORMAT             (-7)  /* FormatException from the 3DS loader (Model.cs:555)     */
#define SOFTRAY_E_TIMEOUT            (-8)  /* softray_host_barrier: a rank process never arrived     */

typedef struct softray_ctx   softray_ctx;    /* opaque: one CUDA device, its streams and buffers     */
typedef struct softray_scene softray_scene;  /* opaque: device-resident SoA geometry + BVHs          */

/* ---- context -------------------------------------------------------------------------------- */

/* device_ordinal: CUDA device index (a torchrun rank passes LOCAL_RANK).  Replaces nothing in
 * the reference (it has no device); lifetime = where Renderer caches geometry_simple /
 * geometry_subdivided / rootGeometry (Renderer.cs:173-175) and Renderer.Dispose (:236-255). */
int  softray_create(int32_t device_ordinal, softray_ctx** out);
void softray_destroy(softray_ctx* ctx);   /* also releases the scenes the context still owns; a handle that
                                            is not live (NULL, already destroyed) is ignored -- finalisers may
                                            run in any order (Renderer.Dispose vs the .NET finaliser thread) */
/* A GROUP context over the first n_devices CUDA devices of this process (0 = every visible device): the drop-in
 * for a single-process host that wants all GPUs of the box behind one softray_render -- the reference fans the
 * rows of one Render() out to rayTraceConcurrency tasks that share surface.Pixels (Renderer.cs:1655-1680).
 * softray_scene_create builds the scene once and replicates it device-to-device; softray_render cuts
 * [start_row,end_row] into interleaved row bands, one set per device, and every device stores its bands into the
 * caller's surface (directly over its own PCIe link when the surface is page-locked: softray_host_register);
 * statistics are summed (times: the slowest device).  softray_render_device on a group is synchronous and needs a
 * framebuffer every device can store into: softray_device_alloc on the group (device 0's memory, peer-mapped).
 * softray_frame.band_count must be <= 1: the group partitions the rows itself. */
int  softray_create_multi(int32_t n_devices, softray_ctx** out);
int  softray_device_count(const softray_ctx* ctx);   /* 1 for a softray_create context */
const char* softray_last_error(const softray_ctx* ctx);
int  softray_abi_version(void);
/* sizeof of the PODs below as this library was compiled: 0 mesh, 1 sphere, 2 scene_desc,
 * 3 instance, 4 frame, 5 stats (a binding checks its own layout against these). */
int  softray_abi_sizeof(int32_t which);

/* ---- scene: Model -> SoA triangles (+BVH), ExtraGeometryToRaytrace -> SoA spheres ------------ */

/* One Model after PostProcessGeometry (Model.cs:750-831): vertices already in the unit cube.
 * Triangle i keeps its position in Model.Triangles => IntersectionInfo.triIndex
 * (Renderer.cs:1452-1469).  tri_argb[i] = Surface.PackColorAndAlpha(tri.diffuseMaterial, 1.0)
 * (Renderer.cs:1463, Surface.cs:131-138).  bbox = Model.Min/Max (Renderer.cs:1487). */
typedef struct softray_mesh {
    const double*   verts_xyz;   /* n_verts * 3                                                    */
    const int32_t*  tri_vidx;    /* n_tris * 3 (vertexIndex1..3, Model.cs:44-46)                   */
    const uint32_t* tri_argb;    /* n_tris, alpha must be 0xFF (Triangle.cs:31)                    */
    int32_t         n_verts;
    int32_t         n_tris;
    double          bbox_min[3];
    double          bbox_max[3];
} softray_mesh;

/* Raytrace.Sphere(center, radius){Color} (Sphere.cs:25-32,50-54); argb = Color.ToARGB()
 * (Color.cs:105-111). */
typedef struct softray_sphere {
    double   cx, cy, cz, r;
    uint32_t argb;
    uint32_t _pad;
} softray_sphere;

#define SOFTRAY_ACCEL_BVH    0   /* library-built BVHs (default); results identical to brute force  */
#define SOFTRAY_ACCEL_BRUTE  1   /* linear scan of every primitive per ray, like
                                    GeometryCollection.IntersectRay (GeometryCollection.cs:44-69)   */
#define SOFTRAY_ACCEL_LBVH   2   /* meshes flattened and their trees built ON THE DEVICE (Morton-code
                                    LBVH, deterministic): scene_create in milliseconds for dynamic
                                    scenes; slower to trace than the host-built SAH tree.  Spheres
                                    still use the host builder.  Results identical.                 */

typedef struct softray_scene_desc {
    const softray_mesh*   meshes;     int32_t n_meshes;   int32_t accel;  /* SOFTRAY_ACCEL_*       */
    /* ExtraGeometryToRaytrace (Renderer.cs:460,1545-1549): tested BEFORE the mesh, so a sphere
     * wins an exact rayFrac tie against a triangle (GeometryCollection.cs:53). */
    const softray_sphere* spheres;    int32_t n_spheres;  int32_t _pad;
} softray_scene_desc;

/* Replaces MakeRayTracableGeometry_simple/_subdivided (Renderer.cs:1452-1494) + the Triangle
 * ctor precompute (Triangle.cs:29-57) + the SpatialSubdivision ctor (SpatialSubdivision.cs:267-315,
 * including its vertex-inside-bbox check).  Deterministic: the same input gives a bit-identical
 * device layout on every call. */
int  softray_scene_create(softray_ctx* ctx, const softray_scene_desc* desc, softray_scene** out);
void softray_scene_destroy(softray_scene* scene);   /* no-op on a scene already released (twice, or with its context) */

/* Layout fingerprint (FNV-1a over every device-resident scene buffer, in upload order) --
 * the "bit-identical layout across runs" check. */
int  softray_scene_fingerprint(const softray_scene* scene, uint64_t* out);

/* ---- frame ---------------------------------------------------------------------------------- */

/* Instance (Instance.cs): the C# side passes the matrices it already builds in InitRender
 * (Instance.cs:134-135) so sin/cos stay .NET's.  Row-major 4x4, M[row*4+col]. */
typedef struct softray_instance {
    double  M[16];      /* _transform        = T(pos) * Roll * Pitch * Yaw                          */
    double  Minv[16];   /* _inverseTransform = Yaw(-) * Pitch(-) * Roll(-) * T(-pos)                */
    double  pos[3];     /* Instance.Position (view space)                                           */
    int32_t mesh_id;    /* index into softray_scene_desc.meshes                                     */
    int32_t _pad;
} softray_instance;

typedef struct softray_frame {
    /* lighting, view space (Renderer.cs:38-41,207-217) */
    double  ambient;                 /* ambientLight_intensity        (0.1)                        */
    double  shininess;               /* specularLight_shininess       (100)                        */
    double  light_dir_view[3];       /* directionalLight_dir          (normalise(-1,-1,1))         */
    double  light_pos_view[3];       /* positionalLight_pos           ((0,0,1.5) - 2*dir)          */
    double  fov_depth;               /* fieldOfViewDepth = 0.5/tan(22.5 deg) (Renderer.cs:97-101)  */
    double  focal_depth;             /* rayTraceFocalDepth            (1.5)                        */
    double  focal_strength;          /* rayTraceFocalBlurStrength     (10.0)                       */
    const softray_instance* instances;  /* n_instances == 1: reference semantics of
                                           RaytraceGeometry(instance).  > 1: nearest hit across
                                           instances (extension)       */
    int32_t n_instances;
    int32_t width, height;           /* SetRenderingSurface (Renderer.cs:593-626)                  */
    int32_t start_row, end_row;      /* rayTraceStartRow/EndRow, inclusive, clamped like :1652-1653;
                                        exactly these rows are written         */
    int32_t sub_pixel_res;           /* rayTraceSubPixelRes           (1)                          */
    int32_t focal_blur;              /* rayTraceFocalBlur; only acts when sub_pixel_res > 1        */
    int32_t subdivision;             /* rayTraceSubdivision (true): the mesh is traced through the
                                        root-box clip of SpatialSubdivision.IntersectRay (:389-416) */
    int32_t shading;                 /* rayTraceShading               (true)                       */
    int32_t shadows;                 /* rayTraceShadows, dynamic mode (false)                      */
    int32_t shadow_samples;          /* softShadowQuality             (100, ShadowMethod.cs:9)     */
    int32_t point_lighting;          /* pointLighting                 (true)                       */
    int32_t specular_lighting;       /* specularLighting              (true)                       */
    int32_t random_seed;             /* rayTraceRandomSeed            (1234567890)                 */
    uint32_t background_argb;        /* BackgroundColor; written as bg | 0xFF000000 (:1860)        */
    /* extensions, 0 = reference behaviour */
    int32_t reflection_depth;        /* mirror bounces, 0..4                                       */
    int32_t texture3d_id;            /* 0 none, 1 = procedural "marble" Texture3D<byte>            */
    /* row-band partition of [start_row,end_row] (the reference's row blocks, Renderer.cs:1659-1670,
     * made explicit for one-process-per-GPU rendering): with band_count > 1 only the rows with
     * ((row - start_row) / band_height) % band_count == band_index are traced and written.
     * band_count <= 1 (or band_height <= 0) means every row. */
    int32_t band_height;
    int32_t band_count;
    int32_t band_index;
    /* how shadow-ray queries are answered (results are identical in every mode; DESIGN.md
     * "Filtered predicates"):  SOFTRAY_FILTER_AUTO  FP32 filter with proven error bounds, FP64
     * reference arithmetic only for the rays the filter cannot decide (default);
     * SOFTRAY_FILTER_OFF  every ray through the FP64 reference arithmetic;
     * SOFTRAY_FILTER_VERIFY  both on every ray, contradictions counted in
     * softray_stats.filter_mismatch (must stay 0). */
    int32_t filter_mode;
    /* != 0: time every stage kernel of the frame with CUDA events (softray_stats.ms_stage; needs stats != NULL).
     * The stage-kernel pipeline then runs its chunks one after the other instead of two at a time, so the
     * frame itself is slower: a measuring aid (bench.py's per-kernel roofline), off by default. */
    int32_t profile_stages;
    int32_t _reserved[2];
} softray_frame;

#define SOFTRAY_N_STAGES 10
#define SOFTRAY_STAGE_SEARCH        0   /* camera rays: ray generation + FP32 filtered closest-hit search        */
#define SOFTRAY_STAGE_HIT           1   /* camera rays: candidates through the reference arithmetic + shading    */
#define SOFTRAY_STAGE_FALLBACK      2   /* camera rays the search could not bracket: full exact walk             */
#define SOFTRAY_STAGE_SEARCH_REF    3   /* the same three for the reflection rays (all bounces)                  */
#define SOFTRAY_STAGE_HIT_REF       4
#define SOFTRAY_STAGE_FALLBACK_REF  5
#define SOFTRAY_STAGE_SHADOW        6   /* ShadowMethod rays (cone tests + filtered any-hit walks)               */
#define SOFTRAY_STAGE_SHADOW_FB     7   /* shadow rays the filter could not decide                               */
#define SOFTRAY_STAGE_COMPOSE       8   /* shadow byte, mirror blend, sub-pixel sums, pixel store                */

#define SOFTRAY_FILTER_AUTO   0
#define SOFTRAY_FILTER_OFF    1
#define SOFTRAY_FILTER_VERIFY 2

/* Counters (the reference's NumRaysFired / NumGeometryTests / NumNodeVisits, Renderer.cs:465-587)
 * and device timings of the last render. */
typedef struct softray_stats {
    uint64_t rays_primary;     /* camera rays                                                      */
    uint64_t rays_shadow;      /* ShadowMethod.TraceRaysForSoftShadows rays                        */
    uint64_t rays_secondary;   /* reflection rays (extension)                                      */
    uint64_t node_visits;      /* BVH nodes popped                                                 */
    uint64_t prim_tests;       /* exact (reference-arithmetic) sphere + triangle tests             */
    uint64_t sphere_tests;     /* of prim_tests, the ray/sphere ones                               */
    uint64_t hits_primary;     /* camera rays that hit geometry                                    */
    uint64_t shaded_hits;      /* hits that went through shading (primary + reflection hits)       */
    uint64_t launches;         /* kernels launched by this call                                    */
    uint64_t filter_tests;     /* FP32 filter primitive tests (not part of prim_tests)             */
    uint64_t filter_unsure;    /* rays the filter could not decide (answered in FP64 instead)      */
    uint64_t filter_mismatch;  /* SOFTRAY_FILTER_VERIFY: sure filter answers the FP64 path contradicts */
    uint64_t rays_bundled;     /* of rays_shadow: answered together, one conservative cone test per
                                  shading point proving that no triangle can occlude any of its rays */
    uint64_t rays_fallback;    /* camera / reflection rays whose candidate search could not bracket the hit (more
                                  than 4 candidate triangles, an axis-parallel direction ...): answered by the
                                  full reference-arithmetic walk (stage-kernel pipeline only)               */
    uint64_t rays_short_listed;/* of rays_shadow: tested against the <= 8 triangles the cone walk of their shading point
                                  could not rule out, without a walk of their own (stage-kernel pipeline only)    */
    double   ms_kernel;        /* CUDA-event time of the render kernel(s)                          */
    double   ms_h2d;           /* frame constants upload                                           */
    double   ms_d2h;           /* framebuffer (+hit ids) readback                                  */
    double   ms_total;         /* host wall time of the call                                       */
    /* softray_frame.profile_stages: CUDA-event time of each stage of the stage-kernel pipeline, summed over the
     * chunks of the frame (0 when not profiled, or when the frame ran the fused kernel: see `launches`) */
    double   ms_stage[SOFTRAY_N_STAGES];
} softray_stats;

/* Host-buffer entry point: the drop-in for the body of Renderer.RaytraceGeometry.
 *   pixels_argb : host, width*height uint32, caller-owned (C# pins surface.Pixels with `fixed`);
 *                 only rows [start_row,end_row] are written.
 *   hit_ids     : optional host width*height int32: triIndex >= 0 (flattened over meshes /
 *                 instances: see DESIGN.md), -1 miss, <= -2 means sphere index -(id+2).
 *                 With sub_pixel_res > 1 it records the LAST sub-ray of the pixel.
 *   stats       : optional. */
int  softray_render(softray_ctx* ctx, const softray_scene* scene, const softray_frame* frame,
                    uint32_t* pixels_argb, int32_t* hit_ids, softray_stats* stats);

/* Device-buffer entry point (the timed "inputs already resident" path and the multi-GPU path:
 * the band stays in HBM for the NCCL gather).  d_pixels/d_hit_ids are device pointers on the
 * context's device with the same full-frame indexing; `stream` is a cudaStream_t (0 = the
 * context's own stream).  Asynchronous with respect to the host unless stats != NULL (or the context is a group).
 * Frames of one context share its per-frame device scratch: a frame enqueued on another stream than the previous
 * one first waits (on the device) for that one, so consecutive asynchronous calls never race; no row of the frame
 * is this call's (start_row > end_row after the clamp, or a band set without rows): nothing is launched and *stats
 * is all zero. */
int  softray_render_device(softray_ctx* ctx, const softray_scene* scene, const softray_frame* frame,
                           uint32_t* d_pixels_argb, int32_t* d_hit_ids, void* stream,
                           softray_stats* stats);

/* ---- multi-GPU: peer-mapped framebuffer --------------------------------------------------------
 * One process per GPU renders its row bands (softray_frame.band_*) of the SAME frame.  The
 * reference's analogue is the row-block fan-out inside one process (Renderer.cs:1655-1680), where
 * every task stores into the one shared surface.Pixels; here rank 0 owns the framebuffer in its
 * HBM, exports it, and the other ranks map it over NVLink and pass the mapped pointer as
 * d_pixels_argb to softray_render_device: the render kernel's own coalesced stores are the gather.
 * (The unfused alternative -- render locally, then NCCL-gather the bands -- needs none of this.) */
#define SOFTRAY_IPC_HANDLE_BYTES 64
int  softray_device_alloc(softray_ctx* ctx, uint64_t bytes, void** d_ptr_out);   /* cudaMalloc'd base pointer */
int  softray_device_free(softray_ctx* ctx, void* d_ptr);
int  softray_ipc_export(softray_ctx* ctx, void* d_ptr /* from softray_device_alloc */, char handle[SOFTRAY_IPC_HANDLE_BYTES]);
int  softray_ipc_open(softray_ctx* ctx, const char handle[SOFTRAY_IPC_HANDLE_BYTES], void** d_ptr_out);
int  softray_ipc_close(softray_ctx* ctx, void* d_ptr);

/* Page-lock a caller-owned host framebuffer (C#: the pinned `int[] surface.Pixels`, Surface.cs:20-30; or a
 * shared-memory section several rank processes map) so that softray_render's kernel stores finished pixels
 * straight into it over PCIe -- no device framebuffer, no D2H copy, and with one process per GPU every GPU
 * writes its own row bands into the one host surface over its own PCIe link (the multi-GPU gather of the
 * host-buffer path).  Optional: an unregistered buffer takes the staged copy.  Unregister before freeing. */
int  softray_host_register(softray_ctx* ctx, void* host_ptr, uint64_t n_bytes);
int  softray_host_unregister(softray_ctx* ctx, void* host_ptr);
/* End-of-frame rendezvous of the rank processes that share a host surface: a sense-reversing spin barrier on two
 * uint32 words (zero-initialised by their owner) in memory all of them map.  Returns when n_ranks callers arrived;
 * no CUDA involved (the rank's own softray_render has already returned, i.e. its bands are in host memory).
 * A rank that never arrives (crashed) makes the others give up with SOFTRAY_E_TIMEOUT after 60 s
 * (SOFTRAY_BARRIER_TIMEOUT_S); the words are then unusable. */
int  softray_host_barrier(volatile uint32_t* two_words, uint32_t n_ranks);

/* ---- diagnostics ------------------------------------------------------------------------------
 * Measured FMA-issue peak of the context's device in TFLOP/s (FMA = 2 flops): the denominator of
 * the FP-issue roofline bench.py reports (MEASURED_PEAKS.json has no FP32/FP64 vector figure).
 * fp64 != 0 measures DFMA, else FFMA. */
int  softray_measure_fma_peak(softray_ctx* ctx, int32_t fp64, double* tflops_out);

/* ---- resolve (the step after the path) ---------------------------------------
 * Renderer.PostProcessImage (Renderer.cs:819-898) then Renderer.AntiAliasImage (:937-978) in one device
 * pass.  `src` is the surface the frame was traced into: (dst_w * aa_res) x (dst_h * aa_res) pixels
 * (AntiAliasResolution multiplies the surface size, Renderer.cs:366-410,593-626); every source pixel goes
 * through the style's colour function, then aa_res x aa_res blocks are summed per channel, divided
 * (truncating) and packed with alpha 0xFF.  aa_res == 1: the styled surface, alpha untouched (dst may
 * alias src).  The depth styles (DepthSmooth / DepthBanded / Normals) read the rasteriser's depth
 * buffers and are not part of the raytrace path: SOFTRAY_E_UNSUPPORTED. */
#define SOFTRAY_STYLE_STANDARD       0   /* Renderer.Style.Standard: nothing to do                  */
#define SOFTRAY_STYLE_COLOR_SHUFFLE  1   /* ZRGB -> 0GBR                                            */
#define SOFTRAY_STYLE_NEGATIVE       2   /* x == BackgroundColor ? x : 0x00ffffff - x (uint)        */
int  softray_resolve(softray_ctx* ctx, const uint32_t* src_argb, int32_t dst_width, int32_t dst_height,
                     int32_t aa_res, int32_t style, uint32_t background_argb, uint32_t* dst_argb);   /* host buffers */
int  softray_resolve_device(softray_ctx* ctx, const uint32_t* d_src_argb, int32_t dst_width, int32_t dst_height,
                            int32_t aa_res, int32_t style, uint32_t background_argb, uint32_t* d_dst_argb,
                            void* stream);

/* ---- model files (the on-disk format feeding the path) --------------------
 * Model.Load3ds (Model.cs:522-653) + the 3dsLoader chunk parser (3dsLoader/ThreeDSFile.cs:132-662) +
 * Model.PostProcessGeometry (Model.cs:750-790) + the per-triangle colour packing of
 * MakeRayTracableGeometry_simple (Renderer.cs:1452-1469): a .3DS byte stream becomes the flattened,
 * unit-cube-normalised arrays softray_scene_create takes.  Host code only (no device needed).
 * Malformed input answers SOFTRAY_E_FORMAT (the FormatException of Model.cs:555). */
typedef struct softray_model softray_model;
int  softray_model_load_3ds(const uint8_t* bytes, uint64_t n_bytes, softray_model** out);
/* fills `out` with pointers into the model, valid until softray_model_destroy */
int  softray_model_get_mesh(const softray_model* model, softray_mesh* out);
void softray_model_destroy(softray_model* model);

/* ---- helpers that mirror small reference functions the shim would otherwise re-implement ---- */

/* Instance.InitRender matrices (Instance.cs:134-135, Matrix.cs:74-168), computed with the C
 * library's sin/cos.  A C# host should pass its own matrices instead. */
void softray_instance_init(softray_instance* inst, const double pos[3],
                           double yaw, double pitch, double roll, int32_t mesh_id);

/* Renderer() constructor defaults (Renderer.cs:207-230,70-85) for a width x height surface. */
void softray_frame_defaults(softray_frame* frame, int32_t width, int32_t height);

#ifdef __cplusplus
}
#endif
#endif /* SOFTRAY_CUDA_H */
