// sr_types.h -- device-resident scene layout and per-frame constants shared by the host side
// (sr_api.cu, sr_bvh.cpp) and the kernels (sr_render.cu).  See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>
#include <vector_types.h>

#if defined(__CUDACC__)
#define SR_HD __host__ __device__
#define SR_ALIGN(n) __align__(n)
#else
#define SR_HD
#define SR_ALIGN(n) alignas(n)
#endif

namespace sr {

// ---- exact (reference-arithmetic) primitive records -----------------------------------------
// One triangle = one 128-byte line, read with eight 128-bit loads.  Holds exactly what
// Triangle's ctor precomputes (Triangle.cs:29-57, Plane.cs:22-29) plus the two per-triangle
// dot products Triangle.IntersectRay re-evaluates on every call (Triangle.cs:90,95).
struct SR_ALIGN(16) TriRec {
    double nx, ny, nz;        // plane._normal (unit)
    double d;                 // plane._originDist = v1 . n
    double v1x, v1y, v1z;     // vertex1
    double den1;              // edge1 . edge2Perp
    double e2px, e2py, e2pz;  // edge2Perp = edge2 x n (un-normalised n)
    double den2;              // edge2 . edge1Perp
    double e1px, e1py, e1pz;  // edge1Perp = edge1 x n
    uint32_t color;           // 0xAARRGGBB
    int32_t  index;           // TriangleIndex (position in Model.Triangles)
};
static_assert(sizeof(TriRec) == 128, "TriRec must be one 128-byte line");

// One sphere = 64 bytes (Sphere.cs:9-11,25-32).
struct SR_ALIGN(16) SphereRec {
    double cx, cy, cz;
    double r;
    double r2;                // radiusSqr = radius * radius
    uint32_t color;
    int32_t  index;           // position in ExtraGeometryToRaytrace
    double _pad[2];
};
static_assert(sizeof(SphereRec) == 64, "SphereRec must be 64 bytes");

// ---- FP32 filter record (DESIGN.md "Filtered predicates"; tri_filter in sr_render.cu) -----------
// The same formulation as the exact test, rounded to FP32: plane (n, d), then
// s = (pos - v1) . a and u = (pos - v1) . b with a = edge2Perp / den1, b = edge1Perp / den2 folded on
// the host in FP64.  a1 / b1 = L1 norms of the ROUNDED a / b, rounded up: they scale the error
// bound of s / u.  a1 < 0 marks a triangle the exact test can never hit (zero area: Triangle.cs:42-43
// makes every quotient NaN); a1 = +inf one the filter must always hand to the exact test.
struct SR_ALIGN(16) TriFilt {
    float nx, ny, nz, d;
    float ax, ay, az, a1;
    float bx, by, bz, b1;
    float v1x, v1y, v1z, _pad;
};
static_assert(sizeof(TriFilt) == 64, "TriFilt must be 64 bytes");

// ---- BVH2 node: both children's boxes in FP32 + links = 64 bytes, four 128-bit loads ----------
// Boxes are rounded outward and padded (sr_bvh.cpp) so that the FP32 slab test can never reject a
// ray whose exact FP64 primitive test would hit.
struct SR_ALIGN(16) BvhNode {
    float lo0x, lo0y, lo0z, hi0x;
    float hi0y, hi0z, lo1x, lo1y;
    float lo1z, hi1x, hi1y, hi1z;
    int32_t child0, child1;     // traversal links: >= 0 node index; < 0 leaf -1 - (first * 16 + count); kNoChild
    int32_t count0, count1;     // primitives in the child leaf (0 = internal child, -1 = no child); host bookkeeping
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

constexpr int32_t kNoChild = -1;    // = a leaf of zero primitives; its box is the point (+FLT_MAX)^3, which no slab test enters
constexpr int kMaxLeafPrims = 15;        // fits the 4-bit count of a packed stack entry
constexpr int kMaxBvhDepth  = 60;        // builder falls back to median splits to stay below this
constexpr int kStackEntries = 64;
constexpr int kTinyMesh = 16;            // filtered walks test this few primitives directly instead of walking their tree
constexpr int kTlasStackEntries = 32;    // instance hierarchy (<= SOFTRAY_MAX_INSTANCES leaves)

struct DevMesh {
    const TriRec*  tris;        // leaf order (BVH) or Model.Triangles order (brute)
    const TriFilt* filt;        // same order as tris
    const BvhNode* nodes;       // nullptr in brute mode
    int32_t n_tris;             // number of records in `tris` (no duplication: one leaf per tri)
    int32_t n_nodes;
    double  bmin[3], bmax[3];   // root AxisAlignedBox = Model.Min/Max (Renderer.cs:1487)
    float   fmin[3], fmax[3];   // the same box rounded to FP32 (nearest)
    float   scale;              // largest |coordinate| of bmin/bmax, rounded up
    int32_t _pad;
    double  bs_center[3], bs_radius;   // a sphere around the box centre that holds every vertex (tighter than the box for round meshes)
};

struct DevScene {
    const DevMesh*   meshes;    int32_t n_meshes;  int32_t accel;
    const SphereRec* spheres;   // leaf order (BVH) or list order (brute)
    const BvhNode*   sphere_nodes;
    int32_t n_spheres;          int32_t n_sphere_nodes;
    double  sph_bmin[3], sph_bmax[3];   // bounds of all spheres (traversal entry clip only)
    const float4*    sph_filt;          // FP32 filter records (cx, cy, cz, r), same order as spheres; BVH mode only
    float   sph_fmin[3], sph_fmax[3];   // sph_bmin / sph_bmax rounded outward to FP32
    float   sph_scale;                  // largest |coordinate| of the sphere bounds, rounded up
    int32_t _pad2;
};

struct DevInstance {
    double M[12];               // rows 0..2 of _transform (3x4)
    double Minv[12];            // rows 0..2 of _inverseTransform
    double pos_z;               // Instance.Position.z
    double start[3];            // start_World (Renderer.cs:1717)
    double light_pos_model[3];  // positionalLight_pos_model (Renderer.cs:1515)
    double light_dir_model[3];  // directionalLight_dir_model (Renderer.cs:1513)
    int32_t mesh;               // index into DevScene.meshes
    int32_t tri_base;           // flattened hit-id base
    int32_t sph_can_shadow;     // 0: no sphere can report rayFrac <= 1 for any shadow ray of this frame
    int32_t _pad;
    double  bs_center_view[3];  // composite frames: the mesh's bounding sphere (DevMesh::bs_*) in view space ...
    double  bs_radius2;         // ... and its squared radius, padded: a view ray that misses it misses the instance
};

// n / d for a divisor fixed per frame and n < 2^31: one multiply-high and a shift instead of the ~20-instruction
// integer division (the stage kernels turn a ray index into tile / pixel / sub-pixel coordinates five divisions deep)
struct FastDiv {
    uint32_t mul, shift, d, _pad;
#if defined(__CUDACC__)
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1u ? n : (__umulhi(n, mul) >> shift); }
#endif
};

struct DevFrame {
    double ambient, shininess;
    double light_dir_view[3], light_pos_view[3];
    double fov_depth, focal_depth, focal_strength, aspect;
    double spec_skip;           // |R.cam| below this: pow(R.cam, shininess) < 2^-82 cannot change ambient + diffuse (shade())
    int32_t width, height, start_row, end_row;
    int32_t sub_pixel_res, focal_blur, subdivision, shading;
    int32_t shadows, shadow_samples, point_lighting, specular_lighting;
    int32_t reflection_depth, texture3d_id, n_instances;
    uint32_t background;        // already | 0xFF000000
    int32_t tiles_x, tiles_y;   // 8x4-pixel warp tiles over this launch's rows
    // row bands (softray_frame.band_*): local band j of this launch is global band
    // band_index + j * band_count, rows start_row + band * band_height ...; an unbanded frame is
    // one band of end_row - start_row + 1 rows
    int32_t band_height, band_count, band_index, tiles_per_band;
    FastDiv fd_n, fd_nn, fd_per_tile, fd_tiles_x, fd_tiles_per_band;   // sub_pixel_res, its square, 32 * that, tiles_x, tiles_per_band
    int32_t filter_mode;        // SOFTRAY_FILTER_*: 0 filter + exact fallback, 1 exact only, 2 verify
    float   light_radius;       // >= the length of every area-light offset (0.2, ShadowMethod.cs:10), rounded up
    int32_t bundle_budget;      // node visits a shadow-bundle cone walk may spend before giving up (0 = no bundles)
    int32_t stage_spheres;      // 1: the blocks of the fused kernel copy the sphere tree + filter records into shared memory
    int32_t phase_sync;         // 1: the warps of a block meet at barriers between the stages of a camera ray
                                // (sr_render.cu "Phase synchronisation"); 0: they run free
    // composite frames (n_instances > 1): a BVH over the view-space boxes of the instances, rebuilt per frame
    const BvhNode* tlas_nodes;  // leaf primitives index tlas_order
    const int32_t* tlas_order;  // instance index of the k-th TLAS leaf primitive
};

struct DevCounters {            // summed over the launch with one atomic per warp per counter
    unsigned long long rays_primary, rays_shadow, rays_secondary, node_visits, prim_tests,
        sphere_tests, hits_primary, shaded_hits, filter_tests, filter_unsure, filter_mismatch, rays_bundled, rays_fallback, rays_short_listed;
};

}  // namespace sr
