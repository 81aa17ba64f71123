/*
 * softray_oracle.h -- CPU restatement of SoftRay's raytrace hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (softray_b200/, libsoftray_cuda.so) never links or calls it.
 *
 * The frame / scene PODs are the ones of the product's C ABI (include/softray_cuda.h) so a test
 * hands the very same structs to both sides.
 */
#ifndef SOFTRAY_ORACLE_H
#define SOFTRAY_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/softray_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- System.Random (.NET Framework 4.x BCL) --------------------------- */
typedef struct orc_random { int32_t a[56]; int32_t inext, inextp; } orc_random;
void    orc_random_init(orc_random* r, int32_t seed);
int32_t orc_random_next(orc_random* r);
double  orc_random_next_double(orc_random* r);

/* ---- primitives (Raytrace/Triangle.cs, Plane.cs, Sphere.cs, AxisAlignedBox.cs) --------------- */
typedef struct orc_hit {
    double   ray_frac;
    double   pos[3];
    double   normal[3];
    uint32_t color;
    int32_t  tri_index;   /* >= 0 triangle, -1 not a triangle (IRayIntersectable.cs:17)           */
    int32_t  prim_id;     /* oracle bookkeeping: >=0 triangle, <= -2 sphere -(id+2)                */
    int32_t  _pad;
} orc_hit;

/* returns 1 on hit */
int orc_triangle_intersect(const double v1[3], const double v2[3], const double v3[3], uint32_t color,
                           const double start[3], const double dir[3], orc_hit* out);
int orc_sphere_intersect(const double center[3], double radius, uint32_t color,
                         const double start[3], const double dir[3], orc_hit* out);
int orc_sphere_contains_point(const double center[3], double radius, const double pt[3]);
int orc_box_contains_point(const double bmin[3], const double bmax[3], const double pt[3]);
/* AxisAlignedBox.ClipLineSegment: start/end updated in place; returns 0 if entirely outside */
int orc_box_clip_line_segment(const double bmin[3], const double bmax[3], double start[3], double end[3]);

/* ---- SpatialSubdivision (Raytrace/SpatialSubdivision.cs) ------------------------------------ */
typedef struct orc_tree orc_tree;
/* verts: n_tris*9 doubles (v1,v2,v3); colors n_tris.  Returns SOFTRAY_E_VERTEX_OUTSIDE_BBOX like
 * the ctor's ArgumentOutOfRangeException. */
int  orc_tree_build(const double* tri_verts, const uint32_t* colors, int32_t n_tris,
                    const double bmin[3], const double bmax[3],
                    int32_t max_tree_depth, int32_t max_geometry_per_node, orc_tree** out);
void orc_tree_free(orc_tree* t);
/* out[0..5] = TreeDepth, NumNodes, NumLeafNodes, NumInternalNodes, triangle references summed
 * over leaves, largest leaf */
void orc_tree_stats(const orc_tree* t, int32_t out[6]);
int  orc_tree_intersect(const orc_tree* t, const double start[3], const double dir[3], orc_hit* out);
/* GeometryCollection.IntersectRay over the same triangles (brute force) */
int  orc_tree_brute_intersect(const orc_tree* t, const double start[3], const double dir[3], orc_hit* out);

/* ---- Model: 3DS loader + PostProcessGeometry (3dsLoader/ThreeDSFile.cs, Model.cs) ------------ */
typedef struct orc_model {
    double*   verts_xyz;  int32_t n_verts;  int32_t n_tris;
    int32_t*  tri_vidx;
    uint32_t* tri_argb;
    double    bbox_min[3], bbox_max[3];
} orc_model;
int  orc_model_load_3ds(const uint8_t* bytes, size_t n_bytes, orc_model** out);
/* The Cloth.cs pattern for procedural models: fill lists, CalcExtent(), PostProcessGeometry()
 * (Cloth.cs:39-40).  normalise=0 skips PostProcessGeometry (bbox = CalcExtent only). */
int  orc_model_from_arrays(const double* verts_xyz, int32_t n_verts, const int32_t* tri_vidx,
                           const uint32_t* tri_argb, int32_t n_tris, int32_t normalise, orc_model** out);
void orc_model_free(orc_model* m);

/* ---- frame driver (Renderer.cs:1501-1925 + the *Method decorators) -------------------------- */
typedef struct orc_options {
    int32_t concurrency;          /* rayTraceConcurrency (4): row blocks, each with its own
                                     System.Random(seed) (Renderer.cs:1659-1670,1693)               */
    int32_t n_threads;            /* host threads used to run it (0 = all cores); never changes
                                     the image                                                     */
    int32_t tree_max_depth;       /* 15  (SpatialSubdivision.cs:269)                               */
    int32_t tree_max_per_node;    /* 25  (SpatialSubdivision.cs:270)                               */
    int32_t path_tracing;         /* rayTracePathTracing: oracle-only, used to pin spheres and the
                                     per-block RNG against the reference's pathTracing_* goldens   */
    int32_t _pad;
    /* test aid: only the columns [col_start, col_end] of each row are produced (col_end < 0: the whole row).  Lets a
     * test compare a window of a FULL-SIZE frame (the rays of a pixel depend on the frame size) in seconds.       */
    int32_t col_start, col_end;
} orc_options;
void orc_options_defaults(orc_options* o);

typedef struct orc_scene orc_scene;
int  orc_scene_create(const softray_scene_desc* desc, const orc_options* opt, orc_scene** out);
void orc_scene_free(orc_scene* s);
/* stats of mesh m's tree, as orc_tree_stats */
void orc_scene_tree_stats(const orc_scene* s, int32_t mesh, int32_t out[6]);

/* optional per-pixel side buffers for the primary ray (last sub-ray when sub_pixel_res > 1) */
typedef struct orc_aux {
    double* ray_frac;   /* width*height or NULL; NaN on miss                                       */
    double* cos_theta;  /* dir.normal / |dir|  (the |cos| > 1e-4 "non-grazing" filter)             */
} orc_aux;

int  orc_render(const orc_scene* s, const softray_frame* f, const orc_options* opt,
                uint32_t* pixels_argb, int32_t* hit_ids, const orc_aux* aux, softray_stats* stats);

/* The 100 area-light offsets of ShadowMethod's ctor (ShadowMethod.cs:63-72) */
void orc_area_light_offsets(int32_t seed, int32_t n, double* out_xyz);
/* Texture3D extension texel (shared definition: DESIGN.md "Texture3D id 1") */
uint8_t orc_texture3d_sample(int32_t id, const double pos[3]);

#ifdef __cplusplus
}
#endif
#endif
