/*
 * softray_oracle.c -- CPU restatement of SoftRay's raytrace hot path in plain C (FP64, no FMA).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for libsoftray_cuda.so and the CPU baseline of bench.py.
 * Nothing under softray_b200/ may call into it.
 *
 * Every function cites the reference file:line (relative to the voidstar69/softray checkout) whose
 * arithmetic it follows.  Build with -ffp-contract=off and without -ffast-math: the reference
 * never fuses multiply-add and evaluates left to right.
 *
 * PARITY PIN: this file reproduces the reference's own golden images bit-exactly (tests/
 * test_oracle_goldens.py against tests/golden/reference_fixtures.npz), its five tree-build
 * known-answer tests and its Triangle/Sphere unit facts.  Extensions the reference does not have
 * (reflection, Texture3D shading, multi-instance compositing) are "parity unpinned": defined here,
 * checked oracle-vs-CUDA only.
 */
#include "softray_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ============================================================================================ */
/* System.Random -- Knuth subtractive generator (.NET BCL, not in the repo) */
/* ============================================================================================ */
#define ORC_MBIG  2147483647
#define ORC_MSEED 161803398

void orc_random_init(orc_random* r, int32_t seed)
{
    int32_t sub = (seed == INT32_MIN) ? ORC_MBIG : (seed < 0 ? -seed : seed);
    int32_t mj = ORC_MSEED - sub;
    int32_t mk = 1;
    memset(r->a, 0, sizeof r->a);
    r->a[55] = mj;
    for (int i = 1; i < 55; i++) {
        int ii = (21 * i) % 55;
        r->a[ii] = mk;
        mk = mj - mk;
        if (mk < 0) mk += ORC_MBIG;
        mj = r->a[ii];
    }
    for (int k = 1; k < 5; k++) {
        for (int i = 1; i < 56; i++) {
            /* 32-bit wrap-around subtraction, as C# unchecked int arithmetic */
            r->a[i] = (int32_t)((uint32_t)r->a[i] - (uint32_t)r->a[1 + (i + 30) % 55]);
            if (r->a[i] < 0) r->a[i] += ORC_MBIG;
        }
    }
    r->inext = 0;
    r->inextp = 21;
}

int32_t orc_random_next(orc_random* r)
{
    int32_t n = r->inext + 1, np = r->inextp + 1;
    if (n >= 56) n = 1;
    if (np >= 56) np = 1;
    int32_t v = (int32_t)((uint32_t)r->a[n] - (uint32_t)r->a[np]);
    if (v == ORC_MBIG) v--;
    if (v < 0) v += ORC_MBIG;
    r->a[n] = v;
    r->inext = n;
    r->inextp = np;
    return v;
}

double orc_random_next_double(orc_random* r)
{
    return orc_random_next(r) * (1.0 / ORC_MBIG);
}

/* ============================================================================================ */
/* Vector (Engine3D/Vector.cs:9-197)                                                            */
/* ============================================================================================ */
typedef struct { double x, y, z; } vec;

static inline vec v3(double x, double y, double z) { vec r = {x, y, z}; return r; }
static inline vec vadd(vec a, vec b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }      /* :55  */
static inline vec vsub(vec a, vec b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }      /* :60  */
static inline vec vmul(vec a, double b) { return v3(a.x * b, a.y * b, a.z * b); }         /* :69  */
static inline vec vneg(vec a) { return v3(-a.x, -a.y, -a.z); }                             /* :92  */
static inline double vdot(vec a, vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      /* :99  */
static inline vec vcross(vec a, vec b)                                                     /* :104 */
{
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline double vlen(vec a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }      /* :121 */
static inline double vlensqr(vec a) { return a.x * a.x + a.y * a.y + a.z * a.z; }         /* :132 */
static inline vec vnormalise(vec a)                                                        /* :177 */
{
    double inv = 1.0 / vlen(a);   /* multiplies by the reciprocal, does not divide */
    return v3(a.x * inv, a.y * inv, a.z * inv);
}
static inline int vis_zero(vec a)                                                          /* :140 */
{
    const double e = 1e-10;
    return -e < a.x && a.x < e && -e < a.y && a.y < e && -e < a.z && a.z < e;
}
static inline vec vfrom(const double p[3]) { return v3(p[0], p[1], p[2]); }
static inline void vto(double p[3], vec a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

/* ============================================================================================ */
/* Matrix (Engine3D/Matrix.cs:34-168), Instance matrices (Engine3D/Instance.cs:134-135)          */
/* ============================================================================================ */
typedef struct { double m[4][4]; } mat4;

static mat4 mat_mul(const mat4* a, const mat4* b)                                          /* :74-92 */
{
    mat4 r;
    for (int row = 0; row < 4; row++)
        for (int col = 0; col < 4; col++) {
            double sum = 0.0;
            for (int i = 0; i < 4; i++) sum += a->m[row][i] * b->m[i][col];
            r.m[row][col] = sum;
        }
    return r;
}
static mat4 mat_zero(void) { mat4 r; memset(&r, 0, sizeof r); return r; }
static mat4 mat_translation(vec p)                                                         /* :94-114 */
{
    mat4 r = mat_zero();
    r.m[0][0] = 1.0; r.m[1][1] = 1.0; r.m[2][2] = 1.0; r.m[3][3] = 1.0;
    r.m[0][3] = p.x; r.m[1][3] = p.y; r.m[2][3] = p.z;
    return r;
}
static mat4 mat_yaw(double a)                                                              /* :116-132 */
{
    mat4 r = mat_zero();
    r.m[0][0] = cos(a);  r.m[2][0] = sin(a);
    r.m[1][1] = 1.0;
    r.m[0][2] = -sin(a); r.m[2][2] = cos(a);
    r.m[3][3] = 1.0;
    return r;
}
static mat4 mat_pitch(double a)                                                            /* :134-150 */
{
    mat4 r = mat_zero();
    r.m[0][0] = 1.0;
    r.m[1][1] = cos(a);  r.m[2][1] = sin(a);
    r.m[1][2] = -sin(a); r.m[2][2] = cos(a);
    r.m[3][3] = 1.0;
    return r;
}
static mat4 mat_roll(double a)                                                             /* :152-168 */
{
    mat4 r = mat_zero();
    r.m[0][0] = cos(a);  r.m[1][0] = sin(a);
    r.m[0][1] = -sin(a); r.m[1][1] = cos(a);
    r.m[2][2] = 1.0;
    r.m[3][3] = 1.0;
    return r;
}

/* Instance.InitRender (Instance.cs:134-135).  Exported under the oracle's own name so tests can
 * compare it with the product's softray_instance_init. */
void orc_instance_init(softray_instance* inst, const double pos[3], double yaw, double pitch,
                       double roll, int32_t mesh_id)
{
    vec p = vfrom(pos);
    mat4 t = mat_translation(p), r = mat_roll(roll), pi = mat_pitch(pitch), y = mat_yaw(yaw);
    mat4 tr = mat_mul(&t, &r);
    mat4 trp = mat_mul(&tr, &pi);
    mat4 M = mat_mul(&trp, &y);
    mat4 yi = mat_yaw(-yaw), pii = mat_pitch(-pitch), ri = mat_roll(-roll), ti = mat_translation(vneg(p));
    mat4 a = mat_mul(&yi, &pii);
    mat4 b = mat_mul(&a, &ri);
    mat4 Minv = mat_mul(&b, &ti);
    memcpy(inst->M, M.m, sizeof inst->M);
    memcpy(inst->Minv, Minv.m, sizeof inst->Minv);
    inst->pos[0] = pos[0]; inst->pos[1] = pos[1]; inst->pos[2] = pos[2];
    inst->mesh_id = mesh_id;
    inst->_pad = 0;
}

/* Renderer() defaults (Renderer.cs:207-230, 70-85, 97-101) */
void orc_frame_defaults(softray_frame* f, int32_t width, int32_t height)
{
    memset(f, 0, sizeof *f);
    f->ambient = 0.1;
    vec d = vnormalise(v3(-1, -1, 1));
    vto(f->light_dir_view, d);
    vec lp = vsub(v3(0.0, 0.0, 1.5), vmul(d, 2));
    vto(f->light_pos_view, lp);
    f->shininess = 100.0;
    {
        const double fov_deg = 45.0;
        const double fov_rad = fov_deg / 180.0 * 3.14159265358979323846;
        f->fov_depth = 0.5 / tan(fov_rad / 2);
    }
    f->focal_depth = 1.5;
    f->focal_strength = 10.0;
    f->width = width; f->height = height;
    f->start_row = 0; f->end_row = height - 1;
    f->sub_pixel_res = 1;
    f->focal_blur = 1;
    f->subdivision = 1;
    f->shading = 1;
    f->shadows = 0;
    f->shadow_samples = 100;
    f->point_lighting = 1;
    f->specular_lighting = 1;
    f->random_seed = 1234567890;
    f->background_argb = 0;
    f->band_height = 0; f->band_count = 1; f->band_index = 0;
}

/* Matrix.Multiply3X3 / Instance.TransformDirection[Reverse] (Matrix.cs:34-41, Instance.cs:216-235) */
static inline vec mul3x3(const double* m, vec v)
{
    return v3(v.x * m[0] + v.y * m[1] + v.z * m[2],
              v.x * m[4] + v.y * m[5] + v.z * m[6],
              v.x * m[8] + v.y * m[9] + v.z * m[10]);
}
/* Matrix.Multiply3X4 (Matrix.cs:50-57) */
static inline vec mul3x4(const double* m, vec v)
{
    return v3(v.x * m[0] + v.y * m[1] + v.z * m[2] + m[3],
              v.x * m[4] + v.y * m[5] + v.z * m[6] + m[7],
              v.x * m[8] + v.y * m[9] + v.z * m[10] + m[11]);
}

/* ============================================================================================ */
/* Colour helpers (Engine3D/Color.cs:31-36,105-111,124-133; Surface.cs:98-108)                  */
/* ============================================================================================ */
static inline uint32_t modulate_packed(uint32_t color, uint8_t amount)                    /* Color.cs:124 */
{
    uint8_t r = (uint8_t)(color >> 16), g = (uint8_t)(color >> 8), b = (uint8_t)color;
    r = (uint8_t)((r * amount) >> 8);
    g = (uint8_t)((g * amount) >> 8);
    b = (uint8_t)((b * amount) >> 8);
    return (255u << 24) + ((uint32_t)r << 16) + ((uint32_t)g << 8) + b;
}
static inline uint8_t to_byte(double v) { return (uint8_t)(int32_t)v; }   /* C# (byte)double, in range */

/* ============================================================================================ */
/* Plane / Triangle / Sphere / AxisAlignedBox                                                   */
/* ============================================================================================ */
typedef struct { vec normal; double origin_dist; } plane_t;

/* Plane ctor (Plane.cs:22-29) */
static plane_t plane_make(vec point, vec normal)
{
    plane_t p;
    p.normal = vnormalise(normal);
    p.origin_dist = vdot(point, p.normal);
    return p;
}

/* Plane.IntersectRay (Plane.cs:67-103): one-sided */
static int plane_intersect_ray(const plane_t* p, vec start, vec dir, orc_hit* out)
{
    double start_dist = vdot(start, p->normal);
    double dir_dist = vdot(dir, p->normal);
    if (dir_dist >= 0.0) return 0;
    double ray_frac = p->origin_dist - start_dist;
    if (ray_frac <= 0.0) {
        ray_frac /= dir_dist;
        vto(out->pos, vadd(start, vmul(dir, ray_frac)));
        vto(out->normal, p->normal);
        out->ray_frac = ray_frac;
        out->tri_index = -1;
        return 1;
    }
    return 0;
}

/* Plane.IntersectLineSegment (Plane.cs:111-138) */
static int plane_intersect_segment(const plane_t* p, vec start, vec end, double* line_frac, vec* pos)
{
    double start_dist = vdot(start, p->normal);
    double end_dist = vdot(end, p->normal);
    double lf = (p->origin_dist - start_dist) / (end_dist - start_dist);
    if (0.0 <= lf && lf <= 1.0) {
        *line_frac = lf;
        *pos = vadd(start, vmul(vsub(end, start), lf));
        return 1;
    }
    return 0;
}

typedef struct {
    vec v1, v2, v3;
    vec edge1, edge2, edge1_perp, edge2_perp;
    plane_t plane;
    uint32_t color;
    int32_t index;
} tri_t;

/* Triangle ctor (Triangle.cs:29-57) */
static void tri_make(tri_t* t, vec v1, vec v2, vec v3_, uint32_t color, int32_t index)
{
    t->v1 = v1; t->v2 = v2; t->v3 = v3_;
    t->color = color;
    t->index = index;
    t->edge1 = vsub(v2, v1);
    t->edge2 = vsub(v3_, v1);
    vec normal = vcross(t->edge1, t->edge2);
    if (vis_zero(normal)) normal = v3(1, 0, 0);
    t->plane = plane_make(v1, normal);
    t->edge1_perp = vcross(t->edge1, normal);   /* un-normalised normal */
    t->edge2_perp = vcross(t->edge2, normal);
}

/* Triangle.IntersectRay (Triangle.cs:83-104) */
static int tri_intersect_ray(const tri_t* t, vec start, vec dir, orc_hit* out)
{
    orc_hit info;
    if (!plane_intersect_ray(&t->plane, start, dir, &info)) return 0;
    vec v1_to = vsub(vfrom(info.pos), t->v1);
    double s = vdot(v1_to, t->edge2_perp) / vdot(t->edge1, t->edge2_perp);
    if (s < 0.0 || s > 1.0) return 0;
    double u = vdot(v1_to, t->edge1_perp) / vdot(t->edge2, t->edge1_perp);
    if (s >= 0.0 && u >= 0.0 && s + u <= 1.0) {
        info.color = t->color;
        info.tri_index = t->index;
        info.prim_id = t->index;
        *out = info;
        return 1;
    }
    return 0;
}

int orc_triangle_intersect(const double v1[3], const double v2[3], const double v3_[3], uint32_t color,
                           const double start[3], const double dir[3], orc_hit* out)
{
    tri_t t;
    tri_make(&t, vfrom(v1), vfrom(v2), vfrom(v3_), color, -1);   /* TriangleIndex = -1 (Triangle.cs:53) */
    return tri_intersect_ray(&t, vfrom(start), vfrom(dir), out);
}

typedef struct { vec center; double radius, radius_sqr; uint32_t color; } sphere_t;

/* Sphere.IntersectRay (Sphere.cs:152-219).  rayFrac is a DISTANCE along the normalised dir. */
static int sphere_intersect_ray(const sphere_t* sp, int32_t index, vec start, vec dir, orc_hit* out)
{
    const double epsilon = 1e-10;
    dir = vnormalise(dir);
    vec s2s = vsub(start, sp->center);
    double proj = vdot(s2s, dir);
    if (proj > sp->radius) return 0;
    double dist_sqr = vlensqr(s2s);
    double term = proj * proj - dist_sqr + sp->radius_sqr;
    if (term < epsilon) return 0;
    double root = sqrt(term);
    double f1 = -proj - root;
    double f2 = -proj + root;
    double ray_frac = (f1 >= 0 ? f1 : f2);
    if (ray_frac < 0) return 0;
    out->ray_frac = ray_frac;
    vec pos = vadd(start, vmul(dir, ray_frac));
    vto(out->pos, pos);
    vto(out->normal, vnormalise(vsub(pos, sp->center)));
    out->color = sp->color;
    out->tri_index = -1;
    out->prim_id = -(index + 2);
    return 1;
}

int orc_sphere_intersect(const double center[3], double radius, uint32_t color,
                         const double start[3], const double dir[3], orc_hit* out)
{
    sphere_t sp = { vfrom(center), radius, radius * radius, color };
    return sphere_intersect_ray(&sp, 0, vfrom(start), vfrom(dir), out);
}

/* Sphere.ContainsPoint (Sphere.cs:56-60): strict < */
int orc_sphere_contains_point(const double center[3], double radius, const double pt[3])
{
    return vlensqr(vsub(vfrom(pt), vfrom(center))) < radius * radius;
}

typedef struct { vec min, max; plane_t planes[6]; } box_t;

/* AxisAlignedBox ctor (AxisAlignedBox.cs:15-28) */
static void box_make(box_t* b, vec mn, vec mx)
{
    b->min = mn; b->max = mx;
    b->planes[0] = plane_make(mn, v3(-1, 0, 0));
    b->planes[1] = plane_make(mn, v3(0, -1, 0));
    b->planes[2] = plane_make(mn, v3(0, 0, -1));
    b->planes[3] = plane_make(mx, v3(+1, 0, 0));
    b->planes[4] = plane_make(mx, v3(0, +1, 0));
    b->planes[5] = plane_make(mx, v3(0, 0, +1));
}

/* AxisAlignedBox.ContainsPoint (AxisAlignedBox.cs:143-149): open box widened by 1e-10 */
static inline int box_contains(const box_t* b, vec p)
{
    const double e = 1e-10;
    return b->min.x - e < p.x && p.x < b->max.x + e &&
           b->min.y - e < p.y && p.y < b->max.y + e &&
           b->min.z - e < p.z && p.z < b->max.z + e;
}

/* AxisAlignedBox.IntersectLineSegment (AxisAlignedBox.cs:111-141) */
static int box_intersect_segment(const box_t* b, vec start, vec end, vec* pos)
{
    double closest = DBL_MAX;
    for (int i = 0; i < 6; i++) {
        double lf; vec p;
        if (plane_intersect_segment(&b->planes[i], start, end, &lf, &p) && lf < closest) {
            if (box_contains(b, p)) { closest = lf; *pos = p; }
        }
    }
    return closest != DBL_MAX;
}

/* AxisAlignedBox.ClipLineSegment (AxisAlignedBox.cs:175-216) */
static int box_clip_segment(const box_t* b, vec* start, vec* end)
{
    int start_inside = box_contains(b, *start);
    int end_inside = box_contains(b, *end);
    if (start_inside && end_inside) return 1;
    vec ipos;
    if (!box_intersect_segment(b, *start, *end, &ipos)) return 0;
    if (start_inside) { *end = ipos; return 1; }
    vec original_start = *start;
    *start = ipos;
    if (!end_inside) {
        /* the reference assumes this second intersection exists (Contract.Assume, :211); when
         * it does not, C# would dereference null -- keep `end` unchanged instead */
        if (box_intersect_segment(b, *end, original_start, &ipos)) *end = ipos;
    }
    return 1;
}

int orc_box_contains_point(const double bmin[3], const double bmax[3], const double pt[3])
{
    box_t b; box_make(&b, vfrom(bmin), vfrom(bmax));
    return box_contains(&b, vfrom(pt));
}
int orc_box_clip_line_segment(const double bmin[3], const double bmax[3], double start[3], double end[3])
{
    box_t b; box_make(&b, vfrom(bmin), vfrom(bmax));
    vec s = vfrom(start), e = vfrom(end);
    int r = box_clip_segment(&b, &s, &e);
    vto(start, s); vto(end, e);
    return r;
}

/* ============================================================================================ */
/* SpatialSubdivision (Raytrace/SpatialSubdivision.cs)                                          */
/* ============================================================================================ */
typedef struct node_t {
    int32_t* geom; int32_t n_geom;     /* triangle ids, in original list order                    */
    box_t box;
    int has_plane; plane_t split;
    struct node_t* normal_side;
    struct node_t* back_side;
} node_t;

struct orc_tree {
    tri_t* tris; int32_t n_tris;
    node_t* root;
    int32_t depth, n_nodes, n_leaves, n_refs, largest_leaf;
};

typedef struct { uint64_t node_visits, prim_tests; } trace_counters;

static node_t* node_new(orc_tree* t, int32_t* geom, int32_t n, vec mn, vec mx)
{
    node_t* nd = (node_t*)calloc(1, sizeof *nd);
    nd->geom = geom; nd->n_geom = n;
    box_make(&nd->box, mn, mx);
    t->n_nodes++;
    return nd;
}

static void node_leaf(orc_tree* t, node_t* nd)
{
    t->n_leaves++;
    t->n_refs += nd->n_geom;
    if (nd->n_geom > t->largest_leaf) t->largest_leaf = nd->n_geom;
}

/* Point.IntersectPlane (Point.cs:35-50): >= goes to the normal side */
static inline int point_on_normal_side(vec p, const plane_t* pl)
{
    return vdot(p, pl->normal) >= pl->origin_dist;
}

/* Node.RecursivePlaneSplit (SpatialSubdivision.cs:49-230), SPLIT_LONGEST_AXIS build */
static void node_split(orc_tree* t, node_t* nd, int depth, int max_depth, int max_geom)
{
    if (depth > t->depth) t->depth = depth;
    if (depth >= max_depth || nd->n_geom <= max_geom) { node_leaf(t, nd); return; }

    vec ext = vsub(nd->box.max, nd->box.min);
    ext = v3(fabs(ext.x), fabs(ext.y), fabs(ext.z));
    int axis;
    if (ext.x > ext.y) axis = (ext.x > ext.z) ? 0 : 2;
    else               axis = (ext.y > ext.z) ? 1 : 2;

    vec split_pt = vmul(vadd(nd->box.min, nd->box.max), 0.5);   /* AxisAlignedBox.Centre (:46-52) */
    vec axis_n = axis == 0 ? v3(1, 0, 0) : axis == 1 ? v3(0, 1, 0) : v3(0, 0, 1);
    plane_t pl = plane_make(split_pt, axis_n);

    int32_t* ng = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nd->n_geom > 0 ? nd->n_geom : 1));
    int32_t* bg = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nd->n_geom > 0 ? nd->n_geom : 1));
    int32_t nn = 0, nb = 0;
    for (int32_t i = 0; i < nd->n_geom; i++) {
        const tri_t* tr = &t->tris[nd->geom[i]];
        /* Triangle.IntersectPlane (Triangle.cs:126-132): union over the three vertices */
        int a = point_on_normal_side(tr->v1, &pl), b = point_on_normal_side(tr->v2, &pl),
            c = point_on_normal_side(tr->v3, &pl);
        if (a || b || c) ng[nn++] = nd->geom[i];
        if (!a || !b || !c) bg[nb++] = nd->geom[i];
    }
    if (nn == nd->n_geom || nb == nd->n_geom) {   /* rejected split (:167-181) */
        free(ng); free(bg);
        node_leaf(t, nd);
        return;
    }
    free(nd->geom); nd->geom = NULL; nd->n_geom = 0;
    nd->has_plane = 1; nd->split = pl;

    vec back_max = nd->box.max, norm_min = nd->box.min;
    if (axis == 0) back_max.x = norm_min.x = split_pt.x;
    else if (axis == 1) back_max.y = norm_min.y = split_pt.y;
    else back_max.z = norm_min.z = split_pt.z;

    depth++;
    if (nn > 0) {
        nd->normal_side = node_new(t, ng, nn, norm_min, nd->box.max);
        node_split(t, nd->normal_side, depth, max_depth, max_geom);
    } else free(ng);
    if (nb > 0) {
        nd->back_side = node_new(t, bg, nb, nd->box.min, back_max);
        node_split(t, nd->back_side, depth, max_depth, max_geom);
    } else free(bg);
    if (!nd->normal_side && !nd->back_side) { t->n_leaves++; nd->has_plane = 0; }   /* (:207-217) */
}

static void node_free(node_t* nd)
{
    if (!nd) return;
    node_free(nd->normal_side);
    node_free(nd->back_side);
    free(nd->geom);
    free(nd);
}

static int tree_build_from_tris(tri_t* tris, int32_t n_tris, vec mn, vec mx, int max_depth,
                                int max_geom, orc_tree** out)
{
    orc_tree* t = (orc_tree*)calloc(1, sizeof *t);
    if (!t) return SOFTRAY_E_OOM;
    t->tris = tris; t->n_tris = n_tris;
    box_t bb; box_make(&bb, mn, mx);
    /* ctor check (SpatialSubdivision.cs:285-295) */
    for (int32_t i = 0; i < n_tris; i++) {
        if (!box_contains(&bb, tris[i].v1) || !box_contains(&bb, tris[i].v2) || !box_contains(&bb, tris[i].v3)) {
            free(t);
            return SOFTRAY_E_VERTEX_OUTSIDE_BBOX;
        }
    }
    int32_t* geom = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_tris > 0 ? n_tris : 1));
    for (int32_t i = 0; i < n_tris; i++) geom[i] = i;
    t->root = node_new(t, geom, n_tris, mn, mx);
    node_split(t, t->root, 1, max_depth, max_geom);
    *out = t;
    return SOFTRAY_OK;
}

int orc_tree_build(const double* tri_verts, const uint32_t* colors, int32_t n_tris,
                   const double bmin[3], const double bmax[3],
                   int32_t max_tree_depth, int32_t max_geometry_per_node, orc_tree** out)
{
    if (!out || n_tris < 0 || (n_tris > 0 && !tri_verts)) return SOFTRAY_E_INVALID_ARG;
    tri_t* tris = (tri_t*)malloc(sizeof(tri_t) * (size_t)(n_tris > 0 ? n_tris : 1));
    if (!tris) return SOFTRAY_E_OOM;
    for (int32_t i = 0; i < n_tris; i++)
        tri_make(&tris[i], vfrom(tri_verts + 9 * i), vfrom(tri_verts + 9 * i + 3), vfrom(tri_verts + 9 * i + 6),
                 colors ? colors[i] : 0xffffffffu, i);
    int rc = tree_build_from_tris(tris, n_tris, vfrom(bmin), vfrom(bmax), max_tree_depth, max_geometry_per_node, out);
    if (rc != SOFTRAY_OK) free(tris);
    return rc;
}

void orc_tree_free(orc_tree* t)
{
    if (!t) return;
    node_free(t->root);
    free(t->tris);
    free(t);
}

void orc_tree_stats(const orc_tree* t, int32_t out[6])
{
    out[0] = t->depth; out[1] = t->n_nodes; out[2] = t->n_leaves; out[3] = t->n_nodes - t->n_leaves;
    out[4] = t->n_refs; out[5] = t->largest_leaf;
}

/* GetClosestIntersection (SpatialSubdivision.cs:629-676).  The per-ray HashSet of tested
 * triangles (:410,636,655) never changes a result: a triangle enters it only when it becomes the
 * leaf's closest hit, and a leaf with a hit ends the traversal (:518-525). */
static int leaf_closest(const orc_tree* t, const node_t* nd, vec start, vec dir, orc_hit* out, trace_counters* c)
{
    double closest = DBL_MAX;
    for (int32_t i = 0; i < nd->n_geom; i++) {
        orc_hit h;
        int hit = tri_intersect_ray(&t->tris[nd->geom[i]], start, dir, &h);
        if (hit && h.ray_frac < closest) {
            if (box_contains(&nd->box, vfrom(h.pos))) { closest = h.ray_frac; *out = h; }
        }
        c->prim_tests++;
    }
    return closest < DBL_MAX;
}

/* RecursiveRayTrace (SpatialSubdivision.cs:458-627) */
static int tree_trace(const orc_tree* t, const node_t* nd, vec start, vec end, vec dir, orc_hit* out, trace_counters* c)
{
    if (!nd) return 0;
    c->node_visits++;
    if (!nd->normal_side && !nd->back_side) {
        return leaf_closest(t, nd, start, dir, out, c);
    }
    int start_normal = point_on_normal_side(start, &nd->split);
    int end_normal = point_on_normal_side(end, &nd->split);
    if (start_normal) {
        if (tree_trace(t, nd->normal_side, start, end, dir, out, c)) return 1;
        if (!end_normal) return tree_trace(t, nd->back_side, start, end, dir, out, c);
    } else {
        if (tree_trace(t, nd->back_side, start, end, dir, out, c)) return 1;
        if (end_normal) return tree_trace(t, nd->normal_side, start, end, dir, out, c);
    }
    return 0;
}

/* SpatialSubdivision.IntersectRay (SpatialSubdivision.cs:381-419) */
static int tree_intersect(const orc_tree* t, vec start, vec dir, orc_hit* out, trace_counters* c)
{
    vec end = vadd(start, vmul(dir, 10000));
    vec original_start = start;
    if (!box_clip_segment(&t->root->box, &start, &end)) return 0;
    double ray_frac_offset = vlen(vsub(original_start, start)) / vlen(dir);   /* Distance()/Length */
    if (tree_trace(t, t->root, start, end, dir, out, c)) {
        out->ray_frac += ray_frac_offset;
        return 1;
    }
    return 0;
}

/* GeometryCollection.IntersectRay over triangles only (GeometryCollection.cs:44-69) */
static int brute_intersect(const orc_tree* t, vec start, vec dir, orc_hit* out, trace_counters* c)
{
    double closest = DBL_MAX;
    for (int32_t i = 0; i < t->n_tris; i++) {
        orc_hit h;
        if (tri_intersect_ray(&t->tris[i], start, dir, &h) && h.ray_frac < closest) { closest = h.ray_frac; *out = h; }
        c->prim_tests++;
    }
    return closest != DBL_MAX;
}

int orc_tree_intersect(const orc_tree* t, const double start[3], const double dir[3], orc_hit* out)
{
    trace_counters c = {0, 0};
    return tree_intersect(t, vfrom(start), vfrom(dir), out, &c);
}
int orc_tree_brute_intersect(const orc_tree* t, const double start[3], const double dir[3], orc_hit* out)
{
    trace_counters c = {0, 0};
    return brute_intersect(t, vfrom(start), vfrom(dir), out, &c);
}

/* ============================================================================================ */
/* Model: 3DS loader (3dsLoader/ThreeDSFile.cs) + Model.Load3ds / PostProcessGeometry (Model.cs) */
/* ============================================================================================ */
typedef struct { const uint8_t* p; size_t n; size_t pos; int err; } rd_t;

static uint8_t rd_u8(rd_t* r) { if (r->pos + 1 > r->n) { r->err = 1; return 0; } return r->p[r->pos++]; }
static uint16_t rd_u16(rd_t* r)
{
    if (r->pos + 2 > r->n) { r->err = 1; r->pos = r->n; return 0; }
    uint16_t v = (uint16_t)(r->p[r->pos] | (r->p[r->pos + 1] << 8)); r->pos += 2; return v;
}
static uint32_t rd_u32(rd_t* r)
{
    if (r->pos + 4 > r->n) { r->err = 1; r->pos = r->n; return 0; }
    uint32_t v = (uint32_t)r->p[r->pos] | ((uint32_t)r->p[r->pos + 1] << 8) | ((uint32_t)r->p[r->pos + 2] << 16) |
                 ((uint32_t)r->p[r->pos + 3] << 24);
    r->pos += 4; return v;
}
static float rd_f32(rd_t* r) { uint32_t u = rd_u32(r); float f; memcpy(&f, &u, 4); return f; }

typedef struct { uint16_t id; uint32_t length; size_t start; int64_t bytes_read; } chunk_t;

/* ThreeDSChunk ctor (ThreeDSFile.cs:664-690) */
static chunk_t chunk_open(rd_t* r)
{
    chunk_t c;
    c.start = r->pos;
    c.id = rd_u16(r);
    c.length = rd_u32(r);
    c.bytes_read = 6;
    /* a chunk shorter than its own header makes the reference's loop spin forever
     * (SkipChunk seeks backwards); the oracle reports a format error instead */
    if (c.length < 6) r->err = 1;
    return c;
}
static void chunk_skip_to_end(rd_t* r, const chunk_t* c) { r->pos = c->start + c->length; if (r->pos > r->n) { r->pos = r->n; } }
/* SkipChunk (ThreeDSFile.cs:575-588) */
static void chunk_skip(rd_t* r, chunk_t* c)
{
    int64_t len = (int64_t)c->length - c->bytes_read;
    int64_t np = (int64_t)r->pos + len;
    if (np < 0) np = 0;
    if ((size_t)np > r->n) { r->err = 1; np = (int64_t)r->n; }
    r->pos = (size_t)np;
    c->bytes_read += len;
}

typedef struct { char name[256]; float diffuse[3]; } mat_t;
typedef struct { int32_t v[3]; int32_t mat; /* -1 = Triangle.defaultMaterial */ } face_t;
typedef struct {
    double* verts; int32_t n_verts; int has_verts;
    face_t* faces; int32_t n_faces; int has_faces;
} entity_t;
typedef struct {
    mat_t* mats; int32_t n_mats, cap_mats;
    entity_t* ents; int32_t n_ents, cap_ents;
} loader_t;

/* ProcessString (ThreeDSFile.cs:590-606) */
static void read_cstr(rd_t* r, chunk_t* c, char* dst, size_t cap)
{
    size_t k = 0; int idx = 0;
    uint8_t b = rd_u8(r);
    while (b != 0 && !r->err) {
        if (k + 1 < cap) dst[k++] = (char)b;
        b = rd_u8(r);
        idx++;
    }
    dst[k] = 0;
    c->bytes_read += idx + 1;
}

/* ProcessColorChunk (ThreeDSFile.cs:420-450): only the first colour sub-chunk is looked at */
static void read_color(rd_t* r, chunk_t* c, float rgb[3])
{
    chunk_t ch = chunk_open(r);
    rgb[0] = rgb[1] = rgb[2] = 1.0f;
    if (ch.id == 0x0010) { rgb[0] = rd_f32(r); rgb[1] = rd_f32(r); rgb[2] = rd_f32(r); }
    else if (ch.id == 0x0011) {
        rgb[0] = (float)rd_u8(r) / 255.0f; rgb[1] = (float)rd_u8(r) / 255.0f; rgb[2] = (float)rd_u8(r) / 255.0f;
    }
    c->bytes_read += (int64_t)ch.length;
    chunk_skip_to_end(r, &ch);
}

/* ProcessPercentageChunk (ThreeDSFile.cs:452-460) */
static void read_percentage(rd_t* r, chunk_t* c)
{
    chunk_t ch = chunk_open(r);
    (void)rd_u16(r);
    ch.bytes_read += 2;
    c->bytes_read += ch.bytes_read;
    chunk_skip_to_end(r, &ch);
}

/* ProcessTexMapChunk (ThreeDSFile.cs:329-418): nothing of it reaches the raytracer */
static void read_texmap(rd_t* r, chunk_t* c)
{
    while (c->bytes_read < (int64_t)c->length && !r->err) {
        chunk_t ch = chunk_open(r);
        if (ch.id == 0xA300) { char tmp[256]; read_cstr(r, &ch, tmp, sizeof tmp); }
        else chunk_skip(r, &ch);
        c->bytes_read += ch.bytes_read;
        chunk_skip_to_end(r, &ch);
    }
}

/* ProcessMaterialChunk (ThreeDSFile.cs:259-327) */
static void read_material(rd_t* r, loader_t* L, chunk_t* c)
{
    mat_t m; float tmp[3];
    m.name[0] = 0;
    m.diffuse[0] = m.diffuse[1] = m.diffuse[2] = 0.0f;   /* Material.cs:33 */
    while (c->bytes_read < (int64_t)c->length && !r->err) {
        chunk_t ch = chunk_open(r);
        switch (ch.id) {
        case 0xA000: read_cstr(r, &ch, m.name, sizeof m.name); break;
        case 0xA010: read_color(r, &ch, tmp); break;
        case 0xA020: read_color(r, &ch, m.diffuse); break;
        case 0xA030: read_color(r, &ch, tmp); break;
        case 0xA040: read_percentage(r, &ch); break;
        case 0xA200: read_percentage(r, &ch); read_texmap(r, &ch); break;
        default: chunk_skip(r, &ch); break;
        }
        c->bytes_read += ch.bytes_read;
        chunk_skip_to_end(r, &ch);
    }
    for (int32_t i = 0; i < L->n_mats; i++)
        if (strcmp(L->mats[i].name, m.name) == 0) return;   /* duplicate names ignored (:323-326) */
    if (L->n_mats == L->cap_mats) {
        L->cap_mats = L->cap_mats ? L->cap_mats * 2 : 8;
        L->mats = (mat_t*)realloc(L->mats, sizeof(mat_t) * (size_t)L->cap_mats);
    }
    L->mats[L->n_mats++] = m;
}

/* ProcessFaceChunk (ThreeDSFile.cs:522-573) */
static void read_face_materials(rd_t* r, loader_t* L, chunk_t* c, entity_t* e)
{
    while (c->bytes_read < (int64_t)c->length && !r->err) {
        chunk_t ch = chunk_open(r);
        if (ch.id == 0x4130) {
            char name[256];
            read_cstr(r, &ch, name, sizeof name);
            int32_t mat = -1;
            for (int32_t i = 0; i < L->n_mats; i++) if (strcmp(L->mats[i].name, name) == 0) { mat = i; break; }
            int nfaces = rd_u16(r);
            ch.bytes_read += 2;
            for (int i = 0; i < nfaces; i++) {
                int fi = rd_u16(r);
                if (fi < e->n_faces) e->faces[fi].mat = mat; else r->err = 1;   /* C#: IndexOutOfRange */
                ch.bytes_read += 2;
            }
            chunk_skip(r, &ch);
        } else chunk_skip(r, &ch);
        c->bytes_read += ch.bytes_read;
        chunk_skip_to_end(r, &ch);
    }
}

/* ProcessObjectChunk (ThreeDSFile.cs:462-520), ReadVertices (:608-632), ReadTriangles (:634-657) */
static void read_object(rd_t* r, loader_t* L, chunk_t* c, entity_t* e)
{
    while (c->bytes_read < (int64_t)c->length && !r->err) {
        chunk_t ch = chunk_open(r);
        switch (ch.id) {
        case 0x4100: read_object(r, L, &ch, e); break;
        case 0x4110: {
            int n = rd_u16(r);
            ch.bytes_read += 2;
            free(e->verts);
            e->verts = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
            e->n_verts = n; e->has_verts = 1;
            for (int i = 0; i < n; i++) {
                float f1 = rd_f32(r), f2 = rd_f32(r), f3 = rd_f32(r);
                e->verts[3 * i + 0] = f1; e->verts[3 * i + 1] = f3; e->verts[3 * i + 2] = -f2;   /* (:627) */
            }
            ch.bytes_read += (int64_t)n * 12;
            break;
        }
        case 0x4120: {
            int n = rd_u16(r);
            ch.bytes_read += 2;
            free(e->faces);
            e->faces = (face_t*)malloc(sizeof(face_t) * (size_t)(n > 0 ? n : 1));
            e->n_faces = n; e->has_faces = 1;
            for (int i = 0; i < n; i++) {
                e->faces[i].v[0] = rd_u16(r); e->faces[i].v[1] = rd_u16(r); e->faces[i].v[2] = rd_u16(r);
                e->faces[i].mat = -1;
                (void)rd_u16(r);   /* face flags */
            }
            ch.bytes_read += (int64_t)n * 8;
            if (ch.bytes_read < (int64_t)ch.length) read_face_materials(r, L, &ch, e);
            break;
        }
        case 0x4140: {
            int n = rd_u16(r);
            ch.bytes_read += 2;
            for (int i = 0; i < n; i++) { (void)rd_f32(r); (void)rd_f32(r); }
            ch.bytes_read += (int64_t)n * 8;
            break;
        }
        default: chunk_skip(r, &ch); break;
        }
        c->bytes_read += ch.bytes_read;
        chunk_skip_to_end(r, &ch);
    }
}

/* ProcessChunk (ThreeDSFile.cs:187-257) */
static void read_chunks(rd_t* r, loader_t* L, chunk_t* c)
{
    while (c->bytes_read < (int64_t)c->length && !r->err) {
        chunk_t ch = chunk_open(r);
        switch (ch.id) {
        case 0x0002: (void)rd_u32(r); ch.bytes_read += 4; break;
        case 0x3D3D: {
            chunk_t blind = chunk_open(r);      /* first sub-chunk skipped unseen (:208-216) */
            chunk_skip(r, &blind);
            ch.bytes_read += blind.bytes_read;
            read_chunks(r, L, &ch);
            break;
        }
        case 0xAFFF: read_material(r, L, &ch); break;
        case 0x4000: {
            char name[256];
            read_cstr(r, &ch, name, sizeof name);
            entity_t e; memset(&e, 0, sizeof e);
            read_object(r, L, &ch, &e);
            if (e.has_verts && e.has_faces) {
                if (L->n_ents == L->cap_ents) {
                    L->cap_ents = L->cap_ents ? L->cap_ents * 2 : 4;
                    L->ents = (entity_t*)realloc(L->ents, sizeof(entity_t) * (size_t)L->cap_ents);
                }
                L->ents[L->n_ents++] = e;
            } else { free(e.verts); free(e.faces); }
            break;
        }
        default: chunk_skip(r, &ch); break;
        }
        c->bytes_read += ch.bytes_read;
        if (ch.id != 0x0002) chunk_skip_to_end(r, &ch);
    }
}

/* Surface.PackColorAndAlpha (Surface.cs:131-138) on a float material (Model.cs:98-100) */
static uint32_t pack_color_and_alpha(const float rgb[3])
{
    uint8_t r = to_byte((double)rgb[0] * 255.0), g = to_byte((double)rgb[1] * 255.0), b = to_byte((double)rgb[2] * 255.0);
    uint8_t a = to_byte(1.0 * 255.0);
    return ((uint32_t)a << 24) + ((uint32_t)r << 16) + ((uint32_t)g << 8) + b;
}

/* Model.PostProcessGeometry (Model.cs:750-790): fit into the unit cube, centred on the origin */
static void model_post_process(orc_model* m)
{
    vec mn = vfrom(m->bbox_min), mx = vfrom(m->bbox_max);
    vec centre = v3((mn.x + mx.x) / 2, (mn.y + mx.y) / 2, (mn.z + mx.z) / 2);
    vec extent = v3(mx.x - mn.x, mx.y - mn.y, mx.z - mn.z);
    double scale = 1.0 / fmax(fmax(extent.x, extent.y), extent.z);
    for (int32_t i = 0; i < m->n_verts; i++) {
        vec p = vmul(vsub(vfrom(m->verts_xyz + 3 * i), centre), scale);
        vto(m->verts_xyz + 3 * i, p);
    }
    vto(m->bbox_min, vmul(vsub(mn, centre), scale));
    vto(m->bbox_max, vmul(vsub(mx, centre), scale));
}

/* Model.CalcExtent (Model.cs:738-748) */
static void model_calc_extent(orc_model* m)
{
    vec mn = v3(DBL_MAX, DBL_MAX, DBL_MAX), mx = v3(-DBL_MAX, -DBL_MAX, -DBL_MAX);
    for (int32_t i = 0; i < m->n_verts; i++) {
        vec p = vfrom(m->verts_xyz + 3 * i);
        mn = v3(fmin(mn.x, p.x), fmin(mn.y, p.y), fmin(mn.z, p.z));
        mx = v3(fmax(mx.x, p.x), fmax(mx.y, p.y), fmax(mx.z, p.z));
    }
    vto(m->bbox_min, mn); vto(m->bbox_max, mx);
}

void orc_model_free(orc_model* m)
{
    if (!m) return;
    free(m->verts_xyz); free(m->tri_vidx); free(m->tri_argb); free(m);
}

/* Model.Load3ds (Model.cs:522-653) */
int orc_model_load_3ds(const uint8_t* bytes, size_t n_bytes, orc_model** out)
{
    if (!bytes || !out) return SOFTRAY_E_INVALID_ARG;
    rd_t r = { bytes, n_bytes, 0, 0 };
    loader_t L; memset(&L, 0, sizeof L);
    int rc = SOFTRAY_OK;
    chunk_t top = chunk_open(&r);
    if (r.err || top.id != 0x4D4D) { rc = SOFTRAY_E_FORMAT; goto done; }   /* "Not a proper 3DS file." */
    read_chunks(&r, &L, &top);
    if (r.err) { rc = SOFTRAY_E_FORMAT; goto done; }                       /* EndOfStreamException */
    if (L.n_ents == 0) { rc = SOFTRAY_E_FORMAT; goto done; }               /* "No entities in model" */
    {
        int32_t nv = 0, nt = 0;
        for (int32_t e = 0; e < L.n_ents; e++) {
            if (L.ents[e].n_verts < 3 || L.ents[e].n_faces == 0) { rc = SOFTRAY_E_FORMAT; goto done; }
            nv += L.ents[e].n_verts; nt += L.ents[e].n_faces;
        }
        orc_model* m = (orc_model*)calloc(1, sizeof *m);
        m->verts_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)nv);
        m->tri_vidx = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)nt);
        m->tri_argb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nt);
        m->n_verts = nv; m->n_tris = nt;
        vec mn = v3(DBL_MAX, DBL_MAX, DBL_MAX), mx = v3(-DBL_MAX, -DBL_MAX, -DBL_MAX);   /* double.MinValue */
        int32_t vo = 0, to = 0;
        const float default_diffuse[3] = {0.0f, 0.0f, 0.0f};
        for (int32_t e = 0; e < L.n_ents; e++) {
            const entity_t* en = &L.ents[e];
            for (int32_t i = 0; i < en->n_verts; i++) {
                double c[3];
                for (int k = 0; k < 3; k++) {
                    double x = en->verts[3 * i + k];
                    if (isnan(x) || isinf(x) || fabs(x) > 1e6) x = 0.0;   /* maxCoordinateSize (:208,593) */
                    c[k] = x;
                }
                memcpy(m->verts_xyz + 3 * (size_t)(vo + i), c, sizeof c);
                mn = v3(fmin(mn.x, c[0]), fmin(mn.y, c[1]), fmin(mn.z, c[2]));
                mx = v3(fmax(mx.x, c[0]), fmax(mx.y, c[1]), fmax(mx.z, c[2]));
            }
            for (int32_t i = 0; i < en->n_faces; i++) {
                for (int k = 0; k < 3; k++) m->tri_vidx[3 * (size_t)(to + i) + k] = vo + en->faces[i].v[k];
                const float* d = en->faces[i].mat >= 0 ? L.mats[en->faces[i].mat].diffuse : default_diffuse;
                m->tri_argb[to + i] = pack_color_and_alpha(d);   /* Renderer.cs:1463 */
            }
            vo += en->n_verts; to += en->n_faces;
        }
        /* C# would throw on an out-of-range vertex index later; report it as a format error */
        for (int32_t i = 0; i < 3 * nt; i++)
            if (m->tri_vidx[i] < 0 || m->tri_vidx[i] >= nv) { orc_model_free(m); rc = SOFTRAY_E_FORMAT; goto done; }
        vto(m->bbox_min, mn); vto(m->bbox_max, mx);
        model_post_process(m);
        *out = m;
    }
done:
    for (int32_t e = 0; e < L.n_ents; e++) { free(L.ents[e].verts); free(L.ents[e].faces); }
    free(L.ents); free(L.mats);
    return rc;
}

int orc_model_from_arrays(const double* verts_xyz, int32_t n_verts, const int32_t* tri_vidx,
                          const uint32_t* tri_argb, int32_t n_tris, int32_t normalise, orc_model** out)
{
    if (!verts_xyz || !tri_vidx || !out || n_verts < 0 || n_tris < 0) return SOFTRAY_E_INVALID_ARG;
    orc_model* m = (orc_model*)calloc(1, sizeof *m);
    m->verts_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n_verts > 0 ? n_verts : 1));
    m->tri_vidx = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(n_tris > 0 ? n_tris : 1));
    m->tri_argb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n_tris > 0 ? n_tris : 1));
    memcpy(m->verts_xyz, verts_xyz, sizeof(double) * 3 * (size_t)n_verts);
    memcpy(m->tri_vidx, tri_vidx, sizeof(int32_t) * 3 * (size_t)n_tris);
    for (int32_t i = 0; i < n_tris; i++) m->tri_argb[i] = tri_argb ? tri_argb[i] : 0xffffffffu;
    m->n_verts = n_verts; m->n_tris = n_tris;
    model_calc_extent(m);
    if (normalise) model_post_process(m);
    *out = m;
    return SOFTRAY_OK;
}

/* ============================================================================================ */
/* Scene                                                                                         */
/* ============================================================================================ */
struct orc_scene {
    int32_t n_meshes;
    orc_tree** trees;          /* per mesh: triangles + SpatialSubdivision                         */
    int32_t n_spheres;
    sphere_t* spheres;
};

void orc_options_defaults(orc_options* o)
{
    o->concurrency = 4;
    o->n_threads = 0;
    o->tree_max_depth = 15;
    o->tree_max_per_node = 25;
    o->path_tracing = 0;
    o->_pad = 0;
    o->col_start = 0; o->col_end = -1;
}

void orc_scene_free(orc_scene* s)
{
    if (!s) return;
    for (int32_t i = 0; i < s->n_meshes; i++) orc_tree_free(s->trees[i]);
    free(s->trees); free(s->spheres); free(s);
}

/* MakeRayTracableGeometry_simple/_subdivided (Renderer.cs:1452-1494) */
int orc_scene_create(const softray_scene_desc* desc, const orc_options* opt, orc_scene** out)
{
    orc_options o;
    if (opt) o = *opt; else orc_options_defaults(&o);
    if (!desc || !out || desc->n_meshes < 0 || desc->n_spheres < 0) return SOFTRAY_E_INVALID_ARG;
    if ((desc->n_meshes > 0 && !desc->meshes) || (desc->n_spheres > 0 && !desc->spheres)) return SOFTRAY_E_INVALID_ARG;
    orc_scene* s = (orc_scene*)calloc(1, sizeof *s);
    s->n_meshes = desc->n_meshes;
    s->trees = (orc_tree**)calloc((size_t)(desc->n_meshes > 0 ? desc->n_meshes : 1), sizeof(orc_tree*));
    for (int32_t mi = 0; mi < desc->n_meshes; mi++) {
        const softray_mesh* m = &desc->meshes[mi];
        if (m->n_tris < 0 || m->n_verts < 0 || (m->n_tris > 0 && (!m->verts_xyz || !m->tri_vidx || !m->tri_argb))) {
            orc_scene_free(s); return SOFTRAY_E_INVALID_ARG;
        }
        tri_t* tris = (tri_t*)malloc(sizeof(tri_t) * (size_t)(m->n_tris > 0 ? m->n_tris : 1));
        for (int32_t i = 0; i < m->n_tris; i++) {
            int32_t a = m->tri_vidx[3 * i], b = m->tri_vidx[3 * i + 1], c = m->tri_vidx[3 * i + 2];
            if (a < 0 || b < 0 || c < 0 || a >= m->n_verts || b >= m->n_verts || c >= m->n_verts) {
                free(tris); orc_scene_free(s); return SOFTRAY_E_INVALID_ARG;
            }
            tri_make(&tris[i], vfrom(m->verts_xyz + 3 * a), vfrom(m->verts_xyz + 3 * b), vfrom(m->verts_xyz + 3 * c),
                     m->tri_argb[i], i);
        }
        int rc = tree_build_from_tris(tris, m->n_tris, vfrom(m->bbox_min), vfrom(m->bbox_max),
                                      o.tree_max_depth, o.tree_max_per_node, &s->trees[mi]);
        if (rc != SOFTRAY_OK) { free(tris); orc_scene_free(s); return rc; }
    }
    s->n_spheres = desc->n_spheres;
    s->spheres = (sphere_t*)malloc(sizeof(sphere_t) * (size_t)(desc->n_spheres > 0 ? desc->n_spheres : 1));
    for (int32_t i = 0; i < desc->n_spheres; i++) {
        const softray_sphere* sp = &desc->spheres[i];
        s->spheres[i].center = v3(sp->cx, sp->cy, sp->cz);
        s->spheres[i].radius = sp->r;
        s->spheres[i].radius_sqr = sp->r * sp->r;     /* Sphere.cs:30 */
        s->spheres[i].color = sp->argb;
    }
    *out = s;
    return SOFTRAY_OK;
}

void orc_scene_tree_stats(const orc_scene* s, int32_t mesh, int32_t out[6])
{
    orc_tree_stats(s->trees[mesh], out);
}

/* ============================================================================================ */
/* Extensions shared by oracle and product: Texture3D id 1, mirror blend             */
/* ============================================================================================ */
/* Index quantisation of Texture3DCache.Sample (Texture3DCache.cs:98-100) with N = 128, clamped;
 * the texel is a closed-form integer pattern (procedural "sampleGenerator", no cache). */
uint8_t orc_texture3d_sample(int32_t id, const double pos[3])
{
    if (id != 1) return 255;
    const int N = 128;
    int q[3];
    for (int k = 0; k < 3; k++) {
        int v = (int)((pos[k] + 0.5) * (N - 1));
        if (v < 0) v = 0;
        if (v > N - 1) v = N - 1;
        q[k] = v;
    }
    int cell = ((q[0] >> 3) ^ (q[1] >> 3) ^ (q[2] >> 3)) & 1;
    int grain = (q[0] * 3 + q[1] * 5 + q[2] * 7) & 31;
    return (uint8_t)(255 - cell * 80 - grain);
}

static inline uint32_t mirror_blend(uint32_t local, uint32_t refl)
{
    uint32_t r = (3 * ((local >> 16) & 0xff) + ((refl >> 16) & 0xff)) >> 2;
    uint32_t g = (3 * ((local >> 8) & 0xff) + ((refl >> 8) & 0xff)) >> 2;
    uint32_t b = (3 * (local & 0xff) + (refl & 0xff)) >> 2;
    return (255u << 24) + (r << 16) + (g << 8) + b;
}

/* ============================================================================================ */
/* The decorator chain                                                                           */
/* ============================================================================================ */
typedef struct {
    uint64_t rays_primary, rays_shadow, rays_secondary, node_visits, prim_tests, hits_primary;
} counters_t;

typedef struct {
    const softray_instance* inst;
    const orc_tree* tree;
    vec light_dir_model;      /* Renderer.cs:1513 */
    vec light_pos_model;      /* Renderer.cs:1515 + Instance.cs:192-203 (un-projection discarded) */
    vec start_world;          /* Renderer.cs:1717 */
    int32_t tri_base;         /* flattened hit-id base for composite frames */
} inst_ctx;

typedef struct {
    const orc_scene* s;
    const softray_frame* f;
    const orc_options* o;
    const inst_ctx* ic; int32_t n_ic;
    const vec* light_offsets;
    orc_random* rng;
    counters_t* c;
} rctx;

/* rootGeometry (Renderer.cs:1534-1549): [ExtraGeometry..., tree | simple list], strict < keeps the
 * earliest entry on ties (GeometryCollection.cs:53). */
static int geometry_intersect(const rctx* x, const inst_ctx* ic, vec start, vec dir, orc_hit* out)
{
    trace_counters tc = {0, 0};
    double closest = DBL_MAX;
    int found = 0;
    if (x->s->n_spheres > 0) {
        for (int32_t i = 0; i < x->s->n_spheres; i++) {
            orc_hit h;
            if (sphere_intersect_ray(&x->s->spheres[i], i, start, dir, &h) && h.ray_frac < closest) {
                closest = h.ray_frac; *out = h; found = 1;
            }
            tc.prim_tests++;
        }
    }
    if (ic->tree) {
        orc_hit h;
        int hit = x->f->subdivision ? tree_intersect(ic->tree, start, dir, &h, &tc)
                                    : brute_intersect(ic->tree, start, dir, &h, &tc);
        if (hit && h.ray_frac < closest) { closest = h.ray_frac; *out = h; found = 1; }
    }
    x->c->node_visits += tc.node_visits;
    x->c->prim_tests += tc.prim_tests;
    return found;
}

/* Instance.TransformPosToView (Instance.cs:168-184): post-projection "view space" */
static vec pos_to_view(const softray_instance* in, double fov, vec pos)
{
    vec v = mul3x4(in->M, pos);
    v.x = v.x / v.z * fov;
    v.y = v.y / v.z * fov;
    v.z = (v.z - in->pos[2] + 1.0) * 0.5;
    return v;
}

/* ShadingMethod.CalcLighting + CalcLightingIntensity (ShadingMethod.cs:90-177), white material */
static double calc_lighting_intensity(const softray_frame* f, vec point, vec normal)
{
    vec to_light;
    if (f->point_lighting) to_light = vnormalise(vsub(vfrom(f->light_pos_view), point));
    else to_light = vneg(vfrom(f->light_dir_view));
    double diffuse = vdot(to_light, normal);
    diffuse = fmax(0.0, diffuse);
    double specular = 0.0;
    if (f->specular_lighting) {
        vec to_cam = vnormalise(vneg(point));
        vec refl = vsub(vmul(normal, 2.0 * vdot(to_light, normal)), to_light);
        double cos_a = vdot(refl, to_cam);
        specular = pow(cos_a, f->shininess);      /* negative base, even exponent => positive */
        specular = fmax(0.0, specular);
    }
    /* Color = white*ambient + white*diffuse + white*specular, each channel (1.0*a + 1.0*d) + 1.0*s */
    double ch = 1.0 * f->ambient + 1.0 * diffuse + 1.0 * specular;
    ch = fmin(ch, 1.0);
    return fmax(fmax(ch, ch), ch);
}

/* geometry -> [Texture3D ext] -> ShadingMethod.IntersectRay (ShadingMethod.cs:36-68) */
static int shading_intersect(const rctx* x, const inst_ctx* ic, vec start, vec dir, orc_hit* out)
{
    if (!geometry_intersect(x, ic, start, dir, out)) return 0;
    if (x->f->texture3d_id) out->color = modulate_packed(out->color, orc_texture3d_sample(x->f->texture3d_id, out->pos));
    if (!x->f->shading) return 1;
    vec pos_view = pos_to_view(ic->inst, x->f->fov_depth, vfrom(out->pos));
    vec normal_view = mul3x3(ic->inst->M, vfrom(out->normal));
    double intensity = calc_lighting_intensity(x->f, pos_view, normal_view);
    uint8_t b = to_byte(255 * intensity);
    out->color = modulate_packed(out->color, b);
    return 1;
}

/* Color(uint) (Color.cs:31-36) / ToARGB (:105-111) for path tracing */
typedef struct { double r, g, b; } col;
static col col_from_argb(uint32_t c)
{
    col k = { (uint8_t)((c >> 16) & 0xff) / 255.0, (uint8_t)((c >> 8) & 0xff) / 255.0, (uint8_t)(c & 0xff) / 255.0 };
    return k;
}

/* PathTracingMethod.IntersectRay (PathTracingMethod.cs:36-101) -- oracle-only decorator */
static int pathtrace_intersect(const rctx* x, const inst_ctx* ic, vec start, vec dir, orc_hit* out)
{
    if (!shading_intersect(x, ic, start, dir, out)) return 0;
    if (!x->o->path_tracing) return 1;
    vec n = vfrom(out->normal);
    vec new_start = vadd(vfrom(out->pos), vmul(n, 0.001));
    /* C# evaluates constructor arguments left to right; C does not promise that */
    double rx = orc_random_next_double(x->rng) * 2 - 1;
    double ry = orc_random_next_double(x->rng) * 2 - 1;
    double rz = orc_random_next_double(x->rng) * 2 - 1;
    vec rd = v3(rx, ry, rz);
    if (vdot(rd, n) < 0.0) rd = vneg(rd);
    rd = vnormalise(rd);
    orc_hit h2;
    col incoming = {0.0, 0.0, 0.0};
    x->c->rays_secondary++;
    if (shading_intersect(x, ic, new_start, rd, &h2)) incoming = col_from_argb(h2.color);
    col emission = col_from_argb(out->color);
    double frac = vdot(n, rd);
    col o = { incoming.r * frac + emission.r, incoming.g * frac + emission.g, incoming.b * frac + emission.b };
    if (o.r > 1.0 || o.g > 1.0 || o.b > 1.0) {
        double inv = 1.0 / sqrt(o.r * o.r + o.g * o.g + o.b * o.b);
        o.r *= inv; o.g *= inv; o.b *= inv;
    }
    out->color = (255u << 24) + ((uint32_t)to_byte(o.r * 255.0) << 16) + ((uint32_t)to_byte(o.g * 255.0) << 8) +
                 to_byte(o.b * 255.0);
    return 1;
}

/* ShadowMethod.TraceRaysForSoftShadows (ShadowMethod.cs:144-180) */
static double soft_shadow_fraction(const rctx* x, const inst_ctx* ic, vec surface_pos, vec surface_normal)
{
    int escaped = 0, q = x->f->shadow_samples;
    for (int i = 0; i < q; i++) {
        vec end = vadd(surface_pos, vmul(surface_normal, 0.001));
        vec dir, start;
        if (x->f->point_lighting) {
            vec light = vadd(ic->light_pos_model, x->light_offsets[i]);
            dir = vsub(end, light);
            start = light;
        } else {
            dir = ic->light_dir_model;
            start = vadd(vadd(end, vmul(dir, 1000.0)), x->light_offsets[i]);
        }
        orc_hit sh;
        x->c->rays_shadow++;
        if (!pathtrace_intersect(x, ic, start, dir, &sh) || sh.ray_frac > 1.0) escaped++;
    }
    return (double)escaped / (double)q;
}

/* ShadowMethod.IntersectRay (ShadowMethod.cs:93-121), dynamic mode */
static int shadow_intersect(const rctx* x, const inst_ctx* ic, vec start, vec dir, orc_hit* out)
{
    if (!pathtrace_intersect(x, ic, start, dir, out)) return 0;
    if (!x->f->shadows) return 1;
    uint8_t b = to_byte(soft_shadow_fraction(x, ic, vfrom(out->pos), vfrom(out->normal)) * 255);
    out->color = modulate_packed(out->color, b);
    return 1;
}

/* The whole chain for one ray plus the extensions: composite instances (nearest hit across
 * instances; ties -> lowest instance) and bounded mirror reflection. */
static int chain_intersect(const rctx* x, vec start_view_dummy, const vec* starts, const vec* dirs,
                           orc_hit* out, int32_t* which)
{
    (void)start_view_dummy;
    int found = 0; double closest = DBL_MAX;
    for (int32_t i = 0; i < x->n_ic; i++) {
        orc_hit h;
        if (x->n_ic == 1) {
            if (shadow_intersect(x, &x->ic[i], starts[i], dirs[i], &h)) { *out = h; *which = 0; return 1; }
            return 0;
        }
        if (shadow_intersect(x, &x->ic[i], starts[i], dirs[i], &h) && h.ray_frac < closest) {
            closest = h.ray_frac; *out = h; *which = i; found = 1;
        }
    }
    return found;
}

/* TraceRayComplex (Renderer.cs:1850-1879) + reflection extension.  Returns packed colour. */
static uint32_t trace_ray_complex(const rctx* x, const vec* starts, const vec* dirs, int depth,
                                  orc_hit* primary, int32_t* which, int* hit_flag)
{
    orc_hit h; int32_t w = 0;
    int hit = chain_intersect(x, v3(0, 0, 0), starts, dirs, &h, &w);
    if (hit_flag) *hit_flag = hit;
    if (!hit) return x->f->background_argb | 0xFF000000u;       /* BackgroundColorWithAlpha (:325-331) */
    if (primary) { *primary = h; *which = w; }
    uint32_t color = h.color;
    if (depth > 0 && x->n_ic == 1) {
        /* r = d - 2(d.n)n from pos + n*0.001, like PathTracingMethod's secondary ray offset */
        vec n = vfrom(h.normal), d = dirs[0];
        vec r = vsub(d, vmul(n, 2.0 * vdot(d, n)));
        vec rs = vadd(vfrom(h.pos), vmul(n, 0.001));
        x->c->rays_secondary++;
        uint32_t refl = trace_ray_complex(x, &rs, &r, depth - 1, NULL, NULL, NULL);
        color = mirror_blend(color, refl);
    }
    return color;
}

void orc_area_light_offsets(int32_t seed, int32_t n, double* out_xyz)
{
    orc_random rng; orc_random_init(&rng, seed);
    for (int32_t i = 0; i < n; i++) {
        double a = orc_random_next_double(&rng) * 2 - 1;
        double b = orc_random_next_double(&rng) * 2 - 1;
        double c = orc_random_next_double(&rng) * 2 - 1;
        vec o = vmul(vnormalise(v3(a, b, c)), 0.2);
        vto(out_xyz + 3 * i, o);
    }
}

/* One pixel of RaytraceBlock (Renderer.cs:1718-1827) */
static void render_pixel(const rctx* x, int col, int row, uint32_t* pixels, int32_t* hit_ids, const orc_aux* aux)
{
    const softray_frame* f = x->f;
    const int W = f->width, H = f->height, n = f->sub_pixel_res;
    const double aspect = (double)H / (double)W;           /* Renderer.cs:621 */
    vec starts[SOFTRAY_MAX_INSTANCES], dirs[SOFTRAY_MAX_INSTANCES];
    const int32_t ni = x->n_ic;
    orc_hit ph; int32_t which = 0; int hit = 0;
    uint32_t out_color;

    if (n == 1) {
        vec dir_view = v3(-((double)col / W - 0.5), -((double)row / H - 0.5) * aspect, f->fov_depth);
        for (int32_t i = 0; i < ni; i++) { starts[i] = x->ic[i].start_world; dirs[i] = mul3x3(x->ic[i].inst->Minv, dir_view); }
        x->c->rays_primary++;
        out_color = trace_ray_complex(x, starts, dirs, f->reflection_depth, &ph, &which, &hit);
    } else {
        int sum_r = 0, sum_g = 0, sum_b = 0;
        vec focal_pt[SOFTRAY_MAX_INSTANCES];
        if (f->focal_blur) {
            vec dir_view = v3(-((double)col / W - 0.5), -((double)row / H - 0.5) * aspect, f->fov_depth);
            for (int32_t i = 0; i < ni; i++) {
                vec dw = mul3x3(x->ic[i].inst->Minv, dir_view);
                focal_pt[i] = vadd(vmul(dw, f->focal_depth), x->ic[i].start_world);
            }
        }
        for (int sx = 0; sx < n; sx++)
            for (int sy = 0; sy < n; sy++) {
                double fx = (double)sx / (n - 1) - 0.5;
                double fy = (double)sy / (n - 1) - 0.5;
                for (int32_t i = 0; i < ni; i++) {
                    if (f->focal_blur) {
                        vec sv = v3(fx / W * f->focal_strength, fy / H * f->focal_strength, -x->ic[i].inst->pos[2]);
                        starts[i] = mul3x3(x->ic[i].inst->Minv, sv);
                        dirs[i] = vsub(focal_pt[i], starts[i]);
                    } else {
                        starts[i] = x->ic[i].start_world;
                        vec dv = v3(-((col + fx) / W - 0.5), -((row + fy) / H - 0.5) * aspect, f->fov_depth);
                        dirs[i] = mul3x3(x->ic[i].inst->Minv, dv);
                    }
                }
                x->c->rays_primary++;
                uint32_t c = trace_ray_complex(x, starts, dirs, f->reflection_depth, &ph, &which, &hit);
                sum_r += (c >> 16) & 0xff; sum_g += (c >> 8) & 0xff; sum_b += c & 0xff;
                if (hit) x->c->hits_primary++;
            }
        sum_r /= n * n; sum_g /= n * n; sum_b /= n * n;
        out_color = (255u << 24) + ((uint32_t)(uint8_t)sum_r << 16) + ((uint32_t)(uint8_t)sum_g << 8) + (uint8_t)sum_b;
    }
    if (n == 1 && hit) x->c->hits_primary++;
    size_t idx = (size_t)row * (size_t)W + (size_t)col;
    pixels[idx] = out_color;                                /* Surface.DrawPixel (Surface.cs:174-181) */
    if (hit_ids) hit_ids[idx] = hit ? (ph.prim_id >= 0 ? x->ic[which].tri_base + ph.prim_id : ph.prim_id) : -1;
    if (aux && aux->ray_frac) aux->ray_frac[idx] = hit ? ph.ray_frac : NAN;
    if (aux && aux->cos_theta) {
        vec d = dirs[which];
        aux->cos_theta[idx] = hit ? vdot(d, vfrom(ph.normal)) / vlen(d) : NAN;
    }
}

/* Row blocks run as tasks (Renderer.cs:1655-1680); here: a pool of pthreads pulling units. */
typedef struct {
    const orc_scene* s; const softray_frame* f; const orc_options* o; const inst_ctx* ic; const vec* offsets;
    uint32_t* pixels; int32_t* hit_ids; const orc_aux* aux;
    int start_row, end_row, block_h, n_units;
    int col_start, col_end, col_chunk, chunks_per_row;   /* non-path-tracing: a unit is a run of col_chunk pixels of one row */
    atomic_int next;
    pthread_mutex_t lock;
    counters_t total;
} render_job;

static void* render_worker(void* arg)
{
    render_job* j = (render_job*)arg;
    counters_t c; memset(&c, 0, sizeof c);
    for (;;) {
        int u = atomic_fetch_add(&j->next, 1);
        if (u >= j->n_units) break;
        orc_random rng; orc_random_init(&rng, j->f->random_seed);   /* per block (Renderer.cs:1693) */
        rctx x = { j->s, j->f, j->o, j->ic, j->f->n_instances, j->offsets, &rng, &c };
        if (j->chunks_per_row > 0) {
            /* no render-time RNG on this path: any partition of the pixels gives the same image */
            const int row = j->start_row + u / j->chunks_per_row;
            const int c0 = j->col_start + (u % j->chunks_per_row) * j->col_chunk;
            if (j->f->band_count > 1 && j->f->band_height > 0 &&
                ((row - j->start_row) / j->f->band_height) % j->f->band_count != j->f->band_index) continue;
            for (int col = c0; col < c0 + j->col_chunk && col <= j->col_end; col++) render_pixel(&x, col, row, j->pixels, j->hit_ids, j->aux);
            continue;
        }
        int top = j->start_row + u * j->block_h;
        /* the reference's last block runs blockHeight rows even past end_row (App. A #16); only
         * rows inside [start_row,end_row] are produced here */
        for (int row = top; row < top + j->block_h && row <= j->end_row; row++) {
            /* row-band partition (softray_frame.band_*): rows of other bands are left untouched */
            if (j->f->band_count > 1 && j->f->band_height > 0 &&
                ((row - j->start_row) / j->f->band_height) % j->f->band_count != j->f->band_index) continue;
            for (int col = 0; col < j->f->width; col++) render_pixel(&x, col, row, j->pixels, j->hit_ids, j->aux);
        }
    }
    pthread_mutex_lock(&j->lock);
    j->total.rays_primary += c.rays_primary; j->total.rays_shadow += c.rays_shadow;
    j->total.rays_secondary += c.rays_secondary; j->total.node_visits += c.node_visits;
    j->total.prim_tests += c.prim_tests; j->total.hits_primary += c.hits_primary;
    pthread_mutex_unlock(&j->lock);
    return NULL;
}

/* RaytraceGeometry (Renderer.cs:1501-1687) from after PreCalculate() */
int orc_render(const orc_scene* s, const softray_frame* f, const orc_options* opt,
               uint32_t* pixels_argb, int32_t* hit_ids, const orc_aux* aux, softray_stats* stats)
{
    orc_options o;
    if (opt) o = *opt; else orc_options_defaults(&o);
    if (!s || !f || !pixels_argb || !f->instances) return SOFTRAY_E_INVALID_ARG;
    if (f->width <= 0 || f->height <= 0 || f->sub_pixel_res < 1 || f->n_instances < 1 || f->n_instances > SOFTRAY_MAX_INSTANCES)
        return SOFTRAY_E_INVALID_ARG;
    if (f->shadows && f->shadow_samples < 1) return SOFTRAY_E_INVALID_ARG;
    if (f->reflection_depth < 0 || f->reflection_depth > 4) return SOFTRAY_E_INVALID_ARG;
    if (f->band_count > 1 && (f->band_index < 0 || f->band_index >= f->band_count)) return SOFTRAY_E_INVALID_ARG;
    if (f->n_instances > 1 && (s->n_spheres > 0 || f->shadows || (f->focal_blur && f->sub_pixel_res > 1) || f->reflection_depth
                               || o.path_tracing))
        return SOFTRAY_E_UNSUPPORTED;

    inst_ctx ic[SOFTRAY_MAX_INSTANCES];
    int32_t base = 0;
    for (int32_t i = 0; i < f->n_instances; i++) {
        const softray_instance* in = &f->instances[i];
        if (in->mesh_id < 0 || in->mesh_id >= s->n_meshes) return SOFTRAY_E_INVALID_ARG;
        ic[i].inst = in;
        ic[i].tree = s->trees[in->mesh_id];
        ic[i].light_dir_model = mul3x3(in->Minv, vfrom(f->light_dir_view));
        ic[i].light_pos_model = mul3x4(in->Minv, vfrom(f->light_pos_view));
        if (f->n_instances == 1) ic[i].start_world = mul3x3(in->Minv, v3(0, 0, -in->pos[2]));
        else ic[i].start_world = mul3x4(in->Minv, v3(0, 0, 0));   /* composite: full translation */
        ic[i].tri_base = base;
        base += ic[i].tree->n_tris;
    }

    vec* offsets = NULL;
    if (f->shadows) {
        offsets = (vec*)malloc(sizeof(vec) * (size_t)f->shadow_samples);
        orc_area_light_offsets(f->random_seed, f->shadow_samples, (double*)offsets);
    }

    /* clamp rows (Renderer.cs:1652-1653) */
    int start_row = f->start_row < 0 ? 0 : f->start_row; if (start_row > f->height - 1) start_row = f->height - 1;
    int end_row = f->end_row < 0 ? 0 : f->end_row;       if (end_row > f->height - 1) end_row = f->height - 1;
    int num_rows = end_row - start_row + 1;
    counters_t total; memset(&total, 0, sizeof total);


    if (num_rows > 0) {
        render_job job;
        memset(&job, 0, sizeof job);
        job.s = s; job.f = f; job.o = &o; job.ic = ic; job.offsets = offsets;
        job.pixels = pixels_argb; job.hit_ids = hit_ids; job.aux = aux;
        job.start_row = start_row; job.end_row = end_row;
        if (o.path_tracing) {
            /* row blocks, each with its own System.Random (Renderer.cs:1659-1670,1693) */
            int conc = o.concurrency > 0 ? o.concurrency : 1;
            job.block_h = (num_rows - 1 + conc) / conc;
            job.n_units = (num_rows - 1 + job.block_h) / job.block_h;
        } else {
            /* no render-time RNG on this path: any partition gives the same image -- runs of 32 pixels */
            job.col_start = o.col_start < 0 ? 0 : o.col_start;
            job.col_end = (o.col_end < 0 || o.col_end > f->width - 1) ? f->width - 1 : o.col_end;
            job.col_chunk = 32;
            job.chunks_per_row = job.col_end >= job.col_start ? (job.col_end - job.col_start + job.col_chunk) / job.col_chunk : 0;
            job.block_h = 1;
            job.n_units = num_rows * job.chunks_per_row;
            if (job.chunks_per_row == 0) job.n_units = 0;
        }
        atomic_init(&job.next, 0);
        pthread_mutex_init(&job.lock, NULL);
        int nthreads = o.n_threads > 0 ? o.n_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
        if (nthreads < 1) nthreads = 1;
        if (nthreads > job.n_units) nthreads = job.n_units;
        if (nthreads > 256) nthreads = 256;
        pthread_t th[256];
        int started = 0;
        for (int i = 1; i < nthreads; i++)
            if (pthread_create(&th[started], NULL, render_worker, &job) == 0) started++;
        render_worker(&job);
        for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
        pthread_mutex_destroy(&job.lock);
        total = job.total;
    }
    free(offsets);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays_primary = total.rays_primary; stats->rays_shadow = total.rays_shadow;
        stats->rays_secondary = total.rays_secondary; stats->node_visits = total.node_visits;
        stats->prim_tests = total.prim_tests; stats->hits_primary = total.hits_primary;
    }
    return SOFTRAY_OK;
}
