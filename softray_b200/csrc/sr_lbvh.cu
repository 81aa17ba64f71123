// sr_lbvh.cu -- device-side scene flatten + deterministic BVH build.
//
// Replaces, on the GPU, what softray_scene_create otherwise does on the host for one mesh:
// the Triangle / Plane constructor precompute (Triangle.cs:29-57, Plane.cs:22-29), the
// vertex-inside-bounding-box contract of the SpatialSubdivision constructor
// (SpatialSubdivision.cs:285-295) and the acceleration structure (the reference's kd-style tree,
// SpatialSubdivision.cs:49-230; here an LBVH: 63-bit Morton codes of the triangle centroids, a stable
// radix sort of (code, index) pairs, Karras' radix tree, bottom-up box fitting).  Everything is a
// function of the input only -- no atomics-ordered allocation, ties broken by triangle index -- so the
// device layout is bit-identical across runs.  The exact records are computed with the same unfused
// FP64 operations in the same order as the host path (make_tri_rec / make_tri_filt in sr_api.cu), so
// images do not depend on which builder ran.
//
// cub::DeviceRadixSort (CUDA toolkit) does the sort; every other step is a kernel below.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <float.h>
#include <math_constants.h>
#include <stdint.h>

#include "sr_types.h"

namespace sr {

namespace {

struct v3 { double x, y, z; };
__device__ __forceinline__ v3 mk3(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 sub3(v3 a, v3 b) { return mk3(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)); }
__device__ __forceinline__ double dot3(v3 a, v3 b)      // Vector.cs:99-102, left to right, never fused
{
    return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z));
}
__device__ __forceinline__ v3 cross3(v3 a, v3 b)        // Vector.cs:104-110
{
    return mk3(__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
               __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}

struct BuildParams {
    double bmin[3], bmax[3];
    float cmin[3], cinv[3];      // centroid quantisation: (c - cmin) * cinv in [0, 1]
    float pad;
    int32_t n_verts, n_tris;
};

// spread the low 21 bits of v so that there are two zero bits between consecutive bits
__device__ __forceinline__ unsigned long long spread21(unsigned int v)
{
    unsigned long long x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// Triangle ctor + Plane ctor + the two denominators (make_tri_rec), the FP32 filter record
// (make_tri_filt), FP32 bounds rounded outward, Morton key.  error: 1 = vertex index out of range,
// 2 = vertex outside the bounding box.
__global__ void __launch_bounds__(256)
tri_prepare_kernel(const double* __restrict__ verts, const int32_t* __restrict__ vidx, const uint32_t* __restrict__ argb,
                   BuildParams p, TriRec* __restrict__ recs, TriFilt* __restrict__ filt, float* __restrict__ blo,
                   float* __restrict__ bhi, unsigned long long* __restrict__ keys, int32_t* __restrict__ ids, int* __restrict__ error)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_tris) return;
    v3 v[3];
    for (int k = 0; k < 3; k++) {
        const int32_t vi = vidx[3 * (size_t)i + k];
        if (vi < 0 || vi >= p.n_verts) { atomicCAS(error, 0, 1); keys[i] = 0; ids[i] = i; return; }
        v[k] = mk3(verts[3 * (size_t)vi], verts[3 * (size_t)vi + 1], verts[3 * (size_t)vi + 2]);
        const double e = 1e-10;                          // AxisAlignedBox.ContainsPoint (AxisAlignedBox.cs:143-149)
        const bool inside = p.bmin[0] - e < v[k].x && v[k].x < p.bmax[0] + e && p.bmin[1] - e < v[k].y && v[k].y < p.bmax[1] + e &&
                            p.bmin[2] - e < v[k].z && v[k].z < p.bmax[2] + e;
        if (!inside) { atomicCAS(error, 0, 2); keys[i] = 0; ids[i] = i; return; }
    }
    const v3 e1 = sub3(v[1], v[0]), e2 = sub3(v[2], v[0]);
    v3 n = cross3(e1, e2);
    {
        const double e = 1e-10;                          // Vector.IsZeroVector (Vector.cs:140-147) -> normal (1,0,0)
        if (-e < n.x && n.x < e && -e < n.y && n.y < e && -e < n.z && n.z < e) n = mk3(1.0, 0.0, 0.0);
    }
    const double inv = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(n.x, n.x), __dmul_rn(n.y, n.y)), __dmul_rn(n.z, n.z))));
    const v3 nn = mk3(__dmul_rn(n.x, inv), __dmul_rn(n.y, inv), __dmul_rn(n.z, inv));
    const v3 e1p = cross3(e1, n), e2p = cross3(e2, n);
    TriRec r;
    r.nx = nn.x; r.ny = nn.y; r.nz = nn.z;
    r.d = dot3(v[0], nn);
    r.v1x = v[0].x; r.v1y = v[0].y; r.v1z = v[0].z;
    r.den1 = dot3(e1, e2p);
    r.e2px = e2p.x; r.e2py = e2p.y; r.e2pz = e2p.z;
    r.den2 = dot3(e2, e1p);
    r.e1px = e1p.x; r.e1py = e1p.y; r.e1pz = e1p.z;
    r.color = argb[i];
    r.index = i;
    recs[i] = r;

    TriFilt f;
    f.nx = __double2float_rn(r.nx); f.ny = __double2float_rn(r.ny); f.nz = __double2float_rn(r.nz); f.d = __double2float_rn(r.d);
    f.v1x = __double2float_rn(r.v1x); f.v1y = __double2float_rn(r.v1y); f.v1z = __double2float_rn(r.v1z); f._pad = 0.0f;
    f.ax = f.ay = f.az = f.bx = f.by = f.bz = 0.0f;
    if (r.den1 == 0.0 || r.den2 == 0.0 || !isfinite(r.den1) || !isfinite(r.den2)) {
        const bool never = (r.den1 == 0.0 || r.den2 == 0.0);
        f.a1 = never ? -1.0f : CUDART_INF_F;
        f.b1 = f.a1;
    } else {
        f.ax = __double2float_rn(__ddiv_rn(r.e2px, r.den1)); f.ay = __double2float_rn(__ddiv_rn(r.e2py, r.den1));
        f.az = __double2float_rn(__ddiv_rn(r.e2pz, r.den1));
        f.bx = __double2float_rn(__ddiv_rn(r.e1px, r.den2)); f.by = __double2float_rn(__ddiv_rn(r.e1py, r.den2));
        f.bz = __double2float_rn(__ddiv_rn(r.e1pz, r.den2));
        f.a1 = __double2float_ru(__dmul_rn(__dadd_rn(__dadd_rn(fabs((double)f.ax), fabs((double)f.ay)), fabs((double)f.az)), 1.0 + 1e-6));
        f.b1 = __double2float_ru(__dmul_rn(__dadd_rn(__dadd_rn(fabs((double)f.bx), fabs((double)f.by)), fabs((double)f.bz)), 1.0 + 1e-6));
        if (!isfinite(f.a1) || !isfinite(f.b1)) { f.a1 = CUDART_INF_F; f.b1 = CUDART_INF_F; }
    }
    filt[i] = f;

    // bounds rounded outward (round_down / round_up of sr_bvh.cpp) and the centroid's Morton key
    float lo[3], hi[3];
    lo[0] = __double2float_rd(fmin(v[0].x, fmin(v[1].x, v[2].x))); hi[0] = __double2float_ru(fmax(v[0].x, fmax(v[1].x, v[2].x)));
    lo[1] = __double2float_rd(fmin(v[0].y, fmin(v[1].y, v[2].y))); hi[1] = __double2float_ru(fmax(v[0].y, fmax(v[1].y, v[2].y)));
    lo[2] = __double2float_rd(fmin(v[0].z, fmin(v[1].z, v[2].z))); hi[2] = __double2float_ru(fmax(v[0].z, fmax(v[1].z, v[2].z)));
    unsigned int q[3];
    for (int k = 0; k < 3; k++) {
        blo[3 * (size_t)i + k] = lo[k]; bhi[3 * (size_t)i + k] = hi[k];
        const float c = __fadd_rn(__fmul_rn(0.5f, lo[k]), __fmul_rn(0.5f, hi[k]));
        float u = __fmul_rn(__fsub_rn(c, p.cmin[k]), p.cinv[k]);
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        q[k] = min((unsigned int)__float2uint_rz(__fmul_rn(u, 2097152.0f)), 2097151u);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    ids[i] = i;
}

// length of the common prefix of the (key, position) pairs at sorted positions i and j; -1 outside
__device__ __forceinline__ int prefix_len(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a != b) return __clzll((long long)(a ^ b));
    return 64 + __clz(i ^ j);                            // equal keys: the sorted position breaks the tie
}

// Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees, and k-d trees":
// internal node i covers a range of sorted leaves determined from the prefix lengths around i.
__global__ void __launch_bounds__(256)
radix_tree_kernel(const unsigned long long* __restrict__ keys, int n, int32_t* __restrict__ left, int32_t* __restrict__ right,
                  int32_t* __restrict__ parent_of_internal, int32_t* __restrict__ parent_of_leaf, int2* __restrict__ range)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (prefix_len(keys, n, i, i + 1) - prefix_len(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = prefix_len(keys, n, i, i - d);
    int lmax = 2;
    while (prefix_len(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (prefix_len(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = prefix_len(keys, n, i, j);
    int s = 0;
    for (int div = 2, t = (l + div - 1) / div;; div *= 2, t = (l + div - 1) / div) {
        if (prefix_len(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    // children: a leaf is stored as -1 - position
    const int32_t lc = (lo == gamma) ? -1 - gamma : gamma;
    const int32_t rc = (hi == gamma + 1) ? -1 - (gamma + 1) : gamma + 1;
    left[i] = lc; right[i] = rc;
    range[i] = make_int2(lo, hi);
    if (lc >= 0) parent_of_internal[lc] = i; else parent_of_leaf[gamma] = i;
    if (rc >= 0) parent_of_internal[rc] = i; else parent_of_leaf[gamma + 1] = i;
    if (i == 0) parent_of_internal[0] = -1;
}

// bottom-up: the second thread to arrive at a node merges its children's boxes (min / max are exact, so
// the result does not depend on which thread that is) and carries on; also the height of every node
__global__ void __launch_bounds__(256)
fit_boxes_kernel(int n, const int32_t* __restrict__ sorted_ids, const float* __restrict__ blo, const float* __restrict__ bhi,
                 const int32_t* __restrict__ left, const int32_t* __restrict__ right, const int32_t* __restrict__ parent_of_internal,
                 const int32_t* __restrict__ parent_of_leaf, float* __restrict__ nlo, float* __restrict__ nhi,
                 int32_t* __restrict__ height, unsigned int* __restrict__ arrivals)
{
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parent_of_leaf[leaf];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&arrivals[node], 1u) == 0u) return;          // first to arrive: the sibling subtree is not done yet
        __threadfence();
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        int h = 0;
        const int32_t ch[2] = {left[node], right[node]};
        for (int c = 0; c < 2; c++) {
            const float* clo; const float* chi;
            if (ch[c] < 0) {
                const int32_t prim = sorted_ids[-1 - ch[c]];
                clo = blo + 3 * (size_t)prim; chi = bhi + 3 * (size_t)prim;
            } else {
                clo = nlo + 3 * (size_t)ch[c]; chi = nhi + 3 * (size_t)ch[c];
                h = max(h, ((volatile int32_t*)height)[ch[c]]);
            }
            for (int k = 0; k < 3; k++) {
                lo[k] = fminf(lo[k], ((volatile const float*)clo)[k]); hi[k] = fmaxf(hi[k], ((volatile const float*)chi)[k]);
            }
        }
        for (int k = 0; k < 3; k++) { nlo[3 * (size_t)node + k] = lo[k]; nhi[3 * (size_t)node + k] = hi[k]; }
        height[node] = h + 1;
        node = parent_of_internal[node];
    }
}

__device__ __forceinline__ void padded_box(const float* lo, const float* hi, float pad, float* out_lo, float* out_hi)
{
    for (int k = 0; k < 3; k++) {                        // write_box of sr_bvh.cpp: pad, rounded outward
        out_lo[k] = __double2float_rd((double)lo[k] - (double)pad);
        out_hi[k] = __double2float_ru((double)hi[k] + (double)pad);
    }
}

// radix tree -> BvhNode array (both children's padded boxes + pre-encoded links).  A subtree of at most
// leaf_max primitives becomes one leaf (its sorted range is contiguous); the radix-tree nodes below it are
// simply never linked.
__global__ void __launch_bounds__(256)
emit_nodes_kernel(int n, float pad, int leaf_max, const int32_t* __restrict__ sorted_ids, const float* __restrict__ blo,
                  const float* __restrict__ bhi, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                  const int2* __restrict__ range, const float* __restrict__ nlo, const float* __restrict__ nhi,
                  BvhNode* __restrict__ nodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    BvhNode nd;
    float lo[2][3], hi[2][3];
    int32_t link[2], count[2];
    const int32_t ch[2] = {left[i], right[i]};
    for (int c = 0; c < 2; c++) {
        if (ch[c] < 0) {
            const int pos = -1 - ch[c];
            const int32_t prim = sorted_ids[pos];
            padded_box(blo + 3 * (size_t)prim, bhi + 3 * (size_t)prim, pad, lo[c], hi[c]);
            link[c] = -1 - (pos * 16 + 1); count[c] = 1;
        } else {
            padded_box(nlo + 3 * (size_t)ch[c], nhi + 3 * (size_t)ch[c], pad, lo[c], hi[c]);
            const int2 r = range[ch[c]];
            const int cnt = r.y - r.x + 1;
            if (cnt <= leaf_max) { link[c] = -1 - (r.x * 16 + cnt); count[c] = cnt; }
            else { link[c] = ch[c]; count[c] = 0; }
        }
    }
    nd.lo0x = lo[0][0]; nd.lo0y = lo[0][1]; nd.lo0z = lo[0][2]; nd.hi0x = hi[0][0]; nd.hi0y = hi[0][1]; nd.hi0z = hi[0][2];
    nd.lo1x = lo[1][0]; nd.lo1y = lo[1][1]; nd.lo1z = lo[1][2]; nd.hi1x = hi[1][0]; nd.hi1y = hi[1][1]; nd.hi1z = hi[1][2];
    nd.child0 = link[0]; nd.child1 = link[1]; nd.count0 = count[0]; nd.count1 = count[1];
    nodes[i] = nd;
}

// a single triangle: the root holds it as its only leaf child (like the host builder)
__global__ void single_leaf_kernel(float pad, const float* __restrict__ blo, const float* __restrict__ bhi, BvhNode* __restrict__ nodes)
{
    BvhNode nd;
    float lo[3], hi[3];
    padded_box(blo, bhi, pad, lo, hi);
    nd.lo0x = lo[0]; nd.lo0y = lo[1]; nd.lo0z = lo[2]; nd.hi0x = hi[0]; nd.hi0y = hi[1]; nd.hi0z = hi[2];
    nd.lo1x = nd.lo1y = nd.lo1z = nd.hi1x = nd.hi1y = nd.hi1z = FLT_MAX;
    nd.child0 = -1 - (0 * 16 + 1); nd.count0 = 1;
    nd.child1 = kNoChild; nd.count1 = -1;
    nodes[0] = nd;
}

// records into leaf (sorted) order
__global__ void __launch_bounds__(256)
gather_records_kernel(int n, const int32_t* __restrict__ sorted_ids, const TriRec* __restrict__ recs, const TriFilt* __restrict__ filt,
                      TriRec* __restrict__ out_recs, TriFilt* __restrict__ out_filt)
{
    // one 16-byte word per thread: 8 words per TriRec, 4 per TriFilt
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_rec_words = (size_t)n * 8, n_filt_words = (size_t)n * 4;
    if (w < n_rec_words) {
        const size_t k = w >> 3, part = w & 7;
        reinterpret_cast<uint4*>(out_recs)[w] = reinterpret_cast<const uint4*>(recs)[(size_t)sorted_ids[k] * 8 + part];
    } else if (w < n_rec_words + n_filt_words) {
        const size_t v = w - n_rec_words, k = v >> 2, part = v & 3;
        reinterpret_cast<uint4*>(out_filt)[v] = reinterpret_cast<const uint4*>(filt)[(size_t)sorted_ids[k] * 4 + part];
    }
}

template <typename T>
cudaError_t dev_alloc(T** p, size_t count)
{
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T) + 16);
}

// carves the build's temporaries out of one allocation (one cudaMalloc / cudaFree instead of twenty)
struct Arena {
    char* base = nullptr; size_t used = 0, cap = 0;
    template <typename T> T* take(size_t count)
    {
        used = (used + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + used) : nullptr;
        used += count * sizeof(T) + 16;
        return p;
    }
};

}  // namespace

// Builds one mesh on the device.  On success *out_tris / *out_filt / *out_nodes are cudaMalloc'd buffers in
// leaf order owned by the caller; *status: 0 ok, 1 vertex index out of range, 2 vertex outside the bounding
// box, 3 tree deeper than the traversal stack.
cudaError_t build_mesh_on_device(const double* h_verts, int32_t n_verts, const int32_t* h_vidx, const uint32_t* h_argb, int32_t n_tris,
                                 const double bmin[3], const double bmax[3], float pad, int leaf_max, cudaStream_t stream, TriRec** out_tris,
                                 TriFilt** out_filt, BvhNode** out_nodes, int32_t* out_n_nodes, int32_t* out_depth, int* status)
{
    *out_tris = nullptr; *out_filt = nullptr; *out_nodes = nullptr; *out_n_nodes = 0; *out_depth = 0; *status = 0;
    const int n = n_tris;
    const int n_internal = n > 1 ? n - 1 : 1;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; return e == cudaSuccess; };

    size_t sort_bytes = 0;
    ok(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                       (const int32_t*)nullptr, (int32_t*)nullptr, n, 0, 63, stream));

    double* d_verts; int32_t* d_vidx; uint32_t* d_argb; TriRec* d_recs; TriFilt* d_filt;
    float *d_blo, *d_bhi, *d_nlo, *d_nhi;
    unsigned long long *d_keys, *d_keys_sorted;
    int32_t *d_ids, *d_ids_sorted, *d_left, *d_right, *d_par_int, *d_par_leaf, *d_height;
    int2* d_range; unsigned int* d_arrivals; int* d_error; char* d_temp;
    Arena arena;
    for (int pass = 0; pass < 2 && ok(cudaSuccess); pass++) {      // pass 0 sizes the arena, pass 1 hands out pointers
        arena.used = 0;
        d_verts = arena.take<double>((size_t)3 * n_verts); d_vidx = arena.take<int32_t>((size_t)3 * n); d_argb = arena.take<uint32_t>((size_t)n);
        d_recs = arena.take<TriRec>((size_t)n); d_filt = arena.take<TriFilt>((size_t)n);
        d_blo = arena.take<float>((size_t)3 * n); d_bhi = arena.take<float>((size_t)3 * n);
        d_nlo = arena.take<float>((size_t)3 * n_internal); d_nhi = arena.take<float>((size_t)3 * n_internal);
        d_keys = arena.take<unsigned long long>((size_t)n); d_keys_sorted = arena.take<unsigned long long>((size_t)n);
        d_ids = arena.take<int32_t>((size_t)n); d_ids_sorted = arena.take<int32_t>((size_t)n);
        d_left = arena.take<int32_t>((size_t)n_internal); d_right = arena.take<int32_t>((size_t)n_internal);
        d_par_int = arena.take<int32_t>((size_t)n_internal); d_par_leaf = arena.take<int32_t>((size_t)n);
        d_height = arena.take<int32_t>((size_t)n_internal); d_range = arena.take<int2>((size_t)n_internal);
        d_arrivals = arena.take<unsigned int>((size_t)n_internal); d_error = arena.take<int>(1);
        d_temp = arena.take<char>(sort_bytes);
        if (pass == 0) { arena.cap = arena.used + 256; ok(cudaMalloc(reinterpret_cast<void**>(&arena.base), arena.cap)); }
    }
    TriRec* d_recs_sorted = nullptr; TriFilt* d_filt_sorted = nullptr; BvhNode* d_nodes = nullptr;
    ok(dev_alloc(&d_nodes, (size_t)n_internal)); ok(dev_alloc(&d_recs_sorted, (size_t)n)); ok(dev_alloc(&d_filt_sorted, (size_t)n));

    BuildParams p;
    for (int k = 0; k < 3; k++) {
        p.bmin[k] = bmin[k]; p.bmax[k] = bmax[k];
        p.cmin[k] = (float)bmin[k];
        const float ext = (float)bmax[k] - (float)bmin[k];
        p.cinv[k] = ext > 0.0f ? 1.0f / ext : 0.0f;
    }
    p.pad = pad; p.n_verts = n_verts; p.n_tris = n;
    const int tpb = 256;
    int h_error = 0;
    int32_t h_height = 1;
    if (ok(cudaSuccess)) {
        // everything below is stream-ordered: one synchronisation at the end
        ok(cudaMemcpyAsync(d_verts, h_verts, sizeof(double) * 3 * (size_t)n_verts, cudaMemcpyHostToDevice, stream));
        ok(cudaMemcpyAsync(d_vidx, h_vidx, sizeof(int32_t) * 3 * (size_t)n, cudaMemcpyHostToDevice, stream));
        ok(cudaMemcpyAsync(d_argb, h_argb, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, stream));
        ok(cudaMemsetAsync(d_error, 0, sizeof(int), stream));
        tri_prepare_kernel<<<(n + tpb - 1) / tpb, tpb, 0, stream>>>(d_verts, d_vidx, d_argb, p, d_recs, d_filt, d_blo, d_bhi, d_keys, d_ids,
                                                                    d_error);
        ok(cudaGetLastError());
        // stable LSD radix sort of (key, index): equal keys keep the triangle order.  (On a bad mesh the
        // kernels below run on partly unwritten keys; every index they form stays inside [0, n).)
        ok(cub::DeviceRadixSort::SortPairs(d_temp, sort_bytes, d_keys, d_keys_sorted, d_ids, d_ids_sorted, n, 0, 63, stream));
        if (n > 1) {
            ok(cudaMemsetAsync(d_arrivals, 0, sizeof(unsigned int) * (size_t)n_internal, stream));
            ok(cudaMemsetAsync(d_height, 0, sizeof(int32_t) * (size_t)n_internal, stream));
            radix_tree_kernel<<<(n - 1 + tpb - 1) / tpb, tpb, 0, stream>>>(d_keys_sorted, n, d_left, d_right, d_par_int, d_par_leaf, d_range);
            fit_boxes_kernel<<<(n + tpb - 1) / tpb, tpb, 0, stream>>>(n, d_ids_sorted, d_blo, d_bhi, d_left, d_right, d_par_int, d_par_leaf,
                                                                      d_nlo, d_nhi, d_height, d_arrivals);
            emit_nodes_kernel<<<(n - 1 + tpb - 1) / tpb, tpb, 0, stream>>>(n, pad, leaf_max, d_ids_sorted, d_blo, d_bhi, d_left, d_right,
                                                                           d_range, d_nlo, d_nhi, d_nodes);
            ok(cudaGetLastError());
            ok(cudaMemcpyAsync(&h_height, d_height, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        } else {
            single_leaf_kernel<<<1, 1, 0, stream>>>(pad, d_blo, d_bhi, d_nodes);
            ok(cudaGetLastError());
        }
        const size_t words = (size_t)n * 12;
        gather_records_kernel<<<(unsigned)((words + tpb - 1) / tpb), tpb, 0, stream>>>(n, d_ids_sorted, d_recs, d_filt, d_recs_sorted,
                                                                                       d_filt_sorted);
        ok(cudaGetLastError());
        ok(cudaMemcpyAsync(&h_error, d_error, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ok(cudaStreamSynchronize(stream));
    }
    if (ok(cudaSuccess) && h_error == 0) {
        *out_depth = h_height + 1;                       // node levels + the leaf level
        if (*out_depth >= kStackEntries) h_error = 3;
    }
    *status = h_error;
    if (e == cudaSuccess && h_error == 0) {
        *out_tris = d_recs_sorted; *out_filt = d_filt_sorted; *out_nodes = d_nodes; *out_n_nodes = n_internal;
        d_recs_sorted = nullptr; d_filt_sorted = nullptr; d_nodes = nullptr;
    }
    cudaFree(arena.base); cudaFree(d_nodes); cudaFree(d_recs_sorted); cudaFree(d_filt_sorted);
    return e;
}

}  // namespace sr
