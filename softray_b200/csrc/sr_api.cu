// sr_api.cu -- the C ABI of libsoftray_cuda.so (include/softray_cuda.h): context, scene flatten +
// upload (replaces MakeRayTracableGeometry_* Renderer.cs:1452-1494, the Triangle ctor precompute
// Triangle.cs:29-57 and the SpatialSubdivision ctor SpatialSubdivision.cs:267-315) and the frame
// entry points (replace the body of Renderer.RaytraceGeometry, Renderer.cs:1501-1687).
//
// Host arithmetic that feeds reference-exact device code (triangle precompute, light / camera
// transforms, area-light offsets) is written in plain double expressions in the reference's
// evaluation order and this file is compiled with -ffp-contract=off, so no FMA is ever formed.
// There is NO CPU rendering path here: every pixel comes from render_kernel (sr_render.cu).
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <atomic>
#include <string>
#include <mutex>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/softray_cuda.h"
#include "sr_bvh.h"
#include "sr_types.h"
#include "sr_wave.h"

namespace sr {
cudaError_t launch_render(const DevFrame& f, const DevScene& sc, const DevInstance* d_insts, const double* d_offsets,
                          uint32_t* d_pixels, int32_t* d_ids, unsigned int* d_tile_counter, DevCounters* d_counters,
                          int grid_blocks, cudaStream_t stream);
int render_kernel_occupancy(int smem_bytes, bool staged);
int render_kernel_block_threads();
cudaError_t measure_fma_peak(bool fp64, int sm_count, cudaStream_t stream, double* tflops);
cudaError_t build_mesh_on_device(const double* h_verts, int32_t n_verts, const int32_t* h_vidx, const uint32_t* h_argb, int32_t n_tris,
                                 const double bmin[3], const double bmax[3], float pad, int leaf_max, cudaStream_t stream, TriRec** out_tris,
                                 TriFilt** out_filt, BvhNode** out_nodes, int32_t* out_n_nodes, int32_t* out_depth, int* status);
cudaError_t launch_resolve(const uint32_t* d_src, uint32_t* d_dst, int dst_w, int dst_h, int aa, int style, uint32_t background,
                           cudaStream_t stream);
}  // namespace sr

using namespace sr;

// ---------------------------------------------------------------------------------------------
// opaque handles
// ---------------------------------------------------------------------------------------------
struct softray_ctx {
    // a GROUP context (softray_create_multi) owns one member context per device and nothing else
    std::vector<softray_ctx*> members;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_staged = nullptr;   // the last frame's constants have left the pinned staging
    bool staging_busy = false;
    bool peer_missing = false;         // group: some member cannot store into member 0's memory (softray_render_device needs it)
    // every frame of a context shares d_insts / d_tile_counter / d_counters: a frame enqueued on another stream
    // than the previous one first waits for that one's kernel (softray_render_device is asynchronous)
    cudaEvent_t ev_frame = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last_frame = false;
    // per-frame device scratch (sized for SOFTRAY_MAX_INSTANCES / SOFTRAY_MAX_SHADOW_SAMPLES)
    DevInstance* d_insts = nullptr;
    double* d_offsets = nullptr;
    unsigned int* d_tile_counter = nullptr;
    DevCounters* d_counters = nullptr;
    // pinned host staging for the frame constants and the counters
    DevInstance* h_insts = nullptr;
    // per-frame TLAS over the instances of a composite frame (pinned staging + device copy)
    BvhNode* h_tlas_nodes = nullptr; BvhNode* d_tlas_nodes = nullptr;
    int32_t* h_tlas_order = nullptr; int32_t* d_tlas_order = nullptr;
    int32_t tlas_nodes_used = 0;
    double* h_offsets = nullptr;
    DevCounters* h_counters = nullptr;
    std::vector<struct softray_scene*> scenes;   // scenes created in this context and not yet destroyed
    // device framebuffer of the host-buffer entry point (grown on demand)
    uint32_t* d_pixels = nullptr;
    int32_t* d_ids = nullptr;
    size_t fb_capacity = 0, ids_capacity = 0;
    int cached_seed = 0, cached_samples = -1;   // area-light offsets currently in h_offsets
    // stage-kernel pipeline (sr_wave.cu): one arena for the records that cross HBM between the stages, grown on demand
    void* wave_base = nullptr;
    size_t wave_bytes = 0;
    uint32_t wave_cap = 0;                      // samples per chunk the arena is laid out for
    int wave_slots = 0;                         // shading points per sample (1 + bounces)
    int last_launches = 1;                      // kernels the last frame launched (softray_stats.launches)
    WaveBufs wave[2];                           // two chunks in flight (wave_render)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    StageTimer stage_timer;                     // softray_frame.profile_stages
    cudaStream_t copy_stream = nullptr;         // chunk-by-chunk DMA of a frame into a page-locked host surface
    std::vector<cudaEvent_t> copy_events;
    std::string err;
};

struct softray_scene {
    // a scene created in a group context: one replica per member (replicas[i] lives in ctx->members[i])
    std::vector<softray_scene*> replicas;
    std::vector<size_t> alloc_bytes;   // size of allocs[i]
    std::vector<DevMesh> host_meshes;  // host copy of dev.meshes (replication patches its pointers)
    softray_ctx* ctx = nullptr;
    DevScene dev;                      // passed to the kernel by value
    std::vector<void*> allocs;         // every device allocation of this scene
    std::vector<int32_t> mesh_tris;    // n_tris per mesh (hit-id bases)
    struct V3 { double v[3]; };
    std::vector<V3> mesh_bmin, mesh_bmax;   // Model.Min/Max per mesh (per-frame instance hierarchy)
    uint64_t fingerprint = 1469598103934665603ull;
    size_t device_bytes = 0;
    double all_min[3] = {1e300, 1e300, 1e300}, all_max[3] = {-1e300, -1e300, -1e300};   // every primitive
    // buffers built on the device (SOFTRAY_ACCEL_LBVH): folded into the fingerprint on first request
    struct DevBuf { const void* ptr; size_t bytes; };
    std::vector<DevBuf> unhashed;
};

static thread_local std::string g_last_error;

// Live handles.  A host runtime with non-deterministic finalisation (the .NET finaliser thread, CPython at
// interpreter exit) may release a context before its scenes, or a scene twice: softray_destroy therefore
// releases the scenes its context still owns, and destroying a handle that is no longer live is a no-op.
static std::mutex g_live_mutex;
static std::unordered_set<const void*> g_live_ctx, g_live_scenes;

static int fail(softray_ctx* ctx, int code, const std::string& msg)
{
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}

static int cuda_fail(softray_ctx* ctx, cudaError_t e, const char* what)
{
    const int code = (e == cudaErrorMemoryAllocation) ? SOFTRAY_E_OOM
                     : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) ? SOFTRAY_E_NO_DEVICE
                                                                                                                   : SOFTRAY_E_CUDA;
    cudaGetLastError();     // the runtime remembers a failed call until asked: the next launch check must not see it
    return fail(ctx, code, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

#define SR_CUDA(ctx, call)                                             \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, #call);   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// exact host math (Engine3D/Vector.cs, Matrix.cs) -- never contracted (see file header)
// ---------------------------------------------------------------------------------------------
namespace {

struct hv { double x, y, z; };
inline hv hmk(double x, double y, double z) { hv r = {x, y, z}; return r; }
inline hv hsub(hv a, hv b) { return hmk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline hv hscale(hv a, double s) { return hmk(a.x * s, a.y * s, a.z * s); }
inline double hdot(hv a, hv b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                 // Vector.cs:99-102
inline hv hcross(hv a, hv b)                                                                   // Vector.cs:104-110
{
    return hmk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline hv hnormalise(hv a)                                                                     // Vector.cs:177-185
{
    const double inv = 1.0 / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return hscale(a, inv);
}
inline bool his_zero(hv a)                                                                     // Vector.cs:140-147
{
    const double e = 1e-10;
    return -e < a.x && a.x < e && -e < a.y && a.y < e && -e < a.z && a.z < e;
}
inline hv hmul3x3(const double* m, hv v)                                                       // Matrix.cs:34-41
{
    return hmk(v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[4] + v.y * m[5] + v.z * m[6],
               v.x * m[8] + v.y * m[9] + v.z * m[10]);
}
inline hv hmul3x4(const double* m, hv v)                                                       // Matrix.cs:50-57
{
    return hmk(v.x * m[0] + v.y * m[1] + v.z * m[2] + m[3], v.x * m[4] + v.y * m[5] + v.z * m[6] + m[7],
               v.x * m[8] + v.y * m[9] + v.z * m[10] + m[11]);
}

// 4x4 row-major product, accumulating from 0.0 over i = 0..3 (Matrix.cs:74-92)
void mat_mul(const double* a, const double* b, double* out)
{
    double r[16];
    for (int row = 0; row < 4; row++)
        for (int col = 0; col < 4; col++) {
            double sum = 0.0;
            for (int i = 0; i < 4; i++) sum += a[row * 4 + i] * b[i * 4 + col];
            r[row * 4 + col] = sum;
        }
    std::memcpy(out, r, sizeof r);
}
void mat_identity(double* m)
{
    std::memset(m, 0, 16 * sizeof(double));
    m[0] = m[5] = m[10] = m[15] = 1.0;
}
void mat_translate(double* m, double x, double y, double z)   // Matrix.cs:94-114
{
    mat_identity(m);
    m[3] = x; m[7] = y; m[11] = z;
}
void mat_yaw(double* m, double a)                             // Matrix.cs:116-132
{
    mat_identity(m);
    m[0] = std::cos(a); m[8] = std::sin(a); m[2] = -std::sin(a); m[10] = std::cos(a);
}
void mat_pitch(double* m, double a)                           // Matrix.cs:134-150
{
    mat_identity(m);
    m[5] = std::cos(a); m[9] = std::sin(a); m[6] = -std::sin(a); m[10] = std::cos(a);
}
void mat_roll(double* m, double a)                            // Matrix.cs:152-168
{
    mat_identity(m);
    m[0] = std::cos(a); m[4] = std::sin(a); m[1] = -std::sin(a); m[5] = std::cos(a);
}

// System.Random(int seed) of the .NET Framework 4.x BCL (Knuth's subtractive generator; the
// algorithm is not in the reference tree -- SURVEY.md Appendix B restates it).  ShadowMethod's
// ctor draws the area-light offsets from it (ShadowMethod.cs:63-72, seeded at Renderer.cs:1624).
class DotNetRandom {
public:
    explicit DotNetRandom(int32_t seed)
    {
        const int32_t kBig = 2147483647, kSeed = 161803398;
        const int32_t sub = seed == INT32_MIN ? kBig : (seed < 0 ? -seed : seed);
        int32_t mj = kSeed - sub, mk = 1;
        std::memset(a_, 0, sizeof a_);
        a_[55] = mj;
        for (int i = 1; i < 55; i++) {
            const int ii = (21 * i) % 55;
            a_[ii] = mk;
            mk = mj - mk;
            if (mk < 0) mk += kBig;
            mj = a_[ii];
        }
        for (int round = 0; round < 4; round++)
            for (int i = 1; i < 56; i++) {
                a_[i] = (int32_t)((uint32_t)a_[i] - (uint32_t)a_[1 + (i + 30) % 55]);
                if (a_[i] < 0) a_[i] += kBig;
            }
        inext_ = 0; inextp_ = 21;
    }
    double next_double()
    {
        const int32_t kBig = 2147483647;
        if (++inext_ >= 56) inext_ = 1;
        if (++inextp_ >= 56) inextp_ = 1;
        int32_t v = (int32_t)((uint32_t)a_[inext_] - (uint32_t)a_[inextp_]);
        if (v == kBig) v--;
        if (v < 0) v += kBig;
        a_[inext_] = v;
        return v * (1.0 / kBig);
    }

private:
    int32_t a_[56];
    int inext_, inextp_;
};

// offsets_i = normalise(2u-1, 2v-1, 2w-1) * 0.2 (ShadowMethod.cs:63-72, lightSourceRadius :10)
void area_light_offsets(int32_t seed, int n, double* out)
{
    DotNetRandom rng(seed);
    for (int i = 0; i < n; i++) {
        const double a = rng.next_double() * 2 - 1;
        const double b = rng.next_double() * 2 - 1;
        const double c = rng.next_double() * 2 - 1;
        const hv o = hscale(hnormalise(hmk(a, b, c)), 0.2);
        out[3 * i] = o.x; out[3 * i + 1] = o.y; out[3 * i + 2] = o.z;
    }
}

// layout fingerprint: FNV-1a over 64-bit words (the tail bytewise) -- every uploaded byte takes part
inline void fnv(uint64_t* h, const void* data, size_t n)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint64_t x = *h;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, p + i, 8); x ^= w; x *= 1099511628211ull; }
    for (; i < n; i++) { x ^= p[i]; x *= 1099511628211ull; }
    *h = x;
}

// run fn(begin, end) over [0, n) on several host threads (per-element work only: no ordering effects)
template <typename F>
void parallel_for(int64_t n, F fn)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = n >= 100000 ? (int)std::min<unsigned>(hw ? hw : 1u, 32u) : 1;
    if (nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> pool;
    const int64_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
        if (b < e) pool.emplace_back([=]() { fn(b, e); });
    }
    for (auto& th : pool) th.join();
}

// upload one host array; records the allocation and folds the bytes into the layout fingerprint
template <typename T>
int upload(softray_scene* sc, const std::vector<T>& host, const T** out)
{
    *out = nullptr;
    if (host.empty()) return SOFTRAY_OK;
    void* d = nullptr;
    const size_t bytes = host.size() * sizeof(T);
    SR_CUDA(sc->ctx, cudaMalloc(&d, bytes));
    sc->allocs.push_back(d); sc->alloc_bytes.push_back(bytes);
    sc->device_bytes += bytes;
    // (pageable source: the call returns once the bytes are staged, so `host` may die right after; the stream is
    // synchronised once, at the end of softray_scene_create)
    SR_CUDA(sc->ctx, cudaMemcpyAsync(d, host.data(), bytes, cudaMemcpyHostToDevice, sc->ctx->stream));
    fnv(&sc->fingerprint, host.data(), bytes);
    *out = static_cast<const T*>(d);
    return SOFTRAY_OK;
}

// Triangle ctor + Plane ctor (Triangle.cs:29-57, Plane.cs:22-29) and the two denominators
// Triangle.IntersectRay evaluates per call (Triangle.cs:90,95)
void make_tri_rec(TriRec* t, hv v1, hv v2, hv v3, uint32_t color, int32_t index)
{
    const hv e1 = hsub(v2, v1), e2 = hsub(v3, v1);
    hv n = hcross(e1, e2);
    if (his_zero(n)) n = hmk(1, 0, 0);
    const hv nn = hnormalise(n);
    const hv e1p = hcross(e1, n), e2p = hcross(e2, n);
    std::memset(t, 0, sizeof *t);
    t->nx = nn.x; t->ny = nn.y; t->nz = nn.z;
    t->d = hdot(v1, nn);
    t->v1x = v1.x; t->v1y = v1.y; t->v1z = v1.z;
    t->den1 = hdot(e1, e2p);
    t->e2px = e2p.x; t->e2py = e2p.y; t->e2pz = e2p.z;
    t->den2 = hdot(e2, e1p);
    t->e1px = e1p.x; t->e1py = e1p.y; t->e1pz = e1p.z;
    t->color = color;
    t->index = index;
}

// FP32 filter record of one triangle (sr_types.h TriFilt): the exact record's quantities with the
// two divisions folded in, rounded to nearest; a1/b1 bound the L1 norms of the rounded vectors.
void make_tri_filt(TriFilt* f, const TriRec& t)
{
    std::memset(f, 0, sizeof *f);
    f->nx = (float)t.nx; f->ny = (float)t.ny; f->nz = (float)t.nz; f->d = (float)t.d;
    f->v1x = (float)t.v1x; f->v1y = (float)t.v1y; f->v1z = (float)t.v1z;
    if (t.den1 == 0.0 || t.den2 == 0.0 || !std::isfinite(t.den1) || !std::isfinite(t.den2)) {
        // zero area: every quotient of Triangle.IntersectRay is NaN or +-inf, the test never passes
        // (Triangle.cs:42-43, TriangleTests.cs:35-44) -- unless a NaN slips through a comparison,
        // so only an exactly-zero denominator is declared "never hit"; anything else odd goes to
        // the exact test
        const bool never = (t.den1 == 0.0 || t.den2 == 0.0);
        f->a1 = never ? -1.0f : INFINITY;
        f->b1 = f->a1;
        return;
    }
    f->ax = (float)(t.e2px / t.den1); f->ay = (float)(t.e2py / t.den1); f->az = (float)(t.e2pz / t.den1);
    f->bx = (float)(t.e1px / t.den2); f->by = (float)(t.e1py / t.den2); f->bz = (float)(t.e1pz / t.den2);
    f->a1 = round_up((std::fabs((double)f->ax) + std::fabs((double)f->ay) + std::fabs((double)f->az)) * (1.0 + 1e-6));
    f->b1 = round_up((std::fabs((double)f->bx) + std::fabs((double)f->by) + std::fabs((double)f->bz)) * (1.0 + 1e-6));
    if (!std::isfinite(f->a1) || !std::isfinite(f->b1)) { f->a1 = INFINITY; f->b1 = INFINITY; }
}

inline bool box_contains(const double* mn, const double* mx, hv p)   // AxisAlignedBox.cs:143-149
{
    const double e = 1e-10;
    return mn[0] - e < p.x && p.x < mx[0] + e && mn[1] - e < p.y && p.y < mx[1] + e && mn[2] - e < p.z && p.z < mx[2] + e;
}

inline double max_abs3(const double* a, const double* b)
{
    double m = 0.0;
    for (int k = 0; k < 3; k++) { m = std::fmax(m, std::fabs(a[k])); m = std::fmax(m, std::fabs(b[k])); }
    return m;
}

// FP32 traversal slack (DESIGN.md "FP32 candidate search"): the slab test evaluates
// fma(plane, 1/d, -o/d) with o, d and the box planes rounded to FP32; all those roundings together
// move a box face by < 4e-7 * (largest coordinate in play).  Boxes are padded by 2^-18 of that
// (~9.5x margin) so the test can only produce false positives.
// SAH primitive-test costs relative to one node visit (measured instruction counts);
// SOFTRAY_SAH_ISECT overrides both for experiments.
constexpr double kTriIsectCost = 2.0, kSphereIsectCost = 2.0;
inline double sah_isect_cost(double dflt)
{
    const char* e = std::getenv("SOFTRAY_SAH_ISECT");
    if (e && *e) { const double v = std::atof(e); if (v > 0.0) return v; }
    return dflt;
}

constexpr long long kPhaseSyncMaxPrims = 65536;
constexpr int kPhaseSyncComposite = 0;
constexpr int kPhaseSyncStages = 3;           // bit 0: tile fetch + ray start; 1: sphere search; 2: exact spheres; 3: mesh search; 4: clip + exact triangles.
                                              // config2 ms: 0 -> 0.717, 1 -> 0.688, 3 -> 0.638, 9 -> 0.640, 19 -> 0.641, 7 -> 0.646, 31 -> 0.654
constexpr int kBundleMinSamples = 16, kBundleBudget = 384;   // budget counts child boxes: 384 = 192 nodes
inline int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return (e && *e) ? std::atoi(e) : dflt;
}

// primitives per leaf of the device-built tree (a radix-tree subtree this small becomes one leaf);
// SOFTRAY_LBVH_LEAF overrides it for experiments.  Measured on one B200 (config3 / config5 ms per frame):
// 1: 65.7 / 20.9, 2: 67.2 / 21.2, 4: 74.5 / 22.9, 8: 90.3 / 26.6 (host SAH tree: 48.7 / 17.1)
constexpr int kLbvhLeaf = 1;
inline int lbvh_leaf_max() { const int v = env_int("SOFTRAY_LBVH_LEAF", kLbvhLeaf); return v < 1 ? 1 : (v > kMaxLeafPrims ? kMaxLeafPrims : v); }

// bounding sphere of a mesh around its box centre: radius = the farthest vertex (every triangle is the hull of its vertices)
inline void mesh_bounding_sphere(const softray_mesh& m, DevMesh* dm)
{
    double c[3], r2 = 0.0;
    for (int k = 0; k < 3; k++) c[k] = 0.5 * (m.bbox_min[k] + m.bbox_max[k]);
    for (int64_t i = 0; i < (int64_t)m.n_verts; i++) {
        double d2 = 0.0;
        for (int k = 0; k < 3; k++) { const double d = m.verts_xyz[3 * i + k] - c[k]; d2 += d * d; }
        r2 = std::fmax(r2, d2);
    }
    for (int k = 0; k < 3; k++) dm->bs_center[k] = c[k];
    dm->bs_radius = std::sqrt(r2);
}

inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f; f.d = d ? d : 1u; f._pad = 0; f.mul = 0; f.shift = 0;
    if (f.d == 1u) return f;
    uint32_t lg = 0;
    while ((1ull << lg) < f.d) lg++;                       // ceil(log2 d)
    const uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + f.d - 1) / f.d);
    f.shift = p - 32;
    return f;
}

inline float traversal_pad(double max_coord) { return round_up(std::ldexp(std::fmax(max_coord, 1e-30), -18)); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int softray_abi_version(void) { return SOFTRAY_ABI_VERSION; }

extern "C" int softray_abi_sizeof(int32_t which)
{
    switch (which) {
    case 0: return (int)sizeof(softray_mesh);
    case 1: return (int)sizeof(softray_sphere);
    case 2: return (int)sizeof(softray_scene_desc);
    case 3: return (int)sizeof(softray_instance);
    case 4: return (int)sizeof(softray_frame);
    case 5: return (int)sizeof(softray_stats);
    default: return -1;
    }
}

extern "C" const char* softray_last_error(const softray_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : g_last_error.c_str();
}

static void release_scene(softray_scene* scene);

extern "C" void softray_destroy(softray_ctx* ctx)
{
    if (!ctx) return;
    std::vector<softray_scene*> orphans;
    {
        std::lock_guard<std::mutex> lock(g_live_mutex);
        if (!g_live_ctx.erase(ctx)) return;             // not (or no longer) a live context
        orphans.swap(ctx->scenes);
        for (softray_scene* sc : orphans) g_live_scenes.erase(sc);
    }
    for (softray_scene* sc : orphans) release_scene(sc);
    if (!ctx->members.empty()) {                        // a group: its members hold every CUDA resource
        for (softray_ctx* m : ctx->members) softray_destroy(m);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_insts); cudaFree(ctx->d_offsets); cudaFree(ctx->d_tile_counter); cudaFree(ctx->d_counters);
    cudaFree(ctx->d_pixels); cudaFree(ctx->d_ids); cudaFree(ctx->wave_base);
    cudaFreeHost(ctx->h_insts); cudaFreeHost(ctx->h_offsets); cudaFreeHost(ctx->h_counters);
    cudaFree(ctx->d_tlas_nodes); cudaFree(ctx->d_tlas_order); cudaFreeHost(ctx->h_tlas_nodes); cudaFreeHost(ctx->h_tlas_order);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->ev_staged) cudaEventDestroy(ctx->ev_staged);
    if (ctx->ev_frame) cudaEventDestroy(ctx->ev_frame);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int softray_create(int32_t device_ordinal, softray_ctx** out)
{
    if (!out) return fail(nullptr, SOFTRAY_E_INVALID_ARG, "softray_create: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, SOFTRAY_E_NO_DEVICE,
                    std::string("softray_create: no CUDA device (there is no CPU fallback)") +
                        (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string()));
    if (device_ordinal < 0 || device_ordinal >= n_dev)
        return fail(nullptr, SOFTRAY_E_INVALID_ARG, "softray_create: device ordinal out of range");
    softray_ctx* ctx = new (std::nothrow) softray_ctx();
    if (!ctx) return fail(nullptr, SOFTRAY_E_OOM, "softray_create: out of host memory");
    ctx->device = device_ordinal;
    int rc = [&]() -> int {
        SR_CUDA(ctx, cudaSetDevice(device_ordinal));
        SR_CUDA(ctx, cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal));
        SR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        for (auto& ev : ctx->ev) SR_CUDA(ctx, cudaEventCreate(&ev));
        SR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_staged, cudaEventDisableTiming));
        SR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_frame, cudaEventDisableTiming));
        SR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        SR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        SR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
        SR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_insts, sizeof(DevInstance) * SOFTRAY_MAX_INSTANCES));
        SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_offsets, sizeof(double) * 3 * SOFTRAY_MAX_SHADOW_SAMPLES));
        SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_tile_counter, sizeof(unsigned int)));
        SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_counters, sizeof(DevCounters)));
        SR_CUDA(ctx, cudaMallocHost((void**)&ctx->h_insts, sizeof(DevInstance) * SOFTRAY_MAX_INSTANCES));
        SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_tlas_nodes, sizeof(BvhNode) * SOFTRAY_MAX_INSTANCES));
        SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_tlas_order, sizeof(int32_t) * SOFTRAY_MAX_INSTANCES));
        SR_CUDA(ctx, cudaMallocHost((void**)&ctx->h_tlas_nodes, sizeof(BvhNode) * SOFTRAY_MAX_INSTANCES));
        SR_CUDA(ctx, cudaMallocHost((void**)&ctx->h_tlas_order, sizeof(int32_t) * SOFTRAY_MAX_INSTANCES));
        SR_CUDA(ctx, cudaMallocHost((void**)&ctx->h_offsets, sizeof(double) * 3 * SOFTRAY_MAX_SHADOW_SAMPLES));
        SR_CUDA(ctx, cudaMallocHost((void**)&ctx->h_counters, sizeof(DevCounters)));
        return SOFTRAY_OK;
    }();
    { std::lock_guard<std::mutex> lock(g_live_mutex); g_live_ctx.insert(ctx); }
    if (rc != SOFTRAY_OK) {
        g_last_error = ctx->err;
        softray_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SOFTRAY_OK;
}

// ---------------------------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------------------------
extern "C" void softray_scene_destroy(softray_scene* scene)
{
    if (!scene) return;
    {
        std::lock_guard<std::mutex> lock(g_live_mutex);
        if (!g_live_scenes.erase(scene)) return;        // already released (with its context, or twice)
        std::vector<softray_scene*>& v = scene->ctx->scenes;
        v.erase(std::remove(v.begin(), v.end(), scene), v.end());
    }
    release_scene(scene);
}

static void release_scene(softray_scene* scene)
{
    if (!scene->replicas.empty()) {                     // a group's scene owns its per-device replicas
        for (softray_scene* r : scene->replicas) softray_scene_destroy(r);
        delete scene;
        return;
    }
    if (scene->ctx) {
        cudaSetDevice(scene->ctx->device);
        cudaStreamSynchronize(scene->ctx->stream);
    }
    for (void* p : scene->allocs) cudaFree(p);
    delete scene;
}

static int build_scene(softray_scene* sc, const softray_scene_desc* desc)
{
    softray_ctx* ctx = sc->ctx;
    const bool brute = desc->accel == SOFTRAY_ACCEL_BRUTE;
    const bool lbvh = desc->accel == SOFTRAY_ACCEL_LBVH;
    std::vector<DevMesh> meshes((size_t)desc->n_meshes);
    for (int32_t mi = 0; mi < desc->n_meshes; mi++) {
        const softray_mesh& m = desc->meshes[mi];
        if (m.n_tris < 0 || m.n_verts < 0 || (m.n_tris > 0 && (!m.verts_xyz || !m.tri_vidx || !m.tri_argb)))
            return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_scene_create: mesh arrays missing");
        if (lbvh) {
            // flatten + tree on the device (sr_lbvh.cu); same records, its own tree
            DevMesh& dm = meshes[(size_t)mi];
            std::memset(&dm, 0, sizeof dm);
            dm.n_tris = m.n_tris;
            for (int k = 0; k < 3; k++) {
                dm.bmin[k] = m.bbox_min[k]; dm.bmax[k] = m.bbox_max[k];
                dm.fmin[k] = (float)m.bbox_min[k]; dm.fmax[k] = (float)m.bbox_max[k];
                sc->all_min[k] = std::fmin(sc->all_min[k], m.bbox_min[k]); sc->all_max[k] = std::fmax(sc->all_max[k], m.bbox_max[k]);
            }
            dm.scale = round_up(max_abs3(m.bbox_min, m.bbox_max));
            mesh_bounding_sphere(m, &dm);
            sc->mesh_tris.push_back(m.n_tris);
            { softray_scene::V3 a, b; for (int k = 0; k < 3; k++) { a.v[k] = m.bbox_min[k]; b.v[k] = m.bbox_max[k]; }
              sc->mesh_bmin.push_back(a); sc->mesh_bmax.push_back(b); }
            if (m.n_tris == 0) continue;
            TriRec* d_tris = nullptr; TriFilt* d_filt = nullptr; BvhNode* d_nodes = nullptr;
            int32_t n_nodes = 0, depth = 0; int status = 0;
            const auto t_build = std::chrono::steady_clock::now();
            SR_CUDA(ctx, build_mesh_on_device(m.verts_xyz, m.n_verts, m.tri_vidx, m.tri_argb, m.n_tris, m.bbox_min, m.bbox_max,
                                              traversal_pad(max_abs3(m.bbox_min, m.bbox_max)), lbvh_leaf_max(), ctx->stream, &d_tris, &d_filt, &d_nodes,
                                              &n_nodes, &depth, &status));
            if (env_int("SOFTRAY_BUILD_TIMING", 0))
                std::fprintf(stderr, "softray: mesh %d (%d triangles) flattened + tree built on the device in %.2f ms\n", mi, m.n_tris,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build).count());
            if (status == 1) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_scene_create: vertex index out of range");
            if (status == 2) return fail(ctx, SOFTRAY_E_VERTEX_OUTSIDE_BBOX, "A triangle vertex is outside the bounding box");
            if (status == 3) return fail(ctx, SOFTRAY_E_UNSUPPORTED, "softray_scene_create: device-built tree too deep (use SOFTRAY_ACCEL_BVH)");
            const size_t b_tris = sizeof(TriRec) * (size_t)m.n_tris, b_filt = sizeof(TriFilt) * (size_t)m.n_tris,
                         b_nodes = sizeof(BvhNode) * (size_t)n_nodes;
            sc->allocs.push_back(d_tris); sc->allocs.push_back(d_filt); sc->allocs.push_back(d_nodes);
            sc->alloc_bytes.push_back(b_tris); sc->alloc_bytes.push_back(b_filt); sc->alloc_bytes.push_back(b_nodes);
            sc->device_bytes += b_tris + b_filt + b_nodes;
            sc->unhashed.push_back({d_tris, b_tris}); sc->unhashed.push_back({d_nodes, b_nodes}); sc->unhashed.push_back({d_filt, b_filt});
            dm.tris = d_tris; dm.filt = d_filt; dm.nodes = d_nodes; dm.n_nodes = n_nodes;
            continue;
        }
        std::vector<TriRec> recs((size_t)m.n_tris);
        std::vector<PrimBounds> bounds((size_t)m.n_tris);
        std::atomic<int> bad(0);            // 1: vertex index out of range, 2: vertex outside the bounding box
        parallel_for(m.n_tris, [&](int64_t i0, int64_t i1) {
            for (int64_t i = i0; i < i1; i++) {
                hv v[3];
                for (int k = 0; k < 3; k++) {
                    const int32_t vi = m.tri_vidx[3 * (size_t)i + k];
                    if (vi < 0 || vi >= m.n_verts) { bad = 1; return; }
                    v[k] = hmk(m.verts_xyz[3 * (size_t)vi], m.verts_xyz[3 * (size_t)vi + 1], m.verts_xyz[3 * (size_t)vi + 2]);
                    // SpatialSubdivision ctor: every vertex inside the bounding box (SpatialSubdivision.cs:285-295)
                    if (!box_contains(m.bbox_min, m.bbox_max, v[k])) { int z = 0; bad.compare_exchange_strong(z, 2); return; }
                }
                make_tri_rec(&recs[(size_t)i], v[0], v[1], v[2], m.tri_argb[i], (int32_t)i);
                PrimBounds& b = bounds[(size_t)i];
                const double xs[3] = {v[0].x, v[1].x, v[2].x}, ys[3] = {v[0].y, v[1].y, v[2].y}, zs[3] = {v[0].z, v[1].z, v[2].z};
                b.lo[0] = round_down(std::fmin(xs[0], std::fmin(xs[1], xs[2]))); b.hi[0] = round_up(std::fmax(xs[0], std::fmax(xs[1], xs[2])));
                b.lo[1] = round_down(std::fmin(ys[0], std::fmin(ys[1], ys[2]))); b.hi[1] = round_up(std::fmax(ys[0], std::fmax(ys[1], ys[2])));
                b.lo[2] = round_down(std::fmin(zs[0], std::fmin(zs[1], zs[2]))); b.hi[2] = round_up(std::fmax(zs[0], std::fmax(zs[1], zs[2])));
            }
        });
        if (bad == 1) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_scene_create: vertex index out of range");
        if (bad == 2) return fail(ctx, SOFTRAY_E_VERTEX_OUTSIDE_BBOX, "A triangle vertex is outside the bounding box");
        DevMesh& dm = meshes[(size_t)mi];
        std::memset(&dm, 0, sizeof dm);
        dm.n_tris = m.n_tris;
        for (int k = 0; k < 3; k++) {
            dm.bmin[k] = m.bbox_min[k]; dm.bmax[k] = m.bbox_max[k];
            dm.fmin[k] = (float)m.bbox_min[k]; dm.fmax[k] = (float)m.bbox_max[k];
            sc->all_min[k] = std::fmin(sc->all_min[k], m.bbox_min[k]); sc->all_max[k] = std::fmax(sc->all_max[k], m.bbox_max[k]);
        }
        dm.scale = round_up(max_abs3(m.bbox_min, m.bbox_max));
        mesh_bounding_sphere(m, &dm);
        sc->mesh_tris.push_back(m.n_tris);
        { softray_scene::V3 a, b; for (int k = 0; k < 3; k++) { a.v[k] = m.bbox_min[k]; b.v[k] = m.bbox_max[k]; }
          sc->mesh_bmin.push_back(a); sc->mesh_bmax.push_back(b); }
        if (m.n_tris == 0) continue;
        if (brute) {
            int rc = upload(sc, recs, &dm.tris);
            if (rc != SOFTRAY_OK) return rc;
        } else {
            BvhBuild bvh;
            build_bvh(bounds, traversal_pad(max_abs3(m.bbox_min, m.bbox_max)), kMaxLeafPrims, sah_isect_cost(kTriIsectCost), &bvh);
            if (bvh.depth >= kStackEntries) return fail(ctx, SOFTRAY_E_UNSUPPORTED, "softray_scene_create: BVH too deep");
            std::vector<TriRec> ordered((size_t)m.n_tris);
            std::vector<TriFilt> filt((size_t)m.n_tris);
            parallel_for(m.n_tris, [&](int64_t k0, int64_t k1) {
                for (int64_t k = k0; k < k1; k++) {
                    ordered[(size_t)k] = recs[(size_t)bvh.order[(size_t)k]];
                    make_tri_filt(&filt[(size_t)k], ordered[(size_t)k]);
                }
            });
            std::vector<TriRec>().swap(recs);
            int rc = upload(sc, ordered, &dm.tris);
            if (rc != SOFTRAY_OK) return rc;
            rc = upload(sc, bvh.nodes, &dm.nodes);
            if (rc != SOFTRAY_OK) return rc;
            dm.n_nodes = (int32_t)bvh.nodes.size();
            rc = upload(sc, filt, &dm.filt);
            if (rc != SOFTRAY_OK) return rc;
        }
    }

    DevScene& ds = sc->dev;
    std::memset(&ds, 0, sizeof ds);
    ds.n_meshes = desc->n_meshes;
    ds.accel = desc->accel;
    ds.n_spheres = desc->n_spheres;
    if (desc->n_spheres > 0) {
        std::vector<SphereRec> recs((size_t)desc->n_spheres);
        std::vector<PrimBounds> bounds((size_t)desc->n_spheres);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int32_t i = 0; i < desc->n_spheres; i++) {
            const softray_sphere& s = desc->spheres[i];
            if (!(s.r >= 0.0) || !std::isfinite(s.cx) || !std::isfinite(s.cy) || !std::isfinite(s.cz) || !std::isfinite(s.r))
                return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_scene_create: sphere is not finite");
            SphereRec& r = recs[(size_t)i];
            std::memset(&r, 0, sizeof r);
            r.cx = s.cx; r.cy = s.cy; r.cz = s.cz; r.r = s.r;
            r.r2 = s.r * s.r;                                   // Sphere.cs:30
            r.color = s.argb; r.index = i;
            const double c[3] = {s.cx, s.cy, s.cz};
            for (int k = 0; k < 3; k++) {
                bounds[(size_t)i].lo[k] = round_down(c[k] - s.r);
                bounds[(size_t)i].hi[k] = round_up(c[k] + s.r);
                lo[k] = std::fmin(lo[k], c[k] - s.r); hi[k] = std::fmax(hi[k], c[k] + s.r);
            }
        }
        for (int k = 0; k < 3; k++) {
            ds.sph_bmin[k] = lo[k]; ds.sph_bmax[k] = hi[k];
            sc->all_min[k] = std::fmin(sc->all_min[k], lo[k]); sc->all_max[k] = std::fmax(sc->all_max[k], hi[k]);
        }
        if (brute) {
            int rc = upload(sc, recs, &ds.spheres);
            if (rc != SOFTRAY_OK) return rc;
        } else {
            BvhBuild bvh;
            const float pad = traversal_pad(max_abs3(lo, hi));
            build_bvh(bounds, pad, kMaxLeafPrims, sah_isect_cost(kSphereIsectCost), &bvh);
            if (bvh.depth >= kStackEntries) return fail(ctx, SOFTRAY_E_UNSUPPORTED, "softray_scene_create: BVH too deep");
            std::vector<SphereRec> ordered((size_t)desc->n_spheres);
            for (int32_t k = 0; k < desc->n_spheres; k++) ordered[(size_t)k] = recs[(size_t)bvh.order[(size_t)k]];
            int rc = upload(sc, ordered, &ds.spheres);
            if (rc != SOFTRAY_OK) return rc;
            rc = upload(sc, bvh.nodes, &ds.sphere_nodes);
            if (rc != SOFTRAY_OK) return rc;
            ds.n_sphere_nodes = (int32_t)bvh.nodes.size();
            // the entry clip works on the padded bounds
            for (int k = 0; k < 3; k++) { ds.sph_bmin[k] = lo[k] - (double)pad; ds.sph_bmax[k] = hi[k] + (double)pad; }
            std::vector<float4> filt((size_t)desc->n_spheres);
            for (int32_t k = 0; k < desc->n_spheres; k++) {
                const SphereRec& r = ordered[(size_t)k];
                filt[(size_t)k] = make_float4((float)r.cx, (float)r.cy, (float)r.cz, (float)r.r);
            }
            rc = upload(sc, filt, &ds.sph_filt);
            if (rc != SOFTRAY_OK) return rc;
            for (int k = 0; k < 3; k++) { ds.sph_fmin[k] = round_down(ds.sph_bmin[k]); ds.sph_fmax[k] = round_up(ds.sph_bmax[k]); }
            ds.sph_scale = round_up(max_abs3(ds.sph_bmin, ds.sph_bmax));
        }
    }
    if (!meshes.empty()) {
        // the DevMesh table itself holds device pointers: fold only its layout-relevant fields
        for (const DevMesh& dm : meshes) {
            fnv(&sc->fingerprint, &dm.n_tris, sizeof dm.n_tris);
            fnv(&sc->fingerprint, &dm.n_nodes, sizeof dm.n_nodes);
            fnv(&sc->fingerprint, dm.bmin, sizeof dm.bmin);
            fnv(&sc->fingerprint, dm.bmax, sizeof dm.bmax);
        }
        const uint64_t keep = sc->fingerprint;
        sc->host_meshes = meshes;
        int rc = upload(sc, meshes, &ds.meshes);
        sc->fingerprint = keep;
        if (rc != SOFTRAY_OK) return rc;
    }
    return SOFTRAY_OK;
}

static int group_scene_create(softray_ctx* group, const softray_scene_desc* desc, softray_scene** out);

extern "C" int softray_scene_create(softray_ctx* ctx, const softray_scene_desc* desc, softray_scene** out)
{
    if (!ctx || !desc || !out) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_scene_create: NULL argument");
    *out = nullptr;
    if (!ctx->members.empty()) return group_scene_create(ctx, desc, out);
    if (desc->n_meshes < 0 || desc->n_spheres < 0 || (desc->n_meshes > 0 && !desc->meshes) ||
        (desc->n_spheres > 0 && !desc->spheres))
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_scene_create: bad counts or NULL arrays");
    if (desc->accel != SOFTRAY_ACCEL_BVH && desc->accel != SOFTRAY_ACCEL_BRUTE && desc->accel != SOFTRAY_ACCEL_LBVH)
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_scene_create: unknown accel");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    softray_scene* sc = new (std::nothrow) softray_scene();
    if (!sc) return fail(ctx, SOFTRAY_E_OOM, "softray_scene_create: out of host memory");
    sc->ctx = ctx;
    int rc;
    try {
        rc = build_scene(sc, desc);
    } catch (const std::bad_alloc&) {
        rc = fail(ctx, SOFTRAY_E_OOM, "softray_scene_create: out of host memory");
    }
    if (rc == SOFTRAY_OK) {
        const cudaError_t e = cudaStreamSynchronize(ctx->stream);        // every upload of the scene has landed
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "softray_scene_create: upload");
    }
    if (rc != SOFTRAY_OK) {
        release_scene(sc);
        return rc;
    }
    { std::lock_guard<std::mutex> lock(g_live_mutex); g_live_scenes.insert(sc); ctx->scenes.push_back(sc); }
    *out = sc;
    return SOFTRAY_OK;
}

extern "C" int softray_scene_fingerprint(const softray_scene* scene, uint64_t* out)
{
    if (!scene || !out) return fail(nullptr, SOFTRAY_E_INVALID_ARG, "softray_scene_fingerprint: NULL argument");
    if (!scene->replicas.empty()) return softray_scene_fingerprint(scene->replicas[0], out);     // (replicas are byte copies)
    softray_scene* sc = const_cast<softray_scene*>(scene);
    if (!sc->unhashed.empty()) {           // device-built buffers: read them back once
        SR_CUDA(sc->ctx, cudaSetDevice(sc->ctx->device));
        std::vector<unsigned char> host;
        for (const softray_scene::DevBuf& b : sc->unhashed) {
            host.resize(b.bytes);
            SR_CUDA(sc->ctx, cudaMemcpy(host.data(), b.ptr, b.bytes, cudaMemcpyDeviceToHost));
            fnv(&sc->fingerprint, host.data(), b.bytes);
        }
        sc->unhashed.clear();
    }
    *out = scene->fingerprint;
    return SOFTRAY_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU inside ONE process: a group context (`softray_create(n_gpus)`; reference: the row-block
// fan-out inside one Render(), Renderer.cs:1655-1680).  One member context per device; scenes are built once and
// replicated device-to-device; a frame is cut into interleaved row bands and every device stores its bands into
// the caller's surface.  A single host thread drives all devices: every launch and copy is asynchronous.
// ---------------------------------------------------------------------------------------------
extern "C" int softray_create_multi(int32_t n_devices, softray_ctx** out)
{
    if (!out) return fail(nullptr, SOFTRAY_E_INVALID_ARG, "softray_create_multi: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, SOFTRAY_E_NO_DEVICE, "softray_create_multi: no CUDA device (there is no CPU fallback)");
    if (n_devices < 0 || n_devices > n_dev) return fail(nullptr, SOFTRAY_E_INVALID_ARG, "softray_create_multi: more devices asked for than visible");
    const int n = n_devices == 0 ? n_dev : n_devices;
    softray_ctx* group = new (std::nothrow) softray_ctx();
    if (!group) return fail(nullptr, SOFTRAY_E_OOM, "softray_create_multi: out of host memory");
    { std::lock_guard<std::mutex> lock(g_live_mutex); g_live_ctx.insert(group); }
    for (int i = 0; i < n; i++) {
        softray_ctx* m = nullptr;
        const int rc = softray_create(i, &m);
        if (rc != SOFTRAY_OK) { softray_destroy(group); return rc; }
        group->members.push_back(m);
    }
    group->device = group->members[0]->device;
    group->sm_count = group->members[0]->sm_count;
    // every member may store into member 0's memory (softray_device_alloc on a group: the NVLink gather of
    // softray_render_device is the kernels' own stores, as with the IPC-mapped framebuffer across processes)
    for (int i = 1; i < n; i++) {
        int can = 0;
        cudaSetDevice(i);
        if (cudaDeviceCanAccessPeer(&can, i, 0) == cudaSuccess && can) {
            e = cudaDeviceEnablePeerAccess(0, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
        }
        cudaGetLastError();
        if (!can) group->peer_missing = true;
    }
    *out = group;
    return SOFTRAY_OK;
}

extern "C" int softray_device_count(const softray_ctx* ctx)
{
    if (!ctx) return 0;
    return ctx->members.empty() ? 1 : (int)ctx->members.size();
}

// a byte copy of `src` on another device, pointers patched
static int clone_scene(const softray_scene* src, softray_ctx* dst, softray_scene** out)
{
    *out = nullptr;
    SR_CUDA(dst, cudaSetDevice(dst->device));
    softray_scene* c = new (std::nothrow) softray_scene();
    if (!c) return fail(dst, SOFTRAY_E_OOM, "softray_scene_create: out of host memory");
    c->ctx = dst;
    c->mesh_tris = src->mesh_tris; c->mesh_bmin = src->mesh_bmin; c->mesh_bmax = src->mesh_bmax;
    c->fingerprint = src->fingerprint; c->device_bytes = src->device_bytes;
    std::memcpy(c->all_min, src->all_min, sizeof c->all_min); std::memcpy(c->all_max, src->all_max, sizeof c->all_max);
    auto remap = [&](const void* p) -> void* {
        if (!p) return nullptr;
        for (size_t k = 0; k < src->allocs.size(); k++) if (src->allocs[k] == p) return c->allocs[k];
        return nullptr;
    };
    int rc = [&]() -> int {
        for (size_t k = 0; k < src->allocs.size(); k++) {
            void* d = nullptr;
            SR_CUDA(dst, cudaMalloc(&d, src->alloc_bytes[k]));
            c->allocs.push_back(d); c->alloc_bytes.push_back(src->alloc_bytes[k]);
            SR_CUDA(dst, cudaMemcpyPeerAsync(d, dst->device, src->allocs[k], src->ctx->device, src->alloc_bytes[k], dst->stream));
        }
        c->dev = src->dev;
        c->dev.spheres = static_cast<const SphereRec*>(remap(src->dev.spheres));
        c->dev.sphere_nodes = static_cast<const BvhNode*>(remap(src->dev.sphere_nodes));
        c->dev.sph_filt = static_cast<const float4*>(remap(src->dev.sph_filt));
        c->dev.meshes = static_cast<const DevMesh*>(remap(src->dev.meshes));
        c->host_meshes = src->host_meshes;
        for (DevMesh& m : c->host_meshes) {
            m.tris = static_cast<const TriRec*>(remap(m.tris)); m.filt = static_cast<const TriFilt*>(remap(m.filt));
            m.nodes = static_cast<const BvhNode*>(remap(m.nodes));
        }
        if (!c->host_meshes.empty())
            SR_CUDA(dst, cudaMemcpyAsync(const_cast<DevMesh*>(c->dev.meshes), c->host_meshes.data(), sizeof(DevMesh) * c->host_meshes.size(),
                                         cudaMemcpyHostToDevice, dst->stream));
        SR_CUDA(dst, cudaStreamSynchronize(dst->stream));
        return SOFTRAY_OK;
    }();
    if (rc != SOFTRAY_OK) { release_scene(c); return rc; }
    { std::lock_guard<std::mutex> lock(g_live_mutex); g_live_scenes.insert(c); dst->scenes.push_back(c); }
    *out = c;
    return SOFTRAY_OK;
}

static int group_scene_create(softray_ctx* group, const softray_scene_desc* desc, softray_scene** out)
{
    softray_scene* first = nullptr;
    int rc = softray_scene_create(group->members[0], desc, &first);      // flatten + build + upload once
    if (rc != SOFTRAY_OK) { group->err = group->members[0]->err; return rc; }
    uint64_t fp = 0;
    softray_scene_fingerprint(first, &fp);                               // (folds device-built buffers in before they are copied)
    softray_scene* g = new (std::nothrow) softray_scene();
    if (!g) { softray_scene_destroy(first); return fail(group, SOFTRAY_E_OOM, "softray_scene_create: out of host memory"); }
    g->ctx = group;
    g->replicas.push_back(first);
    g->dev = first->dev; g->fingerprint = first->fingerprint; g->device_bytes = first->device_bytes; g->mesh_tris = first->mesh_tris;
    for (size_t i = 1; i < group->members.size(); i++) {
        softray_scene* r = nullptr;
        rc = clone_scene(first, group->members[i], &r);
        if (rc != SOFTRAY_OK) {
            group->err = group->members[i]->err;
            for (softray_scene* x : g->replicas) softray_scene_destroy(x);
            delete g;
            return rc;
        }
        g->replicas.push_back(r);
    }
    { std::lock_guard<std::mutex> lock(g_live_mutex); g_live_scenes.insert(g); group->scenes.push_back(g); }
    *out = g;
    return SOFTRAY_OK;
}

// ---------------------------------------------------------------------------------------------
// frame
// ---------------------------------------------------------------------------------------------
namespace {

struct Prepared {
    DevFrame f;
    int grid = 0;
    size_t smem = 0;
    int start_row = 0, end_row = -1;
    bool profile = false;              // softray_frame.profile_stages
    bool wave = false;                 // stage kernels (sr_wave.cu) instead of the fused kernel
    bool empty = false;                // no row to trace (start_row > end_row after the clamp): nothing is launched
};

// A sphere reports rayFrac = distance from the ray start to the hit point (Sphere.cs:160,197) and a
// shadow ray is occluded only by rayFrac <= 1.0 (ShadowMethod.cs:171), so a sphere further than 1.0
// from every shadow-ray start of the frame can never occlude.  Conservative: 0 only when proven.
int32_t spheres_can_shadow(const softray_scene* scene, const DevFrame& f, const DevInstance& d, const double* offsets)
{
    const DevScene& ds = scene->dev;
    const double margin = 1e-6;
    if (f.point_lighting) {
        for (int i = 0; i < f.shadow_samples; i++) {
            double dist2 = 0.0;   // squared distance from start_i = light + offset_i to the sphere bounds
            for (int k = 0; k < 3; k++) {
                const double s = d.light_pos_model[k] + offsets[3 * i + k];
                const double g = s < ds.sph_bmin[k] ? ds.sph_bmin[k] - s : (s > ds.sph_bmax[k] ? s - ds.sph_bmax[k] : 0.0);
                dist2 += g * g;
            }
            if (!(std::sqrt(dist2) > 1.0 + margin)) return 1;
        }
        return 0;
    }
    // directional: start = end + dir * 1000 + offset with `end` within 0.001 of some primitive
    double diag2 = 0.0;
    for (int k = 0; k < 3; k++) { const double e = scene->all_max[k] - scene->all_min[k]; diag2 += e * e; }
    const double len = std::sqrt(d.light_dir_model[0] * d.light_dir_model[0] + d.light_dir_model[1] * d.light_dir_model[1] +
                                 d.light_dir_model[2] * d.light_dir_model[2]);
    const double nearest = 1000.0 * len - 0.2 * (1.0 + 1e-9) - 0.001 * (1.0 + 1e-9) - std::sqrt(diag2);
    return nearest > 1.0 + margin ? 0 : 1;
}

int prepare_frame(softray_ctx* ctx, const softray_scene* scene, const softray_frame* fr, Prepared* p)
{
    if (!fr->instances) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: frame.instances is NULL");
    if (fr->width <= 0 || fr->height <= 0 || fr->sub_pixel_res < 1 || fr->n_instances < 1 ||
        fr->n_instances > SOFTRAY_MAX_INSTANCES)
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: bad surface size, sub_pixel_res or instance count");
    if (fr->shadows && (fr->shadow_samples < 1 || fr->shadow_samples > SOFTRAY_MAX_SHADOW_SAMPLES))
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: shadow_samples out of range");
    if (fr->reflection_depth < 0 || fr->reflection_depth > 4)
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: reflection_depth out of range");
    if (fr->band_count > 1 && (fr->band_index < 0 || fr->band_index >= fr->band_count))
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: band_index out of range");
    if (fr->texture3d_id != 0 && fr->texture3d_id != 1)
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: unknown texture3d_id");
    if (fr->filter_mode < SOFTRAY_FILTER_AUTO || fr->filter_mode > SOFTRAY_FILTER_VERIFY)
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: unknown filter_mode");
    if (fr->n_instances > 1 && (scene->dev.n_spheres > 0 || fr->shadows || (fr->focal_blur && fr->sub_pixel_res > 1) ||
                                fr->reflection_depth))
        return fail(ctx, SOFTRAY_E_UNSUPPORTED,
                    "softray_render: composite (multi-instance) frames support primary rays + shading only");

    if (ctx->staging_busy) {   // an asynchronous earlier frame may still be reading h_insts / h_offsets
        SR_CUDA(ctx, cudaEventSynchronize(ctx->ev_staged));
        ctx->staging_busy = false;
    }
    DevFrame& f = p->f;
    std::memset(&f, 0, sizeof f);
    f.ambient = fr->ambient; f.shininess = fr->shininess;
    for (int k = 0; k < 3; k++) { f.light_dir_view[k] = fr->light_dir_view[k]; f.light_pos_view[k] = fr->light_pos_view[k]; }
    f.fov_depth = fr->fov_depth; f.focal_depth = fr->focal_depth; f.focal_strength = fr->focal_strength;
    f.aspect = (double)fr->height / (double)fr->width;                 // Renderer.cs:621
    // shade(): largest |cos| whose shininess-th power is surely below 2^-82 (0: never skip)
    f.spec_skip = (fr->shininess >= 1.0 && std::isfinite(fr->shininess)) ? std::exp2(-82.0 / fr->shininess) * (1.0 - 1e-9) : 0.0;
    f.width = fr->width; f.height = fr->height;
    // clamp rows like Renderer.cs:1652-1653
    int s = fr->start_row < 0 ? 0 : fr->start_row; if (s > fr->height - 1) s = fr->height - 1;
    int e = fr->end_row < 0 ? 0 : fr->end_row;     if (e > fr->height - 1) e = fr->height - 1;
    f.start_row = s; f.end_row = e;
    p->start_row = s; p->end_row = e;
    // start_row > end_row after the clamp: the reference's row loop (Renderer.cs:1659-1670) runs zero times
    p->empty = e < s;
    if (p->empty) { f.tiles_x = 0; f.tiles_y = 0; p->grid = 0; return SOFTRAY_OK; }
    f.sub_pixel_res = fr->sub_pixel_res;
    f.focal_blur = fr->focal_blur ? 1 : 0;
    f.subdivision = fr->subdivision ? 1 : 0;
    f.shading = fr->shading ? 1 : 0;
    f.shadows = fr->shadows ? 1 : 0;
    f.shadow_samples = fr->shadows ? fr->shadow_samples : 0;
    f.point_lighting = fr->point_lighting ? 1 : 0;
    f.specular_lighting = fr->specular_lighting ? 1 : 0;
    f.reflection_depth = fr->reflection_depth;
    f.texture3d_id = fr->texture3d_id;
    f.n_instances = fr->n_instances;
    f.background = fr->background_argb | 0xFF000000u;                  // Renderer.cs:325-331,1860
    const int rows = e - s + 1;
    const bool banded = fr->band_count > 1 && fr->band_height > 0;
    f.band_height = banded ? fr->band_height : rows;
    f.band_count = banded ? fr->band_count : 1;
    f.band_index = banded ? fr->band_index : 0;
    // the FP32 filter needs the BVH layout (filter records live in leaf order)
    f.filter_mode = scene->dev.accel != SOFTRAY_ACCEL_BRUTE ? fr->filter_mode : SOFTRAY_FILTER_OFF;

    if (f.shadows && (ctx->cached_samples != f.shadow_samples || ctx->cached_seed != fr->random_seed)) {
        area_light_offsets(fr->random_seed, f.shadow_samples, ctx->h_offsets);
        ctx->cached_samples = f.shadow_samples; ctx->cached_seed = fr->random_seed;
    }
    if (f.shadows) {
        double r2 = 0.0;
        for (int i = 0; i < f.shadow_samples; i++) {
            const double* o = ctx->h_offsets + 3 * i;
            r2 = std::fmax(r2, o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
        }
        f.light_radius = round_up(std::sqrt(r2) * (1.0 + 1e-6));
        // a cone walk pays off when it replaces many rays and stays cheap: bounded to a few single-ray walks
        f.bundle_budget = f.shadow_samples >= kBundleMinSamples ? env_int("SOFTRAY_BUNDLE_BUDGET", kBundleBudget) : 0;
    }

    // Stage barriers (sr_render.cu "Phase synchronisation") need every thread of a block in every stage: single-
    // instance frames only (a composite frame calls the per-instance search from inside its hierarchy walk).  They pay
    // where instruction fetch, not the walks, bounds the camera rays: a small cache-resident scene (short walks of
    // similar length) under a frame with at least two tiles per resident warp (config1, 512x512: 0.0586 vs 0.0591 ms, no loss).  SOFTRAY_PHASE_SYNC=<mask> overrides.
    {
        long long prims = scene->dev.n_spheres;
        for (int32_t t : scene->mesh_tris) prims += t;
        const long long tiles = (long long)((f.width + 7) / 8) * ((rows + 3) / 4) / (f.band_count > 0 ? f.band_count : 1);
        const bool small_scene = prims <= kPhaseSyncMaxPrims && tiles >= 2LL * ctx->sm_count * (768 / 32);      // 768 resident threads per SM
        // (a composite frame can only have bit 0: tile fetch + start of a camera ray, both outside its hierarchy walk)
        f.phase_sync = fr->n_instances == 1 ? env_int("SOFTRAY_PHASE_SYNC", small_scene ? kPhaseSyncStages : 0)
                                            : (env_int("SOFTRAY_PHASE_SYNC", kPhaseSyncComposite) & 1);
        if (f.phase_sync) f.phase_sync |= 1;           // (the stage barriers need the synchronised tile fetch)
    }

    int32_t base = 0;
    for (int32_t i = 0; i < fr->n_instances; i++) {
        const softray_instance& in = fr->instances[i];
        if (in.mesh_id < 0 || in.mesh_id >= scene->dev.n_meshes)
            return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: instance.mesh_id out of range");
        DevInstance& d = ctx->h_insts[i];
        std::memset(&d, 0, sizeof d);
        std::memcpy(d.M, in.M, sizeof d.M);            // rows 0..2
        std::memcpy(d.Minv, in.Minv, sizeof d.Minv);
        d.pos_z = in.pos[2];
        // start_World = TransformDirectionReverse((0,0,-Position.z)) (Renderer.cs:1717); composite
        // frames use the full inverse transform of the view origin
        const hv st = fr->n_instances == 1 ? hmul3x3(in.Minv, hmk(0, 0, -in.pos[2])) : hmul3x4(in.Minv, hmk(0, 0, 0));
        d.start[0] = st.x; d.start[1] = st.y; d.start[2] = st.z;
        // Renderer.cs:1513-1515 with Instance.TransformPosFromView = Minv * pos (Instance.cs:203)
        const hv ld = hmul3x3(in.Minv, hmk(fr->light_dir_view[0], fr->light_dir_view[1], fr->light_dir_view[2]));
        const hv lp = hmul3x4(in.Minv, hmk(fr->light_pos_view[0], fr->light_pos_view[1], fr->light_pos_view[2]));
        d.light_dir_model[0] = ld.x; d.light_dir_model[1] = ld.y; d.light_dir_model[2] = ld.z;
        d.light_pos_model[0] = lp.x; d.light_pos_model[1] = lp.y; d.light_pos_model[2] = lp.z;
        d.mesh = in.mesh_id;
        d.tri_base = base;
        base += scene->mesh_tris[(size_t)in.mesh_id];
        d.sph_can_shadow = (f.shadows && scene->dev.n_spheres > 0) ? spheres_can_shadow(scene, f, d, ctx->h_offsets) : 0;
        {   // the mesh's bounding sphere in view space (rigid transform: same radius); padded well beyond FP64 rounding
            const DevMesh& hm = scene->host_meshes[(size_t)in.mesh_id];
            const hv cv = hmul3x4(in.M, hmk(hm.bs_center[0], hm.bs_center[1], hm.bs_center[2]));
            d.bs_center_view[0] = cv.x; d.bs_center_view[1] = cv.y; d.bs_center_view[2] = cv.z;
            const double r = hm.bs_radius * (1.0 + 1e-9) + 1e-9 * (1.0 + std::fabs(cv.x) + std::fabs(cv.y) + std::fabs(cv.z));
            d.bs_radius2 = r * r;
        }
    }

    // composite frame: BVH over the view-space boxes of the instances (the reference traces
    // instances one after the other over the whole frame, Renderer.cs:746-755)
    f.tlas_nodes = nullptr; f.tlas_order = nullptr; ctx->tlas_nodes_used = 0;
    if (fr->n_instances > 1 && scene->dev.accel != SOFTRAY_ACCEL_BRUTE) {
        std::vector<PrimBounds> boxes((size_t)fr->n_instances);
        double big = 0.0;
        for (int32_t i = 0; i < fr->n_instances; i++) {
            const softray_instance& in = fr->instances[i];
            const double* mn = scene->mesh_bmin[(size_t)in.mesh_id].v; const double* mx = scene->mesh_bmax[(size_t)in.mesh_id].v;
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int corner = 0; corner < 8; corner++) {
                const hv p = hmk((corner & 1) ? mx[0] : mn[0], (corner & 2) ? mx[1] : mn[1], (corner & 4) ? mx[2] : mn[2]);
                const hv q = hmul3x4(in.M, p);
                const double c[3] = {q.x, q.y, q.z};
                for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], c[k]); hi[k] = std::fmax(hi[k], c[k]); }
            }
            for (int k = 0; k < 3; k++) {
                // 1e-9: the box tolerance of the clip (1e-10) and the FP64 rounding of the transform
                boxes[(size_t)i].lo[k] = round_down(lo[k] - 1e-9 - 1e-12 * std::fabs(lo[k]));
                boxes[(size_t)i].hi[k] = round_up(hi[k] + 1e-9 + 1e-12 * std::fabs(hi[k]));
                big = std::fmax(big, std::fmax(std::fabs(lo[k]), std::fabs(hi[k])));
            }
        }
        BvhBuild tlas;
        build_bvh(boxes, traversal_pad(big), 2, 4.0, &tlas);
        if ((int)tlas.nodes.size() > SOFTRAY_MAX_INSTANCES || tlas.depth >= kTlasStackEntries)
            return fail(ctx, SOFTRAY_E_UNSUPPORTED, "softray_render: instance hierarchy too large");
        std::memcpy(ctx->h_tlas_nodes, tlas.nodes.data(), tlas.nodes.size() * sizeof(BvhNode));
        std::memcpy(ctx->h_tlas_order, tlas.order.data(), tlas.order.size() * sizeof(int32_t));
        ctx->tlas_nodes_used = (int32_t)tlas.nodes.size();
        f.tlas_nodes = ctx->d_tlas_nodes; f.tlas_order = ctx->d_tlas_order;
    }

    // one warp per 8x4-pixel tile, pulled from an atomic queue by persistent warps
    const int n_bands = (rows + f.band_height - 1) / f.band_height;
    const int my_bands = f.band_index < n_bands ? (n_bands - f.band_index + f.band_count - 1) / f.band_count : 0;
    f.tiles_x = (f.width + 7) / 8;
    f.tiles_per_band = (f.band_height + 3) / 4;
    f.tiles_y = my_bands * f.tiles_per_band;
    f.fd_n = make_fastdiv((uint32_t)f.sub_pixel_res); f.fd_nn = make_fastdiv((uint32_t)(f.sub_pixel_res * f.sub_pixel_res));
    f.fd_per_tile = make_fastdiv(32u * (uint32_t)(f.sub_pixel_res * f.sub_pixel_res));
    f.fd_tiles_x = make_fastdiv((uint32_t)f.tiles_x); f.fd_tiles_per_band = make_fastdiv((uint32_t)f.tiles_per_band);
    p->smem = sizeof(DevInstance) * (size_t)f.n_instances + sizeof(double) * 3 * (size_t)f.shadow_samples;
    {   // Small sphere sets are staged in shared memory by every block of the fused kernel (single-instance frames, filter
        // on) -- as long as that does not cost a resident block: config2's 1000 spheres (80 KB of tree + records) measured
        // 0.526 -> 0.692 ms staged (2 instead of 3 blocks per SM, and L1 already serves the walk), so the default limit
        // is 32 KB (~350 spheres); SOFTRAY_STAGE_SPHERES_MAX overrides it.
        const size_t stage_bytes = sizeof(BvhNode) * (size_t)scene->dev.n_sphere_nodes + sizeof(float4) * (size_t)scene->dev.n_spheres;
        f.stage_spheres = (scene->dev.n_spheres > kTinyMesh && scene->dev.sphere_nodes != nullptr && scene->dev.sph_filt != nullptr && fr->n_instances == 1 &&
                           f.filter_mode != SOFTRAY_FILTER_OFF && stage_bytes <= (size_t)env_int("SOFTRAY_STAGE_SPHERES_MAX", 32 * 1024)) ? 1 : 0;
        if (f.stage_spheres) p->smem = ((p->smem + 63) & ~(size_t)63) + stage_bytes;
    }
    int occ = render_kernel_occupancy((int)p->smem, f.stage_spheres != 0);
    if (occ < 1) occ = 1;
    { const int cap = env_int("SOFTRAY_BLOCKS_PER_SM", 0); if (cap > 0 && cap < occ) occ = cap; }   // experiments
    const long long n_tiles = (long long)f.tiles_x * f.tiles_y;
    long long grid = (long long)ctx->sm_count * occ;
    const int warps_per_block = render_kernel_block_threads() / 32;
    const long long need = (n_tiles + warps_per_block - 1) / warps_per_block;     // one tile per warp at a time
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    p->grid = (int)grid;

    // Which pipeline.  The stage kernels cover what the big frames use: library-built trees, the filtered search,
    // meshes only.  Small cache-resident scenes stay on the fused kernel (one launch, stage barriers); everything the
    // stage kernels do not implement (spheres, brute force, FILTER_OFF / _VERIFY) does too.  SOFTRAY_PIPELINE=fused|wave
    // overrides the size rule, never the capability rule.
    {
        long long prims = scene->dev.n_spheres;
        for (int32_t t : scene->mesh_tris) prims += t;
        const bool capable = scene->dev.accel != SOFTRAY_ACCEL_BRUTE && f.filter_mode == SOFTRAY_FILTER_AUTO && scene->dev.n_spheres == 0 &&
                             scene->dev.n_meshes > 0;
        const char* e = std::getenv("SOFTRAY_PIPELINE");
        bool want = prims > kPhaseSyncMaxPrims || fr->n_instances > 1;
        if (e && !std::strcmp(e, "fused")) want = false;
        if (e && !std::strcmp(e, "wave")) want = true;
        p->wave = capable && want;
        p->profile = fr->profile_stages != 0;
    }
    return SOFTRAY_OK;
}

// device-side alias of a page-locked host buffer, or nullptr if `host` is pageable / not mappable
template <typename T>
T* zero_copy_pointer(T* host)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return nullptr;
    return static_cast<T*>(a.devicePointer);
}

int enqueue_frame(softray_ctx* ctx, const softray_scene* scene, const Prepared& p, uint32_t* d_pixels, int32_t* d_ids,
                  cudaStream_t stream, bool timed, uint32_t* h_pixels = nullptr, int32_t* h_ids = nullptr)
{
    const DevFrame& f = p.f;
    if (ctx->have_last_frame && ctx->last_stream != stream) SR_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_frame, 0));
    if (timed) SR_CUDA(ctx, cudaEventRecord(ctx->ev[0], stream));
    SR_CUDA(ctx, cudaMemcpyAsync(ctx->d_insts, ctx->h_insts, sizeof(DevInstance) * (size_t)f.n_instances,
                                 cudaMemcpyHostToDevice, stream));
    if (f.shadows)
        SR_CUDA(ctx, cudaMemcpyAsync(ctx->d_offsets, ctx->h_offsets, sizeof(double) * 3 * (size_t)f.shadow_samples,
                                     cudaMemcpyHostToDevice, stream));
    if (ctx->tlas_nodes_used > 0) {
        SR_CUDA(ctx, cudaMemcpyAsync(ctx->d_tlas_nodes, ctx->h_tlas_nodes, sizeof(BvhNode) * (size_t)ctx->tlas_nodes_used,
                                     cudaMemcpyHostToDevice, stream));
        SR_CUDA(ctx, cudaMemcpyAsync(ctx->d_tlas_order, ctx->h_tlas_order, sizeof(int32_t) * (size_t)f.n_instances,
                                     cudaMemcpyHostToDevice, stream));
    }
    SR_CUDA(ctx, cudaEventRecord(ctx->ev_staged, stream));
    ctx->staging_busy = true;
    SR_CUDA(ctx, cudaMemsetAsync(ctx->d_tile_counter, 0, sizeof(unsigned int), stream));
    SR_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(DevCounters), stream));
    if (timed) SR_CUDA(ctx, cudaEventRecord(ctx->ev[1], stream));
    ctx->last_launches = 1;
    if (p.wave) {
        const int nn = f.sub_pixel_res * f.sub_pixel_res;
        const int slots = 1 + ((f.reflection_depth > 0 && f.n_instances == 1) ? f.reflection_depth : 0);
        // samples per chunk: SOFTRAY_WAVE_CHUNK (default 8 M), at most half the frame (two chunks in flight), whole tiles
        const long long per_tile = 32LL * nn, frame_tiles = (long long)f.tiles_x * f.tiles_y;
        long long cap_tiles = env_int("SOFTRAY_WAVE_CHUNK", 1 << 23) / per_tile;
        if (cap_tiles > (frame_tiles + 1) / 2) cap_tiles = (frame_tiles + 1) / 2;
        if (cap_tiles < 1) cap_tiles = 1;
        const long long cap = cap_tiles * per_tile;
        if (ctx->wave_cap < (uint32_t)cap || ctx->wave_slots < slots) {
            // (grow only: frames of one host keep their size; a bigger frame reallocates once)
            SR_CUDA(ctx, cudaStreamSynchronize(stream));
            if (ctx->stream != stream) SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const uint32_t new_cap = (uint32_t)std::max<long long>(cap, ctx->wave_cap);
            const int new_slots = std::max(slots, ctx->wave_slots);
            cudaFree(ctx->wave_base); ctx->wave_base = nullptr; ctx->wave_cap = 0; ctx->wave_bytes = 0;
            WaveLayout lay;
            const size_t bytes = wave_buffer_bytes(new_cap, new_slots, &lay);
            SR_CUDA(ctx, cudaMalloc(&ctx->wave_base, 2 * bytes));
            wave_bind(ctx->wave_base, lay, &ctx->wave[0]);
            wave_bind(static_cast<unsigned char*>(ctx->wave_base) + bytes, lay, &ctx->wave[1]);
            ctx->wave_bytes = 2 * bytes; ctx->wave_cap = new_cap; ctx->wave_slots = new_slots;
        }
        int launches = 0;
        HostCopy hc;
        if (h_pixels) {
            hc.h_pixels = h_pixels; hc.h_ids = h_ids; hc.d_pixels = d_pixels; hc.d_ids = d_ids;
            hc.copy_stream = ctx->copy_stream; hc.events = &ctx->copy_events;
        }
        ctx->stage_timer.on = timed && p.profile;
        SR_CUDA(ctx, wave_render(f, scene->dev, ctx->d_insts, ctx->d_offsets, ctx->wave, (uint32_t)cap, d_pixels, d_ids, ctx->d_counters,
                                 ctx->sm_count, stream, ctx->side_stream, ctx->ev_fork, ctx->ev_join, &launches,
                                 (timed && p.profile) ? &ctx->stage_timer : nullptr, h_pixels ? &hc : nullptr));
        ctx->last_launches = launches;
    } else
    SR_CUDA(ctx, launch_render(f, scene->dev, ctx->d_insts, ctx->d_offsets, d_pixels, d_ids, ctx->d_tile_counter,
                               ctx->d_counters, p.grid, stream));
    if (timed) SR_CUDA(ctx, cudaEventRecord(ctx->ev[2], stream));
    SR_CUDA(ctx, cudaEventRecord(ctx->ev_frame, stream));
    ctx->last_stream = stream; ctx->have_last_frame = true;
    return SOFTRAY_OK;
}

int collect_stats(softray_ctx* ctx, cudaStream_t stream, softray_stats* st, bool have_d2h_event)
{
    SR_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(DevCounters), cudaMemcpyDeviceToHost, stream));
    SR_CUDA(ctx, cudaStreamSynchronize(stream));
    std::memset(st, 0, sizeof *st);
    const DevCounters& c = *ctx->h_counters;
    st->rays_primary = c.rays_primary; st->rays_shadow = c.rays_shadow; st->rays_secondary = c.rays_secondary;
    st->node_visits = c.node_visits; st->prim_tests = c.prim_tests; st->sphere_tests = c.sphere_tests;
    st->hits_primary = c.hits_primary; st->shaded_hits = c.shaded_hits;
    st->filter_tests = c.filter_tests; st->filter_unsure = c.filter_unsure; st->filter_mismatch = c.filter_mismatch;
    st->rays_bundled = c.rays_bundled; st->rays_fallback = c.rays_fallback; st->rays_short_listed = c.rays_short_listed;
    st->launches = (uint64_t)ctx->last_launches;
    if (ctx->stage_timer.on) { ctx->stage_timer.collect(st->ms_stage, SOFTRAY_N_STAGES); ctx->stage_timer.on = false; }
    float ms = 0.f;
    SR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); st->ms_h2d = ms;
    SR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); st->ms_kernel = ms;
    if (have_d2h_event) { SR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); st->ms_d2h = ms; }
    return SOFTRAY_OK;
}

}  // namespace

namespace {

// One frame on one device, split so that a group context can start every member before it waits for any:
// render_*_begin enqueues everything (asynchronous), render_end waits and reads the counters back.
struct InFlight { bool launched = false; bool host_path = false; cudaStream_t s = nullptr; };

int render_device_begin(softray_ctx* ctx, const softray_scene* scene, const softray_frame* frame, uint32_t* d_pixels_argb,
                        int32_t* d_hit_ids, cudaStream_t stream, bool timed, InFlight* fl)
{
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    fl->s = stream ? stream : ctx->stream;
    Prepared p;
    int rc = prepare_frame(ctx, scene, frame, &p);
    if (rc != SOFTRAY_OK) return rc;
    if (p.empty || p.f.tiles_y == 0) return SOFTRAY_OK;          // no row of the frame is ours: nothing to launch
    rc = enqueue_frame(ctx, scene, p, d_pixels_argb, d_hit_ids, fl->s, timed);
    if (rc != SOFTRAY_OK) return rc;
    fl->launched = true;
    return SOFTRAY_OK;
}

int render_host_begin(softray_ctx* ctx, const softray_scene* scene, const softray_frame* frame, uint32_t* pixels_argb, int32_t* hit_ids,
                      InFlight* fl)
{
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    fl->s = ctx->stream; fl->host_path = true;
    Prepared p;
    int rc = prepare_frame(ctx, scene, frame, &p);
    if (rc != SOFTRAY_OK) return rc;
    if (p.empty || p.f.tiles_y == 0) return SOFTRAY_OK;
    const size_t n_px = (size_t)frame->width * (size_t)frame->height;
    // A page-locked (CUDA-registered / cudaHostAlloc'd) caller buffer is mapped into the device's address
    // space: the kernel then stores finished pixels straight into it over PCIe, overlapped with tracing,
    // and there is no device framebuffer and no D2H copy.  Pageable memory takes the staging path below.
    uint32_t* zc_pixels = zero_copy_pointer(pixels_argb);
    int32_t* zc_ids = hit_ids ? zero_copy_pointer(hit_ids) : nullptr;
    const bool zero_copy = zc_pixels != nullptr && (!hit_ids || zc_ids != nullptr) && !env_int("SOFTRAY_NO_ZERO_COPY", 0);
    cudaStream_t s = ctx->stream;
    // A big frame of the stage-kernel pipeline leaves chunk by chunk through the copy engines instead: its compose
    // kernels write HBM at HBM speed and the DMA of chunk i runs under the tracing of chunk i + 1 (kernels storing
    // 132 MB over PCIe themselves kept their blocks resident for the length of the transfer: config5 e2e 15.2 ms
    // against 12.1 on the device).
    const int nn_ = p.f.sub_pixel_res * p.f.sub_pixel_res;
    const bool chunk_dma = zero_copy && p.wave && !p.profile && n_px >= ((size_t)1 << 21) && p.f.tiles_y >= 4 && !env_int("SOFTRAY_NO_CHUNK_DMA", 0) &&
                           (long long)p.f.tiles_x * 32 * nn_ <= (long long)env_int("SOFTRAY_WAVE_CHUNK", 1 << 23);
    if (zero_copy && !chunk_dma) {
        rc = enqueue_frame(ctx, scene, p, zc_pixels, zc_ids, s, true);
        if (rc != SOFTRAY_OK) return rc;
    } else {
        if (ctx->fb_capacity < n_px) {
            cudaFree(ctx->d_pixels); ctx->d_pixels = nullptr; ctx->fb_capacity = 0;
            SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_pixels, n_px * sizeof(uint32_t)));
            ctx->fb_capacity = n_px;
        }
        if (hit_ids && ctx->ids_capacity < n_px) {
            cudaFree(ctx->d_ids); ctx->d_ids = nullptr; ctx->ids_capacity = 0;
            SR_CUDA(ctx, cudaMalloc((void**)&ctx->d_ids, n_px * sizeof(int32_t)));
            ctx->ids_capacity = n_px;
        }
        rc = enqueue_frame(ctx, scene, p, ctx->d_pixels, hit_ids ? ctx->d_ids : nullptr, s, true, chunk_dma ? pixels_argb : nullptr,
                           chunk_dma ? hit_ids : nullptr);
        if (rc != SOFTRAY_OK) return rc;
        // read back exactly the rows that were rendered; banded frames copy each
        // band of this rank separately so the caller's other rows stay untouched
        const size_t W = (size_t)frame->width;
        const bool banded = p.f.band_count > 1;
        const int bh = banded ? p.f.band_height : (p.end_row - p.start_row + 1);
        for (int top = p.start_row, b = 0; top <= p.end_row && !chunk_dma; top += bh, b++) {
            if (banded && b % p.f.band_count != p.f.band_index) continue;
            const int last = top + bh - 1 > p.end_row ? p.end_row : top + bh - 1;
            const size_t off = (size_t)top * W, cnt = (size_t)(last - top + 1) * W;
            SR_CUDA(ctx, cudaMemcpyAsync(pixels_argb + off, ctx->d_pixels + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            if (hit_ids)
                SR_CUDA(ctx, cudaMemcpyAsync(hit_ids + off, ctx->d_ids + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        }
    }
    SR_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));
    fl->launched = true;
    return SOFTRAY_OK;
}

// wait == false (softray_render_device without stats): leave the frame in flight
int render_end(softray_ctx* ctx, const InFlight& fl, softray_stats* stats, bool wait)
{
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (!fl.launched) return SOFTRAY_OK;
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stats) return collect_stats(ctx, fl.s, stats, fl.host_path);
    if (wait) SR_CUDA(ctx, cudaStreamSynchronize(fl.s));
    return SOFTRAY_OK;
}

// ---- a frame on a group context: interleaved row bands, one set per member device ---------------------------
// The reference fans the rows of one Render() out to rayTraceConcurrency tasks that all store into the one
// surface.Pixels (Renderer.cs:1655-1680); here the tasks are the GPUs of the box.
int group_band_height(const softray_frame* fr, int n)
{
    int s = fr->start_row < 0 ? 0 : fr->start_row; if (s > fr->height - 1) s = fr->height - 1;
    int e = fr->end_row < 0 ? 0 : fr->end_row;     if (e > fr->height - 1) e = fr->height - 1;
    const int rows = e - s + 1;
    int bh = rows / (n * 64);                         // ~64 bands per device: expensive rows come in strips, spread them
    bh = bh / 4 * 4;                                  // whole 4-row tiles
    return bh < 4 ? 4 : bh;
}

void add_stats(softray_stats* total, const softray_stats& st)
{
    total->rays_primary += st.rays_primary; total->rays_shadow += st.rays_shadow; total->rays_secondary += st.rays_secondary;
    total->node_visits += st.node_visits; total->prim_tests += st.prim_tests; total->sphere_tests += st.sphere_tests;
    total->hits_primary += st.hits_primary; total->shaded_hits += st.shaded_hits; total->launches += st.launches;
    total->filter_tests += st.filter_tests; total->filter_unsure += st.filter_unsure; total->filter_mismatch += st.filter_mismatch;
    total->rays_bundled += st.rays_bundled; total->rays_fallback += st.rays_fallback; total->rays_short_listed += st.rays_short_listed;
    total->ms_kernel = std::fmax(total->ms_kernel, st.ms_kernel); total->ms_h2d = std::fmax(total->ms_h2d, st.ms_h2d);
    total->ms_d2h = std::fmax(total->ms_d2h, st.ms_d2h);
    for (int k = 0; k < SOFTRAY_N_STAGES; k++) total->ms_stage[k] = std::fmax(total->ms_stage[k], st.ms_stage[k]);
}

int group_render(softray_ctx* group, const softray_scene* scene, const softray_frame* frame, uint32_t* pixels, int32_t* ids, bool host,
                 cudaStream_t stream0, softray_stats* stats)
{
    const auto t0 = std::chrono::steady_clock::now();
    const int n = (int)group->members.size();
    if ((int)scene->replicas.size() != n) return fail(group, SOFTRAY_E_INVALID_ARG, "softray_render: scene was not created in this group context");
    if (frame->band_count > 1) return fail(group, SOFTRAY_E_INVALID_ARG, "softray_render: a group context partitions the rows itself (band_count must be <= 1)");
    if (frame->width <= 0 || frame->height <= 0) return fail(group, SOFTRAY_E_INVALID_ARG, "softray_render: bad surface size");
    const int bh = group_band_height(frame, n);
    std::vector<InFlight> fl((size_t)n);
    int rc = SOFTRAY_OK;
    for (int i = 0; i < n && rc == SOFTRAY_OK; i++) {
        softray_frame fi = *frame;
        if (n > 1) { fi.band_height = bh; fi.band_count = n; fi.band_index = i; }
        softray_ctx* m = group->members[(size_t)i];
        rc = host ? render_host_begin(m, scene->replicas[(size_t)i], &fi, pixels, ids, &fl[(size_t)i])
                  : render_device_begin(m, scene->replicas[(size_t)i], &fi, pixels, ids, i == 0 ? stream0 : nullptr, stats != nullptr, &fl[(size_t)i]);
        if (rc != SOFTRAY_OK) group->err = m->err;
    }
    // (wait for whatever was started, also after an error: the caller's buffers must be quiet when we return)
    if (stats) std::memset(stats, 0, sizeof *stats);
    for (int i = 0; i < n; i++) {
        softray_stats st;
        const int rc2 = render_end(group->members[(size_t)i], fl[(size_t)i], stats ? &st : nullptr, true);
        if (rc2 != SOFTRAY_OK && rc == SOFTRAY_OK) { rc = rc2; group->err = group->members[(size_t)i]->err; }
        if (stats && rc2 == SOFTRAY_OK) add_stats(stats, st);
    }
    if (stats) stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

}  // namespace

extern "C" int softray_render_device(softray_ctx* ctx, const softray_scene* scene, const softray_frame* frame,
                                     uint32_t* d_pixels_argb, int32_t* d_hit_ids, void* stream, softray_stats* stats)
{
    if (!ctx || !scene || !frame || !d_pixels_argb) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render_device: NULL argument");
    if (scene->ctx != ctx) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render_device: scene belongs to another context");
    if (!frame->instances) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: frame.instances is NULL");
    if (!ctx->members.empty())          // (synchronous on a group: every member has finished its bands on return)
        return group_render(ctx, scene, frame, d_pixels_argb, d_hit_ids, false, static_cast<cudaStream_t>(stream), stats);
    const auto t0 = std::chrono::steady_clock::now();
    InFlight fl;
    int rc = render_device_begin(ctx, scene, frame, d_pixels_argb, d_hit_ids, static_cast<cudaStream_t>(stream), stats != nullptr, &fl);
    if (rc != SOFTRAY_OK) return rc;
    rc = render_end(ctx, fl, stats, false);
    if (rc == SOFTRAY_OK && stats) stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

extern "C" int softray_render(softray_ctx* ctx, const softray_scene* scene, const softray_frame* frame,
                              uint32_t* pixels_argb, int32_t* hit_ids, softray_stats* stats)
{
    if (!ctx || !scene || !frame || !pixels_argb) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: NULL argument");
    if (scene->ctx != ctx) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: scene belongs to another context");
    if (!frame->instances) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_render: frame.instances is NULL");
    if (!ctx->members.empty()) return group_render(ctx, scene, frame, pixels_argb, hit_ids, true, nullptr, stats);
    const auto t0 = std::chrono::steady_clock::now();
    InFlight fl;
    int rc = render_host_begin(ctx, scene, frame, pixels_argb, hit_ids, &fl);
    if (rc != SOFTRAY_OK) return rc;
    rc = render_end(ctx, fl, stats, true);
    if (rc == SOFTRAY_OK && stats) stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: peer-mapped framebuffer
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == SOFTRAY_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int softray_device_alloc(softray_ctx* ctx, uint64_t bytes, void** d_ptr_out)
{
    if (ctx && !ctx->members.empty()) return softray_device_alloc(ctx->members[0], bytes, d_ptr_out);
    if (!ctx || !d_ptr_out || bytes == 0) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_device_alloc: bad argument");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    SR_CUDA(ctx, cudaMalloc(d_ptr_out, (size_t)bytes));
    return SOFTRAY_OK;
}

extern "C" int softray_device_free(softray_ctx* ctx, void* d_ptr)
{
    if (ctx && !ctx->members.empty()) return softray_device_free(ctx->members[0], d_ptr);
    if (!ctx) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_device_free: NULL context");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    SR_CUDA(ctx, cudaFree(d_ptr));
    return SOFTRAY_OK;
}

extern "C" int softray_ipc_export(softray_ctx* ctx, void* d_ptr, char handle[SOFTRAY_IPC_HANDLE_BYTES])
{
    if (ctx && !ctx->members.empty()) return softray_ipc_export(ctx->members[0], d_ptr, handle);
    if (!ctx || !d_ptr || !handle) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_ipc_export: NULL argument");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    SR_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle, &h, sizeof h);
    return SOFTRAY_OK;
}

extern "C" int softray_ipc_open(softray_ctx* ctx, const char handle[SOFTRAY_IPC_HANDLE_BYTES], void** d_ptr_out)
{
    if (ctx && !ctx->members.empty()) return softray_ipc_open(ctx->members[0], handle, d_ptr_out);
    if (!ctx || !handle || !d_ptr_out) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_ipc_open: NULL argument");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    SR_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SOFTRAY_OK;
}

extern "C" int softray_host_register(softray_ctx* ctx, void* host_ptr, uint64_t n_bytes)
{
    if (ctx && !ctx->members.empty()) return softray_host_register(ctx->members[0], host_ptr, n_bytes);
    if (!ctx || !host_ptr || n_bytes == 0) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_host_register: NULL argument");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    SR_CUDA(ctx, cudaHostRegister(host_ptr, (size_t)n_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return SOFTRAY_OK;
}

extern "C" int softray_host_unregister(softray_ctx* ctx, void* host_ptr)
{
    if (!ctx || !host_ptr) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_host_unregister: NULL argument");
    if (!ctx->members.empty()) {
        for (size_t i = 1; i < ctx->members.size(); i++) { cudaSetDevice(ctx->members[i]->device); cudaStreamSynchronize(ctx->members[i]->stream); }
        return softray_host_unregister(ctx->members[0], host_ptr);
    }
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->stream) SR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SR_CUDA(ctx, cudaHostUnregister(host_ptr));
    return SOFTRAY_OK;
}

extern "C" int softray_host_barrier(volatile uint32_t* w, uint32_t n_ranks)
{
    if (!w || n_ranks == 0) return fail(nullptr, SOFTRAY_E_INVALID_ARG, "softray_host_barrier: NULL argument");
    uint32_t* count = const_cast<uint32_t*>(w);
    uint32_t* generation = const_cast<uint32_t*>(w) + 1;
    const uint32_t gen = __atomic_load_n(generation, __ATOMIC_ACQUIRE);
    if (__atomic_add_fetch(count, 1u, __ATOMIC_ACQ_REL) == n_ranks) {
        __atomic_store_n(count, 0u, __ATOMIC_RELAXED);
        __atomic_add_fetch(generation, 1u, __ATOMIC_RELEASE);
    } else {
        unsigned long long spins = 0;
        const auto t0 = std::chrono::steady_clock::now();
        const double limit_s = (double)env_int("SOFTRAY_BARRIER_TIMEOUT_S", 60);
        while (__atomic_load_n(generation, __ATOMIC_ACQUIRE) == gen) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            if ((++spins & 0xfffff) == 0) {
                std::this_thread::yield();                               // a rank that lost its core must not starve the others
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s)
                    return fail(nullptr, SOFTRAY_E_TIMEOUT, "softray_host_barrier: a rank did not arrive");
            }
        }
    }
    return SOFTRAY_OK;
}

extern "C" int softray_ipc_close(softray_ctx* ctx, void* d_ptr)
{
    if (ctx && !ctx->members.empty()) return softray_ipc_close(ctx->members[0], d_ptr);
    if (!ctx || !d_ptr) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_ipc_close: NULL argument");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    SR_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return SOFTRAY_OK;
}

// ---------------------------------------------------------------------------------------------
// resolve: PostProcessImage + AntiAliasImage (Renderer.cs:819-898,937-978)
// ---------------------------------------------------------------------------------------------
static int check_resolve(softray_ctx* ctx, const void* src, const void* dst, int32_t w, int32_t h, int32_t aa, int32_t style)
{
    if (!ctx || !src || !dst) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_resolve: NULL argument");
    if (w <= 0 || h <= 0 || aa < 1 || aa > 64 || (int64_t)w * aa > 65535 * 4 || (int64_t)h * aa > 65535 * 4)
        return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_resolve: bad surface size or AntiAliasResolution");
    if (style != SOFTRAY_STYLE_STANDARD && style != SOFTRAY_STYLE_COLOR_SHUFFLE && style != SOFTRAY_STYLE_NEGATIVE)
        return fail(ctx, SOFTRAY_E_UNSUPPORTED, "softray_resolve: the depth styles need the rasteriser's depth buffers");
    return SOFTRAY_OK;
}

extern "C" int softray_resolve_device(softray_ctx* ctx, const uint32_t* d_src, int32_t w, int32_t h, int32_t aa, int32_t style,
                                      uint32_t background, uint32_t* d_dst, void* stream)
{
    if (ctx && !ctx->members.empty()) return softray_resolve_device(ctx->members[0], d_src, w, h, aa, style, background, d_dst, stream);
    int rc = check_resolve(ctx, d_src, d_dst, w, h, aa, style);
    if (rc != SOFTRAY_OK) return rc;
    if (aa > 1 && d_src == d_dst) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_resolve: dst may alias src only when aa_res == 1");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    SR_CUDA(ctx, launch_resolve(d_src, d_dst, w, h, aa, style, background, stream ? static_cast<cudaStream_t>(stream) : ctx->stream));
    return SOFTRAY_OK;
}

extern "C" int softray_resolve(softray_ctx* ctx, const uint32_t* src, int32_t w, int32_t h, int32_t aa, int32_t style,
                               uint32_t background, uint32_t* dst)
{
    if (ctx && !ctx->members.empty()) return softray_resolve(ctx->members[0], src, w, h, aa, style, background, dst);
    int rc = check_resolve(ctx, src, dst, w, h, aa, style);
    if (rc != SOFTRAY_OK) return rc;
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n_dst = (size_t)w * (size_t)h, n_src = n_dst * (size_t)aa * (size_t)aa;
    uint32_t *d_src = nullptr, *d_dst = nullptr;
    SR_CUDA(ctx, cudaMalloc((void**)&d_src, n_src * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc((void**)&d_dst, n_dst * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_src, src, n_src * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_resolve(d_src, d_dst, w, h, aa, style, background, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, d_dst, n_dst * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_src); cudaFree(d_dst);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "softray_resolve");
    return SOFTRAY_OK;
}

// ---------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------
extern "C" int softray_measure_fma_peak(softray_ctx* ctx, int32_t fp64, double* tflops_out)
{
    if (ctx && !ctx->members.empty()) return softray_measure_fma_peak(ctx->members[0], fp64, tflops_out);
    if (!ctx || !tflops_out) return fail(ctx, SOFTRAY_E_INVALID_ARG, "softray_measure_fma_peak: NULL argument");
    SR_CUDA(ctx, cudaSetDevice(ctx->device));
    SR_CUDA(ctx, measure_fma_peak(fp64 != 0, ctx->sm_count, ctx->stream, tflops_out));
    return SOFTRAY_OK;
}

// ---------------------------------------------------------------------------------------------
// helpers mirroring small reference functions
// ---------------------------------------------------------------------------------------------
extern "C" void softray_instance_init(softray_instance* inst, const double pos[3], double yaw, double pitch, double roll,
                                      int32_t mesh_id)
{
    if (!inst || !pos) return;
    // _transform = T(pos) * Roll * Pitch * Yaw; _inverseTransform = Yaw(-) * Pitch(-) * Roll(-) * T(-pos)
    // (Instance.cs:134-135)
    double t[16], r[16], p[16], y[16], a[16], b[16];
    mat_translate(t, pos[0], pos[1], pos[2]); mat_roll(r, roll); mat_pitch(p, pitch); mat_yaw(y, yaw);
    mat_mul(t, r, a); mat_mul(a, p, b); mat_mul(b, y, inst->M);
    mat_yaw(y, -yaw); mat_pitch(p, -pitch); mat_roll(r, -roll); mat_translate(t, -pos[0], -pos[1], -pos[2]);
    mat_mul(y, p, a); mat_mul(a, r, b); mat_mul(b, t, inst->Minv);
    inst->pos[0] = pos[0]; inst->pos[1] = pos[1]; inst->pos[2] = pos[2];
    inst->mesh_id = mesh_id;
    inst->_pad = 0;
}

extern "C" void softray_frame_defaults(softray_frame* f, int32_t width, int32_t height)
{
    if (!f) return;
    std::memset(f, 0, sizeof *f);
    f->ambient = 0.1;                                                // Renderer.cs:207-217
    const hv d = hnormalise(hmk(-1, -1, 1));
    f->light_dir_view[0] = d.x; f->light_dir_view[1] = d.y; f->light_dir_view[2] = d.z;
    const hv lp = hsub(hmk(0.0, 0.0, 1.5), hscale(d, 2));
    f->light_pos_view[0] = lp.x; f->light_pos_view[1] = lp.y; f->light_pos_view[2] = lp.z;
    f->shininess = 100.0;
    const double fov_rad = 45.0 / 180.0 * 3.14159265358979323846;    // Renderer.cs:97-101
    f->fov_depth = 0.5 / std::tan(fov_rad / 2);
    f->focal_depth = 1.5; f->focal_strength = 10.0;                  // Renderer.cs:79-83
    f->width = width; f->height = height;
    f->start_row = 0; f->end_row = height - 1;
    f->sub_pixel_res = 1; f->focal_blur = 1; f->subdivision = 1; f->shading = 1; f->shadows = 0;
    f->shadow_samples = 100;                                         // ShadowMethod.cs:9
    f->point_lighting = 1; f->specular_lighting = 1;
    f->random_seed = 1234567890;                                     // Renderer.cs:85
    f->band_height = 0; f->band_count = 1; f->band_index = 0;
}
