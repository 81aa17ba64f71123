// sr_bvh.cpp -- deterministic binned-SAH BVH2 builder (host).  See sr_bvh.h.
#include "sr_bvh.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <thread>
#include <cmath>
#include <cstring>

namespace sr {

float round_down(double x)
{
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}

float round_up(double x)
{
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}

namespace {

struct Box {
    float lo[3], hi[3];
    void reset()
    {
        for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
    }
    void grow(const float* l, const float* h)
    {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); }
    }
    void grow_point(const float* p)
    {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
    }
    double half_area() const
    {
        double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0;
        return dx * dy + dy * dz + dz * dx;
    }
};

constexpr int kBins = 16;
constexpr int32_t kParallelMinPrims = 200000;   // below this one thread builds the whole tree

struct Builder {
    const std::vector<PrimBounds>& prims;
    std::vector<float> cent;   // 3 per prim
    std::vector<int32_t> idx;
    BvhBuild* out;
    float pad;
    int max_leaf;
    double isect_cost;   // SAH cost of one primitive test in units of one node visit

    Builder(const std::vector<PrimBounds>& p, float pad_, int max_leaf_, double isect_cost_, BvhBuild* o)
        : prims(p), out(o), pad(pad_), max_leaf(max_leaf_), isect_cost(isect_cost_)
    {
    }

    struct Child {
        int32_t ref;     // node index or first primitive
        int32_t count;   // 0 internal, >0 leaf, -1 none
        Box box;
    };

    // traversal link (sr_types.h BvhNode): internal child = node index, leaf = -1 - (first * 16 + count),
    // no child = kNoChild (an empty leaf behind an unreachable box)
    static int32_t encode(const Child& c)
    {
        if (c.count < 0) return kNoChild;
        return c.count == 0 ? c.ref : -1 - (c.ref * 16 + c.count);
    }

    void write_box(float* lo3, float* hi3, const Box& b) const
    {
        for (int k = 0; k < 3; k++) {
            lo3[k] = round_down((double)b.lo[k] - (double)pad);
            hi3[k] = round_up((double)b.hi[k] + (double)pad);
        }
    }

    Box bounds_of(int32_t begin, int32_t end) const
    {
        Box b; b.reset();
        for (int32_t i = begin; i < end; i++) b.grow(prims[(size_t)idx[(size_t)i]].lo, prims[(size_t)idx[(size_t)i]].hi);
        return b;
    }

    // Returns the split position (begin < mid < end) or -1 to make a leaf.
    int32_t find_split(int32_t begin, int32_t end, const Box& box, int depth, std::vector<int32_t>& scratch)
    {
        if (scratch.size() < (size_t)(end - begin)) scratch.resize((size_t)(end - begin));
        const int32_t n = end - begin;
        Box cb; cb.reset();
        for (int32_t i = begin; i < end; i++) cb.grow_point(&cent[3 * (size_t)idx[(size_t)i]]);

        double best_cost = DBL_MAX; int best_axis = -1, best_bin = -1;
        const double parent_area = box.half_area();
        const bool sah_ok = depth < kMaxBvhDepth - 24;   // deep trees: balanced median splits only
        if (sah_ok && parent_area > 0.0) {
            for (int axis = 0; axis < 3; axis++) {
                const float c0 = cb.lo[axis], c1 = cb.hi[axis];
                if (!(c1 > c0)) continue;
                const double scale = (double)kBins / ((double)c1 - (double)c0);
                Box bb[kBins]; int32_t cnt[kBins];
                for (int b = 0; b < kBins; b++) { bb[b].reset(); cnt[b] = 0; }
                for (int32_t i = begin; i < end; i++) {
                    const int32_t p = idx[(size_t)i];
                    int b = (int)(((double)cent[3 * (size_t)p + axis] - (double)c0) * scale);
                    if (b < 0) b = 0;
                    if (b >= kBins) b = kBins - 1;
                    bb[b].grow(prims[(size_t)p].lo, prims[(size_t)p].hi);
                    cnt[b]++;
                }
                double right_area[kBins]; int32_t right_cnt[kBins];
                Box acc; acc.reset(); int32_t c = 0;
                for (int b = kBins - 1; b > 0; b--) {
                    if (cnt[b]) acc.grow(bb[b].lo, bb[b].hi);
                    c += cnt[b];
                    right_area[b] = acc.half_area(); right_cnt[b] = c;
                }
                acc.reset(); c = 0;
                for (int b = 0; b < kBins - 1; b++) {
                    if (cnt[b]) acc.grow(bb[b].lo, bb[b].hi);
                    c += cnt[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    const double cost = acc.half_area() * c + right_area[b + 1] * right_cnt[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
                }
            }
        }
        if (best_axis >= 0) {
            // leaf cost n * C_isect vs split cost C_trav + C_isect * cost/area (C_trav = 1)
            const double split_cost = 1.0 + isect_cost * best_cost / parent_area;
            if (n <= max_leaf && isect_cost * n <= split_cost) return -1;
            const float c0 = cb.lo[best_axis], c1 = cb.hi[best_axis];
            const double scale = (double)kBins / ((double)c1 - (double)c0);
            // stable partition keeps the input order inside each side (deterministic layout)
            int32_t nl = 0, nr = 0;
            for (int32_t i = begin; i < end; i++) {
                const int32_t p = idx[(size_t)i];
                int b = (int)(((double)cent[3 * (size_t)p + best_axis] - (double)c0) * scale);
                if (b < 0) b = 0;
                if (b >= kBins) b = kBins - 1;
                if (b <= best_bin) idx[(size_t)(begin + nl++)] = p; else scratch[(size_t)nr++] = p;
            }
            std::memcpy(&idx[(size_t)(begin + nl)], scratch.data(), sizeof(int32_t) * (size_t)nr);
            if (nl > 0 && nr > 0) return begin + nl;
        }
        if (n <= max_leaf) return -1;
        // median split on the widest centroid axis (falls back to index order when all equal)
        int axis = 0;
        float ext = cb.hi[0] - cb.lo[0];
        for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > ext) { ext = cb.hi[k] - cb.lo[k]; axis = k; }
        std::stable_sort(idx.begin() + begin, idx.begin() + end, [&](int32_t a, int32_t b) {
            return cent[3 * (size_t)a + axis] < cent[3 * (size_t)b + axis];
        });
        return begin + n / 2;
    }

    // One subtree, depth first, into `nodes` (indices local to that vector).  Used directly for small
    // inputs and per task for large ones.
    struct Local {
        std::vector<BvhNode> nodes;
        std::vector<int32_t> scratch;
        int32_t depth = 0, n_leaves = 0;
    };

    void set_children_in(std::vector<BvhNode>& nodes, int32_t node, const Child& a, const Child& b) const
    {
        BvhNode& n = nodes[(size_t)node];
        std::memset(&n, 0, sizeof n);
        float lo[3], hi[3];
        if (a.count >= 0) write_box(lo, hi, a.box);
        else { lo[0] = lo[1] = lo[2] = FLT_MAX; hi[0] = hi[1] = hi[2] = FLT_MAX; }
        n.lo0x = lo[0]; n.lo0y = lo[1]; n.lo0z = lo[2]; n.hi0x = hi[0]; n.hi0y = hi[1]; n.hi0z = hi[2];
        if (b.count >= 0) write_box(lo, hi, b.box);
        else { lo[0] = lo[1] = lo[2] = FLT_MAX; hi[0] = hi[1] = hi[2] = FLT_MAX; }
        n.lo1x = lo[0]; n.lo1y = lo[1]; n.lo1z = lo[2]; n.hi1x = hi[0]; n.hi1y = hi[1]; n.hi1z = hi[2];
        n.child0 = encode(a); n.count0 = a.count;
        n.child1 = encode(b); n.count1 = b.count;
    }

    Child build_range(Local& L, int32_t begin, int32_t end, int depth)
    {
        Child c;
        c.box = bounds_of(begin, end);
        if (depth > L.depth) L.depth = depth;
        const int32_t mid = find_split(begin, end, c.box, depth, L.scratch);
        if (mid < 0) {
            c.ref = begin; c.count = end - begin;
            L.n_leaves++;
            return c;
        }
        const int32_t node = (int32_t)L.nodes.size();
        L.nodes.emplace_back();
        std::memset(&L.nodes.back(), 0, sizeof(BvhNode));
        Child a = build_range(L, begin, mid, depth + 1);
        Child b = build_range(L, mid, end, depth + 1);
        set_children_in(L.nodes, node, a, b);
        c.ref = node; c.count = 0;
        return c;
    }

    // Large inputs: the top of the tree is split on this thread until there are enough independent
    // ranges, every range is then built by a worker into its own node vector, and the vectors are
    // appended in range order.  Which thread builds what never shows in the result: the layout depends
    // on the input only ("bit-identical layout across runs").
    struct Pending { Child c[2]; };
    struct Task { int32_t begin, end, depth, node, slot; Child result; Local local; };

    void run()
    {
        const int32_t n = (int32_t)prims.size();
        out->nodes.clear(); out->order.clear(); out->depth = 0; out->n_leaves = 0;
        idx.resize((size_t)n); cent.resize(3 * (size_t)n);
        for (int32_t i = 0; i < n; i++) {
            idx[(size_t)i] = i;
            for (int k = 0; k < 3; k++) cent[3 * (size_t)i + k] = 0.5f * prims[(size_t)i].lo[k] + 0.5f * prims[(size_t)i].hi[k];
        }
        Child none; none.ref = -1; none.count = -1; none.box.reset();
        // the root is always an internal node: slot 0
        out->nodes.emplace_back();
        std::memset(&out->nodes.back(), 0, sizeof(BvhNode));
        if (n == 0) {
            set_children_in(out->nodes, 0, none, none);
            return;
        }
        Box root = bounds_of(0, n);
        for (int k = 0; k < 3; k++) { out->root_lo[k] = root.lo[k]; out->root_hi[k] = root.hi[k]; }
        out->depth = 1;

        std::vector<Pending> pending(1);
        std::vector<Task> open;             // ranges not split yet, each hangs off (node, slot)
        std::vector<int32_t> scratch;
        {
            const int32_t mid = find_split(0, n, root, 1, scratch);
            if (mid < 0) {
                Child leaf; leaf.ref = 0; leaf.count = n; leaf.box = root;
                out->n_leaves = 1;
                set_children_in(out->nodes, 0, leaf, none);
                out->order = idx;
                return;
            }
            Task a; a.begin = 0; a.end = mid; a.depth = 2; a.node = 0; a.slot = 0;
            Task b; b.begin = mid; b.end = n; b.depth = 2; b.node = 0; b.slot = 1;
            open.push_back(std::move(a)); open.push_back(std::move(b));
        }
        const unsigned hw = std::thread::hardware_concurrency();
        const int n_threads = n >= kParallelMinPrims ? (int)std::min<unsigned>(hw ? hw : 1u, 32u) : 1;
        const size_t want_tasks = n_threads > 1 ? (size_t)n_threads * 8 : 0;
        const int32_t min_task = 4096;
        // split the largest open range until there are enough of them (ties: the lower range first)
        while (open.size() < want_tasks) {
            size_t pick = open.size();
            for (size_t i = 0; i < open.size(); i++)
                if (open[i].end - open[i].begin > min_task &&
                    (pick == open.size() || open[i].end - open[i].begin > open[pick].end - open[pick].begin))
                    pick = i;
            if (pick == open.size()) break;
            Task t = std::move(open[pick]);
            open.erase(open.begin() + (long)pick);
            Child c; c.box = bounds_of(t.begin, t.end);
            if (t.depth > out->depth) out->depth = t.depth;
            const int32_t mid = find_split(t.begin, t.end, c.box, t.depth, scratch);
            if (mid < 0) {
                c.ref = t.begin; c.count = t.end - t.begin;
                out->n_leaves++;
            } else {
                const int32_t node = (int32_t)out->nodes.size();
                out->nodes.emplace_back();
                std::memset(&out->nodes.back(), 0, sizeof(BvhNode));
                pending.emplace_back();
                c.ref = node; c.count = 0;
                Task a; a.begin = t.begin; a.end = mid; a.depth = t.depth + 1; a.node = node; a.slot = 0;
                Task b; b.begin = mid; b.end = t.end; b.depth = t.depth + 1; b.node = node; b.slot = 1;
                open.push_back(std::move(a)); open.push_back(std::move(b));
            }
            pending[(size_t)t.node].c[t.slot] = c;
        }
        std::sort(open.begin(), open.end(), [](const Task& x, const Task& y) { return x.begin < y.begin; });

        // build every open range (in parallel when it pays)
        std::atomic<size_t> next(0);
        auto worker = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= open.size()) break;
                Task& t = open[i];
                t.result = build_range(t.local, t.begin, t.end, t.depth);
            }
        };
        if (n_threads > 1 && open.size() > 1) {
            std::vector<std::thread> pool;
            for (int k = 0; k < n_threads; k++) pool.emplace_back(worker);
            for (auto& th : pool) th.join();
        } else {
            worker();
        }
        // append the subtrees in range order and re-base their links
        for (Task& t : open) {
            const int32_t base = (int32_t)out->nodes.size();
            for (BvhNode nd : t.local.nodes) {
                if (nd.count0 == 0) nd.child0 += base;
                if (nd.count1 == 0) nd.child1 += base;
                out->nodes.push_back(nd);
            }
            Child c = t.result;
            if (c.count == 0) c.ref += base;
            pending[(size_t)t.node].c[t.slot] = c;
            if (t.local.depth > out->depth) out->depth = t.local.depth;
            out->n_leaves += t.local.n_leaves;
            std::vector<BvhNode>().swap(t.local.nodes);
        }
        for (size_t node = 0; node < pending.size(); node++)
            set_children_in(out->nodes, (int32_t)node, pending[node].c[0], pending[node].c[1]);
        out->order = idx;
    }
};

}  // namespace

void build_bvh(const std::vector<PrimBounds>& prims, float pad, int max_leaf, double isect_cost, BvhBuild* out)
{
    if (max_leaf > kMaxLeafPrims) max_leaf = kMaxLeafPrims;
    if (max_leaf < 1) max_leaf = 1;
    Builder b(prims, pad, max_leaf, isect_cost, out);
    b.run();
    if (out->order.empty() && !prims.empty()) out->order = b.idx;
}

}  // namespace sr
