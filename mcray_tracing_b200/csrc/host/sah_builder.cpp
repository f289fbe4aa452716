// sah_builder.cpp -- host-side BVH builder (binned surface-area heuristic) producing the same 64-byte node /
// 48-byte triangle-slot layout as the device LBVH (kernels/lbvh.cu).  The device LBVH serves a new or changed scene at
// once (it builds in < 1 ms); this tree is visited with ~8 % fewer node fetches per ray, so a context builds it on a
// background thread and swaps it in when it is ready (mcrt_abi.cu, option "bvh_optimise"), or builds it in place with
// mcrt_set_option(ctx, "bvh_builder", 1).  Node ids are the pre-order positions, known before a subtree is built
// (a range of n triangles holds n - 1 nodes), so the two children of a large node are built by different threads
// and the result does not depend on the number of threads.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

#include "sah_builder.h"

namespace mcrt {

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = 3.0e38f; hi[a] = -3.0e38f; } }
    void grow(const Box& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Builder {
    const std::vector<Box>& tri_box;
    std::vector<float> cent;          // 3 per triangle
    std::vector<int32_t> order;       // triangle ids, permuted in place (disjoint ranges per subtree)
    std::vector<HostBvhNode> nodes;   // n_tri - 1, pre-order; every subtree writes its own id range
    std::atomic<int> max_depth{0};
    std::atomic<int> spare_threads{0};
    std::atomic<bool> failed{false};
    const std::atomic<bool>* cancel = nullptr;     // set by the owner to abandon the build

    explicit Builder(const std::vector<Box>& tb) : tri_box(tb) {}

    void note_depth(int d)
    {
        int cur = max_depth.load(std::memory_order_relaxed);
        while (d > cur && !max_depth.compare_exchange_weak(cur, d, std::memory_order_relaxed)) {}
    }

    // up to `want` spare threads for the caller; give them back with release()
    int claim(int want)
    {
        int got = 0;
        while (got < want) {
            if (spare_threads.fetch_sub(1, std::memory_order_relaxed) > 0) got++;
            else { spare_threads.fetch_add(1, std::memory_order_relaxed); break; }
        }
        return got;
    }
    void release(int n) { if (n > 0) spare_threads.fetch_add(n, std::memory_order_relaxed); }

    // fn(k, lo, hi) over extra + 1 contiguous chunks of [first, first + count): chunk 0 on the calling thread, the others on their own
    template <class F>
    void chunked(int first, int count, int extra, F&& fn)
    {
        const int parts = extra + 1;
        auto cut = [&](int k) { return first + (int)((long long)count * k / parts); };
        std::vector<std::thread> th;
        for (int k = 1; k < parts; k++)
            th.emplace_back([&, k]() { try { fn(k, cut(k), cut(k + 1)); } catch (...) { failed.store(true); } });
        try { fn(0, cut(0), cut(1)); } catch (...) { failed.store(true); }
        for (auto& t : th) t.join();
    }

    static constexpr int NB = 16;              // bins per axis
    static constexpr int kChunkMin = 65536;    // nodes from this size up bin their triangles in parallel chunks ...
    static constexpr int kChunkMax = 7;        // ... on at most this many extra threads
    struct Bins { Box box[3][NB]; int cnt[3][NB]; };

    // builds the subtree of order[first, first + count) into nodes[id ...]; returns the child reference of its root:
    // >= 0 internal node index, < 0 leaf -(1 + slot*4)
    int build(int first, int count, int depth, int id, Box* out_box)
    {
        // Large nodes (the top levels, where a single thread would otherwise walk every triangle of the scene) take their two passes over the
        // triangles in parallel chunks.  Chunk results are merged with min / max / integer adds only, so the bins -- and the tree -- are the
        // same for any number of chunks.
        const int extra = count >= kChunkMin ? claim(kChunkMax) : 0;
        Box bb; bb.reset();
        Box cb; cb.reset();
        {
            Box pbb[kChunkMax + 1], pcb[kChunkMax + 1];
            chunked(first, count, extra, [&](int k, int lo, int hi) {
                Box b1; b1.reset();
                Box c1; c1.reset();
                for (int i = lo; i < hi; i++) {
                    const int t = order[i];
                    b1.grow(tri_box[t]);
                    for (int a = 0; a < 3; a++) { c1.lo[a] = std::min(c1.lo[a], cent[3 * (size_t)t + a]); c1.hi[a] = std::max(c1.hi[a], cent[3 * (size_t)t + a]); }
                }
                pbb[k] = b1; pcb[k] = c1;
            });
            for (int k = 0; k <= extra; k++) { bb.grow(pbb[k]); cb.grow(pcb[k]); }
        }
        *out_box = bb;
        if (count == 1) { release(extra); return -(1 + first * 4); }
        if (count >= 256 && cancel && cancel->load(std::memory_order_relaxed)) { release(extra); throw std::runtime_error("sah builder: cancelled"); }
        note_depth(depth + 1);
        // binned SAH over the centroid bounds
        float scale[3];
        bool live[3];
        for (int a = 0; a < 3; a++) { const float ext = cb.hi[a] - cb.lo[a]; live[a] = ext > 0.0f; scale[a] = live[a] ? NB / ext : 0.0f; }
        Bins own;                                        // chunk 0 (the only one of a small node: no allocation there)
        std::vector<Bins> others((size_t)extra);
        chunked(first, count, extra, [&](int k, int lo, int hi) {
            Bins& bn = k == 0 ? own : others[(size_t)k - 1];
            for (int a = 0; a < 3; a++) for (int b = 0; b < NB; b++) { bn.box[a][b].reset(); bn.cnt[a][b] = 0; }
            for (int i = lo; i < hi; i++) {
                const int t = order[i];
                for (int a = 0; a < 3; a++) {
                    if (!live[a]) continue;
                    int b = (int)((cent[3 * (size_t)t + a] - cb.lo[a]) * scale[a]);
                    b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    bn.box[a][b].grow(tri_box[t]); bn.cnt[a][b]++;
                }
            }
        });
        release(extra);
        for (int k = 0; k < extra; k++)
            for (int a = 0; a < 3; a++) for (int b = 0; b < NB; b++) { own.box[a][b].grow(others[k].box[a][b]); own.cnt[a][b] += others[k].cnt[a][b]; }
        int best_axis = -1, best_bin = -1;
        float best_cost = 3.0e38f;
        for (int a = 0; a < 3; a++) {
            if (!live[a]) continue;
            const Box* bins = own.box[a];
            const int* cnt = own.cnt[a];
            float right_area[NB]; int right_cnt[NB];
            Box acc; acc.reset(); int c = 0;
            for (int b = NB - 1; b >= 1; b--) { acc.grow(bins[b]); c += cnt[b]; right_area[b] = acc.area(); right_cnt[b] = c; }
            acc.reset(); c = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bins[b]); c += cnt[b];
                if (c == 0 || right_cnt[b + 1] == 0) continue;
                const float cost = acc.area() * c + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
            }
        }
        int mid;
        if (best_axis >= 0 && depth < 40) {      // beyond depth 40 fall back to balanced splits: bounds the stack
            const float bscale = scale[best_axis];
            const float lo = cb.lo[best_axis];
            // stable: the triangle order inside a leaf range (and with it the tree) does not depend on the library's partition algorithm
            auto it = std::stable_partition(order.begin() + first, order.begin() + first + count, [&](int32_t t) {
                int b = (int)((cent[3 * (size_t)t + best_axis] - lo) * bscale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                return b <= best_bin;
            });
            mid = (int)(it - order.begin());
        } else {
            mid = first + count / 2;       // all centroids coincide: split by index
        }
        if (mid == first || mid == first + count) mid = first + count / 2;
        const int left = mid - first, right = first + count - mid;
        // pre-order ids: the left subtree (left - 1 nodes) follows this node, the right subtree follows the left one
        const int id_left = id + 1, id_right = id + left;
        Box lb, rb;
        int lc = 0, rc = 0;
        bool fork = false;
        if (left >= kForkMin && right >= kForkMin) {
            if (spare_threads.fetch_sub(1, std::memory_order_relaxed) > 0) fork = true;
            else spare_threads.fetch_add(1, std::memory_order_relaxed);
        }
        if (fork) {
            std::thread th([&]() {
                try { lc = build(first, left, depth + 1, id_left, &lb); } catch (...) { failed.store(true); }
            });
            try { rc = build(mid, right, depth + 1, id_right, &rb); } catch (...) { failed.store(true); }
            th.join();
            spare_threads.fetch_add(1, std::memory_order_relaxed);
        } else {
            lc = build(first, left, depth + 1, id_left, &lb);
            rc = build(mid, right, depth + 1, id_right, &rb);
        }
        HostBvhNode& nd = nodes[id];
        nd.f[0] = lb.lo[0]; nd.f[1] = lb.lo[1]; nd.f[2] = lb.lo[2]; nd.f[3] = lb.hi[0]; nd.f[4] = lb.hi[1]; nd.f[5] = lb.hi[2];
        nd.f[6] = rb.lo[0]; nd.f[7] = rb.lo[1]; nd.f[8] = rb.lo[2]; nd.f[9] = rb.hi[0]; nd.f[10] = rb.hi[1]; nd.f[11] = rb.hi[2];
        nd.child[0] = lc; nd.child[1] = rc; nd.child[2] = nd.child[3] = 0;
        return id;
    }
    static constexpr int kForkMin = 8192;     // both children at least this large before a thread is worth starting
};

}  // namespace

void build_sah_bvh(const float* tri_local, const int32_t* tri_mesh, int n_tri, const float* mesh_origin3, HostBvh* out, int max_threads,
                   const std::atomic<bool>* cancel)
{
    out->nodes.clear(); out->slots.clear(); out->max_depth = 0; out->max_abs = 0.0f;
    if (n_tri <= 0) return;
    std::vector<Box> boxes((size_t)n_tri);
    Builder b(boxes);
    b.cent.resize((size_t)n_tri * 3);
    b.order.resize((size_t)n_tri);
    for (int t = 0; t < n_tri; t++) {
        const float* o = mesh_origin3 + 3 * (size_t)tri_mesh[t];
        Box bx; bx.reset();
        for (int k = 0; k < 3; k++)
            for (int a = 0; a < 3; a++) {
                const float w = tri_local[9 * (size_t)t + 3 * k + a] + o[a];       // same fp32 add as the device build
                bx.lo[a] = std::min(bx.lo[a], w); bx.hi[a] = std::max(bx.hi[a], w);
            }
        boxes[t] = bx;
        for (int a = 0; a < 3; a++) {
            b.cent[3 * (size_t)t + a] = 0.5f * (bx.lo[a] + bx.hi[a]);
            out->max_abs = std::max(out->max_abs, std::max(std::fabs(bx.lo[a]), std::fabs(bx.hi[a])));
        }
        b.order[t] = t;
    }
    if (n_tri >= 2) {
        b.nodes.resize((size_t)n_tri - 1);
        int threads = max_threads > 0 ? max_threads : (int)std::thread::hardware_concurrency();
        threads = threads < 1 ? 1 : (threads > 32 ? 32 : threads);
        b.spare_threads.store(threads - 1);
        b.cancel = cancel;
        Box root;
        const int r = b.build(0, n_tri, 0, 0, &root);
        if (b.failed.load()) throw std::runtime_error("sah builder: a worker thread failed or the build was cancelled");
        if (r != 0) throw std::runtime_error("sah builder: root is not node 0");
    }
    out->nodes.swap(b.nodes);
    out->max_depth = b.max_depth.load();
    out->slots.resize((size_t)n_tri);
    for (int k = 0; k < n_tri; k++) {
        const int t = b.order[k];
        HostTriSlot& s = out->slots[k];
        memcpy(s.v, tri_local + 9 * (size_t)t, sizeof(float) * 9);
        s.mesh = tri_mesh[t]; s.tri = t;
    }
}

}  // namespace mcrt
