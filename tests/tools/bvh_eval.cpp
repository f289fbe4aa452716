// Offline tree-quality evaluator (development / test tool, no GPU): builds the product's host binned-SAH tree (csrc/host/sah_builder.cpp),
// collapses it greedily to the 4-wide tree exactly as k_collapse_bvh4 does (lbvh.cu), and runs the ordered stack traversal of closest_hit
// (mcrt_device.cuh) over a dump of real queries, counting node visits and triangle tests per query.  It reproduces the GPU's own counters
// (mcrt_stats.bvh_node_visits with "count_traversal": 18.69 visits / 2.02 triangle tests per query on ircad11 at one frame per call), so
// builder ideas can be judged without GPU time.  ROT=n additionally applies n passes of Kensler tree rotations to the BVH2.
//   python tests/tools/bvh_eval_dump.py ircad            # scene triangles + the 21 570 queries of one frame (from the oracle) -> /tmp/ev/ircad.bin
//   g++ -O2 -std=c++17 -pthread -w -I mcray_tracing_b200/csrc/host tests/tools/bvh_eval.cpp mcray_tracing_b200/csrc/host/sah_builder.cpp -o /tmp/ev/bvh_eval
//   /tmp/ev/bvh_eval /tmp/ev/ircad.bin label
// Results of the round-2 study: profiles/r02ba_bvh_eval.txt.
#include "sah_builder.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace mcrt;
struct N4 { float lo[4][3], hi[4][3]; int child[4]; int n; };
static const int EMPTY = 0x7fffffff;
static void collapse(const std::vector<HostBvhNode>& n2, std::vector<N4>& n4)
{
    n4.resize(n2.size());
    for (size_t i = 0; i < n2.size(); i++) {
        N4& o = n4[i];
        auto put = [&](int at, const float* f, int ref) { for (int a = 0; a < 3; a++) { o.lo[at][a] = f[a]; o.hi[at][a] = f[3 + a]; } o.child[at] = ref; };
        put(0, n2[i].f, n2[i].child[0]); put(1, n2[i].f + 6, n2[i].child[1]);
        int n = 2;
        for (int round = 0; round < 2; round++) {
            int pick = -1; float best = -1.0f;
            for (int k = 0; k < n; k++) {
                if (o.child[k] < 0) continue;
                const float dx = o.hi[k][0] - o.lo[k][0], dy = o.hi[k][1] - o.lo[k][1], dz = o.hi[k][2] - o.lo[k][2];
                const float area = dx * dy + dy * dz + dz * dx;
                if (area > best) { best = area; pick = k; }
            }
            if (pick < 0) break;
            const HostBvhNode& c = n2[o.child[pick]];
            put(pick, c.f, c.child[0]); put(n, c.f + 6, c.child[1]); n++;
        }
        o.n = n;
        for (; n < 4; n++) o.child[n] = EMPTY;
    }
}

// ---- tree rotations (Kensler 2008) on the BVH2: swap a child with a grandchild of the other side when that shrinks the rotated node's box ----
static inline float area6(const float* b) { const float dx = b[3]-b[0], dy = b[4]-b[1], dz = b[5]-b[2]; return dx*dy+dy*dz+dz*dx; }
static inline void uni(const float* a, const float* b, float* o) { for (int k = 0; k < 3; k++) { o[k] = std::min(a[k], b[k]); o[3+k] = std::max(a[k+3], b[k+3]); } }
static long long rotate_pass(std::vector<HostBvhNode>& n)
{
    long long done = 0;
    // bottom-up: ids are pre-order, so descending id order visits children before parents (approximately, after rotations)
    for (int i = (int)n.size() - 1; i >= 0; i--) {
        HostBvhNode& p = n[i];
        float best_gain = 0.0f; int best_side = -1, best_g = -1;
        for (int side = 0; side < 2; side++) {              // rotate the inner child `side`'s children against the other child
            const int c = p.child[side];
            if (c < 0) continue;
            const float* other = p.f + 6 * (1 - side);       // box of the other child
            const float a_old = area6(p.f + 6 * side);
            for (int g = 0; g < 2; g++) {
                // swap grandchild g of c with `other`: c's new box = union(other, grandchild 1-g)
                float nb[6]; uni(other, n[c].f + 6 * (1 - g), nb);
                const float gain = a_old - area6(nb);
                if (gain > best_gain) { best_gain = gain; best_side = side; best_g = g; }
            }
        }
        if (best_side < 0) continue;
        const int side = best_side, g = best_g, c = p.child[side];
        HostBvhNode& q = n[c];
        // other child of p moves down into q's slot g; q's grandchild g moves up into p's slot (1 - side)
        float ob[6]; memcpy(ob, p.f + 6 * (1 - side), 24); const int oref = p.child[1 - side];
        memcpy(p.f + 6 * (1 - side), q.f + 6 * g, 24); p.child[1 - side] = q.child[g];
        memcpy(q.f + 6 * g, ob, 24); q.child[g] = oref;
        float nb[6]; uni(q.f, q.f + 6, nb); memcpy(p.f + 6 * side, nb, 24);
        done++;
    }
    return done;
}
static int depth_of(const std::vector<HostBvhNode>& n) { int deepest = 0; std::vector<std::pair<int,int>> st; st.emplace_back(0,1); while(!st.empty()){ auto [i,d]=st.back(); st.pop_back(); deepest=std::max(deepest,d); for(int k=0;k<2;k++) if(n[i].child[k]>=0) st.emplace_back(n[i].child[k], d+1);} return deepest; }
struct Tri { double v[3][3]; };
int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb");
    int64_t hd[3]; fread(hd, 8, 3, f);
    const int nt = (int)hd[0], nm = (int)hd[1], nr = (int)hd[2];
    std::vector<float> tri((size_t)nt * 9), org((size_t)nm * 3), from((size_t)nr * 3), to((size_t)nr * 3); std::vector<int32_t> mesh(nt);
    fread(tri.data(), 4, tri.size(), f); fread(mesh.data(), 4, nt, f); fread(org.data(), 4, org.size(), f); fread(from.data(), 4, from.size(), f); fread(to.data(), 4, to.size(), f);
    fclose(f);
    HostBvh hb;
    auto t0 = std::chrono::steady_clock::now();
    build_sah_bvh(tri.data(), mesh.data(), nt, org.data(), &hb, 0);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("ROT")) { const int passes = atoi(getenv("ROT")); for (int p = 0; p < passes; p++) { long long d = rotate_pass(hb.nodes); printf("  rotation pass %d: %lld rotations\n", p, d); if (!d) break; } hb.max_depth = depth_of(hb.nodes); }
    std::vector<N4> n4; collapse(hb.nodes, n4);
    // SAH cost of the BVH2 (sum of child areas / root area)
    double cost = 0; { for (auto& n : hb.nodes) for (int k = 0; k < 2; k++) { const float* b = n.f + 6 * k; double dx = b[3]-b[0], dy = b[4]-b[1], dz = b[5]-b[2]; cost += dx*dy+dy*dz+dz*dx; } }
    long long visits = 0, tests = 0, hits = 0, dead = 0, nhit_hist[5] = {0,0,0,0,0};
    std::vector<int> stack(256);
    for (int r = 0; r < nr; r++) {
        const double o[3] = {from[3*r], from[3*r+1], from[3*r+2]}, d[3] = {to[3*r]-o[0], to[3*r+1]-o[1], to[3*r+2]-o[2]};
        double inv[3]; for (int a = 0; a < 3; a++) inv[a] = 1.0 / (d[a] == 0 ? 1e-30 : d[a]);
        double best = 1.0; int sp = 0; int node = 0;
        while (true) {
            if (node >= 0) {
                visits++;
                const N4& n = n4[node];
                double t[4]; int c[4];
                for (int k = 0; k < 4; k++) {
                    c[k] = n.child[k]; t[k] = 3e38;
                    if (k >= n.n) continue;
                    double tn = 0, tf = best * 1.000002;
                    for (int a = 0; a < 3; a++) {
                        double t0 = (n.lo[k][a] - 1e-4 - o[a]) * inv[a], t1 = (n.hi[k][a] + 1e-4 - o[a]) * inv[a];
                        if (t0 > t1) std::swap(t0, t1);
                        tn = std::max(tn, t0); tf = std::min(tf, t1);
                    }
                    if (tn <= tf) t[k] = tn;
                }
                auto cs = [&](int i, int j) { if (t[j] < t[i]) { std::swap(t[i], t[j]); std::swap(c[i], c[j]); } };
                cs(0,1); cs(2,3); cs(0,2); cs(1,3); cs(1,2);
                { int nh = 0; for (int k = 0; k < 4; k++) nh += t[k] < 3e38; nhit_hist[nh]++; }
                if (t[0] < 3e38) {
                    if (t[3] < 3e38) stack[sp++] = c[3];
                    if (t[2] < 3e38) stack[sp++] = c[2];
                    if (t[1] < 3e38) stack[sp++] = c[1];
                    node = c[0];
                    continue;
                }
            } else {
                const int code = -node - 1; const int first = code >> 2, cnt = (code & 3) + 1;
                for (int slot = first; slot < first + cnt; slot++) {
                tests++;
                const HostTriSlot& s = hb.slots[slot];
                const float* og = &org[3 * s.mesh];
                double v0[3], e1[3], e2[3];
                for (int a = 0; a < 3; a++) { v0[a] = (double)s.v[a] + og[a]; e1[a] = (double)s.v[3+a] + og[a] - v0[a]; e2[a] = (double)s.v[6+a] + og[a] - v0[a]; }
                double p[3] = {d[1]*e2[2]-d[2]*e2[1], d[2]*e2[0]-d[0]*e2[2], d[0]*e2[1]-d[1]*e2[0]};
                double det = e1[0]*p[0]+e1[1]*p[1]+e1[2]*p[2];
                if (std::fabs(det) > 1e-300) {
                    double id = 1.0 / det, tv[3] = {o[0]-v0[0], o[1]-v0[1], o[2]-v0[2]};
                    double u = (tv[0]*p[0]+tv[1]*p[1]+tv[2]*p[2]) * id;
                    double q[3] = {tv[1]*e1[2]-tv[2]*e1[1], tv[2]*e1[0]-tv[0]*e1[2], tv[0]*e1[1]-tv[1]*e1[0]};
                    double v = (d[0]*q[0]+d[1]*q[1]+d[2]*q[2]) * id;
                    double tt = (e2[0]*q[0]+e2[1]*q[1]+e2[2]*q[2]) * id;
                    if (u >= -1e-4 && v >= -1e-4 && u + v <= 1.0001 && tt > 0 && tt < best) { best = tt; }
                }
                }
            }
            if (sp == 0) break;
            node = stack[--sp];
        }
        if (best < 1.0) hits++;
    }
    { // reachable BVH4 nodes: children histogram, leaf depth
        long long ch_hist[5] = {0,0,0,0,0}, leaves = 0, depth_sum = 0, reach = 0;
        std::vector<std::pair<int,int>> st; st.emplace_back(0, 1);
        while (!st.empty()) { auto [nd, dp] = st.back(); st.pop_back(); reach++; ch_hist[n4[nd].n]++;
            for (int k = 0; k < n4[nd].n; k++) { int c = n4[nd].child[k]; if (c >= 0) st.emplace_back(c, dp + 1); else { leaves++; depth_sum += dp; } } }
        printf("  reachable BVH4 nodes %lld of %zu; children 2/3/4: %lld %lld %lld; mean leaf depth %.2f | visits by #hit children 0..4: %.3f %.3f %.3f %.3f %.3f per query\n",
               reach, n4.size(), ch_hist[2], ch_hist[3], ch_hist[4], (double)depth_sum / leaves, (double)nhit_hist[0]/nr, (double)nhit_hist[1]/nr, (double)nhit_hist[2]/nr, (double)nhit_hist[3]/nr, (double)nhit_hist[4]/nr);
    }
    printf("%s: build %.0f ms, depth %d, sah-area %.4g | node visits/query %.3f tri tests/query %.3f hits %lld/%d\n", argc > 2 ? argv[2] : "", ms, hb.max_depth, cost,
           (double)visits / nr, (double)tests / nr, hits, nr);
}
