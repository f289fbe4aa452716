// lbvh.cu -- device-side BVH construction: replaces Bullet's btBvhTriangleMeshShape /
// btQuantizedBvh build per OBJ (scene.cpp:300-334, `new btBvhTriangleMeshShape(tiva, true)`) and
// the btDbvtBroadphase over the bodies (scene.cpp:249-262) with ONE world-space LBVH over all
// triangles of all meshes.
//
// Pipeline (all on the GPU): world boxes + centroids -> 63-bit Morton keys -> radix sort
// (cub::DeviceRadixSort, the CUDA toolkit's primitive; start-up only) -> Karras 2012 radix-tree
// hierarchy -> bottom-up refit with arrival counters -> 64-byte traversal nodes + Morton-ordered
// 48-byte triangle slots.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <utility>
#include <vector>

#include "mcrt_device.cuh"
#include "mcrt_launch.h"

namespace mcrt {

namespace {

__device__ __forceinline__ int f2ord(float f)       // order-preserving float -> int
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ord2f(int i)
{
    const int j = i >= 0 ? i : i ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
    return __int_as_float(j);
#else
    float f; memcpy(&f, &j, 4); return f;
#endif
}

// world-space box of a triangle: fl(v_local + body origin) per vertex
__global__ void k_tri_bounds(const float* __restrict__ tri_local, const int32_t* __restrict__ tri_mesh, const DevMesh* __restrict__ meshes,
                             int n, float4* __restrict__ box_lo, float4* __restrict__ box_hi, int* __restrict__ scene_bounds)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const DevMesh m = meshes[tri_mesh[t]];
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = 0; k < 3; k++) {
        const float w[3] = {tri_local[9 * (size_t)t + 3 * k] + m.ox, tri_local[9 * (size_t)t + 3 * k + 1] + m.oy,
                            tri_local[9 * (size_t)t + 3 * k + 2] + m.oz};
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], w[a]); hi[a] = fmaxf(hi[a], w[a]); }
    }
    box_lo[t] = make_float4(lo[0], lo[1], lo[2], 0.f);
    box_hi[t] = make_float4(hi[0], hi[1], hi[2], 0.f);
    for (int a = 0; a < 3; a++) {
        atomicMin(&scene_bounds[a], f2ord(lo[a]));
        atomicMax(&scene_bounds[3 + a], f2ord(hi[a]));
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v)
{
    v &= 0x1fffffULL;
    v = (v | (v << 32)) & 0x1f00000000ffffULL;
    v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
    v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
    v = (v | (v << 2)) & 0x1249249249249249ULL;
    return v;
}

__global__ void k_morton(const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, const int* __restrict__ scene_bounds, int n,
                         unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float q[3];
    const float lo[3] = {box_lo[t].x, box_lo[t].y, box_lo[t].z}, hi[3] = {box_hi[t].x, box_hi[t].y, box_hi[t].z};
    for (int a = 0; a < 3; a++) {
        const float slo = ord2f(scene_bounds[a]), shi = ord2f(scene_bounds[3 + a]);
        const float c = 0.5f * (lo[a] + hi[a]);
        const float ext = fmaxf(shi - slo, 1e-30f);
        q[a] = fminf(fmaxf((c - slo) / ext, 0.0f), 1.0f);
    }
    const unsigned long long x = (unsigned long long)fminf(q[0] * 2097152.0f, 2097151.0f);
    const unsigned long long y = (unsigned long long)fminf(q[1] * 2097152.0f, 2097151.0f);
    const unsigned long long z = (unsigned long long)fminf(q[2] * 2097152.0f, 2097151.0f);
    keys[t] = (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
    vals[t] = (unsigned int)t;
}

// Karras 2012: delta(i,j) = length of the common prefix of keys i and j (key ties broken by index)
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_karras(const unsigned long long* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                         int* __restrict__ parent_internal, int* __restrict__ parent_leaf, int2* __restrict__ range)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    // child encoding: >= 0 internal, < 0 leaf (~sorted position)
    const int lc = (lo == gamma) ? ~gamma : gamma;
    const int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    range[i] = make_int2(lo, hi);      // sorted-position range covered by this node
    if (lc >= 0) parent_internal[lc] = i; else parent_leaf[~lc] = i;
    if (rc >= 0) parent_internal[rc] = i; else parent_leaf[~rc] = i;
    if (i == 0) parent_internal[0] = -1;
}

// bottom-up refit: the second thread to arrive at a node owns it
__global__ void k_refit(const unsigned int* __restrict__ vals, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, int n,
                        const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent_internal,
                        const int* __restrict__ parent_leaf, int* __restrict__ arrivals, float4* __restrict__ node_lo,
                        float4* __restrict__ node_hi, int* __restrict__ max_depth)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int node = parent_leaf[k];
    int depth = 1;
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&arrivals[node], 1) == 0) return;      // first arrival: sibling subtree not ready
        __threadfence();
        const int lc = left[node], rc = right[node];
        float4 llo, lhi, rlo, rhi;
        if (lc >= 0) { llo = __ldcg(&node_lo[lc]); lhi = __ldcg(&node_hi[lc]); } else { const unsigned int t = vals[~lc]; llo = box_lo[t]; lhi = box_hi[t]; }
        if (rc >= 0) { rlo = __ldcg(&node_lo[rc]); rhi = __ldcg(&node_hi[rc]); } else { const unsigned int t = vals[~rc]; rlo = box_lo[t]; rhi = box_hi[t]; }
        node_lo[node] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
        node_hi[node] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
        depth++;
        node = parent_internal[node];
    }
    atomicMax(max_depth, depth);   // only the thread that completes the root gets here
}

// exact depth of the tree (longest leaf-to-root chain): the traversal stack bound
__global__ void k_leaf_depth(int n, const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf, int* __restrict__ max_depth)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int depth = 0;
    for (int node = parent_leaf[k]; node >= 0; node = parent_internal[node]) depth++;
    atomicMax(max_depth, depth);
}

// Child reference of the traversal layout: >= 0 internal node; < 0 a leaf holding `count` (1..MCRT_LEAF_MAX)
// consecutive Morton-ordered triangle slots starting at `first`: -(1 + first * 4 + (count - 1)).  Radix-tree
// subtrees of at most MCRT_LEAF_MAX triangles are collapsed into one leaf (their inner nodes are never visited).
__device__ __forceinline__ int encode_child(int child, const int2* __restrict__ range)
{
    if (child < 0) return -(1 + (~child) * 4);
    const int2 r = range[child];
    const int count = r.y - r.x + 1;
    if (count <= MCRT_LEAF_MAX) return -(1 + r.x * 4 + (count - 1));
    return child;
}

__global__ void k_emit_nodes(const unsigned int* __restrict__ vals, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                             int n, const int* __restrict__ left, const int* __restrict__ right, const float4* __restrict__ node_lo,
                             const float4* __restrict__ node_hi, const int2* __restrict__ range, BvhNode* __restrict__ nodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int lc = left[i], rc = right[i];
    float4 llo, lhi, rlo, rhi;
    if (lc >= 0) { llo = node_lo[lc]; lhi = node_hi[lc]; } else { const unsigned int t = vals[~lc]; llo = box_lo[t]; lhi = box_hi[t]; }
    if (rc >= 0) { rlo = node_lo[rc]; rhi = node_hi[rc]; } else { const unsigned int t = vals[~rc]; rlo = box_lo[t]; rhi = box_hi[t]; }
    BvhNode nd;
    nd.a = make_float4(llo.x, llo.y, llo.z, lhi.x);
    nd.b = make_float4(lhi.y, lhi.z, rlo.x, rlo.y);
    nd.c = make_float4(rlo.z, rhi.x, rhi.y, rhi.z);
    nd.d = make_int4(encode_child(lc, range), encode_child(rc, range), 0, 0);
    nodes[i] = nd;
}

__global__ void k_emit_tris(const unsigned int* __restrict__ vals, const float* __restrict__ tri_local, const int32_t* __restrict__ tri_mesh,
                            int n, TriSlot* __restrict__ slots)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned int t = vals[k];
    const float* p = tri_local + 9 * (size_t)t;
    TriSlot s;
    s.v0 = make_float4(p[0], p[1], p[2], __int_as_float(tri_mesh[t]));
    s.v1 = make_float4(p[3], p[4], p[5], __int_as_float((int)t));
    s.v2 = make_float4(p[6], p[7], p[8], 0.f);
    slots[k] = s;
}

// BVH2 -> BVH4.  Node i starts from its two BVH2 children and, twice, replaces the inner child with the LARGEST surface area by
// that child's two children (greedy surface-area collapse: the big boxes, which rays hit most often, are the ones opened up;
// MCRT_BVH4_GREEDY=0 gives the fixed shape "each child replaced by its children").  Built for EVERY inner BVH2 node -- any
// inner descendant can then be referenced as a child -- so no depth information is needed and the same kernel serves the
// device LBVH and the host SAH tree.
#ifndef MCRT_BVH4_GREEDY
#define MCRT_BVH4_GREEDY 1
#endif
__global__ void k_collapse_bvh4(const BvhNode* __restrict__ nodes2, const int n_nodes, Bvh4Node* __restrict__ nodes4, const int greedy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    float lo[4][3], hi[4][3];
    int child[4];
    int n = 0;
    auto put = [&](int at, float lx, float ly, float lz, float hx, float hy, float hz, int ref) {
        lo[at][0] = lx; lo[at][1] = ly; lo[at][2] = lz; hi[at][0] = hx; hi[at][1] = hy; hi[at][2] = hz; child[at] = ref;
    };
    const BvhNode nd = nodes2[i];
    put(0, nd.a.x, nd.a.y, nd.a.z, nd.a.w, nd.b.x, nd.b.y, nd.d.x);
    put(1, nd.b.z, nd.b.w, nd.c.x, nd.c.y, nd.c.z, nd.c.w, nd.d.y);
    n = 2;
    if (greedy) {
    for (int round = 0; round < 2; round++) {
        int pick = -1;
        float best = -1.0f;
        for (int k = 0; k < n; k++) {
            if (child[k] < 0) continue;                      // leaves cannot be opened
            const float dx = hi[k][0] - lo[k][0], dy = hi[k][1] - lo[k][1], dz = hi[k][2] - lo[k][2];
            const float area = dx * dy + dy * dz + dz * dx;
            if (area > best) { best = area; pick = k; }
        }
        if (pick < 0) break;
        const BvhNode c = nodes2[child[pick]];
        put(pick, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y, c.d.x);
        put(n, c.b.z, c.b.w, c.c.x, c.c.y, c.c.z, c.c.w, c.d.y);
        n++;
    }
    } else {
        const int r0 = child[0], r1 = child[1];
        if (r1 >= 0) { const BvhNode c = nodes2[r1]; put(1, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y, c.d.x); put(n, c.b.z, c.b.w, c.c.x, c.c.y, c.c.z, c.c.w, c.d.y); n++; }
        if (r0 >= 0) { const BvhNode c = nodes2[r0]; put(0, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y, c.d.x); put(n, c.b.z, c.b.w, c.c.x, c.c.y, c.c.z, c.c.w, c.d.y); n++; }
    }
    for (; n < 4; n++) { lo[n][0] = lo[n][1] = lo[n][2] = 3.0e38f; hi[n][0] = hi[n][1] = hi[n][2] = -3.0e38f; child[n] = MCRT_BVH4_EMPTY; }
    Bvh4Node o;
    o.lox = make_float4(lo[0][0], lo[1][0], lo[2][0], lo[3][0]); o.loy = make_float4(lo[0][1], lo[1][1], lo[2][1], lo[3][1]);
    o.loz = make_float4(lo[0][2], lo[1][2], lo[2][2], lo[3][2]);
    o.hix = make_float4(hi[0][0], hi[1][0], hi[2][0], hi[3][0]); o.hiy = make_float4(hi[0][1], hi[1][1], hi[2][1], hi[3][1]);
    o.hiz = make_float4(hi[0][2], hi[1][2], hi[2][2], hi[3][2]);
    o.child = make_int4(child[0], child[1], child[2], child[3]);
    o.pad = make_int4(0, 0, 0, 0);
    nodes4[i] = o;
}

}  // namespace

static int bvh4_depth(const std::vector<Bvh4Node>& h)
{
    std::vector<std::pair<int, int>> stack;          // (node, depth)
    stack.emplace_back(0, 1);
    int deepest = 0;
    while (!stack.empty()) {
        const std::pair<int, int> top = stack.back();
        stack.pop_back();
        if (top.second > deepest) deepest = top.second;
        const int c[4] = {h[top.first].child.x, h[top.first].child.y, h[top.first].child.z, h[top.first].child.w};
        for (int k = 0; k < 4; k++)
            if (c[k] >= 0 && c[k] != MCRT_BVH4_EMPTY) stack.emplace_back(c[k], top.second + 1);
    }
    return deepest;
}

// Greedy collapse first; if its longest root-to-leaf chain does not fit the traversal stack (3 pushes per level), the fixed-shape
// collapse (half the BVH2 depth) is used instead.  *depth4_out = the chain length of the tree that was kept.
cudaError_t collapse_bvh4(const BvhNode* d_nodes2, int n_nodes, Bvh4Node** d_nodes4_out, cudaStream_t stream, int* depth4_out)
{
    *d_nodes4_out = nullptr;
    if (depth4_out) *depth4_out = 0;
    if (n_nodes <= 0) return cudaSuccess;
    Bvh4Node* d4 = nullptr;
    cudaError_t e = cudaMalloc(&d4, sizeof(Bvh4Node) * (size_t)n_nodes);
    if (e != cudaSuccess) return e;
    std::vector<Bvh4Node> h((size_t)n_nodes);
    int depth = 0;
    for (int greedy = MCRT_BVH4_GREEDY; greedy >= 0; greedy--) {
        k_collapse_bvh4<<<(n_nodes + 255) / 256, 256, 0, stream>>>(d_nodes2, n_nodes, d4, greedy);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) e = cudaMemcpy(h.data(), d4, sizeof(Bvh4Node) * (size_t)n_nodes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { cudaFree(d4); return e; }
        depth = bvh4_depth(h);                             // start-up only
        if (3 * depth + 1 <= MCRT_STACK_DEPTH4) break;
    }
    if (depth4_out) *depth4_out = depth;
    *d_nodes4_out = d4;
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// BVH2 -> BVH8 (round 2).  Top-down, one launch per level of the wide tree.  A work item = (BVH2 node, BVH8 node it
// becomes).  The item starts from the BVH2 node's two children and up to six times replaces the inner child with the
// largest surface area by that child's two children (the greedy collapse of k_collapse_bvh4, taken to 8); the resulting
// 2..8 children are assigned to the 8 slots by octant -- slot bit a set = the child lies towards +axis a of the node's
// centre (greedy on dot(child centre - node centre, slot direction), Ylitie et al. 2017) -- so that a ray can visit the
// slots in the order (slot XOR ray octant) without sorting entry distances.  Inner children get consecutive node indices
// and leaf children consecutive triangle slots (one atomicAdd each per node), in slot order: a child is addressed by the
// node's base + the number of like children in lower slots, and the traversal stack holds one (base, masks) entry per
// node instead of one entry per child.
// ------------------------------------------------------------------------------------------------
struct WideItem { int node2; int node8; };

__global__ void __launch_bounds__(128) k_collapse_bvh8_level(const BvhNode* __restrict__ nodes2, const TriSlot* __restrict__ tris_in,
                                                            const WideItem* __restrict__ items_in, const int n_in,
                                                            WideItem* __restrict__ items_out, int* __restrict__ counters /* {nodes, tris, out} */,
                                                            Bvh8Node* __restrict__ nodes8, TriSlot* __restrict__ tris_out)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_in) return;
    const WideItem it = items_in[w];
    float lo[8][3], hi[8][3];
    int ref[8];
    int n = 2;
    auto put = [&](int at, float lx, float ly, float lz, float hx, float hy, float hz, int r) {
        lo[at][0] = lx; lo[at][1] = ly; lo[at][2] = lz; hi[at][0] = hx; hi[at][1] = hy; hi[at][2] = hz; ref[at] = r;
    };
    {
        const BvhNode nd = nodes2[it.node2];
        put(0, nd.a.x, nd.a.y, nd.a.z, nd.a.w, nd.b.x, nd.b.y, nd.d.x);
        put(1, nd.b.z, nd.b.w, nd.c.x, nd.c.y, nd.c.z, nd.c.w, nd.d.y);
    }
    while (n < 8) {
        int pick = -1;
        float best = -1.0f;
        for (int k = 0; k < n; k++) {
            if (ref[k] < 0) continue;                            // leaves cannot be opened
            const float dx = hi[k][0] - lo[k][0], dy = hi[k][1] - lo[k][1], dz = hi[k][2] - lo[k][2];
            const float area = dx * dy + dy * dz + dz * dx;
            if (area > best) { best = area; pick = k; }
        }
        if (pick < 0) break;
        const BvhNode c = nodes2[ref[pick]];
        put(pick, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y, c.d.x);
        put(n, c.b.z, c.b.w, c.c.x, c.c.y, c.c.z, c.c.w, c.d.y);
        n++;
    }
    // node centre, child centres
    float nlo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, nhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = 0; k < n; k++)
        for (int a = 0; a < 3; a++) { nlo[a] = fminf(nlo[a], lo[k][a]); nhi[a] = fmaxf(nhi[a], hi[k][a]); }
    float cc[8][3];
    for (int k = 0; k < n; k++)
        for (int a = 0; a < 3; a++) cc[k][a] = 0.5f * (lo[k][a] + hi[k][a]) - 0.5f * (nlo[a] + nhi[a]);
    // greedy octant assignment: repeatedly take the (child, free slot) pair with the largest dot(centre offset, slot direction)
    int slot_of[8];
    unsigned used_slots = 0u, done_children = 0u;
    for (int round = 0; round < n; round++) {
        float bestc = -3.0e38f;
        int bk = 0, bs = 0;
        for (int k = 0; k < n; k++) {
            if (done_children & (1u << k)) continue;
            for (int sl = 0; sl < 8; sl++) {
                if (used_slots & (1u << sl)) continue;
                const float cost = ((sl & 1) ? cc[k][0] : -cc[k][0]) + ((sl & 2) ? cc[k][1] : -cc[k][1]) + ((sl & 4) ? cc[k][2] : -cc[k][2]);
                if (cost > bestc) { bestc = cost; bk = k; bs = sl; }
            }
        }
        slot_of[bk] = bs; used_slots |= 1u << bs; done_children |= 1u << bk;
    }
    unsigned imask = 0u, lmask = 0u;
    for (int k = 0; k < n; k++) { if (ref[k] >= 0) imask |= 1u << slot_of[k]; else lmask |= 1u << slot_of[k]; }
    const int ni = __popc(imask), nl = __popc(lmask);
    const int cbase = ni ? atomicAdd(&counters[0], ni) : 0;
    const int tbase = nl ? atomicAdd(&counters[1], nl) : 0;
    const int obase = ni ? atomicAdd(&counters[2], ni) : 0;
    Bvh8Node o;
    for (int sl = 0; sl < 8; sl++) {
        o.lox[sl] = o.loy[sl] = o.loz[sl] = 3.0e38f;             // empty slot: inverted box, never hit
        o.hix[sl] = o.hiy[sl] = o.hiz[sl] = -3.0e38f;
    }
    for (int k = 0; k < n; k++) {
        const int sl = slot_of[k];
        o.lox[sl] = lo[k][0]; o.loy[sl] = lo[k][1]; o.loz[sl] = lo[k][2];
        o.hix[sl] = hi[k][0]; o.hiy[sl] = hi[k][1]; o.hiz[sl] = hi[k][2];
        if (ref[k] >= 0) {
            const int rank = __popc(imask & ((1u << sl) - 1u));
            items_out[obase + rank] = WideItem{ref[k], cbase + rank};
        } else {
            const int rank = __popc(lmask & ((1u << sl) - 1u));
            const int code = -ref[k] - 1;                           // leaf of exactly one triangle (MCRT_LEAF_MAX == 1)
            tris_out[tbase + rank] = tris_in[code >> 2];
        }
    }
    o.child_base = cbase; o.tri_base = tbase; o.masks = imask | (lmask << 8);
    for (int k = 0; k < 13; k++) o.pad[k] = 0;
    nodes8[it.node8] = o;
}

cudaError_t collapse_bvh8(const BvhNode* d_nodes2, int n_nodes2, const TriSlot* d_tris_in, int n_tri, Bvh8Node** d_nodes8_out,
                          TriSlot** d_tris8_out, cudaStream_t stream, int* depth8_out, int* n_nodes8_out)
{
    static_assert(sizeof(Bvh8Node) == 256, "Bvh8Node must be two cache lines");
    static_assert(MCRT_LEAF_MAX == 1, "the 8-wide collapse expects single-triangle leaves");
    *d_nodes8_out = nullptr; *d_tris8_out = nullptr;
    if (depth8_out) *depth8_out = 0;
    if (n_nodes8_out) *n_nodes8_out = 0;
    if (n_nodes2 <= 0 || n_tri <= 0) return cudaSuccess;
    Bvh8Node* d8 = nullptr;
    TriSlot* t8 = nullptr;
    WideItem *qa = nullptr, *qb = nullptr;
    int* d_cnt = nullptr;
    cudaError_t e = cudaMalloc(&d8, sizeof(Bvh8Node) * (size_t)n_nodes2);
    if (e == cudaSuccess) e = cudaMalloc(&t8, sizeof(TriSlot) * (size_t)n_tri);
    if (e == cudaSuccess) e = cudaMalloc(&qa, sizeof(WideItem) * (size_t)n_nodes2);
    if (e == cudaSuccess) e = cudaMalloc(&qb, sizeof(WideItem) * (size_t)n_nodes2);
    if (e == cudaSuccess) e = cudaMalloc(&d_cnt, sizeof(int) * 3);
    int depth = 0, total_nodes = 1;
    if (e == cudaSuccess) {
        const WideItem root{0, 0};
        int h_cnt[3] = {1, 0, 0};                                  // node 0 = the root is taken
        e = cudaMemcpyAsync(qa, &root, sizeof(root), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cnt, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, stream);
        int n_in = 1;
        while (e == cudaSuccess && n_in > 0) {
            depth++;
            k_collapse_bvh8_level<<<(n_in + 127) / 128, 128, 0, stream>>>(d_nodes2, d_tris_in, qa, n_in, qb, d_cnt, d8, t8);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            n_in = h_cnt[2];
            total_nodes = h_cnt[0];
            if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt + 2, 0, sizeof(int), stream);
            WideItem* t = qa; qa = qb; qb = t;
        }
        if (e == cudaSuccess && h_cnt[1] != n_tri) e = cudaErrorUnknown;     // every triangle is the leaf child of exactly one node
    }
    cudaFree(qa); cudaFree(qb); cudaFree(d_cnt);
    if (e != cudaSuccess) { cudaFree(d8); cudaFree(t8); return e; }
    *d_nodes8_out = d8; *d_tris8_out = t8;
    if (depth8_out) *depth8_out = depth;
    if (n_nodes8_out) *n_nodes8_out = total_nodes;
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// PLOC (parallel locally-ordered clustering, Meister & Bittner 2018), round 2: a SAH-quality hierarchy ON THE DEVICE over the
// same Morton-sorted triangles as the Karras tree.  Clusters start as the sorted triangles; every iteration each cluster looks
// `radius` places left and right in the (still Morton-ordered) cluster array for the partner with the smallest merged surface
// area, mutual nearest neighbours merge into a new inner node, and the array is compacted in order.  The binary tree comes out
// in the layout k_collapse_bvh4 / k_collapse_bvh8 expect (BvhNode with both child boxes, root = node 0: node ids are handed
// out downwards from n - 2 and the last merge is the root).  Option bvh_builder = 2; bit-identical results
// (test_traversal_options_do_not_change_results).  MEASURED WORSE than the Karras tree on the ircad11 stand-in meshes
// (20.8 instead of 17.5 node visits per query, trace 2.97 vs 2.57 ms; the host binned-SAH tree: 16.1 visits, 2.44 ms;
// profiles/r02r_ab_ploc.txt): the regular triangulation of those meshes makes the merged-area criterion tie massively, few
// pairs are mutual per iteration and the clusters grow as chains.  Kept as an option, not the default.
// ------------------------------------------------------------------------------------------------
#ifndef MCRT_PLOC_RADIUS
#define MCRT_PLOC_RADIUS 16
#endif
__device__ __forceinline__ float merged_area(float4 alo, float4 ahi, float4 blo, float4 bhi)
{
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y),
                dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_ploc_init(const unsigned int* __restrict__ vals, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, int n,
                            float4* __restrict__ c_lo, float4* __restrict__ c_hi, int* __restrict__ c_ref)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned int t = vals[k];
    c_lo[k] = box_lo[t]; c_hi[k] = box_hi[t];
    c_ref[k] = -(1 + k * 4);                                     // leaf of the single triangle in sorted slot k
}

__global__ void k_ploc_nearest(const float4* __restrict__ c_lo, const float4* __restrict__ c_hi, int m, int* __restrict__ nn)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float4 lo = c_lo[i], hi = c_hi[i];
    float best = 3.0e38f;
    int bj = -1;
    const int j0 = i - MCRT_PLOC_RADIUS < 0 ? 0 : i - MCRT_PLOC_RADIUS, j1 = i + MCRT_PLOC_RADIUS >= m ? m - 1 : i + MCRT_PLOC_RADIUS;
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const float a = merged_area(lo, hi, c_lo[j], c_hi[j]);
        if (a < best) { best = a; bj = j; }                       // ties: the lower index, on both sides -> mutual pairs stay mutual
    }
    nn[i] = bj;
}

__global__ void k_ploc_merge(const float4* __restrict__ c_lo, const float4* __restrict__ c_hi, const int* __restrict__ c_ref, const int* __restrict__ nn,
                             int m, int* __restrict__ next_node, BvhNode* __restrict__ nodes, int* __restrict__ parent_internal,
                             int* __restrict__ parent_leaf, float4* __restrict__ o_lo, float4* __restrict__ o_hi, int* __restrict__ o_ref,
                             int* __restrict__ keep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i];
    float4 lo = c_lo[i], hi = c_hi[i];
    int ref = c_ref[i], k = 1;
    if (j >= 0 && nn[j] == i) {
        if (i < j) {
            const float4 blo = c_lo[j], bhi = c_hi[j];
            const int rj = c_ref[j];
            const int id = atomicSub(next_node, 1);
            BvhNode nd;
            nd.a = make_float4(lo.x, lo.y, lo.z, hi.x);
            nd.b = make_float4(hi.y, hi.z, blo.x, blo.y);
            nd.c = make_float4(blo.z, bhi.x, bhi.y, bhi.z);
            nd.d = make_int4(ref, rj, 0, 0);
            nodes[id] = nd;
            if (ref >= 0) parent_internal[ref] = id; else parent_leaf[(-ref - 1) >> 2] = id;
            if (rj >= 0) parent_internal[rj] = id; else parent_leaf[(-rj - 1) >> 2] = id;
            lo = make_float4(fminf(lo.x, blo.x), fminf(lo.y, blo.y), fminf(lo.z, blo.z), 0.f);
            hi = make_float4(fmaxf(hi.x, bhi.x), fmaxf(hi.y, bhi.y), fmaxf(hi.z, bhi.z), 0.f);
            ref = id;
        } else {
            k = 0;                                                // absorbed by cluster j
        }
    }
    o_lo[i] = lo; o_hi[i] = hi; o_ref[i] = ref; keep[i] = k;
}

__global__ void k_ploc_compact(const float4* __restrict__ o_lo, const float4* __restrict__ o_hi, const int* __restrict__ o_ref,
                               const int* __restrict__ keep, const int* __restrict__ pos, int m, float4* __restrict__ c_lo,
                               float4* __restrict__ c_hi, int* __restrict__ c_ref, int* __restrict__ m_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (keep[i]) { const int p = pos[i]; c_lo[p] = o_lo[i]; c_hi[p] = o_hi[i]; c_ref[p] = o_ref[i]; }
    if (i == m - 1) *m_out = pos[i] + keep[i];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

cudaError_t build_lbvh(const float* h_tri_local, const int32_t* h_tri_mesh, int n_tri, const DevMesh* d_meshes, cudaStream_t stream,
                       LbvhResult* out, int ploc)
{
    cudaError_t err = cudaSuccess;
    memset(out, 0, sizeof(*out));
    float* d_tri_local = nullptr; int32_t* d_tri_mesh = nullptr;
    float4 *d_lo = nullptr, *d_hi = nullptr, *d_nlo = nullptr, *d_nhi = nullptr;
    int2* d_range = nullptr;
    int *d_bounds = nullptr, *d_left = nullptr, *d_right = nullptr, *d_pi = nullptr, *d_pl = nullptr, *d_arr = nullptr, *d_depth = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    unsigned int *d_vals = nullptr, *d_vals2 = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    // PLOC working set
    float4 *d_clo = nullptr, *d_chi = nullptr, *d_olo = nullptr, *d_ohi = nullptr;
    int *d_cref = nullptr, *d_oref = nullptr, *d_nn = nullptr, *d_keep = nullptr, *d_pos = nullptr, *d_ctr = nullptr;
    void* d_scan_tmp = nullptr;
    size_t scan_bytes = 0;
    const int n = n_tri;
    const int B = 256, G = (n + B - 1) / B;
    int h_bounds[6];
    int h_depth[2] = {0, 0};
    if (n <= 0) return cudaSuccess;

    CK(cudaMalloc(&d_tri_local, sizeof(float) * 9 * (size_t)n));
    CK(cudaMalloc(&d_tri_mesh, sizeof(int32_t) * (size_t)n));
    CK(cudaMemcpyAsync(d_tri_local, h_tri_local, sizeof(float) * 9 * (size_t)n, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_tri_mesh, h_tri_mesh, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, stream));
    CK(cudaMalloc(&d_lo, sizeof(float4) * (size_t)n));
    CK(cudaMalloc(&d_hi, sizeof(float4) * (size_t)n));
    CK(cudaMalloc(&d_bounds, sizeof(int) * 6));
    for (int a = 0; a < 3; a++) { h_bounds[a] = 0x7f7fffff; h_bounds[3 + a] = (int)0x80800000 /* ord(-FLT_MAX) */; }
    h_bounds[3] = h_bounds[4] = h_bounds[5] = (int)(0xff7fffffu ^ 0x7fffffffu);
    CK(cudaMemcpyAsync(d_bounds, h_bounds, sizeof(h_bounds), cudaMemcpyHostToDevice, stream));
    k_tri_bounds<<<G, B, 0, stream>>>(d_tri_local, d_tri_mesh, d_meshes, n, d_lo, d_hi, d_bounds);
    CK(cudaGetLastError());
    CK(cudaMalloc(&out->tris, sizeof(TriSlot) * (size_t)n));
    CK(cudaMalloc(&d_vals, sizeof(unsigned int) * (size_t)n));
    if (n >= 2) {
        CK(cudaMalloc(&d_keys, sizeof(unsigned long long) * (size_t)n));
        CK(cudaMalloc(&d_keys2, sizeof(unsigned long long) * (size_t)n));
        CK(cudaMalloc(&d_vals2, sizeof(unsigned int) * (size_t)n));
        k_morton<<<G, B, 0, stream>>>(d_lo, d_hi, d_bounds, n, d_keys2, d_vals2);
        CK(cudaGetLastError());
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys2, d_keys, d_vals2, d_vals, n, 0, 63, stream));
        CK(cudaMalloc(&d_tmp, tmp_bytes));
        CK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys2, d_keys, d_vals2, d_vals, n, 0, 63, stream));
        CK(cudaMalloc(&d_left, sizeof(int) * (size_t)(n - 1)));
        CK(cudaMalloc(&d_right, sizeof(int) * (size_t)(n - 1)));
        CK(cudaMalloc(&d_pi, sizeof(int) * (size_t)(n - 1)));
        CK(cudaMalloc(&d_pl, sizeof(int) * (size_t)n));
        CK(cudaMalloc(&d_range, sizeof(int2) * (size_t)(n - 1)));
        CK(cudaMalloc(&d_arr, sizeof(int) * (size_t)(n - 1)));
        CK(cudaMalloc(&d_depth, sizeof(int) * 2));
        CK(cudaMalloc(&d_nlo, sizeof(float4) * (size_t)(n - 1)));
        CK(cudaMalloc(&d_nhi, sizeof(float4) * (size_t)(n - 1)));
        CK(cudaMemsetAsync(d_arr, 0, sizeof(int) * (size_t)(n - 1), stream));
        CK(cudaMemsetAsync(d_depth, 0, sizeof(int) * 2, stream));
        if (!ploc) {
        k_karras<<<G, B, 0, stream>>>(d_keys, n, d_left, d_right, d_pi, d_pl, d_range);
        CK(cudaGetLastError());
        k_refit<<<G, B, 0, stream>>>(d_vals, d_lo, d_hi, n, d_left, d_right, d_pi, d_pl, d_arr, d_nlo, d_nhi, d_depth);
        CK(cudaGetLastError());
        k_leaf_depth<<<G, B, 0, stream>>>(n, d_pi, d_pl, d_depth + 1);
        CK(cudaGetLastError());
        CK(cudaMalloc(&out->nodes, sizeof(BvhNode) * (size_t)(n - 1)));
        k_emit_nodes<<<G, B, 0, stream>>>(d_vals, d_lo, d_hi, n, d_left, d_right, d_nlo, d_nhi, d_range, out->nodes);
        CK(cudaGetLastError());
        } else {
            // ---- PLOC over the Morton-sorted triangles ----
            CK(cudaMalloc(&out->nodes, sizeof(BvhNode) * (size_t)(n - 1)));
            CK(cudaMalloc(&d_clo, sizeof(float4) * (size_t)n)); CK(cudaMalloc(&d_chi, sizeof(float4) * (size_t)n));
            CK(cudaMalloc(&d_olo, sizeof(float4) * (size_t)n)); CK(cudaMalloc(&d_ohi, sizeof(float4) * (size_t)n));
            CK(cudaMalloc(&d_cref, sizeof(int) * (size_t)n)); CK(cudaMalloc(&d_oref, sizeof(int) * (size_t)n));
            CK(cudaMalloc(&d_nn, sizeof(int) * (size_t)n)); CK(cudaMalloc(&d_keep, sizeof(int) * (size_t)n)); CK(cudaMalloc(&d_pos, sizeof(int) * (size_t)n));
            CK(cudaMalloc(&d_ctr, sizeof(int) * 2));
            CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_keep, d_pos, n, stream));
            CK(cudaMalloc(&d_scan_tmp, scan_bytes ? scan_bytes : 16));
            int h_ctr[2] = {n - 2, n};                             // next node id (handed out downwards), cluster count
            CK(cudaMemcpyAsync(d_ctr, h_ctr, sizeof(h_ctr), cudaMemcpyHostToDevice, stream));
            CK(cudaMemsetAsync(d_pi, 0xff, sizeof(int) * (size_t)(n - 1), stream));       // parent of the root stays -1
            k_ploc_init<<<G, B, 0, stream>>>(d_vals, d_lo, d_hi, n, d_clo, d_chi, d_cref);
            CK(cudaGetLastError());
            int m = n, iterations = 0;
            while (m > 1) {
                const int Gm = (m + B - 1) / B;
                k_ploc_nearest<<<Gm, B, 0, stream>>>(d_clo, d_chi, m, d_nn);
                k_ploc_merge<<<Gm, B, 0, stream>>>(d_clo, d_chi, d_cref, d_nn, m, d_ctr, out->nodes, d_pi, d_pl, d_olo, d_ohi, d_oref, d_keep);
                CK(cudaGetLastError());
                CK(cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_bytes, d_keep, d_pos, m, stream));
                k_ploc_compact<<<Gm, B, 0, stream>>>(d_olo, d_ohi, d_oref, d_keep, d_pos, m, d_clo, d_chi, d_cref, d_ctr + 1);
                CK(cudaGetLastError());
                int m_new = 0;
                CK(cudaMemcpyAsync(&m_new, d_ctr + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                if (m_new >= m || ++iterations > 4096) { err = cudaErrorUnknown; goto done; }     // every iteration merges at least the globally closest pair
                m = m_new;
            }
            k_leaf_depth<<<G, B, 0, stream>>>(n, d_pi, d_pl, d_depth + 1);
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(h_depth, d_depth, sizeof(h_depth), cudaMemcpyDeviceToHost, stream));
    } else {
        const unsigned int zero = 0;
        CK(cudaMemcpyAsync(d_vals, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
    }
    k_emit_tris<<<G, B, 0, stream>>>(d_vals, d_tri_local, d_tri_mesh, n, out->tris);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_bounds, d_bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    out->n_tri = n;
    out->n_nodes = n >= 2 ? n - 1 : 0;
    out->max_depth = h_depth[1];
    out->max_abs = 0.0f;
    for (int a = 0; a < 6; a++) { const float f = fabsf(ord2f(h_bounds[a])); if (f > out->max_abs) out->max_abs = f; }
done:
    cudaFree(d_tri_local); cudaFree(d_tri_mesh); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_nlo); cudaFree(d_nhi);
    cudaFree(d_bounds); cudaFree(d_left); cudaFree(d_right); cudaFree(d_pi); cudaFree(d_pl); cudaFree(d_arr); cudaFree(d_depth);
    cudaFree(d_range); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_vals); cudaFree(d_vals2); cudaFree(d_tmp);
    cudaFree(d_clo); cudaFree(d_chi); cudaFree(d_olo); cudaFree(d_ohi); cudaFree(d_cref); cudaFree(d_oref); cudaFree(d_nn); cudaFree(d_keep);
    cudaFree(d_pos); cudaFree(d_ctr); cudaFree(d_scan_tmp);
    if (err != cudaSuccess) { cudaFree(out->nodes); cudaFree(out->tris); memset(out, 0, sizeof(*out)); }
    return err;
}

}  // namespace mcrt
