// sah_builder.h -- host-side binned-SAH BVH builder (see sah_builder.cpp).
#ifndef MCRT_SAH_BUILDER_H
#define MCRT_SAH_BUILDER_H
#include <atomic>
#include <cstdint>
#include <vector>

namespace mcrt {

struct HostBvhNode { float f[12]; int32_t child[4]; };          // bit-compatible with kernels' BvhNode (64 B)
struct HostTriSlot { float v[9]; int32_t mesh; int32_t tri; };

struct HostBvh {
    std::vector<HostBvhNode> nodes;      // n_tri - 1 nodes, root = 0, pre-order
    std::vector<HostTriSlot> slots;      // triangles in leaf order
    int max_depth = 0;
    float max_abs = 0.0f;
};

// tri_local: 9 floats/triangle (v_obj * scaling); mesh_origin3: body origin per mesh; max_threads <= 0: all host threads.
// The tree does not depend on the number of threads.  cancel (nullable): when it becomes true the build is abandoned (throws).
void build_sah_bvh(const float* tri_local, const int32_t* tri_mesh, int n_tri, const float* mesh_origin3, HostBvh* out, int max_threads = 0,
                   const std::atomic<bool>* cancel = nullptr);

}  // namespace mcrt
#endif
