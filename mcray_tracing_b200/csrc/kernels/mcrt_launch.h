// mcrt_launch.h -- host-callable launchers of the sm_100a kernels (lbvh.cu, trace.cu, image.cu).
#ifndef MCRT_LAUNCH_H
#define MCRT_LAUNCH_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "mcrt_device.cuh"

namespace mcrt {

struct LbvhResult {
    BvhNode* nodes;
    TriSlot* tris;
    int n_tri, n_nodes;
    int max_depth;      // longest leaf-to-root chain = traversal stack bound
    float max_abs;
};
#define MCRT_TRAVERSAL_STACK 62      // BVH2 stack 64; the 4-wide traversal needs 3 * ceil(depth / 2) + 2 <= 96
// 4-wide copy of a BVH2 node array (device LBVH or uploaded host SAH tree); caller frees *d_nodes4_out with cudaFree
// *depth4_out (nullable): longest chain of BVH4 nodes from the root; the traversal needs 3 * depth4 + 1 <= MCRT_STACK_DEPTH4
cudaError_t collapse_bvh4(const BvhNode* d_nodes2, int n_nodes, Bvh4Node** d_nodes4_out, cudaStream_t stream, int* depth4_out = nullptr);

// 8-wide tree from the same BVH2 node array (round 2, the default traversal structure): nodes in breadth-first order with
// the inner children of a node consecutive, triangles re-ordered so the leaf children of a node are consecutive.  Caller
// frees both outputs with cudaFree.  *depth8_out = levels of the wide tree (the traversal stack holds one entry per level).
cudaError_t collapse_bvh8(const BvhNode* d_nodes2, int n_nodes2, const TriSlot* d_tris_in, int n_tri, Bvh8Node** d_nodes8_out,
                          TriSlot** d_tris8_out, cudaStream_t stream, int* depth8_out, int* n_nodes8_out);
// uploads the slot-permutation table of the 8-wide traversal (once per device)
cudaError_t init_trace_kernels();

// builds the device BVH from host triangle data; d_meshes must already be on the device
// ploc = 0: Karras radix tree over the Morton order (LBVH); 1: PLOC agglomerative clustering over the same order (SAH-quality)
cudaError_t build_lbvh(const float* h_tri_local, const int32_t* h_tri_mesh, int n_tri, const DevMesh* d_meshes, cudaStream_t stream,
                       LbvhResult* out, int ploc = 0);

struct FrameDev {
    const PoseTrigDev* poses;            // [n_poses]
    const float2* elem_sincos;           // [elements] (sin a_t, cos a_t)
    const unsigned long long* seed_frame;   // device: {Philox seed, first frame}; kept in HBM so a captured graph stays valid
    int n_poses;
    int frame_offset;                    // index of poses[0] within the call's batch
    int frame_stride;                    // frame (Philox counter) of pose i = first frame + (frame_offset + i) * frame_stride; 1 unless the
                                         // poses of a sweep are dealt out to the GPUs round-robin (option "frame_stride")
};

struct TraceBuffers {
    PathState paths;               // [n_paths]
    DevSegment* segments;          // [n_paths][max_depth]
    int32_t* n_segments;           // [n_paths]
    float* hit_fraction;           // [n_paths][max_depth] or nullptr
    int32_t* hit_mesh;             // [n_paths][max_depth] or nullptr
    int* queue_a;                  // [n_paths]
    int* queue_b;                  // [n_paths]
    int* counters;                 // [max_depth + 1]: counters[b] = live paths entering bounce b; [max_depth] = bounce at which the tail merge started
    int tail_threshold;            // > 0: once at most this many paths are alive, one launch finishes them (no more compaction)
    unsigned long long* trav_counters;   // nullptr, or {BVH node visits, triangle tests} accumulated over the call
    // optional order-preserving compaction (nullptr = off): queue_a is the dense queue every bounce reads, queue_b the sparse
    // one it writes -- every warp compacts the survivors of its 32 paths into its own 32 slots and records how many (no CTA
    // barrier); k_compact then packs the sparse queue into the dense one.  Survivors stay sorted by (pose, element, sample),
    // so the lanes of a warp keep tracing neighbouring rays instead of pieces of unrelated warps glued together by atomic
    // order.  (Round 1 compacted per 128-path CTA chunk behind two CTA barriers and looked paths up by binary search over a
    // chunk prefix: 15 % + 9 % of the bounce kernels' stall samples, profiles/r02a_k_bounce_hot_lines.txt.)
    int* warp_counts;              // [ceil(n_paths / 32)]: survivors of warp chunk w of the current bounce
    int* tile_counts;              // [max_depth][2][n_tiles]: refracted / reflected survivors per tile of 256 warp chunks (8192 paths)
    int n_tiles;
    int group_histories;           // 1: the dense queue holds all refracted survivors before all reflected ones (option "group_histories")
    // optional (nullptr = off): closest hit of bounce 0 per (pose, element).  All samples of an element leave the transducer
    // on the same ray (scene.cpp:84-100), so k_first_hit traces it once and bounce 0 only shades: 2 float4 per element =
    // (fraction, tri_id, mesh, dist_a), (n_raw.xyz, -)
    float4* first_hits;
};

// ---- ray-tree mode: BOTH children of every boundary hit are followed, as in the cited paper, instead
// of the one Monte-Carlo branch this fork of the reference keeps (ray.cpp:84-94).  The wavefront grows level by level; node ids
// are heap indices (root 1, reflected child 2n, refracted child 2n + 1) and key the Philox counter.
// Everything stays ORDERED, so nothing is ever sorted: the rays of a level are processed in (path, node) order; every warp writes the
// children of its 32 rays, in that order, into its own 64 slots of a sparse pool and k_compact_tree packs the slot indices into the
// next level's dense queue (the scheme of k_bounce / k_compact).  A level's segments are stored at [seg_base(level) + queue index],
// so the segments of a scanline at one level are contiguous; level_first / level_end give that range per (level, scanline) and the
// accumulate kernel walks a scanline's segments level by level = in (level, path, node) order.
struct TreeRay {                   // 48 B
    float4 origin_intensity;
    float4 dir_state;              // direction + packed (medium, outside medium, depth)
    double distance;
    int path, node;
};
struct TreeBuffers {
    TreeRay* rays_a;               // sparse pools, [2 * ray_capacity + 64]: 64 slots per warp chunk of the level that wrote them
    TreeRay* rays_b;
    int* queue_a;                  // dense queues, [ray_capacity]: pool slots of the rays entering a level, in (path, node) order
    int* queue_b;
    int* warp_counts;              // [ray_capacity / 32 + 1]: children written by warp chunk w of the current level
    int* tile_counts;              // [max_depth][n_tiles]: children per tile of 256 warp chunks
    int n_tiles;
    DevSegment* segments;          // [seg_capacity]: level 0 first, then level 1, ...; queue order inside a level
    unsigned long long* keys;      // [seg_capacity]: (path << 20 | node) of the segment in the same slot (mcrt_trace_tree_debug)
    int* level_first;              // [max_depth][n_scanlines]: first segment slot of the scanline at that level
    int* level_end;                // [max_depth][n_scanlines]: one past its last slot (0: the scanline has no ray at that level)
    int* counters;                 // [max_depth + 3]: rays entering level l; [max_depth + 1] segments; [max_depth + 2] overflow flag
    unsigned long long* trav_counters;
    int ray_capacity, seg_capacity, n_scanlines;
};
// traces the trees of all paths of the uploaded poses; fills segments / keys / level_first / level_end / counters
void launch_trace_tree(const SceneDev& sc, const AcqDev& aq, const FrameDev& fr, const TreeBuffers& tb, int sm_count, cudaStream_t stream,
                       int* launches);
// echo accumulation of tree segments: a warp per scanline, a lane per SEGMENT (a segment is self-contained: start point, start
// time, initial intensity), rounds of 32 segments through the windowed shared-memory ring of k_accumulate_win; no columns in HBM
cudaError_t launch_accumulate_tree(const SceneDev& sc, const AcqDev& aq, const float2* d_volume, const TreeBuffers& tb, int n_poses,
                                   float* d_rf, unsigned long long* d_steps, cudaStream_t stream, int* launches);

// generate + max_depth x (intersect, shade, compact): scene::cast_rays (scene.cpp:50-183)
void launch_trace(const SceneDev& sc, const AcqDev& aq, const FrameDev& fr, const TraceBuffers& tb, int sm_count, cudaStream_t stream,
                  int* launches);
void launch_closest_hit(const SceneDev& sc, int64_t n, const float* d_from, const float* d_to, int32_t* d_tri, int32_t* d_mesh,
                        float* d_frac, float* d_point, float* d_normal, cudaStream_t stream);
void launch_elements(const AcqDev& aq, const FrameDev& fr, float* d_pos, float* d_dir, cudaStream_t stream);

// main.cpp:106-144: segments -> raw RF, scanline-major [n_poses*elements][rows]
// d_columns: accumulate_columns_bytes() bytes of HBM scratch: one private RF column per path,
// columns[scanline][row][sample].
cudaError_t init_image_kernels();      // once per device, outside any stream capture
size_t accumulate_columns_bytes(const AcqDev& aq, int n_poses);
// kernels launch_post will issue for this geometry (1 = fused shared-memory path, 3 = axial + lateral + envelope)
int post_launch_count(int cols, int rows, int n_lateral, int flags, int n_images);
cudaError_t launch_accumulate(const SceneDev& sc, const AcqDev& aq, const float2* d_volume, const DevSegment* d_segments,
                              const int32_t* d_nseg, int n_poses, float* d_rf, unsigned long long* d_steps, float* d_columns,
                              cudaStream_t stream, int* launches);

// rf_image::convolve + envelope (rfimage.h:93-123, 54-91) on [n_images][cols][rows]; flags bit0 convolve, bit1 envelope.
// d_tmp0/d_tmp1: scratch of the same size as the image batch.  Result always lands in d_out.
void set_long_scanline_ct(bool on);   // A/B switch of the round-2 long-scanline post kernels (default on)
void launch_post(const float* d_in, int n_images, int cols, int rows, const float* d_axial, int n_axial, const float* d_lateral,
                 int n_lateral, int flags, float* d_tmp0, float* d_tmp1, float* d_out, cudaStream_t stream, int* launches,
                 int col_offset = 0, int cols_total = 0,     // scanline-block runs: global index of scanline 0 / global scanline count
                 const float* d_lateral_by_row = nullptr,   // depth-dependent lateral PSF: [n_lateral][rows] taps (unfused kernels)
                 int in_pitch = 0,                          // row stride of d_in in floats (0 = rows); flags & 1 == 0 needs a dense input
                 const float* h_axial = nullptr, const float* h_lateral = nullptr,    // host copies of the taps: enable the TMA-staged
                                                                                      // kernel (taps travel as kernel parameters)
                 const unsigned long long* d_out_target = nullptr);  // device {base pointer, image stride in floats}: the result goes there instead
                                                                     // of d_out -- ONLY honoured when post_writes_through_target() says so
// true when launch_post would take the kernel that can write through a device-resident output target (the TMA-staged fused kernel)
bool post_writes_through_target(int n_images, int cols, int rows, int in_pitch, int n_axial, int n_lateral, int flags, const float* h_axial,
                                const float* h_lateral, bool by_row);
// row pitch the raw RF image should have so that launch_post can stage it with TMA bulk copies (rows rounded up to 4 floats), or rows
int post_preferred_pitch(int rows, int n_axial, int n_lateral);
// exhaustive device check (all 2^32 float bit patterns) that the 3-instruction FMA division reproduces the voxel index of
// coord / resolution for this resolution; enables AcqDev::voxel_fma_division
cudaError_t validate_fma_division(float resolution, bool* ok);
// can launch_accumulate use the windowed kernel for this scene / acquisition (then d_columns may be nullptr)?
bool accumulate_windowed_supported(const SceneDev& sc, const AcqDev& aq);
// elevational PSF: out[img][px] = sum_j in[img * n_planes + j][px] * w[j] (raw RF images of the n_planes ray fans of one frame)
void launch_elevation_combine(const float* d_in, int n_out_images, int64_t px_per_image, int n_planes, const float* d_w, float* d_out,
                              cudaStream_t stream, int* launches);
// rfimage.h:127-136 (commented out in the reference): I = log10(I+1)/log10(max+1) per image, in place
void launch_log_compress(float* d_img, int n_images, int64_t px_per_image, int* d_max_bits, cudaStream_t stream, int* launches);
// B-mode display chain on the envelope image [n][cols][rows]: TGC gain table (per row), log compression to a dynamic range;
// d_out receives values in [0, 1].  launch_quantize8: x255, round, saturate -> uint8
void launch_bmode(const float* d_env, int n_images, int cols, int rows, const float* d_gain, float dynamic_range_db, float* d_out,
                  int* d_max_bits, cudaStream_t stream, int* launches);
void launch_quantize8(const float* d_in, int64_t n, unsigned char* d_out, cudaStream_t stream, int* launches);
// [n][cols][rows] -> [n][rows][cols]
void launch_transpose(const float* d_in, int n_images, int cols, int rows, float* d_out, cudaStream_t stream, int* launches);
// cv::remap (rfimage.h:139) with the precomputed maps; input scanline-major [n][cols][rows]
void launch_scan_convert(const float* d_rf, int n_images, int cols, int rows, const float* d_map_x, const float* d_map_y, int scan_rows,
                         int scan_cols, float* d_out, cudaStream_t stream, int* launches);
void launch_numerics_probe(int op, int64_t n, const double* d_a, const double* d_b, double* d_out, cudaStream_t stream);

}  // namespace mcrt
#endif
