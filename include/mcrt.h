/* mcrt.h -- C ABI of libmcrt.so: the B200-native (sm_100a) replacement for the per-frame
 * simulation hot path of thepochynsons/MCRay-Tracing.
 *
 * The reference exports no plugin/FFI interface: it is one executable (`mattausch <scene>`,
 * src/main.cpp:42-50) whose frame loop (src/main.cpp:92-152) calls
 *     scene::cast_rays<S,E>(transducer&)            src/scene.h:29-30, src/scene.cpp:50-183
 *     the echo accumulation loop                     src/main.cpp:106-144
 *     rf_image::{clear,add_echo,convolve,envelope,postprocess}   src/rfimage.h:33-140,161-164
 * on objects built once by scene::scene(json, transducer&) (src/scene.cpp:16-31),
 * transducer<N>::transducer (src/transducer.h:24-62), psf<...>::psf (src/psf.h:34-58) and
 * volume<...>::volume (src/volume.h:19-35).  This header is that in-process seam as a C ABI:
 * plain pointers and sizes, no C++/torch types, negative return codes instead of exceptions
 * (the reference reports errors as exceptions caught at src/main.cpp:154-159).
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * All compute runs in hand-written CUDA kernels; there is NO CPU fallback: every entry point
 * that needs the GPU fails with MCRT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MCRT_H
#define MCRT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCRT_OK 0
#define MCRT_ERR_INVALID (-1)   /* bad argument */
#define MCRT_ERR_SCENE (-2)     /* "Error while loading scene: ..." (src/scene.cpp:23-26) */
#define MCRT_ERR_CUDA (-3)      /* CUDA runtime / no usable device */
#define MCRT_ERR_NOMEM (-4)

typedef struct mcrt_ctx mcrt_ctx; /* opaque; owns the device BVH, scatterer volume, streams, graphs */

/* probe pose: replaces transducer::setPosition / the const `angles` (src/transducer.h:120-137) */
typedef struct mcrt_pose {
    float pos[3];        /* "transducerPosition", world cm (src/main.cpp:65,72) */
    float angles_deg[3]; /* "transducerAngles" x,y,z degrees (src/main.cpp:68-69) */
} mcrt_pose;

/* Acquisition parameters: the constexpr block of src/main.cpp:23-37,54 made runtime.
 * mcrt_default_params() fills in the reference's values. */
typedef struct mcrt_params {
    int32_t elements;        /* transducer_elements 512          main.cpp:26 */
    int32_t samples;         /* samples_te 5                     main.cpp:27 */
    int32_t max_depth;       /* ray::max_depth 10 (<= 16)        ray.h:23    */
    float frequency_mhz;     /* transducer_frequency 4.5f        main.cpp:24 */
    double radius_cm;        /* transducer_radius 3              main.cpp:29 */
    double fov_deg;          /* transducer_amplitude 60          main.cpp:28 */
    double depth_cm;         /* ultrasound_depth 15              main.cpp:30 */
    uint32_t speed_of_sound; /* 1500                             main.cpp:23 */
    uint32_t resolution_um;  /* 145: psf / scatterer-volume grid main.cpp:33 */
    int32_t psf_axial;       /* 7  (odd)                         main.cpp:34 */
    int32_t psf_lateral;     /* 13 (odd)                         main.cpp:34 */
    float psf_var_x;         /* 0.05f                            main.cpp:54 */
    float psf_var_y;         /* 0.2f                             main.cpp:54 */
    int32_t deterministic;   /* 1: roughness off (cos theta' = 1, thickness q = 0); DESIGN.md "Deterministic mode" */
    int32_t scan_rows;       /* 400                              rfimage.h:26 */
    int32_t scan_cols;       /* 500                              rfimage.h:26 */
    float axial_scale;       /* 1.0f; >1 refines the axial sampling grid (extension) */
    int32_t rf_layout;       /* 0: rf_out[pose][element][row] (scanline-major, native);
                                1: rf_out[pose][row][element] = cv::Mat(max_rows, columns), rfimage.h:24 */
} mcrt_params;

/* A ray segment (ray_physics::segment, src/ray.h:28-36) plus parity-debug fields. */
typedef struct mcrt_segment {
    float from[3];
    float to[3];
    float dir[3];
    float reflected_intensity;
    float initial_intensity;
    float attenuation;
    double distance_traveled; /* mm, from the transducer to the start of the segment */
    int32_t media_id;         /* index into the scene's material array */
    int32_t tri_id;           /* global triangle id hit at the end of the segment, -1 = miss */
    int32_t mesh_id;          /* -1 = miss */
    float hit_fraction;       /* closest-hit fraction along [from+0.1*dir, to]; 1 on a miss */
} mcrt_segment;

/* In-memory scene description (what scene::parse_config + load_mesh_from_obj produce,
 * src/scene.cpp:185-247, src/objloader.h:154-161) for callers that do not go through files. */
typedef struct mcrt_scene_arrays {
    int32_t n_materials;
    const float* materials8;              /* n_materials x {impedance, attenuation, mu0, mu1, sigma,
                                             specularity, shininess, thickness}  (src/mesh.h:7-10) */
    int32_t starting_material;
    int32_t n_meshes;
    const int32_t* mesh_material_inside;  /* src/mesh.h:18 */
    const int32_t* mesh_material_outside; /* src/mesh.h:19 */
    const int32_t* mesh_vascular;         /* src/mesh.h:15 */
    const float* mesh_deltas;             /* n_meshes x 3, src/mesh.h:16 */
    const int64_t* tri_offsets;           /* n_meshes + 1 */
    const float* tri_vertices;            /* 9 floats per triangle, OBJ space, objloader.h order */
    float scaling;                        /* "scaling" */
    float origin[3];                      /* "origin"  */
    float spacing[3];                     /* "spacing" */
} mcrt_scene_arrays;

typedef struct mcrt_info {
    int32_t rows;             /* RF samples per scanline: rfimage.h:180 (465 by default) */
    int32_t cols;             /* = elements */
    int32_t scan_rows, scan_cols;
    int32_t n_materials, n_meshes;
    int64_t n_triangles;
    int64_t n_bvh_nodes;
    int32_t device;
    int32_t sm_count;
    float start_pose[6];      /* the scene file's transducerPosition / transducerAngles */
    double axial_resolution_mm, time_step_us, row_period_us, max_travel_time_us;
    int32_t voxel_fma_division; /* 1: the 3-instruction voxel index passed its exhaustive check for this resolution and is in use */
    int32_t bvh_cache_hit;      /* 1: the acceleration structure (any builder) was loaded from $MCRT_BVH_CACHE instead of being built */
    int32_t bvh_optimised;      /* 1: the background-built SAH tree has replaced the device LBVH (options "bvh_optimise", "bvh_wait") */
    int32_t reserved0;
} mcrt_info;

typedef struct mcrt_stats {   /* of the most recent mcrt_simulate call */
    int64_t poses;
    int64_t segments;         /* closest-hit queries = the reference's `tests` counter (scene.cpp:118) */
    int64_t march_steps;      /* scatterer-volume samples (main.cpp:124-136) */
    int64_t kernel_launches;  /* kernels launched by this library inside the call */
    float ms_total;           /* device time of the call, CUDA events on the library's stream */
    float ms_trace, ms_accumulate, ms_post; /* only filled when profiling stages (mcrt_set_option) */
    int64_t bvh_node_visits;  /* only with option "count_traversal": BVH nodes fetched / triangles tested */
    int64_t bvh_triangle_tests;
    int64_t late_echoes;      /* windowed accumulate: echoes that arrived for an already finished row window (0 unless time runs backwards) */
} mcrt_stats;

int mcrt_default_params(mcrt_params* p);

/* replaces: main.cpp:52-81 (volume, psf, rf_image, json, transducer, scene construction).
 * Accepts the reference's scene files unchanged.  Extensions: missing "shininess"/"thickness"
 * default to 1e6 / 0; a non-existent "workingDirectory" falls back to the scene file's directory. */
int mcrt_create(const char* scene_json_path, const mcrt_params* params, int device, mcrt_ctx** out);
int mcrt_create_from_arrays(const mcrt_scene_arrays* scene, const mcrt_params* params, int device, mcrt_ctx** out);
void mcrt_destroy(mcrt_ctx* ctx);
const char* mcrt_last_error(void); /* thread-local, never NULL */

int mcrt_get_info(const mcrt_ctx* ctx, mcrt_info* info);
int mcrt_get_stats(const mcrt_ctx* ctx, mcrt_stats* stats);
/* options (name = value; unknown names fail with MCRT_ERR_INVALID):
 *   behaviour   "max_batch_poses"=N (poses per internal batch; sizes the workspace), "log_compress"=0/1 (apply the log compression the
 *               reference keeps commented out at rfimage.h:131-136 to the envelope image: affects rf_out and scan_out; default 0),
 *               "ray_tree"=B (> 0: follow BOTH children of every boundary hit with a budget of B segments per path, see
 *               mcrt_trace_tree_debug; 0 = off), "frame_stride"=G (pose i of a call is frame first_frame + i * G of the Philox stream:
 *               the round-robin pose deal of a G-rank sweep; default 1), "rf_out_frame_stride"=G (device rf_out only: frame i of a call
 *               is written at rf_out + i * G frames -- the interleaved slots of that deal in a buffer in global pose order; default 1),
 *               "bvh_builder"=0 device LBVH (default) / 1 host binned-SAH
 *               tree / 2 device PLOC (rebuilds the acceleration structure in place; every builder gives the same results),
 *               "bvh_optimise"=1/0 (default 1; builder 0, scenes of >= 32768 triangles: the device LBVH serves a new or changed scene at
 *               once while a host thread builds the binned-SAH tree of the same triangles; the first compute call after it has finished
 *               adopts it -- mcrt_info.bvh_optimised; after mesh updates the optimiser waits for 8 calls on an unchanged scene; the
 *               environment variable MCRT_BVH_OPTIMISE=0 disables it at mcrt_create), "bvh_wait"=1 (block until a running optimisation
 *               has finished and adopt its tree: keeps the swap out of a timed region)
 *   diagnostics "profile_stages"=0/1 (per-stage events in mcrt_stats, disables the CUDA graph), "count_traversal"=0/1 (BVH work counters
 *               in mcrt_stats), "use_graph"=0/1
 *   A/B switches of measured design choices (results are bit-identical either way; DESIGN.md section 5, profiles/):
 *               "tail_merge" (12, in units of 768 paths per SM; 0 = off), "first_hit_dedup" (1: large calls, 2: always, 0: off), "ordered_compaction" (1: large calls, 2: always,
 *               0: warp-aggregated atomic appends), "group_histories" (0), "accumulate_windowed" (1), "voxel_fma_division" (1 when the
 *               resolution passed its exhaustive check), "post_tma" (1), "long_ct" (1), "overlap" (0), "direct_out" (1: with a device rf_out the
 *               last kernel of the chain writes the frames straight into it; 0: internal image + device-to-device copy) */
int mcrt_set_option(mcrt_ctx* ctx, const char* name, int64_t value);

/* replaces one iteration of main.cpp:92-152 per pose: rf_image.clear(); scene.cast_rays();
 * echo accumulation; rf_image.convolve(psf); rf_image.envelope(); rf_image.postprocess().
 * Pose i is simulated as frame (first_frame + i) of the Philox stream `seed`.
 * rf_out:   n_poses x elements x rows float32 (layout per params.rf_layout), host OR device pointer.
 * scan_out: n_poses x scan_rows x scan_cols float32 (cv::remap of rfimage.h:139), host or device, nullable. */
int mcrt_simulate(mcrt_ctx* ctx, const mcrt_pose* poses, int32_t n_poses, uint64_t seed, uint64_t first_frame,
                  float* rf_out, float* scan_out);

/* Same, but rf_out/scan_out must be DEVICE pointers and nothing is synchronised: the work is
 * enqueued on `cuda_stream` (a cudaStream_t passed as void*; NULL = the library's own stream)
 * so a caller can overlap it or follow it with a collective.
 * Stream contract: a context owns ONE workspace (pose staging, path state, segments, RF images, captured graphs) that every
 * entry point uses.  Calls on one context must come from one host thread at a time; they may use different streams -- each
 * entry point makes its stream wait for the work the previous call enqueued (an event recorded at the end of every call), so
 * mixing mcrt_simulate_async on a user stream with the blocking entry points, or switching user streams, is ordered
 * correctly; what is NOT provided is concurrency between two calls on the same context (use one context per stream for that).
 * The results of an async call may be read once its stream has reached the point of the call. */
int mcrt_simulate_async(mcrt_ctx* ctx, const mcrt_pose* poses, int32_t n_poses, uint64_t seed, uint64_t first_frame,
                        float* rf_out_dev, float* scan_out_dev, void* cuda_stream);

/* One pose, only the scanline block [first_element, first_element + n_elements): the unit of the
 * scanline-block partition of a single frame across GPUs.  The block's right-hand PSF halo (psf_lateral - 1
 * scanlines) is re-traced locally, so the result equals the same scanlines of mcrt_simulate bit for bit.
 * rf_out: n_elements x rows float32, scanline-major, host or device pointer. */
int mcrt_simulate_scanlines(mcrt_ctx* ctx, const mcrt_pose* pose, uint64_t seed, uint64_t frame, int32_t first_element,
                            int32_t n_elements, float* rf_out);

/* parity hook for scene::cast_rays (scene.cpp:50-183): segments[elements][samples][max_depth],
 * n_segments[elements][samples]; host pointers. */
int mcrt_trace_debug(mcrt_ctx* ctx, const mcrt_pose* pose, uint64_t seed, uint64_t frame, mcrt_segment* segments,
                     int32_t* n_segments);

/* ---- peer-memory plumbing for the multi-GPU gather (one process per GPU): the rank that collects the RF lines allocates the
 * receive buffer with mcrt_device_alloc, exports it (CUDA IPC, 64-byte handle to ship through any host channel); every other
 * rank opens it and DEPOSITS its finished lines straight into that memory over NVLink -- either by passing the peer pointer
 * as rf_out_dev to mcrt_simulate_async, or with mcrt_copy_async from a local buffer on a copy stream (overlapping the next
 * step's simulation).  No SM is spent on the transfer and nobody but the collector receives anything. */
int mcrt_device_alloc(int device, size_t bytes, void** dev_ptr);
int mcrt_device_free(int device, void* dev_ptr);
int mcrt_ipc_export(int device, const void* dev_ptr, unsigned char handle64[64]);
int mcrt_ipc_open(int device, const unsigned char handle64[64], void** peer_ptr);
int mcrt_ipc_close(int device, void* peer_ptr);
/* dst / src: device pointers (local or peer-mapped); cuda_stream: a cudaStream_t passed as void* (NULL = default stream) */
int mcrt_copy_async(int device, void* dst, const void* src, size_t bytes, void* cuda_stream);
/* strided variant (cudaMemcpy2DAsync): `height` rows of `width_bytes`; used to deposit the frames of a round-robin pose
 * partition (rank r holds poses r, r + G, ...) into the collector's buffer in global pose order with ONE copy */
int mcrt_copy2d_async(int device, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t height,
                      void* cuda_stream);

/* Ray-tree mode (option "ray_tree" = segment budget per path > 0): BOTH children of every boundary hit
 * are followed, as in the cited paper, instead of the one Monte-Carlo branch this fork of the reference keeps
 * (ray.cpp:84-94).  Parity hook: all segments of one pose sorted by (path = element * samples + sample, node), node = 1 for
 * the root, 2n / 2n+1 for the reflected / refracted child of node n.  capacity = entries the three arrays can hold. */
int mcrt_trace_tree_debug(mcrt_ctx* ctx, const mcrt_pose* pose, uint64_t seed, uint64_t frame, int64_t capacity,
                          mcrt_segment* segments, int32_t* path, int32_t* node, int64_t* n_out);

/* ---- stage-level entry points (each one is the CUDA kernel of that stage; used by the parity
 * tests and by callers that want only part of the chain).  Host pointers unless noted. ---- */

/* btCollisionWorld::rayTest + ClosestRayResultCallback (scene.cpp:115-126) for n segments
 * [from,to]: tri_id (-1 miss), mesh_id, fraction, hit point and origin-facing unit normal. */
int mcrt_closest_hit(mcrt_ctx* ctx, int64_t n, const float* from3, const float* to3, int32_t* tri_id, int32_t* mesh_id,
                     float* fraction, float* point3, float* normal3);
/* transducer<N>::element(i) for a pose (transducer.h:24-67) */
int mcrt_transducer_elements(mcrt_ctx* ctx, const mcrt_pose* pose, float* pos3, float* dir3);
/* main.cpp:106-144 on caller-supplied segments -> raw (un-convolved) RF, scanline-major [elements][rows] */
int mcrt_accumulate(mcrt_ctx* ctx, const mcrt_segment* segments, const int32_t* n_segments, float* rf_out);
/* rf_image::convolve + rf_image::envelope (rfimage.h:93-123, 54-91) on a [cols][rows]
 * scanline-major image; flags bit0 = convolve, bit1 = envelope.  Any rows/cols/tap counts. */
int mcrt_postprocess(mcrt_ctx* ctx, const float* rf_in, int32_t cols, int32_t rows, const float* axial, int32_t n_axial,
                     const float* lateral, int32_t n_lateral, int32_t flags, float* rf_out);
/* rf_image::create_mapping + cv::remap (rfimage.h:183-215, 139) on the ctx's geometry */
int mcrt_scan_convert(mcrt_ctx* ctx, const float* rf_in /* [cols][rows] */, float* scan_out);
/* Moving / deforming meshes (the reference's `rigid` flag, mesh.h:15, anticipates non-rigid organs but
 * nothing ever moves, scene.cpp:318-333).  Updates are staged and applied by ONE acceleration-structure rebuild at the next
 * compute call (the device LBVH build is ~0.3 ms of kernels for 624 640 triangles, cheaper and better than a refit).
 * origin3: the body origin in world cm (= deltas * scaling, scene.cpp:320-324).  tri_local9: the mesh's triangles in its
 * local frame (v_obj * scaling), 9 floats each, same count and order as loaded (mcrt_get_scene). */
int mcrt_set_mesh_origin(mcrt_ctx* ctx, int32_t mesh, const float* origin3);
int mcrt_set_mesh_vertices(mcrt_ctx* ctx, int32_t mesh, const float* tri_local9, int64_t n_triangles);

/* Depth-dependent lateral PSF (psf.h:11-25 describes lateral ranges that "vary according to distance to the
 * transducer" while the reference fills one lateral kernel, psf.h:52-57): RF row r is convolved laterally with a Gaussian of
 * variance var_y * w^2, w = 1 + spread * |depth(r) - focus_cm| / focus_cm.  spread = 0 restores the reference PSF.
 * table_out (nullable): the taps, [psf_lateral][rows] floats. */
int mcrt_set_psf_depth_profile(mcrt_ctx* ctx, float focus_cm, float spread, float* table_out);

/* B-mode display chain on an envelope image (the reference stops at the envelope and keeps its
 * log compression commented out, rfimage.h:127-136, then writes the scan-converted image x255 as 8 bit, rfimage.h:142-148):
 *   v = |E| * 10^((gain_db + tgc_db_per_cm * depth_cm(row)) / 20),   y = clamp(1 + 20 log10(v / max v) / dynamic_range_db, 0, 1),
 * then create_mapping + cv::remap and x255 -> uint8.  env_in: n x [cols][rows] (rf_out of mcrt_simulate, rf_layout 0), host
 * or device.  compressed_out (nullable): n x [cols][rows] float in [0,1].  bmode8_out (nullable): n x scan_rows x scan_cols
 * uint8.  Both outputs host or device. */
typedef struct mcrt_bmode_params {
    float gain_db;           /* overall gain */
    float tgc_db_per_cm;     /* time-gain compensation slope; depth_cm(row) = row * depth_cm / rows */
    float dynamic_range_db;  /* > 0, e.g. 60 */
    float reserved0;
} mcrt_bmode_params;
int mcrt_bmode(mcrt_ctx* ctx, const float* env_in, int32_t n_images, const mcrt_bmode_params* bp, float* compressed_out,
               uint8_t* bmode8_out);

/* Elevational PSF.  psf.h:16-18 describes three ranges -- axial, lateral and elevation -- and psf.h:42,77
 * declares an elevation_kernel the reference never fills.  With n_planes > 1 (odd) every frame is traced as n_planes ray fans,
 * fan j offset by z_j = j * resolution - n_planes * resolution / 2 [mm] along the transducer's elevation axis (the normal of its
 * fan plane), Philox frame counter (frame * n_planes + j); the raw RF images of the fans are combined with the elevation taps
 * exp(-0.5 z_j^2 / var_z) (the form of psf.h:87-92) before the axial / lateral passes.  n_planes = 1 switches it off.
 * taps_out / z_mm_out (nullable): n_planes floats each.  mcrt_elevation_pose: the pose of fan `plane` (test hook). */
int mcrt_set_elevation(mcrt_ctx* ctx, int32_t n_planes, float var_z, float* taps_out, float* z_mm_out);
int mcrt_elevation_pose(const mcrt_ctx* ctx, const mcrt_pose* pose, int32_t plane, mcrt_pose* out);
int mcrt_get_psf_taps(const mcrt_ctx* ctx, float* axial, float* lateral);
/* scene as loaded (for loader parity): local-frame vertices (v_obj*scaling) 9 floats/triangle in
 * objloader order, mesh id per triangle, body origin per mesh (scene.cpp:313-324) */
int mcrt_get_scene(const mcrt_ctx* ctx, float* tri_local9, int32_t* tri_mesh, float* mesh_origin3, float* materials8);
/* scatterer volume as uploaded: 256^3 x {texture_noise, scattering_probability} (volume.h:19-35) */
int mcrt_get_volume(const mcrt_ctx* ctx, float* out);

/* ---- host-only entry points (no GPU needed): the drop-in surface's loaders and start-up tables ---- */

/* load_mesh_from_obj (objloader.h:154-161 = tinyobj::LoadObj + btgCreateGraphicsShapeFromWavefrontObj):
 * un-welded triangle soup, 9 floats per triangle in OBJ space.  Writes min(n, capacity) triangles. */
int mcrt_load_obj(const char* obj_path, float* out9, int64_t capacity_tris, int64_t* n_tris);
/* scene::parse_config + mesh loading (scene.cpp:185-247, 300-334) without creating a device context:
 * counts, and the same error codes / messages mcrt_create reports for a bad scene. */
int mcrt_scene_probe(const char* scene_json_path, int64_t* n_triangles, int32_t* n_meshes, int32_t* n_materials, float* start_pose6);
/* start-up tables for a parameter set: derived sizes (info: rows, cols, timing constants), the
 * (sin a_t, cos a_t) element table (transducer.h:41-59), PSF taps (psf.h:34-58) and the scan-conversion
 * maps (rfimage.h:183-215).  Any output pointer may be NULL. */
int mcrt_host_tables(const mcrt_params* params, mcrt_info* info, float* elem_sincos2, float* axial, float* lateral, float* map_x,
                     float* map_y);

/* the host tree builder of the background optimisation (option "bvh_optimise") and of "bvh_builder"=1: binned-SAH BVH2 over
 * n_triangles triangles (9 floats each, mesh-local; world = local + mesh_origin3[mesh]).  nodes16: (n_triangles - 1) x 16 words =
 * {child0 lo.xyz hi.xyz, child1 lo.xyz hi.xyz} as float + {child0, child1, 0, 0} as int32 (child >= 0: node index, pre-order, root 0;
 * child < 0: leaf slot -(1 + 4 * slot)); slot_triangle: the triangle of every leaf slot.  threads <= 0: all host threads; the
 * tree does not depend on it.  Replaces the Bullet btBvhTriangleMeshShape build of scene.cpp:309. */
int mcrt_host_build_sah(const float* tri_local9, const int32_t* tri_mesh, int64_t n_triangles, const float* mesh_origin3, int32_t n_meshes,
                        int32_t threads, float* nodes16, int32_t* slot_triangle, int32_t* max_depth);

/* numerics contract self-test: evaluates the shared transcendentals ON THE DEVICE.
 * op 0 expf, 1 logf, 2 powf(a,b), 3 sin (double), 4 cos (double), 5 philox (a=counter as float bits) */
int mcrt_numerics_probe(int device, int32_t op, int64_t n, const double* a, const double* b, double* out);

#ifdef __cplusplus
}
#endif
#endif /* MCRT_H */
