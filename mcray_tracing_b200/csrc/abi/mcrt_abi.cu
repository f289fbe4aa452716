// mcrt_abi.cu -- the C ABI of include/mcrt.h: context ownership, per-frame pipeline orchestration
// (upload poses -> wavefront trace -> accumulate -> PSF -> envelope -> scan conversion), CUDA-graph
// capture of that pipeline, and the stage-level entry points the parity tests call.
// No CPU fallback anywhere: every compute entry point launches the kernels of csrc/kernels/.
#include <algorithm>
#include <atomic>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/mcrt.h"
#include "../host/mcrt_host.h"
#include "../host/sah_builder.h"
#include "../kernels/mcrt_launch.h"

using namespace mcrt;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define CUDA_TRY_NOTHROW(expr) do { (void)(expr); (void)cudaGetLastError(); } while (0)
#define CUDA_TRY(expr)                                                                                            \
    do {                                                                                                          \
        cudaError_t e__ = (expr);                                                                                 \
        if (e__ != cudaSuccess) {                                                                                 \
            (void)cudaGetLastError();                                                                             \
            throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__));                                 \
        }                                                                                                         \
    } while (0)

// process-wide, per-device copy of the 128 MiB scatterer volume (volume.h: `static const volume_`)
std::mutex g_volume_mutex;
std::map<int, float2*> g_volume_dev;

float2* device_volume(int device, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_volume_mutex);
    auto it = g_volume_dev.find(device);
    if (it != g_volume_dev.end()) return it->second;
    const std::vector<float>& h = scatterer_volume();
    float2* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    g_volume_dev[device] = d;
    return d;
}

template <typename T>
void dev_alloc(T*& p, size_t n)
{
    p = nullptr;
    if (n == 0) return;
    CUDA_TRY(cudaMalloc(&p, n * sizeof(T)));
}

template <typename T>
void dev_free(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

bool is_device_pointer(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace

struct mcrt_ctx {
    int device = 0;
    int sm_count = 148;
    mcrt_params params;
    Derived dv;
    HostScene scene;
    AcqDev aq;
    SceneDev sc;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;        // second branch of the captured two-stream pipeline
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sub[4] = {nullptr, nullptr, nullptr, nullptr};
    bool overlap = false;                  // software-pipeline sub-batches inside the graph (measured slower: off)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_up = nullptr;
    bool upload_pending = false;
    // cross-stream ordering of the shared workspace: every entry point records ev_last on the stream it used; an entry point that
    // runs on a DIFFERENT stream first makes that stream wait for it (begin_call / end_call)
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool last_valid = false;

    DevMesh* d_meshes = nullptr;
    DevMaterial* d_materials = nullptr;
    LbvhResult bvh{};
    float2* d_volume = nullptr;        // owned by the process-wide cache
    float2* d_elem_sincos = nullptr;
    bool voxel_fma_validated = false;
    Bvh4Node* d_nodes4 = nullptr;      // 4-wide copy of bvh.nodes (collapse_bvh4; only when built with MCRT_BVH8=0)
    Bvh8Node* d_nodes8 = nullptr;      // 8-wide tree (collapse_bvh8), the default traversal structure
    TriSlot* d_tris8 = nullptr;        // triangle slots in the 8-wide tree's leaf order
    int n_nodes_wide = 0, depth_wide = 0;
    int tree_budget = 0;               // > 0: ray-tree mode with this many segments per path (option "ray_tree")
    TreeBuffers tree{};
    int* h_tree = nullptr;             // pinned: per batch {segments, overflow}
    int bvh_builder = 0;               // 0 device LBVH, 1 host binned SAH (option "bvh_builder")
    bool bvh_cache_hit = false;        // the last SAH build came from $MCRT_BVH_CACHE
    bool scene_dirty = false;          // staged mesh updates wait for a rebuild
    // Background tree optimisation (option "bvh_optimise", default on, builder 0 only): the device LBVH serves a new or changed
    // scene at once; a host thread builds the binned-SAH tree of the same triangles and the first compute call after it has
    // finished swaps it in (results do not depend on the tree: closest hits tie-break by triangle id).
    struct BgTree {
        std::thread th;
        std::atomic<int> state{0};         // 0 idle, 1 building, 2 ready, 3 failed
        std::atomic<bool> cancel{false};   // abandon the running build (mcrt_destroy, builder / option changes)
        uint64_t version = 0;              // scene_version the snapshot below was taken at
        std::vector<float> tri_local, origins;
        std::vector<int32_t> tri_mesh;
        HostBvh tree;
        ~BgTree() { cancel.store(true); if (th.joinable()) th.join(); }
    };
    std::unique_ptr<BgTree> bg;
    uint64_t scene_version = 0;        // bumped by every staged mesh update
    int bvh_optimise = 1;
    bool bvh_optimised = false;        // the traversal tree in use is the optimised one
    int quiet_calls = 0;               // compute calls since the last mesh update (a deforming scene is not re-optimised every frame)
    float* d_axial = nullptr;
    float* d_lateral = nullptr;
    float* d_lat_by_row = nullptr;     // depth-dependent lateral PSF table [psf_lateral][rows] (mcrt_set_psf_depth_profile), or nullptr
    std::vector<float> h_lat_by_row;
    float* d_map_x = nullptr;
    float* d_map_y = nullptr;
    std::vector<float> h_axial, h_lateral;

    // per-batch workspace
    int cap_poses = 0;
    TraceBuffers tb{};
    PoseTrigDev* d_poses = nullptr;
    unsigned long long* d_seed_frame = nullptr;
    unsigned long long* d_steps = nullptr;
    unsigned long long* d_trav = nullptr;      // {node visits, triangle tests}, only when count_traversal
    bool count_traversal = false;
    bool log_compress = false;             // rfimage.h:127-136, commented out in the reference
    int* d_max_bits = nullptr;             // [cap_poses] per-image maximum (ordered-int encoding)
    // mcrt_bmode workspace (grow-only: a cudaMalloc / cudaFree pair of ~0.75 GB per call cost more than the kernels)
    float *bm_gain = nullptr, *bm_env = nullptr, *bm_cmp = nullptr, *bm_scan = nullptr;
    unsigned char* bm_q = nullptr;
    int* bm_max = nullptr;
    int bm_cap = 0;
    // elevational PSF (mcrt_set_elevation): every output frame is traced as elev_n ray fans offset along the elevation axis
    int elev_n = 1;
    std::vector<float> h_elev_w, h_elev_z;
    float* d_elev_w = nullptr;
    float* d_rf_elev = nullptr;            // raw RF images of the output frames after the elevational combine ([frames][E][rf_pitch])
    int frame_stride = 1;                  // frame index of pose i of a call = first_frame + i * frame_stride (option "frame_stride")
    bool direct_out = true;                // option "direct_out": the post kernel writes the frames straight into the caller's device buffer
    bool cur_direct = false;               // ... decided per call (simulate_impl)
    int out_frame_stride = 1;              // option "rf_out_frame_stride": frame i of a call lands at rf_out + i * stride frames
    bool post_tma = true;                  // TMA-staged fused post kernel (option "post_tma"; 0 = round 1's k_post_fused, for A/B and equivalence tests)
    int first_hit_dedup = 1;               // bounce 0 traced once per element (TraceBuffers::first_hits) when samples >= 4: 0 off, 1 large calls, 2 always
    int ordered_compaction = 1;            // order-preserving compaction between bounces (TraceBuffers::warp_counts): 0 off, 1 large calls, 2 always
    bool group_histories = false;          // ... with the refracted survivors before the reflected ones (option "group_histories"; measured slower: profiles/r02ab_ab_group_histories.txt)
    float* d_rf_acc = nullptr;
    float* d_rf_tmp0 = nullptr;
    float* d_rf_tmp1 = nullptr;
    float* d_rf_final = nullptr;
    float* d_rf_t = nullptr;           // transposed copy when rf_layout == 1
    float* d_scan = nullptr;
    float* d_columns = nullptr;        // HBM accumulate columns (long scanlines only)
    size_t columns_bytes = 0;
    PoseTrig* h_poses = nullptr;       // pinned staging
    unsigned long long* h_seed_frame = nullptr;
    int* h_counters = nullptr;         // pinned, [MAX_BATCHES][32]
    unsigned long long* h_steps = nullptr;
    unsigned long long* h_trav = nullptr;

    std::map<std::pair<int, int>, cudaGraphExec_t> graphs;   // (n_poses in batch, want_scan) -> exec
    std::map<std::pair<int, int>, int> graph_launches;        // kernels in that graph, counted while capturing
    bool use_graph = true;
    bool profile_stages = false;
    int max_batch_poses = 256;
    mcrt_stats stats{};
    bool stats_pending = false;
    int pending_batches = 0;
    int pending_launches = 0;
};

extern "C" {   // defined next to the entry points below
static void update_scene_bounds(mcrt_ctx* c);
static void ensure_scene_current(mcrt_ctx* c);
static void rebuild_bvh(mcrt_ctx* c);
static void make_bvh2(mcrt_ctx* c, mcrt::LbvhResult* nb);
static void bg_tree_start(mcrt_ctx* c);
static void bg_tree_adopt(mcrt_ctx* c, bool wait);
static void bg_tree_cancel(mcrt_ctx* c);
}

namespace {

const int kMaxBatchesPerCall = 4096;
const int64_t kOrderedMinPaths = 262144;      // paths per call from which the order-preserving compaction is used
const int64_t kFirstHitMinElements = 4096;   // (pose, element) pairs per call from which bounce 0 is traced once per element
const int kMaxSub = 4;               // sub-batches of the two-stream pipeline
const int kMinPosesPerSub = 8;
const int kCounterSlot = 128;        // ints reserved per batch in the pinned counters mirror

void free_workspace(mcrt_ctx* c)
{
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
    dev_free(c->tb.paths.origin_intensity); dev_free(c->tb.paths.dir_state); dev_free(c->tb.paths.distance);
    dev_free(c->tb.segments); dev_free(c->tb.n_segments); dev_free(c->tb.hit_fraction); dev_free(c->tb.hit_mesh);
    dev_free(c->tb.queue_a); dev_free(c->tb.queue_b); dev_free(c->tb.counters);
    dev_free(c->tb.warp_counts); dev_free(c->tb.tile_counts); c->tb.n_tiles = 0; dev_free(c->tb.first_hits);
    dev_free(c->tree.rays_a); dev_free(c->tree.rays_b); dev_free(c->tree.queue_a); dev_free(c->tree.queue_b); dev_free(c->tree.warp_counts);
    dev_free(c->tree.tile_counts); dev_free(c->tree.segments); dev_free(c->tree.keys); dev_free(c->tree.level_first); dev_free(c->tree.level_end);
    dev_free(c->tree.counters);
    c->tree = TreeBuffers{};
    dev_free(c->d_poses); dev_free(c->d_rf_acc); dev_free(c->d_rf_tmp0); dev_free(c->d_rf_tmp1); dev_free(c->d_rf_final);
    dev_free(c->d_rf_t); dev_free(c->d_scan); dev_free(c->d_columns); dev_free(c->d_max_bits); dev_free(c->d_rf_elev);
    if (c->h_poses) cudaFreeHost(c->h_poses);
    c->h_poses = nullptr;
    c->cap_poses = 0;
}

void ensure_workspace(mcrt_ctx* c, int n_poses)
{
    if (n_poses <= c->cap_poses) return;
    free_workspace(c);
    const size_t n_paths = (size_t)n_poses * c->aq.elements * c->aq.samples;
    const size_t n_px = (size_t)n_poses * c->aq.elements * c->aq.rows;
    const size_t n_px_acc = (size_t)n_poses * c->aq.elements * c->aq.rf_pitch;      // raw image: 16-byte row pitch (AcqDev::rf_pitch)
    if (n_paths * c->aq.max_depth > 0x7fffffffULL) throw std::invalid_argument("batch too large: reduce max_batch_poses");
    dev_alloc(c->tb.paths.origin_intensity, n_paths);
    dev_alloc(c->tb.paths.dir_state, n_paths);
    dev_alloc(c->tb.paths.distance, n_paths);
    dev_alloc(c->tb.segments, n_paths * c->aq.max_depth);
    dev_alloc(c->tb.n_segments, n_paths);
    dev_alloc(c->tb.queue_a, n_paths);
    dev_alloc(c->tb.queue_b, n_paths);
    dev_alloc(c->tb.counters, (size_t)kMaxSub * (c->aq.max_depth + 1));
    dev_alloc(c->d_poses, (size_t)n_poses);
    if (c->elev_n > 1) dev_alloc(c->d_rf_elev, n_px_acc / (size_t)c->elev_n + 4);
    dev_alloc(c->d_rf_acc, n_px_acc);
    CUDA_TRY(cudaMemset(c->d_rf_acc, 0, sizeof(float) * n_px_acc));                 // the pad words of every row stay 0
    dev_alloc(c->d_rf_tmp0, n_px);
    dev_alloc(c->d_rf_tmp1, n_px);
    dev_alloc(c->d_rf_final, n_px);
    if (c->params.rf_layout == 1) dev_alloc(c->d_rf_t, n_px);
    dev_alloc(c->d_scan, (size_t)n_poses * c->params.scan_rows * c->params.scan_cols);
    dev_alloc(c->d_max_bits, (size_t)n_poses);
    // the windowed accumulate kernel keeps its columns in shared memory: no HBM columns at all
    c->columns_bytes = (c->aq.accumulate_windowed || c->tree_budget > 0) ? 0 : accumulate_columns_bytes(c->aq, n_poses);
    if (c->tree_budget > 0) {
        const size_t cap = n_paths * (size_t)c->tree_budget;
        if (cap > 0x3fffffffULL) throw std::invalid_argument("ray_tree: batch too large for the segment pool; reduce max_batch_poses or the budget");
        TreeBuffers& t = c->tree;
        t.seg_capacity = (int)cap; t.ray_capacity = (int)(cap / 2 + n_paths);
        t.n_scanlines = n_poses * c->aq.elements;
        const size_t n_chunks = (size_t)t.ray_capacity / 32 + 1;
        t.n_tiles = (int)((n_chunks + 255) / 256);
        dev_alloc(t.rays_a, 64 * n_chunks); dev_alloc(t.rays_b, 64 * n_chunks);
        dev_alloc(t.queue_a, (size_t)t.ray_capacity); dev_alloc(t.queue_b, (size_t)t.ray_capacity);
        dev_alloc(t.warp_counts, n_chunks); dev_alloc(t.tile_counts, (size_t)c->aq.max_depth * t.n_tiles);
        dev_alloc(t.segments, cap); dev_alloc(t.keys, cap);
        dev_alloc(t.level_first, (size_t)c->aq.max_depth * t.n_scanlines); dev_alloc(t.level_end, (size_t)c->aq.max_depth * t.n_scanlines);
        dev_alloc(t.counters, (size_t)c->aq.max_depth + 3);
    }
    if (c->columns_bytes) dev_alloc(c->d_columns, c->columns_bytes / sizeof(float));
    CUDA_TRY(cudaMallocHost(&c->h_poses, sizeof(PoseTrig) * (size_t)n_poses));
    c->tb.trav_counters = c->count_traversal ? c->d_trav : nullptr;
    if (c->first_hit_dedup && c->aq.samples >= 4) dev_alloc(c->tb.first_hits, 2 * (size_t)n_poses * c->aq.elements);
    if (c->ordered_compaction) {
        const size_t n_chunks = (n_paths + 31) / 32;
        c->tb.n_tiles = (int)((n_chunks + 255) / 256);
        dev_alloc(c->tb.warp_counts, n_chunks + 1);
        dev_alloc(c->tb.tile_counts, 2 * (size_t)c->aq.max_depth * c->tb.n_tiles);
        c->tb.group_histories = c->group_histories ? 1 : 0;
    }
    c->cap_poses = n_poses;
}

// One workspace (d_poses, tb.*, d_rf_*, the pinned pose / seed staging, the captured graphs) serves every entry point.  Work of a
// previous call may still be in flight on another stream (mcrt_simulate_async returns without synchronising): the stream of
// this call waits for it, and the pinned staging is not rewritten before its last upload has been consumed.
void begin_call(mcrt_ctx* c, cudaStream_t s)
{
    if (c->last_valid && c->last_stream != s) CUDA_TRY(cudaStreamWaitEvent(s, c->ev_last, 0));
    if (c->upload_pending) { CUDA_TRY(cudaEventSynchronize(c->ev_up)); c->upload_pending = false; }
}
void end_call(mcrt_ctx* c, cudaStream_t s)
{
    CUDA_TRY(cudaEventRecord(c->ev_last, s));
    c->last_stream = s; c->last_valid = true;
}

// the wide traversal tree over a freshly built BVH2 (device LBVH or uploaded host SAH tree)
struct WideTree { Bvh4Node* n4 = nullptr; Bvh8Node* n8 = nullptr; TriSlot* t8 = nullptr; int n_nodes = 0, depth = 0; };
void build_wide_tree(const LbvhResult& b, cudaStream_t stream, WideTree* w)
{
#if MCRT_BVH8
    const cudaError_t ce = collapse_bvh8(b.nodes, b.n_nodes, b.tris, b.n_tri, &w->n8, &w->t8, stream, &w->depth, &w->n_nodes);
    if (ce != cudaSuccess) throw CudaError(std::string("collapse_bvh8: ") + cudaGetErrorString(ce));
    if (w->depth > MCRT_STACK_DEPTH8) {
        cudaFree(w->n8); cudaFree(w->t8);
        throw std::invalid_argument("8-wide BVH deeper than the traversal stack (degenerate mesh?)");
    }
#else
    const cudaError_t ce = collapse_bvh4(b.nodes, b.n_nodes, &w->n4, stream, &w->depth);
    if (ce != cudaSuccess) throw CudaError(std::string("collapse_bvh4: ") + cudaGetErrorString(ce));
    if (3 * w->depth + 1 > MCRT_STACK_DEPTH4) { cudaFree(w->n4); throw std::invalid_argument("4-wide BVH deeper than the traversal stack (degenerate mesh?)"); }
    w->n_nodes = b.n_nodes;
#endif
}

// wavefront trace of poses [pose0, pose0 + n) of the uploaded batch; `slot` selects its compaction counters
void enqueue_trace(mcrt_ctx* c, int pose0, int n, int slot, cudaStream_t s, int* launches)
{
    const size_t p0 = (size_t)pose0 * c->aq.elements * c->aq.samples;
    FrameDev fr;
    fr.poses = c->d_poses + pose0; fr.elem_sincos = c->d_elem_sincos; fr.seed_frame = c->d_seed_frame; fr.n_poses = n; fr.frame_offset = pose0; fr.frame_stride = c->frame_stride;
    TraceBuffers tb = c->tb;
    tb.paths.origin_intensity += p0; tb.paths.dir_state += p0; tb.paths.distance += p0;
    tb.segments += p0 * c->aq.max_depth; tb.n_segments += p0;
    if (tb.hit_fraction) tb.hit_fraction += p0 * c->aq.max_depth;
    if (tb.hit_mesh) tb.hit_mesh += p0 * c->aq.max_depth;
    tb.queue_a += p0; tb.queue_b += p0;
    // measured (profiles/r01u_ab_firsthit.txt): -3.6 % trace time at 256 frames per call, +2 % at one frame (one more dependent launch)
    if (tb.first_hits && c->first_hit_dedup == 1 && (int64_t)n * c->aq.elements < kFirstHitMinElements) tb.first_hits = nullptr;
    if (tb.first_hits) tb.first_hits += 2 * (size_t)pose0 * c->aq.elements;
    // sub-batches keep the atomic compaction; small calls too (the 9 extra scan launches cost more than the order returns:
    // +1.6 % frames/s at 256 frames per call, profiles/r01o_ab_ordered.txt)
    if (pose0 != 0 || slot != 0 || (c->ordered_compaction == 1 && (int64_t)n * c->aq.elements * c->aq.samples < kOrderedMinPaths)) {
        tb.warp_counts = nullptr; tb.tile_counts = nullptr; tb.n_tiles = 0;
    }
    tb.counters += (size_t)slot * (c->aq.max_depth + 1);
    launch_trace(c->sc, c->aq, fr, tb, c->sm_count, s, launches);
}

// accumulate -> PSF -> envelope (-> transpose, scan conversion) of poses [pose0, pose0 + n)
// `n` counts SUB-frames (elev_n ray fans per output frame; 1 without the elevational PSF), pose0 likewise.
void enqueue_image(mcrt_ctx* c, int pose0, int n, bool want_scan, cudaStream_t s, int* launches, cudaEvent_t ev_after_accumulate = nullptr)
{
    const int N = c->elev_n;
    const size_t p0 = (size_t)pose0 * c->aq.elements * c->aq.samples;
    const size_t ax0 = (size_t)pose0 * c->aq.elements * c->aq.rf_pitch;
    CUDA_TRY(launch_accumulate(c->sc, c->aq, c->d_volume, c->tb.segments + p0 * c->aq.max_depth, c->tb.n_segments + p0, n, c->d_rf_acc + ax0,
                               c->d_steps, c->d_columns + p0 * c->aq.rows, s, launches));
    // output frames of this piece
    const int f0 = pose0 / N, nf = n / N;
    const size_t px0 = (size_t)f0 * c->aq.elements * c->aq.rows;
    const float* raw = c->d_rf_acc + ax0;
    if (N > 1) {
        const int64_t px_img = (int64_t)c->aq.elements * c->aq.rf_pitch;
        float* comb = c->d_rf_elev + (size_t)f0 * px_img;
        launch_elevation_combine(raw, nf, px_img, N, c->d_elev_w, comb, s, launches);
        raw = comb;
    }
    if (ev_after_accumulate) CUDA_TRY(cudaEventRecord(ev_after_accumulate, s));
    launch_post(raw, nf, c->aq.elements, c->aq.rows, c->d_axial, c->params.psf_axial, c->d_lateral, c->params.psf_lateral, 3,
                c->d_rf_tmp0 + px0, c->d_rf_tmp1 + px0, c->d_rf_final + px0, s, launches, 0, 0, c->d_lat_by_row, c->aq.rf_pitch,
                c->post_tma ? c->h_axial.data() : nullptr, c->h_lateral.data(), c->cur_direct ? c->d_seed_frame + 2 : nullptr);
    if (c->log_compress) launch_log_compress(c->d_rf_final + px0, nf, (int64_t)c->aq.elements * c->aq.rows, c->d_max_bits + f0, s, launches);
    if (c->params.rf_layout == 1) launch_transpose(c->d_rf_final + px0, nf, c->aq.elements, c->aq.rows, c->d_rf_t + px0, s, launches);
    if (want_scan)
        launch_scan_convert(c->d_rf_final + px0, nf, c->aq.elements, c->aq.rows, c->d_map_x, c->d_map_y, c->params.scan_rows, c->params.scan_cols,
                            c->d_scan + (size_t)f0 * c->params.scan_rows * c->params.scan_cols, s, launches);
}

// enqueue the whole per-frame chain for `n` (sub-)poses already uploaded to d_poses / d_seed_frame, one stream
void enqueue_pipeline(mcrt_ctx* c, int n, bool want_scan, cudaStream_t s, int* launches, bool stage_events)
{
    CUDA_TRY(cudaMemsetAsync(c->tb.counters, 0, sizeof(int) * (size_t)kMaxSub * (c->aq.max_depth + 1), s));
    CUDA_TRY(cudaMemsetAsync(c->d_steps, 0, 2 * sizeof(unsigned long long), s));
    if (stage_events) CUDA_TRY(cudaEventRecord(c->ev_a, s));
    enqueue_trace(c, 0, n, 0, s, launches);
    if (stage_events) CUDA_TRY(cudaEventRecord(c->ev_b, s));
    // stage split for the profile: accumulate (+ sample reduction, + elevational combine) | PSF + envelope (+ transpose, scan)
    enqueue_image(c, 0, n, want_scan, s, launches, stage_events ? c->ev_c : nullptr);
    CUDA_TRY(cudaGetLastError());
}

// The same chain software-pipelined over `nsub` pose sub-batches on two streams (only ever captured into a
// CUDA graph): stream s1 traces sub-batch k+1 while stream s2 accumulates / post-processes sub-batch k.
// The trace kernels are latency-bound and the accumulate kernel issue-bound, so they fill each other's
// idle issue slots.  Results are bit-identical to the single-stream chain (every write is keyed by path).
void enqueue_pipeline_overlapped(mcrt_ctx* c, int n, int nsub, bool want_scan, cudaStream_t s1, cudaStream_t s2, int* launches)
{
    CUDA_TRY(cudaMemsetAsync(c->tb.counters, 0, sizeof(int) * (size_t)kMaxSub * (c->aq.max_depth + 1), s1));
    CUDA_TRY(cudaMemsetAsync(c->d_steps, 0, 2 * sizeof(unsigned long long), s1));
    CUDA_TRY(cudaEventRecord(c->ev_fork, s1));
    CUDA_TRY(cudaStreamWaitEvent(s2, c->ev_fork, 0));
    for (int k = 0; k < nsub; k++) {
        const int pose0 = (int)((long long)n * k / nsub), pose1 = (int)((long long)n * (k + 1) / nsub);
        enqueue_trace(c, pose0, pose1 - pose0, k, s1, launches);
        CUDA_TRY(cudaEventRecord(c->ev_sub[k], s1));
        CUDA_TRY(cudaStreamWaitEvent(s2, c->ev_sub[k], 0));
        enqueue_image(c, pose0, pose1 - pose0, want_scan, s2, launches);
    }
    CUDA_TRY(cudaEventRecord(c->ev_join, s2));
    CUDA_TRY(cudaStreamWaitEvent(s1, c->ev_join, 0));
    CUDA_TRY(cudaGetLastError());
}

int pipeline_sub_batches(const mcrt_ctx* c, int n)
{
    if (!c->overlap || n < 2 * kMinPosesPerSub || c->elev_n > 1) return 1;
    int nsub = n / kMinPosesPerSub;
    return nsub > kMaxSub ? kMaxSub : nsub;
}

void run_batch(mcrt_ctx* c, int n, bool want_scan, cudaStream_t s, int* launches)
{
    const bool graph_ok = c->use_graph && !c->profile_stages;
    if (!graph_ok) {
        enqueue_pipeline(c, n, want_scan, s, launches, c->profile_stages);
        return;
    }
    const auto key = std::make_pair(n, (want_scan ? 1 : 0) | (c->cur_direct ? 2 : 0));
    const int nsub = pipeline_sub_batches(c, n);
    auto it = c->graphs.find(key);
    if (it == c->graphs.end()) {
        // capture on the library's own stream (the caller's stream may be the legacy default stream)
        cudaGraph_t graph = nullptr;
        int dummy = 0;
        CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        try {
            if (nsub > 1) enqueue_pipeline_overlapped(c, n, nsub, want_scan, c->stream, c->stream2, &dummy);
            else enqueue_pipeline(c, n, want_scan, c->stream, &dummy, false);
        } catch (...) {
            cudaStreamEndCapture(c->stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        CUDA_TRY(cudaStreamEndCapture(c->stream, &graph));
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) throw CudaError(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        it = c->graphs.emplace(key, exec).first;
        c->graph_launches[key] = dummy;
    }
    CUDA_TRY(cudaGraphLaunch(it->second, s));
    if (launches) *launches += c->graph_launches[key];
}

// ray-tree mode: trace the trees level by level (ordered, nothing is sorted), accumulate a lane per segment, then the usual
// PSF / envelope / scan chain.  Not graph-captured (rarely used; the launch sequence itself is static).
void run_tree_batch(mcrt_ctx* c, int n, bool want_scan, cudaStream_t s, int* launches)
{
    CUDA_TRY(cudaMemsetAsync(c->d_steps, 0, 2 * sizeof(unsigned long long), s));
    CUDA_TRY(cudaMemsetAsync(c->tb.counters, 0, sizeof(int) * (size_t)kMaxSub * (c->aq.max_depth + 1), s));
    FrameDev fr;
    fr.poses = c->d_poses; fr.elem_sincos = c->d_elem_sincos; fr.seed_frame = c->d_seed_frame; fr.n_poses = n; fr.frame_offset = 0; fr.frame_stride = c->frame_stride;
    TreeBuffers t = c->tree;
    t.trav_counters = c->count_traversal ? c->d_trav : nullptr;
    if (c->profile_stages) CUDA_TRY(cudaEventRecord(c->ev_a, s));
    launch_trace_tree(c->sc, c->aq, fr, t, c->sm_count, s, launches);
    if (c->profile_stages) CUDA_TRY(cudaEventRecord(c->ev_b, s));
    CUDA_TRY(launch_accumulate_tree(c->sc, c->aq, c->d_volume, t, n, c->d_rf_acc, c->d_steps, s, launches));
    if (c->profile_stages) CUDA_TRY(cudaEventRecord(c->ev_c, s));
    launch_post(c->d_rf_acc, n, c->aq.elements, c->aq.rows, c->d_axial, c->params.psf_axial, c->d_lateral, c->params.psf_lateral, 3,
                c->d_rf_tmp0, c->d_rf_tmp1, c->d_rf_final, s, launches, 0, 0, c->d_lat_by_row, c->aq.rf_pitch, c->post_tma ? c->h_axial.data() : nullptr,
                c->h_lateral.data());
    if (c->log_compress) launch_log_compress(c->d_rf_final, n, (int64_t)c->aq.elements * c->aq.rows, c->d_max_bits, s, launches);
    if (c->params.rf_layout == 1) launch_transpose(c->d_rf_final, n, c->aq.elements, c->aq.rows, c->d_rf_t, s, launches);
    if (want_scan)
        launch_scan_convert(c->d_rf_final, n, c->aq.elements, c->aq.rows, c->d_map_x, c->d_map_y, c->params.scan_rows, c->params.scan_cols,
                            c->d_scan, s, launches);
    CUDA_TRY(cudaGetLastError());
}

int simulate_impl(mcrt_ctx* c, const mcrt_pose* poses, int32_t n_poses, uint64_t seed, uint64_t first_frame, float* rf_out,
                  float* scan_out, cudaStream_t user_stream, bool async)
{
    if (!c || !poses || n_poses < 0 || !rf_out) return fail(MCRT_ERR_INVALID, "mcrt_simulate: null argument");
    if (n_poses == 0) return MCRT_OK;
    try {
        CUDA_TRY(cudaSetDevice(c->device));
        const bool rf_dev = async ? true : is_device_pointer(rf_out);
        const bool scan_dev = scan_out ? (async ? true : is_device_pointer(scan_out)) : true;
        cudaStream_t s = (async && user_stream) ? user_stream : c->stream;
        // elevational PSF: every output frame = N ray fans (sub-frames); batches hold whole frames
        const int N = c->elev_n;
        if (N > 1 && c->frame_stride != 1) return fail(MCRT_ERR_INVALID, "mcrt_simulate: the elevational PSF cannot be combined with frame_stride != 1");
        if (N > 1 && c->tree_budget > 0) return fail(MCRT_ERR_INVALID, "mcrt_simulate: the elevational PSF is not available in ray-tree mode");
        int batch_cap = c->max_batch_poses < 1 ? 1 : c->max_batch_poses;          // in sub-frames
        batch_cap = batch_cap / N < 1 ? N : (batch_cap / N) * N;
        const int frames_per_batch = batch_cap / N;
        const int n_batches = (n_poses + frames_per_batch - 1) / frames_per_batch;
        if (n_batches > kMaxBatchesPerCall) return fail(MCRT_ERR_INVALID, "mcrt_simulate: too many batches; raise max_batch_poses");
        ensure_scene_current(c);
        ensure_workspace(c, (n_poses < frames_per_batch ? n_poses : frames_per_batch) * N);
        const size_t px_per_pose = (size_t)c->aq.elements * c->aq.rows;
        const size_t scan_per_pose = (size_t)c->params.scan_rows * c->params.scan_cols;
        int launches = 0;
        // frame i of the call lands at rf_out + i * ostride frames (option "rf_out_frame_stride": the interleaved slots of a round-robin sweep)
        const int ostride = c->out_frame_stride;
        if (ostride != 1 && !rf_dev) return fail(MCRT_ERR_INVALID, "mcrt_simulate: rf_out_frame_stride needs a device rf_out");
        // the post kernel can write the frames straight into rf_out (no internal image + device-to-device copy) when the chain ends there
        const int first_nf = n_poses < frames_per_batch ? n_poses : frames_per_batch;
        c->cur_direct = c->direct_out && rf_dev && c->params.rf_layout == 0 && !c->log_compress && !scan_out && c->tree_budget == 0 &&
                        pipeline_sub_batches(c, first_nf * N) == 1 &&
                        post_writes_through_target(first_nf, c->aq.elements, c->aq.rows, c->aq.rf_pitch, c->params.psf_axial, c->params.psf_lateral, 3,
                                                   c->post_tma ? c->h_axial.data() : nullptr, c->h_lateral.data(), c->d_lat_by_row != nullptr);
        begin_call(c, s);
        CUDA_TRY(cudaEventRecord(c->ev0, s));
        if (c->count_traversal) CUDA_TRY(cudaMemsetAsync(c->d_trav, 0, 2 * sizeof(unsigned long long), s));
        for (int b = 0; b < n_batches; b++) {
            const int p0 = b * frames_per_batch;                                     // first output frame of the batch
            const int nf = (n_poses - p0) < frames_per_batch ? (n_poses - p0) : frames_per_batch;
            const int n = nf * N;                                                    // sub-frames
            // the pinned pose staging is reused: wait for the previous upload (not the previous frame) to finish
            if (c->upload_pending) { CUDA_TRY(cudaEventSynchronize(c->ev_up)); c->upload_pending = false; }
            if (N == 1) {
                for (int i = 0; i < n; i++) c->h_poses[i] = pose_trig(poses[p0 + i]);
            } else {
                for (int i = 0; i < nf; i++)
                    for (int j = 0; j < N; j++) c->h_poses[i * N + j] = pose_trig(elevation_pose(poses[p0 + i], c->h_elev_z[j]));
            }
            c->h_seed_frame[0] = seed;
            // frame (Philox counter) of sub-frame (i, j) = (first_frame + i) * N + j
            c->h_seed_frame[1] = N == 1 ? first_frame + (uint64_t)p0 * (uint64_t)c->frame_stride : (first_frame + (uint64_t)p0) * (uint64_t)N;
            // where the post kernel writes this batch's frames when it writes through the output target (cur_direct)
            float* const rf_dst = rf_out + (size_t)p0 * px_per_pose * (size_t)ostride;
            c->h_seed_frame[2] = (unsigned long long)reinterpret_cast<uintptr_t>(rf_dst);
            c->h_seed_frame[3] = (unsigned long long)(px_per_pose * (size_t)ostride);
            CUDA_TRY(cudaMemcpyAsync(c->d_poses, c->h_poses, sizeof(PoseTrig) * (size_t)n, cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(c->d_seed_frame, c->h_seed_frame, 4 * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaEventRecord(c->ev_up, s));
            c->upload_pending = true;
            if (c->tree_budget > 0) {
                run_tree_batch(c, n, scan_out != nullptr, s, &launches);
                CUDA_TRY(cudaMemcpyAsync(c->h_tree + 2 * b, c->tree.counters + c->aq.max_depth + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            } else {
                run_batch(c, n, scan_out != nullptr, s, &launches);
            }
            if (!c->cur_direct) {
                const float* rf_src = c->params.rf_layout == 1 ? c->d_rf_t : c->d_rf_final;
                if (ostride == 1)
                    CUDA_TRY(cudaMemcpyAsync(rf_dst, rf_src, sizeof(float) * px_per_pose * nf, rf_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
                else
                    CUDA_TRY(cudaMemcpy2DAsync(rf_dst, sizeof(float) * px_per_pose * (size_t)ostride, rf_src, sizeof(float) * px_per_pose,
                                               sizeof(float) * px_per_pose, (size_t)nf, rf_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
            }
            if (scan_out)
                CUDA_TRY(cudaMemcpyAsync(scan_out + (size_t)p0 * scan_per_pose, c->d_scan, sizeof(float) * scan_per_pose * nf,
                                         scan_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(c->h_counters + kCounterSlot * b, c->tb.counters, sizeof(int) * (size_t)kMaxSub * (c->aq.max_depth + 1),
                                     cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(c->h_steps + 2 * b, c->d_steps, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        }
        if (c->count_traversal) CUDA_TRY(cudaMemcpyAsync(c->h_trav, c->d_trav, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaEventRecord(c->ev1, s));
        end_call(c, s);
        c->stats_pending = true;
        c->pending_batches = n_batches;
        c->pending_launches = launches;
        c->stats = mcrt_stats{};
        c->stats.poses = (int64_t)n_poses * N;             // simulated ray fans (= frames without the elevational PSF)
        if (!async || c->tree_budget > 0) CUDA_TRY(cudaStreamSynchronize(s));
        if (c->tree_budget > 0)
            for (int b = 0; b < n_batches; b++)
                if (c->h_tree[2 * b + 1] || c->h_tree[2 * b] > c->tree.seg_capacity)
                    return fail(MCRT_ERR_NOMEM, "ray_tree: the segment budget per path was exceeded; raise option ray_tree");
    } catch (const CudaError& e) {
        return fail(MCRT_ERR_CUDA, e.what());
    } catch (const std::exception& e) {
        return fail(MCRT_ERR_INVALID, e.what());
    }
    return MCRT_OK;
}

void finalize_stats(mcrt_ctx* c)
{
    if (!c->stats_pending) return;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) { (void)cudaGetLastError(); return; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->stats.ms_total = ms;
    const int64_t paths_per_pose = (int64_t)c->aq.elements * c->aq.samples;
    int64_t segs = c->stats.poses * paths_per_pose;      // bounce 0 traces every path
    int64_t steps = 0, late = 0;
    for (int b = 0; b < c->pending_batches; b++) {
        for (int k = 0; k < kMaxSub; k++)
            for (int d = 1; d < c->aq.max_depth; d++) segs += c->h_counters[kCounterSlot * b + k * (c->aq.max_depth + 1) + d];
        steps += (int64_t)c->h_steps[2 * b];
        late += (int64_t)c->h_steps[2 * b + 1];
    }
    if (c->tree_budget > 0) { segs = 0; for (int b = 0; b < c->pending_batches; b++) segs += c->h_tree[2 * b]; }
    c->stats.segments = segs;
    c->stats.march_steps = steps;
    c->stats.late_echoes = late;
    c->stats.kernel_launches = c->pending_launches;
    if (c->count_traversal) { c->stats.bvh_node_visits = (int64_t)c->h_trav[0]; c->stats.bvh_triangle_tests = (int64_t)c->h_trav[1]; }
    if (c->profile_stages && c->pending_batches == 1) {
        cudaEventElapsedTime(&c->stats.ms_trace, c->ev_a, c->ev_b);
        cudaEventElapsedTime(&c->stats.ms_accumulate, c->ev_b, c->ev_c);
        cudaEventElapsedTime(&c->stats.ms_post, c->ev_c, c->ev1);
    }
    c->stats_pending = false;
}

int create_impl(HostScene&& scene, const mcrt_params* params, int device, mcrt_ctx** out)
{
    std::unique_ptr<mcrt_ctx> c(new mcrt_ctx());
    c->params = *params;
    validate_params(c->params);
    c->dv = derive(c->params);
    c->scene = std::move(scene);
    if ((int)c->scene.meshes.size() > MCRT_MAX_SMEM_MESHES) throw std::invalid_argument("more than 64 meshes are not supported");
    if ((int)c->scene.materials.size() > MCRT_MAX_SMEM_MATERIALS) throw std::invalid_argument("more than 64 materials are not supported");
    c->device = device;
    int n_dev = 0;
    CUDA_TRY(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) throw CudaError("no CUDA device " + std::to_string(device) + " (libmcrt has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) throw CudaError(std::string("device '") + prop.name + "' is not sm_100 (B200); libmcrt only carries sm_100a code");
    c->sm_count = prop.multiProcessorCount;
    // option "tail_merge" (units of sm_count * 768 paths): once at most this many paths are alive one launch walks them to their end.  12 units
    // (1.36 M paths on a B200) measured best with the round-2 kernels: a call of up to 256 frames of 256 x 16 is traced by ONE launch
    // (+17 % at 32 frames per call, +13 % at 64, +5 % at 128, +1.6 % at 256, config 4 +3..10 %), larger calls compact until they fit;
    // 20 units already lose at 512 frames per call (profiles/r02at_ab_tail_threshold.txt)
    c->tb.tail_threshold = 12 * c->sm_count * 6 * 128;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (int k = 0; k < 4; k++) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_sub[k], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&c->ev0)); CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaEventCreate(&c->ev_a)); CUDA_TRY(cudaEventCreate(&c->ev_b)); CUDA_TRY(cudaEventCreate(&c->ev_c));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));

    // acquisition constants
    AcqDev& aq = c->aq;
    memset(&aq, 0, sizeof(aq));
    aq.elements = c->params.elements; aq.samples = c->params.samples; aq.max_depth = c->params.max_depth; aq.rows = c->dv.rows;
    aq.frequency = c->params.frequency_mhz; aq.axres_f = c->dv.axial_resolution_f;
    aq.radius_f = (float)c->params.radius_cm;                         // radius.to<float>(), transducer.h:51
    aq.vol_resolution = c->params.resolution_um / 1000.0f;            // volume.h:49
    aq.axres_mm = c->dv.axial_resolution_mm; aq.time_step_us = c->dv.time_step_us; aq.row_period_us = c->dv.row_period_us;
    aq.inv_row_period = 1.0 / c->dv.row_period_us; aq.max_travel_time_us = c->dv.max_travel_time_us;
    aq.speed = (double)c->params.speed_of_sound; aq.deterministic = c->params.deterministic;
    aq.rf_pitch = post_preferred_pitch(aq.rows, c->params.psf_axial, c->params.psf_lateral);

    // scene tables
    const HostScene& hs = c->scene;
    std::vector<DevMesh> meshes(hs.meshes.size());
    for (size_t m = 0; m < hs.meshes.size(); m++) {
        DevMesh& d = meshes[m];
        d.ox = hs.meshes[m].origin[0]; d.oy = hs.meshes[m].origin[1]; d.oz = hs.meshes[m].origin[2];
        d.mat_in = hs.meshes[m].material_inside; d.mat_out = hs.meshes[m].material_outside; d.vascular = hs.meshes[m].is_vascular ? 1 : 0;
        d.pad0 = d.pad1 = 0;
    }
    dev_alloc(c->d_meshes, meshes.size() ? meshes.size() : 1);
    if (!meshes.empty()) CUDA_TRY(cudaMemcpy(c->d_meshes, meshes.data(), sizeof(DevMesh) * meshes.size(), cudaMemcpyHostToDevice));
    dev_alloc(c->d_materials, hs.materials.size());
    CUDA_TRY(cudaMemcpy(c->d_materials, hs.materials.data(), sizeof(DevMaterial) * hs.materials.size(), cudaMemcpyHostToDevice));
    make_bvh2(c.get(), &c->bvh);                       // device LBVH, or a validated tree from $MCRT_BVH_CACHE
    if (c->bvh.max_depth > MCRT_TRAVERSAL_STACK) throw std::invalid_argument("BVH deeper than the traversal stack (degenerate mesh?)");
    SceneDev& sc = c->sc;
    memset(&sc, 0, sizeof(sc));
    {
        CUDA_TRY(init_trace_kernels());
        WideTree w;
        build_wide_tree(c->bvh, c->stream, &w);
        c->d_nodes4 = w.n4; c->d_nodes8 = w.n8; c->d_tris8 = w.t8; c->n_nodes_wide = w.n_nodes; c->depth_wide = w.depth;
    }
    sc.nodes = c->bvh.nodes; sc.nodes4 = c->d_nodes4; sc.nodes8 = c->d_nodes8; sc.tris = c->d_tris8 ? c->d_tris8 : c->bvh.tris;
    sc.meshes = c->d_meshes; sc.materials = c->d_materials;
    sc.n_tri = c->bvh.n_tri; sc.n_mesh = (int)hs.meshes.size(); sc.n_mat = (int)hs.materials.size();
    sc.starting_material = hs.starting_material;
    for (int a = 0; a < 3; a++) sc.spacing[a] = hs.spacing[a];
    c->aq.accumulate_windowed = accumulate_windowed_supported(sc, c->aq) ? 1 : 0;
    sc.max_abs = c->bvh.max_abs;
    update_scene_bounds(c.get());
    {   // the better tree is built on a host thread while the rest of the start-up runs (MCRT_BVH_OPTIMISE=0: never)
        const char* e = getenv("MCRT_BVH_OPTIMISE");
        if (e && *e == '0') c->bvh_optimise = 0;
        c->quiet_calls = 8;
        bg_tree_start(c.get());
    }

    // transducer table, psf taps, scan maps, scatterer volume
    std::vector<float> sincos;
    element_angle_table(c->params, c->dv, sincos);
    dev_alloc(c->d_elem_sincos, (size_t)c->params.elements);
    CUDA_TRY(cudaMemcpy(c->d_elem_sincos, sincos.data(), sizeof(float) * sincos.size(), cudaMemcpyHostToDevice));
    psf_taps(c->params, c->h_axial, c->h_lateral);
    dev_alloc(c->d_axial, c->h_axial.size());
    dev_alloc(c->d_lateral, c->h_lateral.size());
    CUDA_TRY(cudaMemcpy(c->d_axial, c->h_axial.data(), sizeof(float) * c->h_axial.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_lateral, c->h_lateral.data(), sizeof(float) * c->h_lateral.size(), cudaMemcpyHostToDevice));
    std::vector<float> mx, my;
    scan_mapping(c->params, c->dv, mx, my);
    dev_alloc(c->d_map_x, mx.size());
    dev_alloc(c->d_map_y, my.size());
    CUDA_TRY(cudaMemcpy(c->d_map_x, mx.data(), sizeof(float) * mx.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_map_y, my.data(), sizeof(float) * my.size(), cudaMemcpyHostToDevice));
    c->d_volume = device_volume(device, c->stream);
    CUDA_TRY(init_image_kernels());
    {   // enable the 3-instruction voxel index only if it is provably the reference's for this resolution
        bool ok = false;
        CUDA_TRY(validate_fma_division(c->aq.vol_resolution, &ok));
        c->voxel_fma_validated = ok;
        c->aq.voxel_fma_division = ok ? 1 : 0;
    }

    dev_alloc(c->d_seed_frame, 4);                      // seed, first frame | output target: base pointer, image stride (floats)
    dev_alloc(c->d_trav, 2);
    dev_alloc(c->d_steps, 2);
    CUDA_TRY(cudaMallocHost(&c->h_seed_frame, 4 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMallocHost(&c->h_counters, sizeof(int) * kCounterSlot * kMaxBatchesPerCall));
    CUDA_TRY(cudaMallocHost(&c->h_steps, sizeof(unsigned long long) * 2 * kMaxBatchesPerCall));
    CUDA_TRY(cudaMallocHost(&c->h_trav, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMallocHost(&c->h_tree, sizeof(int) * 2 * kMaxBatchesPerCall));
    memset(c->h_tree, 0, sizeof(int) * 2 * kMaxBatchesPerCall);
    c->h_trav[0] = c->h_trav[1] = 0;
    memset(c->h_counters, 0, sizeof(int) * kCounterSlot * kMaxBatchesPerCall);
    memset(c->h_steps, 0, sizeof(unsigned long long) * 2 * kMaxBatchesPerCall);
    *out = c.release();
    return MCRT_OK;
}

void destroy_impl(mcrt_ctx* c)
{
    if (!c) return;
    if (c->bg) { c->bg->cancel.store(true); if (c->bg->th.joinable()) c->bg->th.join(); }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_workspace(c);
    dev_free(c->d_meshes); dev_free(c->d_materials); dev_free(c->bvh.nodes); dev_free(c->bvh.tris); dev_free(c->d_nodes4);
    dev_free(c->d_nodes8); dev_free(c->d_tris8);
    dev_free(c->d_elev_w);
    dev_free(c->bm_gain); dev_free(c->bm_env); dev_free(c->bm_cmp); dev_free(c->bm_scan); dev_free(c->bm_q); dev_free(c->bm_max);
    dev_free(c->d_elem_sincos); dev_free(c->d_axial); dev_free(c->d_lateral); dev_free(c->d_lat_by_row); dev_free(c->d_map_x); dev_free(c->d_map_y);
    dev_free(c->d_seed_frame); dev_free(c->d_steps); dev_free(c->d_trav);
    if (c->h_seed_frame) cudaFreeHost(c->h_seed_frame);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_steps) cudaFreeHost(c->h_steps);
    if (c->h_tree) cudaFreeHost(c->h_tree);
    if (c->h_trav) cudaFreeHost(c->h_trav);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->ev_c) cudaEventDestroy(c->ev_c);
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (int k = 0; k < 4; k++) if (c->ev_sub[k]) cudaEventDestroy(c->ev_sub[k]);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    (void)cudaGetLastError();
    delete c;
}

template <typename F>
int guarded(const char* what, F&& f)
{
    try {
        return f();
    } catch (const CudaError& e) {
        return fail(MCRT_ERR_CUDA, std::string(what) + ": " + e.what());
    } catch (const std::bad_alloc&) {
        return fail(MCRT_ERR_NOMEM, std::string(what) + ": out of host memory");
    } catch (const std::invalid_argument& e) {
        return fail(MCRT_ERR_INVALID, std::string(what) + ": " + e.what());
    } catch (const std::exception& e) {
        return fail(MCRT_ERR_SCENE, e.what());
    }
}

DevSegment to_dev_segment(const mcrt_segment& s)
{
    DevSegment d;
    d.s0 = make_float4(s.from[0], s.from[1], s.from[2], s.reflected_intensity);
    d.s1 = make_float4(s.dir[0], s.dir[1], s.dir[2], s.initial_intensity);
    d.s2 = make_float4(s.to[0], s.to[1], s.to[2], s.attenuation);
    unsigned long long bits;
    memcpy(&bits, &s.distance_traveled, 8);
    d.s3 = make_int4((int)(bits & 0xffffffffu), (int)(bits >> 32), s.media_id, s.tri_id);
    return d;
}

}  // namespace

// ================================================================================================
extern "C" {

const char* mcrt_last_error(void) { return g_last_error.c_str(); }

int mcrt_default_params(mcrt_params* p)
{
    if (!p) return fail(MCRT_ERR_INVALID, "mcrt_default_params: null");
    memset(p, 0, sizeof(*p));
    p->elements = 512; p->samples = 5; p->max_depth = 10; p->frequency_mhz = 4.5f; p->radius_cm = 3; p->fov_deg = 60; p->depth_cm = 15;
    p->speed_of_sound = 1500; p->resolution_um = 145; p->psf_axial = 7; p->psf_lateral = 13; p->psf_var_x = 0.05f; p->psf_var_y = 0.2f;
    p->deterministic = 0; p->scan_rows = 400; p->scan_cols = 500; p->axial_scale = 1.0f; p->rf_layout = 0;
    return MCRT_OK;
}

int mcrt_create(const char* scene_json_path, const mcrt_params* params, int device, mcrt_ctx** out)
{
    if (!scene_json_path || !out) return fail(MCRT_ERR_INVALID, "mcrt_create: null argument");
    *out = nullptr;
    mcrt_params defaults;
    mcrt_default_params(&defaults);
    const mcrt_params* p = params ? params : &defaults;
    return guarded("mcrt_create", [&]() { return create_impl(load_scene_file(scene_json_path), p, device, out); });
}

int mcrt_create_from_arrays(const mcrt_scene_arrays* scene, const mcrt_params* params, int device, mcrt_ctx** out)
{
    if (!scene || !out) return fail(MCRT_ERR_INVALID, "mcrt_create_from_arrays: null argument");
    *out = nullptr;
    mcrt_params defaults;
    mcrt_default_params(&defaults);
    const mcrt_params* p = params ? params : &defaults;
    return guarded("mcrt_create_from_arrays", [&]() { return create_impl(scene_from_arrays(*scene), p, device, out); });
}

void mcrt_destroy(mcrt_ctx* ctx) { destroy_impl(ctx); }

// 64-bit FNV-1a over the data that determines the SAH tree (the cache key)
static uint64_t fnv1a64(const void* data, size_t n, uint64_t h)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

// Tree cache: $MCRT_BVH_CACHE/<builder>_<hash>.bvh = {magic, n_nodes, n_slots, 0, max_depth, max_abs, nodes, slots}
// for every builder (lbvh = the default device build, sah = the host binned-SAH tree).  The key hashes the triangle data, the
// body origins, the builder and a format / builder version, so a file written by another version is simply never looked up;
// a file that IS found is still validated in full before use (child references, leaf slots, every slot's vertices against
// the scene, the depth): a stale, truncated or hostile file in a shared cache directory falls back to rebuilding.
static const uint64_t kTreeCacheVersion = 3;          // bump when a builder or the file layout changes
static std::string tree_cache_path(const HostScene& hs, const std::vector<float>& origins, int builder)
{
    const char* dir = getenv("MCRT_BVH_CACHE");
    if (!dir || !*dir) return std::string();
    uint64_t h = 1469598103934665603ULL;
    const uint64_t tag[3] = {kTreeCacheVersion, (uint64_t)builder, (uint64_t)MCRT_LEAF_MAX};
    h = fnv1a64(tag, sizeof(tag), h);
    h = fnv1a64(hs.tri_local.data(), sizeof(float) * hs.tri_local.size(), h);
    h = fnv1a64(hs.tri_mesh.data(), sizeof(int32_t) * hs.tri_mesh.size(), h);
    h = fnv1a64(origins.data(), sizeof(float) * origins.size(), h);
    char name[64];
    snprintf(name, sizeof(name), "/%s_%016llx.bvh", builder == 0 ? "lbvh" : (builder == 1 ? "sah" : "ploc"), (unsigned long long)h);
    return std::string(dir) + name;
}
static const uint64_t kTreeCacheMagic = 0x3242564854544d43ULL;   // "CMTTHVB2"
static bool tree_cache_valid(const HostBvh& hb, const HostScene& hs)
{
    const size_t n_tri = hs.tri_mesh.size(), n_nodes = hb.nodes.size();
    if (hb.slots.size() != n_tri || n_nodes + 1 != (n_tri ? n_tri : 1)) return false;
    std::vector<unsigned char> seen(n_tri, 0);
    for (size_t k = 0; k < n_tri; k++) {
        const HostTriSlot& t = hb.slots[k];
        if (t.tri < 0 || (size_t)t.tri >= n_tri || seen[t.tri] || t.mesh != hs.tri_mesh[t.tri]) return false;
        if (memcmp(t.v, hs.tri_local.data() + 9 * (size_t)t.tri, sizeof(float) * 9) != 0) return false;
        seen[t.tri] = 1;
    }
    if (n_nodes == 0) return true;
    // every node reachable exactly once from the root, every leaf slot referenced exactly once, depth as recorded
    std::vector<unsigned char> node_seen(n_nodes, 0), slot_seen(n_tri, 0);
    std::vector<std::pair<int, int>> stack;
    stack.emplace_back(0, 1);
    int deepest = 0;
    size_t visited = 0;
    while (!stack.empty()) {
        const std::pair<int, int> top = stack.back();
        stack.pop_back();
        if (top.first < 0 || (size_t)top.first >= n_nodes || node_seen[top.first]) return false;
        node_seen[top.first] = 1; visited++;
        for (int k = 0; k < 2; k++) {
            const int ch = hb.nodes[top.first].child[k];
            if (ch >= 0) { stack.emplace_back(ch, top.second + 1); continue; }
            const int code = -ch - 1, first = code >> 2, count = (code & 3) + 1;
            if (count > MCRT_LEAF_MAX || first < 0 || (size_t)(first + count) > n_tri) return false;
            for (int j = 0; j < count; j++) { if (slot_seen[first + j]) return false; slot_seen[first + j] = 1; }
            if (top.second + 1 > deepest) deepest = top.second + 1;
        }
        for (int k = 0; k < 12; k++) if (!(hb.nodes[top.first].f[k] == hb.nodes[top.first].f[k])) return false;      // NaN planes
    }
    for (size_t k = 0; k < n_tri; k++) if (!slot_seen[k]) return false;
    // deepest counts the root as 1 and the leaf as one more: deepest - 1 inner nodes on the longest chain
    return visited == n_nodes && deepest - 1 <= hb.max_depth && hb.max_depth <= MCRT_TRAVERSAL_STACK && hb.max_abs >= 0.0f && hb.max_abs < 3.0e38f;
}
static bool tree_cache_load(const std::string& path, const HostScene& hs, HostBvh* hb)
{
    FILE* f = path.empty() ? nullptr : fopen(path.c_str(), "rb");
    if (!f) return false;
    const size_t n_tri = hs.tri_mesh.size();
    uint64_t head[4] = {0, 0, 0, 0};
    bool ok = fread(head, sizeof(head), 1, f) == 1 && head[0] == kTreeCacheMagic && head[2] == n_tri && head[1] + 1 == (n_tri ? n_tri : 1);
    int32_t depth = 0; float max_abs = 0.0f;
    ok = ok && fread(&depth, sizeof(depth), 1, f) == 1 && fread(&max_abs, sizeof(max_abs), 1, f) == 1;
    if (ok) {
        hb->nodes.resize(head[1]); hb->slots.resize(head[2]);
        ok = (head[1] == 0 || fread(hb->nodes.data(), sizeof(HostBvhNode), head[1], f) == head[1]) &&
             (head[2] == 0 || fread(hb->slots.data(), sizeof(HostTriSlot), head[2], f) == head[2]);
        hb->max_depth = depth; hb->max_abs = max_abs;
    }
    fclose(f);
    return ok && tree_cache_valid(*hb, hs);
}
static void tree_cache_store(const std::string& path, const HostBvh& hb)
{
    if (path.empty()) return;
    const std::string tmp = path + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                   // an unwritable cache directory is not an error
    const uint64_t head[4] = {kTreeCacheMagic, hb.nodes.size(), hb.slots.size(), 0};
    const int32_t depth = hb.max_depth;
    bool ok = fwrite(head, sizeof(head), 1, f) == 1 && fwrite(&depth, sizeof(depth), 1, f) == 1 && fwrite(&hb.max_abs, sizeof(float), 1, f) == 1;
    ok = ok && (hb.nodes.empty() || fwrite(hb.nodes.data(), sizeof(HostBvhNode), hb.nodes.size(), f) == hb.nodes.size());
    ok = ok && (hb.slots.empty() || fwrite(hb.slots.data(), sizeof(HostTriSlot), hb.slots.size(), f) == hb.slots.size());
    fclose(f);
    if (ok) rename(tmp.c_str(), path.c_str()); else remove(tmp.c_str());
}
// host tree -> device arrays (a cache hit of either builder, or a fresh host SAH build)
static void upload_host_tree(const HostBvh& hb, LbvhResult* nb)
{
    const size_t n = hb.slots.size();
    std::vector<TriSlot> slots(n);
    for (size_t k = 0; k < n; k++) {
        const HostTriSlot& t = hb.slots[k];
        int mbits = t.mesh, tbits = t.tri;
        float mf, tf;
        memcpy(&mf, &mbits, 4); memcpy(&tf, &tbits, 4);
        slots[k].v0 = make_float4(t.v[0], t.v[1], t.v[2], mf);
        slots[k].v1 = make_float4(t.v[3], t.v[4], t.v[5], tf);
        slots[k].v2 = make_float4(t.v[6], t.v[7], t.v[8], 0.f);
    }
    static_assert(sizeof(HostBvhNode) == sizeof(BvhNode), "node layouts must match");
    if (n) { dev_alloc(nb->tris, n); CUDA_TRY(cudaMemcpy(nb->tris, slots.data(), sizeof(TriSlot) * n, cudaMemcpyHostToDevice)); }
    if (!hb.nodes.empty()) {
        dev_alloc(nb->nodes, hb.nodes.size());
        CUDA_TRY(cudaMemcpy(nb->nodes, hb.nodes.data(), sizeof(BvhNode) * hb.nodes.size(), cudaMemcpyHostToDevice));
    }
    nb->n_tri = (int)n; nb->n_nodes = (int)hb.nodes.size(); nb->max_depth = hb.max_depth; nb->max_abs = hb.max_abs;
}
// device LBVH -> host tree (to store it)
static void download_device_tree(const LbvhResult& b, HostBvh* hb)
{
    hb->nodes.resize((size_t)b.n_nodes); hb->slots.resize((size_t)b.n_tri);
    std::vector<TriSlot> slots((size_t)b.n_tri);
    if (b.n_nodes) CUDA_TRY(cudaMemcpy(hb->nodes.data(), b.nodes, sizeof(BvhNode) * (size_t)b.n_nodes, cudaMemcpyDeviceToHost));
    if (b.n_tri) CUDA_TRY(cudaMemcpy(slots.data(), b.tris, sizeof(TriSlot) * (size_t)b.n_tri, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < slots.size(); k++) {
        HostTriSlot& t = hb->slots[k];
        const float v[9] = {slots[k].v0.x, slots[k].v0.y, slots[k].v0.z, slots[k].v1.x, slots[k].v1.y, slots[k].v1.z, slots[k].v2.x, slots[k].v2.y, slots[k].v2.z};
        memcpy(t.v, v, sizeof(v));
        memcpy(&t.mesh, &slots[k].v0.w, 4); memcpy(&t.tri, &slots[k].v1.w, 4);
    }
    hb->max_depth = b.max_depth; hb->max_abs = b.max_abs;
}

// scene bounds for the coherence-sort keys
static void update_scene_bounds(mcrt_ctx* c)
{
    const HostScene& hs = c->scene;
    SceneDev& sc = c->sc;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (size_t t = 0; t < hs.tri_mesh.size(); t++)
        for (int k = 0; k < 3; k++)
            for (int a = 0; a < 3; a++) {
                const float w = hs.tri_local[9 * t + 3 * k + a] + hs.meshes[hs.tri_mesh[t]].origin[a];
                lo[a] = w < lo[a] ? w : lo[a]; hi[a] = w > hi[a] ? w : hi[a];
            }
    for (int a = 0; a < 3; a++) {
        const float ext = hi[a] - lo[a];
        sc.bounds_lo[a] = hs.tri_mesh.empty() ? 0.0f : lo[a];
        sc.bounds_inv[a] = (!hs.tri_mesh.empty() && ext > 0.0f) ? 1.0f / ext : 0.0f;
    }
}

// the BVH2 of the current scene with the selected builder: from the validated disk cache when there is one, else built
static void make_bvh2(mcrt_ctx* c, LbvhResult* nb)
{
    const HostScene& hs = c->scene;
    memset(nb, 0, sizeof(*nb));
    std::vector<float> origins(hs.meshes.size() * 3 + 3);
    for (size_t m = 0; m < hs.meshes.size(); m++)
        for (int a = 0; a < 3; a++) origins[3 * m + a] = hs.meshes[m].origin[a];
    const std::string cache = tree_cache_path(hs, origins, c->bvh_builder);
    HostBvh hb;
    c->bvh_cache_hit = tree_cache_load(cache, hs, &hb);
    if (c->bvh_cache_hit) {
        upload_host_tree(hb, nb);
    } else if (c->bvh_builder == 0 || c->bvh_builder == 2) {
        const cudaError_t be = build_lbvh(hs.tri_local.data(), hs.tri_mesh.data(), (int)hs.tri_mesh.size(), c->d_meshes, c->stream, nb, c->bvh_builder == 2);
        if (be != cudaSuccess) throw CudaError(std::string("build_lbvh: ") + cudaGetErrorString(be));
        if (!cache.empty()) { download_device_tree(*nb, &hb); tree_cache_store(cache, hb); }     // only when a cache directory is set
    } else {
        build_sah_bvh(hs.tri_local.data(), hs.tri_mesh.data(), (int)hs.tri_mesh.size(), origins.data(), &hb);
        tree_cache_store(cache, hb);
        upload_host_tree(hb, nb);
    }
}

// (Re)build the acceleration structure from c->scene with the selected builder and swap it in.  Used by the bvh_builder
// option and after mesh updates (mcrt_set_mesh_origin / mcrt_set_mesh_vertices): the device LBVH build is ~0.3 ms of
// kernels for 624 640 triangles, so moving or deforming meshes are handled by rebuilding, not by refitting.
// a freshly built BVH2 becomes the traversal structure of the context (wide tree built, old trees freed, graphs dropped: they hold
// the old pointers).  Takes ownership of nb's device arrays.
static void install_bvh2(mcrt_ctx* c, LbvhResult nb)
{
    if (nb.max_depth > MCRT_TRAVERSAL_STACK) { cudaFree(nb.nodes); cudaFree(nb.tris); throw std::invalid_argument("BVH deeper than the traversal stack"); }
    WideTree w;
    try {
        build_wide_tree(nb, c->stream, &w);
    } catch (...) { cudaFree(nb.nodes); cudaFree(nb.tris); throw; }
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
    dev_free(c->bvh.nodes); dev_free(c->bvh.tris); dev_free(c->d_nodes4); dev_free(c->d_nodes8); dev_free(c->d_tris8);
    c->bvh = nb;
    c->d_nodes4 = w.n4; c->d_nodes8 = w.n8; c->d_tris8 = w.t8; c->n_nodes_wide = w.n_nodes; c->depth_wide = w.depth;
    c->sc.nodes = nb.nodes; c->sc.nodes4 = w.n4; c->sc.nodes8 = w.n8; c->sc.tris = w.t8 ? w.t8 : nb.tris; c->sc.n_tri = nb.n_tri; c->sc.max_abs = nb.max_abs;
}

static void rebuild_bvh(mcrt_ctx* c)
{
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    LbvhResult nb{};
    make_bvh2(c, &nb);
    install_bvh2(c, nb);
    update_scene_bounds(c);
    c->scene_dirty = false;
    c->bvh_optimised = false;
}

// ---- background tree optimisation (see mcrt_ctx::BgTree) ----
static const size_t kOptimiseMinTriangles = 32768;     // below this a closest-hit query visits so few nodes that the better tree buys nothing
static std::vector<float> scene_origins(const HostScene& hs)
{
    std::vector<float> origins(hs.meshes.size() * 3 + 3);
    for (size_t m = 0; m < hs.meshes.size(); m++)
        for (int a = 0; a < 3; a++) origins[3 * m + a] = hs.meshes[m].origin[a];
    return origins;
}
static void bg_tree_cancel(mcrt_ctx* c)
{
    if (!c->bg) return;
    c->bg->cancel.store(true);
    if (c->bg->th.joinable()) c->bg->th.join();
    c->bg->cancel.store(false);
    c->bg->state.store(0);
    c->bg->tree = HostBvh();
}
// start the optimisation of the CURRENT scene unless it is pointless, disabled, already done or already running
static void bg_tree_start(mcrt_ctx* c)
{
    const HostScene& hs = c->scene;
    if (!c->bvh_optimise || c->bvh_builder != 0 || c->bvh_optimised || c->scene_dirty || hs.tri_mesh.size() < kOptimiseMinTriangles) return;
    if (c->bg && c->bg->state.load() != 0) return;       // building, or a result waits to be adopted
    if (!c->bg) c->bg.reset(new mcrt_ctx::BgTree());
    mcrt_ctx::BgTree* bg = c->bg.get();
    if (bg->th.joinable()) bg->th.join();
    bg->version = c->scene_version;
    bg->origins = scene_origins(hs);
    // a tree of this very scene in $MCRT_BVH_CACHE (validated) needs no thread
    const std::string cache = tree_cache_path(hs, bg->origins, 1);
    if (tree_cache_load(cache, hs, &bg->tree)) { bg->state.store(2); return; }
    bg->tri_local = hs.tri_local; bg->tri_mesh = hs.tri_mesh;          // snapshot: mesh updates may be staged while the thread runs
    bg->state.store(1);
    bg->th = std::thread([bg]() {
        try {
            build_sah_bvh(bg->tri_local.data(), bg->tri_mesh.data(), (int)bg->tri_mesh.size(), bg->origins.data(), &bg->tree, 0, &bg->cancel);
            bg->state.store(2);
        } catch (...) { bg->state.store(3); }
    });
}
// swap the finished tree in (wait: block until the running build has finished).  Called at the start of compute entry points.
static void bg_tree_adopt(mcrt_ctx* c, bool wait)
{
    mcrt_ctx::BgTree* bg = c->bg.get();
    if (!bg) return;
    int st = bg->state.load();
    if (st == 1 && wait) { if (bg->th.joinable()) bg->th.join(); st = bg->state.load(); }
    if (st != 2 && st != 3) return;
    if (bg->th.joinable()) bg->th.join();
    bg->state.store(0);
    const bool usable = st == 2 && bg->version == c->scene_version && !c->scene_dirty && c->bvh_builder == 0 && c->bvh_optimise;
    if (usable) {
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (c->last_valid) CUDA_TRY(cudaEventSynchronize(c->ev_last));          // work of an asynchronous call may still read the old tree
        LbvhResult nb{};
        upload_host_tree(bg->tree, &nb);
        install_bvh2(c, nb);
        c->bvh_optimised = true;
        tree_cache_store(tree_cache_path(c->scene, bg->origins, 1), bg->tree);  // no-op without $MCRT_BVH_CACHE
    }
    bg->tree = HostBvh(); bg->tri_local.clear(); bg->tri_local.shrink_to_fit(); bg->tri_mesh.clear(); bg->tri_mesh.shrink_to_fit();
}

// mesh updates are staged on the host and applied by ONE rebuild at the next compute call
static void ensure_scene_current(mcrt_ctx* c)
{
    if (!c->scene_dirty) {
        bg_tree_adopt(c, false);
        // after mesh updates the optimiser waits for 8 compute calls on an unchanged scene before it starts again
        if (!c->bvh_optimised && c->bvh_optimise && c->bvh_builder == 0) {
            if (c->quiet_calls < 8) c->quiet_calls++;
            if (c->quiet_calls >= 8) bg_tree_start(c);
        }
        return;
    }
    c->quiet_calls = 0;
    const HostScene& hs = c->scene;
    std::vector<DevMesh> meshes(hs.meshes.size());
    for (size_t m = 0; m < hs.meshes.size(); m++) {
        DevMesh& d = meshes[m];
        d.ox = hs.meshes[m].origin[0]; d.oy = hs.meshes[m].origin[1]; d.oz = hs.meshes[m].origin[2];
        d.mat_in = hs.meshes[m].material_inside; d.mat_out = hs.meshes[m].material_outside; d.vascular = hs.meshes[m].is_vascular ? 1 : 0;
        d.pad0 = d.pad1 = 0;
    }
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (!meshes.empty()) CUDA_TRY(cudaMemcpy(c->d_meshes, meshes.data(), sizeof(DevMesh) * meshes.size(), cudaMemcpyHostToDevice));
    rebuild_bvh(c);
}

int mcrt_set_mesh_origin(mcrt_ctx* c, int32_t mesh, const float* origin3)
{
    if (!c || !origin3) return fail(MCRT_ERR_INVALID, "mcrt_set_mesh_origin: null argument");
    if (mesh < 0 || (size_t)mesh >= c->scene.meshes.size()) return fail(MCRT_ERR_INVALID, "mcrt_set_mesh_origin: no such mesh");
    for (int a = 0; a < 3; a++) c->scene.meshes[mesh].origin[a] = origin3[a];
    c->scene_dirty = true; c->scene_version++;
    return MCRT_OK;
}

int mcrt_set_mesh_vertices(mcrt_ctx* c, int32_t mesh, const float* tri_local9, int64_t n_triangles)
{
    if (!c || !tri_local9) return fail(MCRT_ERR_INVALID, "mcrt_set_mesh_vertices: null argument");
    if (mesh < 0 || (size_t)mesh >= c->scene.meshes.size()) return fail(MCRT_ERR_INVALID, "mcrt_set_mesh_vertices: no such mesh");
    HostScene& hs = c->scene;
    size_t first = hs.tri_mesh.size(), count = 0;
    for (size_t t = 0; t < hs.tri_mesh.size(); t++)
        if (hs.tri_mesh[t] == mesh) { if (count == 0) first = t; count++; }
    if ((int64_t)count != n_triangles) return fail(MCRT_ERR_INVALID, "mcrt_set_mesh_vertices: the triangle count of a mesh cannot change");
    if (count) memcpy(hs.tri_local.data() + 9 * first, tri_local9, sizeof(float) * 9 * count);
    c->scene_dirty = true; c->scene_version++;
    return MCRT_OK;
}

int mcrt_get_info(const mcrt_ctx* c, mcrt_info* info)
{
    if (!c || !info) return fail(MCRT_ERR_INVALID, "mcrt_get_info: null argument");
    memset(info, 0, sizeof(*info));
    info->rows = c->dv.rows; info->cols = c->dv.cols; info->scan_rows = c->params.scan_rows; info->scan_cols = c->params.scan_cols;
    info->n_materials = (int)c->scene.materials.size(); info->n_meshes = (int)c->scene.meshes.size();
    info->n_triangles = (int64_t)c->scene.tri_mesh.size(); info->n_bvh_nodes = c->bvh.n_nodes; info->device = c->device;
    info->sm_count = c->sm_count;
    for (int i = 0; i < 6; i++) info->start_pose[i] = c->scene.start_pose[i];
    info->axial_resolution_mm = c->dv.axial_resolution_mm; info->time_step_us = c->dv.time_step_us;
    info->row_period_us = c->dv.row_period_us; info->max_travel_time_us = c->dv.max_travel_time_us;
    info->voxel_fma_division = c->aq.voxel_fma_division;
    info->bvh_cache_hit = c->bvh_cache_hit ? 1 : 0;
    info->bvh_optimised = c->bvh_optimised ? 1 : 0;
    return MCRT_OK;
}

int mcrt_get_stats(const mcrt_ctx* c, mcrt_stats* stats)
{
    if (!c || !stats) return fail(MCRT_ERR_INVALID, "mcrt_get_stats: null argument");
    finalize_stats(const_cast<mcrt_ctx*>(c));
    *stats = c->stats;
    return MCRT_OK;
}

int mcrt_set_option(mcrt_ctx* c, const char* name, int64_t value)
{
    if (!c || !name) return fail(MCRT_ERR_INVALID, "mcrt_set_option: null argument");
    const std::string n(name);
    if (n == "profile_stages") c->profile_stages = value != 0;
    else if (n == "use_graph") c->use_graph = value != 0;
    else if (n == "overlap") {
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        c->overlap = value != 0;
    }
    else if (n == "bvh_builder") {
        // 0: device LBVH (default, lbvh.cu) + background optimisation (option "bvh_optimise"); 1: host binned-SAH tree
        // (sah_builder.cpp) built in place; 2: device PLOC.  Every tree is cached on disk when the environment variable
        // MCRT_BVH_CACHE names a directory.  Rebuilds in place.
        return guarded("mcrt_set_option", [&]() {
            if (value < 0 || value > 2) throw std::invalid_argument("bvh_builder: 0 device LBVH, 1 host binned SAH, 2 device PLOC");
            bg_tree_cancel(c);
            c->bvh_builder = (int)value;
            rebuild_bvh(c);
            c->quiet_calls = 8;
            bg_tree_start(c);
            return MCRT_OK;
        });
    }
    else if (n == "bvh_optimise") {
        // 1 (default): with builder 0, a host thread builds the binned-SAH tree of the scene in the background and the first compute
        // call after it has finished adopts it; 0: keep (or go back to) the plain device LBVH
        return guarded("mcrt_set_option", [&]() {
            c->bvh_optimise = value != 0 ? 1 : 0;
            if (!c->bvh_optimise) { bg_tree_cancel(c); if (c->bvh_optimised) rebuild_bvh(c); }
            else { c->quiet_calls = 8; bg_tree_start(c); }
            return MCRT_OK;
        });
    }
    else if (n == "bvh_wait") {
        // block until a running background optimisation has finished and adopt its tree (benchmarks: no swap inside a timed region)
        return guarded("mcrt_set_option", [&]() {
            if (c->scene_dirty) ensure_scene_current(c);
            // (a build that was started before the last mesh update is stale: its result is dropped and a second round builds the current scene)
            for (int round = 0; value != 0 && round < 2 && !c->bvh_optimised; round++) { c->quiet_calls = 8; bg_tree_start(c); bg_tree_adopt(c, true); }
            return MCRT_OK;
        });
    }
    else if (n == "log_compress") {
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        c->log_compress = value != 0;
    }
    else if (n == "ray_tree") {
        // > 0: follow BOTH children of every boundary hit (TreeBuffers); the value is the segment budget per path
        if (value < 0 || value > 4096) return fail(MCRT_ERR_INVALID, "ray_tree: budget must be in [0, 4096] segments per path");
        CUDA_TRY_NOTHROW(cudaStreamSynchronize(c->stream));
        free_workspace(c);
        c->tree_budget = (int)value;
    }
    else if (n == "tail_merge") {
        // N (default 1): once no more paths are alive than N resident waves of k_bounce threads, one launch finishes them
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        if (value < 0 || value > 64) return fail(MCRT_ERR_INVALID, "tail_merge: 0 (off) .. 64 resident waves");
        c->tb.tail_threshold = (int)value * c->sm_count * 6 * 128;
    }
    else if (n == "first_hit_dedup") {
        CUDA_TRY_NOTHROW(cudaStreamSynchronize(c->stream));
        free_workspace(c);
        c->first_hit_dedup = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
    }
    else if (n == "ordered_compaction") {
        // changes the workspace and the captured graphs: drop both, they are rebuilt on the next call
        CUDA_TRY_NOTHROW(cudaStreamSynchronize(c->stream));
        free_workspace(c);
        c->ordered_compaction = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
    }
    else if (n == "count_traversal") {
        // changes the kernel arguments baked into captured graphs: drop them
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        c->count_traversal = value != 0;
        c->tb.trav_counters = c->count_traversal ? c->d_trav : nullptr;
    }
    else if (n == "accumulate_windowed") {
        // changes the workspace (HBM columns or not) and the captured graphs
        CUDA_TRY_NOTHROW(cudaStreamSynchronize(c->stream));
        free_workspace(c);
        c->aq.accumulate_windowed = (value != 0 && accumulate_windowed_supported(c->sc, c->aq)) ? 1 : 0;
    }
    else if (n == "frame_stride") {
        // frame (Philox counter) of pose i of a call = first_frame + i * stride: a sweep dealt out round-robin to G GPUs (rank r
        // simulates poses r, r + G, ...) keeps its global frame indices with stride G, so N GPUs stay bit-identical to one
        if (value < 1 || value > 65536) return fail(MCRT_ERR_INVALID, "frame_stride: 1 .. 65536");
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        c->frame_stride = (int)value;
    }
    else if (n == "post_tma") {
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        c->post_tma = value != 0;
    }
    else if (n == "direct_out") c->direct_out = value != 0;          // A/B switch; decided per call, part of the graph key
    else if (n == "rf_out_frame_stride") {
        if (value < 1 || value > 65536) return fail(MCRT_ERR_INVALID, "rf_out_frame_stride: 1 .. 65536 frames");
        c->out_frame_stride = (int)value;
    }
    else if (n == "group_histories") {
        // changes a kernel argument baked into captured graphs: drop them (the workspace keeps its size)
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        c->group_histories = value != 0;
        c->tb.group_histories = c->group_histories ? 1 : 0;
    }
    else if (n == "long_ct") {
        // A/B switch (process-wide): compile-time-tap axial / lateral kernels + shared-memory envelope on the long-scanline path
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        set_long_scanline_ct(value != 0);
    }
    else if (n == "voxel_fma_division") {
        // A/B switch; can only be turned on for a resolution that passed the exhaustive check at mcrt_create
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        c->aq.voxel_fma_division = (value != 0 && c->voxel_fma_validated) ? 1 : 0;
    }
    else if (n == "max_batch_poses") { if (value < 1) return fail(MCRT_ERR_INVALID, "max_batch_poses must be >= 1"); c->max_batch_poses = (int)value; }
    else return fail(MCRT_ERR_INVALID, "unknown option '" + n + "'");
    return MCRT_OK;
}

int mcrt_simulate(mcrt_ctx* ctx, const mcrt_pose* poses, int32_t n_poses, uint64_t seed, uint64_t first_frame, float* rf_out,
                  float* scan_out)
{
    return simulate_impl(ctx, poses, n_poses, seed, first_frame, rf_out, scan_out, nullptr, false);
}

int mcrt_simulate_async(mcrt_ctx* ctx, const mcrt_pose* poses, int32_t n_poses, uint64_t seed, uint64_t first_frame, float* rf_out_dev,
                        float* scan_out_dev, void* cuda_stream)
{
    return simulate_impl(ctx, poses, n_poses, seed, first_frame, rf_out_dev, scan_out_dev, (cudaStream_t)cuda_stream, true);
}

int mcrt_simulate_scanlines(mcrt_ctx* c, const mcrt_pose* pose, uint64_t seed, uint64_t frame, int32_t first_element, int32_t n_elements,
                            float* rf_out)
{
    if (!c || !pose || !rf_out) return fail(MCRT_ERR_INVALID, "mcrt_simulate_scanlines: null argument");
    if (first_element < 0 || n_elements < 1 || first_element + n_elements > c->aq.elements)
        return fail(MCRT_ERR_INVALID, "mcrt_simulate_scanlines: scanline block out of range");
    if (c->log_compress) return fail(MCRT_ERR_INVALID, "mcrt_simulate_scanlines: log_compress needs the whole frame (its maximum)");
    return guarded("mcrt_simulate_scanlines", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_scene_current(c);
        ensure_workspace(c, 1);
        const int E = c->aq.elements, kl = c->params.psf_lateral;
        // the forward-looking lateral taps of the block's last scanlines reach Kl-1 scanlines to the right:
        // recompute that halo locally instead of exchanging it (Philox keys make it bit-identical)
        const int e0 = first_element;
        const int e1 = (e0 + n_elements + kl - 1 < E) ? e0 + n_elements + kl - 1 : E;
        const int n_local = e1 - e0;
        AcqDev aq = c->aq;
        aq.elements = n_local;
        aq.element_offset = e0;
        cudaStream_t s = c->stream;
        begin_call(c, s);
        c->h_poses[0] = pose_trig(*pose);
        c->h_seed_frame[0] = seed; c->h_seed_frame[1] = frame;
        CUDA_TRY(cudaMemcpyAsync(c->d_poses, c->h_poses, sizeof(PoseTrig), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c->d_seed_frame, c->h_seed_frame, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        FrameDev fr;
        fr.poses = c->d_poses; fr.elem_sincos = c->d_elem_sincos + e0; fr.seed_frame = c->d_seed_frame; fr.n_poses = 1; fr.frame_offset = 0; fr.frame_stride = 1;
        int launches = 0;
        CUDA_TRY(cudaMemsetAsync(c->tb.counters, 0, sizeof(int) * (size_t)kMaxSub * (c->aq.max_depth + 1), s));
        CUDA_TRY(cudaMemsetAsync(c->d_steps, 0, 2 * sizeof(unsigned long long), s));
        launch_trace(c->sc, aq, fr, c->tb, c->sm_count, s, &launches);
        CUDA_TRY(launch_accumulate(c->sc, aq, c->d_volume, c->tb.segments, c->tb.n_segments, 1, c->d_rf_acc, c->d_steps, c->d_columns, s, &launches));
        launch_post(c->d_rf_acc, 1, n_local, aq.rows, c->d_axial, c->params.psf_axial, c->d_lateral, kl, 3, c->d_rf_tmp0, c->d_rf_tmp1,
                    c->d_rf_final, s, &launches, e0, E, c->d_lat_by_row, aq.rf_pitch, c->post_tma ? c->h_axial.data() : nullptr, c->h_lateral.data());
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(rf_out, c->d_rf_final, sizeof(float) * (size_t)n_elements * aq.rows,
                                 is_device_pointer(rf_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        c->stats = mcrt_stats{};
        c->stats.poses = 1; c->stats.kernel_launches = launches;
        c->stats_pending = false;
        return MCRT_OK;
    });
}

int mcrt_trace_debug(mcrt_ctx* c, const mcrt_pose* pose, uint64_t seed, uint64_t frame, mcrt_segment* segments, int32_t* n_segments)
{
    if (!c || !pose || !segments || !n_segments) return fail(MCRT_ERR_INVALID, "mcrt_trace_debug: null argument");
    return guarded("mcrt_trace_debug", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_scene_current(c);
        ensure_workspace(c, 1);
        const size_t n_paths = (size_t)c->aq.elements * c->aq.samples;
        const size_t n_seg = n_paths * c->aq.max_depth;
        if (!c->tb.hit_fraction) { dev_alloc(c->tb.hit_fraction, (size_t)c->cap_poses * n_seg); dev_alloc(c->tb.hit_mesh, (size_t)c->cap_poses * n_seg); }
        begin_call(c, c->stream);
        c->h_poses[0] = pose_trig(*pose);
        c->h_seed_frame[0] = seed; c->h_seed_frame[1] = frame;
        CUDA_TRY(cudaMemcpyAsync(c->d_poses, c->h_poses, sizeof(PoseTrig), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->d_seed_frame, c->h_seed_frame, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        FrameDev fr;
        fr.poses = c->d_poses; fr.elem_sincos = c->d_elem_sincos; fr.seed_frame = c->d_seed_frame; fr.n_poses = 1; fr.frame_offset = 0; fr.frame_stride = 1;
        int launches = 0;
        launch_trace(c->sc, c->aq, fr, c->tb, c->sm_count, c->stream, &launches);
        CUDA_TRY(cudaGetLastError());
        std::vector<DevSegment> hs(n_seg);
        std::vector<float> hf(n_seg);
        std::vector<int32_t> hm(n_seg);
        CUDA_TRY(cudaMemcpyAsync(hs.data(), c->tb.segments, sizeof(DevSegment) * n_seg, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(hf.data(), c->tb.hit_fraction, sizeof(float) * n_seg, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(hm.data(), c->tb.hit_mesh, sizeof(int32_t) * n_seg, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(n_segments, c->tb.n_segments, sizeof(int32_t) * n_paths, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        // the debug arrays make later graph captures carry two extra stores per segment: drop them again
        dev_free(c->tb.hit_fraction); dev_free(c->tb.hit_mesh);
        memset(segments, 0, sizeof(mcrt_segment) * n_seg);
        for (size_t p = 0; p < n_paths; p++)
            for (int k = 0; k < n_segments[p]; k++) {
                const size_t i = p * c->aq.max_depth + k;
                const DevSegment& d = hs[i];
                mcrt_segment& o = segments[i];
                o.from[0] = d.s0.x; o.from[1] = d.s0.y; o.from[2] = d.s0.z; o.reflected_intensity = d.s0.w;
                o.dir[0] = d.s1.x; o.dir[1] = d.s1.y; o.dir[2] = d.s1.z; o.initial_intensity = d.s1.w;
                o.to[0] = d.s2.x; o.to[1] = d.s2.y; o.to[2] = d.s2.z; o.attenuation = d.s2.w;
                const unsigned long long bits = (unsigned long long)(unsigned int)d.s3.x | ((unsigned long long)(unsigned int)d.s3.y << 32);
                memcpy(&o.distance_traveled, &bits, 8);
                o.media_id = d.s3.z; o.tri_id = d.s3.w; o.mesh_id = hm[i]; o.hit_fraction = hf[i];
            }
        return MCRT_OK;
    });
}

int mcrt_device_alloc(int device, size_t bytes, void** dev_ptr)
{
    if (!dev_ptr || bytes == 0) return fail(MCRT_ERR_INVALID, "mcrt_device_alloc: bad argument");
    return guarded("mcrt_device_alloc", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaMalloc(dev_ptr, bytes));
        return MCRT_OK;
    });
}

int mcrt_device_free(int device, void* dev_ptr)
{
    return guarded("mcrt_device_free", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaFree(dev_ptr));
        return MCRT_OK;
    });
}

int mcrt_ipc_export(int device, const void* dev_ptr, unsigned char handle64[64])
{
    if (!dev_ptr || !handle64) return fail(MCRT_ERR_INVALID, "mcrt_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    return guarded("mcrt_ipc_export", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        cudaIpcMemHandle_t h;
        CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
        memcpy(handle64, &h, 64);
        return MCRT_OK;
    });
}

int mcrt_ipc_open(int device, const unsigned char handle64[64], void** peer_ptr)
{
    if (!handle64 || !peer_ptr) return fail(MCRT_ERR_INVALID, "mcrt_ipc_open: null argument");
    return guarded("mcrt_ipc_open", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
        return MCRT_OK;
    });
}

int mcrt_ipc_close(int device, void* peer_ptr)
{
    return guarded("mcrt_ipc_close", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
        return MCRT_OK;
    });
}

int mcrt_copy_async(int device, void* dst, const void* src, size_t bytes, void* cuda_stream)
{
    if (!dst || !src) return fail(MCRT_ERR_INVALID, "mcrt_copy_async: null argument");
    return guarded("mcrt_copy_async", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
        return MCRT_OK;
    });
}

int mcrt_copy2d_async(int device, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t height, void* cuda_stream)
{
    if (!dst || !src) return fail(MCRT_ERR_INVALID, "mcrt_copy2d_async: null argument");
    return guarded("mcrt_copy2d_async", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
        return MCRT_OK;
    });
}

int mcrt_trace_tree_debug(mcrt_ctx* c, const mcrt_pose* pose, uint64_t seed, uint64_t frame, int64_t capacity, mcrt_segment* segments,
                          int32_t* path, int32_t* node, int64_t* n_out)
{
    if (!c || !pose || !segments || !path || !node || !n_out) return fail(MCRT_ERR_INVALID, "mcrt_trace_tree_debug: null argument");
    if (c->tree_budget <= 0) return fail(MCRT_ERR_INVALID, "mcrt_trace_tree_debug: option ray_tree is off");
    return guarded("mcrt_trace_tree_debug", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_scene_current(c);
        ensure_workspace(c, 1);
        cudaStream_t s = c->stream;
        begin_call(c, s);
        c->h_poses[0] = pose_trig(*pose);
        c->h_seed_frame[0] = seed; c->h_seed_frame[1] = frame;
        CUDA_TRY(cudaMemcpyAsync(c->d_poses, c->h_poses, sizeof(PoseTrig), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c->d_seed_frame, c->h_seed_frame, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        FrameDev fr;
        fr.poses = c->d_poses; fr.elem_sincos = c->d_elem_sincos; fr.seed_frame = c->d_seed_frame; fr.n_poses = 1; fr.frame_offset = 0; fr.frame_stride = 1;
        int launches = 0;
        TreeBuffers t = c->tree;
        t.trav_counters = nullptr;
        launch_trace_tree(c->sc, c->aq, fr, t, c->sm_count, s, &launches);
        CUDA_TRY(cudaGetLastError());
        int cnt[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(cnt, t.counters + c->aq.max_depth + 1, sizeof(cnt), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (cnt[1] || cnt[0] > t.seg_capacity) return fail(MCRT_ERR_NOMEM, "ray_tree: the segment budget per path was exceeded; raise option ray_tree");
        *n_out = cnt[0];
        if (cnt[0] > capacity) return fail(MCRT_ERR_INVALID, "mcrt_trace_tree_debug: capacity too small");
        const size_t n = (size_t)cnt[0];
        // the device keeps the segments level by level; this debug hook returns them ordered by (path, node): sorted here, on the host
        std::vector<DevSegment> hs(n);
        std::vector<unsigned long long> keys_raw(n), keys(n);
        std::vector<unsigned> slots(n);
        if (n) {
            CUDA_TRY(cudaMemcpy(hs.data(), t.segments, sizeof(DevSegment) * n, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(keys_raw.data(), t.keys, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
        }
        for (size_t i = 0; i < n; i++) slots[i] = (unsigned)i;
        std::sort(slots.begin(), slots.end(), [&](unsigned a, unsigned b) { return keys_raw[a] < keys_raw[b]; });
        for (size_t i = 0; i < n; i++) keys[i] = keys_raw[slots[i]];
        memset(segments, 0, sizeof(mcrt_segment) * n);
        for (size_t i = 0; i < n; i++) {
            const DevSegment& d = hs[slots[i]];
            mcrt_segment& o = segments[i];
            o.from[0] = d.s0.x; o.from[1] = d.s0.y; o.from[2] = d.s0.z; o.reflected_intensity = d.s0.w;
            o.dir[0] = d.s1.x; o.dir[1] = d.s1.y; o.dir[2] = d.s1.z; o.initial_intensity = d.s1.w;
            o.to[0] = d.s2.x; o.to[1] = d.s2.y; o.to[2] = d.s2.z; o.attenuation = d.s2.w;
            const unsigned long long bits = (unsigned long long)(unsigned int)d.s3.x | ((unsigned long long)(unsigned int)d.s3.y << 32);
            memcpy(&o.distance_traveled, &bits, 8);
            o.media_id = d.s3.z; o.tri_id = d.s3.w; o.mesh_id = -1; o.hit_fraction = 0.0f;
            path[i] = (int32_t)(keys[i] >> 20); node[i] = (int32_t)(keys[i] & 0xfffffULL);
        }
        return MCRT_OK;
    });
}

int mcrt_closest_hit(mcrt_ctx* c, int64_t n, const float* from3, const float* to3, int32_t* tri_id, int32_t* mesh_id, float* fraction,
                     float* point3, float* normal3)
{
    if (!c || n < 0 || !from3 || !to3 || !tri_id || !mesh_id || !fraction || !point3 || !normal3)
        return fail(MCRT_ERR_INVALID, "mcrt_closest_hit: null argument");
    if (n == 0) return MCRT_OK;
    return guarded("mcrt_closest_hit", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_scene_current(c);
        float *d_from = nullptr, *d_to = nullptr, *d_frac = nullptr, *d_pt = nullptr, *d_nr = nullptr;
        int32_t *d_tri = nullptr, *d_mesh = nullptr;
        int rc = MCRT_OK;
        try {
            dev_alloc(d_from, (size_t)n * 3); dev_alloc(d_to, (size_t)n * 3); dev_alloc(d_frac, (size_t)n); dev_alloc(d_pt, (size_t)n * 3);
            dev_alloc(d_nr, (size_t)n * 3); dev_alloc(d_tri, (size_t)n); dev_alloc(d_mesh, (size_t)n);
            CUDA_TRY(cudaMemcpyAsync(d_from, from3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(d_to, to3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
            launch_closest_hit(c->sc, n, d_from, d_to, d_tri, d_mesh, d_frac, d_pt, d_nr, c->stream);
            CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(tri_id, d_tri, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(mesh_id, d_mesh, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(fraction, d_frac, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(point3, d_pt, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(normal3, d_nr, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            // traversal micro-benchmark: device time of the closest-hit kernel alone
            c->stats = mcrt_stats{};
            c->stats.segments = n; c->stats.kernel_launches = 1;
            CUDA_TRY(cudaEventElapsedTime(&c->stats.ms_total, c->ev0, c->ev1));
            c->stats.ms_trace = c->stats.ms_total;
            c->stats_pending = false;
        } catch (...) {
            dev_free(d_from); dev_free(d_to); dev_free(d_frac); dev_free(d_pt); dev_free(d_nr); dev_free(d_tri); dev_free(d_mesh);
            throw;
        }
        dev_free(d_from); dev_free(d_to); dev_free(d_frac); dev_free(d_pt); dev_free(d_nr); dev_free(d_tri); dev_free(d_mesh);
        return rc;
    });
}

int mcrt_transducer_elements(mcrt_ctx* c, const mcrt_pose* pose, float* pos3, float* dir3)
{
    if (!c || !pose || !pos3 || !dir3) return fail(MCRT_ERR_INVALID, "mcrt_transducer_elements: null argument");
    return guarded("mcrt_transducer_elements", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_scene_current(c);
        ensure_workspace(c, 1);
        float *d_pos = nullptr, *d_dir = nullptr;
        const size_t n = (size_t)c->aq.elements * 3;
        dev_alloc(d_pos, n); dev_alloc(d_dir, n);
        begin_call(c, c->stream);
        c->h_poses[0] = pose_trig(*pose);
        cudaError_t e = cudaMemcpyAsync(c->d_poses, c->h_poses, sizeof(PoseTrig), cudaMemcpyHostToDevice, c->stream);
        FrameDev fr;
        fr.poses = c->d_poses; fr.elem_sincos = c->d_elem_sincos; fr.seed_frame = c->d_seed_frame; fr.n_poses = 1; fr.frame_offset = 0; fr.frame_stride = 1;
        if (e == cudaSuccess) { launch_elements(c->aq, fr, d_pos, d_dir, c->stream); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(pos3, d_pos, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dir3, d_dir, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        dev_free(d_pos); dev_free(d_dir);
        CUDA_TRY(e);
        return MCRT_OK;
    });
}

int mcrt_accumulate(mcrt_ctx* c, const mcrt_segment* segments, const int32_t* n_segments, float* rf_out)
{
    if (!c || !segments || !n_segments || !rf_out) return fail(MCRT_ERR_INVALID, "mcrt_accumulate: null argument");
    return guarded("mcrt_accumulate", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_scene_current(c);
        ensure_workspace(c, 1);
        const size_t n_paths = (size_t)c->aq.elements * c->aq.samples;
        const size_t n_seg = n_paths * c->aq.max_depth;
        for (size_t p = 0; p < n_paths; p++)
            if (n_segments[p] < 0 || n_segments[p] > c->aq.max_depth) throw std::invalid_argument("n_segments out of range");
        std::vector<DevSegment> hs(n_seg);
        for (size_t i = 0; i < n_seg; i++) {
            if (segments[i].media_id < 0 || segments[i].media_id >= c->sc.n_mat) {
                if (n_segments[i / c->aq.max_depth] > (int)(i % c->aq.max_depth)) throw std::invalid_argument("segment media_id out of range");
                hs[i] = DevSegment{};
                continue;
            }
            hs[i] = to_dev_segment(segments[i]);
        }
        begin_call(c, c->stream);
        CUDA_TRY(cudaMemcpyAsync(c->tb.segments, hs.data(), sizeof(DevSegment) * n_seg, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->tb.n_segments, n_segments, sizeof(int32_t) * n_paths, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemsetAsync(c->d_steps, 0, 2 * sizeof(unsigned long long), c->stream));
        int launches = 0;
        CUDA_TRY(launch_accumulate(c->sc, c->aq, c->d_volume, c->tb.segments, c->tb.n_segments, 1, c->d_rf_acc, c->d_steps, c->d_columns,
                                   c->stream, &launches));
        CUDA_TRY(cudaMemcpy2DAsync(rf_out, sizeof(float) * c->aq.rows, c->d_rf_acc, sizeof(float) * c->aq.rf_pitch, sizeof(float) * c->aq.rows,
                                   (size_t)c->aq.elements, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->h_steps, c->d_steps, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->stats = mcrt_stats{};
        c->stats.march_steps = (int64_t)c->h_steps[0];
        c->stats.late_echoes = (int64_t)c->h_steps[1];
        c->stats.kernel_launches = launches;
        c->stats_pending = false;
        return MCRT_OK;
    });
}

int mcrt_postprocess(mcrt_ctx* c, const float* rf_in, int32_t cols, int32_t rows, const float* axial, int32_t n_axial, const float* lateral,
                     int32_t n_lateral, int32_t flags, float* rf_out)
{
    if (!c || !rf_in || !rf_out || cols < 1 || rows < 2) return fail(MCRT_ERR_INVALID, "mcrt_postprocess: bad argument");
    if ((flags & 1) && (!axial || !lateral || n_axial < 1 || n_lateral < 1 || n_axial > 255 || n_lateral > 255))
        return fail(MCRT_ERR_INVALID, "mcrt_postprocess: bad taps");
    if (rows > 32768) return fail(MCRT_ERR_INVALID, "mcrt_postprocess: rows > 32768");
    return guarded("mcrt_postprocess", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        const size_t n = (size_t)cols * rows;
        // the shape the fused TMA-staged kernel takes (both passes, 7 x 13 taps, short scanlines) is uploaded with its 16-byte row pitch
        const int pitch = (flags == 3 && c->post_tma) ? post_preferred_pitch(rows, n_axial, n_lateral) : rows;
        float *d_in = nullptr, *d_t0 = nullptr, *d_t1 = nullptr, *d_out = nullptr, *d_ax = nullptr, *d_lat = nullptr;
        try {
            dev_alloc(d_in, (size_t)cols * pitch); dev_alloc(d_t0, n); dev_alloc(d_t1, n); dev_alloc(d_out, n);
            dev_alloc(d_ax, (size_t)(n_axial > 0 ? n_axial : 1)); dev_alloc(d_lat, (size_t)(n_lateral > 0 ? n_lateral : 1));
            if (pitch != rows) CUDA_TRY(cudaMemsetAsync(d_in, 0, sizeof(float) * (size_t)cols * pitch, c->stream));
            CUDA_TRY(cudaMemcpy2DAsync(d_in, sizeof(float) * pitch, rf_in, sizeof(float) * rows, sizeof(float) * rows, (size_t)cols, cudaMemcpyHostToDevice, c->stream));
            if (flags & 1) {
                CUDA_TRY(cudaMemcpyAsync(d_ax, axial, sizeof(float) * n_axial, cudaMemcpyHostToDevice, c->stream));
                CUDA_TRY(cudaMemcpyAsync(d_lat, lateral, sizeof(float) * n_lateral, cudaMemcpyHostToDevice, c->stream));
            }
            int launches = 0;
            launch_post(d_in, 1, cols, rows, d_ax, n_axial, d_lat, n_lateral, flags, d_t0, d_t1, d_out, c->stream, &launches, 0, 0, nullptr, pitch,
                        c->post_tma ? axial : nullptr, lateral);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(rf_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        } catch (...) {
            dev_free(d_in); dev_free(d_t0); dev_free(d_t1); dev_free(d_out); dev_free(d_ax); dev_free(d_lat);
            throw;
        }
        dev_free(d_in); dev_free(d_t0); dev_free(d_t1); dev_free(d_out); dev_free(d_ax); dev_free(d_lat);
        return MCRT_OK;
    });
}

int mcrt_scan_convert(mcrt_ctx* c, const float* rf_in, float* scan_out)
{
    if (!c || !rf_in || !scan_out) return fail(MCRT_ERR_INVALID, "mcrt_scan_convert: null argument");
    return guarded("mcrt_scan_convert", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        ensure_scene_current(c);
        ensure_workspace(c, 1);
        const size_t n = (size_t)c->aq.elements * c->aq.rows;
        const size_t ns = (size_t)c->params.scan_rows * c->params.scan_cols;
        begin_call(c, c->stream);
        CUDA_TRY(cudaMemcpyAsync(c->d_rf_final, rf_in, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
        int launches = 0;
        launch_scan_convert(c->d_rf_final, 1, c->aq.elements, c->aq.rows, c->d_map_x, c->d_map_y, c->params.scan_rows, c->params.scan_cols,
                            c->d_scan, c->stream, &launches);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(scan_out, c->d_scan, sizeof(float) * ns, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return MCRT_OK;
    });
}

int mcrt_set_psf_depth_profile(mcrt_ctx* c, float focus_cm, float spread, float* table_out)
{
    if (!c) return fail(MCRT_ERR_INVALID, "mcrt_set_psf_depth_profile: null argument");
    if (!(spread >= 0.0f) || (spread > 0.0f && !(focus_cm > 0.0f))) return fail(MCRT_ERR_INVALID, "mcrt_set_psf_depth_profile: spread >= 0 and focus_cm > 0");
    return guarded("mcrt_set_psf_depth_profile", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
        c->graphs.clear();
        dev_free(c->d_lat_by_row);
        c->h_lat_by_row.clear();
        if (spread > 0.0f) {
            psf_lateral_depth_table(c->params, c->aq.rows, focus_cm, spread, c->h_lat_by_row);
            dev_alloc(c->d_lat_by_row, c->h_lat_by_row.size());
            CUDA_TRY(cudaMemcpy(c->d_lat_by_row, c->h_lat_by_row.data(), sizeof(float) * c->h_lat_by_row.size(), cudaMemcpyHostToDevice));
            if (table_out) memcpy(table_out, c->h_lat_by_row.data(), sizeof(float) * c->h_lat_by_row.size());
        }
        return MCRT_OK;
    });
}

int mcrt_set_elevation(mcrt_ctx* c, int32_t n_planes, float var_z, float* taps_out, float* z_mm_out)
{
    if (!c) return fail(MCRT_ERR_INVALID, "mcrt_set_elevation: null argument");
    if (n_planes < 1 || n_planes > 63 || (n_planes % 2) == 0) return fail(MCRT_ERR_INVALID, "mcrt_set_elevation: n_planes must be odd, 1 .. 63 (psf.h:31)");
    if (n_planes > 1 && !(var_z > 0.0f)) return fail(MCRT_ERR_INVALID, "mcrt_set_elevation: var_z must be > 0");
    return guarded("mcrt_set_elevation", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        free_workspace(c);                                  // the workspace holds n_planes sub-frames per frame; graphs are dropped with it
        dev_free(c->d_elev_w);
        c->elev_n = n_planes;
        c->h_elev_w.clear(); c->h_elev_z.clear();
        if (n_planes > 1) {
            psf_elevation_taps(c->params, n_planes, var_z, c->h_elev_w, c->h_elev_z);
            dev_alloc(c->d_elev_w, (size_t)n_planes);
            CUDA_TRY(cudaMemcpy(c->d_elev_w, c->h_elev_w.data(), sizeof(float) * n_planes, cudaMemcpyHostToDevice));
            if (taps_out) memcpy(taps_out, c->h_elev_w.data(), sizeof(float) * n_planes);
            if (z_mm_out) memcpy(z_mm_out, c->h_elev_z.data(), sizeof(float) * n_planes);
        }
        return MCRT_OK;
    });
}

int mcrt_elevation_pose(const mcrt_ctx* c, const mcrt_pose* pose, int32_t plane, mcrt_pose* out)
{
    if (!c || !pose || !out) return fail(MCRT_ERR_INVALID, "mcrt_elevation_pose: null argument");
    if (plane < 0 || plane >= c->elev_n || c->elev_n < 2) return fail(MCRT_ERR_INVALID, "mcrt_elevation_pose: no such plane (mcrt_set_elevation first)");
    *out = elevation_pose(*pose, c->h_elev_z[plane]);
    return MCRT_OK;
}

int mcrt_bmode(mcrt_ctx* c, const float* env_in, int32_t n_images, const mcrt_bmode_params* bp, float* compressed_out, uint8_t* bmode8_out)
{
    if (!c || !env_in || !bp || n_images < 1) return fail(MCRT_ERR_INVALID, "mcrt_bmode: null argument");
    if (!(bp->dynamic_range_db > 0.0f)) return fail(MCRT_ERR_INVALID, "mcrt_bmode: dynamic_range_db must be > 0");
    return guarded("mcrt_bmode", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        const int rows = c->aq.rows, cols = c->aq.elements;
        const size_t px = (size_t)rows * cols, ns = (size_t)c->params.scan_rows * c->params.scan_cols;
        // gain table from the shared numerics (so the oracle reproduces it bit for bit): 10^(dB / 20) = exp(dB * ln10 / 20)
        std::vector<float> gain(rows);
        for (int r = 0; r < rows; r++) {
            const double depth_cm = (double)r * c->params.depth_cm / (double)rows;
            gain[r] = (float)mc_exp(((double)bp->gain_db + (double)bp->tgc_db_per_cm * depth_cm) * (2.30258509299404568402 / 20.0));
        }
        cudaStream_t s = c->stream;
        if (n_images > c->bm_cap) {
            CUDA_TRY(cudaStreamSynchronize(s));
            dev_free(c->bm_gain); dev_free(c->bm_env); dev_free(c->bm_cmp); dev_free(c->bm_scan); dev_free(c->bm_q); dev_free(c->bm_max);
            c->bm_cap = 0;
            dev_alloc(c->bm_gain, (size_t)rows); dev_alloc(c->bm_env, px * n_images); dev_alloc(c->bm_cmp, px * n_images);
            dev_alloc(c->bm_scan, ns * n_images); dev_alloc(c->bm_q, ns * n_images); dev_alloc(c->bm_max, (size_t)n_images);
            c->bm_cap = n_images;
        }
        CUDA_TRY(cudaMemcpyAsync(c->bm_gain, gain.data(), sizeof(float) * rows, cudaMemcpyHostToDevice, s));
        const float* src = env_in;
        if (!is_device_pointer(env_in)) {
            CUDA_TRY(cudaMemcpyAsync(c->bm_env, env_in, sizeof(float) * px * n_images, cudaMemcpyHostToDevice, s));
            src = c->bm_env;
        }
        int launches = 0;
        launch_bmode(src, n_images, cols, rows, c->bm_gain, bp->dynamic_range_db, c->bm_cmp, c->bm_max, s, &launches);
        if (compressed_out)
            CUDA_TRY(cudaMemcpyAsync(compressed_out, c->bm_cmp, sizeof(float) * px * n_images,
                                     is_device_pointer(compressed_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        if (bmode8_out) {
            launch_scan_convert(c->bm_cmp, n_images, cols, rows, c->d_map_x, c->d_map_y, c->params.scan_rows, c->params.scan_cols, c->bm_scan, s, &launches);
            launch_quantize8(c->bm_scan, (int64_t)(ns * n_images), c->bm_q, s, &launches);
            CUDA_TRY(cudaMemcpyAsync(bmode8_out, c->bm_q, ns * n_images, is_device_pointer(bmode8_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(s));        // gain.data() is a stack-owned staging buffer: the upload must be complete
        c->stats = mcrt_stats{};
        c->stats.poses = n_images; c->stats.kernel_launches = launches;
        c->stats_pending = false;
        return MCRT_OK;
    });
}

int mcrt_get_psf_taps(const mcrt_ctx* c, float* axial, float* lateral)
{
    if (!c || !axial || !lateral) return fail(MCRT_ERR_INVALID, "mcrt_get_psf_taps: null argument");
    memcpy(axial, c->h_axial.data(), sizeof(float) * c->h_axial.size());
    memcpy(lateral, c->h_lateral.data(), sizeof(float) * c->h_lateral.size());
    return MCRT_OK;
}

int mcrt_get_scene(const mcrt_ctx* c, float* tri_local9, int32_t* tri_mesh, float* mesh_origin3, float* materials8)
{
    if (!c) return fail(MCRT_ERR_INVALID, "mcrt_get_scene: null argument");
    if (tri_local9) memcpy(tri_local9, c->scene.tri_local.data(), sizeof(float) * c->scene.tri_local.size());
    if (tri_mesh) memcpy(tri_mesh, c->scene.tri_mesh.data(), sizeof(int32_t) * c->scene.tri_mesh.size());
    if (mesh_origin3)
        for (size_t m = 0; m < c->scene.meshes.size(); m++)
            for (int a = 0; a < 3; a++) mesh_origin3[3 * m + a] = c->scene.meshes[m].origin[a];
    if (materials8) memcpy(materials8, c->scene.materials.data(), sizeof(HostMaterial) * c->scene.materials.size());
    return MCRT_OK;
}

int mcrt_get_volume(const mcrt_ctx* c, float* out)
{
    if (!c || !out) return fail(MCRT_ERR_INVALID, "mcrt_get_volume: null argument");
    return guarded("mcrt_get_volume", [&]() {
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaMemcpy(out, c->d_volume, sizeof(float) * 2 * (size_t)256 * 256 * 256, cudaMemcpyDeviceToHost));
        return MCRT_OK;
    });
}

int mcrt_load_obj(const char* obj_path, float* out9, int64_t capacity_tris, int64_t* n_tris)
{
    if (!obj_path || !n_tris) return fail(MCRT_ERR_INVALID, "mcrt_load_obj: null argument");
    return guarded("mcrt_load_obj", [&]() {
        std::vector<float> soup;
        load_obj_soup(obj_path, soup);
        const int64_t n = (int64_t)(soup.size() / 9);
        *n_tris = n;
        if (out9 && capacity_tris > 0) memcpy(out9, soup.data(), sizeof(float) * 9 * (size_t)(n < capacity_tris ? n : capacity_tris));
        return MCRT_OK;
    });
}

int mcrt_host_build_sah(const float* tri_local9, const int32_t* tri_mesh, int64_t n_triangles, const float* mesh_origin3, int32_t n_meshes,
                        int32_t threads, float* nodes16, int32_t* slot_triangle, int32_t* max_depth)
{
    if (!tri_local9 || !tri_mesh || !mesh_origin3 || n_triangles < 0 || n_triangles > 0x1fffffff || n_meshes <= 0)
        return fail(MCRT_ERR_INVALID, "mcrt_host_build_sah: bad argument");
    for (int64_t t = 0; t < n_triangles; t++)
        if (tri_mesh[t] < 0 || tri_mesh[t] >= n_meshes) return fail(MCRT_ERR_INVALID, "mcrt_host_build_sah: mesh index out of range");
    return guarded("mcrt_host_build_sah", [&]() {
        HostBvh hb;
        build_sah_bvh(tri_local9, tri_mesh, (int)n_triangles, mesh_origin3, &hb, threads);
        static_assert(sizeof(HostBvhNode) == 64, "16 words per node");
        if (nodes16 && !hb.nodes.empty()) memcpy(nodes16, hb.nodes.data(), sizeof(HostBvhNode) * hb.nodes.size());
        if (slot_triangle) for (size_t k = 0; k < hb.slots.size(); k++) slot_triangle[k] = hb.slots[k].tri;
        if (max_depth) *max_depth = hb.max_depth;
        return MCRT_OK;
    });
}

int mcrt_scene_probe(const char* scene_json_path, int64_t* n_triangles, int32_t* n_meshes, int32_t* n_materials, float* start_pose6)
{
    if (!scene_json_path) return fail(MCRT_ERR_INVALID, "mcrt_scene_probe: null argument");
    return guarded("mcrt_scene_probe", [&]() {
        const HostScene s = load_scene_file(scene_json_path);
        if (n_triangles) *n_triangles = (int64_t)s.tri_mesh.size();
        if (n_meshes) *n_meshes = (int32_t)s.meshes.size();
        if (n_materials) *n_materials = (int32_t)s.materials.size();
        if (start_pose6) memcpy(start_pose6, s.start_pose, sizeof(float) * 6);
        return MCRT_OK;
    });
}

int mcrt_host_tables(const mcrt_params* params, mcrt_info* info, float* elem_sincos2, float* axial, float* lateral, float* map_x,
                     float* map_y)
{
    if (!params) return fail(MCRT_ERR_INVALID, "mcrt_host_tables: null argument");
    return guarded("mcrt_host_tables", [&]() {
        validate_params(*params);
        const Derived d = derive(*params);
        if (info) {
            memset(info, 0, sizeof(*info));
            info->rows = d.rows; info->cols = d.cols; info->scan_rows = params->scan_rows; info->scan_cols = params->scan_cols;
            info->device = -1;
            info->axial_resolution_mm = d.axial_resolution_mm; info->time_step_us = d.time_step_us; info->row_period_us = d.row_period_us;
            info->max_travel_time_us = d.max_travel_time_us;
        }
        if (elem_sincos2) {
            std::vector<float> t;
            element_angle_table(*params, d, t);
            memcpy(elem_sincos2, t.data(), sizeof(float) * t.size());
        }
        if (axial || lateral) {
            std::vector<float> a, l;
            psf_taps(*params, a, l);
            if (axial) memcpy(axial, a.data(), sizeof(float) * a.size());
            if (lateral) memcpy(lateral, l.data(), sizeof(float) * l.size());
        }
        if (map_x || map_y) {
            std::vector<float> mx, my;
            scan_mapping(*params, d, mx, my);
            if (map_x) memcpy(map_x, mx.data(), sizeof(float) * mx.size());
            if (map_y) memcpy(map_y, my.data(), sizeof(float) * my.size());
        }
        return MCRT_OK;
    });
}

int mcrt_numerics_probe(int device, int32_t op, int64_t n, const double* a, const double* b, double* out)
{
    if (n < 0 || !a || !b || !out) return fail(MCRT_ERR_INVALID, "mcrt_numerics_probe: bad argument");
    if (n == 0) return MCRT_OK;
    return guarded("mcrt_numerics_probe", [&]() {
        CUDA_TRY(cudaSetDevice(device));
        double *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
        cudaError_t e = cudaSuccess;
        dev_alloc(d_a, (size_t)n); dev_alloc(d_b, (size_t)n); dev_alloc(d_o, (size_t)n);
        e = cudaMemcpy(d_a, a, sizeof(double) * n, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_b, b, sizeof(double) * n, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { launch_numerics_probe(op, n, d_a, d_b, d_o, 0); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaMemcpy(out, d_o, sizeof(double) * n, cudaMemcpyDeviceToHost);
        dev_free(d_a); dev_free(d_b); dev_free(d_o);
        CUDA_TRY(e);
        return MCRT_OK;
    });
}

}  // extern "C"
