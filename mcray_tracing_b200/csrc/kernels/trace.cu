// trace.cu -- the Monte-Carlo path loop of scene::cast_rays<S,E> (scene.cpp:50-183) as a wavefront:
// one launch per bounce; each launch does generate (bounce 0) / intersect (device BVH instead of
// btCollisionWorld::rayTest) / shade (ray_physics::hit_boundary, ray.cpp:11-97) / compact.  This fork of the
// reference keeps exactly one of {reflection, refraction} per hit (ray.cpp:84-94), so the single-path wavefront
// never grows: compaction only -- warp-aggregated atomic appends, or (large calls) an order-preserving per-chunk
// compaction + k_scan_chunks; bounce 0 is traced once per element by k_first_hit (all samples share the ray).
// The ray-tree mode (k_tree_level) follows both children: the wavefront grows level by level, child rays and
// segments are appended with warp-aggregated atomics and sorted by (path, node) afterwards.

#include "mcrt_device.cuh"
#include "mcrt_launch.h"

namespace mcrt {

namespace {

// 8-wide traversal: perm[oinv][h] = the hit mask h (bit = slot) with every bit moved to position (slot XOR oinv), the
// visiting priority of that slot for a ray of octant oinv (init_trace_kernels fills it; every CTA copies it to shared memory)
__device__ uint4 g_perm8[128];

struct SharedScene {
    float4 mesh_origin[MCRT_MAX_SMEM_MESHES];
    int4 mesh_info[MCRT_MAX_SMEM_MESHES];      // (mat_in, mat_out, vascular, -)
    DevMaterial materials[MCRT_MAX_SMEM_MATERIALS];
#if MCRT_BVH8
    uint4 perm4[128];                          // unsigned char [8][256]
#endif
#if MCRT_SMEM_STACK > 0
    int stack[MCRT_SMEM_STACK * MCRT_TRACE_THREADS];   // [entry][thread]: the top of every thread's traversal stack
#endif
#if MCRT_PACKET
    PacketShared pk;                                   // warp-coherent traversal: one stack and one staged node per warp
#endif
    __device__ __forceinline__ int* stack_column()
    {
#if MCRT_SMEM_STACK > 0
        return stack + threadIdx.x;
#else
        return nullptr;
#endif
    }
    __device__ __forceinline__ const unsigned char* perm() const
    {
#if MCRT_BVH8
        return reinterpret_cast<const unsigned char*>(perm4);
#else
        return nullptr;
#endif
    }
};

__device__ __forceinline__ void load_shared_scene(const SceneDev& sc, SharedScene& sh)
{
#if MCRT_BVH8
    for (int i = threadIdx.x; i < 128; i += blockDim.x) sh.perm4[i] = g_perm8[i];
#endif
    for (int i = threadIdx.x; i < sc.n_mesh && i < MCRT_MAX_SMEM_MESHES; i += blockDim.x) {
        const DevMesh m = sc.meshes[i];
        sh.mesh_origin[i] = make_float4(m.ox, m.oy, m.oz, 0.f);
        sh.mesh_info[i] = make_int4(m.mat_in, m.mat_out, m.vascular, 0);
    }
    for (int i = threadIdx.x; i < sc.n_mat && i < MCRT_MAX_SMEM_MATERIALS; i += blockDim.x) sh.materials[i] = sc.materials[i];
    __syncthreads();
}

// transducer<N>::transducer (transducer.h:45-59): element direction and position for one pose.
__device__ __forceinline__ void element_pose(const AcqDev& aq, const PoseTrigDev& pt, float2 sc_t, float3& pos, float3& dir)
{
    float3 d = make_float3(sc_t.x, sc_t.y, 0.0f);                       // (sin a, cos a, 0)
    d = v_rotate(d, make_float3(0.f, 0.f, 1.f), pt.cz, pt.sz);
    d = v_rotate(d, make_float3(1.f, 0.f, 0.f), pt.cx, pt.sx);
    d = v_rotate(d, make_float3(0.f, 1.f, 0.f), pt.cy, pt.sy);
    dir = d;
    pos = v_add(make_float3(pt.px, pt.py, pt.pz), v_scl(d, aq.radius_f));
}

__device__ __forceinline__ unsigned pack_state(int media, int outside, int depth)
{
    return (unsigned)(media & 0xff) | ((unsigned)((outside + 2) & 0xff) << 8) | ((unsigned)(depth & 0xff) << 16);
}

// Everything scene::cast_rays and ray_physics::hit_boundary compute at a boundary hit (scene.cpp:128-150, ray.cpp:11-97) up
// to, but not including, the choice between the two children: the single-path wavefront keeps one of them
// (ray.cpp:84-94), the ray-tree mode follows both.
struct ShadeResult {
    float3 hit_point, inside_point;
    float back;                    // echo towards the transducer (Eq. 8) * cos(theta')
    float intensity_at_hit;        // after ray_physics::travel
    double distance_after;         // distance_traveled after ray_physics::travel
    float3 refl_dir, refr_dir;
    float i_refl, i_refr;          // child intensities before the epsilon cut
    int mac, mav;                  // medium / outside-medium of the refracted child
    float x;                       // the reflect / refract uniform
};

__device__ __forceinline__ void shade_hit(const SceneDev& sc, const AcqDev& aq, const SharedScene& sh, const float3 from, const float3 dir,
                                          float intensity, const double distance_traveled, const int media, const int outside,
                                          const DevMaterial& med, const HitRec& h, const float3 from_test, const float3 to, const uint64_t seed,
                                          const uint32_t frame, const uint32_t element, const uint32_t sample, const uint32_t rng_index,
                                          ShadeResult& r)
{
    const float frequency = aq.frequency;
    const int4 organ = sh.mesh_info[h.mesh];
    const float3 hit_point = v_interpolate3(from_test, to, h.fraction);      // m_hitPointWorld
    const float3 normal = hit_normal(h);
    const mc_u32x4 b0 = mc_rng_block(seed, frame, element, sample, rng_index, 0);
    // scene.cpp:132-139: penetration q = |N(0, thickness_inside)| (Box-Muller on block 0 words 0,1)
    float q = 0.0f;
    const float thickness = sh.materials[organ.x].thickness;
    if (!aq.deterministic && thickness != 0.0f) {
        const double u1 = mc_u01d(b0.v[0]), u2 = mc_u01d(b0.v[1]);
        double sn, cs;
        mc_sincos(2 * MC_PI_D * u2, &sn, &cs);
        const double z = sqrt(-2.0 * mc_log(u1)) * cs;
        q = (float)fabs(z * (double)thickness);
    }
    const float3 inside_point = v_add(v_scl(dir, q), hit_point);
    // ray_physics::travel (ray.cpp:99-103)
    const double mm = rp_distance_in_mm(sc.spacing, from, inside_point);
    r.distance_after = distance_traveled + mm;
    intensity = intensity * mc_expf(-med.attenuation * ((float)mm * 0.01f) * frequency);
    // medium state machine, ray.cpp:14-47 as it behaves
    int mac, mav;
    if (outside != MCRT_OUTSIDE_NULL) {
        if (organ.z) { mav = MCRT_OUTSIDE_NULL; mac = (outside == MCRT_OUTSIDE_SELF) ? media : outside; }
        else { mav = (outside == organ.x) ? organ.y : organ.x; mac = media; }
    } else {
        if (organ.z) { mav = MCRT_OUTSIDE_SELF; mac = organ.x; }
        else { mav = MCRT_OUTSIDE_NULL; mac = organ.x; }
    }
    const DevMaterial after = sh.materials[mac];
    // ray.cpp:49-50: power-cosine jitter of the normal
    float random_angle = 1.0f;
    float3 random_normal = normal;
    if (!aq.deterministic) {
        random_angle = rp_power_cosine_variate((int)after.shininess, mc_u01d(b0.v[2]));
        bool ok = false;
        for (uint32_t attempt = 0; attempt < MC_RNG_BLOCKS_PER_BOUNCE - 1 && !ok; attempt++) {
            const mc_u32x4 b = mc_rng_block(seed, frame, element, sample, rng_index, 1 + attempt);
            ok = rp_random_unit_vector_attempt(normal, random_angle, mc_u01d(b.v[0]), mc_u01d(b.v[1]), random_normal);
        }
        if (!ok) random_normal = normal;
    }
    // ray.cpp:53-82
    float incidence_angle = v_dot(dir, v_neg(random_normal));
    if (incidence_angle < 0) incidence_angle = v_dot(dir, random_normal);
    const float refr_ratio = med.impedance / after.impedance;
    float refraction_angle = 1 - refr_ratio * refr_ratio * (1 - incidence_angle * incidence_angle);
    const bool total_internal_reflection = refraction_angle < 0;
    refraction_angle = sqrtf(refraction_angle);
    float3 refraction_direction = rp_snells_law(dir, random_normal, incidence_angle, refraction_angle, refr_ratio);
    refraction_direction = v_normalized(refraction_direction);
    float3 reflection_direction = v_add(dir, v_scl(random_normal, 2 * incidence_angle));
    reflection_direction = v_normalized(reflection_direction);
    const float intensity_refl = total_internal_reflection
                                     ? intensity
                                     : rp_reflection_intensity(intensity, med.impedance, incidence_angle, after.impedance, refraction_angle);
    r.i_refl = intensity_refl;
    r.i_refr = intensity - intensity_refl;
    r.back = rp_reflected_intensity_eq8(dir, refraction_direction, reflection_direction, after.specularity) * random_angle;
    r.x = mc_u01f(b0.v[3]);
    r.intensity_at_hit = intensity;
    r.hit_point = hit_point; r.inside_point = inside_point;
    r.refl_dir = reflection_direction; r.refr_dir = refraction_direction;
    r.mac = mac; r.mav = mav;
}

// One bounce of one path.  Returns true if the path survives into the next bounce.
template <bool FIRST>
__device__ __forceinline__ bool bounce_path(const SceneDev& sc, const AcqDev& aq, const FrameDev& fr, const TraceBuffers& tb,
                                            SharedScene& sh, int p, int bounce, int& node_visits, int& tri_tests, bool& reflected,
                                            const unsigned wm /* MCRT_PACKET: the lanes of the warp inside this call */)
{
    const int ES = aq.elements * aq.samples;
    const int pose = p / ES;
    const int rem = p - pose * ES;
    const int element = rem / aq.samples;
    const int sample = rem - element * aq.samples;
    const uint64_t seed = __ldg(&fr.seed_frame[0]);
    const uint32_t frame = (uint32_t)(__ldg(&fr.seed_frame[1]) + ((uint64_t)fr.frame_offset + (uint64_t)pose) * (uint64_t)fr.frame_stride);

    float3 from, dir;
    float intensity;
    double distance_traveled;
    int media, outside;
    if (FIRST) {                                                          // scene.cpp:84-100
        element_pose(aq, fr.poses[pose], __ldg(&fr.elem_sincos[element]), from, dir);
        intensity = 1.0f / (float)(unsigned)aq.samples;
        distance_traveled = 0.0;
        media = sc.starting_material;
        outside = MCRT_OUTSIDE_NULL;
    } else {
        const float4 oi = tb.paths.origin_intensity[p];
        const float4 ds = tb.paths.dir_state[p];
        from = make_float3(oi.x, oi.y, oi.z); intensity = oi.w;
        dir = make_float3(ds.x, ds.y, ds.z);
        const unsigned st = __float_as_uint(ds.w);
        media = (int)(st & 0xff); outside = (int)((st >> 8) & 0xff) - 2;
        distance_traveled = tb.paths.distance[p];
    }
    const float frequency = aq.frequency;
    const DevMaterial med = sh.materials[media];

    // scene.cpp:112-117
    const float r_length = rp_max_ray_length(med.attenuation, intensity, frequency);
    const float3 to = v_add(from, v_scl(make_float3(sc.spacing[0] * dir.x, sc.spacing[1] * dir.y, sc.spacing[2] * dir.z), r_length / 100.0f));
    const float3 from_test = v_add(from, v_scl(dir, 0.1f));
    HitRec h;
    if (FIRST && tb.first_hits) {
        // traced once per element by k_first_hit (same ray for every sample)
        const float4 a = __ldg(&tb.first_hits[2 * (size_t)(p / aq.samples)]), b = __ldg(&tb.first_hits[2 * (size_t)(p / aq.samples) + 1]);
        h.fraction = a.x; h.tri_id = __float_as_int(a.y); h.mesh = __float_as_int(a.z); h.dist_a = a.w;
        h.n_raw = make_float3(b.x, b.y, b.z);
    } else {
#if MCRT_PACKET >= 2
        closest_hit_packet(sc, sh.mesh_origin, sh.pk, wm, from_test, to, h, node_visits, tri_tests);
#else
        closest_hit(sc, sh.mesh_origin, sh.perm(), sh.stack_column(), from_test, to, h, node_visits, tri_tests);
#endif
    }

    DevSegment seg;
    bool alive = false;
    const size_t seg_idx = (size_t)p * aq.max_depth + bounce;
    if (h.tri_id >= 0) {
        ShadeResult r;
        shade_hit(sc, aq, sh, from, dir, intensity, distance_traveled, media, outside, med, h, from_test, to, seed, frame,
                  (uint32_t)(element + aq.element_offset), (uint32_t)sample, (uint32_t)bounce, r);
        // ray.cpp:84-94: keep exactly one branch
        const float reflection_probability = r.i_refl / r.intensity_at_hit;
        float3 ndir;
        float nint;
        int nmedia, noutside;
        if (reflection_probability > r.x) {
            reflected = true;
            ndir = r.refl_dir; nmedia = media; noutside = outside;
            nint = r.i_refl > MCRT_INTENSITY_EPSILON ? r.i_refl : 0.0f;
        } else {
            ndir = r.refr_dir; nmedia = r.mac; noutside = r.mav;
            nint = r.i_refr > MCRT_INTENSITY_EPSILON ? r.i_refr : 0.0f;
        }
        // scene.cpp:148
        seg.s0 = make_float4(from.x, from.y, from.z, r.back);
        seg.s1 = make_float4(dir.x, dir.y, dir.z, intensity);
        seg.s2 = make_float4(r.inside_point.x, r.inside_point.y, r.inside_point.z, med.attenuation);
        seg.s3 = make_int4(__double2loint(distance_traveled), __double2hiint(distance_traveled), media, h.tri_id);
        // scene.cpp:151-157
        alive = nint > MCRT_INTENSITY_EPSILON;
        if (alive && bounce + 1 < aq.max_depth) {
            tb.paths.origin_intensity[p] = make_float4(r.hit_point.x, r.hit_point.y, r.hit_point.z, nint);
            tb.paths.dir_state[p] = make_float4(ndir.x, ndir.y, ndir.z, __uint_as_float(pack_state(nmedia, noutside, bounce + 1)));
            tb.paths.distance[p] = r.distance_after;
        }
    } else {
        // scene.cpp:163-164
        seg.s0 = make_float4(from.x, from.y, from.z, 0.0f);
        seg.s1 = make_float4(dir.x, dir.y, dir.z, intensity);
        seg.s2 = make_float4(to.x, to.y, to.z, med.attenuation);
        seg.s3 = make_int4(__double2loint(distance_traveled), __double2hiint(distance_traveled), media, -1);
    }
    tb.segments[seg_idx] = seg;
    tb.n_segments[p] = bounce + 1;
    if (tb.hit_fraction) tb.hit_fraction[seg_idx] = h.fraction;
    if (tb.hit_mesh) tb.hit_mesh[seg_idx] = h.mesh;
    return alive;
}

#ifndef MCRT_BOUNCE_GRID_CTAS_PER_SM
#define MCRT_BOUNCE_GRID_CTAS_PER_SM 64   // measured: an (effectively) uncapped grid beats a persistent 8-CTA/SM grid by 15 % at 256 frames
#endif
#ifndef MCRT_BOUNCE_MIN_CTAS
#define MCRT_BOUNCE_MIN_CTAS 8      // 64 registers, 32 warps/SM.  Round 1 measured 6 (80 registers) best; with the round-2 kernel 8 is 1-2 % faster at
                                    // 1024 frames, one frame and on config 4, 0.5 % slower at 64 frames; 10 (48 registers) -11 % (profiles/r02ap_ab_bounce_regs.txt)
#endif
// ORDERED: order-preserving compaction (see TraceBuffers::warp_counts): every warp compacts its survivors into its own
// 32-slot piece of the sparse queue and records how many, k_compact turns that into the dense queue of the next bounce.
// Otherwise survivors are appended to the next queue with one warp-aggregated atomicAdd per warp.
template <bool FIRST, bool ORDERED>
__global__ void __launch_bounds__(128, MCRT_BOUNCE_MIN_CTAS) k_bounce(const SceneDev sc, const AcqDev aq, const FrameDev fr, const TraceBuffers tb, const int bounce)
{
    __shared__ SharedScene sh;
    load_shared_scene(sc, sh);
    // ORDERED: queue_a is always the dense input queue, queue_b the sparse output (k_compact runs between the bounces)
    const int* __restrict__ qin = ORDERED ? tb.queue_a : ((bounce & 1) ? tb.queue_b : tb.queue_a);
    int* __restrict__ qout = ORDERED ? tb.queue_b : ((bounce & 1) ? tb.queue_a : tb.queue_b);
    // Tail merge: the late bounces have few live paths and each launch costs the latency of one full bounce (~40 us) however
    // few they are.  Once at most tail_threshold paths (one resident wave) are alive, THIS launch walks each of them to its
    // end in-thread (no compaction between the merged bounces) and the remaining bounce launches return immediately.
    const int tail_mark = FIRST ? 0 : tb.counters[aq.max_depth];          // (bounce at which the tail started) + 1, 0 = not yet
    if (tail_mark && bounce >= tail_mark) return;
    const int n_in = FIRST ? fr.n_poses * aq.elements * aq.samples : tb.counters[bounce];
    const bool tail = tb.tail_threshold > 0 && n_in <= tb.tail_threshold && bounce + 1 < aq.max_depth;
    if (tail && blockIdx.x == 0 && threadIdx.x == 0) tb.counters[aq.max_depth] = bounce + 1;
    const int n_round = (n_in + 31) & ~31;
    const unsigned lane = threadIdx.x & 31;
    const bool last = bounce + 1 >= aq.max_depth;
    int node_visits = 0, tri_tests = 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += gridDim.x * blockDim.x) {
        int p = -1;
        bool alive = false, reflected = false;
#if MCRT_PACKET >= 2
        unsigned wm = __ballot_sync(0xffffffffu, idx < n_in);
#else
        const unsigned wm = 0u;
#endif
        if (idx < n_in) {
            p = FIRST ? idx : qin[idx];
            if (FIRST) alive = bounce_path<true>(sc, aq, fr, tb, sh, p, bounce, node_visits, tri_tests, reflected, wm);
            else alive = true;                                     // traced by the loop below (ONE inlined copy of bounce_path<false>)
        }
        if (!FIRST || tail) {
            // bounce `bounce` (non-first kernels) and, in tail mode, every later bounce of the path; counters[b] still receives
            // the number of paths entering bounce b (mcrt_stats.segments).  The body is shared so that the kernel holds a single
            // copy of the ~50 KB bounce code (instruction-cache footprint, profiles/r01_traversal_ab.txt).
            for (int b = FIRST ? bounce + 1 : bounce; b < aq.max_depth; b++) {
                if (b > bounce) {
                    const unsigned ma = __ballot_sync(0xffffffffu, alive);
                    if (!ma) break;
                    if ((int)lane == __ffs(ma) - 1) atomicAdd(&tb.counters[b], __popc(ma));
#if MCRT_PACKET >= 2
                    wm = ma;
#endif
                }
                if (alive) { reflected = false; alive = bounce_path<false>(sc, aq, fr, tb, sh, p, b, node_visits, tri_tests, reflected, wm); }
                if (!tail) break;
            }
            if (tail) continue;                                     // `tail` is uniform over the launch: no barrier is skipped by part of a CTA
        }
        if (last) continue;
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (ORDERED) {
            // the warp's 32 paths of this iteration are one chunk (idx is warp-aligned): compact them into the chunk's own
            // 32 slots, in order -- no CTA barrier, nobody waits for a slower warp
            // Option group_histories (off): survivors that were REFRACTED go first, the REFLECTED ones behind them (both in order) and
            // k_compact places all refracted survivors of the launch before all reflected ones; applied at every bounce this keeps
            // paths with the same reflect / refract history together.  Measured 0.5 % (ircad11) to 1.6 % (config 4) SLOWER than plain
            // (pose, element, sample) order: reflections are rare, and the reflected rays gathered from many elements are less coherent
            // among themselves than next to their refracted siblings (profiles/r02ab_ab_group_histories.txt).
            const int chunk = idx >> 5;
            const unsigned lt = (1u << lane) - 1u;
            const unsigned mr = tb.group_histories ? __ballot_sync(0xffffffffu, alive && reflected) : 0u;
            const unsigned mt = m & ~mr;
            if (alive) qout[chunk * 32 + ((mr >> lane) & 1u ? __popc(mt) + __popc(mr & lt) : __popc(mt & lt))] = p;
            if (lane == 0) {
                tb.warp_counts[chunk] = __popc(mt) | (__popc(mr) << 8);
                if (mt) atomicAdd(&tb.tile_counts[((size_t)bounce * 2) * tb.n_tiles + (chunk >> 8)], __popc(mt));
                if (mr) atomicAdd(&tb.tile_counts[((size_t)bounce * 2 + 1) * tb.n_tiles + (chunk >> 8)], __popc(mr));
            }
        } else if (m) {
            // compact: warp-aggregated queue append (one atomic per warp)
            const int leader = __ffs(m) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(&tb.counters[bounce + 1], __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (alive) {
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                qout[pos] = p;
            }
        }
    }
    if (tb.trav_counters) {      // optional work counters (mcrt_set_option "count_traversal")
        for (int off = 16; off > 0; off >>= 1) {
            node_visits += __shfl_xor_sync(0xffffffffu, node_visits, off);
            tri_tests += __shfl_xor_sync(0xffffffffu, tri_tests, off);
        }
        if (lane == 0) {
            atomicAdd(&tb.trav_counters[0], (unsigned long long)node_visits);
            atomicAdd(&tb.trav_counters[1], (unsigned long long)tri_tests);
        }
    }
}

// Bounce 0 for one (pose, element): every sample of the element starts on this ray with intensity 1 / samples in the
// starting medium (scene.cpp:84-100), so its closest hit is found once here; k_bounce<true> then shades each sample.
__global__ void __launch_bounds__(128, MCRT_BOUNCE_MIN_CTAS) k_first_hit(const SceneDev sc, const AcqDev aq, const FrameDev fr, const TraceBuffers tb)
{
    __shared__ SharedScene sh;
    load_shared_scene(sc, sh);
    const int n = fr.n_poses * aq.elements;
    int node_visits = 0, tri_tests = 0;
#if MCRT_PACKET >= 1
    // warp-uniform trip count: the whole warp establishes the mask of the lanes that trace
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31) & ~31); i += gridDim.x * blockDim.x) {
        const unsigned wm = __ballot_sync(0xffffffffu, i < n);
        if (i >= n) continue;
#else
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#endif
        const int pose = i / aq.elements, element = i - pose * aq.elements;
        float3 from, dir;
        element_pose(aq, fr.poses[pose], __ldg(&fr.elem_sincos[element]), from, dir);
        const float intensity = 1.0f / (float)(unsigned)aq.samples;
        const DevMaterial med = sh.materials[sc.starting_material];
        const float r_length = rp_max_ray_length(med.attenuation, intensity, aq.frequency);
        const float3 to = v_add(from, v_scl(make_float3(sc.spacing[0] * dir.x, sc.spacing[1] * dir.y, sc.spacing[2] * dir.z), r_length / 100.0f));
        const float3 from_test = v_add(from, v_scl(dir, 0.1f));
        HitRec h;
#if MCRT_PACKET >= 1
        closest_hit_packet(sc, sh.mesh_origin, sh.pk, wm, from_test, to, h, node_visits, tri_tests);
#else
        closest_hit(sc, sh.mesh_origin, sh.perm(), sh.stack_column(), from_test, to, h, node_visits, tri_tests);
#endif
        tb.first_hits[2 * (size_t)i] = make_float4(h.fraction, __int_as_float(h.tri_id), __int_as_float(h.mesh), h.dist_a);
        tb.first_hits[2 * (size_t)i + 1] = make_float4(h.n_raw.x, h.n_raw.y, h.n_raw.z, 0.0f);
    }
    if (tb.trav_counters) {
        for (int off = 16; off > 0; off >>= 1) {
            node_visits += __shfl_xor_sync(0xffffffffu, node_visits, off);
            tri_tests += __shfl_xor_sync(0xffffffffu, tri_tests, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&tb.trav_counters[0], (unsigned long long)node_visits);
            atomicAdd(&tb.trav_counters[1], (unsigned long long)tri_tests);
        }
    }
}

// ORDERED compaction, between bounce b and b + 1: sparse queue (32 slots per warp chunk: the chunk's refracted survivors, then its
// reflected ones; warp_counts[chunk] = refracted | reflected << 8) -> dense queue: ALL refracted survivors in (pose, element, sample)
// order, then all reflected ones.  A tile = 256 consecutive warp chunks = one CTA.  A tile's bases are sums of the per-tile counts
// k_bounce accumulated (tile_counts[0][..] refracted, [1][..] reflected; one atomic per warp chunk and kind), so there is no scan
// pass and no look-back chain: every CTA is independent.  counters[b + 1] receives the number of survivors.
__global__ void __launch_bounds__(256) k_compact(const int* __restrict__ sparse, int* __restrict__ dense, const int* __restrict__ warp_counts,
                                                const int* __restrict__ tile_counts, const int n_tiles, int* __restrict__ counters, const int bounce,
                                                const int n_paths_first, const int tail_index)
{
    __shared__ int s_red[3][8];
    __shared__ int s_scan[2][8];
    // after the tail merge (k_bounce) the bounces >= tail_from do not compact and keep counters[] themselves
    const int tail_mark = counters[tail_index];                           // (bounce at which the tail started) + 1
    if (tail_mark && bounce + 1 >= tail_mark) return;
    const int n_in = bounce == 0 ? n_paths_first : counters[bounce];
    const int n_chunks = (n_in + 31) >> 5;
    const int tile = blockIdx.x;
    if (tile * 256 >= n_chunks) return;
    const int tiles_used = (n_chunks + 255) >> 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // bases of this tile: refracted before it, reflected before it, all refracted
    int part_t = 0, part_r = 0, all_t = 0;
    for (int i = threadIdx.x; i < tiles_used; i += 256) {
        const int t = __ldg(&tile_counts[i]);
        all_t += t;
        if (i < tile) { part_t += t; part_r += __ldg(&tile_counts[n_tiles + i]); }
    }
    for (int off = 16; off > 0; off >>= 1) {
        part_t += __shfl_xor_sync(0xffffffffu, part_t, off); part_r += __shfl_xor_sync(0xffffffffu, part_r, off);
        all_t += __shfl_xor_sync(0xffffffffu, all_t, off);
    }
    if (lane == 0) { s_red[0][warp] = part_t; s_red[1][warp] = part_r; s_red[2][warp] = all_t; }
    // exclusive scans of the tile's 256 chunk counts
    const int chunk = tile * 256 + threadIdx.x;
    const int cnt = chunk < n_chunks ? __ldg(&warp_counts[chunk]) : 0;
    const int ct = cnt & 0xff, cr = cnt >> 8;
    int incl_t = ct, incl_r = cr;
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl_t, off), z = __shfl_up_sync(0xffffffffu, incl_r, off);
        if (lane >= off) { incl_t += y; incl_r += z; }
    }
    if (lane == 31) { s_scan[0][warp] = incl_t; s_scan[1][warp] = incl_r; }
    __syncthreads();
    int base_t = 0, base_r = 0, total_t = 0, before_t = 0, before_r = 0, sum_t = 0, sum_r = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        base_t += s_red[0][w]; base_r += s_red[1][w]; total_t += s_red[2][w];
        const int t = s_scan[0][w], r = s_scan[1][w];
        if (w < warp) { before_t += t; before_r += r; }
        sum_t += t; sum_r += r;
    }
    const int excl_t = base_t + before_t + incl_t - ct;
    const int excl_r = total_t + base_r + before_r + incl_r - cr;
    // gather: the warp walks its 32 chunks, lanes copy the chunk's survivors (coalesced on both sides)
#pragma unroll 4
    for (int c = 0; c < 32; c++) {
        const int n_t = __shfl_sync(0xffffffffu, ct, c), n_r = __shfl_sync(0xffffffffu, cr, c);
        const int o_t = __shfl_sync(0xffffffffu, excl_t, c), o_r = __shfl_sync(0xffffffffu, excl_r, c);
        if (lane < n_t + n_r) {
            const int v = __ldg(&sparse[(size_t)(tile * 256 + warp * 32 + c) * 32 + lane]);
            dense[lane < n_t ? o_t + lane : o_r + lane - n_t] = v;
        }
    }
    if (threadIdx.x == 0 && sum_t + sum_r) atomicAdd(&counters[bounce + 1], sum_t + sum_r);
}

#ifndef MCRT_CH_MIN_CTAS
#define MCRT_CH_MIN_CTAS 1
#endif
// ------------------------------------------------------------------------------------------------
// Ray-tree mode: one launch per tree level, everything ORDERED (see TreeBuffers): ray idx of the level's dense queue emits segment
// seg_base + idx; the children of a warp's 32 rays go, in ray order and reflected before refracted, into the warp's own 64 slots of
// the sparse output pool; k_compact_tree packs the slot indices into the next level's queue.  No atomics on the data path besides
// one add per warp into its tile's child count, no sort.
// ------------------------------------------------------------------------------------------------
template <bool FIRST>
__global__ void __launch_bounds__(128, 5) k_tree_level(const SceneDev sc, const AcqDev aq, const FrameDev fr, const TreeBuffers tb, const int level)
{
    __shared__ SharedScene sh;
    load_shared_scene(sc, sh);
    const TreeRay* __restrict__ rin = (level & 1) ? tb.rays_b : tb.rays_a;
    TreeRay* __restrict__ rout = (level & 1) ? tb.rays_a : tb.rays_b;
    const int* __restrict__ qin = (level & 1) ? tb.queue_b : tb.queue_a;
    const int ES = aq.elements * aq.samples;
    const int n_paths = fr.n_poses * ES;
    // rays entering this level and the first segment slot of the level
    int n_in = FIRST ? n_paths : tb.counters[level];
    bool overflow = false;
    if (n_in > tb.ray_capacity) { n_in = tb.ray_capacity; overflow = true; }
    int seg_base = 0;
    for (int l = 0; l < level; l++) { const int c = l == 0 ? n_paths : tb.counters[l]; seg_base += c < tb.ray_capacity ? c : tb.ray_capacity; }
    if (seg_base + n_in > tb.seg_capacity) { n_in = tb.seg_capacity - seg_base > 0 ? tb.seg_capacity - seg_base : 0; overflow = true; }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        tb.counters[aq.max_depth + 1] = seg_base + n_in;
        if (overflow) tb.counters[aq.max_depth + 2] = 1;
    }
    const int n_round = (n_in + 31) & ~31;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint64_t seed = __ldg(&fr.seed_frame[0]);
    int* __restrict__ lvl_first = tb.level_first + (size_t)level * tb.n_scanlines;
    int* __restrict__ lvl_end = tb.level_end + (size_t)level * tb.n_scanlines;
    int node_visits = 0, tri_tests = 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += gridDim.x * blockDim.x) {
        const bool valid = idx < n_in;
        bool c_refl = false, c_refr = false;
        TreeRay kid_refl, kid_refr;
        int scanline = -2;
        if (valid) {
            int path, node, media, outside, depth;
            float3 from, dir;
            float intensity;
            double distance_traveled;
            if (FIRST) {
                path = idx; node = 1; depth = 0;
                const int pose = path / ES, element = (path - pose * ES) / aq.samples;
                element_pose(aq, fr.poses[pose], __ldg(&fr.elem_sincos[element]), from, dir);
                intensity = 1.0f / (float)(unsigned)aq.samples;
                distance_traveled = 0.0;
                media = sc.starting_material; outside = MCRT_OUTSIDE_NULL;
            } else {
                const TreeRay r = rin[qin[idx]];
                path = r.path; node = r.node;
                from = make_float3(r.origin_intensity.x, r.origin_intensity.y, r.origin_intensity.z); intensity = r.origin_intensity.w;
                dir = make_float3(r.dir_state.x, r.dir_state.y, r.dir_state.z);
                const unsigned st = __float_as_uint(r.dir_state.w);
                media = (int)(st & 0xff); outside = (int)((st >> 8) & 0xff) - 2; depth = (int)((st >> 16) & 0xff);
                distance_traveled = r.distance;
            }
            const int pose = path / ES;
            const int rem = path - pose * ES;
            const int element = rem / aq.samples, sample = rem - element * aq.samples;
            scanline = pose * aq.elements + element;
            const uint32_t frame = (uint32_t)(__ldg(&fr.seed_frame[1]) + ((uint64_t)fr.frame_offset + (uint64_t)pose) * (uint64_t)fr.frame_stride);
            const DevMaterial med = sh.materials[media];
            const float r_length = rp_max_ray_length(med.attenuation, intensity, aq.frequency);
            const float3 to = v_add(from, v_scl(make_float3(sc.spacing[0] * dir.x, sc.spacing[1] * dir.y, sc.spacing[2] * dir.z), r_length / 100.0f));
            const float3 from_test = v_add(from, v_scl(dir, 0.1f));
            HitRec h;
            closest_hit(sc, sh.mesh_origin, sh.perm(), sh.stack_column(), from_test, to, h, node_visits, tri_tests);
            DevSegment seg;
            if (h.tri_id >= 0) {
                ShadeResult r;
                // the Philox counter takes the node id where the single-path mode puts the bounce index
                shade_hit(sc, aq, sh, from, dir, intensity, distance_traveled, media, outside, med, h, from_test, to, seed, frame,
                          (uint32_t)(element + aq.element_offset), (uint32_t)sample, (uint32_t)node, r);
                seg.s0 = make_float4(from.x, from.y, from.z, r.back);
                seg.s1 = make_float4(dir.x, dir.y, dir.z, intensity);
                seg.s2 = make_float4(r.inside_point.x, r.inside_point.y, r.inside_point.z, med.attenuation);
                seg.s3 = make_int4(__double2loint(distance_traveled), __double2hiint(distance_traveled), media, h.tri_id);
                if (depth + 1 < aq.max_depth) {
                    c_refl = r.i_refl > MCRT_INTENSITY_EPSILON;
                    c_refr = r.i_refr > MCRT_INTENSITY_EPSILON;
                    kid_refl.origin_intensity = make_float4(r.hit_point.x, r.hit_point.y, r.hit_point.z, r.i_refl);
                    kid_refl.dir_state = make_float4(r.refl_dir.x, r.refl_dir.y, r.refl_dir.z, __uint_as_float(pack_state(media, outside, depth + 1)));
                    kid_refl.distance = r.distance_after; kid_refl.path = path; kid_refl.node = 2 * node;
                    kid_refr.origin_intensity = make_float4(r.hit_point.x, r.hit_point.y, r.hit_point.z, r.i_refr);
                    kid_refr.dir_state = make_float4(r.refr_dir.x, r.refr_dir.y, r.refr_dir.z, __uint_as_float(pack_state(r.mac, r.mav, depth + 1)));
                    kid_refr.distance = r.distance_after; kid_refr.path = path; kid_refr.node = 2 * node + 1;
                }
            } else {
                seg.s0 = make_float4(from.x, from.y, from.z, 0.0f);
                seg.s1 = make_float4(dir.x, dir.y, dir.z, intensity);
                seg.s2 = make_float4(to.x, to.y, to.z, med.attenuation);
                seg.s3 = make_int4(__double2loint(distance_traveled), __double2hiint(distance_traveled), media, -1);
            }
            tb.segments[seg_base + idx] = seg;
            tb.keys[seg_base + idx] = ((unsigned long long)(unsigned)path << 20) | (unsigned long long)(unsigned)node;
        }
        // the scanline's segment range at this level: rays are in (path, node) order, so a scanline's rays are consecutive
        {
            int prev = __shfl_up_sync(0xffffffffu, scanline, 1), next = __shfl_down_sync(0xffffffffu, scanline, 1);
            if (valid) {
                if (lane == 0) prev = idx == 0 ? -1 : (FIRST ? (idx - 1) / aq.samples : rin[qin[idx - 1]].path / aq.samples);
                if (lane == 31) next = idx + 1 >= n_in ? -1 : (FIRST ? (idx + 1) / aq.samples : rin[qin[idx + 1]].path / aq.samples);
                if (prev != scanline) lvl_first[scanline] = seg_base + idx;
                if (next != scanline) lvl_end[scanline] = seg_base + idx + 1;
            }
        }
        // children -> this warp chunk's 64 slots of the sparse pool, in ray order, reflected before refracted
        const unsigned m1 = __ballot_sync(0xffffffffu, c_refl), m2 = __ballot_sync(0xffffffffu, c_refr);
        const int chunk = idx >> 5;
        const int before = __popc(m1 & lt) + __popc(m2 & lt);
        if (c_refl) rout[(size_t)chunk * 64 + before] = kid_refl;
        if (c_refr) rout[(size_t)chunk * 64 + before + (c_refl ? 1 : 0)] = kid_refr;
        if (lane == 0) {
            const int n = __popc(m1) + __popc(m2);
            tb.warp_counts[chunk] = n;
            if (n) atomicAdd(&tb.tile_counts[(size_t)level * tb.n_tiles + (chunk >> 8)], n);
        }
    }
    if (tb.trav_counters) {
        for (int off = 16; off > 0; off >>= 1) {
            node_visits += __shfl_xor_sync(0xffffffffu, node_visits, off);
            tri_tests += __shfl_xor_sync(0xffffffffu, tri_tests, off);
        }
        if (lane == 0) {
            atomicAdd(&tb.trav_counters[0], (unsigned long long)node_visits);
            atomicAdd(&tb.trav_counters[1], (unsigned long long)tri_tests);
        }
    }
}

// sparse pool of level `level` (64 slots per warp chunk, warp_counts[chunk] of them used) -> dense queue of level + 1 (the slot
// indices, in order).  One CTA per tile of 256 warp chunks; the tile's base is the sum of the tile counts before it (k_compact).
__global__ void __launch_bounds__(256) k_compact_tree(int* __restrict__ dense, const int* __restrict__ warp_counts, const int* __restrict__ tile_counts,
                                                     int* __restrict__ counters, const int level, const int n_paths_first, const int ray_capacity)
{
    __shared__ int s_red[8];
    __shared__ int s_scan[8];
    int n_in = level == 0 ? n_paths_first : counters[level];
    if (n_in > ray_capacity) n_in = ray_capacity;
    const int n_chunks = (n_in + 31) >> 5;
    const int tile = blockIdx.x;
    if (tile * 256 >= n_chunks) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int part = 0;
    for (int i = threadIdx.x; i < tile; i += 256) part += __ldg(&tile_counts[i]);
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) s_red[warp] = part;
    const int chunk = tile * 256 + threadIdx.x;
    const int cnt = chunk < n_chunks ? __ldg(&warp_counts[chunk]) : 0;
    int incl = cnt;
    for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += y; }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int base = 0, before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { base += s_red[w]; const int t = s_scan[w]; if (w < warp) before += t; total += t; }
    const int excl = base + before + incl - cnt;
#pragma unroll 4
    for (int c = 0; c < 32; c++) {
        const int n_c = __shfl_sync(0xffffffffu, cnt, c);
        const int o_c = __shfl_sync(0xffffffffu, excl, c);
        const int slot0 = (tile * 256 + warp * 32 + c) * 64;
        if (lane < n_c && o_c + lane < ray_capacity) dense[o_c + lane] = slot0 + lane;
        if (lane + 32 < n_c && o_c + lane + 32 < ray_capacity) dense[o_c + lane + 32] = slot0 + lane + 32;
    }
    if (threadIdx.x == 0 && total) atomicAdd(&counters[level + 1], total);
}

__global__ void __launch_bounds__(128, MCRT_CH_MIN_CTAS) k_closest_hit(const SceneDev sc, const int64_t n, const float* __restrict__ from3,
                                                    const float* __restrict__ to3, int32_t* __restrict__ tri, int32_t* __restrict__ mesh,
                                                    float* __restrict__ frac, float* __restrict__ point3, float* __restrict__ normal3)
{
    __shared__ SharedScene sh;
    load_shared_scene(sc, sh);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float3 f = make_float3(from3[3 * i], from3[3 * i + 1], from3[3 * i + 2]);
        const float3 t = make_float3(to3[3 * i], to3[3 * i + 1], to3[3 * i + 2]);
        HitRec h;
        int nv = 0, nt = 0;
        closest_hit(sc, sh.mesh_origin, sh.perm(), sh.stack_column(), f, t, h, nv, nt);
        tri[i] = h.tri_id;
        mesh[i] = h.mesh;
        frac[i] = h.fraction;
        const float3 pt = v_interpolate3(f, t, h.fraction);
        float3 nr = make_float3(0.f, 0.f, 0.f);
        if (h.tri_id >= 0) nr = hit_normal(h);
        point3[3 * i] = pt.x; point3[3 * i + 1] = pt.y; point3[3 * i + 2] = pt.z;
        normal3[3 * i] = nr.x; normal3[3 * i + 1] = nr.y; normal3[3 * i + 2] = nr.z;
    }
}

__global__ void k_elements(const AcqDev aq, const FrameDev fr, float* __restrict__ pos3, float* __restrict__ dir3)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fr.n_poses * aq.elements) return;
    const int pose = i / aq.elements, e = i - pose * aq.elements;
    float3 pos, dir;
    element_pose(aq, fr.poses[pose], fr.elem_sincos[e], pos, dir);
    pos3[3 * i] = pos.x; pos3[3 * i + 1] = pos.y; pos3[3 * i + 2] = pos.z;
    dir3[3 * i] = dir.x; dir3[3 * i + 1] = dir.y; dir3[3 * i + 2] = dir.z;
}

__global__ void k_numerics_probe(const int op, const int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                 double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = 0.0;
    switch (op) {
        case 0: r = (double)mc_expf((float)a[i]); break;
        case 1: r = (double)mc_logf((float)a[i]); break;
        case 2: r = (double)mc_powf((float)a[i], (float)b[i]); break;
        case 3: { double s, c; mc_sincos(a[i], &s, &c); r = s; break; }
        case 4: { double s, c; mc_sincos(a[i], &s, &c); r = c; break; }
        case 5: {
            const mc_u32x4 w = mc_rng_block(0x0123456789abcdefULL, (uint32_t)a[i], (uint32_t)b[i], 3u, 2u, 1u);
            r = (double)w.v[0] + 4294967296.0 * (double)(w.v[3] & 0xfffffu);
            break;
        }
        case 6: r = mc_pow(a[i], b[i]); break;
        case 7: r = mc_exp(a[i]); break;
        case 8: r = mc_log(a[i]); break;
        default: break;
    }
    out[i] = r;
}

}  // namespace

static int grid_for(int64_t n_threads, int block, int sm_count, int ctas_per_sm)
{
    int64_t g = (n_threads + block - 1) / block;
    const int64_t cap = (int64_t)sm_count * ctas_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t init_trace_kernels()
{
    unsigned char lut[8][256];
    for (unsigned o = 0; o < 8; o++)
        for (unsigned h = 0; h < 256; h++) {
            unsigned p = 0;
            for (unsigned sl = 0; sl < 8; sl++)
                if (h & (1u << sl)) p |= 1u << (sl ^ o);
            lut[o][h] = (unsigned char)p;
        }
    return cudaMemcpyToSymbol(g_perm8, lut, sizeof(lut));
}

void launch_trace(const SceneDev& sc, const AcqDev& aq, const FrameDev& fr, const TraceBuffers& tb, int sm_count, cudaStream_t stream,
                  int* launches)
{
    const int64_t n_paths = (int64_t)fr.n_poses * aq.elements * aq.samples;
    // counters[0] is informational; counters[1..] are the compaction cursors
    cudaMemsetAsync(tb.counters, 0, sizeof(int) * (size_t)(aq.max_depth + 1), stream);
    if (tb.warp_counts) cudaMemsetAsync(tb.tile_counts, 0, sizeof(int) * 2 * (size_t)aq.max_depth * tb.n_tiles, stream);
    const int block = 128;
    // persistent-style grid: a multiple of the SM count, grid-stride loop inside
    const int grid = grid_for(n_paths, block, sm_count, MCRT_BOUNCE_GRID_CTAS_PER_SM);
    if (tb.first_hits) {
        const int n_elem = fr.n_poses * aq.elements;
        k_first_hit<<<grid_for(n_elem, block, sm_count, MCRT_BOUNCE_GRID_CTAS_PER_SM), block, 0, stream>>>(sc, aq, fr, tb);
        if (launches) (*launches)++;
    }
    for (int b = 0; b < aq.max_depth; b++) {
        const bool ordered = tb.warp_counts != nullptr;
        if (ordered) {
            if (b == 0) k_bounce<true, true><<<grid, block, 0, stream>>>(sc, aq, fr, tb, b);
            else k_bounce<false, true><<<grid, block, 0, stream>>>(sc, aq, fr, tb, b);
            if (b + 1 < aq.max_depth) {
                k_compact<<<tb.n_tiles, 256, 0, stream>>>(tb.queue_b, tb.queue_a, tb.warp_counts, tb.tile_counts + (size_t)b * 2 * tb.n_tiles, tb.n_tiles,
                                                          tb.counters, b, (int)n_paths, aq.max_depth);
                if (launches) (*launches)++;
            }
        } else if (b == 0) k_bounce<true, false><<<grid, block, 0, stream>>>(sc, aq, fr, tb, b);
        else k_bounce<false, false><<<grid, block, 0, stream>>>(sc, aq, fr, tb, b);
        if (launches) (*launches)++;
    }
}

void launch_trace_tree(const SceneDev& sc, const AcqDev& aq, const FrameDev& fr, const TreeBuffers& tb, int sm_count, cudaStream_t stream,
                       int* launches)
{
    const int64_t n_paths = (int64_t)fr.n_poses * aq.elements * aq.samples;
    cudaMemsetAsync(tb.counters, 0, sizeof(int) * (size_t)(aq.max_depth + 3), stream);
    cudaMemsetAsync(tb.tile_counts, 0, sizeof(int) * (size_t)aq.max_depth * tb.n_tiles, stream);
    cudaMemsetAsync(tb.level_end, 0, sizeof(int) * (size_t)aq.max_depth * tb.n_scanlines, stream);
    const int block = 128;
    for (int l = 0; l < aq.max_depth; l++) {
        // level l holds at most min(2^l * n_paths, ray_capacity) rays
        int64_t bound = n_paths;
        for (int k = 0; k < l && bound < tb.ray_capacity; k++) bound *= 2;
        if (bound > tb.ray_capacity) bound = tb.ray_capacity;
        const int grid = grid_for(bound, block, sm_count, MCRT_BOUNCE_GRID_CTAS_PER_SM);
        if (l == 0) k_tree_level<true><<<grid, block, 0, stream>>>(sc, aq, fr, tb, l);
        else k_tree_level<false><<<grid, block, 0, stream>>>(sc, aq, fr, tb, l);
        if (launches) (*launches)++;
        if (l + 1 < aq.max_depth) {
            const int tiles = (int)(((bound + 31) / 32 + 255) / 256);
            k_compact_tree<<<tiles, 256, 0, stream>>>((l & 1) ? tb.queue_a : tb.queue_b, tb.warp_counts, tb.tile_counts + (size_t)l * tb.n_tiles, tb.counters, l,
                                                      (int)n_paths, tb.ray_capacity);
            if (launches) (*launches)++;
        }
    }
}

void launch_closest_hit(const SceneDev& sc, int64_t n, const float* d_from, const float* d_to, int32_t* d_tri, int32_t* d_mesh,
                        float* d_frac, float* d_point, float* d_normal, cudaStream_t stream)
{
    if (n <= 0) return;
    const int block = 128;
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 64) g = 148 * 64;
    k_closest_hit<<<(int)g, block, 0, stream>>>(sc, n, d_from, d_to, d_tri, d_mesh, d_frac, d_point, d_normal);
}

void launch_elements(const AcqDev& aq, const FrameDev& fr, float* d_pos, float* d_dir, cudaStream_t stream)
{
    const int n = fr.n_poses * aq.elements;
    k_elements<<<(n + 127) / 128, 128, 0, stream>>>(aq, fr, d_pos, d_dir);
}

void launch_numerics_probe(int op, int64_t n, const double* d_a, const double* d_b, double* d_out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_numerics_probe<<<(int)((n + 255) / 256), 256, 0, stream>>>(op, n, d_a, d_b, d_out);
}

}  // namespace mcrt
