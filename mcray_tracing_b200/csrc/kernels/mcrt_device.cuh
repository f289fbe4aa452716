// mcrt_device.cuh -- device-side data layout and the ray-physics device functions shared by the
// trace kernels.  Arithmetic follows the reference's evaluation order in fp32/fp64 (citations to
// thepochynsons/MCRay-Tracing); this translation unit family is compiled with -fmad=false so no
// multiply-add is ever contracted, and all transcendentals come from mcrt_numerics.h.
#ifndef MCRT_DEVICE_CUH
#define MCRT_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/mcrt_numerics.h"

namespace mcrt {

// ------------------------------------------------------------------------------------------------
// HBM layout
// ------------------------------------------------------------------------------------------------
struct DevMaterial { float impedance, attenuation, mu0, mu1, sigma, specularity, shininess, thickness; };   // 32 B
struct DevMesh { float ox, oy, oz; int mat_in; int mat_out; int vascular; int pad0, pad1; };                // 32 B

// BVH2 node, 64 B = one half cache line pair, fetched as 4 x LDG.128.
//   a = (c0.lo.x, c0.lo.y, c0.lo.z, c0.hi.x)   b = (c0.hi.y, c0.hi.z, c1.lo.x, c1.lo.y)
//   c = (c1.lo.z, c1.hi.x, c1.hi.y, c1.hi.z)   d = (child0, child1, -, -)
// child >= 0: internal node index; child < 0: leaf of 1..MCRT_LEAF_MAX consecutive triangle slots,
// encoded -(1 + first * 4 + (count - 1)).
#ifndef MCRT_LEAF_MAX
#define MCRT_LEAF_MAX 1       // measured on ircad11 (profiles/r01_traversal_ab.txt): 1 beats 2 and 4
#endif
#ifndef MCRT_STACK_CULL
#define MCRT_STACK_CULL 0     // measured: storing the entry parameter on the stack costs more than it saves
#endif
struct __align__(16) BvhNode { float4 a, b, c; int4 d; };

// BVH4 node, 128 B = one cache line: every inner BVH2 node collapsed with its inner children (k_collapse_bvh4).  Planes are
// stored per axis for the 4 children, so a ray picks its near / far planes by ADDRESS (sign of the direction) instead of
// by selects, and one FFMA per plane does the slab test.  Empty slots hold an inverted box and child = MCRT_BVH4_EMPTY.
//   lox, loy, loz, hix, hiy, hiz = planes of children 0..3;  child = the 4 child references (same encoding as BVH2)
#ifndef MCRT_BVH4
#define MCRT_BVH4 1           // measured against the BVH2 traversal in profiles/r01ab_ab_bvh4.txt
#endif
#define MCRT_BVH4_EMPTY 0x7fffffff
#ifndef MCRT_PREFETCH
#define MCRT_PREFETCH 0       // 1: prefetch the cache line of every reference pushed on the traversal stack.  Measured slower
                              // (trace 2.64 vs 2.57 ms per 512 frames, config 4 4.79 vs 4.57 ms; profiles/r02d_ab_prefetch_smemstack.txt):
                              // the loop is bound by ALU issue, not by the latency of its node fetches
#endif
#ifndef MCRT_SMEM_STACK
#define MCRT_SMEM_STACK 0     // > 0: the first N entries of the 4-wide traversal stack live in shared memory ([entry][thread], conflict-free),
                              // deeper entries in local memory.  Measured slower with N = 8 (trace +6.7 %, same file): the range check per
                              // push / pop costs more ALU issue slots than the L1-resident local-memory stack ever waited for
#endif
#ifndef MCRT_STACK_CULL4
#define MCRT_STACK_CULL4 0    // 1 (4-wide, sorted traversal): a stack entry is the 64-bit pair (child reference, entry parameter), written with one
                              // STL.64; a popped entry whose entry parameter already lies beyond the current best hit -- by the very test its parent's
                              // visit applied -- is dropped without fetching the node
#endif
#ifndef MCRT_FOLD_SLACK
#define MCRT_FOLD_SLACK 0     // 1: the relative slack of the interval test (1 + 2e-6) is folded into the ray's far-plane constants at set-up
                              // instead of one FFMA per child box and visit
#endif
#define MCRT_TRACE_THREADS 128            // CTA size of every kernel that traverses (stride of the shared-memory stack)
#ifndef MCRT_BVH4_SORT
#define MCRT_BVH4_SORT 1      // 1: the hit children of a node are visited in entry order; 0: nearest first, the rest in slot order
#endif
struct __align__(16) Bvh4Node { float4 lox, loy, loz, hix, hiy, hiz; int4 child; int4 pad; };

// BVH8 node, 256 B = two cache lines (round 2, the default): built by k_collapse_bvh8_level.  Planes per axis for the 8
// slots (empty slot = inverted box); slot bit a set = the child lies towards +axis a of the node centre, so a ray visits
// the slots in the order of descending (slot XOR oinv), oinv bit a = the ray travels towards +a, without sorting.  The
// inner children are nodes child_base + (number of inner slots below this one), the leaf children the triangle slots
// tri_base + (number of leaf slots below): the stack holds ONE entry per visited node, not one per child.
#ifndef MCRT_BVH8
#define MCRT_BVH8 0           // measured against the sorted 4-wide traversal (profiles/r02c_ab_bvh8.txt): 13.4 instead of 17.5 node
                              // visits per query, but 107 instead of 70 child boxes tested and 2.2 instead of 1.7 triangle tests
                              // (octant order is approximate) -- on a kernel bound by the ALU pipe that is 13 % slower (trace 2.91 vs
                              // 2.56 ms per 512 frames, config 4: 5.13 vs 4.57 ms).  Kept as a build option (-DMCRT_BVH8=1), bit-identical.
#endif
struct __align__(16) Bvh8Node {
    float lox[8], loy[8], loz[8], hix[8], hiy[8], hiz[8];
    int child_base, tri_base;
    unsigned masks;           // bits 0..7: slot holds an inner node; bits 8..15: slot holds a triangle
    int pad[13];
};
#define MCRT_STACK_DEPTH8 40               // one entry per level of the wide tree (checked at build time)

// Triangle slot (48 B, Morton order): local-frame vertices v_obj*scaling; v0.w = mesh id bits,
// v1.w = original (objloader-order) triangle id bits.
struct __align__(16) TriSlot { float4 v0, v1, v2; };

// One emitted segment, 64 B (ray.h:28-36 minus the dangling media reference, B-1).
//   s0 = (from.xyz, reflected_intensity)  s1 = (dir.xyz, initial_intensity)
//   s2 = (to.xyz, attenuation)            s3 = (distance_traveled lo, hi [double bits], media_id, tri_id)
struct __align__(16) DevSegment { float4 s0, s1, s2; int4 s3; };

#define MCRT_MAX_SMEM_MESHES 64
#define MCRT_MAX_SMEM_MATERIALS 64

struct SceneDev {
    const BvhNode* nodes;
    const Bvh4Node* nodes4;  // the same tree, 4-wide (nullptr when the scene has < 2 triangles or MCRT_BVH8)
    const Bvh8Node* nodes8;  // the same tree, 8-wide (MCRT_BVH8; `tris` is then in the wide tree's leaf order)
    const TriSlot* tris;
    const DevMesh* meshes;
    const DevMaterial* materials;
    int n_tri, n_mesh, n_mat;
    int starting_material;
    float spacing[3];
    float max_abs;          // largest |coordinate| of any world-space triangle box
    float bounds_lo[3];     // world-space bounding box of the scene and 1/extent (coherence sort keys)
    float bounds_inv[3];
};

struct AcqDev {
    int elements, samples, max_depth, rows;
    float frequency;
    float axres_f;          // axial_resolution.to<float>()
    float radius_f;         // transducer_radius.to<float>() [cm]
    float vol_resolution;   // resolution_um / 1000.0f
    double axres_mm;
    double time_step_us;
    double row_period_us;
    double inv_row_period;
    double max_travel_time_us;
    double speed;           // (double)speed_of_sound
    int deterministic;
    int element_offset;     // global index of local element 0 (scanline-block runs; 0 otherwise): RNG key, PSF borders
    int voxel_fma_division; // 1: coord / vol_resolution via div_fma was validated exhaustively for this resolution (image.cu)
    int accumulate_windowed;// 1: row-window synchronous accumulate (no columns in HBM) when the scene allows it (image.cu)
    int rf_pitch;           // row stride (floats) of the raw RF image the accumulate stage writes: rows, or rows rounded up to a
                            // multiple of 4 so that every scanline starts 16-byte aligned (TMA bulk copies / 128-bit staging in the post kernel)
};

struct PoseTrigDev { float px, py, pz, cz, sz, cx, sx, cy, sy, pad0, pad1, pad2; };   // = mcrt::PoseTrig, 48 B

// Path state between bounces (ray.h:13-26), SoA.
struct PathState {
    float4* origin_intensity;   // (from.xyz, intensity)
    float4* dir_state;          // (dir.xyz, packed: media | (outside+2)<<8 | depth<<16)
    double* distance;           // distance_traveled [mm]
};

#define MCRT_OUTSIDE_NULL (-1)
#define MCRT_OUTSIDE_SELF (-2)
#define MCRT_INTENSITY_EPSILON 1e-10f     // ray.h:24
#define MCRT_STACK_DEPTH 64                // >= depth of the BVH (checked at build time)
#define MCRT_STACK_DEPTH4 96               // BVH4: up to 3 pushes per level, ceil(depth / 2) levels (checked at build time)

// ------------------------------------------------------------------------------------------------
// btVector3, scalar path
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 v_add(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 v_sub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 v_neg(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 v_scl(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float v_dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 v_cross(float3 a, float3 b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float v_length(float3 a) { return sqrtf(v_dot(a, a)); }
__device__ __forceinline__ float3 v_normalized(float3 a) { return v_scl(a, 1.0f / v_length(a)); }
__device__ __forceinline__ float3 v_interpolate3(float3 v0, float3 v1, float rt)     // btVector3::setInterpolate3
{
    const float s = 1.0f - rt;
    return make_float3(s * v0.x + rt * v1.x, s * v0.y + rt * v1.y, s * v0.z + rt * v1.z);
}
// btVector3::rotate with the host-supplied cos/sin of the angle
__device__ __forceinline__ float3 v_rotate(float3 v, float3 axis, float c, float s)
{
    const float3 o = v_scl(axis, v_dot(axis, v));
    const float3 x_ = v_sub(v, o);
    const float3 y_ = v_cross(axis, v);
    return v_add(v_add(o, v_scl(x_, c)), v_scl(y_, s));
}

// ------------------------------------------------------------------------------------------------
// ray_physics (ray.cpp)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rp_max_ray_length(float attenuation, float intensity, float frequency)   // ray.cpp:110-113
{
    return 10.f * mc_logf(MCRT_INTENSITY_EPSILON / intensity) / -attenuation * frequency;
}

__device__ __forceinline__ float3 rp_snells_law(float3 l, float3 n, float c, float refraction_angle, float r)   // ray.cpp:115-124
{
    return v_add(v_scl(l, r), v_scl(n, r * c - refraction_angle));
}

__device__ __forceinline__ float rp_reflection_intensity(float intensity_in, float media_1, float incidence_angle, float media_2,
                                                         float refracted_angle)                                    // ray.cpp:126-132
{
    const float num = media_1 * incidence_angle - media_2 * refracted_angle;
    const float denom = media_1 * incidence_angle + media_2 * refracted_angle;
    const double q = (double)(num / denom);
    return (float)((double)intensity_in * (q * q));
}

__device__ __forceinline__ float rp_reflected_intensity_eq8(float3 d, float3 refr, float3 refl, float specularity)  // ray.cpp:154-164
{
    const float refraction_angle = v_dot(d, refr);
    float refraction_factor = mc_powf(refraction_angle, specularity);
    const float reflection_angle = v_dot(d, refl);
    float reflection_factor = mc_powf(reflection_angle, specularity);
    if (refraction_factor != refraction_factor) refraction_factor = 0.0f;      // B-5
    if (reflection_factor != reflection_factor) reflection_factor = 0.0f;
    return fmaxf(refraction_factor, 0.0f) + fmaxf(reflection_factor, 0.0f);
}

__device__ __forceinline__ float rp_power_cosine_variate(int v, double number)    // ray.cpp:213-224 (int v: B-8)
{
    const int indice = v + 1;
    const float exponente = (float)((double)1.0 / indice);
    return (float)mc_pow(number, (double)exponente);
}

// ray.cpp:167-211, one disk-sampling attempt
__device__ __forceinline__ bool rp_random_unit_vector_attempt(float3 v, float cos_theta, double u_az, double u_rad, float3& out)
{
    const double a = u_az * 2 * MC_PI_D;
    const double r = 0.5 * sqrt(u_rad);
    double sn, cs;
    mc_sincos(a, &sn, &cs);
    float px = (float)(r * cs);
    float py = (float)(r * sn);
    const float p = px * px + py * py;
    if (!(p <= 0.25f)) return false;
    bool flag = false;
    float vx = v.x, vy = v.y;
    const float vz = v.z;
    if (fabsf(vx) > fabsf(vy)) { vx = vy; vy = v.x; flag = true; }
    const float b = 1 - vx * vx;
    float radicando = 1 - cos_theta * cos_theta;
    radicando = radicando / (p * b);
    const float c = sqrtf(radicando);
    px = px * c;
    py = py * c;
    const float d = cos_theta - vx * px;
    float wx = vx * cos_theta - b * px;
    float wy = vy * d + vz * py;
    const float wz = vz * d - vy * py;
    if (flag) { const float aux = wy; wy = wx; wx = aux; }
    out = make_float3(wx, wy, wz);
    return true;
}

// scene.cpp:281-290
__device__ __forceinline__ double rp_distance_in_mm(const float* spacing, float3 v1, float3 v2)
{
    const float x_dist = fabsf(v1.x - v2.x) * spacing[0];
    const float y_dist = fabsf(v1.y - v2.y) * spacing[1];
    const float z_dist = fabsf(v1.z - v2.z) * spacing[2];
    const double xd = x_dist, yd = y_dist, zd = z_dist;
    return sqrt(xd * xd + yd * yd + zd * zd) * 10;
}

// ------------------------------------------------------------------------------------------------
// closest hit: device BVH traversal + Bullet's triangle ray-cast (btTriangleRaycastCallback)
// ------------------------------------------------------------------------------------------------
struct HitRec {
    float fraction;     // m_closestHitFraction (1.0f = no hit)
    int tri_id;         // original triangle id, -1 = none
    int mesh;
    float3 n_raw;       // un-normalised triangle normal of the best hit
    float dist_a;       // signed plane distance of the ray origin (normal flip rule)
};

__device__ __forceinline__ void tri_test(const TriSlot* __restrict__ tris, int slot, const float4* __restrict__ s_mesh, float3 from_w,
                                         float3 to_w, HitRec& best)
{
    const float4 q0 = __ldg(&tris[slot].v0);
    const float4 q1 = __ldg(&tris[slot].v1);
    const float4 q2 = __ldg(&tris[slot].v2);
    const int mesh = __float_as_int(q0.w);
    const int tri = __float_as_int(q1.w);
    const float4 mo = s_mesh[mesh];
    // worldTocollisionObject * p with an identity basis = p - body origin
    const float3 from_l = make_float3(from_w.x - mo.x, from_w.y - mo.y, from_w.z - mo.z);
    const float3 to_l = make_float3(to_w.x - mo.x, to_w.y - mo.y, to_w.z - mo.z);
    const float3 vert0 = make_float3(q0.x, q0.y, q0.z), vert1 = make_float3(q1.x, q1.y, q1.z), vert2 = make_float3(q2.x, q2.y, q2.z);
    const float3 v10 = v_sub(vert1, vert0);
    const float3 v20 = v_sub(vert2, vert0);
    const float3 n = v_cross(v10, v20);
    const float dist = v_dot(vert0, n);
    float dist_a = v_dot(n, from_l);
    dist_a -= dist;
    float dist_b = v_dot(n, to_l);
    dist_b -= dist;
    if (dist_a * dist_b >= 0.0f) return;
    const float proj_length = dist_a - dist_b;
    const float distance = dist_a / proj_length;
    if (!(distance < best.fraction || (distance == best.fraction && tri < best.tri_id))) return;
    float edge_tolerance = v_dot(n, n);
    edge_tolerance *= -0.0001f;
    const float3 point = v_interpolate3(from_l, to_l, distance);
    const float3 v0p = v_sub(vert0, point);
    const float3 v1p = v_sub(vert1, point);
    const float3 cp0 = v_cross(v0p, v1p);
    if (v_dot(cp0, n) >= edge_tolerance) {
        const float3 v2p = v_sub(vert2, point);
        const float3 cp1 = v_cross(v1p, v2p);
        if (v_dot(cp1, n) >= edge_tolerance) {
            const float3 cp2 = v_cross(v2p, v0p);
            if (v_dot(cp2, n) >= edge_tolerance) {
                best.fraction = distance;
                best.tri_id = tri;
                best.mesh = mesh;
                best.n_raw = n;
                best.dist_a = dist_a;
            }
        }
    }
}

// Conservative slab test of one child box against the segment (see DESIGN.md "Traversal"): the
// ray origin is widened by +-e per axis (o_near / o_far) and the interval compare carries a
// relative slack, so the BVH can never cull a triangle the exact per-triangle arithmetic accepts.
#ifndef MCRT_WHILE_WHILE
#define MCRT_WHILE_WHILE 0
#endif
#ifndef MCRT_BOX_FMA
#define MCRT_BOX_FMA 1   // slab planes as one FFMA each (plane * inv - origin * inv); measured in profiles/r01_traversal_ab.txt
#endif
struct RayBox {
#if MCRT_BOX_FMA
    float cnx, cny, cnz;   // -(o_near * inv): o_near = o + e where inv >= 0 else o - e
    float cfx, cfy, cfz;   // -(o_far * inv)
#else
    float onx, ony, onz;   // origin for the near plane (o + e where inv >= 0 else o - e)
    float ofx, ofy, ofz;   // origin for the far plane
#endif
    float ix, iy, iz;      // 1 / (to - from)
#if MCRT_FOLD_SLACK
    float sfx, sfy, sfz;   // ix, iy, iz times (1 + 2e-6) and ...
    float scx, scy, scz;   // ... cfx, cfy, cfz times (1 + 2e-6): far planes that carry the interval test's relative slack (4-wide traversal)
#endif
    bool px, py, pz;       // inv >= 0
};

__device__ __forceinline__ RayBox make_raybox(float3 from_w, float3 to_w, float scene_max_abs)
{
    RayBox r;
    const float dx = to_w.x - from_w.x, dy = to_w.y - from_w.y, dz = to_w.z - from_w.z;
    r.ix = 1.0f / dx; r.iy = 1.0f / dy; r.iz = 1.0f / dz;
#if MCRT_BOX_FMA
    // an axis-parallel ray (d == 0) gets a huge finite reciprocal: plane * inv - origin * inv must not become
    // inf - inf; t = (plane - origin) * 1e30 still separates "inside the slab" from "outside" for any gap > 1e-30
    if (!(fabsf(r.ix) <= 1e30f)) r.ix = copysignf(1e30f, dx);
    if (!(fabsf(r.iy) <= 1e30f)) r.iy = copysignf(1e30f, dy);
    if (!(fabsf(r.iz) <= 1e30f)) r.iz = copysignf(1e30f, dz);
#endif
    r.px = r.ix >= 0.0f; r.py = r.iy >= 0.0f; r.pz = r.iz >= 0.0f;
    const float mo = fmaxf(fabsf(from_w.x), fmaxf(fabsf(from_w.y), fabsf(from_w.z)));
    const float e = 4e-6f * (scene_max_abs + mo) + 1e-7f;
#if MCRT_BOX_FMA
    // rounding of the two-term form: <= 2^-24 (|o * inv| + |t|), i.e. <= 1.2e-7 (max_abs + |o|) in space -- 30x inside e
    r.cnx = -((r.px ? from_w.x + e : from_w.x - e) * r.ix); r.cfx = -((r.px ? from_w.x - e : from_w.x + e) * r.ix);
    r.cny = -((r.py ? from_w.y + e : from_w.y - e) * r.iy); r.cfy = -((r.py ? from_w.y - e : from_w.y + e) * r.iy);
    r.cnz = -((r.pz ? from_w.z + e : from_w.z - e) * r.iz); r.cfz = -((r.pz ? from_w.z - e : from_w.z + e) * r.iz);
#if MCRT_FOLD_SLACK
    r.sfx = r.ix * 1.000002f; r.sfy = r.iy * 1.000002f; r.sfz = r.iz * 1.000002f;
    r.scx = r.cfx * 1.000002f; r.scy = r.cfy * 1.000002f; r.scz = r.cfz * 1.000002f;
#endif
#else
    r.onx = r.px ? from_w.x + e : from_w.x - e; r.ofx = r.px ? from_w.x - e : from_w.x + e;
    r.ony = r.py ? from_w.y + e : from_w.y - e; r.ofy = r.py ? from_w.y - e : from_w.y + e;
    r.onz = r.pz ? from_w.z + e : from_w.z - e; r.ofz = r.pz ? from_w.z - e : from_w.z + e;
#endif
    return r;
}

__device__ __forceinline__ bool box_test(const RayBox& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float tbest,
                                         float& tnear)
{
    const float nx = r.px ? lox : hix, fx = r.px ? hix : lox;
    const float ny = r.py ? loy : hiy, fy = r.py ? hiy : loy;
    const float nz = r.pz ? loz : hiz, fz = r.pz ? hiz : loz;
#if MCRT_BOX_FMA
    // the box test only has to be conservative, not bit-exact: explicit FFMA (the library is built -fmad=false)
    const float t0x = __fmaf_rn(nx, r.ix, r.cnx), t1x = __fmaf_rn(fx, r.ix, r.cfx);
    const float t0y = __fmaf_rn(ny, r.iy, r.cny), t1y = __fmaf_rn(fy, r.iy, r.cfy);
    const float t0z = __fmaf_rn(nz, r.iz, r.cnz), t1z = __fmaf_rn(fz, r.iz, r.cfz);
    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tbest));
    tnear = tn;
    return tn <= __fmaf_rn(tf, 1.000002f, 1e-37f);
#else
    const float t0x = (nx - r.onx) * r.ix, t1x = (fx - r.ofx) * r.ix;
    const float t0y = (ny - r.ony) * r.iy, t1y = (fy - r.ofy) * r.iy;
    const float t0z = (nz - r.onz) * r.iz, t1z = (fz - r.ofz) * r.iz;
    // fmaxf / fminf drop NaN operands (0 * inf when the origin sits exactly on a slab plane)
    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tbest));
    tnear = tn;
    return tn <= tf * 1.000002f + 1e-37f;
#endif
}

// Stack-based closest-hit traversal, near child first.  `node_visits` / `tri_tests` are per-thread work
// counters (two integer adds per iteration; reported through mcrt_stats when "count_traversal" is on).
__device__ __forceinline__ void closest_hit(const SceneDev& sc, const float4* __restrict__ s_mesh, const unsigned char* __restrict__ s_perm,
                                            int* __restrict__ s_stack /* this thread's column of the shared-memory stack, or nullptr */,
                                            float3 from_w, float3 to_w, HitRec& best, int& node_visits, int& tri_tests)
{
    best.fraction = 1.0f; best.tri_id = -1; best.mesh = -1; best.n_raw = make_float3(0.f, 0.f, 0.f); best.dist_a = 0.0f;
    if (sc.n_tri <= 0) return;
    if (sc.n_tri == 1) { tri_test(sc.tris, 0, s_mesh, from_w, to_w, best); tri_tests++; return; }
    const RayBox rb = make_raybox(from_w, to_w, sc.max_abs);
#if MCRT_BVH8 && MCRT_BOX_FMA
    {
        // 8-wide traversal in octant order with a node-group stack.  Per node: 12 LDG.128 of planes (near / far picked by
        // address) + 1 of the header, one FFMA per plane, the 8 hit bits, one table look-up that moves hit bit `slot` to
        // bit (slot XOR oinv) -- the visiting priority -- and at most one stack push.
        const unsigned long long base = (unsigned long long)reinterpret_cast<uintptr_t>(sc.nodes8);
        const unsigned onx = rb.px ? 0u : 96u, ofx = rb.px ? 96u : 0u;     // byte offsets of lox / hix
        const unsigned ony = rb.py ? 32u : 128u, ofy = rb.py ? 128u : 32u;
        const unsigned onz = rb.pz ? 64u : 160u, ofz = rb.pz ? 160u : 64u;
        const unsigned oinv = (rb.px ? 1u : 0u) | (rb.py ? 2u : 0u) | (rb.pz ? 4u : 0u);
        const unsigned char* perm = s_perm + (oinv << 8);
        uint2 stack8[MCRT_STACK_DEPTH8];
        int sp = 0;
        unsigned cur_base = 0u, cur_imask = 1u, cur_hp = 1u << oinv;        // the root: "slot 0" of a virtual parent
        while (true) {
            if (cur_hp == 0u) {
                if (sp == 0) break;
                const uint2 e = stack8[--sp];
                cur_base = e.x; cur_imask = e.y >> 8; cur_hp = e.y & 0xffu;
            }
            // next child of the current group: highest priority bit
            const unsigned bit = 31u - (unsigned)__clz((int)cur_hp);
            cur_hp ^= 1u << bit;                                             // the rest of the group
            const unsigned slot = bit ^ oinv;
            const unsigned node = cur_base + (unsigned)__popc(cur_imask & ((1u << slot) - 1u));
            node_visits++;
            const unsigned long long nd = base + (unsigned long long)node * (unsigned long long)sizeof(Bvh8Node);
            const float4* pnx = reinterpret_cast<const float4*>(nd | onx); const float4* pfx = reinterpret_cast<const float4*>(nd | ofx);
            const float4* pny = reinterpret_cast<const float4*>(nd | ony); const float4* pfy = reinterpret_cast<const float4*>(nd | ofy);
            const float4* pnz = reinterpret_cast<const float4*>(nd | onz); const float4* pfz = reinterpret_cast<const float4*>(nd | ofz);
            const int4 hd = __ldg(reinterpret_cast<const int4*>(nd) + 12);
            const float tb = best.fraction * 1.000002f;
            unsigned h = 0u;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const float4 nx = __ldg(pnx + half), fx = __ldg(pfx + half), ny = __ldg(pny + half), fy = __ldg(pfy + half);
                const float4 nz = __ldg(pnz + half), fz = __ldg(pfz + half);
                const float n0[4] = {nx.x, nx.y, nx.z, nx.w}, n1[4] = {ny.x, ny.y, ny.z, ny.w}, n2[4] = {nz.x, nz.y, nz.z, nz.w};
                const float f0[4] = {fx.x, fx.y, fx.z, fx.w}, f1[4] = {fy.x, fy.y, fy.z, fy.w}, f2[4] = {fz.x, fz.y, fz.z, fz.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float t0x = __fmaf_rn(n0[k], rb.ix, rb.cnx), t1x = __fmaf_rn(f0[k], rb.ix, rb.cfx);
                    const float t0y = __fmaf_rn(n1[k], rb.iy, rb.cny), t1y = __fmaf_rn(f1[k], rb.iy, rb.cfy);
                    const float t0z = __fmaf_rn(n2[k], rb.iz, rb.cnz), t1z = __fmaf_rn(f2[k], rb.iz, rb.cfz);
                    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
                    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tb));
                    if (tn <= __fmaf_rn(tf, 1.000002f, 1e-37f)) h |= 1u << (half * 4 + k);      // an empty slot's inverted box never passes
                }
            }
            const unsigned masks = (unsigned)hd.z;
            // the node's triangles first: a hit shortens the ray for everything that follows
            unsigned hl = h & (masks >> 8);
            while (hl) {
                const unsigned sl = 31u - (unsigned)__clz((int)hl);
                hl ^= 1u << sl;
                tri_tests++;
                tri_test(sc.tris, hd.y + __popc((masks >> 8) & ((1u << sl) - 1u)), s_mesh, from_w, to_w, best);
            }
            const unsigned hi = h & masks & 0xffu;
            if (hi) {
                // descend: the rest of the old group goes on the stack, this node's hit children become the current group
                if (cur_hp) stack8[sp++] = make_uint2(cur_base, (cur_imask << 8) | cur_hp);
                cur_base = (unsigned)hd.x; cur_imask = masks & 0xffu; cur_hp = perm[hi];
            }
        }
        return;
    }
#endif
#if MCRT_BVH4 && MCRT_BOX_FMA
    {
        // 4-wide traversal: half the dependent node fetches of the BVH2 loop below.  Near / far planes are picked by address.
        // A node is 128-byte aligned and the plane offsets are < 128, so (node address | offset) == (node address + offset):
        // ONE 64-bit multiply-add forms the node address and each of the seven loads only ORs its offset into the low word
        // (round 1: a 64-bit add per load, 21 of the ~135 instructions of a node visit).  Forming all seven addresses on the
        // FMA pipe instead (six per-ray plane bases, one IMAD.WIDE each, no ALU instruction at all) measured 3.7 % SLOWER:
        // the visit is as sensitive to the length of its dependent chain as to its instruction count
        // (profiles/r02f_ab_fma_addresses.txt).
        const unsigned long long base = (unsigned long long)reinterpret_cast<uintptr_t>(sc.nodes4);
        const unsigned onx = rb.px ? 0u : 48u, ofx = rb.px ? 48u : 0u;   // byte offsets of lox / hix
        const unsigned ony = rb.py ? 16u : 64u, ofy = rb.py ? 64u : 16u;
        const unsigned onz = rb.pz ? 32u : 80u, ofz = rb.pz ? 80u : 32u;
#if MCRT_STACK_CULL4
        unsigned long long stack4[MCRT_STACK_DEPTH4];   // (entry parameter bits << 32) | child reference
#else
        int stack4[MCRT_STACK_DEPTH4];
#endif
        int sp4 = 0;
        int node4 = 0;
#if MCRT_SMEM_STACK > 0
#define MCRT_PUSH(v) { if (sp4 < MCRT_SMEM_STACK) s_stack[sp4 * MCRT_TRACE_THREADS] = (v); else stack4[sp4 - MCRT_SMEM_STACK] = (v); sp4++; }
#define MCRT_POP() (--sp4, sp4 < MCRT_SMEM_STACK ? s_stack[sp4 * MCRT_TRACE_THREADS] : stack4[sp4 - MCRT_SMEM_STACK])
#elif MCRT_STACK_CULL4
#define MCRT_PUSH_T(v, tt) { stack4[sp4++] = ((unsigned long long)__float_as_uint(tt) << 32) | (unsigned long long)(unsigned)(v); }
#else
#define MCRT_PUSH(v) { stack4[sp4++] = (v); }
#define MCRT_POP() (stack4[--sp4])
#endif
#if MCRT_PREFETCH
        const unsigned long long tri_base_addr = (unsigned long long)reinterpret_cast<uintptr_t>(sc.tris);
#define MCRT_PREFETCH_REF(ref) { const int r_ = (ref); \
            const unsigned long long a_ = r_ >= 0 ? base + (unsigned long long)(unsigned)r_ * 128ull : tri_base_addr + (unsigned long long)(unsigned)((-r_ - 1) >> 2) * 48ull; \
            asm volatile("prefetch.global.L1 [%0];" :: "l"(a_)); }
#else
#define MCRT_PREFETCH_REF(ref) {}
#endif
        while (true) {
            if (node4 >= 0) {
                node_visits++;
                const unsigned long long nd = base + (unsigned long long)(unsigned)node4 * (unsigned long long)sizeof(Bvh4Node);
                const float4 nx = __ldg(reinterpret_cast<const float4*>(nd | onx)), fx = __ldg(reinterpret_cast<const float4*>(nd | ofx));
                const float4 ny = __ldg(reinterpret_cast<const float4*>(nd | ony)), fy = __ldg(reinterpret_cast<const float4*>(nd | ofy));
                const float4 nz = __ldg(reinterpret_cast<const float4*>(nd | onz)), fz = __ldg(reinterpret_cast<const float4*>(nd | ofz));
                const int4 ch = __ldg(reinterpret_cast<const int4*>(nd) + 6);
#if MCRT_FOLD_SLACK
                const float tb = best.fraction * 1.0000041f;     // (1 + 2e-6)^2 rounded up: the slack of the best hit and of the interval test
#else
                const float tb = best.fraction * 1.000002f;
#endif
                float t[4];
                int c[4] = {ch.x, ch.y, ch.z, ch.w};
                {
                    const float n0[4] = {nx.x, nx.y, nx.z, nx.w}, n1[4] = {ny.x, ny.y, ny.z, ny.w}, n2[4] = {nz.x, nz.y, nz.z, nz.w};
                    const float f0[4] = {fx.x, fx.y, fx.z, fx.w}, f1[4] = {fy.x, fy.y, fy.z, fy.w}, f2[4] = {fz.x, fz.y, fz.z, fz.w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
#if MCRT_FOLD_SLACK
                        // the far planes are evaluated with constants that already carry the factor (1 + 2e-6): tf here IS tf * 1.000002 of the
                        // other branch up to one rounding (the slack is 17 ulp), and a box whose far side lies behind the origin still misses
                        const float t0x = __fmaf_rn(n0[k], rb.ix, rb.cnx), t1x = __fmaf_rn(f0[k], rb.sfx, rb.scx);
                        const float t0y = __fmaf_rn(n1[k], rb.iy, rb.cny), t1y = __fmaf_rn(f1[k], rb.sfy, rb.scy);
                        const float t0z = __fmaf_rn(n2[k], rb.iz, rb.cnz), t1z = __fmaf_rn(f2[k], rb.sfz, rb.scz);
                        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
                        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tb));
                        t[k] = (tn <= tf) ? tn : 3.0e38f;
#else
                        // (the clamp of the entry parameter at 0 cannot ride on an FFMA as its saturate modifier: an axis-parallel ray
                        // outside a box's slab has t0 = +1e30, which .SAT would turn into 1 -- a hit while nothing closer is known:
                        // 17.5 -> 21.7 node visits per query, profiles/r02e_ab_fma_addresses_sat.txt)
                        const float t0x = __fmaf_rn(n0[k], rb.ix, rb.cnx), t1x = __fmaf_rn(f0[k], rb.ix, rb.cfx);
                        const float t0y = __fmaf_rn(n1[k], rb.iy, rb.cny), t1y = __fmaf_rn(f1[k], rb.iy, rb.cfy);
                        const float t0z = __fmaf_rn(n2[k], rb.iz, rb.cnz), t1z = __fmaf_rn(f2[k], rb.iz, rb.cfz);
                        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
                        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tb));
                        // empty slots (inverted box) give tn = +inf / NaN-free miss; a miss sorts last
                        t[k] = (tn <= __fmaf_rn(tf, 1.000002f, 1e-37f)) ? tn : 3.0e38f;      // an empty slot's inverted box never passes
#endif
                    }
                }
#if MCRT_BVH4_SORT
                // sort the 4 (t, child) pairs by entry parameter (5-comparator network); misses carry t = 3e38
#define MCRT_CSWAP(i, j) { const bool sw = t[j] < t[i]; const float tt = sw ? t[j] : t[i]; t[j] = sw ? t[i] : t[j]; t[i] = tt; \
                           const int cc = sw ? c[j] : c[i]; c[j] = sw ? c[i] : c[j]; c[i] = cc; }
                MCRT_CSWAP(0, 1) MCRT_CSWAP(2, 3) MCRT_CSWAP(0, 2) MCRT_CSWAP(1, 3) MCRT_CSWAP(1, 2)
#undef MCRT_CSWAP
                if (t[0] < 3.0e38f) {
                    // nearest first; the others go on the stack far-to-near so the nearer one is popped first.  Every pushed
                    // reference WILL be fetched when it is popped (there is no culling on the stack), so its cache line is
                    // requested now: one LSU instruction that turns an L2 round trip at pop time into an L1 hit.
#if MCRT_STACK_CULL4
                    if (t[3] < 3.0e38f) MCRT_PUSH_T(c[3], t[3]);
                    if (t[2] < 3.0e38f) MCRT_PUSH_T(c[2], t[2]);
                    if (t[1] < 3.0e38f) MCRT_PUSH_T(c[1], t[1]);
#else
                    if (t[3] < 3.0e38f) { MCRT_PUSH(c[3]); MCRT_PREFETCH_REF(c[3]); }
                    if (t[2] < 3.0e38f) { MCRT_PUSH(c[2]); MCRT_PREFETCH_REF(c[2]); }
                    if (t[1] < 3.0e38f) { MCRT_PUSH(c[1]); MCRT_PREFETCH_REF(c[1]); }
#endif
                    node4 = c[0];
                    continue;
                }
#else
                // nearest child first (slot index packed into the low mantissa bits of its entry parameter, t >= 0 so the
                // bit pattern orders like the value); the other hit children are pushed in slot order
                const unsigned k0 = (__float_as_uint(t[0]) & ~3u) | 0u, k1 = (__float_as_uint(t[1]) & ~3u) | 1u;
                const unsigned k2 = (__float_as_uint(t[2]) & ~3u) | 2u, k3 = (__float_as_uint(t[3]) & ~3u) | 3u;
                const unsigned kmin = min(min(k0, k1), min(k2, k3));
                if (kmin < (__float_as_uint(3.0e38f) & ~3u)) {
                    const int near_slot = (int)(kmin & 3u);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (k != near_slot && t[k] < 3.0e38f) MCRT_PUSH(c[k]);
                    node4 = near_slot == 0 ? c[0] : (near_slot == 1 ? c[1] : (near_slot == 2 ? c[2] : c[3]));
                    continue;
                }
#endif
            } else {
                const int code = -node4 - 1;
                const int first = code >> 2, count = (code & 3) + 1;
                tri_tests += count;
                for (int k = 0; k < count; k++) tri_test(sc.tris, first + k, s_mesh, from_w, to_w, best);
            }
#if MCRT_STACK_CULL4
            // pop; an entry whose box the ray enters beyond the best hit found since it was pushed is dropped unfetched.  The bound is
            // the one its parent's visit would apply now (entry parameter <= min(far planes, best * s) * s implies <= best * s * s),
            // so nothing is culled here that the visit itself would not cull child by child
            {
#if MCRT_FOLD_SLACK
                const float tcull = best.fraction * 1.0000041f;
#else
                const float tcull = __fmaf_rn(best.fraction * 1.000002f, 1.000002f, 1e-37f);
#endif
                bool found = false;
                while (sp4 > 0) {
                    const unsigned long long e = stack4[--sp4];
                    if (__uint_as_float((unsigned)(e >> 32)) <= tcull) { node4 = (int)(unsigned)e; found = true; break; }
                }
                if (!found) break;
            }
#else
            if (sp4 == 0) break;
            node4 = MCRT_POP();
#endif
        }
        return;
    }
#endif
    int stack[MCRT_STACK_DEPTH];
#if MCRT_STACK_CULL
    float stack_t[MCRT_STACK_DEPTH];     // entry parameter of the pushed subtree: lets a pop be culled without a fetch
#endif
    int sp = 0;
    int node = 0;   // root
#if MCRT_WHILE_WHILE
    // while-while: every lane first descends to its next leaf (or runs out of nodes), the warp reconverges, then the
    // leaves are tested with (nearly) all lanes active instead of a few lanes per loop iteration
    const int kDone = (int)0x80000000;
    while (true) {
        while (node >= 0) {
            node_visits++;
            const BvhNode* nd = sc.nodes + node;
            const float4 a = __ldg(&nd->a), b = __ldg(&nd->b), c = __ldg(&nd->c);
            const int4 d = __ldg(&nd->d);
            const float tb = best.fraction * 1.000002f;
            float t0, t1;
            const bool h0 = box_test(rb, a.x, a.y, a.z, a.w, b.x, b.y, tb, t0);
            const bool h1 = box_test(rb, b.z, b.w, c.x, c.y, c.z, c.w, tb, t1);
            if (h0 && h1) {
                const bool swap = t1 < t0;
                node = swap ? d.y : d.x;
                stack[sp++] = swap ? d.x : d.y;
            } else if (h0) {
                node = d.x;
            } else if (h1) {
                node = d.y;
            } else {
                node = sp > 0 ? stack[--sp] : kDone;
            }
        }
        if (node == kDone) break;
        {
            const int code = -node - 1;
            const int first = code >> 2, count = (code & 3) + 1;
            tri_tests += count;
            for (int k = 0; k < count; k++) tri_test(sc.tris, first + k, s_mesh, from_w, to_w, best);
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
#else
    while (true) {
        if (node >= 0) {
            node_visits++;
            const BvhNode* nd = sc.nodes + node;
            const float4 a = __ldg(&nd->a), b = __ldg(&nd->b), c = __ldg(&nd->c);
            const int4 d = __ldg(&nd->d);
            const float tb = best.fraction * 1.000002f;
            float t0, t1;
            const bool h0 = box_test(rb, a.x, a.y, a.z, a.w, b.x, b.y, tb, t0);
            const bool h1 = box_test(rb, b.z, b.w, c.x, c.y, c.z, c.w, tb, t1);
            if (h0 && h1) {
                const bool swap = t1 < t0;
                node = swap ? d.y : d.x;
                stack[sp] = swap ? d.x : d.y;
#if MCRT_STACK_CULL
                stack_t[sp] = swap ? t0 : t1;
#endif
                sp++;
                continue;
            }
            if (h0) { node = d.x; continue; }
            if (h1) { node = d.y; continue; }
        } else {
            const int code = -node - 1;
            const int first = code >> 2, count = (code & 3) + 1;
            tri_tests += count;
            for (int k = 0; k < count; k++) tri_test(sc.tris, first + k, s_mesh, from_w, to_w, best);
        }
#if MCRT_STACK_CULL
        // pop, skipping subtrees that start beyond the current closest hit (with the same slack)
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stack_t[sp] <= best.fraction * 1.000002f) { node = stack[sp]; found = true; break; }
        }
        if (!found) break;
#else
        if (sp == 0) break;
        node = stack[--sp];
#endif
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// Warp-coherent ("packet") traversal with the node staged through shared memory -- the north star's literal design point, built as
// a compile-time option (MCRT_PACKET: 0 off, 1 the bounce-0 kernel k_first_hit only, 2 every traversal) and measured against the
// per-lane traversal above (profiles/r02ae_ab_packet.txt).  The warp walks ONE path through the tree with ONE stack (shared memory,
// entries = node reference + mask of the lanes whose rays entered it); a visited node's 128 bytes are fetched once, by eight lanes,
// into a per-warp shared-memory buffer, and every lane of the entry's mask tests its own ray (own direction signs, own closest hit
// so far) against the four children from there.  The children any lane hit are visited in the entry order of the mask's first lane.
// Leaves are tested by the lanes of their mask only.  Results are identical to the per-lane traversal (ties resolve by triangle id).
// The cost is the UNION of the nodes the 32 rays need, so it pays only while the rays of a warp are neighbours.
// ------------------------------------------------------------------------------------------------
#ifndef MCRT_PACKET
#define MCRT_PACKET 0
#endif
#define MCRT_PK_DEPTH 128
struct PacketShared {
    int2 stack[MCRT_TRACE_THREADS / 32][MCRT_PK_DEPTH];
    float4 node[MCRT_TRACE_THREADS / 32][8];
};

// wm: the lanes that take part -- established by the caller with a __ballot_sync executed by the whole warp, so the named lanes converge
// at the first __syncwarp(wm) even if the warp was split when it got here.
__device__ __forceinline__ void closest_hit_packet(const SceneDev& sc, const float4* __restrict__ s_mesh, PacketShared& ps, const unsigned wm,
                                                   float3 from_w, float3 to_w, HitRec& best, int& node_visits, int& tri_tests)
{
    best.fraction = 1.0f; best.tri_id = -1; best.mesh = -1; best.n_raw = make_float3(0.f, 0.f, 0.f); best.dist_a = 0.0f;
    if (sc.n_tri <= 0) return;
    if (sc.n_tri == 1) { tri_test(sc.tris, 0, s_mesh, from_w, to_w, best); tri_tests++; return; }
    __syncwarp(wm);
    const unsigned lane = threadIdx.x & 31u;
    const int w = threadIdx.x >> 5;
    const int n_act = __popc(wm), rank = __popc(wm & ((1u << lane) - 1u));
    const RayBox rb = make_raybox(from_w, to_w, sc.max_abs);
    const int inx = rb.px ? 0 : 3, ifx = rb.px ? 3 : 0;    // float4 index of lox / hix within the staged node
    const int iny = rb.py ? 1 : 4, ify = rb.py ? 4 : 1;
    const int inz = rb.pz ? 2 : 5, ifz = rb.pz ? 5 : 2;
    int2* __restrict__ stack = ps.stack[w];
    float4* __restrict__ nd = ps.node[w];
    const float4* __restrict__ nodes = reinterpret_cast<const float4*>(sc.nodes4);
    int sp = 0;
    int cur = 0;
    unsigned cur_mask = wm;
    while (true) {
        const bool mine = (cur_mask >> lane) & 1u;
        if (cur >= 0) {
            if (mine) node_visits++;
            for (int i = rank; i < 8; i += n_act) nd[i] = __ldg(nodes + (size_t)cur * 8 + i);
            __syncwarp(wm);
            const float4 nx = nd[inx], fx = nd[ifx], ny = nd[iny], fy = nd[ify], nz = nd[inz], fz = nd[ifz];
            const float4 chf = nd[6];
            __syncwarp(wm);                                    // everybody has read the buffer before the next node overwrites it
            const int c[4] = {__float_as_int(chf.x), __float_as_int(chf.y), __float_as_int(chf.z), __float_as_int(chf.w)};
            const float tb = best.fraction * 1.000002f;
            float t[4];
            {
                const float n0[4] = {nx.x, nx.y, nx.z, nx.w}, n1[4] = {ny.x, ny.y, ny.z, ny.w}, n2[4] = {nz.x, nz.y, nz.z, nz.w};
                const float f0[4] = {fx.x, fx.y, fx.z, fx.w}, f1[4] = {fy.x, fy.y, fy.z, fy.w}, f2[4] = {fz.x, fz.y, fz.z, fz.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float t0x = __fmaf_rn(n0[k], rb.ix, rb.cnx), t1x = __fmaf_rn(f0[k], rb.ix, rb.cfx);
                    const float t0y = __fmaf_rn(n1[k], rb.iy, rb.cny), t1y = __fmaf_rn(f1[k], rb.iy, rb.cfy);
                    const float t0z = __fmaf_rn(n2[k], rb.iz, rb.cnz), t1z = __fmaf_rn(f2[k], rb.iz, rb.cfz);
                    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
                    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tb));
                    t[k] = (mine && tn <= __fmaf_rn(tf, 1.000002f, 1e-37f)) ? tn : 3.0e38f;
                }
            }
            unsigned m[4];
            float tl[4];
            const int leader = __ffs(cur_mask) - 1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                m[k] = __ballot_sync(wm, t[k] < 3.0e38f);
                tl[k] = __shfl_sync(wm, t[k], leader);
                if (m[k] == 0u) tl[k] = 3.0e38f;                 // nobody enters: sorts last
                else if (!(tl[k] < 3.0e38f)) tl[k] = 2.9e38f;    // entered by others only: behind the leader's own children
            }
            int cc[4] = {c[0], c[1], c[2], c[3]};
#define MCRT_PSWAP(i, j) { const bool sw = tl[j] < tl[i]; const float tt = sw ? tl[j] : tl[i]; tl[j] = sw ? tl[i] : tl[j]; tl[i] = tt; \
                           const int q = sw ? cc[j] : cc[i]; cc[j] = sw ? cc[i] : cc[j]; cc[i] = q; \
                           const unsigned u = sw ? m[j] : m[i]; m[j] = sw ? m[i] : m[j]; m[i] = u; }
            MCRT_PSWAP(0, 1) MCRT_PSWAP(2, 3) MCRT_PSWAP(0, 2) MCRT_PSWAP(1, 3) MCRT_PSWAP(1, 2)
#undef MCRT_PSWAP
            if (m[0]) {
                if (m[3]) { stack[sp] = make_int2(cc[3], (int)m[3]); sp++; }
                if (m[2]) { stack[sp] = make_int2(cc[2], (int)m[2]); sp++; }
                if (m[1]) { stack[sp] = make_int2(cc[1], (int)m[1]); sp++; }
                cur = cc[0]; cur_mask = m[0];
                continue;
            }
        } else if (mine) {
            const int code = -cur - 1;
            const int first = code >> 2, count = (code & 3) + 1;
            tri_tests += count;
            for (int k = 0; k < count; k++) tri_test(sc.tris, first + k, s_mesh, from_w, to_w, best);
        }
        if (sp == 0) break;
        --sp;
        const int2 e = stack[sp];
        cur = e.x; cur_mask = (unsigned)e.y;
    }
    __syncwarp(wm);
}

// final normal of the best hit: normalise, face the ray origin (btTriangleRaycastCallback)
__device__ __forceinline__ float3 hit_normal(const HitRec& h)
{
    const float3 n = v_normalized(h.n_raw);
    return (h.dist_a <= 0.0f) ? v_neg(n) : n;
}

}  // namespace mcrt
#endif
