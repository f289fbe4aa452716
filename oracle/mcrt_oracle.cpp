// mcrt_oracle.cpp -- CPU ORACLE: a restatement of the reference's per-frame hot path.
// TEST INFRASTRUCTURE ONLY (see mcrt_oracle.h).  Every function cites the reference lines it
// follows (paths relative to thepochynsons/MCRay-Tracing).  Arithmetic is written in the
// reference's evaluation order, fp32 where the reference is fp32, fp64 where the units library
// makes it fp64; compile with -ffp-contract=off.
//
// Deliberate deviations: B-1 medium copied at segment emission; B-5 NaN
// from total internal reflection / pow of negative base mapped to 0; B-11 RNG = Philox keyed by
// (seed; frame, element, sample, bounce); B-14 march step count held in 64 bits; transcendentals
// on the ray path come from the shared numerics contract (mcrt_numerics.h) instead of glibc.
#include "mcrt_oracle.h"
#include "../mcray_tracing_b200/csrc/common/mcrt_numerics.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------------
// btVector3 (scalar, non-SSE path) restated as free functions
// ------------------------------------------------------------------------------------------------
struct v3 { float x, y, z; };
inline v3 mk(float x, float y, float z) { v3 r{x, y, z}; return r; }
inline v3 add(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline v3 sub(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline v3 neg(v3 a) { return mk(-a.x, -a.y, -a.z); }
inline v3 scl(v3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline v3 cross(v3 a, v3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(v3 a) { return sqrtf(dot(a, a)); }
inline v3 normalized(v3 a) { return scl(a, 1.0f / length(a)); }                 // *this *= 1/length()
inline v3 interpolate3(v3 v0, v3 v1, float rt)                                   // btVector3::setInterpolate3
{
    const float s = 1.0f - rt;
    return mk(s * v0.x + rt * v1.x, s * v0.y + rt * v1.y, s * v0.z + rt * v1.z);
}
inline v3 rotate(v3 v, v3 axis, float angle)                                     // btVector3::rotate
{
    const v3 o = scl(axis, dot(axis, v));
    const v3 x_ = sub(v, o);
    const v3 y_ = cross(axis, v);
    return add(add(o, scl(x_, cosf(angle))), scl(y_, sinf(angle)));
}

struct material_t { float impedance, attenuation, mu0, mu1, sigma, specularity, shininess, thickness; };   // mesh.h:7-10
struct mesh_t { int mat_in, mat_out; bool vascular; v3 deltas; v3 origin; int64_t tri_begin, tri_end; };    // mesh.h:12-20

struct bvh_node { float lo[3], hi[3]; int32_t left, right; int32_t first, count; };

constexpr int OUTSIDE_NULL = -1;
constexpr int OUTSIDE_SELF = -2;

int g_threads = 1;

}  // namespace

struct orc_scene {
    std::vector<material_t> materials;
    int starting_material = 0;
    std::vector<mesh_t> meshes;
    std::vector<float> tri_local;      // 9 floats per triangle: v_obj * scaling
    std::vector<int32_t> tri_mesh;
    float scaling = 1.0f;
    float origin[3] = {0, 0, 0};
    float spacing[3] = {1, 1, 1};
    // the oracle's own acceleration structure (conservative; results are traversal-order independent)
    std::vector<bvh_node> nodes;
    std::vector<int32_t> bvh_tris;
    float max_abs = 0.0f;
};

struct orc_volume {
    std::vector<float> data;   // [256][256][256][2] = {texture_noise, scattering_probability}
};

namespace {

// ------------------------------------------------------------------------------------------------
// Closest hit: btCollisionWorld::rayTest + ClosestRayResultCallback (scene.cpp:115-126), restated
// from bullet3's btTriangleRaycastCallback::processTriangle.  PARITY UNPINNED.
// ------------------------------------------------------------------------------------------------
struct hit_t { float fraction; int32_t tri; int32_t mesh; v3 normal; };

// One triangle, in the body's local frame.  Ties on the fraction resolve to the smallest global
// triangle id so that no traversal order can matter (Bullet itself keeps the first one visited).
inline void test_triangle(const orc_scene& s, int32_t tri, v3 from_l, v3 to_l, hit_t& best)
{
    const float* p = &s.tri_local[(size_t)tri * 9];
    const v3 vert0 = mk(p[0], p[1], p[2]), vert1 = mk(p[3], p[4], p[5]), vert2 = mk(p[6], p[7], p[8]);
    const v3 v10 = sub(vert1, vert0);
    const v3 v20 = sub(vert2, vert0);
    v3 n = cross(v10, v20);
    const float dist = dot(vert0, n);
    float dist_a = dot(n, from_l);
    dist_a -= dist;
    float dist_b = dot(n, to_l);
    dist_b -= dist;
    if (dist_a * dist_b >= 0.0f) return;                      // same side
    const float proj_length = dist_a - dist_b;
    const float distance = dist_a / proj_length;
    if (!(distance < best.fraction || (distance == best.fraction && tri < best.tri))) return;
    float edge_tolerance = dot(n, n);
    edge_tolerance *= -0.0001f;
    const v3 point = interpolate3(from_l, to_l, distance);
    const v3 v0p = sub(vert0, point);
    const v3 v1p = sub(vert1, point);
    const v3 cp0 = cross(v0p, v1p);
    if (dot(cp0, n) >= edge_tolerance) {
        const v3 v2p = sub(vert2, point);
        const v3 cp1 = cross(v1p, v2p);
        if (dot(cp1, n) >= edge_tolerance) {
            const v3 cp2 = cross(v2p, v0p);
            if (dot(cp2, n) >= edge_tolerance) {
                n = normalized(n);
                best.fraction = distance;
                best.tri = tri;
                best.mesh = s.tri_mesh[tri];
                best.normal = (dist_a <= 0.0f) ? neg(n) : n;  // face the ray origin
            }
        }
    }
}

inline void to_local(const orc_scene& s, int mesh, v3 from_w, v3 to_w, v3& from_l, v3& to_l)
{
    // worldTocollisionObject * p with an identity basis = p - body origin (one rounding per component)
    const v3 o = s.meshes[mesh].origin;
    from_l = sub(from_w, o);
    to_l = sub(to_w, o);
}

hit_t closest_hit_brute(const orc_scene& s, v3 from_w, v3 to_w)
{
    hit_t best; best.fraction = 1.0f; best.tri = -1; best.mesh = -1; best.normal = mk(0, 0, 0);
    for (size_t m = 0; m < s.meshes.size(); m++) {
        v3 fl, tl; to_local(s, (int)m, from_w, to_w, fl, tl);
        for (int64_t t = s.meshes[m].tri_begin; t < s.meshes[m].tri_end; t++) test_triangle(s, (int32_t)t, fl, tl, best);
    }
    return best;
}

// ---- oracle BVH (median split over world-space boxes, padded so culling is conservative) --------
void tri_world_box(const orc_scene& s, int32_t tri, float lo[3], float hi[3])
{
    const float* p = &s.tri_local[(size_t)tri * 9];
    const v3 o = s.meshes[s.tri_mesh[tri]].origin;
    for (int a = 0; a < 3; a++) { lo[a] = 3.0e38f; hi[a] = -3.0e38f; }
    for (int k = 0; k < 3; k++) {
        const float w[3] = {p[3 * k] + o.x, p[3 * k + 1] + o.y, p[3 * k + 2] + o.z};
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], w[a]); hi[a] = std::max(hi[a], w[a]); }
    }
}

int32_t build_node(orc_scene& s, std::vector<float>& cent, int32_t first, int32_t count)
{
    bvh_node nd; nd.left = nd.right = -1; nd.first = first; nd.count = count;
    for (int a = 0; a < 3; a++) { nd.lo[a] = 3.0e38f; nd.hi[a] = -3.0e38f; }
    float clo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, chi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int32_t i = first; i < first + count; i++) {
        float lo[3], hi[3]; tri_world_box(s, s.bvh_tris[i], lo, hi);
        for (int a = 0; a < 3; a++) {
            nd.lo[a] = std::min(nd.lo[a], lo[a]); nd.hi[a] = std::max(nd.hi[a], hi[a]);
            const float c = cent[(size_t)s.bvh_tris[i] * 3 + a];
            clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c);
        }
    }
    const int32_t id = (int32_t)s.nodes.size();
    s.nodes.push_back(nd);
    if (count > 4) {
        int axis = 0;
        if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
        if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
        const int32_t mid = first + count / 2;
        std::nth_element(s.bvh_tris.begin() + first, s.bvh_tris.begin() + mid, s.bvh_tris.begin() + first + count,
                         [&](int32_t a, int32_t b) {
                             const float ca = cent[(size_t)a * 3 + axis], cb = cent[(size_t)b * 3 + axis];
                             return ca < cb || (ca == cb && a < b);
                         });
        const int32_t l = build_node(s, cent, first, mid - first);
        const int32_t r = build_node(s, cent, mid, first + count - mid);
        s.nodes[id].left = l; s.nodes[id].right = r; s.nodes[id].count = 0;
    }
    return id;
}

void build_bvh(orc_scene& s)
{
    const int64_t n = (int64_t)s.tri_mesh.size();
    s.bvh_tris.resize(n);
    std::vector<float> cent((size_t)n * 3);
    s.max_abs = 0.0f;
    for (int64_t t = 0; t < n; t++) {
        s.bvh_tris[t] = (int32_t)t;
        float lo[3], hi[3]; tri_world_box(s, (int32_t)t, lo, hi);
        for (int a = 0; a < 3; a++) {
            cent[(size_t)t * 3 + a] = 0.5f * (lo[a] + hi[a]);
            s.max_abs = std::max(s.max_abs, std::max(std::fabs(lo[a]), std::fabs(hi[a])));
        }
    }
    s.nodes.clear();
    if (n > 0) { s.nodes.reserve((size_t)n); build_node(s, cent, 0, (int32_t)n); }
}

// Conservative segment/box test: the box is inflated by `pad` (absolute) and the parametric
// interval by a relative slack, far more than any rounding of the exact per-triangle arithmetic.
inline bool box_overlap(const bvh_node& nd, const double o[3], const double inv[3], double pad, double tmax)
{
    double t0 = 0.0, t1 = tmax;
    for (int a = 0; a < 3; a++) {
        double lo = ((double)nd.lo[a] - pad - o[a]) * inv[a];
        double hi = ((double)nd.hi[a] + pad - o[a]) * inv[a];
        if (lo != lo || hi != hi) continue;      // 0 * inf: the origin lies exactly on the slab plane
        if (lo > hi) std::swap(lo, hi);
        t0 = std::max(t0, lo); t1 = std::min(t1, hi);
    }
    return t0 <= t1 * (1.0 + 1e-6) + 1e-30;
}

hit_t closest_hit_bvh(const orc_scene& s, v3 from_w, v3 to_w)
{
    hit_t best; best.fraction = 1.0f; best.tri = -1; best.mesh = -1; best.normal = mk(0, 0, 0);
    if (s.nodes.empty()) return best;
    const double o[3] = {from_w.x, from_w.y, from_w.z};
    const double d[3] = {(double)to_w.x - from_w.x, (double)to_w.y - from_w.y, (double)to_w.z - from_w.z};
    double inv[3];
    for (int a = 0; a < 3; a++) inv[a] = 1.0 / d[a];
    const double mo = std::max(std::fabs(o[0]), std::max(std::fabs(o[1]), std::fabs(o[2])));
    const double pad = 1e-5 * ((double)s.max_abs + mo) + 1e-6;
    // cached per-mesh local rays
    std::vector<v3> fl(s.meshes.size()), tl(s.meshes.size());
    for (size_t m = 0; m < s.meshes.size(); m++) to_local(s, (int)m, from_w, to_w, fl[m], tl[m]);
    int32_t stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const bvh_node& nd = s.nodes[stack[--sp]];
        const double tmax = (double)best.fraction * (1.0 + 1e-5) + 1e-30;
        if (!box_overlap(nd, o, inv, pad, tmax)) continue;
        if (nd.left < 0) {
            for (int32_t i = nd.first; i < nd.first + nd.count; i++) {
                const int32_t t = s.bvh_tris[i];
                const int m = s.tri_mesh[t];
                test_triangle(s, t, fl[m], tl[m], best);
            }
        } else {
            stack[sp++] = nd.left; stack[sp++] = nd.right;
        }
    }
    return best;
}

// ------------------------------------------------------------------------------------------------
// ray_physics (ray.cpp)
// ------------------------------------------------------------------------------------------------
constexpr float INTENSITY_EPSILON = 1e-10;   // ray.h:24 (double literal narrowed to float)

struct path_t {                 // ray_physics::ray, ray.h:13-26
    v3 from, direction;
    int depth;
    int media;                  // material id; the reference keeps the material by value
    int media_outside;          // OUTSIDE_NULL, OUTSIDE_SELF (= &own media slot, ray.cpp:38) or material id
    float intensity, frequency;
    double distance_traveled;   // mm
    bool null;
};

inline float max_ray_length(float attenuation, float intensity, float frequency)          // ray.cpp:110-113
{
    return 10.f * mc_logf(INTENSITY_EPSILON / intensity) / -attenuation * frequency;
}

inline void travel(float attenuation, float frequency, float& intensity, double& dist, double mm)   // ray.cpp:99-103
{
    dist = dist + mm;
    intensity = intensity * mc_expf(-attenuation * ((float)mm * 0.01f) * frequency);
}

inline v3 snells_law(v3 l, v3 n, float c, float refraction_angle, float r)                // ray.cpp:115-124
{
    return add(scl(l, r), scl(n, r * c - refraction_angle));
}

inline float reflection_intensity(float intensity_in, float media_1, float incidence_angle, float media_2, float refracted_angle)
{                                                                                          // ray.cpp:126-132
    const float num = media_1 * incidence_angle - media_2 * refracted_angle;
    const float denom = media_1 * incidence_angle + media_2 * refracted_angle;
    const double q = (double)(num / denom);
    return (float)((double)intensity_in * (q * q));     // intensity_in * pow(num/denom, 2) in double
}

// Eq. 8 (ray.cpp:154-164).  B-5: a NaN factor (TIR refraction direction, negative base with a
// non-integer specularity) contributes 0 instead of poisoning the image.
inline float reflected_intensity_eq8(v3 direction, v3 refraction_direction, v3 reflection_direction, float specularity)
{
    const float refraction_angle = dot(direction, refraction_direction);
    float refraction_factor = mc_powf(refraction_angle, specularity);
    const float reflection_angle = dot(direction, reflection_direction);
    float reflection_factor = mc_powf(reflection_angle, specularity);
    if (refraction_factor != refraction_factor) refraction_factor = 0.0f;
    if (reflection_factor != reflection_factor) reflection_factor = 0.0f;
    return std::max(refraction_factor, 0.0f) + std::max(reflection_factor, 0.0f);
}

// ray.cpp:213-224 with the uniform supplied by the caller; `v` is the truncated shininess (B-8)
inline float power_cosine_variate(int v, double number)
{
    const int indice = v + 1;
    const float exponente = (float)((double)1.0 / indice);
    return (float)mc_pow(number, (double)exponente);
}

// ray.cpp:167-211 with the two uniforms of one disk-sampling attempt supplied by the caller.
// Returns false if the attempt is rejected by `while (!(p <= 0.25))`.
inline bool random_unit_vector_attempt(v3 v, float cos_theta, double u_az, double u_rad, v3& out)
{
    const double a = u_az * 2 * MC_PI_D;
    const double r = 0.5 * sqrt(u_rad);
    double sn, cs; mc_sincos(a, &sn, &cs);
    float px = (float)(r * cs);
    float py = (float)(r * sn);
    const float p = px * px + py * py;
    if (!(p <= 0.25f)) return false;
    bool flag = false;
    float vx = v.x, vy = v.y;
    const float vz = v.z;
    if (fabsf(vx) > fabsf(vy)) { vx = vy; vy = v.x; flag = true; }     // B-7: float abs
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
    out = mk(wx, wy, wz);
    return true;
}

struct boundary_result { float reflected_intensity; path_t returned; int branch; };

// ray.cpp:11-97.  rng: cos_theta (already drawn, or 1 in deterministic mode), the jittered normal,
// and the reflect/refract uniform are supplied by the caller so the function itself is pure.
boundary_result hit_boundary(const orc_scene& s, const path_t& r, v3 hit_point, v3 random_normal, float random_angle,
                             const mesh_t& collided_mesh, int material_after_collision, int material_after_vascularities,
                             float x, int force_branch)
{
    const material_t& mac = s.materials[material_after_collision];
    const material_t& media = s.materials[r.media];
    float incidence_angle = dot(r.direction, neg(random_normal));           // ray.cpp:53
    if (incidence_angle < 0) incidence_angle = dot(r.direction, random_normal);   // B-6
    const float refr_ratio = media.impedance / mac.impedance;
    float refraction_angle = 1 - refr_ratio * refr_ratio * (1 - incidence_angle * incidence_angle);
    const bool total_internal_reflection = refraction_angle < 0;
    refraction_angle = sqrtf(refraction_angle);
    v3 refraction_direction = snells_law(r.direction, random_normal, incidence_angle, refraction_angle, refr_ratio);
    refraction_direction = normalized(refraction_direction);
    v3 reflection_direction = add(r.direction, scl(random_normal, 2 * incidence_angle));
    reflection_direction = normalized(reflection_direction);
    const float intensity_refl = total_internal_reflection
                                     ? r.intensity
                                     : reflection_intensity(r.intensity, media.impedance, incidence_angle, mac.impedance, refraction_angle);
    const float intensity_refr = r.intensity - intensity_refl;
    const float back = reflected_intensity_eq8(r.direction, refraction_direction, reflection_direction, mac.specularity) * random_angle;
    const float reflection_probability = intensity_refl / r.intensity;
    bool reflect = reflection_probability > x;
    if (force_branch >= 0) reflect = force_branch != 0;
    boundary_result out;
    out.reflected_intensity = back;
    out.branch = reflect ? 1 : 0;
    path_t& q = out.returned;
    q.from = hit_point;
    q.depth = r.depth + 1;
    q.frequency = r.frequency;
    q.distance_traveled = r.distance_traveled;
    q.null = false;
    if (reflect) {
        q.direction = reflection_direction; q.media = r.media; q.media_outside = r.media_outside;
        q.intensity = intensity_refl > INTENSITY_EPSILON ? intensity_refl : 0.0f;
    } else {
        q.direction = refraction_direction; q.media = material_after_collision; q.media_outside = material_after_vascularities;
        q.intensity = intensity_refr > INTENSITY_EPSILON ? intensity_refr : 0.0f;
    }
    (void)collided_mesh;
    return out;
}

// The medium state machine of ray.cpp:14-47 as it actually behaves:
// `&r.media == &collided_mesh.material_inside` is never true, and `&r.media` stored as
// media_outside aliases the path's own slot.
inline void medium_after(const path_t& r, const mesh_t& m, int& after_collision, int& after_vascularities)
{
    if (r.media_outside != OUTSIDE_NULL) {                 // in a vessel
        if (m.vascular) {
            after_vascularities = OUTSIDE_NULL;
            after_collision = (r.media_outside == OUTSIDE_SELF) ? r.media : r.media_outside;
        } else {
            after_vascularities = (r.media_outside == m.mat_in) ? m.mat_out : m.mat_in;   // SELF never equals a map address
            after_collision = r.media;
        }
    } else {
        if (m.vascular) {
            after_vascularities = OUTSIDE_SELF;
            after_collision = m.mat_in;
        } else {
            after_vascularities = OUTSIDE_NULL;
            after_collision = m.mat_in;                    // B-2: the address compare is always false
        }
    }
}

inline double distance_in_mm(const orc_scene& s, v3 v1, v3 v2)                               // scene.cpp:281-290
{
    const float x_dist = fabsf(v1.x - v2.x) * s.spacing[0];
    const float y_dist = fabsf(v1.y - v2.y) * s.spacing[1];
    const float z_dist = fabsf(v1.z - v2.z) * s.spacing[2];
    const double xd = x_dist, yd = y_dist, zd = z_dist;
    return sqrt(xd * xd + yd * yd + zd * zd) * 10;
}

inline v3 enlarge(const orc_scene& s, v3 versor, float mm)                                    // scene.cpp:292-298
{
    return scl(mk(s.spacing[0] * versor.x, s.spacing[1] * versor.y, s.spacing[2] * versor.z), mm / 100.0f);
}

// ------------------------------------------------------------------------------------------------
// transducer.h:24-62
// ------------------------------------------------------------------------------------------------
inline double deg_to_rad(double deg) { return ((deg * (MC_PI_D * 1.0) * 1) / 180); }   // units.h:1375, Ratio 1/180, PiRatio 1

void transducer_elements(const orc_params& p, const orc_derived& dv, const float* pos3, const float* ang3, float* out_pos, float* out_dir)
{
    const float x_angle = (float)deg_to_rad((double)ang3[0]);
    const float y_angle = (float)deg_to_rad((double)ang3[1]);
    const float z_angle = (float)deg_to_rad((double)ang3[2]);
    // amp = transducer_element_separation / radius : mm / cm, converted to a scalar with ratio 1/10
    const double amp_raw = dv.element_separation_mm / p.radius_cm;
    const float amp_f = (float)((amp_raw * 1) / 10);
    const double amplitude = (double)amp_f;
    const double angle_center_of_element = amplitude / 2.0f;
    double angle = -(amplitude * (double)(size_t)p.elements / 2) + angle_center_of_element;
    const float radius_f = (float)p.radius_cm;
    const v3 P = mk(pos3[0], pos3[1], pos3[2]);
    for (int t = 0; t < p.elements; t++) {
        const float af = (float)angle;
        v3 d = mk(sinf(af), cosf(af), 0);
        d = rotate(d, mk(0, 0, 1), z_angle);
        d = rotate(d, mk(1, 0, 0), x_angle);
        d = rotate(d, mk(0, 1, 0), y_angle);
        const v3 q = add(P, scl(d, radius_f));
        out_pos[3 * t] = q.x; out_pos[3 * t + 1] = q.y; out_pos[3 * t + 2] = q.z;
        out_dir[3 * t] = d.x; out_dir[3 * t + 1] = d.y; out_dir[3 * t + 2] = d.z;
        angle = angle + amplitude;
    }
}

// ------------------------------------------------------------------------------------------------
// volume.h:46-61
// ------------------------------------------------------------------------------------------------
inline uint32_t voxel_index(float coord, float resolution)
{
    // static_cast<unsigned>(negative float) is UB; x86-64/GCC emits cvttss2si (64-bit) and keeps the
    // low 32 bits.  `% 256` of that is `& 255`.
    const float q = coord / resolution;
    const int64_t w = (int64_t)q;
    return (uint32_t)w & 255u;
}

inline float get_scattering(const orc_volume& v, float resolution, float density, float mu, float sigma, float x, float y, float z)
{
    const uint32_t xi = voxel_index(x, resolution), yi = voxel_index(y, resolution), zi = voxel_index(z, resolution);
    const float* vox = &v.data[(((size_t)xi * 256 + yi) * 256 + zi) * 2];
    return vox[1] >= density ? vox[0] * sigma + mu : 0.0f;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

void orc_set_threads(int32_t n) { g_threads = n < 1 ? 1 : n; }
int32_t orc_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_default_params(orc_params* p)
{
    memset(p, 0, sizeof(*p));
    p->elements = 512; p->samples = 5; p->max_depth = 10; p->frequency_mhz = 4.5f;
    p->radius_cm = 3; p->fov_deg = 60; p->depth_cm = 15; p->speed_of_sound = 1500; p->resolution_um = 145;
    p->psf_axial = 7; p->psf_lateral = 13; p->psf_var_x = 0.05f; p->psf_var_y = 0.2f;
    p->deterministic = 0; p->scan_rows = 400; p->scan_cols = 500; p->axial_scale = 1.0f;
}

void orc_derive(const orc_params* p, orc_derived* d)
{
    const float axres_f = (1.45f / p->frequency_mhz) / (p->axial_scale > 0.0f ? p->axial_scale : 1.0f);   // main.cpp:25
    d->axial_resolution_mm = (double)axres_f;
    d->axial_resolution_f = (float)d->axial_resolution_mm;
    // microsecond_t(centimeter_t / meters_per_second_t): raw quotient, then ratio 10000/1 (units.h:1365)
    d->max_travel_time_us = ((p->depth_cm / (double)p->speed_of_sound) * 10000) / 1;
    d->max_travel_time_u = (uint32_t)d->max_travel_time_us;
    d->rf_axial_um = (uint32_t)(d->axial_resolution_f * 1000.0f);
    d->rows = (int32_t)((p->speed_of_sound * d->max_travel_time_u) / d->rf_axial_um);           // rfimage.h:180
    d->cols = p->elements;
    // main.cpp:66: transducer_amplitude.to<float>() * transducer_radius / transducer_elements  -> mm
    const double amplitude_rad = ((p->fov_deg * (MC_PI_D * 1.0) * 1) / 180);
    const float amplitude_f = (float)amplitude_rad;
    const double sep_cm = ((double)amplitude_f * p->radius_cm) / (double)(size_t)p->elements;
    d->element_separation_mm = (sep_cm * 10) / 1;
    d->time_step_us = ((d->axial_resolution_mm * 1000) / 1) / (double)p->speed_of_sound;       // rfimage.h:48-51
    d->row_period_us = (double)d->rf_axial_um / (double)p->speed_of_sound;                     // rfimage.h:35
}

orc_scene* orc_scene_create(int32_t n_mat, const float* materials8, int32_t starting_material, int32_t n_mesh,
                            const int32_t* mesh_in, const int32_t* mesh_out, const int32_t* mesh_vascular,
                            const float* mesh_deltas, const int64_t* tri_offsets, const float* tri_vertices_obj,
                            float scaling, const float* origin3, const float* spacing3)
{
    orc_scene* s = new orc_scene();
    s->materials.resize(n_mat);
    for (int i = 0; i < n_mat; i++) memcpy(&s->materials[i], materials8 + 8 * i, sizeof(material_t));
    s->starting_material = starting_material;
    s->scaling = scaling;
    for (int a = 0; a < 3; a++) { s->origin[a] = origin3[a]; s->spacing[a] = spacing3[a]; }
    const int64_t n_tri = tri_offsets[n_mesh];
    s->tri_local.resize((size_t)n_tri * 9);
    s->tri_mesh.resize((size_t)n_tri);
    s->meshes.resize(n_mesh);
    for (int m = 0; m < n_mesh; m++) {
        mesh_t& me = s->meshes[m];
        me.mat_in = mesh_in[m]; me.mat_out = mesh_out[m]; me.vascular = mesh_vascular[m] != 0;
        me.deltas = mk(mesh_deltas[3 * m], mesh_deltas[3 * m + 1], mesh_deltas[3 * m + 2]);
        // scene.cpp:322-323: pos = deltas*scaling*scaling; position = pos + origin
        const float px = me.deltas.x * scaling * scaling, py = me.deltas.y * scaling * scaling, pz = me.deltas.z * scaling * scaling;
        me.origin = mk(px + origin3[0], py + origin3[1], pz + origin3[2]);
        me.tri_begin = tri_offsets[m]; me.tri_end = tri_offsets[m + 1];
        for (int64_t t = me.tri_begin; t < me.tri_end; t++) {
            s->tri_mesh[t] = m;
            // btBvhTriangleMeshShape fetches vertices as v_obj * localScaling (scene.cpp:313-316)
            for (int k = 0; k < 9; k++) s->tri_local[(size_t)t * 9 + k] = tri_vertices_obj[(size_t)t * 9 + k] * scaling;
        }
    }
    build_bvh(*s);
    return s;
}

void orc_scene_destroy(orc_scene* s) { delete s; }
int64_t orc_scene_num_triangles(const orc_scene* s) { return (int64_t)s->tri_mesh.size(); }
void orc_scene_get_local_vertices(const orc_scene* s, float* out9) { memcpy(out9, s->tri_local.data(), s->tri_local.size() * sizeof(float)); }
void orc_scene_get_mesh_origins(const orc_scene* s, float* out3)
{
    for (size_t m = 0; m < s->meshes.size(); m++) { out3[3 * m] = s->meshes[m].origin.x; out3[3 * m + 1] = s->meshes[m].origin.y; out3[3 * m + 2] = s->meshes[m].origin.z; }
}

int32_t orc_closest_hit(const orc_scene* s, const float* from3, const float* to3, int32_t use_bvh, float* out_f7, int32_t* out_mesh)
{
    const v3 f = mk(from3[0], from3[1], from3[2]), t = mk(to3[0], to3[1], to3[2]);
    const hit_t h = use_bvh ? closest_hit_bvh(*s, f, t) : closest_hit_brute(*s, f, t);
    if (out_f7) {
        out_f7[0] = h.fraction;
        const v3 p = interpolate3(f, t, h.fraction);        // ClosestRayResultCallback::addSingleResult
        out_f7[1] = p.x; out_f7[2] = p.y; out_f7[3] = p.z;
        out_f7[4] = h.normal.x; out_f7[5] = h.normal.y; out_f7[6] = h.normal.z;
    }
    if (out_mesh) *out_mesh = h.mesh;
    return h.tri;
}

void orc_transducer_elements(const orc_params* p, const float* pos3, const float* angles_deg3, float* out_pos, float* out_dir)
{
    orc_derived d; orc_derive(p, &d);
    transducer_elements(*p, d, pos3, angles_deg3, out_pos, out_dir);
}

void orc_psf_taps(const orc_params* p, float* axial, float* lateral)
{
    // psf.h:34-58, 80-92 (M_PI redefined to 3.14159, psf.h:9); libm exp/cos in double as the reference
    const float half_axial = (size_t)p->psf_axial * (size_t)p->resolution_um / 1000.0f / 2.0f;
    const float half_lateral = (size_t)p->psf_lateral * (size_t)p->resolution_um / 1000.0f / 2.0f;
    const float resolution = p->resolution_um / 1000.0f;
    const float freq = p->frequency_mhz;
    for (int i = 0; i < p->psf_axial; i++) {
        const float x = (size_t)i * resolution - half_axial;
        axial[i] = (float)(exp(-0.5f * (((double)x * (double)x) / p->psf_var_x)) * cos(2 * MC_PI_REDEFINED * freq * x));
    }
    for (int i = 0; i < p->psf_lateral; i++) {
        const float y = (size_t)i * resolution - half_lateral;
        lateral[i] = (float)exp(-0.5f * (((double)y * (double)y) / p->psf_var_y));
    }
}

// Elevational PSF: psf.h:42,77 declares elevation_kernel and never fills it; filled here by analogy with
// lateral_function (psf.h:87-92) on the grid of psf.h:46-57.  z_mm[i] = elevational offset of ray-fan plane i.
void orc_elevation_taps(const orc_params* p, int32_t n, float var_z, float* taps, float* z_mm)
{
    const float half_elevation = (size_t)n * (size_t)p->resolution_um / 1000.0f / 2.0f;
    const float resolution = p->resolution_um / 1000.0f;
    for (int i = 0; i < n; i++) {
        const float z = (size_t)i * resolution - half_elevation;
        taps[i] = (float)exp(-0.5f * (((double)z * (double)z) / var_z));
        z_mm[i] = z;
    }
}

// position of the probe whose fan lies z_mm off the imaging plane: the fan plane's normal is the transducer's local z axis
// pushed through the same three btVector3::rotate calls as the element directions (transducer.h:51-56)
void orc_elevation_position(const float* pos3, const float* angles_deg3, float z_mm, float* out_pos3)
{
    const float x_angle = (float)deg_to_rad((double)angles_deg3[0]);
    const float y_angle = (float)deg_to_rad((double)angles_deg3[1]);
    const float z_angle = (float)deg_to_rad((double)angles_deg3[2]);
    v3 e = mk(0, 0, 1);
    e = rotate(e, mk(0, 0, 1), z_angle);
    e = rotate(e, mk(1, 0, 0), x_angle);
    e = rotate(e, mk(0, 1, 0), y_angle);
    const float s = z_mm * 0.1f;                        // mm -> cm world units
    out_pos3[0] = pos3[0] + e.x * s; out_pos3[1] = pos3[1] + e.y * s; out_pos3[2] = pos3[2] + e.z * s;
}

orc_volume* orc_volume_get(void)
{
    // volume.h:19-35: default-seeded std::default_random_engine + std::normal_distribution<double>,
    // fill order i,j,k with texture_noise then scattering_probability.  The stream is
    // libstdc++-specific, so it is produced with the very same <random> calls.
    static orc_volume* vol = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        vol = new orc_volume();
        const size_t n = (size_t)256 * 256 * 256;
        vol->data.resize(n * 2);
        std::default_random_engine generator;
        std::normal_distribution<double> distribution(0.0, 1.0);
        for (size_t i = 0; i < n; i++) {
            vol->data[2 * i] = (float)distribution(generator);
            vol->data[2 * i + 1] = (float)distribution(generator);
        }
    });
    return vol;
}
const float* orc_volume_raw(const orc_volume* v) { return v->data.data(); }
float orc_volume_get_scattering(const orc_volume* v, float density, float mu, float sigma, float x, float y, float z)
{
    return get_scattering(*v, 145 / 1000.0f, density, mu, sigma, x, y, z);
}

float orc_max_ray_length(float attenuation, float intensity, float frequency) { return max_ray_length(attenuation, intensity, frequency); }
void orc_travel(float attenuation, float intensity, float frequency, double dist0, double mm, float* out_i, double* out_d)
{
    float i = intensity; double d = dist0; travel(attenuation, frequency, i, d, mm); *out_i = i; *out_d = d;
}
float orc_reflection_intensity(float i_in, float z1, float c1, float z2, float c2) { return reflection_intensity(i_in, z1, c1, z2, c2); }
float orc_reflected_intensity_eq8(const float* d, const float* a, const float* b, float specularity)
{
    return reflected_intensity_eq8(mk(d[0], d[1], d[2]), mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]), specularity);
}
void orc_snells_law(const float* l, const float* n, float c1, float c2, float ratio, float* out)
{
    const v3 r = snells_law(mk(l[0], l[1], l[2]), mk(n[0], n[1], n[2]), c1, c2, ratio);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_random_unit_vector(const float* v, float cos_theta, double u_az, double u_rad, float* out)
{
    v3 w = mk(0, 0, 0);
    if (!random_unit_vector_attempt(mk(v[0], v[1], v[2]), cos_theta, u_az, u_rad, w)) w = mk(NAN, NAN, NAN);
    out[0] = w.x; out[1] = w.y; out[2] = w.z;
}
double orc_distance_in_mm(const orc_scene* s, const float* a, const float* b)
{
    return distance_in_mm(*s, mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]));
}

void orc_hit_boundary(const orc_scene* s, const float* from3, const float* dir3, float intensity, const int32_t* state_i3,
                      const float* hit_point3, const float* normal3, int32_t mesh_id, int32_t deterministic,
                      const double* rng_u4, int32_t force_branch, float* out_f8, int32_t* out_i3, int32_t* out_branch)
{
    path_t r;
    r.from = mk(from3[0], from3[1], from3[2]); r.direction = mk(dir3[0], dir3[1], dir3[2]);
    r.media = state_i3[0]; r.media_outside = state_i3[1]; r.depth = state_i3[2];
    r.intensity = intensity; r.frequency = 4.5f; r.distance_traveled = 0; r.null = false;
    const mesh_t& m = s->meshes[mesh_id];
    int mac, mav; medium_after(r, m, mac, mav);
    const v3 n = mk(normal3[0], normal3[1], normal3[2]);
    float random_angle = 1.0f; v3 rn = n;
    if (!deterministic) {
        random_angle = power_cosine_variate((int)s->materials[mac].shininess, rng_u4[0]);
        if (!random_unit_vector_attempt(n, random_angle, rng_u4[2], rng_u4[3], rn)) rn = n;
    }
    const boundary_result br = hit_boundary(*s, r, mk(hit_point3[0], hit_point3[1], hit_point3[2]), rn, random_angle, m, mac, mav,
                                            (float)rng_u4[1], force_branch);
    out_f8[0] = br.reflected_intensity;
    out_f8[1] = br.returned.from.x; out_f8[2] = br.returned.from.y; out_f8[3] = br.returned.from.z;
    out_f8[4] = br.returned.direction.x; out_f8[5] = br.returned.direction.y; out_f8[6] = br.returned.direction.z;
    out_f8[7] = br.returned.intensity;
    out_i3[0] = br.returned.depth; out_i3[1] = br.returned.media; out_i3[2] = br.returned.media_outside;
    *out_branch = br.branch;
}

// ------------------------------------------------------------------------------------------------
// scene::cast_rays<S,E> (scene.cpp:50-183)
// ------------------------------------------------------------------------------------------------
int64_t orc_cast_rays(const orc_scene* sp, const orc_params* pp, const float* pos3, const float* angles_deg3, uint64_t seed,
                      uint32_t frame, int32_t use_bvh, orc_segment* segments, int32_t* n_segments)
{
    const orc_scene& s = *sp; const orc_params& p = *pp;
    orc_derived dv; orc_derive(&p, &dv);
    const int E = p.elements, S = p.samples, D = p.max_depth;
    std::vector<float> epos((size_t)E * 3), edir((size_t)E * 3);
    transducer_elements(p, dv, pos3, angles_deg3, epos.data(), edir.data());
    int64_t tests = 0;
    const float initial_intensity = 1.0f;                                           // scene.h:49
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(g_threads) reduction(+ : tests) if (g_threads > 1)
#endif
    for (int ray_i = 0; ray_i < E; ray_i++) {
        std::vector<path_t> samples(S);
        path_t first_ray;                                                           // scene.cpp:84-96
        first_ray.from = mk(epos[3 * ray_i], epos[3 * ray_i + 1], epos[3 * ray_i + 2]);
        first_ray.direction = mk(edir[3 * ray_i], edir[3 * ray_i + 1], edir[3 * ray_i + 2]);
        first_ray.depth = 0;
        first_ray.media = s.starting_material;
        first_ray.media_outside = OUTSIDE_NULL;
        first_ray.intensity = initial_intensity / (float)(unsigned)S;
        first_ray.frequency = p.frequency_mhz;
        first_ray.distance_traveled = 0;
        first_ray.null = false;
        for (int sample_i = 0; sample_i < S; sample_i++) { samples[sample_i] = first_ray; n_segments[(size_t)ray_i * S + sample_i] = 0; }
        for (int i = 0; i < D; i++) {
            for (int sample_i = 0; sample_i < S; sample_i++) {
                path_t& ray_ = samples[sample_i];
                if (ray_.null) continue;
                const material_t& media = s.materials[ray_.media];
                orc_segment* seg_out = &segments[((size_t)ray_i * S + sample_i) * D + n_segments[(size_t)ray_i * S + sample_i]];
                const float r_length = max_ray_length(media.attenuation, ray_.intensity, ray_.frequency);
                const v3 to = add(ray_.from, enlarge(s, ray_.direction, r_length));
                const v3 from_test = add(ray_.from, scl(ray_.direction, 0.1f));     // scene.cpp:115-117
                const hit_t h = use_bvh ? closest_hit_bvh(s, from_test, to) : closest_hit_brute(s, from_test, to);
                tests++;
                auto put = [&](v3 seg_to, float refl, float init_i, double dist_before, int tri, int mesh, float frac) {
                    seg_out->from[0] = ray_.from.x; seg_out->from[1] = ray_.from.y; seg_out->from[2] = ray_.from.z;
                    seg_out->to[0] = seg_to.x; seg_out->to[1] = seg_to.y; seg_out->to[2] = seg_to.z;
                    seg_out->dir[0] = ray_.direction.x; seg_out->dir[1] = ray_.direction.y; seg_out->dir[2] = ray_.direction.z;
                    seg_out->reflected_intensity = refl; seg_out->initial_intensity = init_i;
                    seg_out->attenuation = media.attenuation; seg_out->distance_traveled = dist_before;
                    seg_out->media_id = ray_.media; seg_out->tri_id = tri; seg_out->mesh_id = mesh; seg_out->hit_fraction = frac;
                    n_segments[(size_t)ray_i * S + sample_i]++;
                };
                if (h.tri >= 0) {
                    const double distance_before_hit = ray_.distance_traveled;
                    const float intensity_before_hit = ray_.intensity;
                    const mesh_t& organ = s.meshes[h.mesh];
                    const v3 hit_point = interpolate3(from_test, to, h.fraction);   // m_hitPointWorld
                    // RNG (B-11): block 0 = {thickness u1, thickness u2, shininess u, choice u}
                    const mc_u32x4 b0 = mc_rng_block(seed, frame, (uint32_t)ray_i, (uint32_t)sample_i, (uint32_t)i, 0);
                    // scene.cpp:132-135: q = |N(0, thickness_inside)|, Box-Muller on (u1,u2)
                    float q = 0.0f;
                    const float thickness = s.materials[organ.mat_in].thickness;
                    if (!p.deterministic && thickness != 0.0f) {                    // B-13: sigma 0 -> q = 0
                        const double u1 = mc_u01d(b0.v[0]), u2 = mc_u01d(b0.v[1]);
                        double sn, cs; mc_sincos(2 * MC_PI_D * u2, &sn, &cs);
                        const double z = sqrt(-2.0 * mc_log(u1)) * cs;
                        q = (float)fabs(z * (double)thickness);
                    }
                    const v3 inside_point = add(scl(ray_.direction, q), hit_point); // scene.cpp:139
                    travel(media.attenuation, ray_.frequency, ray_.intensity, ray_.distance_traveled, distance_in_mm(s, ray_.from, inside_point));
                    int mac, mav; medium_after(ray_, organ, mac, mav);
                    float random_angle = 1.0f; v3 random_normal = h.normal;
                    if (!p.deterministic) {
                        random_angle = power_cosine_variate((int)s.materials[mac].shininess, mc_u01d(b0.v[2]));
                        bool ok = false;
                        for (uint32_t attempt = 0; attempt < MC_RNG_BLOCKS_PER_BOUNCE - 1 && !ok; attempt++) {
                            const mc_u32x4 b = mc_rng_block(seed, frame, (uint32_t)ray_i, (uint32_t)sample_i, (uint32_t)i, 1 + attempt);
                            ok = random_unit_vector_attempt(h.normal, random_angle, mc_u01d(b.v[0]), mc_u01d(b.v[1]), random_normal);
                        }
                        if (!ok) random_normal = h.normal;
                    }
                    const boundary_result result = hit_boundary(s, ray_, hit_point, random_normal, random_angle, organ, mac, mav, mc_u01f(b0.v[3]), -1);
                    put(inside_point, result.reflected_intensity, intensity_before_hit, distance_before_hit, h.tri, h.mesh, h.fraction);
                    if (result.returned.intensity > INTENSITY_EPSILON) ray_ = result.returned;   // scene.cpp:151-157
                    else ray_.null = true;
                } else {
                    put(to, 0.0f, ray_.intensity, ray_.distance_traveled, -1, -1, 1.0f);         // scene.cpp:163-164
                    ray_.null = true;
                }
            }
        }
    }
    return tests;
}

// ------------------------------------------------------------------------------------------------
// Ray-tree mode (new functionality): the loop of scene::cast_rays with BOTH children of every
// boundary hit followed (as in the cited paper) instead of the one branch ray.cpp:84-94 keeps.  Node ids: root 1, reflected
// child 2n, refracted child 2n+1; the Philox counter takes the node id where the single-path loop puts the bounce index.
// Output: segments of all paths in (path, node) order; returns the number written (or -1 if capacity is too small).
// ------------------------------------------------------------------------------------------------
int64_t orc_cast_rays_tree(const orc_scene* sp, const orc_params* pp, const float* pos3, const float* angles_deg3, uint64_t seed,
                           uint32_t frame, int32_t use_bvh, int64_t capacity, orc_segment* segments, int32_t* seg_path, int32_t* seg_node)
{
    const orc_scene& s = *sp; const orc_params& p = *pp;
    orc_derived dv; orc_derive(&p, &dv);
    const int E = p.elements, S = p.samples, D = p.max_depth;
    std::vector<float> epos((size_t)E * 3), edir((size_t)E * 3);
    transducer_elements(p, dv, pos3, angles_deg3, epos.data(), edir.data());
    struct node_t { path_t ray; int node; };
    int64_t n_out = 0;
    for (int ray_i = 0; ray_i < E; ray_i++) {
        for (int sample_i = 0; sample_i < S; sample_i++) {
            path_t first_ray;
            first_ray.from = mk(epos[3 * ray_i], epos[3 * ray_i + 1], epos[3 * ray_i + 2]);
            first_ray.direction = mk(edir[3 * ray_i], edir[3 * ray_i + 1], edir[3 * ray_i + 2]);
            first_ray.depth = 0;
            first_ray.media = s.starting_material;
            first_ray.media_outside = OUTSIDE_NULL;
            first_ray.intensity = 1.0f / (float)(unsigned)S;
            first_ray.frequency = p.frequency_mhz;
            first_ray.distance_traveled = 0;
            first_ray.null = false;
            // breadth-first: node ids of a level are ascending if the parents are, so the output is sorted by node id
            std::vector<node_t> level{{first_ray, 1}}, next;
            while (!level.empty()) {
                next.clear();
                for (const node_t& nd : level) {
                    path_t ray_ = nd.ray;
                    const material_t& media = s.materials[ray_.media];
                    if (n_out >= capacity) return -1;
                    orc_segment* seg_out = &segments[n_out];
                    seg_path[n_out] = ray_i * S + sample_i; seg_node[n_out] = nd.node;
                    n_out++;
                    const float r_length = max_ray_length(media.attenuation, ray_.intensity, ray_.frequency);
                    const v3 to = add(ray_.from, enlarge(s, ray_.direction, r_length));
                    const v3 from_test = add(ray_.from, scl(ray_.direction, 0.1f));
                    const hit_t h = use_bvh ? closest_hit_bvh(s, from_test, to) : closest_hit_brute(s, from_test, to);
                    auto put = [&](v3 seg_to, float refl, float init_i, double dist_before, int tri, int mesh, float frac) {
                        seg_out->from[0] = ray_.from.x; seg_out->from[1] = ray_.from.y; seg_out->from[2] = ray_.from.z;
                        seg_out->to[0] = seg_to.x; seg_out->to[1] = seg_to.y; seg_out->to[2] = seg_to.z;
                        seg_out->dir[0] = ray_.direction.x; seg_out->dir[1] = ray_.direction.y; seg_out->dir[2] = ray_.direction.z;
                        seg_out->reflected_intensity = refl; seg_out->initial_intensity = init_i;
                        seg_out->attenuation = media.attenuation; seg_out->distance_traveled = dist_before;
                        seg_out->media_id = ray_.media; seg_out->tri_id = tri; seg_out->mesh_id = mesh; seg_out->hit_fraction = frac;
                    };
                    if (h.tri < 0) { put(to, 0.0f, ray_.intensity, ray_.distance_traveled, -1, -1, 1.0f); continue; }
                    const double distance_before_hit = ray_.distance_traveled;
                    const float intensity_before_hit = ray_.intensity;
                    const mesh_t& organ = s.meshes[h.mesh];
                    const v3 hit_point = interpolate3(from_test, to, h.fraction);
                    const mc_u32x4 b0 = mc_rng_block(seed, frame, (uint32_t)ray_i, (uint32_t)sample_i, (uint32_t)nd.node, 0);
                    float q = 0.0f;
                    const float thickness = s.materials[organ.mat_in].thickness;
                    if (!p.deterministic && thickness != 0.0f) {
                        const double u1 = mc_u01d(b0.v[0]), u2 = mc_u01d(b0.v[1]);
                        double sn, cs; mc_sincos(2 * MC_PI_D * u2, &sn, &cs);
                        const double z = sqrt(-2.0 * mc_log(u1)) * cs;
                        q = (float)fabs(z * (double)thickness);
                    }
                    const v3 inside_point = add(scl(ray_.direction, q), hit_point);
                    travel(media.attenuation, ray_.frequency, ray_.intensity, ray_.distance_traveled, distance_in_mm(s, ray_.from, inside_point));
                    int mac, mav; medium_after(ray_, organ, mac, mav);
                    float random_angle = 1.0f; v3 random_normal = h.normal;
                    if (!p.deterministic) {
                        random_angle = power_cosine_variate((int)s.materials[mac].shininess, mc_u01d(b0.v[2]));
                        bool ok = false;
                        for (uint32_t attempt = 0; attempt < MC_RNG_BLOCKS_PER_BOUNCE - 1 && !ok; attempt++) {
                            const mc_u32x4 b = mc_rng_block(seed, frame, (uint32_t)ray_i, (uint32_t)sample_i, (uint32_t)nd.node, 1 + attempt);
                            ok = random_unit_vector_attempt(h.normal, random_angle, mc_u01d(b.v[0]), mc_u01d(b.v[1]), random_normal);
                        }
                        if (!ok) random_normal = h.normal;
                    }
                    // both branches of hit_boundary (force_branch 1 = reflection, 0 = refraction)
                    const boundary_result refl = hit_boundary(s, ray_, hit_point, random_normal, random_angle, organ, mac, mav, 0.0f, 1);
                    const boundary_result refr = hit_boundary(s, ray_, hit_point, random_normal, random_angle, organ, mac, mav, 0.0f, 0);
                    put(inside_point, refl.reflected_intensity, intensity_before_hit, distance_before_hit, h.tri, h.mesh, h.fraction);
                    if (ray_.depth + 1 < D) {
                        if (refl.returned.intensity > INTENSITY_EPSILON) next.push_back({refl.returned, 2 * nd.node});
                        if (refr.returned.intensity > INTENSITY_EPSILON) next.push_back({refr.returned, 2 * nd.node + 1});
                    }
                }
                level.swap(next);
            }
        }
    }
    return n_out;
}

// main.cpp:106-144 on a flat segment list (ray-tree mode): segments are accumulated in the given order into column
// seg_path / samples.
int64_t orc_accumulate_flat(const orc_scene* sp, const orc_params* pp, const orc_volume* vol, const orc_segment* segments,
                            const int32_t* seg_path, int64_t n_segments, float* rf)
{
    const orc_scene& s = *sp; const orc_params& p = *pp;
    orc_derived dv; orc_derive(&p, &dv);
    const int S = p.samples;
    const int rows = dv.rows, cols = dv.cols;
    const float vol_resolution = p.resolution_um / 1000.0f;
    const float axres_f = dv.axial_resolution_f;
    const double time_step = dv.time_step_us;
    const double max_travel_time = dv.max_travel_time_us;
    int64_t total_steps = 0;
    auto add_echo = [&](int column, float echo, double micros) {
        const double row = micros / dv.row_period_us;
        if (row < (double)(unsigned)rows) rf[(size_t)(int)row * cols + column] += echo;
    };
    for (int64_t i = 0; i < n_segments; i++) {
        const orc_segment& seg = segments[i];
        const int column = seg_path[i] / S;
        const material_t& media = s.materials[seg.media_id];
        const double starting_micros = ((seg.distance_traveled * 1000) / 1) / (double)p.speed_of_sound;
        const v3 from = mk(seg.from[0], seg.from[1], seg.from[2]), to = mk(seg.to[0], seg.to[1], seg.to[2]);
        const double distance = (double)(length(sub(to, from)) * 10.0f);
        const double steps_d = distance / dv.axial_resolution_mm;
        uint64_t steps64;
        if (!(steps_d >= 0.0)) steps64 = 0;
        else if (steps_d >= 9.0e18) steps64 = (uint64_t)9000000000000000000ULL;
        else steps64 = (uint64_t)steps_d;
        const uint32_t steps32 = (uint32_t)steps64;
        const v3 delta_step = scl(mk(seg.dir[0], seg.dir[1], seg.dir[2]), axres_f);
        v3 point = from;
        double time_elapsed = starting_micros;
        float intensity = seg.initial_intensity;
        const float decay = mc_expf(-seg.attenuation * axres_f * 0.01f * p.frequency_mhz * 1.0f);
        for (uint64_t step = 0; step < steps64 && time_elapsed < max_travel_time; step++) {
            const float scattering = get_scattering(*vol, vol_resolution, media.mu1, media.mu0, media.sigma, point.x, point.y, point.z);
            add_echo(column, intensity * scattering, time_elapsed);
            point = add(point, delta_step);
            time_elapsed = time_elapsed + time_step;
            intensity *= decay;
            total_steps++;
        }
        add_echo(column, seg.reflected_intensity / (float)(size_t)S, starting_micros + time_step * (double)(uint32_t)(steps32 - 1u));
    }
    return total_steps;
}

// ------------------------------------------------------------------------------------------------
// main.cpp:106-144 + rf_image::add_echo (rfimage.h:33-40)
// ------------------------------------------------------------------------------------------------
int64_t orc_accumulate(const orc_scene* sp, const orc_params* pp, const orc_volume* vol, const orc_segment* segments,
                       const int32_t* n_segments, float* rf)
{
    const orc_scene& s = *sp; const orc_params& p = *pp;
    orc_derived dv; orc_derive(&p, &dv);
    const int E = p.elements, S = p.samples, D = p.max_depth;
    const int rows = dv.rows, cols = dv.cols;
    const float vol_resolution = p.resolution_um / 1000.0f;
    const float axres_f = dv.axial_resolution_f;
    const double time_step = dv.time_step_us;
    const double max_travel_time = dv.max_travel_time_us;
    int64_t total_steps = 0;
    auto add_echo = [&](int column, float echo, double micros) {
        const double row = micros / dv.row_period_us;
        if (row < (double)(unsigned)rows) rf[(size_t)(int)row * cols + column] += echo;
    };
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(g_threads) reduction(+ : total_steps) if (g_threads > 1)
#endif
    for (int ray_i = 0; ray_i < E; ray_i++) {
        for (int sample_i = 0; sample_i < S; sample_i++) {
            const int ns = n_segments[(size_t)ray_i * S + sample_i];
            for (int k = 0; k < ns; k++) {
                const orc_segment& seg = segments[((size_t)ray_i * S + sample_i) * D + k];
                const material_t& media = s.materials[seg.media_id];
                const double starting_micros = ((seg.distance_traveled * 1000) / 1) / (double)p.speed_of_sound;
                const v3 from = mk(seg.from[0], seg.from[1], seg.from[2]), to = mk(seg.to[0], seg.to[1], seg.to[2]);
                const double distance = (double)(length(sub(to, from)) * 10.0f);                     // scene.cpp:342-346
                const double steps_d = distance / dv.axial_resolution_mm;
                uint64_t steps64;                                                                    // B-14
                if (!(steps_d >= 0.0)) steps64 = 0;
                else if (steps_d >= 9.0e18) steps64 = (uint64_t)9000000000000000000ULL;
                else steps64 = (uint64_t)steps_d;
                const uint32_t steps32 = (uint32_t)steps64;
                const v3 delta_step = scl(mk(seg.dir[0], seg.dir[1], seg.dir[2]), axres_f);
                v3 point = from;
                double time_elapsed = starting_micros;
                float intensity = seg.initial_intensity;
                const float decay = mc_expf(-seg.attenuation * axres_f * 0.01f * p.frequency_mhz * 1.0f);
                for (uint64_t step = 0; step < steps64 && time_elapsed < max_travel_time; step++) {
                    const float scattering = get_scattering(*vol, vol_resolution, media.mu1, media.mu0, media.sigma, point.x, point.y, point.z);
                    add_echo(ray_i, intensity * scattering, time_elapsed);
                    point = add(point, delta_step);
                    time_elapsed = time_elapsed + time_step;
                    intensity *= decay;
                    total_steps++;
                }
                add_echo(ray_i, seg.reflected_intensity / (float)(size_t)S, starting_micros + time_step * (double)(uint32_t)(steps32 - 1u));
            }
        }
    }
    return total_steps;
}

// rfimage.h:93-123
void orc_convolve(float* rf, int32_t rows, int32_t cols, const float* axial, int32_t n_axial, const float* lateral, int32_t n_lateral)
{
    std::vector<float> buf((size_t)rows * cols, 0.0f);
    for (int col = 0; col < cols; col++)
        for (int row = n_axial; row < rows - n_axial; row++) {
            float convolution = 0;
            for (int k = 0; k < n_axial; k++) convolution += rf[(size_t)(row + k) * cols + col] * axial[k];
            buf[(size_t)row * cols + col] = convolution;
        }
    for (int row = n_axial; row < rows - n_axial; row++)
        for (int col = n_lateral / 2; col < cols - n_lateral; col++) {
            float convolution = 0;
            for (int k = 0; k < n_lateral; k++) convolution += buf[(size_t)row * cols + col + k] * lateral[k];
            rf[(size_t)row * cols + col] = convolution;
        }
}

// rfimage.h:54-91
void orc_envelope(float* rf, int32_t rows, int32_t cols)
{
    auto I = [&](size_t r, size_t c) -> float& { return rf[r * cols + c]; };
    if (rows < 2) return;
    for (size_t column = 0; column < (size_t)cols; column++) {
        bool ascending = I(0, column) < I(1, column);
        size_t last_peak_pos = 0;
        float last_peak = I(last_peak_pos, column);
        for (size_t i = 1; i + 1 < (size_t)rows; i++) {
            if (I(i, column) < I(i + 1, column)) {
                ascending = true;
            } else if (ascending) {
                ascending = false;
                const float new_peak = std::abs(I(i, column));
                for (size_t j = last_peak_pos; j < i; j++) {
                    const float alpha = (static_cast<float>(j) - static_cast<float>(last_peak_pos)) /
                                        (static_cast<float>(i) - static_cast<float>(last_peak_pos));
                    I(j, column) = last_peak * (1 - alpha) + new_peak * alpha;
                }
                last_peak_pos = i;
                last_peak = new_peak;
            }
        }
    }
}

// rfimage.h:127-136 (commented out in the reference): max = minMaxLoc; I = log10(I + 1) / log10(max + 1)
// Depth-dependent lateral PSF (new functionality on top of psf.h:52-57 / rfimage.h:111-122): taps of row r
// are exp(-0.5 y^2 / (var_y w^2)), w = 1 + spread |depth(r) - focus| / focus; table [n_lateral][rows].
void orc_psf_depth_table(const orc_params* p, int32_t rows, float focus_cm, float spread, float* table)
{
    const float half_lateral = (size_t)p->psf_lateral * (size_t)p->resolution_um / 1000.0f / 2.0f;
    const float resolution = p->resolution_um / 1000.0f;
    for (int32_t r = 0; r < rows; r++) {
        const double depth = (double)r * p->depth_cm / (double)rows;
        const float w = (float)(1.0 + (double)spread * std::fabs(depth - (double)focus_cm) / (double)focus_cm);
        const float var = p->psf_var_y * w * w;
        for (int32_t i = 0; i < p->psf_lateral; i++) {
            const float y = (size_t)i * resolution - half_lateral;
            const double yy = (double)y * (double)y;
            table[(size_t)i * rows + r] = (float)std::exp(-0.5f * (yy / var));
        }
    }
}

// rf_image::convolve (rfimage.h:93-123) with per-row lateral taps lateral_by_row[k * rows + row]
void orc_convolve_depth(float* rf, int32_t rows, int32_t cols, const float* axial, int32_t n_axial, const float* lateral_by_row, int32_t n_lateral)
{
    std::vector<float> buf((size_t)rows * cols, 0.0f);
    for (int col = 0; col < cols; col++)
        for (int row = n_axial; row < rows - n_axial; row++) {
            float convolution = 0;
            for (int k = 0; k < n_axial; k++) convolution += rf[(size_t)(row + k) * cols + col] * axial[k];
            buf[(size_t)row * cols + col] = convolution;
        }
    for (int row = n_axial; row < rows - n_axial; row++)
        for (int col = n_lateral / 2; col < cols - n_lateral; col++) {
            float convolution = 0;
            for (int k = 0; k < n_lateral; k++) convolution += buf[(size_t)row * cols + col + k] * lateral_by_row[(size_t)k * rows + row];
            rf[(size_t)row * cols + col] = convolution;
        }
}

// B-mode display chain (new functionality on top of rfimage.h:127-148): TGC gain, log compression to a
// dynamic range.  rf: [rows][cols] envelope image (oracle layout), in place -> [0, 1].
void orc_bmode(float* rf, int32_t rows, int32_t cols, double depth_cm, float gain_db, float tgc_db_per_cm, float dynamic_range_db)
{
    const double ln10 = 2.30258509299404568402;
    const size_t n = (size_t)rows * cols;
    float mx = 0.0f;
    for (int32_t r = 0; r < rows; r++) {
        const double d = (double)r * depth_cm / (double)rows;
        const float g = (float)mc_exp(((double)gain_db + (double)tgc_db_per_cm * d) * (ln10 / 20.0));
        for (int32_t c = 0; c < cols; c++) {
            const float v = std::fabs(rf[(size_t)r * cols + c]) * g;
            rf[(size_t)r * cols + c] = v;
            mx = v > mx ? v : mx;
        }
    }
    const double scale = 20.0 / (double)dynamic_range_db;
    for (size_t i = 0; i < n; i++) {
        const float v = rf[i];
        float y = 0.0f;
        if ((double)mx > 0.0 && v > 0.0f) {
            y = (float)(1.0 + scale * (mc_log((double)v / (double)mx) / ln10));
            y = y < 0.0f ? 0.0f : (y > 1.0f ? 1.0f : y);
        }
        rf[i] = y;
    }
}

void orc_log_compress(float* rf, int32_t rows, int32_t cols)
{
    const double ln10 = 2.30258509299404568402;
    float mx = -3.0e38f;
    const size_t n = (size_t)rows * cols;
    for (size_t i = 0; i < n; i++) mx = rf[i] > mx ? rf[i] : mx;           // NaN samples never win, like fmaxf
    const double den = mc_log((double)mx + 1) / ln10;
    for (size_t i = 0; i < n; i++) {
        const float num = (float)(mc_log((double)(rf[i] + 1)) / ln10);
        rf[i] = (float)((double)num / den);
    }
}

// rfimage.h:183-215
void orc_create_mapping(const orc_params* p, float* map_x, float* map_y)
{
    orc_derived dv; orc_derive(p, &dv);
    const int srows = p->scan_rows, scols = p->scan_cols;
    const float radius_f = (float)((p->radius_cm * 10) / 1);                 // rf_image ctor takes millimeter_t
    const double total_angle = ((p->fov_deg * (MC_PI_D * 1.0) * 1) / 180);
    const float total_angle_f = (float)total_angle;
    const float depth_f = (dv.max_travel_time_u * p->speed_of_sound) * 0.001f;   // unsigned*unsigned -> float
    const float ratio = (float)(((depth_f + radius_f) - radius_f * std::cos(total_angle_f / 2.0)) / srows);
    const double shift_y = ((double)radius_f) * (double)std::cos(total_angle_f / 2.0f);
    const float half_width = (float)scols / 2.0f;
    for (int j = 0; j < scols; j++)
        for (int i = 0; i < srows; i++) {
            const float fi = static_cast<float>(i) + (float)shift_y / ratio;
            const float fj = static_cast<float>(j) - half_width;
            const float r = std::sqrt(std::pow(fi, 2.0f) + std::pow(fj, 2.0f));
            const double angle = (double)std::atan2(fj, fi);
            map_x[(size_t)i * scols + j] = (r * ratio - radius_f) / depth_f * (float)dv.rows;
            map_y[(size_t)i * scols + j] = (float)(((angle - (-(total_angle / 2))) / total_angle) * (float)dv.cols);
        }
}

// cv::remap(src, dst, map1 = map_y (x / column), map2 = map_x (y / row), INTER_LINEAR,
// BORDER_CONSTANT, 0) as OpenCV computes it for CV_32FC1 maps: coordinates are rounded to 1/32
// pixel (INTER_BITS = 5) and the four taps are weighted with a float table.
void orc_scan_convert(const float* rf, int32_t rows, int32_t cols, const float* map_x, const float* map_y, int32_t scan_rows,
                      int32_t scan_cols, float* out)
{
    for (int i = 0; i < scan_rows; i++)
        for (int j = 0; j < scan_cols; j++) {
            const float x = map_y[(size_t)i * scan_cols + j], y = map_x[(size_t)i * scan_cols + j];
            // saturate_cast<int>(x * INTER_TAB_SIZE) = cvRound with saturation; NaN -> INT_MIN
            auto fix = [](float v) -> int {
                const double t = (double)v * 32.0;
                if (!(t == t)) return INT32_MIN;
                if (t <= -2147483648.0) return INT32_MIN;
                if (t >= 2147483647.0) return INT32_MAX;
                return (int)lrint(t);
            };
            const int sx = fix(x), sy = fix(y);
            const int ix = sx >> 5, iy = sy >> 5;
            const int fx = sx & 31, fy = sy & 31;
            const float ax = fx * (1.0f / 32), ay = fy * (1.0f / 32);
            const float w[4] = {(1.0f - ay) * (1.0f - ax), (1.0f - ay) * ax, ay * (1.0f - ax), ay * ax};
            auto px = [&](int r, int c) -> float { return (r >= 0 && r < rows && c >= 0 && c < cols) ? rf[(size_t)r * cols + c] : 0.0f; };
            float v;
            if ((unsigned)ix < (unsigned)(cols - 1) && (unsigned)iy < (unsigned)(rows - 1)) {
                const float* S = rf + (size_t)iy * cols + ix;
                v = S[0] * w[0] + S[1] * w[1] + S[cols] * w[2] + S[cols + 1] * w[3];
            } else if (ix >= cols || ix + 1 < 0 || iy >= rows || iy + 1 < 0) {
                v = 0.0f;
            } else {
                v = px(iy, ix) * w[0] + px(iy, ix + 1) * w[1] + px(iy + 1, ix) * w[2] + px(iy + 1, ix + 1) * w[3];
            }
            out[(size_t)i * scan_cols + j] = v;
        }
}

// host evaluation of the shared numerics contract (same op codes as mcrt_numerics_probe)
void orc_numerics(int32_t op, int64_t n, const double* a, const double* b, double* out)
{
    for (int64_t i = 0; i < n; i++) {
        double r = 0.0;
        switch (op) {
            case 0: r = (double)mc_expf((float)a[i]); break;
            case 1: r = (double)mc_logf((float)a[i]); break;
            case 2: r = (double)mc_powf((float)a[i], (float)b[i]); break;
            case 3: { double sn, cs; mc_sincos(a[i], &sn, &cs); r = sn; break; }
            case 4: { double sn, cs; mc_sincos(a[i], &sn, &cs); r = cs; break; }
            case 5: {
                const mc_u32x4 w = mc_rng_block(0x0123456789abcdefULL, (uint32_t)a[i], (uint32_t)b[i], 3u, 2u, 1u);
                r = (double)w.v[0] + 4294967296.0 * (double)(w.v[3] & 0xfffffu);
                break;
            }
            case 6: r = mc_pow(a[i], b[i]); break;
            case 7: r = mc_exp(a[i]); break;
            case 8: r = mc_log(a[i]); break;
            case 9: {
                // Markstein's division step in fp32 with a true IEEE fma, as the device evaluates the envelope's
                // alpha = (j - p) / (q - p) (k_post_tma, div_small_int): q0 = a y, r = fma(-q0, b, a), q = fma(r, y, q0), y = fl(1 / b)
                const float fa = (float)a[i], fb = (float)b[i];
                const float y = 1.0f / fb;
                const float q0 = fa * y;
                const float rr = __builtin_fmaf(-q0, fb, fa);
                r = (double)__builtin_fmaf(rr, y, q0);
                break;
            }
            default: break;
        }
        out[i] = r;
    }
}

int64_t orc_simulate_frame(const orc_scene* s, const orc_params* p, const float* pos3, const float* angles_deg3, uint64_t seed,
                           uint32_t frame, float* rf, float* scan_out, double* stage_seconds4, int64_t* steps_out)
{
    using clk = std::chrono::steady_clock;
    orc_derived dv; orc_derive(p, &dv);
    const size_t nseg = (size_t)p->elements * p->samples * p->max_depth;
    std::vector<orc_segment> segs(nseg);
    std::vector<int32_t> nsegs((size_t)p->elements * p->samples);
    const auto t0 = clk::now();
    memset(rf, 0, sizeof(float) * (size_t)dv.rows * dv.cols);                           // rf_image.clear(), main.cpp:102
    const int64_t tests = orc_cast_rays(s, p, pos3, angles_deg3, seed, frame, 1, segs.data(), nsegs.data());
    const auto t1 = clk::now();
    const int64_t steps = orc_accumulate(s, p, orc_volume_get(), segs.data(), nsegs.data(), rf);
    const auto t2 = clk::now();
    std::vector<float> ax(p->psf_axial), lat(p->psf_lateral);
    orc_psf_taps(p, ax.data(), lat.data());
    orc_convolve(rf, dv.rows, dv.cols, ax.data(), p->psf_axial, lat.data(), p->psf_lateral);
    orc_envelope(rf, dv.rows, dv.cols);
    const auto t3 = clk::now();
    if (scan_out) {
        std::vector<float> mx((size_t)p->scan_rows * p->scan_cols), my((size_t)p->scan_rows * p->scan_cols);
        orc_create_mapping(p, mx.data(), my.data());
        orc_scan_convert(rf, dv.rows, dv.cols, mx.data(), my.data(), p->scan_rows, p->scan_cols, scan_out);
    }
    const auto t4 = clk::now();
    if (stage_seconds4) {
        stage_seconds4[0] = std::chrono::duration<double>(t1 - t0).count();
        stage_seconds4[1] = std::chrono::duration<double>(t2 - t1).count();
        stage_seconds4[2] = std::chrono::duration<double>(t3 - t2).count();
        stage_seconds4[3] = std::chrono::duration<double>(t4 - t3).count();
    }
    if (steps_out) *steps_out = steps;
    return tests;
}

}  // extern "C"
