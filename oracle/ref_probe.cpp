// ref_probe.cpp -- TEST INFRASTRUCTURE ONLY.  Builds (oracle/Makefile, target _ref) into
// oracle/_ref/libref_probe.so by compiling the reference's own headers / translation units from
// where they lie under /root/reference (never copied into this repo) against the tiny shims in
// oracle/ref_shims/.  Exposes the reference's real psf / volume / transducer / ray_physics /
// rf_image / tinyobj+objloader code through a C ABI so tests/golden/make_golden.py can record
// known-answer vectors and tests can pin the oracle restatement to the reference.
//
// What cannot be probed: scene.cpp's cast_rays (needs Bullet's collision world).  main.cpp's constants block (lines 17-37), its
// echo-accumulation loop (lines 106-144) and scene::distance (scene.cpp:341-346) are extracted verbatim at build time into
// oracle/_ref/*.inc by the Makefile and compiled here (ref_accumulate_loop).
#define private public      // reach rf_image::intensities / map_x / map_y (test-only)
#define protected public
#include "rfimage.h"        // pulls psf.h, units.h and the cv::Mat shim
#undef private
#undef protected
#include "volume.h"
#include "transducer.h"
#include "ray.h"
#include "mesh.h"
#include "objloader.h"

#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <memory>

#include "_ref/main_constants.inc"

using psf_ref = psf_;
using rf_ref = rf_image_;

// ---- a stand-in for the Bullet-bound parts of class scene (scene.h:19-76) -------------------------------------------------
// The reference's scene::cast_rays<S,E> (scene.cpp:50-183), distance_in_mm / enlarge (:281-298) and distance (:341-346) are
// compiled VERBATIM from the extracted text below; only the collision world behind m_dynamicsWorld->rayTest is replaced: it
// forwards the query to a caller-supplied closest-hit function (the oracle's), so what gets pinned is everything the reference
// does AROUND the ray test (ray set-up, max_ray_length / enlarge, the 1 mm start offset, travel, hit_boundary, segment
// emission, path termination, the bounce / sample / element loop order), not Bullet's triangle arithmetic.
#include <ctime>
#include <iostream>
#include <sstream>
#include <cassert>
#include <random>
struct btCollisionObject { void* m_user = nullptr; void* getUserPointer() const { return m_user; } };
struct btCollisionWorld {
    struct ClosestRayResultCallback {
        ClosestRayResultCallback(const btVector3& f, const btVector3& t) : m_rayFromWorld(f), m_rayToWorld(t) {}
        bool hasHit() const { return m_collisionObject != nullptr; }
        btVector3 m_rayFromWorld, m_rayToWorld, m_hitPointWorld, m_hitNormalWorld;
        const btCollisionObject* m_collisionObject = nullptr;
    };
};
typedef int (*ref_hit_fn)(const void* ctx, const float* from3, const float* to3, int use_bvh, float* out_f7, int* out_mesh);   // = orc_closest_hit
struct shim_world {
    ref_hit_fn fn = nullptr;
    const void* ctx = nullptr;
    std::vector<btCollisionObject> bodies;
    void rayTest(const btVector3& from, const btVector3& to, btCollisionWorld::ClosestRayResultCallback& cb) const
    {
        const float f[3] = {from.x(), from.y(), from.z()}, t[3] = {to.x(), to.y(), to.z()};
        float o[7];
        int m = -1;
        if (fn(ctx, f, t, 1, o, &m) >= 0) {
            cb.m_hitPointWorld = btVector3(o[1], o[2], o[3]);
            cb.m_hitNormalWorld = btVector3(o[4], o[5], o[6]);
            cb.m_collisionObject = &bodies[(size_t)m];
        }
    }
};
class scene
{
    using transducer_ = transducer<512>;
public:
    template<unsigned int sample_count,unsigned int ray_count>
    std::array<std::array<std::vector<ray_physics::segment>,sample_count>, ray_count>cast_rays(transducer_ & transducer);
    units::length::millimeter_t distance(const btVector3 & from, const btVector3 & to) const;
    units::length::millimeter_t distance_in_mm(const btVector3 & v1, const btVector3 & v2) const;
    btVector3 enlarge(const btVector3 & versor, float mm) const;
    std::unordered_map<std::string, material> materials;
    std::string starting_material;
    std::vector<mesh> meshes;
    const float initial_intensity { 1.0f };
    std::array<float,3> spacing;
    std::unique_ptr<shim_world> m_dynamicsWorld;
    clock_t frame_start;
};
#include "_ref/scene_distance.inc"
#include "_ref/scene_helpers.inc"
#include "_ref/scene_cast_rays.inc"

extern "C" {

// ---- constants (main.cpp:23-37, rfimage.h:43-51,180) -----------------------------------------
void ref_constants(double* out)
{
    rf_ref rf{transducer_radius, transducer_amplitude};
    millimeter_t sep = transducer_amplitude.to<float>() * transducer_radius / transducer_elements;   // main.cpp:66
    out[0] = axial_resolution.to<double>();
    out[1] = max_travel_time.to<double>();
    out[2] = (double)max_travel_time.to<unsigned int>();
    out[3] = (double)static_cast<unsigned int>(axial_resolution.to<float>() * 1000.0f);
    out[4] = (double)rf.intensities.rows;
    out[5] = (double)rf.intensities.cols;
    out[6] = sep.to<double>();
    out[7] = rf.micros_traveled(axial_resolution).to<double>();          // march time step, main.cpp:118
    out[8] = 0.0;   // rf_image::get_dt() (rfimage.h:43-46) does not compile when instantiated; unused by the reference
    out[9] = rf.micros_traveled(millimeter_t(37.25)).to<double>();       // main.cpp:114 with 37.25 mm
    out[10] = (double)(unsigned int)(millimeter_t(123.456) / axial_resolution);   // main.cpp:116
    out[11] = (double)transducer_radius.to<float>();
    out[12] = transducer_amplitude.to<double>();
    out[13] = (double)(axial_resolution.to<float>());
}

// ---- psf.h ------------------------------------------------------------------------------------
void ref_psf_taps(float* axial7, float* lateral13)
{
    const psf_ref p{transducer_frequency, 0.05f, 0.2f, 0.1f};      // main.cpp:54
    for (int i = 0; i < 7; i++) axial7[i] = p.axial_kernel[i];
    for (int i = 0; i < 13; i++) lateral13[i] = p.lateral_kernel[i];
}

// ---- volume.h ---------------------------------------------------------------------------------
static const volume_* ref_vol()
{
    static const volume_* v = new volume_();
    return v;
}
const float* ref_volume_raw() { return reinterpret_cast<const float*>(ref_vol()); }
float ref_volume_get_scattering(float density, float mu, float sigma, float x, float y, float z)
{
    return ref_vol()->get_scattering(density, mu, sigma, x, y, z);
}

// ---- transducer.h -----------------------------------------------------------------------------
void ref_transducer_elements(const float* pos, const float* angles_deg, float* out_pos, float* out_dir)
{
    using namespace units::angle;
    millimeter_t sep = transducer_amplitude.to<float>() * transducer_radius / transducer_elements;
    std::array<degree_t, 3> ang = {degree_t((float)angles_deg[0]), degree_t((float)angles_deg[1]), degree_t((float)angles_deg[2])};
    static std::unique_ptr<transducer_> t;
    t.reset(new transducer_(transducer_frequency, transducer_radius, sep, btVector3(pos[0], pos[1], pos[2]), ang));
    for (size_t i = 0; i < transducer_elements; i++) {
        auto e = t->element(i);
        for (int k = 0; k < 3; k++) { out_pos[3 * i + k] = e.position[k]; out_dir[3 * i + k] = e.direction[k]; }
    }
}

// ---- ray.cpp ----------------------------------------------------------------------------------
static material mk_mat(const float* m) { return material{m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7]}; }

float ref_max_ray_length(const float* mat8, float intensity, float frequency)
{
    ray_physics::ray r{btVector3(0, 0, 0), btVector3(1, 0, 0), 0, mk_mat(mat8), nullptr, intensity, frequency, millimeter_t(0), 0};
    return ray_physics::max_ray_length(r);
}
void ref_travel(const float* mat8, float intensity, float frequency, double dist0, double mm, float* out_intensity, double* out_dist)
{
    ray_physics::ray r{btVector3(0, 0, 0), btVector3(1, 0, 0), 0, mk_mat(mat8), nullptr, intensity, frequency, millimeter_t(dist0), 0};
    ray_physics::travel(r, millimeter_t(mm));
    *out_intensity = r.intensity; *out_dist = r.distance_traveled.to<double>();
}
float ref_reflection_intensity(float i_in, float z1, float c1, float z2, float c2)
{
    return ray_physics::reflection_intensity(i_in, z1, c1, z2, c2);
}
float ref_reflected_intensity_eq8(const float* d, const float* refr, const float* refl, const float* mat8)
{
    return ray_physics::reflected_intensity(btVector3(d[0], d[1], d[2]), btVector3(refr[0], refr[1], refr[2]),
                                            btVector3(refl[0], refl[1], refl[2]), mk_mat(mat8));
}
void ref_snells_law(const float* l, const float* n, float c1, float c2, float ratio, float* out)
{
    btVector3 v = ray_physics::snells_law(btVector3(l[0], l[1], l[2]), btVector3(n[0], n[1], n[2]), c1, c2, ratio);
    out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
}
void ref_random_unit_vector(const float* v, float cos_theta, float* out)
{
    btVector3 w = ray_physics::random_unit_vector(btVector3(v[0], v[1], v[2]), cos_theta);
    out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
}
float ref_power_cosine_variate(int v) { return ray_physics::power_cosine_variate(v); }

// A miniature of scene.cpp:80-157's per-path slot: the reference keeps each path in a std::array
// slot and assigns result.returned back into it, which is what gives `media_outside = &r.media`
// (ray.cpp:38) its "points at my own slot" meaning.  Materials live in an unordered_map and meshes
// hold references into it exactly as scene.h:41,44 / scene.cpp:209-241 do.
struct ref_world {
    std::unordered_map<int, material> materials;
    std::vector<mesh> meshes;
    ray_physics::ray slot;
};
ref_world* ref_world_create(int n_mat, const float* mats8, int n_mesh, const int* mesh_in, const int* mesh_out, const int* vascular)
{
    ref_world* w = new ref_world();
    for (int i = 0; i < n_mat; i++) w->materials[i] = mk_mat(mats8 + 8 * i);
    w->meshes.reserve(n_mesh);
    for (int i = 0; i < n_mesh; i++)
        w->meshes.emplace_back(mesh{"m", true, vascular[i] != 0, {0.f, 0.f, 0.f}, true, w->materials.at(mesh_in[i]), w->materials.at(mesh_out[i])});
    return w;
}
void ref_world_destroy(ref_world* w) { delete w; }
void ref_world_start(ref_world* w, int start_mat, const float* from, const float* dir, float intensity, float frequency)
{
    // scene.cpp:84-100
    w->slot = ray_physics::ray{btVector3(from[0], from[1], from[2]), btVector3(dir[0], dir[1], dir[2]), 0, w->materials.at(start_mat),
                               nullptr, intensity, frequency, millimeter_t(0), 0};
}
// media_outside code: -1 null, -2 SELF (points into the slot), >=0 material id, -3 unknown pointer
static int classify_outside(const ref_world* w, const ray_physics::ray& r, const ray_physics::ray* slot)
{
    if (!r.media_outside) return -1;
    if (r.media_outside == &slot->media) return -2;
    for (auto& kv : w->materials) if (&kv.second == r.media_outside) return kv.first;
    return -3;
}
static int classify_media(const ref_world* w, const material& m)
{
    for (auto& kv : w->materials) if (memcmp(&kv.second, &m, sizeof(material)) == 0) return kv.first;
    return -3;
}
// One ray_physics::hit_boundary call on the slot (ray.cpp:11-97).  If `commit`, the returned ray
// replaces the slot as scene.cpp:154 does.  out_f: [0]=reflected_intensity, [1..3]=returned.from,
// [4..6]=returned.direction, [7]=returned.intensity; out_i: [0]=depth, [1]=media id (by value
// match; ids with identical parameters alias to the lowest), [2]=media_outside code.
void ref_world_hit(ref_world* w, const float* hit_point, const float* normal, int mesh_id, int commit, float* out_f, int* out_i)
{
    auto res = ray_physics::hit_boundary(w->slot, btVector3(hit_point[0], hit_point[1], hit_point[2]),
                                         btVector3(normal[0], normal[1], normal[2]), w->meshes.at(mesh_id));
    out_f[0] = res.reflected_intensity;
    for (int k = 0; k < 3; k++) { out_f[1 + k] = res.returned.from[k]; out_f[4 + k] = res.returned.direction[k]; }
    out_f[7] = res.returned.intensity;
    out_i[0] = (int)res.returned.depth;
    out_i[2] = classify_outside(w, res.returned, &w->slot);
    if (commit) { w->slot = res.returned; out_i[1] = classify_media(w, w->slot.media); }
    else out_i[1] = classify_media(w, res.returned.media);
}
void ref_world_set_intensity(ref_world* w, float intensity) { w->slot.intensity = intensity; }

// ---- rfimage.h --------------------------------------------------------------------------------
static rf_ref* g_rf = nullptr;
static rf_ref* rf() { if (!g_rf) g_rf = new rf_ref{transducer_radius, transducer_amplitude}; return g_rf; }
void ref_rf_clear() { rf()->clear(); }
void ref_rf_set(const float* img) { memcpy(rf()->intensities.data.data(), img, sizeof(float) * rf()->intensities.data.size()); }
void ref_rf_get(float* img) { memcpy(img, rf()->intensities.data.data(), sizeof(float) * rf()->intensities.data.size()); }
void ref_rf_add_echo(unsigned int column, float echo, double micros) { rf()->add_echo(column, echo, microsecond_t(micros)); }
void ref_rf_convolve() { const psf_ref p{transducer_frequency, 0.05f, 0.2f, 0.1f}; rf()->convolve(p); }
void ref_rf_envelope() { rf()->envelope(); }
void ref_rf_mapping(float* map_x, float* map_y)
{
    memcpy(map_x, rf()->map_x.data.data(), sizeof(float) * 400 * 500);
    memcpy(map_y, rf()->map_y.data.data(), sizeof(float) * 400 * 500);
}

// ---- main.cpp:106-144, the echo-accumulation loop, verbatim -----------------------------------
// segs: [E][S][D][12] = from3, to3, dir3, reflected_intensity, initial_intensity, attenuation; dist_mm: [E][S][D];
// media3: [E][S][D][3] = mu0, mu1, sigma of the segment's medium (copied at emission: SURVEY B-1); nseg: [E][S].
// Runs the reference's own loop over its own rf_image / volume and returns `intensities` ([465][512]).
void ref_accumulate_loop(const float* segs, const double* dist_mm, const float* media3, const int* nseg, int max_depth, float* rf_out)
{
    using rays_t = std::array<std::array<std::vector<ray_physics::segment>, samples_te>, transducer_elements>;
    std::unique_ptr<rays_t> rays_ptr(new rays_t());
    std::vector<material> mats((size_t)transducer_elements * samples_te * max_depth);
    for (size_t e = 0; e < transducer_elements; e++)
        for (size_t s = 0; s < samples_te; s++)
            for (int k = 0; k < nseg[e * samples_te + s]; k++) {
                const size_t i = (e * samples_te + s) * max_depth + k;
                const float* f = segs + 12 * i;
                material& m = mats[i];
                m = material{};
                m.mu0 = media3[3 * i]; m.mu1 = media3[3 * i + 1]; m.sigma = media3[3 * i + 2]; m.attenuation = f[11];
                (*rays_ptr)[e][s].emplace_back(ray_physics::segment{btVector3(f[0], f[1], f[2]), btVector3(f[3], f[4], f[5]), btVector3(f[6], f[7], f[8]),
                                                                     f[9], f[10], f[11], units::length::millimeter_t(dist_mm[i]), m});
            }
    auto& rays = *rays_ptr;
    auto& rf_image = *rf();
    rf_image.clear();                                   // main.cpp:102
    scene scene;
    const auto& texture_volume = *ref_vol();
#include "_ref/main_accumulate_loop.inc"
    memcpy(rf_out, rf_image.intensities.data.data(), sizeof(float) * rf_image.intensities.data.size());
}

// ---- scene.cpp:50-183, scene::cast_rays<5,512>, verbatim ------------------------------------------
// materials8: [n_mat][8]; mesh tables as in the scene JSON; hit_fn / hit_ctx: closest-hit oracle.  Outputs per segment
// [512][5][10]: seg12 = from3, to3, dir3, reflected_intensity, initial_intensity, attenuation; dist_mm; nseg [512][5].
// segment::media is NOT read: it dangles once cast_rays returns (SURVEY B-1).
int64_t ref_cast_rays(const float* materials8, int n_mat, const int* mesh_in, const int* mesh_out, const int* mesh_vasc, int n_mesh,
                      int starting_material, const float* spacing3, const float* pos3, const float* angles_deg3, void* hit_fn,
                      const void* hit_ctx, float* seg12, double* dist_mm, int* nseg)
{
    using namespace units::angle;
    scene sc;
    for (int i = 0; i < n_mat; i++) sc.materials[std::to_string(i)] = mk_mat(materials8 + 8 * i);
    sc.starting_material = std::to_string(starting_material);
    sc.meshes.reserve(n_mesh);
    for (int i = 0; i < n_mesh; i++)
        sc.meshes.emplace_back(mesh{"m", true, mesh_vasc[i] != 0, {0.f, 0.f, 0.f}, true, sc.materials.at(std::to_string(mesh_in[i])),
                                    sc.materials.at(std::to_string(mesh_out[i]))});
    sc.spacing = {spacing3[0], spacing3[1], spacing3[2]};
    sc.m_dynamicsWorld.reset(new shim_world());
    sc.m_dynamicsWorld->fn = (ref_hit_fn)hit_fn;
    sc.m_dynamicsWorld->ctx = hit_ctx;
    sc.m_dynamicsWorld->bodies.resize(n_mesh);
    for (int i = 0; i < n_mesh; i++) sc.m_dynamicsWorld->bodies[i].m_user = &sc.meshes[i];          // scene.cpp:46 setUserPointer(&mesh)
    sc.frame_start = clock();
    millimeter_t sep = transducer_amplitude.to<float>() * transducer_radius / transducer_elements;   // main.cpp:66
    std::array<degree_t, 3> ang = {degree_t((float)angles_deg3[0]), degree_t((float)angles_deg3[1]), degree_t((float)angles_deg3[2])};
    transducer_ tr(transducer_frequency, transducer_radius, sep, btVector3(pos3[0], pos3[1], pos3[2]), ang);   // main.cpp:65-72
    std::ostringstream sink;                                  // scene.cpp:141-142,179 print two lines per hit and one per frame
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    auto rays = sc.cast_rays<samples_te, transducer_elements>(tr);
    std::cout.rdbuf(old);
    int64_t total = 0;
    const int D = (int)ray_physics::ray::max_depth;
    for (size_t e = 0; e < transducer_elements; e++)
        for (size_t s = 0; s < samples_te; s++) {
            const auto& v = rays[e][s];
            nseg[e * samples_te + s] = (int)v.size();
            total += (int64_t)v.size();
            for (size_t k = 0; k < v.size() && (int)k < D; k++) {
                const size_t i = (e * samples_te + s) * D + k;
                float* f = seg12 + 12 * i;
                for (int a = 0; a < 3; a++) { f[a] = v[k].from[a]; f[3 + a] = v[k].to[a]; f[6 + a] = v[k].direction[a]; }
                f[9] = v[k].reflected_intensity; f[10] = v[k].initial_intensity; f[11] = v[k].attenuation;
                dist_mm[i] = v[k].distance_traveled.to<double>();
            }
        }
    return total;
}

// ---- tinyobj + objloader.h --------------------------------------------------------------------
// returns number of triangles; if out != null writes 9 floats per triangle (objloader.h:23-139 order)
int ref_load_obj(const char* path, float* out, int max_tris)
{
    GLInstanceGraphicsShape* g = load_mesh_from_obj(path, "");
    int ntri = g->m_numIndices / 3;
    if (out) {
        for (int t = 0; t < ntri && t < max_tris; t++)
            for (int k = 0; k < 3; k++) {
                const GLInstanceVertex& v = g->m_vertices->at(g->m_indices->at(3 * t + k));
                out[9 * t + 3 * k + 0] = v.xyzw[0]; out[9 * t + 3 * k + 1] = v.xyzw[1]; out[9 * t + 3 * k + 2] = v.xyzw[2];
            }
    }
    delete g;
    return ntri;
}
}  // extern "C"
