// mcrt_host.h -- host-side (C++) half of the drop-in surface: scene JSON + OBJ loading, transducer
// geometry tables, PSF taps, scan-conversion maps, scatterer volume.  Everything here runs once at
// context creation (the reference does it in main.cpp:52-81) or once per pose (3 sin/cos pairs).
#ifndef MCRT_HOST_H
#define MCRT_HOST_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/mcrt.h"

namespace mcrt {

struct HostMaterial { float impedance, attenuation, mu0, mu1, sigma, specularity, shininess, thickness; };   // mesh.h:7-10

struct HostMesh {                    // mesh.h:12-20 (+ the rigid body origin of scene.cpp:322-324)
    std::string filename;
    bool is_rigid = true, is_vascular = false, outside_normals = true;
    float deltas[3] = {0, 0, 0};
    int material_inside = 0, material_outside = 0;
    float origin[3] = {0, 0, 0};
    int64_t tri_begin = 0, tri_end = 0;
};

struct HostScene {
    std::string working_dir;
    std::vector<std::string> material_names;
    std::vector<HostMaterial> materials;
    int starting_material = 0;
    std::vector<HostMesh> meshes;
    std::vector<float> tri_local;    // 9 floats / triangle: v_obj * scaling (scene.cpp:313-316)
    std::vector<int32_t> tri_mesh;
    float spacing[3] = {1, 1, 1};
    float origin[3] = {0, 0, 0};
    float scaling = 1.0f;
    float start_pose[6] = {0, 0, 0, 0, 0, 0};
};

// Quantities derived from mcrt_params the way main.cpp:25,31,36,66 / rfimage.h:35,48-51,180 derive them.
struct Derived {
    double axial_resolution_mm;
    float axial_resolution_f;
    double max_travel_time_us;
    uint32_t max_travel_time_u;
    uint32_t rf_axial_um;
    int rows, cols;
    double element_separation_mm;
    double time_step_us;
    double row_period_us;
};

// Per-pose trigonometry computed on the host (keep transcendentals off the device).
struct PoseTrig { float pos[3]; float cz, sz, cx, sx, cy, sy; float pad[3]; };   // 48 bytes

void validate_params(const mcrt_params& p);                                    // throws std::invalid_argument
Derived derive(const mcrt_params& p);
// tinyobj 0.9.5 semantics + objloader.h un-welding: appends 9 floats per triangle (OBJ space)
void load_obj_soup(const std::string& path, std::vector<float>& out9);
HostScene load_scene_file(const std::string& scene_path);                      // scene.cpp:185-247 + :300-334
HostScene scene_from_arrays(const mcrt_scene_arrays& a);
// (sin a_t, cos a_t) for every element, transducer.h:41-59
void element_angle_table(const mcrt_params& p, const Derived& d, std::vector<float>& sincos2);
PoseTrig pose_trig(const mcrt_pose& pose);                                     // transducer.h:37-39,51-53
void psf_taps(const mcrt_params& p, std::vector<float>& axial, std::vector<float>& lateral);    // psf.h:34-58
// depth-dependent lateral PSF taps [psf_lateral][rows] (extension; see mcrt_host.cpp)
// Elevational PSF (psf.h:42,77 declares elevation_kernel and never fills it; psf.h:16-18 "three ranges: axial, lateral and
// elevation"): filled by analogy with the lateral kernel, taps[i] = exp(-0.5 z_i^2 / var_z), z_i = i * resolution - n * resolution / 2
// (not normalised, not centred -- psf.h:40-57), plus the elevational offset z_i [mm] of the ray-fan plane the tap weighs.
void psf_elevation_taps(const mcrt_params& p, int n, float var_z, std::vector<float>& taps, std::vector<float>& z_mm);
// the probe pose whose fan lies z_mm off the imaging plane along the transducer's elevation axis (its local z, rotated by the pose)
mcrt_pose elevation_pose(const mcrt_pose& pose, float z_mm);
void psf_lateral_depth_table(const mcrt_params& p, int rows, float focus_cm, float spread, std::vector<float>& table);
void scan_mapping(const mcrt_params& p, const Derived& d, std::vector<float>& map_x, std::vector<float>& map_y);   // rfimage.h:183-215
const std::vector<float>& scatterer_volume();                                  // volume.h:19-35, process-wide

}  // namespace mcrt
#endif
