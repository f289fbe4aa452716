// mcrt_host.cpp -- see mcrt_host.h.  Citations are to thepochynsons/MCRay-Tracing.
#include "mcrt_host.h"

#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <random>
#include <sstream>
#include <stdexcept>

#include "json_min.h"

namespace mcrt {

static const double kPi = 3.14159265358979323846;      // units.h:360 PI_VAL (== glibc M_PI as a double)
static const double kPiRedefined = 3.14159;             // "#define M_PI 3.14159" psf.h:9

void validate_params(const mcrt_params& p)
{
    auto bad = [](const char* m) { throw std::invalid_argument(std::string("mcrt_params: ") + m); };
    if (p.elements < 1 || p.elements > (1 << 20)) bad("elements out of range");
    if (p.samples < 1 || p.samples > 4096) bad("samples out of range");
    if (p.max_depth < 1 || p.max_depth > 16) bad("max_depth must be in [1,16]");
    if (!(p.frequency_mhz > 0.0f)) bad("frequency_mhz must be > 0");
    if (!(p.radius_cm > 0.0) || !(p.fov_deg > 0.0) || !(p.depth_cm > 0.0)) bad("radius/fov/depth must be > 0");
    if (p.speed_of_sound == 0 || p.resolution_um == 0) bad("speed_of_sound/resolution_um must be > 0");
    if (p.psf_axial < 1 || p.psf_axial > 255 || (p.psf_axial % 2) == 0) bad("psf_axial must be odd, 1..255");     // psf.h:29
    if (p.psf_lateral < 1 || p.psf_lateral > 255 || (p.psf_lateral % 2) == 0) bad("psf_lateral must be odd, 1..255");
    if (!(p.psf_var_x > 0.0f) || !(p.psf_var_y > 0.0f)) bad("psf variances must be > 0");
    if (p.scan_rows < 1 || p.scan_cols < 1) bad("scan size must be positive");
    if (!(p.axial_scale > 0.0f)) bad("axial_scale must be > 0");
    if (p.rf_layout != 0 && p.rf_layout != 1) bad("rf_layout must be 0 or 1");
}

Derived derive(const mcrt_params& p)
{
    Derived d;
    // main.cpp:25  axial_resolution = millimeter_t(1.45f / transducer_frequency)
    d.axial_resolution_f = (1.45f / p.frequency_mhz) / p.axial_scale;
    d.axial_resolution_mm = (double)d.axial_resolution_f;
    // main.cpp:31  microsecond_t(ultrasound_depth / speed_of_sound): cm/(m/s) -> us is ratio 10000 (units.h:1365)
    d.max_travel_time_us = ((p.depth_cm / (double)p.speed_of_sound) * 10000) / 1;
    d.max_travel_time_u = (uint32_t)d.max_travel_time_us;
    // main.cpp:36  static_cast<unsigned>(axial_resolution.to<float>() * 1000.0f)
    d.rf_axial_um = (uint32_t)(d.axial_resolution_f * 1000.0f);
    if (d.rf_axial_um == 0) throw std::invalid_argument("mcrt_params: axial resolution below 1 um");
    d.rows = (int)((p.speed_of_sound * d.max_travel_time_u) / d.rf_axial_um);                 // rfimage.h:180
    d.cols = p.elements;
    // main.cpp:66  transducer_amplitude.to<float>() * transducer_radius / transducer_elements  [mm]
    const double amplitude_rad = ((p.fov_deg * (kPi * 1.0) * 1) / 180);                        // units.h:1375
    const float amplitude_f = (float)amplitude_rad;
    const double sep_cm = ((double)amplitude_f * p.radius_cm) / (double)(size_t)p.elements;
    d.element_separation_mm = (sep_cm * 10) / 1;
    d.time_step_us = ((d.axial_resolution_mm * 1000) / 1) / (double)p.speed_of_sound;         // rfimage.h:48-51, main.cpp:118
    d.row_period_us = (double)d.rf_axial_um / (double)p.speed_of_sound;                       // rfimage.h:35
    if (d.rows < 2) throw std::invalid_argument("mcrt_params: fewer than 2 RF rows");
    return d;
}

// ------------------------------------------------------------------------------------------------
// Wavefront OBJ, tinyobj 0.9.5 semantics (tiny_obj_loader.cpp:504-717) reduced to what reaches the
// collision mesh: `v` and `f`; polygon -> triangle fan (:272-285); 1-based, 0 and negative
// (relative) indices (:97-109); faces keep file order across g/o groups, and objloader.h:23-139
// un-welds each triangle into three fresh vertices.  Floats via atof, indices via atoi, as tinyobj.
// ------------------------------------------------------------------------------------------------
static inline bool is_space(char c) { return c == ' ' || c == '\t'; }

void load_obj_soup(const std::string& path, std::vector<float>& out9)
{
    std::ifstream ifs(path, std::ios::binary);
    if (!ifs) throw std::runtime_error("Cannot open file [" + path + "]");
    std::stringstream ss;
    ss << ifs.rdbuf();
    const std::string text = ss.str();
    std::vector<float> v;
    v.reserve(1 << 16);
    std::vector<int> face;
    size_t pos = 0;
    const size_t n = text.size();
    while (pos < n) {
        size_t eol = pos;
        while (eol < n && text[eol] != '\n' && text[eol] != '\r') eol++;
        const char* token = text.c_str() + pos;
        const char* end = text.c_str() + eol;
        // advance to the next line (handles \n, \r\n and lone \r like tinyobj's safeGetline)
        pos = eol;
        if (pos < n) pos += (text[pos] == '\r' && pos + 1 < n && text[pos + 1] == '\n') ? 2 : 1;
        while (token < end && is_space(*token)) token++;
        if (token >= end || *token == '#') continue;
        if (token[0] == 'v' && token + 1 < end && is_space(token[1])) {
            // strtod stops at the first non-numeric character; the buffer is NUL-terminated at its end
            const char* q = token + 2;
            float xyz[3];
            for (int k = 0; k < 3; k++) {
                while (q < end && is_space(*q)) q++;
                xyz[k] = (q < end) ? (float)atof(q) : 0.0f;
                while (q < end && !is_space(*q)) q++;
            }
            v.push_back(xyz[0]); v.push_back(xyz[1]); v.push_back(xyz[2]);
            continue;
        }
        if (token[0] == 'f' && token + 1 < end && is_space(token[1])) {
            const char* q = token + 2;
            face.clear();
            const int vsize = (int)(v.size() / 3);
            while (true) {
                while (q < end && is_space(*q)) q++;
                if (q >= end) break;
                const int idx = atoi(q);
                int i;
                if (idx > 0) i = idx - 1;
                else if (idx == 0) i = 0;
                else i = vsize + idx;
                face.push_back(i);
                while (q < end && !is_space(*q)) q++;
            }
            for (size_t k = 2; k < face.size(); k++) {
                const int tri[3] = {face[0], face[k - 1], face[k]};
                for (int c = 0; c < 3; c++) {
                    if (tri[c] < 0 || (size_t)tri[c] * 3 + 2 >= v.size())
                        throw std::runtime_error("face references a vertex that does not exist in [" + path + "]");
                    out9.push_back(v[(size_t)tri[c] * 3 + 0]);
                    out9.push_back(v[(size_t)tri[c] * 3 + 1]);
                    out9.push_back(v[(size_t)tri[c] * 3 + 2]);
                }
            }
            continue;
        }
        // vn, vt, g, o, usemtl, mtllib, unknown commands: irrelevant to the collision mesh
    }
}

static void place_meshes(HostScene& s, const std::vector<std::vector<float>>& soups)
{
    int64_t total = 0;
    for (const auto& t : soups) total += (int64_t)(t.size() / 9);
    if (total > 0x7fffffff) throw std::runtime_error("too many triangles");
    s.tri_local.resize((size_t)total * 9);
    s.tri_mesh.resize((size_t)total);
    int64_t at = 0;
    for (size_t m = 0; m < s.meshes.size(); m++) {
        HostMesh& me = s.meshes[m];
        // scene.cpp:322-324: pos = deltas * scaling * scaling; position = pos + origin
        for (int a = 0; a < 3; a++) {
            const float p = me.deltas[a] * s.scaling * s.scaling;
            me.origin[a] = p + s.origin[a];
        }
        const size_t nt = soups[m].size() / 9;
        if (nt == 0) throw std::runtime_error("mesh [" + me.filename + "] has no triangles");   // reference: at(0) on empty array, scene.cpp:305
        me.tri_begin = at;
        for (size_t t = 0; t < nt; t++, at++) {
            s.tri_mesh[(size_t)at] = (int32_t)m;
            // btBvhTriangleMeshShape fetches vertices as v_obj * localScaling (scene.cpp:313-316)
            for (int k = 0; k < 9; k++) s.tri_local[(size_t)at * 9 + k] = soups[m][t * 9 + k] * s.scaling;
        }
        me.tri_end = at;
    }
}

static bool dir_exists(const std::string& p)
{
    struct stat st;
    return !p.empty() && stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

HostScene load_scene_file(const std::string& scene_path)
{
    HostScene s;
    try {
        std::ifstream ifs(scene_path, std::ios::binary);
        if (!ifs) throw std::runtime_error("cannot open scene file [" + scene_path + "]");
        std::stringstream ss;
        ss << ifs.rdbuf();
        const mcrt_json::Value cfg = mcrt_json::parse(ss.str());
        // scene.cpp:189
        const mcrt_json::Value* wd = cfg.find("workingDirectory");
        s.working_dir = wd ? wd->as_string() : "";
        if (!dir_exists(s.working_dir)) {    // extension: every example points at /home/santiago/...
            const size_t slash = scene_path.find_last_of('/');
            s.working_dir = slash == std::string::npos ? std::string("./") : scene_path.substr(0, slash + 1);
        }
        const mcrt_json::Value& t_pos = cfg.at("transducerPosition");        // scene.cpp:191, main.cpp:65
        for (int a = 0; a < 3; a++) s.start_pose[a] = t_pos.at(a).as_float();
        if (const mcrt_json::Value* t_dir = cfg.find("transducerAngles"))     // main.cpp:68 (at() there; optional here)
            for (int a = 0; a < 3; a++) s.start_pose[3 + a] = t_dir->at(a).as_float();
        const mcrt_json::Value& orig = cfg.at("origin");
        const mcrt_json::Value& spac = cfg.at("spacing");
        for (int a = 0; a < 3; a++) { s.origin[a] = orig.at(a).as_float(); s.spacing[a] = spac.at(a).as_float(); }
        const std::string starting = cfg.at("startingMaterial").as_string();
        s.scaling = cfg.at("scaling").as_float();
        const mcrt_json::Value& mats = cfg.at("materials");                   // scene.cpp:204-225
        if (!mats.is_array()) throw std::runtime_error("materials must be an array");
        for (const auto& mat : mats.arr) {
            HostMaterial m;
            m.impedance = mat.at("impedance").as_float();
            m.attenuation = mat.at("attenuation").as_float();
            m.mu0 = mat.at("mu0").as_float();
            m.mu1 = mat.at("mu1").as_float();
            m.sigma = mat.at("sigma").as_float();
            m.specularity = mat.at("specularity").as_float();
            // extension: examples/ircad11/ircad11.scene has neither key
            const mcrt_json::Value* sh = mat.find("shininess");
            const mcrt_json::Value* th = mat.find("thickness");
            m.shininess = sh ? sh->as_float() : 1000000.0f;
            m.thickness = th ? th->as_float() : 0.0f;
            const std::string name = mat.at("name").as_string();
            bool replaced = false;           // unordered_map::operator[]: a repeated name overwrites
            for (size_t i = 0; i < s.material_names.size(); i++)
                if (s.material_names[i] == name) { s.materials[i] = m; replaced = true; }
            if (!replaced) { s.material_names.push_back(name); s.materials.push_back(m); }
        }
        auto mat_id = [&](const std::string& name) -> int {
            for (size_t i = 0; i < s.material_names.size(); i++)
                if (s.material_names[i] == name) return (int)i;
            throw std::out_of_range("material '" + name + "' not found");
        };
        s.starting_material = mat_id(starting);                                // materials.at(starting_material), scene.cpp:90
        const mcrt_json::Value& meshes = cfg.at("meshes");                    // scene.cpp:227-246
        if (!meshes.is_array()) throw std::runtime_error("meshes must be an array");
        std::vector<std::vector<float>> soups;
        for (const auto& me : meshes.arr) {
            HostMesh hm;
            const mcrt_json::Value& deltas = me.at("deltas");
            hm.filename = me.at("file").as_string();
            hm.is_rigid = me.at("rigid").as_bool();
            hm.is_vascular = me.at("vascular").as_bool();
            for (int a = 0; a < 3; a++) hm.deltas[a] = deltas.at(a).as_float();
            hm.outside_normals = me.at("outsideNormals").as_bool();
            hm.material_inside = mat_id(me.at("material").as_string());
            hm.material_outside = mat_id(me.at("outsideMaterial").as_string());
            s.meshes.push_back(hm);
            soups.emplace_back();
            load_obj_soup(s.working_dir + hm.filename, soups.back());         // scene.cpp:42
        }
        place_meshes(s, soups);
    } catch (const std::exception& ex) {
        throw std::runtime_error("Error while loading scene: " + std::string(ex.what()));    // scene.cpp:23-26
    }
    return s;
}

HostScene scene_from_arrays(const mcrt_scene_arrays& a)
{
    HostScene s;
    try {
        if (a.n_materials < 1 || !a.materials8) throw std::runtime_error("materials must be a non-empty array");
        if (a.n_meshes < 0 || (a.n_meshes > 0 && (!a.tri_offsets || !a.tri_vertices))) throw std::runtime_error("meshes must be an array");
        s.scaling = a.scaling;
        for (int k = 0; k < 3; k++) { s.origin[k] = a.origin[k]; s.spacing[k] = a.spacing[k]; }
        for (int i = 0; i < a.n_materials; i++) {
            HostMaterial m;
            memcpy(&m, a.materials8 + 8 * i, sizeof(m));
            s.materials.push_back(m);
            s.material_names.push_back("material" + std::to_string(i));
        }
        auto chk = [&](int id) { if (id < 0 || id >= a.n_materials) throw std::out_of_range("material index out of range"); return id; };
        s.starting_material = chk(a.starting_material);
        std::vector<std::vector<float>> soups;
        for (int m = 0; m < a.n_meshes; m++) {
            HostMesh hm;
            hm.filename = "mesh" + std::to_string(m);
            hm.is_vascular = a.mesh_vascular ? a.mesh_vascular[m] != 0 : false;
            for (int k = 0; k < 3; k++) hm.deltas[k] = a.mesh_deltas ? a.mesh_deltas[3 * m + k] : 0.0f;
            hm.material_inside = chk(a.mesh_material_inside[m]);
            hm.material_outside = chk(a.mesh_material_outside[m]);
            s.meshes.push_back(hm);
            const int64_t b = a.tri_offsets[m], e = a.tri_offsets[m + 1];
            if (e < b) throw std::runtime_error("tri_offsets must be non-decreasing");
            soups.emplace_back(a.tri_vertices + b * 9, a.tri_vertices + e * 9);
        }
        place_meshes(s, soups);
    } catch (const std::exception& ex) {
        throw std::runtime_error("Error while loading scene: " + std::string(ex.what()));
    }
    return s;
}

// ------------------------------------------------------------------------------------------------
// transducer.h:24-62
// ------------------------------------------------------------------------------------------------
void element_angle_table(const mcrt_params& p, const Derived& d, std::vector<float>& sincos2)
{
    // amp = transducer_element_separation / radius: millimetres over centimetres, scalar ratio 1/10
    const double amp_raw = d.element_separation_mm / p.radius_cm;
    const float amp_f = (float)((amp_raw * 1) / 10);
    const double amplitude = (double)amp_f;                       // radian_t amplitude { amp.to<float>() }
    const double angle_center_of_element = amplitude / 2.0f;
    double angle = -(amplitude * (double)(size_t)p.elements / 2) + angle_center_of_element;
    sincos2.resize((size_t)p.elements * 2);
    for (int t = 0; t < p.elements; t++) {
        const float af = (float)angle;
        sincos2[2 * t] = std::sin(af);                            // std::sin(float) -> sinf
        sincos2[2 * t + 1] = std::cos(af);
        angle = angle + amplitude;
    }
}

PoseTrig pose_trig(const mcrt_pose& pose)
{
    PoseTrig t;
    memset(&t, 0, sizeof(t));
    for (int a = 0; a < 3; a++) t.pos[a] = pose.pos[a];
    // radian_t x_angle { degree_t }: ((deg * PI) * 1) / 180 in double, then .to<float>()
    const float xa = (float)((((double)pose.angles_deg[0]) * (kPi * 1.0) * 1) / 180);
    const float ya = (float)((((double)pose.angles_deg[1]) * (kPi * 1.0) * 1) / 180);
    const float za = (float)((((double)pose.angles_deg[2]) * (kPi * 1.0) * 1) / 180);
    t.cz = cosf(za); t.sz = sinf(za);                              // btCos / btSin in btVector3::rotate
    t.cx = cosf(xa); t.sx = sinf(xa);
    t.cy = cosf(ya); t.sy = sinf(ya);
    return t;
}

// ------------------------------------------------------------------------------------------------
// psf.h:34-58, 80-92
// ------------------------------------------------------------------------------------------------
void psf_taps(const mcrt_params& p, std::vector<float>& axial, std::vector<float>& lateral)
{
    const float half_axial = (size_t)p.psf_axial * (size_t)p.resolution_um / 1000.0f / 2.0f;
    const float half_lateral = (size_t)p.psf_lateral * (size_t)p.resolution_um / 1000.0f / 2.0f;
    const float resolution = p.resolution_um / 1000.0f;
    axial.resize(p.psf_axial);
    lateral.resize(p.psf_lateral);
    for (int i = 0; i < p.psf_axial; i++) {
        const float x = (size_t)i * resolution - half_axial;
        const double xx = (double)x * (double)x;                                  // pow(x, 2): float^int -> double
        axial[i] = (float)(std::exp(-0.5f * (xx / p.psf_var_x)) * std::cos(2 * kPiRedefined * p.frequency_mhz * x));
    }
    for (int i = 0; i < p.psf_lateral; i++) {
        const float y = (size_t)i * resolution - half_lateral;
        const double yy = (double)y * (double)y;
        lateral[i] = (float)std::exp(-0.5f * (yy / p.psf_var_y));
    }
}

void psf_elevation_taps(const mcrt_params& p, int n, float var_z, std::vector<float>& taps, std::vector<float>& z_mm)
{
    const float half_elevation = (size_t)n * (size_t)p.resolution_um / 1000.0f / 2.0f;       // psf.h:46
    const float resolution = p.resolution_um / 1000.0f;
    taps.resize(n); z_mm.resize(n);
    for (int i = 0; i < n; i++) {
        const float z = (size_t)i * resolution - half_elevation;
        const double zz = (double)z * (double)z;
        taps[i] = (float)std::exp(-0.5f * (zz / var_z));                                   // the form of lateral_function, psf.h:87-92
        z_mm[i] = z;
    }
}

mcrt_pose elevation_pose(const mcrt_pose& pose, float z_mm)
{
    // the fan lies in the transducer's local xy plane (transducer.h:45-59: directions (sin a, cos a, 0) rotated about Z, X, Y);
    // its normal, local (0, 0, 1), rotated the same way is (cx sy, -sx, cx cy) -- btVector3::rotate term by term
    const PoseTrig t = pose_trig(pose);
    const float ez[3] = {t.cx * t.sy, -t.sx, t.cx * t.cy};
    const float s = z_mm * 0.1f;                                                             // mm -> world units (cm)
    mcrt_pose q = pose;
    for (int a = 0; a < 3; a++) q.pos[a] = pose.pos[a] + ez[a] * s;
    return q;
}

// Depth-dependent lateral PSF (psf.h:11-25 says "lateral and elevation ranges vary according to distance to
// the transducer" but the reference fills ONE lateral kernel): the lateral Gaussian of RF row r has the variance
//   var_y(r) = var_y * w^2,  w = 1 + spread * |depth(r) - focus| / focus,  depth(r) = r * depth_cm / rows,
// i.e. the beam is narrowest at the focus.  Same float / double mix as psf_taps (psf.h:80-92).  table: [kl][rows].
void psf_lateral_depth_table(const mcrt_params& p, int rows, float focus_cm, float spread, std::vector<float>& table)
{
    const float half_lateral = (size_t)p.psf_lateral * (size_t)p.resolution_um / 1000.0f / 2.0f;
    const float resolution = p.resolution_um / 1000.0f;
    table.assign((size_t)p.psf_lateral * rows, 0.0f);
    for (int r = 0; r < rows; r++) {
        const double depth = (double)r * p.depth_cm / (double)rows;
        const float w = (float)(1.0 + (double)spread * std::fabs(depth - (double)focus_cm) / (double)focus_cm);
        const float var = p.psf_var_y * w * w;
        for (int i = 0; i < p.psf_lateral; i++) {
            const float y = (size_t)i * resolution - half_lateral;
            const double yy = (double)y * (double)y;
            table[(size_t)i * rows + r] = (float)std::exp(-0.5f * (yy / var));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// rf_image::create_mapping (rfimage.h:183-215): map_x = source row, map_y = source column
// ------------------------------------------------------------------------------------------------
void scan_mapping(const mcrt_params& p, const Derived& d, std::vector<float>& map_x, std::vector<float>& map_y)
{
    const int srows = p.scan_rows, scols = p.scan_cols;
    const float radius_f = (float)((p.radius_cm * 10) / 1);                       // millimeter_t radius
    const double total_angle = ((p.fov_deg * (kPi * 1.0) * 1) / 180);
    const float total_angle_f = (float)total_angle;
    const float depth_f = (d.max_travel_time_u * p.speed_of_sound) * 0.001f;
    const float ratio = (float)(((depth_f + radius_f) - radius_f * std::cos(total_angle_f / 2.0)) / srows);
    const double shift_y = ((double)radius_f) * (double)std::cos(total_angle_f / 2.0f);
    const float half_width = (float)scols / 2.0f;
    map_x.resize((size_t)srows * scols);
    map_y.resize((size_t)srows * scols);
    for (int j = 0; j < scols; j++)
        for (int i = 0; i < srows; i++) {
            const float fi = static_cast<float>(i) + (float)shift_y / ratio;
            const float fj = static_cast<float>(j) - half_width;
            const float r = std::sqrt(std::pow(fi, 2.0f) + std::pow(fj, 2.0f));
            const double angle = (double)std::atan2(fj, fi);
            map_x[(size_t)i * scols + j] = (r * ratio - radius_f) / depth_f * (float)d.rows;
            map_y[(size_t)i * scols + j] = (float)(((angle - (-(total_angle / 2))) / total_angle) * (float)d.cols);
        }
}

// ------------------------------------------------------------------------------------------------
// volume.h:19-35.  The stream is libstdc++-specific (minstd_rand0 + Marsaglia polar with a cached
// second variate), so it is generated with the same <random> calls on the host and uploaded once.
// ------------------------------------------------------------------------------------------------
const std::vector<float>& scatterer_volume()
{
    static std::vector<float> vol;
    static std::once_flag once;
    std::call_once(once, [] {
        const size_t n = (size_t)256 * 256 * 256;
        vol.resize(n * 2);
        std::default_random_engine generator;
        std::normal_distribution<double> distribution(0.0, 1.0);
        for (size_t i = 0; i < n; i++) {
            vol[2 * i] = (float)distribution(generator);          // texture_noise
            vol[2 * i + 1] = (float)distribution(generator);      // scattering_probability
        }
    });
    return vol;
}

}  // namespace mcrt
