/* mcrt_oracle.h -- C ABI of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a CPU restatement of the reference's per-frame simulation hot path
 * (thepochynsons/MCRay-Tracing: src/scene.cpp:50-183, src/ray.cpp, src/transducer.h,
 * src/main.cpp:102-148, src/volume.h, src/psf.h, src/rfimage.h:33-140,183-215).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The shipped library (libmcrt.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: the pieces of the reference that compile here (psf.h, volume.h, transducer.h,
 * ray.cpp, rfimage.h, tinyobj+objloader.h) are pinned bit-exactly through oracle/_ref
 * (golden vectors under tests/golden, tests/test_oracle_pins.py).  The closest-hit query (Bullet's
 * btCollisionWorld::rayTest, scene.cpp:115-117) and cv::remap (rfimage.h:139) live in
 * un-vendored, un-pinned third-party code: Bullet's triangle ray-cast is restated from its
 * published algorithm (PARITY UNPINNED for that call), cv::remap is pinned against cv2.remap
 * of opencv-python-headless 4.13.
 */
#ifndef MCRT_ORACLE_H
#define MCRT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Acquisition parameters: main.cpp:23-37 made runtime (defaults = the reference's constexprs). */
typedef struct orc_params {
    int32_t elements;          /* transducer_elements   512   main.cpp:26 */
    int32_t samples;           /* samples_te            5     main.cpp:27 */
    int32_t max_depth;         /* ray::max_depth        10    ray.h:23    */
    float   frequency_mhz;     /* transducer_frequency  4.5f  main.cpp:24 */
    double  radius_cm;         /* transducer_radius     3     main.cpp:29 */
    double  fov_deg;           /* transducer_amplitude  60    main.cpp:28 */
    double  depth_cm;          /* ultrasound_depth      15    main.cpp:30 */
    uint32_t speed_of_sound;   /* 1500                        main.cpp:23, rfimage.h:19 */
    uint32_t resolution_um;    /* 145 (psf / volume grid)     main.cpp:33 */
    int32_t psf_axial;         /* 7                           main.cpp:34 */
    int32_t psf_lateral;       /* 13 */
    float   psf_var_x;         /* 0.05f                       main.cpp:54 */
    float   psf_var_y;         /* 0.2f */
    int32_t deterministic;     /* cos(theta')=1, q=0 */
    int32_t scan_rows;         /* 400                         rfimage.h:26 */
    int32_t scan_cols;         /* 500 */
    float   axial_scale;       /* 1.0f; >1 refines the axial grid (extension for BASELINE config 5):
                                  axial_resolution = (1.45f/frequency)/axial_scale */
    int32_t reserved;
} orc_params;

/* Quantities derived from orc_params exactly as main.cpp:25,31,36 / rfimage.h:180 / main.cpp:66 do */
typedef struct orc_derived {
    double  axial_resolution_mm;   /* (double)(1.45f/f)            main.cpp:25 */
    float   axial_resolution_f;    /* axial_resolution.to<float>() */
    double  max_travel_time_us;    /* main.cpp:31 */
    uint32_t max_travel_time_u;    /* .to<unsigned int>()          main.cpp:36 */
    uint32_t rf_axial_um;          /* (unsigned)(axres_f*1000.0f)  main.cpp:36 */
    int32_t rows;                  /* rfimage.h:180 */
    int32_t cols;
    double  element_separation_mm; /* main.cpp:66 */
    double  time_step_us;          /* rf_image.micros_traveled(axial_resolution) main.cpp:118 */
    double  row_period_us;         /* axial_resolution_/speed_of_sound_  rfimage.h:35 */
} orc_derived;

/* One emitted ray segment (ray.h:28-36) plus the parity-debug fields the reference does not keep. */
typedef struct orc_segment {
    float from[3];
    float to[3];
    float dir[3];
    float reflected_intensity;
    float initial_intensity;
    float attenuation;
    double distance_traveled;      /* mm */
    int32_t media_id;              /* SURVEY B-1: medium copied by value (as id) at emplace time */
    int32_t tri_id;                /* global triangle id of the hit closing this segment, -1 = miss */
    int32_t mesh_id;               /* -1 = miss */
    float hit_fraction;            /* Bullet's m_closestHitFraction; 1.0f on a miss */
} orc_segment;

typedef struct orc_scene orc_scene;
typedef struct orc_volume orc_volume;

void orc_default_params(orc_params* p);
void orc_derive(const orc_params* p, orc_derived* d);

/* materials: n_mat x 8 floats {impedance, attenuation, mu0, mu1, sigma, specularity, shininess,
 * thickness} (mesh.h:7-10).  Meshes: raw OBJ-space triangle soup, 9 floats per triangle, in the
 * order objloader.h:23-139 produces; tri_offsets has n_mesh+1 entries.  deltas: n_mesh x 3. */
orc_scene* orc_scene_create(int32_t n_mat, const float* materials8, int32_t starting_material,
                            int32_t n_mesh, const int32_t* mesh_material_inside,
                            const int32_t* mesh_material_outside, const int32_t* mesh_vascular,
                            const float* mesh_deltas, const int64_t* tri_offsets,
                            const float* tri_vertices_obj, float scaling, const float* origin3,
                            const float* spacing3);
void orc_scene_destroy(orc_scene* s);
int64_t orc_scene_num_triangles(const orc_scene* s);
/* local-frame vertices (v_obj * scaling, scene.cpp:313-316) and per-mesh body origins (scene.cpp:322-324) */
void orc_scene_get_local_vertices(const orc_scene* s, float* out9);
void orc_scene_get_mesh_origins(const orc_scene* s, float* out3);

/* closest hit of the segment [from,to] exactly as scene.cpp:115-126 queries it (caller passes the
 * already offset `from`).  use_bvh=0: brute force over every triangle; 1: the oracle's own BVH.
 * Returns tri id or -1; out_f = {fraction, point xyz, normal xyz}; out_mesh = mesh id. */
int32_t orc_closest_hit(const orc_scene* s, const float* from3, const float* to3, int32_t use_bvh,
                        float* out_f7, int32_t* out_mesh);

/* transducer.h:24-62 */
void orc_transducer_elements(const orc_params* p, const float* pos3, const float* angles_deg3,
                             float* out_pos, float* out_dir);

/* psf.h:34-58 */
void orc_psf_taps(const orc_params* p, float* axial, float* lateral);

/* volume.h */
orc_volume* orc_volume_get(void);                       /* process-wide singleton, 256^3 x 2 floats */
const float* orc_volume_raw(const orc_volume* v);
float orc_volume_get_scattering(const orc_volume* v, float density, float mu, float sigma, float x, float y, float z);

/* ray.cpp pieces, exposed for pinning */
float orc_max_ray_length(float attenuation, float intensity, float frequency);
void  orc_travel(float attenuation, float intensity, float frequency, double dist0, double mm, float* out_i, double* out_d);
float orc_reflection_intensity(float i_in, float z1, float c1, float z2, float c2);
float orc_reflected_intensity_eq8(const float* d3, const float* refr3, const float* refl3, float specularity);
void  orc_snells_law(const float* l3, const float* n3, float c1, float c2, float ratio, float* out3);
void  orc_random_unit_vector(const float* v3, float cos_theta, double u_azimuth, double u_radius, float* out3);
double orc_distance_in_mm(const orc_scene* s, const float* a3, const float* b3);
/* hit_boundary (ray.cpp:11-97) on an explicit path state.  state_i = {media_id, outside_code
 * (-1 null, -2 SELF, >=0 id), depth}; rng_u = {u_shininess, u_choice(float as double), u_az, u_rad}.
 * force_branch: -1 use u_choice, 0 refraction, 1 reflection.  out_f as ref_world_hit. */
void orc_hit_boundary(const orc_scene* s, const float* from3, const float* dir3, float intensity,
                      const int32_t* state_i3, const float* hit_point3, const float* normal3,
                      int32_t mesh_id, int32_t deterministic, const double* rng_u4, int32_t force_branch,
                      float* out_f8, int32_t* out_i3, int32_t* out_branch);

/* scene.cpp:50-183.  segments: [elements][samples][max_depth]; n_segments: [elements][samples].
 * use_bvh as orc_closest_hit.  Returns total number of closest-hit queries (the reference's `tests`). */
int64_t orc_cast_rays(const orc_scene* s, const orc_params* p, const float* pos3, const float* angles_deg3,
                      uint64_t seed, uint32_t frame, int32_t use_bvh, orc_segment* segments, int32_t* n_segments);

/* main.cpp:106-144.  rf: rows x cols float32 row-major (cv::Mat(max_rows, columns)); accumulated
 * into (caller clears).  Returns number of march steps taken. */
int64_t orc_accumulate(const orc_scene* s, const orc_params* p, const orc_volume* v,
                       const orc_segment* segments, const int32_t* n_segments, float* rf);

/* rfimage.h:93-123 / :54-91 on a rows x cols row-major image, in place */
void orc_convolve(float* rf, int32_t rows, int32_t cols, const float* axial, int32_t n_axial, const float* lateral, int32_t n_lateral);
void orc_envelope(float* rf, int32_t rows, int32_t cols);
/* rfimage.h:127-136, the log compression the reference keeps commented out; in place */
void orc_log_compress(float* rf, int32_t rows, int32_t cols);
/* elevational PSF: taps / fan offsets, and the probe position of the fan z_mm off the imaging plane */
void orc_elevation_taps(const orc_params* p, int32_t n, float var_z, float* taps, float* z_mm);
void orc_elevation_position(const float* pos3, const float* angles_deg3, float z_mm, float* out_pos3);
void orc_psf_depth_table(const orc_params* p, int32_t rows, float focus_cm, float spread, float* table);
void orc_convolve_depth(float* rf, int32_t rows, int32_t cols, const float* axial, int32_t n_axial, const float* lateral_by_row, int32_t n_lateral);
int64_t orc_cast_rays_tree(const orc_scene* s, const orc_params* p, const float* pos3, const float* angles_deg3, uint64_t seed, uint32_t frame,
                           int32_t use_bvh, int64_t capacity, orc_segment* segments, int32_t* seg_path, int32_t* seg_node);
int64_t orc_accumulate_flat(const orc_scene* s, const orc_params* p, const orc_volume* vol, const orc_segment* segments, const int32_t* seg_path,
                            int64_t n_segments, float* rf);
void orc_bmode(float* rf, int32_t rows, int32_t cols, double depth_cm, float gain_db, float tgc_db_per_cm, float dynamic_range_db);
/* rfimage.h:183-215: map_x (source row), map_y (source column), each scan_rows x scan_cols */
void orc_create_mapping(const orc_params* p, float* map_x, float* map_y);
/* rfimage.h:139: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) restated (OpenCV 5-bit fixed-point weights) */
void orc_scan_convert(const float* rf, int32_t rows, int32_t cols, const float* map_x, const float* map_y,
                      int32_t scan_rows, int32_t scan_cols, float* out);

/* whole frame: clear, cast, accumulate, convolve, envelope, (scan convert if scan_out != NULL).
 * stage_seconds (nullable) = {cast, accumulate, convolve+envelope, scan}.  Returns `tests`. */
int64_t orc_simulate_frame(const orc_scene* s, const orc_params* p, const float* pos3, const float* angles_deg3,
                           uint64_t seed, uint32_t frame, float* rf, float* scan_out, double* stage_seconds4,
                           int64_t* steps_out);

/* the shared numerics contract evaluated on the host; op codes as mcrt_numerics_probe (include/mcrt.h) */
void orc_numerics(int32_t op, int64_t n, const double* a, const double* b, double* out);

/* number of OpenMP threads used over elements (1 = the reference's single thread, scene.cpp:74) */
void orc_set_threads(int32_t n);
int32_t orc_get_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
