// mattausch -- headless drop-in for the reference executable (`mattausch <scene-file>`,
// main.cpp:42-50): loads the scene JSON, simulates frames on the GPU through the C ABI and writes
// the RF image (raw little-endian float32) and the scan-converted B-mode image (8-bit PGM, the
// x255 conversion of rf_image::save, rfimage.h:142-148) instead of opening an imshow window.
// Extensions (all optional, reference defaults otherwise):
//   --frames N  --seed S  --elements E  --samples S  --deterministic  --out DIR  --device D  --log-compress
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/mcrt.h"

static bool write_pgm(const std::string& path, const float* img, int rows, int cols)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P5\n%d %d\n255\n", cols, rows);
    std::vector<unsigned char> line(cols);
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) {
            const float v = img[(size_t)r * cols + c] * 255.0f;      // convertTo(CV_8U, 255.0): round + saturate
            const long q = std::isnan(v) ? 0 : lrintf(v);
            line[c] = (unsigned char)(q < 0 ? 0 : (q > 255 ? 255 : q));
        }
        fwrite(line.data(), 1, cols, f);
    }
    fclose(f);
    return true;
}

int main(int argc, char** argv)
{
    if (argc < 2 || argv[1][0] == '-') {
        printf("Incorrect argument list.\n");                        // main.cpp:46-50
        return 0;
    }
    mcrt_params p;
    mcrt_default_params(&p);
    int frames = 1, device = 0, log_compress = 0;
    unsigned long long seed = 0;
    std::string out_dir = ".";
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : "0"; };
        if (a == "--frames") frames = atoi(next());
        else if (a == "--seed") seed = strtoull(next(), nullptr, 10);
        else if (a == "--elements") p.elements = atoi(next());
        else if (a == "--samples") p.samples = atoi(next());
        else if (a == "--deterministic") p.deterministic = 1;
        else if (a == "--out") out_dir = next();
        else if (a == "--device") device = atoi(next());
        else if (a == "--log-compress") log_compress = 1;        // rfimage.h:131-136 (commented out in the reference)
        else { printf("Incorrect argument list.\n"); return 0; }
    }
    mcrt_ctx* ctx = nullptr;
    if (mcrt_create(argv[1], &p, device, &ctx) != MCRT_OK) {
        // main.cpp:154-159
        printf("The program found an error and will terminate.\nReason:\n%s\n", mcrt_last_error());
        return 0;
    }
    if (log_compress) mcrt_set_option(ctx, "log_compress", 1);
    mcrt_info info;
    mcrt_get_info(ctx, &info);
    printf("%g us\n", info.max_travel_time_us);                      // main.cpp:76
    printf("rf_image: %d, %d\n", info.rows, info.cols);              // rfimage.h:29
    mcrt_pose pose;
    for (int k = 0; k < 3; k++) { pose.pos[k] = info.start_pose[k]; pose.angles_deg[k] = info.start_pose[3 + k]; }
    std::vector<float> rf((size_t)info.rows * info.cols), scan((size_t)info.scan_rows * info.scan_cols);
    for (int f = 0; f < frames; f++) {
        if (mcrt_simulate(ctx, &pose, 1, seed, (uint64_t)f, rf.data(), scan.data()) != MCRT_OK) {
            printf("The program found an error and will terminate.\nReason:\n%s\n", mcrt_last_error());
            break;
        }
        mcrt_stats st;
        mcrt_get_stats(ctx, &st);
        // scene.cpp:178-179 prints "fps tests total_collisions"
        printf("%g %lld %d\n", st.ms_total > 0 ? 1000.0 / st.ms_total : 0.0, (long long)st.segments, p.samples * p.elements);
        char name[64];
        snprintf(name, sizeof(name), "/rf_%04d.f32", f);
        if (FILE* fh = fopen((out_dir + name).c_str(), "wb")) { fwrite(rf.data(), sizeof(float), rf.size(), fh); fclose(fh); }
        snprintf(name, sizeof(name), "/mattausch_%04d.pgm", f);
        write_pgm(out_dir + name, scan.data(), info.scan_rows, info.scan_cols);
    }
    mcrt_destroy(ctx);
    return 0;
}
