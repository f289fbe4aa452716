// mattausch -- headless drop-in for the reference executable (`mattausch <scene-file>`,
// main.cpp:42-50): loads the scene JSON, simulates frames on the GPU through the C ABI and writes
// the RF image (raw little-endian float32) and the scan-converted B-mode image (8-bit PGM, the
// x255 conversion of rf_image::save, rfimage.h:142-148) instead of opening an imshow window.
// Extensions (all optional, reference defaults otherwise):
//   --frames N  --seed S  --elements E  --samples S  --deterministic  --out DIR  --device D  --log-compress
//   --png                      also write the scan-converted image as 8-bit PNG (frame 0 under the reference's name,
//                              prelog.png, rfimage.h:147)
//   --bmode DR [--gain dB] [--tgc dB/cm]   B-mode display chain (mcrt_bmode): TGC, log compression to DR dB -> bmode_NNNN.png
//   --poses FILE [--gpus G] [--batch B]    probe sweep: one pose per line (x y z ax ay az), frame index = line index; the poses are
//                              split into G contiguous blocks, one context + host thread per GPU (devices D .. D+G-1), B poses per
//                              call; all frames go to sweep_rf.f32 ([pose][scanline][sample] float32)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <chrono>
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

// Minimal 8-bit grayscale PNG: zlib stream of *stored* deflate blocks (no compression library needed).
static unsigned crc32_update(unsigned crc, const unsigned char* p, size_t n)
{
    static unsigned table[256];
    static bool init = false;
    if (!init) {
        for (unsigned i = 0; i < 256; i++) { unsigned c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
static void put_be32(std::vector<unsigned char>& v, unsigned x) { for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s)); }
static void png_chunk(FILE* f, const char* type, const std::vector<unsigned char>& data)
{
    std::vector<unsigned char> head; put_be32(head, (unsigned)data.size());
    fwrite(head.data(), 1, 4, f);
    std::vector<unsigned char> body(type, type + 4);
    body.insert(body.end(), data.begin(), data.end());
    fwrite(body.data(), 1, body.size(), f);
    std::vector<unsigned char> tail; put_be32(tail, crc32_update(0xffffffffu, body.data(), body.size()) ^ 0xffffffffu);
    fwrite(tail.data(), 1, 4, f);
}
static bool write_png8(const std::string& path, const unsigned char* img, int rows, int cols)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr; put_be32(ihdr, (unsigned)cols); put_be32(ihdr, (unsigned)rows);
    const unsigned char rest[5] = {8, 0, 0, 0, 0};                 // 8 bit, grayscale, deflate, no filter, no interlace
    ihdr.insert(ihdr.end(), rest, rest + 5);
    png_chunk(f, "IHDR", ihdr);
    std::vector<unsigned char> raw;                                // scanlines, each prefixed by filter type 0
    raw.reserve((size_t)rows * (cols + 1));
    for (int r = 0; r < rows; r++) { raw.push_back(0); raw.insert(raw.end(), img + (size_t)r * cols, img + (size_t)(r + 1) * cols); }
    std::vector<unsigned char> z = {0x78, 0x01};
    unsigned a = 1, b = 0;                                         // adler32
    for (unsigned char c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        const size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + n >= raw.size() ? 1 : 0);               // BFINAL, BTYPE = 00 (stored)
        z.push_back((unsigned char)(n & 0xff)); z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xff)); z.push_back((unsigned char)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        if (raw.empty()) break;
    }
    put_be32(z, (b << 16) | a);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
    fclose(f);
    return true;
}

static void quantize8(const float* img, size_t n, std::vector<unsigned char>& out)
{
    out.resize(n);
    for (size_t i = 0; i < n; i++) {
        const float v = img[i] * 255.0f;                             // convertTo(CV_8U, 255.0): round + saturate
        const long q = std::isnan(v) ? 0 : lrintf(v);
        out[i] = (unsigned char)(q < 0 ? 0 : (q > 255 ? 255 : q));
    }
}

int main(int argc, char** argv)
{
    if (argc < 2 || argv[1][0] == '-') {
        printf("Incorrect argument list.\n");                        // main.cpp:46-50
        return 0;
    }
    mcrt_params p;
    mcrt_default_params(&p);
    int frames = 1, device = 0, log_compress = 0, png = 0, bmode = 0, gpus = 1, batch = 64;
    int elevation = 1, ray_tree = 0, psf_depth = 0;
    float elevation_var = 0.0f, psf_focus = 0.0f, psf_spread = 0.0f;
    std::string poses_file;
    mcrt_bmode_params bp = {0.0f, 0.0f, 60.0f, 0.0f};
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
        else if (a == "--png") png = 1;
        else if (a == "--poses") poses_file = next();
        else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--batch") batch = atoi(next());
        else if (a == "--bmode") { bmode = 1; bp.dynamic_range_db = (float)atof(next()); }
        else if (a == "--gain") bp.gain_db = (float)atof(next());
        else if (a == "--tgc") bp.tgc_db_per_cm = (float)atof(next());
        else if (a == "--elevation") { elevation = atoi(next()); elevation_var = (float)atof(next()); }    // fans, variance [mm^2] (psf.h:16-18)
        else if (a == "--psf-depth") { psf_depth = 1; psf_focus = (float)atof(next()); psf_spread = (float)atof(next()); }   // focus [cm], spread
        else if (a == "--ray-tree") ray_tree = atoi(next());         // follow both children of every hit; segment budget per path
        else { printf("Incorrect argument list.\n"); return 0; }
    }
    // the optional extensions of the display / physics chain; false: the library's message is in mcrt_last_error()
    auto configure = [&](mcrt_ctx* c) -> bool {
        if (log_compress && mcrt_set_option(c, "log_compress", 1) != MCRT_OK) return false;
        if (ray_tree > 0 && mcrt_set_option(c, "ray_tree", ray_tree) != MCRT_OK) return false;
        if (psf_depth && mcrt_set_psf_depth_profile(c, psf_focus, psf_spread, nullptr) != MCRT_OK) return false;
        if (elevation > 1 && mcrt_set_elevation(c, elevation, elevation_var, nullptr, nullptr) != MCRT_OK) return false;
        return true;
    };
    if (!poses_file.empty()) {
        // ---- probe sweep sharded over the GPUs of this box (BASELINE config 3), C++ host: one context and one thread per GPU ----
        std::vector<mcrt_pose> poses;
        if (FILE* f = fopen(poses_file.c_str(), "r")) {
            char line[512];
            while (fgets(line, sizeof(line), f)) {
                mcrt_pose q;
                if (line[0] == '#') continue;
                if (sscanf(line, "%f %f %f %f %f %f", &q.pos[0], &q.pos[1], &q.pos[2], &q.angles_deg[0], &q.angles_deg[1], &q.angles_deg[2]) == 6)
                    poses.push_back(q);
            }
            fclose(f);
        }
        if (poses.empty() || gpus < 1 || batch < 1) { printf("Incorrect argument list.\n"); return 0; }
        std::vector<mcrt_ctx*> ctxs(gpus, nullptr);
        for (int g = 0; g < gpus; g++) {                               // contexts are created one after the other
            if (mcrt_create(argv[1], &p, device + g, &ctxs[g]) != MCRT_OK || !configure(ctxs[g])) {
                printf("The program found an error and will terminate.\nReason:\n%s\n", mcrt_last_error());
                for (mcrt_ctx* c : ctxs) if (c) mcrt_destroy(c);
                return 0;
            }
            mcrt_set_option(ctxs[g], "max_batch_poses", batch);
        }
        mcrt_info info;
        mcrt_get_info(ctxs[0], &info);
        printf("%g us\n", info.max_travel_time_us);
        printf("rf_image: %d, %d\n", info.rows, info.cols);
        const size_t px = (size_t)info.rows * info.cols, n = poses.size();
        std::vector<float> all(px * n);
        std::vector<int> failed(gpus, 0);
        std::vector<long long> segs(gpus, 0);
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> workers;
        for (int g = 0; g < gpus; g++) {
            workers.emplace_back([&, g]() {
                // contiguous block of GPU g: the first (n % gpus) blocks get one pose more (sweep.py::shard_bounds)
                const size_t base = n / gpus, extra = n % gpus;
                const size_t b = g * base + ((size_t)g < extra ? (size_t)g : extra), e = b + base + ((size_t)g < extra ? 1 : 0);
                for (size_t i = b; i < e; i += (size_t)batch) {
                    const int m = (int)((e - i) < (size_t)batch ? (e - i) : (size_t)batch);
                    if (mcrt_simulate(ctxs[g], &poses[i], m, seed, (uint64_t)i, all.data() + i * px, nullptr) != MCRT_OK) {
                        fprintf(stderr, "GPU %d: %s\n", device + g, mcrt_last_error());
                        failed[g] = 1;
                        return;
                    }
                    mcrt_stats st;
                    mcrt_get_stats(ctxs[g], &st);
                    segs[g] += (long long)st.segments;
                }
            });
        }
        for (auto& w : workers) w.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        long long total_segs = 0;
        int any_failed = 0;
        for (int g = 0; g < gpus; g++) { total_segs += segs[g]; any_failed |= failed[g]; }
        for (mcrt_ctx* c : ctxs) mcrt_destroy(c);
        if (any_failed) { printf("The program found an error and will terminate.\n"); return 0; }
        // "fps tests total_collisions" of scene.cpp:178-179, for the whole sweep
        printf("%g %lld %d\n", dt > 0 ? (double)n / dt : 0.0, total_segs, p.samples * p.elements);
        if (FILE* fh = fopen((out_dir + "/sweep_rf.f32").c_str(), "wb")) { fwrite(all.data(), sizeof(float), all.size(), fh); fclose(fh); }
        return 0;
    }
    mcrt_ctx* ctx = nullptr;
    if (mcrt_create(argv[1], &p, device, &ctx) != MCRT_OK || !configure(ctx)) {
        // main.cpp:154-159
        printf("The program found an error and will terminate.\nReason:\n%s\n", mcrt_last_error());
        if (ctx) mcrt_destroy(ctx);
        return 0;
    }
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
        std::vector<unsigned char> img8;
        if (png) {
            quantize8(scan.data(), scan.size(), img8);
            if (f == 0) write_png8(out_dir + "/prelog.png", img8.data(), info.scan_rows, info.scan_cols);        // rfimage.h:147
            snprintf(name, sizeof(name), "/mattausch_%04d.png", f);
            write_png8(out_dir + name, img8.data(), info.scan_rows, info.scan_cols);
        }
        if (bmode) {
            img8.resize(scan.size());
            if (mcrt_bmode(ctx, rf.data(), 1, &bp, nullptr, img8.data()) != MCRT_OK) {
                printf("The program found an error and will terminate.\nReason:\n%s\n", mcrt_last_error());
                break;
            }
            snprintf(name, sizeof(name), "/bmode_%04d.png", f);
            write_png8(out_dir + name, img8.data(), info.scan_rows, info.scan_cols);
        }
    }
    mcrt_destroy(ctx);
    return 0;
}
