// grt_headless - the host side of the hot path in C++ over the C ABI (include/grt.h), without a window.
//
// Stands where the reference's frame loop stands (/root/reference/main.cpp:1503-2871): content folders
// (main.cpp:1562-1586), `-bench <metric>` / `-start <metric>` (main.cpp:956-986), the default camera
// (main.cpp:664-673), the per-frame sequence (main.cpp:2244-2526) as one grt_render call, the
// "Frametime Elapsed: %f" line of -bench (main.cpp:2864-2871) and the screenshot path
// (main.cpp:2762-2831: linear -> sRGB, written here as a binary PPM instead of PNG).
// The window, ImGui, input and Steam parts of main.cpp are out of scope.
//
//   grt_headless -content <folder> -bench <metric name> [-size W H] [-frames N] [-inflight K]
//                [-param name value]... [-static] [-no-adaptive] [-redshift] [-geodesic tau] [-background file.png|file.ppm] [-out frame.ppm]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "grt.h"

namespace {

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "grt_headless: %s: %s\n", what, grt_last_error());
    exit(1);
}
void check(int rc, const char* what) { if (rc < 0) die(what); }

// a procedural RGBA8 sky (bands, a lat/long grid): backgrounds enter the ABI as pixels, PNG decoding is the caller's
std::vector<uint8_t> procedural_sky(int w, int h) {
    std::vector<uint8_t> px((size_t)w * h * 4);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float u = (float)x / w, v = (float)y / h;
            float r = 0.25f + 0.25f * sinf(6.2831853f * (3 * u + v)), g = 0.25f + 0.25f * sinf(6.2831853f * (5 * v - u) + 1.0f), b = 0.35f + 0.30f * cosf(6.2831853f * (2 * u + 2 * v));
            bool grid = fabsf(fmodf(u * 24, 1.0f) - 0.5f) > 0.47f || fabsf(fmodf(v * 12, 1.0f) - 0.5f) > 0.47f;
            if (grid) r = g = b = 0.9f;
            uint8_t* p = &px[((size_t)y * w + x) * 4];
            p[0] = (uint8_t)(r * 255); p[1] = (uint8_t)(g * 255); p[2] = (uint8_t)(b * 255); p[3] = 255;
        }
    return px;
}

bool read_ppm(const char* path, std::vector<uint8_t>& rgba, int& w, int& h) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    int maxv = 0;
    if (fscanf(f, "P6 %d %d %d", &w, &h, &maxv) != 3 || maxv != 255) { fclose(f); return false; }
    fgetc(f);
    std::vector<uint8_t> rgb((size_t)w * h * 3);
    bool ok = fread(rgb.data(), 1, rgb.size(), f) == rgb.size();
    fclose(f);
    rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) { rgba[i * 4] = rgb[i * 3]; rgba[i * 4 + 1] = rgb[i * 3 + 1]; rgba[i * 4 + 2] = rgb[i * 3 + 2]; rgba[i * 4 + 3] = 255; }
    return ok;
}

void write_ppm(const char* path, const std::vector<uint8_t>& rgba, int w, int h) {
    FILE* f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "grt_headless: cannot write %s\n", path); return; }
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    for (size_t i = 0; i < (size_t)w * h; i++) fwrite(&rgba[i * 4], 1, 3, f);
    fclose(f);
}

}  // namespace

int main(int argc, char** argv) {
    std::string content_dir, metric_name = "kerr_schild", out_path, background_path;
    int width = 1920, height = 1080, frames = 8, inflight = 2;
    bool is_static = false, adaptive = true, redshift = false, print_frametime = false, on_geodesic = false;
    float geodesic_tau = 0.0f;
    std::vector<std::pair<std::string, float>> params;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](int k = 1) -> const char* { if (i + k >= argc) { fprintf(stderr, "grt_headless: %s needs a value\n", a.c_str()); exit(2); } return argv[i + k]; };
        if (a == "-content") { content_dir = next(); i++; }
        else if (a == "-bench") { metric_name = next(); print_frametime = true; i++; }       // main.cpp:975-980
        else if (a == "-start") { metric_name = next(); i++; }                               // main.cpp:982-985
        else if (a == "-size") { width = atoi(next(1)); height = atoi(next(2)); i += 2; }
        else if (a == "-frames") { frames = atoi(next()); i++; }
        else if (a == "-inflight") { inflight = atoi(next()); i++; }
        else if (a == "-param") { params.push_back({next(1), (float)atof(next(2))}); i += 2; }
        else if (a == "-static") is_static = true;
        else if (a == "-no-adaptive") adaptive = false;
        else if (a == "-redshift") redshift = true;
        else if (a == "-geodesic") { on_geodesic = true; geodesic_tau = (float)atof(next()); i++; }
        else if (a == "-background") { background_path = next(); i++; }
        else if (a == "-out") { out_path = next(); i++; }
        else { fprintf(stderr, "grt_headless: unknown argument %s\n", a.c_str()); return 2; }
    }
    if (content_dir.empty()) { fprintf(stderr, "usage: grt_headless -content <folder> -bench <metric name> [-size W H] [-frames N] [-out frame.ppm]\n"); return 2; }

    printf("Started (%s)\n", grt_version());
    grt_content* content = grt_content_create();
    check(grt_content_add_folder(content, content_dir.c_str()), "content folder");
    grt_metric* metric = grt_metric_load_by_name(content, metric_name.c_str());
    if (!metric) die("metric");                                   // the reference prints the script error and keeps the old metric
    for (auto& p : params) {                                      // the $cfg sliders, main.cpp:2120-2160
        int n = grt_metric_param_count(metric), found = -1;
        for (int k = 0; k < n; k++) if (p.first == grt_metric_param_name(metric, k)) found = k;
        if (found < 0) { fprintf(stderr, "grt_headless: metric has no parameter %s\n", p.first.c_str()); return 2; }
        check(grt_metric_param_set(metric, found, p.second), "parameter");
    }
    grt_ctx* ctx = grt_create(0);
    if (!ctx) die("context");                                     // no CPU fallback: fails without an sm_100 GPU
    auto t_build = std::chrono::steady_clock::now();
    check(grt_set_metric(ctx, metric, is_static ? 1 : 0), "program build");
    printf("Built %s program for %s in %.0f ms\n", is_static ? "static" : "dynamic", grt_metric_name(metric),
           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build).count());

    std::vector<uint8_t> sky;
    int bw = 4096, bh = 2048;
    const bool is_png = background_path.size() > 4 && background_path.compare(background_path.size() - 4, 4, ".png") == 0;
    if (is_png) {                                                 // backgrounds/nasa.png in the reference (main.cpp:233-234)
        uint8_t* px = grt_image_load_png(background_path.c_str(), &bw, &bh);
        if (!px) die("background");
        sky.assign(px, px + (size_t)bw * bh * 4);
        grt_image_free(px);
    } else if (background_path.empty() || !read_ppm(background_path.c_str(), sky, bw, bh)) sky = procedural_sky(bw, bh);
    check(grt_set_background(ctx, sky.data(), bw, bh, 0), "background");

    grt_features feat;
    check(grt_get_features(ctx, &feat), "features");
    feat.adaptive_sampling = adaptive ? 1 : 0;
    feat.redshift = redshift ? 1 : 0;
    check(grt_set_features(ctx, &feat), "features");
    const float pos[4] = {0.0f, 0.0f, -4.0f, 0.0f};                                   // main.cpp:669
    const float s = sinf(-3.14159265f / 4), c = cosf(-3.14159265f / 4);
    const float quat[4] = {s, 0.0f, 0.0f, c};                                          // axis (1,0,0), angle -pi/2 (main.cpp:670-673)
    check(grt_set_camera(ctx, pos, quat, 0), "camera");
    if (on_geodesic) {                                                                 // main.cpp:2675-2760, 2264-2297
        int n = grt_snapshot_geodesic(ctx, 0.0f, 1);
        check(n, "geodesic snapshot");
        check(grt_set_camera_on_geodesic(ctx, geodesic_tau), "camera on geodesic");
        printf("Camera on its geodesic at proper time %g (%d path points)\n", geodesic_tau, n);
    }
    check(grt_set_frames_in_flight(ctx, inflight), "frames in flight");

    // one synchronous frame with statistics, then `frames` frames queued back to back into device buffers
    std::vector<uint8_t> host((size_t)width * height * 4);
    grt_stats st;
    check(grt_render(ctx, width, height, host.data(), GRT_FORMAT_RGBA8, GRT_OUT_HOST, &st), "render");
    printf("First frame: %.3f ms on the device (prepass %.3f, integrate %.3f, render %.3f), %llu ray-steps, %llu rays (%llu refinement)\n", st.ms_total, st.ms_prepass,
           st.ms_integrate, st.ms_render, (unsigned long long)st.trial_steps, (unsigned long long)st.rays_integrated, (unsigned long long)st.extra_rays);
    std::vector<void*> dev(inflight, nullptr);
    for (auto& d : dev) if (cudaMalloc(&d, host.size()) != cudaSuccess) { fprintf(stderr, "grt_headless: cudaMalloc failed\n"); return 1; }
    // frames go out in groups of `inflight` (one per frame slot, overlapping on the GPU); the host waits per group
    for (int k = 0; k < frames; k += inflight) {
        const int n = std::min(inflight, frames - k);
        auto t0 = std::chrono::steady_clock::now();
        for (int j = 0; j < n; j++) {
            check(grt_set_camera(ctx, pos, quat, 0), "camera");                       // the per-frame camera upload, main.cpp:2244-2262
            check(grt_render(ctx, width, height, dev[j], GRT_FORMAT_RGBA8, GRT_OUT_DEVICE, nullptr), "render");
        }
        check(grt_synchronize(ctx), "synchronize");
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / n;
        if (print_frametime)
            for (int j = 0; j < n; j++) printf("Frametime Elapsed: %f\n", ms);       // main.cpp:2864-2871 (do-not-change string)
    }
    check(grt_synchronize(ctx), "synchronize");
    if (!out_path.empty()) {                                                           // screenshot, main.cpp:2762-2831
        write_ppm(out_path.c_str(), host, width, height);
        printf("Wrote %s (%dx%d)\n", out_path.c_str(), width, height);
    }
    for (auto& d : dev) cudaFree(d);
    grt_destroy(ctx);
    grt_metric_destroy(metric);
    grt_content_destroy(content);
    return 0;
}
