// grt_headless - the host side of the hot path in C++ over the C ABI (include/grt.h), without a window.
//
// Stands where the reference's frame loop stands (/root/reference/main.cpp:1503-2871): content folders
// (main.cpp:1562-1586), `-bench <metric>` / `-start <metric>` (main.cpp:956-986), the default camera
// (main.cpp:664-673), the per-frame sequence (main.cpp:2244-2526) as one grt_render call, the
// "Frametime Elapsed: %f" line of -bench (main.cpp:2864-2871), the screenshot path (main.cpp:2762-2831: clamp, linear -> sRGB,
// clamp, * 255, saved as ./screenshots/<metric>_<milliseconds>.png) and the settings file (main.cpp:240-254: settings.json with the
// fields of graphics_settings.cpp:10-33).  With -gpus N the frame is tiled in row stripes over N devices of this box and assembled
// on device 0 (grt_group_*: one context per device, peer stores into device 0's frame) - the single-context frame loop of
// main.cpp:1069-1095 over several GPUs.  The window, ImGui, input and Steam parts of main.cpp are out of scope.
//
//   grt_headless -content <folder> -bench <metric name> [-size W H] [-frames N] [-inflight K] [-gpus N | -devices 0,1,...] [-no-balance]
//                [-param name value]... [-static] [-no-adaptive] [-redshift] [-geodesic tau] [-background file.png|file.ppm]
//                [-settings settings.json] [-save-settings settings.json] [-screenshot] [-out frame.png|frame.ppm]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <sys/stat.h>

#include <cuda_runtime.h>

#include "grt.h"
#include "mini_json.hpp"

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

// the fields of struct graphics_settings that reach the hot path or the screenshot (graphics_settings.hpp:8-47, same defaults)
struct Settings {
    int width = 1920, height = 1080;
    float field_of_view = 90.0f;
    int screenshot_width = 1920, screenshot_height = 1080;
    int supersample_factor = 2;
    bool supersample = false;
    int anisotropy = 8;
    bool use_old_redshift = false, use_adaptive_sampling = true;
    float adaptive_sampling_threshold = 32.0f;
};

// main.cpp:240-254: a missing file keeps the defaults, a broken one prints the reference's message and keeps the defaults
bool load_settings(const std::string& path, Settings& st) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    try {
        mini_json::Value js = mini_json::parse(text);
        auto num = [&](const char* k, auto& out) { if (const mini_json::Value* v = js.find(k)) out = (std::remove_reference_t<decltype(out)>)v->num; };
        auto flag = [&](const char* k, bool& out) { if (const mini_json::Value* v = js.find(k)) out = v->b || v->num != 0; };
        num("width", st.width); num("height", st.height); num("field_of_view", st.field_of_view);
        num("screenshot_width", st.screenshot_width); num("screenshot_height", st.screenshot_height);
        num("supersample_factor", st.supersample_factor); flag("supersample", st.supersample);
        num("anisotropy", st.anisotropy); flag("use_old_redshift", st.use_old_redshift);
        flag("use_adaptive_sampling", st.use_adaptive_sampling); num("adaptive_sampling_threshold", st.adaptive_sampling_threshold);
        return true;
    } catch (std::exception& ex) {
        printf("Failed to load settings.json %s\n", ex.what());
        return false;
    }
}
void save_settings(const std::string& path, const Settings& st) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "grt_headless: cannot write %s\n", path.c_str()); return; }
    fprintf(f, "{\"width\":%d,\"height\":%d,\"field_of_view\":%g,\"screenshot_width\":%d,\"screenshot_height\":%d,\"supersample\":%s,\"supersample_factor\":%d,"
               "\"anisotropy\":%d,\"use_old_redshift\":%s,\"use_adaptive_sampling\":%s,\"adaptive_sampling_threshold\":%g}\n",
            st.width, st.height, st.field_of_view, st.screenshot_width, st.screenshot_height, st.supersample ? "true" : "false", st.supersample_factor, st.anisotropy,
            st.use_old_redshift ? "true" : "false", st.use_adaptive_sampling ? "true" : "false", st.adaptive_sampling_threshold);
    fclose(f);
}

bool ends_with(const std::string& s, const char* suffix) {
    size_t n = strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
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
    std::string content_dir, metric_name = "kerr_schild", out_path, background_path, settings_path = "settings.json", save_settings_path;
    int width = 0, height = 0, frames = 8, inflight = 2, gpus = 1;
    bool is_static = false, adaptive = true, redshift = false, print_frametime = false, on_geodesic = false, screenshot = false, adaptive_given = false, balance = true;
    float geodesic_tau = 0.0f;
    std::vector<std::pair<std::string, float>> params;
    std::vector<int> devices;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](int k = 1) -> const char* { if (i + k >= argc) { fprintf(stderr, "grt_headless: %s needs a value\n", a.c_str()); exit(2); } return argv[i + k]; };
        if (a == "-content") { content_dir = next(); i++; }
        else if (a == "-bench") { metric_name = next(); print_frametime = true; i++; }       // main.cpp:975-980
        else if (a == "-start") { metric_name = next(); i++; }                               // main.cpp:982-985
        else if (a == "-size") { width = atoi(next(1)); height = atoi(next(2)); i += 2; }
        else if (a == "-frames") { frames = atoi(next()); i++; }
        else if (a == "-inflight") { inflight = atoi(next()); i++; }
        else if (a == "-gpus") { gpus = atoi(next()); i++; }
        else if (a == "-balance") balance = true;                                             // the default: rows shared out by measured cost (grt_group_rebalance)
        else if (a == "-no-balance") balance = false;                                         // plain interleaved stripes
        else if (a == "-devices") {                                                          // explicit device list, e.g. 0,1,2,3 (a device may repeat: ranks then share it)
            devices.clear();
            for (const char* q = next(); *q;) { devices.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q == ',') q++; }
            gpus = (int)devices.size();
            i++;
        }
        else if (a == "-param") { params.push_back({next(1), (float)atof(next(2))}); i += 2; }
        else if (a == "-static") is_static = true;
        else if (a == "-no-adaptive") { adaptive = false; adaptive_given = true; }
        else if (a == "-redshift") redshift = true;
        else if (a == "-geodesic") { on_geodesic = true; geodesic_tau = (float)atof(next()); i++; }
        else if (a == "-background") { background_path = next(); i++; }
        else if (a == "-settings") { settings_path = next(); i++; }
        else if (a == "-save-settings") { save_settings_path = next(); i++; }
        else if (a == "-screenshot") screenshot = true;
        else if (a == "-out") { out_path = next(); i++; }
        else { fprintf(stderr, "grt_headless: unknown argument %s\n", a.c_str()); return 2; }
    }
    if (content_dir.empty()) { fprintf(stderr, "usage: grt_headless -content <folder> -bench <metric name> [-size W H] [-frames N] [-gpus N] [-screenshot] [-out frame.png]\n"); return 2; }
    if (gpus < 1 || gpus > 16) { fprintf(stderr, "grt_headless: -gpus must be 1..16\n"); return 2; }

    printf("Started (%s)\n", grt_version());
    Settings sett;
    if (load_settings(settings_path, sett)) printf("Loaded %s\n", settings_path.c_str());       // main.cpp:240-254
    if (width <= 0 || height <= 0) { width = sett.width; height = sett.height; }
    if (!adaptive_given) adaptive = sett.use_adaptive_sampling;
    if (!save_settings_path.empty()) { sett.width = width; sett.height = height; save_settings(save_settings_path, sett); }

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

    std::vector<uint8_t> sky;
    int bw = 4096, bh = 2048;
    if (ends_with(background_path, ".png")) {                     // backgrounds/nasa.png in the reference (main.cpp:233-234)
        uint8_t* px = grt_image_load_png(background_path.c_str(), &bw, &bh);
        if (!px) die("background");
        sky.assign(px, px + (size_t)bw * bh * 4);
        grt_image_free(px);
    } else if (background_path.empty() || !read_ppm(background_path.c_str(), sky, bw, bh)) sky = procedural_sky(bw, bh);

    const float pos[4] = {0.0f, 0.0f, -4.0f, 0.0f};                                   // main.cpp:669
    const float s = sinf(-3.14159265f / 4), c = cosf(-3.14159265f / 4);
    const float quat[4] = {s, 0.0f, 0.0f, c};                                          // axis (1,0,0), angle -pi/2 (main.cpp:670-673)

    // the devices as one renderer (grt_group_*): one context per device, rank 0 owns the assembled frame
    if (devices.empty()) for (int r = 0; r < gpus; r++) devices.push_back(r);
    grt_group* group = grt_group_create(devices.data(), gpus);
    if (!group) die("context");                                   // no CPU fallback: fails without an sm_100 GPU
    grt_ctx* ctx = grt_group_ctx(group, 0);
    auto t_build = std::chrono::steady_clock::now();
    check(grt_group_set_metric(group, metric, is_static ? 1 : 0), "program build");
    printf("Built %s program for %s in %.0f ms\n", is_static ? "static" : "dynamic", grt_metric_name(metric),
           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build).count());
    check(grt_group_set_background(group, sky.data(), bw, bh, 0), "background");
    grt_features feat;
    check(grt_get_features(ctx, &feat), "features");
    feat.adaptive_sampling = adaptive ? 1 : 0;
    feat.redshift = redshift ? 1 : 0;
    feat.field_of_view = sett.field_of_view;
    feat.max_probes = sett.anisotropy;
    feat.use_old_redshift = sett.use_old_redshift ? 1 : 0;
    feat.adaptive_sampling_threshold = sett.adaptive_sampling_threshold;
    check(grt_group_set_features(group, &feat), "features");
    check(grt_group_set_camera(group, pos, quat, 0), "camera");
    if (on_geodesic) {                                                                 // main.cpp:2675-2760, 2264-2297: every device integrates the same path
        int n = 0;
        for (int r = 0; r < gpus; r++) {
            n = grt_snapshot_geodesic(grt_group_ctx(group, r), 0.0f, 1);
            check(n, "geodesic snapshot");
            check(grt_set_camera_on_geodesic(grt_group_ctx(group, r), geodesic_tau), "camera on geodesic");
        }
        printf("Camera on its geodesic at proper time %g (%d path points)\n", geodesic_tau, n);
    }
    check(grt_group_set_frames_in_flight(group, inflight), "frames in flight");
    if (gpus > 1 && balance) check(grt_group_rebalance(group, width, height), "stripe costs");   // one block of rows per device, boundaries at equal measured cost

    // one synchronous frame with statistics on device 0 (the whole frame), then `frames` frames queued back to back
    std::vector<uint8_t> host((size_t)width * height * 4);
    grt_stats st;
    check(grt_render(ctx, width, height, host.data(), GRT_FORMAT_RGBA8, GRT_OUT_HOST, &st), "render");
    printf("First frame: %.3f ms on the device (prepass %.3f, integrate %.3f, render %.3f), %llu ray-steps, %llu rays (%llu refinement)\n", st.ms_total, st.ms_prepass,
           st.ms_integrate, st.ms_render, (unsigned long long)st.trial_steps, (unsigned long long)st.rays_integrated, (unsigned long long)st.extra_rays);

    // the frame loop: every device renders its stripes straight into device 0's frame buffer; the assembled frame lands in pinned
    // host memory behind the last device's stores
    if (gpus > 1) printf("Frame tiled over %d devices\n", gpus);
    std::vector<uint8_t*> pinned(std::max(2, inflight), nullptr);
    for (auto& p : pinned) if (cudaHostAlloc((void**)&p, host.size(), cudaHostAllocDefault) != cudaSuccess) { fprintf(stderr, "grt_headless: cudaHostAlloc failed\n"); return 1; }
    unsigned seq = 0;
    // frames go out in groups of `inflight` (one per frame slot, overlapping on the GPU); the host waits per group
    for (int k = 0; k < frames; k += inflight) {
        const int n = std::min(inflight, frames - k);
        auto t0 = std::chrono::steady_clock::now();
        for (int j = 0; j < n; j++, seq++) {
            if (!on_geodesic) check(grt_group_set_camera(group, pos, quat, 0), "camera");   // the per-frame camera upload, main.cpp:2244-2262
            check(grt_group_render(group, width, height, GRT_FORMAT_RGBA8, pinned[seq % pinned.size()]), "render");
        }
        check(grt_group_synchronize(group), "synchronize");
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / n;
        if (print_frametime)
            for (int j = 0; j < n; j++) printf("Frametime Elapsed: %f\n", ms);       // main.cpp:2864-2871 (do-not-change string)
    }
    if (seq > 0) {
        const uint8_t* last = pinned[(seq - 1) % pinned.size()];
        if (memcmp(last, host.data(), host.size()) != 0) printf("note: the last frame of the loop differs from the first frame\n");
        memcpy(host.data(), last, host.size());
    }

    // screenshot (main.cpp:2762-2831): screenshot_width x screenshot_height (times the supersample factor when supersampling is on),
    // clamp -> lin_to_srgb -> clamp -> * 255 (the GRT_FORMAT_RGBA8 encoding), ./screenshots/<metric>_<milliseconds>.png
    if (screenshot) {
        const int sw = sett.screenshot_width * (sett.supersample ? sett.supersample_factor : 1), sh = sett.screenshot_height * (sett.supersample ? sett.supersample_factor : 1);
        std::vector<uint8_t> shot((size_t)sw * sh * 4);
        check(grt_render(ctx, sw, sh, shot.data(), GRT_FORMAT_RGBA8, GRT_OUT_HOST, nullptr), "screenshot render");
        mkdir("./screenshots", 0755);
        const long long millis = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
        std::string fname = "./screenshots/" + std::string(grt_metric_name(metric)) + "_" + std::to_string(millis) + ".png";
        check(grt_image_save_png(fname.c_str(), shot.data(), sw, sh), "screenshot");
        printf("Wrote %s (%dx%d)\n", fname.c_str(), sw, sh);
    }
    if (!out_path.empty()) {
        if (ends_with(out_path, ".png")) check(grt_image_save_png(out_path.c_str(), host.data(), width, height), "png");
        else write_ppm(out_path.c_str(), host, width, height);
        printf("Wrote %s (%dx%d)\n", out_path.c_str(), width, height);
    }
    for (auto& p : pinned) cudaFreeHost(p);
    grt_group_destroy(group);
    grt_metric_destroy(metric);
    grt_content_destroy(content);
    return 0;
}
