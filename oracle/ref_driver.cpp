// =============================================================================================
// TEST INFRASTRUCTURE ONLY.  Host driver for oracle/_ref/libclref_<metric>.so: runs the
// reference's OWN kernels - /root/reference/cl.cl compiled as C++ where it lies (ref_build.py,
// ref_shim/) - on host cores, one OpenMP iteration per OpenCL work-item, in the order the
// reference's frame loop enqueues them (/root/reference/main.cpp:2299-2331 camera,
// :2356-2526 frame, :139-205 execute_kernel).  Same C structs as oracle_capi.cpp so the Python
// side shares one binding.  Nothing here is part of the product path.
// =============================================================================================
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// the host-side texture build (graphics_settings.cpp:152-212) has no device source to compile:
// the restatement in geodesic_oracle.hpp supplies the mip chain; it is expanded below into the
// reference's padded full-size array layers, which the reference's own read_mipmap then samples.
#include "geodesic_oracle.hpp"

#include "ref_shim/opencl_as_cxx.hpp"
#undef printf

// trial ray-steps: GEO_ACCEL0 is evaluated exactly once per step_verlet call (cl.cl:3299); the
// counter lives in the cl.cl translation unit (ref_shim/cl_macro_abi.hpp, cl_trailer.inc)
extern "C" long long clref_take_trial_steps();

// kernel prototypes (cl.cl line of each __kernel in the comment); struct bodies live in cl.cl
struct lightray;
struct render_data;
struct dynamic_config;
struct dynamic_feature_config;
extern "C" {
void cart_to_generic_kernel(const float4*, float4*, int, float, const dynamic_config*);                                              // cl.cl:6018
void init_basis_vectors(const float4*, int, float3, float4*, float4*, float4*, float4*, const dynamic_config*);                      // cl.cl:2483
void clear_termination_buffer(int*, int, int);                                                                                       // cl.cl:4997
void calculate_singularities(const lightray*, const int*, int*, int, int);                                                           // cl.cl:5008
void init_rays_generic(const float4*, const float4*, lightray*, int*, int, int, const int*, int, int, int, const float4*, const float4*, const float4*,
                       const float4*, const dynamic_config*, const dynamic_feature_config*, int);                                    // cl.cl:3143
void do_generic_rays(lightray*, const int*, int*, int*, const dynamic_config*, const dynamic_feature_config*, int, int, int, int, float4*, int*, int);  // cl.cl:3954
void calculate_render_data(const lightray*, const int*, render_data*, int*, int, int, const dynamic_config*, const dynamic_feature_config*);            // cl.cl:5135
void handle_adaptive_sampling(const lightray*, const int*, render_data*, int*, lightray*, int*, float4*, float4*, const float4*, const float4*, const float4*,
                              const float4*, int, int, const dynamic_config*, const dynamic_feature_config*);                        // cl.cl:5223
void render(const render_data*, const int*, image2d_t, image2d_array_t, image2d_array_t, int, int, int, const dynamic_config*, const dynamic_feature_config*);  // cl.cl:5453
void boost_tetrad(float4*, int, float3*, float4*, float4*, float4*, float4*, const dynamic_config*);                                   // cl.cl:2441
void init_inertial_ray(float4*, int, lightray*, int*, float4*, float4*, float4*, float4*, float3*, const dynamic_config*);            // cl.cl:3117
void get_geodesic_path(lightray*, float4*, float4*, float*, int*, int, const dynamic_config*, const dynamic_feature_config*, int*);   // cl.cl:4735
void parallel_transport_quantity(float4*, float4*, float*, float4*, int*, int, float4*, const dynamic_config*);                       // cl.cl:2569
void handle_interpolating_geodesic(const float4*, const float4*, const float*, float4*, const float4*, const float4*, const float4*, const float4*, float4*, float4*,
                                   float4*, float4*, float, const int*, int, const float3*, float4*, const dynamic_config*);          // cl.cl:2738
int clref_sizeof_lightray();
int clref_sizeof_render_data();
int clref_sizeof_features();
}

namespace {

// struct dynamic_feature_config of the dynamic program (cl_macro_abi.hpp: floats then ints, alphabetical)
struct FeatureBlock {
    float adaptive_sampling_threshold, field_of_view, max_acceleration_change, max_precision_radius, min_step, ray_skip, universe_size;
    int adaptive_sampling, redshift, reparameterisation, use_old_redshift, use_triangle_rendering;
};

template <typename T> struct Buf {  // 64-byte aligned, zero-filled device-buffer stand-in
    T* p = nullptr;
    explicit Buf(size_t n) { size_t bytes = ((n ? n : 1) * sizeof(T) + 63) / 64 * 64; p = (T*)aligned_alloc(64, bytes); memset((void*)p, 0, bytes); }
    ~Buf() { free((void*)p); }
    Buf(const Buf&) = delete;
};
struct RawBuf {
    char* p;
    RawBuf(size_t n, size_t elem) { size_t bytes = ((n ? n : 1) * elem + 63) / 64 * 64; p = (char*)aligned_alloc(64, bytes); memset(p, 0, bytes); }
    ~RawBuf() { free(p); }
    RawBuf(const RawBuf&) = delete;
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// clEnqueueNDRangeKernel over a 1-D range: global size rounded up to the work-group size, as an
// OpenCL 1.2 launch needs; every kernel bounds-checks its id.
template <typename F> void launch_1d(size_t gsize, size_t wgs, long long* steps, F&& body) {
    size_t n = (gsize + wgs - 1) / wgs * wgs;
    long long total = 0;
#pragma omp parallel reduction(+ : total)
    {
#pragma omp for schedule(dynamic, 16)
        for (long long id = 0; id < (long long)n; id++) {
            clshim::work_item.gid[0] = (size_t)id; clshim::work_item.gid[1] = 0;
            clshim::work_item.lid[0] = (size_t)id % wgs; clshim::work_item.grp[0] = (size_t)id / wgs;
            clshim::work_item.gsz[0] = n; clshim::work_item.lsz[0] = wgs;
            body();
        }
        total += clref_take_trial_steps();
    }
    if (steps) *steps += total;
}
// a chosen subset of the work-items of a 1-D launch of global size gsize (id_of maps 0..count-1 to global ids)
template <typename M, typename F> void launch_1d_ids(size_t count, size_t gsize, size_t wgs, M&& id_of, F&& body) {
    size_t n = (gsize + wgs - 1) / wgs * wgs;
#pragma omp parallel for schedule(dynamic, 16)
    for (long long k = 0; k < (long long)count; k++) {
        size_t id = id_of((size_t)k);
        clshim::work_item.gid[0] = id; clshim::work_item.gid[1] = 0;
        clshim::work_item.lid[0] = id % wgs; clshim::work_item.grp[0] = id / wgs;
        clshim::work_item.gsz[0] = n; clshim::work_item.lsz[0] = wgs;
        body();
    }
}
// 2-D launch; only the work-items with x0 <= gid0 < x1, y0 <= gid1 < y1 are run (the whole range by default)
template <typename F> void launch_2d(size_t gx, size_t gy, size_t lx, size_t ly, size_t x0, size_t y0, size_t x1, size_t y1, F&& body) {
    size_t nx = (gx + lx - 1) / lx * lx, ny = (gy + ly - 1) / ly * ly;
    x1 = std::min(x1, nx); y1 = std::min(y1, ny);
#pragma omp parallel for schedule(dynamic, 1)
    for (long long y = (long long)y0; y < (long long)y1; y++)
        for (size_t x = x0; x < x1; x++) {
            clshim::work_item.gid[0] = x; clshim::work_item.gid[1] = (size_t)y;
            clshim::work_item.lid[0] = x % lx; clshim::work_item.lid[1] = (size_t)y % ly;
            clshim::work_item.grp[0] = x / lx; clshim::work_item.grp[1] = (size_t)y / ly;
            clshim::work_item.gsz[0] = nx; clshim::work_item.gsz[1] = ny; clshim::work_item.lsz[0] = lx; clshim::work_item.lsz[1] = ly;
            body();
        }
}

// the padded array image of load_mipped_image (graphics_settings.cpp:164-211)
std::vector<uint8_t> padded_layers(const oracle::MipImage& m) {
    std::vector<uint8_t> out((size_t)m.levels * m.width * m.height * 4);
    for (int l = 0; l < m.levels; l++) {
        const uint8_t* src = m.level[l].data();
        uint8_t* dst = out.data() + (size_t)l * m.width * m.height * 4;
        for (int y = 0; y < m.height; y++)
            for (int x = 0; x < m.width; x++) {
                int lx = std::min(x, m.lw[l] - 1), ly = std::min(y, m.lh[l] - 1);
                memcpy(dst + ((size_t)y * m.width + x) * 4, src + ((size_t)ly * m.lw[l] + lx) * 4, 4);
            }
    }
    return out;
}

}  // namespace

extern "C" {

// identical to oracle_capi.cpp
struct OracleFeatures {
    float universe_size, max_acceleration_change, max_precision_radius, min_step, adaptive_sampling_threshold, field_of_view;
    int redshift, reparameterisation, use_old_redshift, adaptive_sampling;
};
struct OracleFrameArgs {
    int width, height;
    float camera_pos[4];
    float camera_quat[4];
    float basis_speed[3];
    int flip;
    int use_given_camera;
    float camera_generic[4];
    float tetrad[16];
    OracleFeatures features;
    float params[256];   // struct grt_params: named + derived values (grt_metric_packed_params)
    int max_probes;
    int use_prepass;
    const uint8_t* bg1; int bg1_w, bg1_h;
    const uint8_t* bg2; int bg2_w, bg2_h;
};
struct OracleFrameOut {
    float* image;
    void* rays;
    void* rdata;
    int* termination;
    long long trial_steps, accepted_steps;
    long long rays_traced, extra_rays;
    double t_init_s, t_integrate_s, t_render_data_s, t_adaptive_s, t_render_s, t_total_s;
    float camera_generic[4];
    float tetrad[16];
};

static FeatureBlock to_block(const OracleFeatures& f) {
    FeatureBlock b;
    b.adaptive_sampling_threshold = f.adaptive_sampling_threshold; b.field_of_view = f.field_of_view;
    b.max_acceleration_change = f.max_acceleration_change; b.max_precision_radius = f.max_precision_radius;
    b.min_step = f.min_step; b.ray_skip = 4.f; b.universe_size = f.universe_size;
    b.adaptive_sampling = f.adaptive_sampling; b.redshift = f.redshift; b.reparameterisation = f.reparameterisation;
    b.use_old_redshift = f.use_old_redshift; b.use_triangle_rendering = 0;
    return b;
}

void clref_info(int* out) {
    out[0] = GRT_BIG_METRIC; out[1] = GRT_CONSTANT_THETA; out[2] = GRT_SINGULAR; out[3] = GRT_ADAPTIVE_PRECISION;
    out[4] = GRT_SINGULARITY_DETECTION; out[5] = GRT_USE_PREPASS; out[6] = GRT_NUM_PARAMS; out[7] = GRT_STATIC_PARAMS;
    out[8] = clref_sizeof_lightray(); out[9] = clref_sizeof_render_data();
#ifdef _OPENMP
    out[10] = omp_get_max_threads();
#else
    out[10] = 1;
#endif
    out[11] = clref_sizeof_features() == (int)sizeof(FeatureBlock);
}

// bench.py's reference arm runs under torchrun, which exports OMP_NUM_THREADS=1: the thread count is set explicitly
void clref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

struct CameraState { Buf<float4> generic{1}, quat{1}, e0{1}, e1{1}, e2{1}, e3{1}; };

static void setup_camera(const OracleFrameArgs* a, const dynamic_config* cfg, CameraState& c) {
    *c.quat.p = float4(a->camera_quat[0], a->camera_quat[1], a->camera_quat[2], a->camera_quat[3]);
    if (a->use_given_camera) {
        *c.generic.p = float4(a->camera_generic[0], a->camera_generic[1], a->camera_generic[2], a->camera_generic[3]);
        float4* e[4] = {c.e0.p, c.e1.p, c.e2.p, c.e3.p};
        for (int i = 0; i < 4; i++) *e[i] = float4(a->tetrad[i * 4], a->tetrad[i * 4 + 1], a->tetrad[i * 4 + 2], a->tetrad[i * 4 + 3]);
        return;
    }
    // main.cpp:2299-2331
    Buf<float4> cart(1);
    *cart.p = float4(a->camera_pos[0], a->camera_pos[1], a->camera_pos[2], a->camera_pos[3]);
    float clflip = a->flip ? 1.f : 0.f;
    launch_1d(1, 1, nullptr, [&] { cart_to_generic_kernel(cart.p, c.generic.p, 1, clflip, cfg); });
    float3 speed(a->basis_speed[0], a->basis_speed[1], a->basis_speed[2]);
    launch_1d(1, 1, nullptr, [&] { init_basis_vectors(c.generic.p, 1, speed, c.e0.p, c.e1.p, c.e2.p, c.e3.p, cfg); });
}

// One frame, main.cpp:2356-2526 (triangle rendering off).
// `win` = {x0, y0, x1, y1}: a window of the frame (see oracle_capi.cpp frame_impl) - the same kernels, but only the rays the window
// depends on go through do_generic_rays, handle_adaptive_sampling and render are enqueued over the window's sub-range of work-items,
// and o->image / o->rdata receive the window only.  This is how frames of the BASELINE sizes are produced in seconds.
static int frame_impl(const OracleFrameArgs* a, OracleFrameOut* o, const int* win) {
    const int width = a->width, height = a->height;
    int ex0 = 0, ey0 = 0, ex1 = width, ey1 = height;
    if (win) {
        ex0 = std::max(0, (win[0] - 6) & ~1); ey0 = std::max(0, (win[1] - 6) & ~1);
        ex1 = std::min(width, (win[2] + 7) & ~1); ey1 = std::min(height, (win[3] + 7) & ~1);
    }
    const size_t LR = (size_t)clref_sizeof_lightray(), RD = (size_t)clref_sizeof_render_data();
    double t_start = now_s();
    Buf<float> cfg_buf(256);
    memcpy(cfg_buf.p, a->params, sizeof(float) * 256);
    const dynamic_config* cfg = (const dynamic_config*)cfg_buf.p;
    Buf<FeatureBlock> dfg_buf(1);
    *dfg_buf.p = to_block(a->features);
    const dynamic_feature_config* dfg = (const dynamic_feature_config*)dfg_buf.p;

    CameraState cam;
    setup_camera(a, cfg, cam);
    memcpy(o->camera_generic, cam.generic.p, 16);
    memcpy(o->tetrad + 0, cam.e0.p, 16); memcpy(o->tetrad + 4, cam.e1.p, 16); memcpy(o->tetrad + 8, cam.e2.p, 16); memcpy(o->tetrad + 12, cam.e3.p, 16);

    const int prepass_width = width / 16, prepass_height = height / 16;   // main.cpp:2378-2379
    const size_t npx = (size_t)width * height;
    RawBuf rays_in(npx, LR), rays_adaptive(npx, LR), rdata(npx, RD);
    Buf<int> rays_count_in(1), rays_adaptive_count(1), render_data_count(1), termination_buffer(npx), stored_ray_counts(npx), tmin(npx), tmax(npx);
    const int isnap = 0;
    long long steps = 0;
    o->rays_traced = 0; o->t_integrate_s = 0; o->extra_rays = 0; o->t_adaptive_s = 0; o->t_render_s = 0;
    const size_t wg = 64;   // workgroup_size 8x8 (graphics_settings.hpp:42)

    long long last_traced = 0;
    auto execute_kernel = [&](char* rays, int* count, bool windowed) {   // main.cpp:139-205; global size {width*height}
        memset(stored_ray_counts.p, 0, sizeof(int) * npx);
        double t = now_s();
        if (!windowed) {
            launch_1d(npx, wg, &steps, [&] { do_generic_rays((lightray*)rays, count, tmin.p, tmax.p, cfg, dfg, width, height, 0, 0, nullptr, stored_ray_counts.p, 0); });
            last_traced = *count;
        } else {
            // gather the rays of the extended window (lightray::sx, sy at bytes 76, 80: cl.cl:813-824), run the kernel on them, scatter back
            std::vector<int> idx;
            for (int i = 0; i < *count; i++) {
                int sx = *(const int*)(rays + (size_t)i * LR + 76), sy = *(const int*)(rays + (size_t)i * LR + 80);
                if (sx >= ex0 && sx < ex1 && sy >= ey0 && sy < ey1) idx.push_back(i);
            }
            RawBuf sub(idx.size(), LR);
            Buf<int> sub_count(1);
            for (size_t i = 0; i < idx.size(); i++) memcpy(sub.p + i * LR, rays + (size_t)idx[i] * LR, LR);
            *sub_count.p = (int)idx.size();
            launch_1d(idx.size(), wg, &steps, [&] { do_generic_rays((lightray*)sub.p, sub_count.p, tmin.p, tmax.p, cfg, dfg, width, height, 0, 0, nullptr, stored_ray_counts.p, 0); });
            for (size_t i = 0; i < idx.size(); i++) memcpy(rays + (size_t)idx[i] * LR, sub.p + i * LR, LR);
            last_traced = (long long)idx.size();
        }
        o->t_integrate_s += now_s() - t;
        o->rays_traced += last_traced;
    };

    double t0 = now_s();
    int use_prepass = a->use_prepass < 0 ? GRT_USE_PREPASS : a->use_prepass;
    if (use_prepass && prepass_width > 0 && prepass_height > 0) {   // main.cpp:2387-2436
        size_t np = (size_t)prepass_width * prepass_height;
        launch_1d(np, 256, nullptr, [&] { clear_termination_buffer(termination_buffer.p, prepass_width, prepass_height); });
        *rays_count_in.p = 0;
        launch_1d(np, 256, nullptr, [&] {
            init_rays_generic(cam.generic.p, cam.quat.p, (lightray*)rays_in.p, rays_count_in.p, prepass_width, prepass_height, termination_buffer.p, prepass_width,
                              prepass_height, isnap, cam.e0.p, cam.e1.p, cam.e2.p, cam.e3.p, cfg, dfg, 1);
        });
        execute_kernel(rays_in.p, rays_count_in.p, false);
        launch_1d(np, 256, nullptr, [&] { calculate_singularities((const lightray*)rays_in.p, rays_count_in.p, termination_buffer.p, prepass_width, prepass_height); });
    }
    if (o->termination) memcpy(o->termination, termination_buffer.p, sizeof(int) * (size_t)prepass_width * prepass_height);

    launch_1d(npx, 256, nullptr, [&] {   // main.cpp:2438-2461
        init_rays_generic(cam.generic.p, cam.quat.p, (lightray*)rays_in.p, rays_count_in.p, width, height, termination_buffer.p, prepass_width, prepass_height, isnap,
                          cam.e0.p, cam.e1.p, cam.e2.p, cam.e3.p, cfg, dfg, 0);
    });
    o->t_init_s = now_s() - t0 - o->t_integrate_s;
    execute_kernel(rays_in.p, rays_count_in.p, win != nullptr);

    *render_data_count.p = 0;
    double t1 = now_s();
    launch_1d(npx, 256, nullptr, [&] { calculate_render_data((const lightray*)rays_in.p, rays_count_in.p, (render_data*)rdata.p, render_data_count.p, width, height, cfg, dfg); });
    o->t_render_data_s = now_s() - t1;

    if (a->features.adaptive_sampling) {   // main.cpp:2478-2511
        *rays_adaptive_count.p = 0;
        double t2 = now_s();
        launch_2d((size_t)width / 2, (size_t)height / 2, 8, 8, (size_t)ex0 / 2, (size_t)ey0 / 2, (size_t)ex1 / 2, (size_t)ey1 / 2, [&] {
            handle_adaptive_sampling((const lightray*)rays_in.p, rays_count_in.p, (render_data*)rdata.p, render_data_count.p, (lightray*)rays_adaptive.p,
                                     rays_adaptive_count.p, cam.generic.p, cam.quat.p, cam.e0.p, cam.e1.p, cam.e2.p, cam.e3.p, width, height, cfg, dfg);
        });
        o->t_adaptive_s = now_s() - t2;
        execute_kernel(rays_adaptive.p, rays_adaptive_count.p, false);   // every appended ray comes from a cell of the window
        double t4 = now_s();
        launch_1d(npx, 256, nullptr, [&] {
            calculate_render_data((const lightray*)rays_adaptive.p, rays_adaptive_count.p, (render_data*)rdata.p, render_data_count.p, width, height, cfg, dfg);
        });
        o->t_render_data_s += now_s() - t4;
        o->extra_rays = *rays_adaptive_count.p;
    }
    if (o->rays && !win) memcpy(o->rays, rays_in.p, LR * (size_t)*rays_count_in.p);
    if (o->rdata && !win) memcpy(o->rdata, rdata.p, RD * npx);
    const int wx0 = win ? win[0] : 0, wy0 = win ? win[1] : 0, wx1 = win ? win[2] : width, wy1 = win ? win[3] : height, ww = wx1 - wx0, wh = wy1 - wy0;
    if (o->rdata && win)
        for (int y = wy0; y < wy1; y++) memcpy((char*)o->rdata + (size_t)(y - wy0) * ww * RD, rdata.p + ((size_t)y * width + wx0) * RD, RD * (size_t)ww);

    if (o->image && a->bg1) {   // main.cpp:2513-2525
        oracle::MipImage m1 = oracle::build_mip_image(a->bg1, a->bg1_w, a->bg1_h);
        std::vector<uint8_t> l1 = padded_layers(m1), l2;
        image2d_array_t i1{l1.data(), m1.width, m1.height, m1.levels}, i2 = i1;
        if (a->bg2) {
            oracle::MipImage m2 = oracle::build_mip_image(a->bg2, a->bg2_w, a->bg2_h);
            l2 = padded_layers(m2);
            i2 = image2d_array_t{l2.data(), m2.width, m2.height, m2.levels};
        }
        double t6 = now_s();
        if (!win) {
            image2d_t out{o->image, width, height};
            launch_1d(npx, 256, nullptr, [&] { render((const render_data*)rdata.p, render_data_count.p, out, i1, i2, width, height, a->max_probes, cfg, dfg); });
        } else {
            // render writes pixel (sx, sy) of its record: give it the full-size target, enqueue only the window's work-items, copy the window out
            std::vector<float> full((size_t)npx * 4, 0.f);
            image2d_t out{full.data(), width, height};
            launch_1d_ids((size_t)ww * wh, npx, 256, [&](size_t k) { return (size_t)(wy0 + k / ww) * width + wx0 + k % ww; },
                          [&] { render((const render_data*)rdata.p, render_data_count.p, out, i1, i2, width, height, a->max_probes, cfg, dfg); });
            for (int y = wy0; y < wy1; y++) memcpy(o->image + (size_t)(y - wy0) * ww * 4, full.data() + ((size_t)y * width + wx0) * 4, sizeof(float) * 4 * (size_t)ww);
        }
        o->t_render_s = now_s() - t6;
    }
    o->trial_steps = steps;
    o->accepted_steps = -1;   // the reference does not count them
    o->t_total_s = now_s() - t_start;
    return 0;
}

int clref_frame(const OracleFrameArgs* a, OracleFrameOut* o) { return frame_impl(a, o, nullptr); }
int clref_frame_window(const OracleFrameArgs* a, int x0, int y0, int x1, int y1, OracleFrameOut* o) {
    if (x0 < 0 || y0 < 0 || x1 > a->width || y1 > a->height || x1 <= x0 || y1 <= y0) return -1;
    const int win[4] = {x0, y0, x1, y1};
    return frame_impl(a, o, win);
}

// Bounded CPU-baseline sample for bench.py: rays of the pixels (x0 + i*stride, y0 + j*stride) of the
// full width x height grid, produced by the reference's init_rays_generic (full grid, adaptive
// sampling and prepass off so that every pixel's ray exists), gathered, then integrated by the
// reference's do_generic_rays; `seconds` covers the integration launch only.
int clref_integrate_sample(const OracleFrameArgs* a, int x0, int y0, int cw, int ch, int stride, long long* trial_steps, long long* accepted_steps, double* seconds,
                           int* terminated_out, int* steps_out) {
    const int width = a->width, height = a->height;
    const size_t LR = (size_t)clref_sizeof_lightray();
    Buf<float> cfg_buf(256);
    memcpy(cfg_buf.p, a->params, sizeof(float) * 256);
    const dynamic_config* cfg = (const dynamic_config*)cfg_buf.p;
    Buf<FeatureBlock> dfg_buf(1);
    *dfg_buf.p = to_block(a->features);
    dfg_buf.p->adaptive_sampling = 0;
    const dynamic_feature_config* dfg = (const dynamic_feature_config*)dfg_buf.p;
    CameraState cam;
    setup_camera(a, cfg, cam);
    const size_t npx = (size_t)width * height, n = (size_t)cw * ch;
    RawBuf all(npx, LR), sample(n, LR);
    Buf<int> count(1), termination_buffer(npx), stored(n), tmin(n), tmax(n);
    launch_1d(npx, 256, nullptr, [&] {
        init_rays_generic(cam.generic.p, cam.quat.p, (lightray*)all.p, count.p, width, height, termination_buffer.p, width / 16, height / 16, 0, cam.e0.p, cam.e1.p,
                          cam.e2.p, cam.e3.p, cfg, dfg, 0);
    });
    for (size_t i = 0; i < n; i++) {
        int cx = x0 + (int)(i % cw) * stride, cy = y0 + (int)(i / cw) * stride;
        memcpy(sample.p + i * LR, all.p + ((size_t)cy * width + cx) * LR, LR);
    }
    *count.p = (int)n;
    long long steps = 0;
    double t = now_s();
    launch_1d(n, 64, &steps, [&] { do_generic_rays((lightray*)sample.p, count.p, tmin.p, tmax.p, cfg, dfg, cw, ch, 0, 0, nullptr, stored.p, 0); });
    *seconds = now_s() - t;
    *trial_steps = steps;
    if (accepted_steps) *accepted_steps = -1;
    if (terminated_out)
        for (size_t i = 0; i < n; i++) terminated_out[i] = *(const int*)(sample.p + i * LR + 72);   // lightray::terminated (cl.cl:813-824)
    if (steps_out) memset(steps_out, 0, sizeof(int) * n);
    return 0;
}

// main.cpp:2675-2760: boost_tetrad -> init_inertial_ray -> get_geodesic_path -> 4x parallel_transport_quantity
int clref_geodesic_snapshot(const float* camera_generic, const float* tetrad_in, const float* basis_speed, const OracleFeatures* f, const float* params,
                            int max_path_length, float* positions, float* velocities, float* ds, float* transported /* 4 * max * 4 */, float* boosted_tetrad) {
    Buf<float> cfg_buf(256);
    memcpy(cfg_buf.p, params, sizeof(float) * 256);
    const dynamic_config* cfg = (const dynamic_config*)cfg_buf.p;
    Buf<FeatureBlock> dfg_buf(1);
    *dfg_buf.p = to_block(*f);
    const dynamic_feature_config* dfg = (const dynamic_feature_config*)dfg_buf.p;
    Buf<float4> cam(1), e0(1), e1(1), e2(1), e3(1);
    Buf<float3> speed(1);
    *cam.p = float4(camera_generic[0], camera_generic[1], camera_generic[2], camera_generic[3]);
    float4* e[4] = {e0.p, e1.p, e2.p, e3.p};
    for (int i = 0; i < 4; i++) *e[i] = float4(tetrad_in[i * 4], tetrad_in[i * 4 + 1], tetrad_in[i * 4 + 2], tetrad_in[i * 4 + 3]);
    *speed.p = float3(basis_speed[0], basis_speed[1], basis_speed[2]);
    launch_1d(1, 1, nullptr, [&] { boost_tetrad(cam.p, 1, speed.p, e0.p, e1.p, e2.p, e3.p, cfg); });
    for (int i = 0; i < 4; i++) memcpy(boosted_tetrad + i * 4, e[i], 16);
    RawBuf ray(1, (size_t)clref_sizeof_lightray());
    Buf<int> ray_count(1), count_out(1);
    launch_1d(1, 1, nullptr, [&] { init_inertial_ray(cam.p, 1, (lightray*)ray.p, ray_count.p, e0.p, e1.p, e2.p, e3.p, speed.p, cfg); });
    Buf<float4> pos((size_t)max_path_length), vel((size_t)max_path_length), q((size_t)max_path_length);
    Buf<float> dsv((size_t)max_path_length);
    launch_1d(1, 1, nullptr, [&] { get_geodesic_path((lightray*)ray.p, pos.p, vel.p, dsv.p, ray_count.p, max_path_length, cfg, dfg, count_out.p); });
    int cnt = *count_out.p;
    memcpy(positions, pos.p, sizeof(float) * 4 * (size_t)cnt);
    memcpy(velocities, vel.p, sizeof(float) * 4 * (size_t)cnt);
    memcpy(ds, dsv.p, sizeof(float) * (size_t)cnt);
    for (int k = 0; k < 4; k++) {
        launch_1d(1, 1, nullptr, [&] { parallel_transport_quantity(pos.p, vel.p, dsv.p, e[k], count_out.p, 1, q.p, cfg); });
        memcpy(transported + (size_t)k * max_path_length * 4, q.p, sizeof(float) * 4 * (size_t)cnt);
    }
    return cnt;
}

// main.cpp:2264-2296
void clref_geodesic_interpolate(const float* positions, const float* velocities, const float* ds, const float* transported, int max_path_length, int cnt,
                                float target_time, int parallel_transport_observer, const float* basis_speed, const float* params, float* camera_generic_out,
                                float* tetrad_out, float* velocity_out) {
    Buf<float> cfg_buf(256);
    memcpy(cfg_buf.p, params, sizeof(float) * 256);
    const dynamic_config* cfg = (const dynamic_config*)cfg_buf.p;
    size_t n = (size_t)std::max(cnt, 1);
    Buf<float4> pos(n), vel(n), t0(n), t1(n), t2(n), t3(n), cam(1), e0(1), e1(1), e2(1), e3(1), v(1);
    Buf<float> dsv(n);
    Buf<int> count(1);
    Buf<float3> speed(1);
    *count.p = cnt;
    *speed.p = float3(basis_speed[0], basis_speed[1], basis_speed[2]);
    float4* te[4] = {t0.p, t1.p, t2.p, t3.p};
    memcpy(pos.p, positions, sizeof(float) * 4 * (size_t)cnt);
    memcpy(vel.p, velocities, sizeof(float) * 4 * (size_t)cnt);
    memcpy(dsv.p, ds, sizeof(float) * (size_t)cnt);
    for (int k = 0; k < 4; k++) memcpy(te[k], transported + (size_t)k * max_path_length * 4, sizeof(float) * 4 * (size_t)cnt);
    *e0.p = float4(1, 0, 0, 0); *e1.p = float4(0, 1, 0, 0); *e2.p = float4(0, 0, 1, 0); *e3.p = float4(0, 0, 0, 1);
    launch_1d(1, 1, nullptr, [&] {
        handle_interpolating_geodesic(pos.p, vel.p, dsv.p, cam.p, t0.p, t1.p, t2.p, t3.p, e0.p, e1.p, e2.p, e3.p, target_time, count.p, parallel_transport_observer,
                                      speed.p, v.p, cfg);
    });
    memcpy(camera_generic_out, cam.p, 16);
    memcpy(velocity_out, v.p, 16);
    memcpy(tetrad_out + 0, e0.p, 16); memcpy(tetrad_out + 4, e1.p, 16); memcpy(tetrad_out + 8, e2.p, 16); memcpy(tetrad_out + 12, e3.p, 16);
}

}  // extern "C"
