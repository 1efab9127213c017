// =============================================================================================
// CPU ORACLE - TEST INFRASTRUCTURE ONLY (see geodesic_oracle.hpp).  C entry points for ctypes.
// Frame orchestration follows /root/reference/main.cpp:2244-2526 (per-frame kernel sequence) and
// main.cpp:2675-2760 (geodesic snapshot).  Pinned against the reference's own kernels - see geodesic_oracle.hpp.
// =============================================================================================
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "geodesic_oracle.hpp"

using namespace oracle;

extern "C" {

struct OracleFeatures {
    float universe_size, max_acceleration_change, max_precision_radius, min_step, adaptive_sampling_threshold, field_of_view;
    int redshift, reparameterisation, use_old_redshift, adaptive_sampling;
};

struct OracleFrameArgs {
    int width, height;
    float camera_pos[4];    // (t, x, y, z) cartesian, main.cpp:669-673
    float camera_quat[4];   // (x, y, z, w)
    float basis_speed[3];
    int flip;
    int use_given_camera;   // 1: camera_generic + tetrad below are used as is (camera on a geodesic)
    float camera_generic[4];
    float tetrad[16];
    OracleFeatures features;
    float params[256];   // struct grt_params: named + derived values (grt_metric_packed_params)
    int max_probes;
    int use_prepass;        // -1: metric default (GRT_USE_PREPASS)
    const uint8_t* bg1; int bg1_w, bg1_h;
    const uint8_t* bg2; int bg2_w, bg2_h;   // null: same as bg1
};

struct OracleFrameOut {
    float* image;            // w*h*4 linear float RGBA (may be null)
    lightray* rays;          // w*h first-pass rays after integration (w*h/4 used with adaptive sampling); may be null
    render_data* rdata;      // w*h; may be null
    int* termination;        // (w/16)*(h/16) prepass mask; may be null
    long long trial_steps, accepted_steps;
    long long rays_traced, extra_rays;
    double t_init_s, t_integrate_s, t_render_data_s, t_adaptive_s, t_render_s, t_total_s;
    float camera_generic[4];
    float tetrad[16];
};

static Features to_features(const OracleFeatures& f) {
    Features d;
    d.universe_size = f.universe_size;
    d.max_acceleration_change = f.max_acceleration_change;
    d.max_precision_radius = f.max_precision_radius;
    d.min_step = f.min_step;
    d.adaptive_sampling_threshold = f.adaptive_sampling_threshold;
    d.field_of_view = f.field_of_view;
    d.redshift = f.redshift;
    d.reparameterisation = f.reparameterisation;
    d.use_old_redshift = f.use_old_redshift;
    d.adaptive_sampling = f.adaptive_sampling;
    return d;
}
static grt_params to_params(const float* p) {
    grt_params cfg;
    memset(&cfg, 0, sizeof cfg);
    memcpy(&cfg, p, sizeof(float) * (GRT_NUM_PARAMS > 0 ? GRT_NUM_PARAMS : 1));
    return cfg;
}
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void oracle_info(int* out) {
    out[0] = GRT_BIG_METRIC; out[1] = GRT_CONSTANT_THETA; out[2] = GRT_SINGULAR; out[3] = GRT_ADAPTIVE_PRECISION;
    out[4] = GRT_SINGULARITY_DETECTION; out[5] = GRT_USE_PREPASS; out[6] = GRT_NUM_PARAMS; out[7] = GRT_STATIC_PARAMS;
    out[8] = (int)sizeof(lightray); out[9] = (int)sizeof(render_data);
#ifdef _OPENMP
    out[10] = omp_get_max_threads();
#else
    out[10] = 1;
#endif
}

// bench.py's reference arm runs under torchrun, which exports OMP_NUM_THREADS=1: the thread count is set explicitly
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void oracle_metric_eval(const float* pos, const float* params, float* g16, float* dg64) {
    grt_params cfg = to_params(params);
    calculate_metric_big(from_arr(pos), g16, cfg);
    calculate_partial_derivatives_big(from_arr(pos), dg64, cfg);
}

// a_ref: GEO_ACCEL (symbolic Christoffel), a_num: calculate_acceleration[_big] (numeric Christoffel),
// a_low: the lowered-index formulation the CUDA kernels use
void oracle_accel(const float* pos, const float* vel, const float* params, float* a_ref, float* a_num, float* a_low) {
    grt_params cfg = to_params(params);
    grt_accel_ref(pos, vel, cfg, a_ref);
    grt_accel(pos, vel, cfg, a_low);
    float g[16], dg[64];
    calculate_metric_big(from_arr(pos), g, cfg);
    calculate_partial_derivatives_big(from_arr(pos), dg, cfg);
    to_arr(calculate_acceleration_numeric(from_arr(vel), g, dg), a_num);
}

void oracle_coords(const float* pos, const float* vel, const float* params, float* to_polar, float* to_polar_vel, float* from_polar, float* from_polar_vel,
                   float* distance, float* period) {
    grt_params cfg = to_params(params);
    to_arr(generic_to_spherical(from_arr(pos), cfg), to_polar);
    to_arr(generic_velocity_to_spherical_velocity(from_arr(pos), from_arr(vel), cfg), to_polar_vel);
    to_arr(spherical_to_generic(from_arr(pos), cfg), from_polar);
    to_arr(spherical_velocity_to_generic_velocity(from_arr(pos), from_arr(vel), cfg), from_polar_vel);
    *distance = get_distance_to_object(from_arr(pos), cfg);
    to_arr(get_coordinate_period(cfg), period);
}

// cart_to_generic_kernel + init_basis_vectors (main.cpp:2299-2331)
void oracle_camera(const float* cam_pos, int flip, const float* basis_speed, const float* params, float* generic_out, float* tetrad_out) {
    grt_params cfg = to_params(params);
    float4 gen = cart_to_generic(from_arr(cam_pos), flip ? 1.f : 0.f, cfg);
    float4 e[4];
    calculate_tetrads(gen, f3(basis_speed[0], basis_speed[1], basis_speed[2]), e, cfg, 1);
    to_arr(gen, generic_out);
    for (int i = 0; i < 4; i++) to_arr(e[i], tetrad_out + i * 4);
}

// tetrad at an arbitrary generic position (orient 0/1) - used to check calculate_tetrads on its own
void oracle_tetrad_at(const float* generic_pos, const float* basis_speed, int orient, const float* params, float* tetrad_out) {
    grt_params cfg = to_params(params);
    float4 e[4];
    calculate_tetrads(from_arr(generic_pos), f3(basis_speed[0], basis_speed[1], basis_speed[2]), e, cfg, orient);
    for (int i = 0; i < 4; i++) to_arr(e[i], tetrad_out + i * 4);
}

void oracle_mip_level(const uint8_t* rgba, int w, int h, int level, uint8_t* out, int* lw, int* lh, int* levels) {
    MipImage img = build_mip_image(rgba, w, h);
    *levels = img.levels;
    if (level < img.levels) {
        *lw = img.lw[level]; *lh = img.lh[level];
        if (out) memcpy(out, img.level[level].data(), img.level[level].size());
    }
}

// `win` = {x0, y0, x1, y1}: only the rays a window of the frame depends on are integrated (the window grown by 6 pixels: Jacobian
// +1, interpolation +2, refine stencil +-1 cell), the 1/16 prepass runs in full, and o->image / o->rdata receive the window only.
// Inside the window every value equals that of the full frame - this is how frames of the BASELINE sizes are compared.
static int frame_impl(const OracleFrameArgs* a, OracleFrameOut* o, const int* win) {
    const int width = a->width, height = a->height;
    int ex0 = 0, ey0 = 0, ex1 = width, ey1 = height;
    if (win) {
        ex0 = std::max(0, (win[0] - 6) & ~1); ey0 = std::max(0, (win[1] - 6) & ~1);
        ex1 = std::min(width, (win[2] + 7) & ~1); ey1 = std::min(height, (win[3] + 7) & ~1);
    }
    auto in_ext = [&](const lightray& r) { return r.sx >= ex0 && r.sx < ex1 && r.sy >= ey0 && r.sy < ey1; };
    Features dfg = to_features(a->features);
    grt_params cfg = to_params(a->params);
    double t_start = now_s();
    // integrate the rays of `all` that lie in the extended window (all of them for a full frame)
    auto integrate = [&](lightray* all, int count, Counters* cnt) -> int {
        if (!win) { do_generic_rays(all, count, cfg, dfg, cnt); return count; }
        std::vector<int> idx;
        for (int i = 0; i < count; i++) if (in_ext(all[i])) idx.push_back(i);
        std::vector<lightray> sub(idx.size());
        for (size_t i = 0; i < idx.size(); i++) sub[i] = all[idx[i]];
        do_generic_rays(sub.data(), (int)sub.size(), cfg, dfg, cnt);
        for (size_t i = 0; i < idx.size(); i++) all[idx[i]] = sub[i];
        return (int)sub.size();
    };

    // camera (main.cpp:2244-2331)
    float4 camera_quat = from_arr(a->camera_quat);
    float4 camera_generic;
    float4 e[4];
    if (a->use_given_camera) {
        camera_generic = from_arr(a->camera_generic);
        for (int i = 0; i < 4; i++) e[i] = from_arr(a->tetrad + i * 4);
    } else {
        camera_generic = cart_to_generic(from_arr(a->camera_pos), a->flip ? 1.f : 0.f, cfg);
        calculate_tetrads(camera_generic, f3(a->basis_speed[0], a->basis_speed[1], a->basis_speed[2]), e, cfg, 1);
    }
    to_arr(camera_generic, o->camera_generic);
    for (int i = 0; i < 4; i++) to_arr(e[i], o->tetrad + i * 4);

    Counters cnt;
    const int prepass_width = width / 16, prepass_height = height / 16;
    std::vector<int> termination_buffer((size_t)std::max(1, prepass_width * prepass_height), 0);
    std::vector<lightray> rays((size_t)width * height);
    int use_prepass = a->use_prepass < 0 ? GRT_USE_PREPASS : a->use_prepass;
    o->rays_traced = 0;
    o->t_integrate_s = 0;
    double t0 = now_s();
    if (use_prepass && prepass_width > 0 && prepass_height > 0) {
        // main.cpp:2387-2436
        clear_termination_buffer(termination_buffer.data(), prepass_width, prepass_height);
        int n = init_rays_generic(camera_generic, camera_quat, rays.data(), prepass_width, prepass_height, termination_buffer.data(),
                                  prepass_width, prepass_height, 0, e, cfg, dfg, 1);
        double ti = now_s();
        do_generic_rays(rays.data(), n, cfg, dfg, &cnt);
        o->t_integrate_s += now_s() - ti;
        o->rays_traced += n;
        calculate_singularities(rays.data(), n, termination_buffer.data(), prepass_width, prepass_height);
    }
    if (o->termination) memcpy(o->termination, termination_buffer.data(), sizeof(int) * (size_t)prepass_width * prepass_height);

    // main.cpp:2438-2476
    int ray_count = init_rays_generic(camera_generic, camera_quat, rays.data(), width, height, termination_buffer.data(), prepass_width, prepass_height,
                                      0, e, cfg, dfg, 0);
    o->t_init_s = now_s() - t0 - o->t_integrate_s;
    double ti = now_s();
    o->rays_traced += integrate(rays.data(), ray_count, &cnt);
    o->t_integrate_s += now_s() - ti;

    std::vector<render_data> rdata((size_t)width * height);
    memset(rdata.data(), 0, sizeof(render_data) * rdata.size());
    double t1 = now_s();
    calculate_render_data(rays.data(), ray_count, rdata.data(), width, height, cfg, dfg);
    o->t_render_data_s = now_s() - t1;

    o->extra_rays = 0;
    o->t_adaptive_s = 0;
    if (dfg.adaptive_sampling) {
        // main.cpp:2478-2511
        double t2 = now_s();
        std::vector<lightray> extra((size_t)width * height);
        int extra_count = handle_adaptive_sampling(rays.data(), rdata.data(), extra.data(), camera_generic, camera_quat, e, width, height, cfg, dfg);
        o->t_adaptive_s = now_s() - t2;
        double t3 = now_s();
        int extra_traced = integrate(extra.data(), extra_count, &cnt);
        o->t_integrate_s += now_s() - t3;
        double t4 = now_s();
        calculate_render_data(extra.data(), extra_count, rdata.data(), width, height, cfg, dfg);
        o->t_render_data_s += now_s() - t4;
        o->extra_rays = win ? extra_traced : extra_count;
        o->rays_traced += extra_traced;
    }
    if (o->rays && !win) memcpy(o->rays, rays.data(), sizeof(lightray) * (size_t)ray_count);
    if (o->rdata && !win) memcpy(o->rdata, rdata.data(), sizeof(render_data) * rdata.size());
    const int wx0 = win ? win[0] : 0, wy0 = win ? win[1] : 0, wx1 = win ? win[2] : width, wy1 = win ? win[3] : height, ww = wx1 - wx0;
    if (o->rdata && win)
        for (int y = wy0; y < wy1; y++) memcpy(o->rdata + (size_t)(y - wy0) * ww, rdata.data() + (size_t)y * width + wx0, sizeof(render_data) * (size_t)ww);

    o->t_render_s = 0;
    if (o->image && a->bg1) {
        double t5 = now_s();
        MipImage bg1 = build_mip_image(a->bg1, a->bg1_w, a->bg1_h);
        MipImage bg2_storage;
        const MipImage* bg2 = &bg1;
        if (a->bg2) { bg2_storage = build_mip_image(a->bg2, a->bg2_w, a->bg2_h); bg2 = &bg2_storage; }
        double t6 = now_s();
        (void)t5;
        // the framebuffer is cleared to nothing in the reference; every pixel id is written by `render`
#pragma omp parallel for schedule(dynamic, 256)
        for (int k = 0; k < ww * (wy1 - wy0); k++) {
            int id = (wy0 + k / ww) * width + wx0 + k % ww;
            int sx, sy;
            float4 px = render_pixel(rdata.data(), id, bg1, *bg2, width, height, a->max_probes, dfg, &sx, &sy);
            if (sx < wx0 || sx >= wx1 || sy < wy0 || sy >= wy1) continue;   // (a record that was never written names pixel 0,0)
            float* dst = o->image + ((size_t)(sy - wy0) * ww + (sx - wx0)) * 4;
            dst[0] = px.x; dst[1] = px.y; dst[2] = px.z; dst[3] = px.w;
        }
        o->t_render_s = now_s() - t6;
    }
    o->trial_steps = cnt.trial_steps;
    o->accepted_steps = cnt.accepted_steps;
    o->t_total_s = now_s() - t_start;
    return 0;
}

int oracle_frame(const OracleFrameArgs* a, OracleFrameOut* o) { return frame_impl(a, o, nullptr); }
int oracle_frame_window(const OracleFrameArgs* a, int x0, int y0, int x1, int y1, OracleFrameOut* o) {
    if (x0 < 0 || y0 < 0 || x1 > a->width || y1 > a->height || x1 <= x0 || y1 <= y0) return -1;
    const int win[4] = {x0, y0, x1, y1};
    return frame_impl(a, o, win);
}

// Bounded CPU-baseline sample: trace cw x ch rays at pixels (x0 + i*stride, y0 + j*stride) of the full
// width x height pixel grid (same pixel directions as the full frame), integration stage only.
int oracle_integrate_sample(const OracleFrameArgs* a, int x0, int y0, int cw, int ch, int stride, long long* trial_steps, long long* accepted_steps, double* seconds,
                            int* terminated_out /* cw*ch or null */, int* steps_out /* cw*ch or null */) {
    Features dfg = to_features(a->features);
    grt_params cfg = to_params(a->params);
    float4 camera_quat = from_arr(a->camera_quat);
    float4 camera_generic = cart_to_generic(from_arr(a->camera_pos), a->flip ? 1.f : 0.f, cfg);
    float4 e[4];
    calculate_tetrads(camera_generic, f3(a->basis_speed[0], a->basis_speed[1], a->basis_speed[2]), e, cfg, 1);
    std::vector<lightray> rays((size_t)cw * ch);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < cw * ch; i++) {
        int cx = x0 + (i % cw) * stride, cy = y0 + (i / cw) * stride;
        rays[i] = make_pixel_ray(cx, cy, a->width, a->height, camera_generic, camera_quat, e, 0, cfg, dfg, true);
    }
    Counters cnt;
    double t0 = now_s();
    do_generic_rays(rays.data(), cw * ch, cfg, dfg, &cnt, steps_out);
    *seconds = now_s() - t0;
    *trial_steps = cnt.trial_steps;
    *accepted_steps = cnt.accepted_steps;
    if (terminated_out)
        for (int i = 0; i < cw * ch; i++) terminated_out[i] = rays[i].terminated;
    return 0;
}

// Trace one explicit ray and record every accepted state (physics pins: conserved quantities).
int oracle_trace_path(const float* pos, const float* vel, const OracleFeatures* f, const float* params, int max_points, float* positions, float* velocities,
                      int* terminated) {
    Features dfg = to_features(*f);
    grt_params cfg = to_params(params);
    lightray ray = geodesic_to_ray(0, 0, from_arr(pos), from_arr(vel), nullptr, cfg);
    // re-run the integration loop step by step: same code path as do_generic_ray, recording states
    float4 position = ray.position, velocity = ray.velocity, acceleration = ray.acceleration;
    float next_ds = 0.00001f;
#if GRT_ADAPTIVE_PRECISION
    (void)acceleration_to_precision(acceleration, dfg.max_acceleration_change, &next_ds);
#endif
    int n = 0;
    *terminated = 0;
    for (int i = 0; i < 4096 * 4 && n < max_points; i++) {
        to_arr(position, positions + n * 4);
        to_arr(velocity, velocities + n * 4);
        n++;
        float4 polar_position = generic_to_spherical(position, cfg);
        float r_value = get_distance_to_object(polar_position, cfg);
        float ds = linear_val(fabsf(r_value), 3, dfg.max_precision_radius, 0.2f, 0.5f);
#if GRT_ADAPTIVE_PRECISION
        ds = next_ds;
#endif
        if (fabsf(r_value) < dfg.max_precision_radius) ds = fminf(ds, 0.2f);
        else ds = 0.1f * (fabsf(r_value) - dfg.max_precision_radius) + 0.2f;
        bool should_terminate = fabsf(polar_position.y) >= dfg.universe_size;
#if GRT_SINGULAR
        should_terminate |= fabsf(polar_position.y) < GRT_SINGULAR_TERMINATOR;
#endif
        if (should_terminate) { *terminated = 1; break; }
        while (true) {
            float4 np, nv, na;
            float dl = 1;
            step_verlet(position, velocity, acceleration, ds, &np, &nv, &na, &dl, cfg, dfg);
#if GRT_ADAPTIVE_PRECISION
            if (fabsf(r_value) < dfg.max_precision_radius) {
                int res = calculate_ds_error(ds, na, dfg.max_acceleration_change, &next_ds, dfg);
                if (res == DS_RETURN) return n;
                if (res == DS_SKIP) { ds = fminf(next_ds, 0.2f); continue; }
            }
#endif
            position = np; velocity = nv; acceleration = na;
            break;
        }
        if (any_degenerate(position) || any_degenerate(velocity) || any_degenerate(acceleration)) break;
    }
    return n;
}

// main.cpp:2675-2760: boost_tetrad -> init_inertial_ray -> get_geodesic_path -> 4x parallel_transport_quantity
int oracle_geodesic_snapshot(const float* camera_generic, const float* tetrad_in, const float* basis_speed, const OracleFeatures* f, const float* params,
                             int max_path_length, float* positions, float* velocities, float* ds, float* transported /* 4 * max * 4 */, float* boosted_tetrad) {
    Features dfg = to_features(*f);
    grt_params cfg = to_params(params);
    float4 e[4];
    for (int i = 0; i < 4; i++) e[i] = from_arr(tetrad_in + i * 4);
    float3 speed = f3(basis_speed[0], basis_speed[1], basis_speed[2]);
    boost_tetrad(from_arr(camera_generic), speed, e, cfg);
    for (int i = 0; i < 4; i++) to_arr(e[i], boosted_tetrad + i * 4);
    // init_inertial_ray (cl.cl:3117-3141): velocity from the *boosted* tetrad and the same speed
    float4 velocity = get_timelike_vector(speed, 1, e);
    lightray trace = geodesic_to_ray(0, 0, from_arr(camera_generic), velocity, nullptr, cfg);
    std::vector<float4> pos(max_path_length), vel(max_path_length);
    std::vector<float> dsv(max_path_length);
    int cnt = get_geodesic_path(trace, pos.data(), vel.data(), dsv.data(), max_path_length, cfg, dfg);
    for (int i = 0; i < cnt; i++) { to_arr(pos[i], positions + i * 4); to_arr(vel[i], velocities + i * 4); ds[i] = dsv[i]; }
    std::vector<float4> q(std::max(cnt, 1));
    for (int k = 0; k < 4; k++) {
        parallel_transport_quantity(pos.data(), vel.data(), dsv.data(), e[k], cnt, q.data(), cfg);
        for (int i = 0; i < cnt; i++) to_arr(q[i], transported + ((size_t)k * max_path_length + i) * 4);
    }
    return cnt;
}

void oracle_geodesic_interpolate(const float* positions, const float* velocities, const float* ds, const float* transported, int max_path_length, int cnt,
                                 float target_time, int parallel_transport_observer, const float* basis_speed, const float* params, float* camera_generic_out,
                                 float* tetrad_out, float* velocity_out) {
    grt_params cfg = to_params(params);
    std::vector<float4> pos(std::max(cnt, 1)), vel(std::max(cnt, 1));
    std::vector<std::vector<float4>> te(4, std::vector<float4>(std::max(cnt, 1)));
    for (int i = 0; i < cnt; i++) {
        pos[i] = from_arr(positions + i * 4);
        vel[i] = from_arr(velocities + i * 4);
        for (int k = 0; k < 4; k++) te[k][i] = from_arr(transported + ((size_t)k * max_path_length + i) * 4);
    }
    const float4* t_e[4] = {te[0].data(), te[1].data(), te[2].data(), te[3].data()};
    float4 cam = f4(0, 0, 0, 0), v = f4(0, 0, 0, 0), e[4] = {f4(1, 0, 0, 0), f4(0, 1, 0, 0), f4(0, 0, 1, 0), f4(0, 0, 0, 1)};
    handle_interpolating_geodesic(pos.data(), vel.data(), ds, &cam, t_e, e, target_time, cnt, parallel_transport_observer,
                                  f3(basis_speed[0], basis_speed[1], basis_speed[2]), &v, cfg);
    to_arr(cam, camera_generic_out);
    to_arr(v, velocity_out);
    for (int i = 0; i < 4; i++) to_arr(e[i], tetrad_out + i * 4);
}

}  // extern "C"
