/* geodesic_raytracing_b200 - C ABI of libgrt.so
 *
 * The drop-in boundary for the reference's hot path.  Two halves:
 *
 *  (b1) plugin boundary - loads the reference's metric scripts unchanged.  Replaces, in
 *       /root/reference: content_manager.{hpp,cpp} (content folders, config inheritance),
 *       js_interop.{hpp,cpp} (script sandbox, CShim/CMath/$cfg/$pin), metric.hpp
 *       (metric_descriptor::load, build_argument_string), metric_manager.hpp:19-170 (program build).
 *  (b2) device boundary - replaces the OpenCL-by-name calls of the frame loop
 *       (`cqueue.exec("<kernel>", args, global, local)`, main.cpp:139-205, 2244-2526, 2675-2760)
 *       and render_state.hpp (device buffers) with one call per frame.
 *
 * Conventions: plain C, no exceptions cross the ABI; functions returning int return 0 on
 * success and a negative GRT_E_* code on failure, pointer-returning functions return NULL on
 * failure; the message is available from grt_last_error() (thread local).  All calls for one
 * context come from one thread (the reference has a single issuing thread).
 * There is NO CPU fallback: device functions fail with GRT_E_CUDA when no sm_100 GPU is usable.
 */
#ifndef GRT_H_INCLUDED
#define GRT_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRT_OK 0
#define GRT_E_INVALID -1   /* bad argument */
#define GRT_E_SCRIPT -2    /* script / JSON error (message carries the QuickJS-style text) */
#define GRT_E_COMPILE -3   /* NVRTC failed; grt_last_error() holds the log */
#define GRT_E_CUDA -4      /* CUDA runtime error, or no usable GPU */
#define GRT_E_STATE -5     /* call order (no metric bound, no background, ...) */

typedef struct grt_content grt_content;   /* content_manager */
typedef struct grt_metric grt_metric;     /* metrics::metric {desc, metric_cfg, sand} */
typedef struct grt_ctx grt_ctx;           /* one GPU: program + render_state + camera */

const char* grt_last_error(void);
const char* grt_version(void);

/* ---- (b1) plugin boundary ------------------------------------------------------------- */

/* content_manager::add_content_folder (content_manager.cpp:302-314): a folder with
 * <m>.js + <m>.json, coordinates/<c>.js, origins/<o>.js and optionally sorting.json. */
grt_content* grt_content_create(void);
void grt_content_destroy(grt_content*);
int grt_content_add_folder(grt_content*, const char* folder);
/* metric friendly names (JSON "name") in UI order (sorting.json first, content_manager.cpp:193-253) */
int grt_content_metric_count(const grt_content*);
const char* grt_content_metric_name(const grt_content*, int index);

/* load_metric_from_script (content_manager.cpp:9-53): runs the metric script, its coordinate,
 * origin and periodicity scripts in one shared sandbox, derives g, dg, coordinate differentials,
 * the acceleration and the classification flags.  `by_name` looks the script up through the
 * JSON "name" as `-bench <name>` does (main.cpp:967-986, content_manager.cpp:316-379). */
grt_metric* grt_metric_load(grt_content*, const char* js_path);
grt_metric* grt_metric_load_by_name(grt_content*, const char* friendly_name);
void grt_metric_destroy(grt_metric*);

/* $cfg parameters in first-access order = field order of `struct dynamic_config`
 * (js_interop.cpp:38-68, 795-815; cl.cl:940-945) */
int grt_metric_param_count(const grt_metric*);
const char* grt_metric_param_name(const grt_metric*, int index);
float grt_metric_param_default(const grt_metric*, int index);
float grt_metric_param_get(const grt_metric*, int index);
int grt_metric_param_set(grt_metric*, int index, float value);

typedef struct grt_metric_info {
    int is_big;                 /* GENERIC_BIG_METRIC: off-diagonal terms (metric.hpp:665-708) */
    int is_constant_theta;      /* GENERIC_CONSTANT_THETA (metric.hpp:557-622, 740) */
    int use_prepass;
    int singular;
    int traversable_event_horizon;
    int adaptive_precision;
    int detect_singularities;
    int coordinate_system;      /* 0 X_Y_THETA_PHI, 1 CARTESIAN, 2 CYLINDRICAL, 3 OTHER */
    int has_periodicity;
    int unconditionally_nonsingular;
    float max_acceleration_change;
    float singular_terminator;
    float w_v[4];               /* step-error weights W_V1..4 (metric.hpp:849-869) */
    int flops_expr;             /* ops in the emitted g + dg DAG (SURVEY.md 8(d) F_expr) */
    int flops_accel;            /* ops in the emitted lowered-index acceleration */
    int flops_accel_reference;  /* ops in the emitted reference-formulation acceleration (GEO_ACCEL0..3) */
} grt_metric_info;
int grt_metric_get_info(const grt_metric*, grt_metric_info* out);
const char* grt_metric_name(const grt_metric*);
const char* grt_metric_description(const grt_metric*);

/* The generated-code ABI (replaces build_argument_string, metric.hpp:725-959): C source defining
 * struct grt_params and grt_metric / grt_metric_partials / grt_to_polar / ... / grt_accel.
 * is_static != 0 substitutes the current parameter values as literals (metric.hpp:494-547).
 * The string stays valid until the next call for the same metric. */
const char* grt_metric_source(grt_metric*, int is_static);

/* The parameter block the generated functions read (`struct grt_params` of grt_metric_source): the named `$cfg` values followed by
 * the metric's derived parameters - sub-expressions of the parameters alone, evaluated on the host whenever a value changes so
 * that no ray recomputes them (none for most metrics; the double-Kerr family has dozens).  Returns the number of floats;
 * copies at most max_count of them when out is not NULL.  Code that runs the generated source elsewhere (the CPU oracle) uses it. */
int grt_metric_packed_params(const grt_metric*, float* out, int max_count);

/* Plugin-layer diagnostic, host only, no device involved: evaluates the metric's symbolic expressions (what the
 * emitted functions compute) in double precision at one point with the current parameter values - g (16, row-major),
 * dg (64, [k*16 + i*4 + j] = d_k g_ij), the acceleration -Gamma^i_ab v^a v^b in the emitted arrangement and in the
 * reference formulation (metric.hpp:184-244), to_polar / from_polar (4 each) and the origin distance of `pos` taken as
 * polar coordinates.  Any output pointer may be NULL.  The tests use it to pin the script engine and the symbolic
 * layer against the reference's own QuickJS engine running the same scripts; rendering never calls it. */
int grt_metric_eval_symbolic(const grt_metric*, const double pos[4], const double vel[4], double* g16, double* dg64, double* accel4, double* accel_ref4,
                             double* to_polar4, double* from_polar4, double* distance1);

/* ---- (b2) device boundary ------------------------------------------------------------- */

/* dynamic_feature_config (main.cpp:1123-1158, dynamic_feature_config.cpp:182-237); the
 * triangle-renderer features (use_triangle_rendering, ray_skip) are out of scope. */
typedef struct grt_features {
    float universe_size;                 /* 20 */
    float max_acceleration_change;       /* overwritten per metric on grt_set_metric (metric_manager.hpp:50) */
    float max_precision_radius;          /* 10 */
    float min_step;                      /* 1e-6 */
    float adaptive_sampling_threshold;   /* 32 (graphics_settings.hpp:40, main.cpp:1559) */
    float field_of_view;                 /* 90 */
    int redshift;                        /* 0 */
    int reparameterisation;              /* 0 */
    int use_old_redshift;                /* 0 */
    int adaptive_sampling;               /* 1 */
    int max_probes;                      /* anisotropy, 8 (graphics_settings.hpp:34) */
    int use_prepass;                     /* -1: follow the metric's JSON */
} grt_features;
void grt_features_default(grt_features* out);

typedef struct grt_stats {
    uint64_t trial_steps;       /* Verlet bodies executed incl. rejected retries (SURVEY.md 8(d) unit of work) */
    uint64_t accepted_steps;
    uint64_t rays_integrated;   /* rays that entered the integration loop (prepass + first pass + refinement) */
    uint64_t rays_finished;
    uint64_t extra_rays;        /* refinement rays spawned by adaptive sampling */
    float ms_camera, ms_prepass, ms_init, ms_integrate, ms_render_data, ms_adaptive, ms_render, ms_total;
    int integrate_launches, total_launches;
} grt_stats;

grt_ctx* grt_create(int device);
void grt_destroy(grt_ctx*);

/* Builds (NVRTC -> sm_100a cubin, cached on disk by source hash) and binds the program for a
 * metric (metric_manager::check_recompile, metric_manager.hpp:19-170).  is_static: parameters
 * folded as literals; otherwise they are kernel arguments and grt_update_params() re-reads them
 * without a rebuild.  Copies max_acceleration_change from the metric (metric_manager.hpp:50). */
int grt_set_metric(grt_ctx*, grt_metric*, int is_static);
int grt_update_params(grt_ctx*);
/* Soft recompile (metric_manager.hpp:153-219): NVRTC runs on a background thread while the bound program
 * keeps rendering; grt_build_poll returns 0 while the build runs, 1 once the new program has been swapped
 * in (or when nothing is pending), a negative code if the build failed (the old program stays bound).
 * A new request supersedes one still in flight.  The metric must stay alive until the poll returns != 0. */
int grt_set_metric_async(grt_ctx*, grt_metric*, int is_static);
int grt_build_poll(grt_ctx*);
/* compile only (no GPU needed): used by build checks to pre-populate the cubin cache */
int grt_compile_metric(grt_metric*, int is_static, int fast_math, const char** log_out);

/* NVRTC `--use_fast_math` for programs built after this call (default on: the reference builds
 * with -cl-unsafe-math-optimizations and native_* functions, metric_manager.hpp:70, cl.cl:115-117) */
int grt_set_fast_math(grt_ctx*, int enabled);

/* Scheduling of the integration kernel (results do not depend on it): `blocks_per_sm` 128-thread
 * blocks of the persistent grid per SM and frame slot (0 = as many as fit: throughput; fewer = lower
 * per-step latency for the longest rays).  Default 3: measured best on the 4K kerr_schild frame
 * (profiles/round1_tuning.md).  `donate_threshold` (0..31) is accepted and ignored: it steered the
 * in-kernel exchange of surviving rays between warps, which never paid for its atomics and was taken
 * out of the loop in round 2 (profiles/round2_tuning.md); the argument stays so callers need not change. */
int grt_set_tuning(grt_ctx*, int blocks_per_sm, int donate_threshold);

/* Frame slots that asynchronous frames rotate through (1..16, default 2 as the reference's two
 * render_state objects, main.cpp:1463-1469).  More slots fill the latency-bound stretches of a frame
 * (prepass, tails of the integration launches) with other frames' rays: throughput up, latency up. */
int grt_set_frames_in_flight(grt_ctx*, int n);

int grt_set_features(grt_ctx*, const grt_features*);
int grt_get_features(const grt_ctx*, grt_features* out);

/* load_mipped_image (graphics_settings.cpp:152-212): RGBA8 pixels -> mip pyramid on the device.
 * which = 0: the r >= 0 side (mip_background), 1: the r < 0 side (mip_background2);
 * if side 1 is never set it shares side 0 (graphics_settings.cpp:214-243). */
int grt_set_background(grt_ctx*, const uint8_t* rgba, int width, int height, int which);
/* Background file -> RGBA8 pixels for grt_set_background (the reference loads its backgrounds/ PNG files with sf::Image,
 * graphics_settings.cpp:214-243, main.cpp:233-234): non-interlaced PNG, any colour type, 8/16 bit.  Host only.
 * Returns a malloc'ed width*height*4 buffer the caller releases with grt_image_free, or NULL (grt_last_error). */
uint8_t* grt_image_load_png(const char* path, int* width, int* height);
void grt_image_free(uint8_t* pixels);
/* RGBA8 pixels -> PNG file: the screenshot of main.cpp:2783-2812 (sf::Image::saveToFile). GRT_FORMAT_RGBA8 frames already carry the
 * reference's screenshot encoding (clamp, linear -> sRGB, clamp, * 255 truncated, main.cpp:2791-2799). Host only. */
int grt_image_save_png(const char* path, const uint8_t* rgba, int width, int height);
/* device-resident variant (pixels already in HBM) */
int grt_set_background_device(grt_ctx*, const void* rgba_device, int width, int height, int which);

/* camera (main.cpp:664-673, 2244-2262): position (t,x,y,z) cartesian, quaternion (x,y,z,w),
 * flip_r = far side of a wormhole (cart_to_generic_kernel `flip`), observer 3-velocity in the
 * camera tetrad frame (cartesian_basis_speed). */
int grt_set_camera(grt_ctx*, const float pos_txyz[4], const float quat_xyzw[4], int flip_r);
int grt_set_observer_speed(grt_ctx*, const float speed_xyz[3]);

/* camera attached to a geodesic (main.cpp:2675-2760, 2264-2297): snapshot integrates the
 * observer's timelike geodesic from the current camera and parallel transports the tetrad;
 * returns the number of recorded points (<= 64000, main.cpp:1230).  set_camera_on_geodesic
 * places the camera at proper time tau on it (handle_interpolating_geodesic). */
int grt_snapshot_geodesic(grt_ctx*, float linear_basis_speed, int parallel_transport_observer);
int grt_set_camera_on_geodesic(grt_ctx*, float proper_time);
int grt_clear_camera_geodesic(grt_ctx*);

#define GRT_FORMAT_RGBA32F 0
#define GRT_FORMAT_RGBA8 1
#define GRT_OUT_HOST 0         /* host pointer, one D2H copy inside the call, returns when the frame is there */
#define GRT_OUT_DEVICE 1       /* device pointer, no copy, asynchronous */
#define GRT_OUT_HOST_ASYNC 2   /* pinned host pointer, D2H copy queued behind the frame, asynchronous */
/* One frame (main.cpp:2244-2526): camera/tetrad -> [prepass] -> init_rays -> integrate ->
 * render_data -> [adaptive sampling -> integrate -> render_data] -> render.
 * out: width*height pixels in `format`, row-major; out_is_device is one of GRT_OUT_*.  stats may be
 * NULL; asking for stats makes the call wait for the frame.  Asynchronous frames alternate between
 * two frame slots with their own stream and buffers (the reference keeps two render_state objects
 * in flight, main.cpp:1463-1469), so two consecutive frames overlap on the GPU; grt_synchronize
 * waits for all of them. */
int grt_render(grt_ctx*, int width, int height, void* out, int format, int out_is_device, grt_stats* stats);
/* Rows [y0, y1) of a width x height frame (row stripes for multi-GPU tiling);
 * out receives (y1 - y0) * width pixels.  y0, y1 must be even when adaptive sampling is on. */
int grt_render_rows(grt_ctx*, int width, int height, int y0, int y1, void* out, int format, int out_is_device, grt_stats* stats);
/* Several row stripes [y0s[i], y1s[i]) in one pass (sorted, disjoint): the interleaved stripes one
 * GPU owns when a frame is tiled over several GPUs.  All stripes share one integration launch so
 * the divergent tail of the slowest rays is paid once.  out receives the stripes back to back. */
int grt_render_stripes(grt_ctx*, int width, int height, const int* y0s, const int* y1s, int n_stripes, void* out, int format, int out_is_device, grt_stats* stats);
/* grt_synchronize waits for every frame in flight; grt_stream returns the cudaStream_t of the most
 * recently submitted frame (for ordering a consumer, e.g. the NCCL gather, behind it). */
int grt_synchronize(grt_ctx*);
void* grt_stream(grt_ctx*);

/* ---- one frame over several GPUs (replaces the single-context frame loop of main.cpp:1069-1095, 2244-2526) ----
 * The frame is tiled in row stripes (grt_render_stripes); rank 0 OWNS the assembled frames.  Every rank's render kernel stores its
 * stripes straight into the owner's frame buffer through a peer-mapped pointer (NVLink), and the hand-shake - "my stripes of frame
 * seq are in", "buffer f may be written again" - is one word per buffer moved by stream memory operations (cuStreamWriteValue32 to
 * the peer-mapped word, cuStreamWaitValue32 on local words): no collective kernel competes with the persistent integration blocks for SMs.
 *   one process per GPU (torchrun): grt_comm_init on every rank -> exchange the GRT_COMM_HANDLE_BYTES handles (any byte transport,
 *       e.g. torch.distributed.all_gather) -> grt_comm_connect(handles of all ranks in rank order);
 *   one process, several contexts (grt_headless -gpus N): grt_comm_init on every context (handle_out may be NULL) ->
 *       grt_comm_connect_local(contexts in rank order).
 * grt_comm_render(rank's stripes, seq) is asynchronous on every rank; frames must be submitted with seq = 0, 1, 2, ... on all ranks.
 * n_frames buffers rotate (seq % n_frames); a rank starts storing frame seq only after the owner has consumed frame seq - n_frames.
 * On the owner, host_out (pinned memory, may be NULL) receives the assembled frame behind the last rank's stores; grt_comm_frame
 * gives the device pointer of the buffer of frame seq, valid once the owner's stream has passed its wait (grt_synchronize, or order
 * a consumer on grt_stream).  In the single-process form call the owner LAST for each frame if host_out is not pinned. */
#define GRT_COMM_HANDLE_BYTES 64
int grt_comm_init(grt_ctx*, int rank, int world, int width, int height, int format, int n_frames, uint8_t* handle_out /* GRT_COMM_HANDLE_BYTES or NULL */);
int grt_comm_connect(grt_ctx*, const uint8_t* handles /* world * GRT_COMM_HANDLE_BYTES, rank order */);
int grt_comm_connect_local(grt_ctx** contexts_in_rank_order, int n);
/* Unmap the other ranks' regions (waits for this context's frames first). With one process per GPU every rank must disconnect
 * BEFORE any rank destroys its context: freeing a region another process still has mapped is undefined in CUDA IPC. */
int grt_comm_disconnect(grt_ctx*);
int grt_comm_render(grt_ctx*, const int* y0s, const int* y1s, int n_stripes, unsigned seq, void* host_out);
void* grt_comm_frame(grt_ctx*, unsigned seq);
/* The stripes rank `rank` of `world` renders: `stripe_rows`-row stripes dealt one per rank within every group of `world` consecutive
 * stripes, the position rotating by three from group to group, so that the cheap and the expensive rows of a frame are shared out
 * evenly.  Writes up to max_stripes (y0, y1) pairs in ascending order, returns the count. */
int grt_interleaved_stripes(int height, int rank, int world, int stripe_rows, int* y0s, int* y1s, int max_stripes);

/* Load balancing of the stripes by measured cost (sort-first rendering with feedback from a frame already rendered: a camera moves
 * little from one frame to the next, so the cost profile of the last frame predicts the next one).
 * grt_measure_stripe_costs renders every HALF stripe of the frame by itself on this context (with its halo, statistics read back)
 * and writes its trial ray-steps: cost_out[2 i] and cost_out[2 i + 1] are the upper and the lower half of stripe i.  A few
 * milliseconds per half, meant for set-up time and for an occasional refresh; returns the number of costs written (2 x stripes).
 * The result is a deterministic function of the bound program, the features and the camera: every rank that calls it gets the same
 * numbers, or one rank measures and the host distributes them.
 * grt_balanced_stripes deals the stripes to the ranks by longest-processing-time-first - a stripe heavier than 3/10 of a rank's
 * fair share as its two halves - followed by pairwise exchanges that lower the heaviest rank, keeps the number of pieces per rank
 * nearly equal, merges neighbouring pieces of one rank and returns this rank's stripes in ascending order.  With cost == NULL it is
 * grt_interleaved_stripes. */
int grt_measure_stripe_costs(grt_ctx*, int width, int height, int stripe_rows, double* cost_out, int max_costs);
int grt_balanced_stripes(int height, int rank, int world, int stripe_rows, const double* cost, int n_cost, int* y0s, int* y1s, int max_stripes);
/* The other way to share a frame out: ONE contiguous block of rows per rank (rank r renders rows [bounds[r], bounds[r + 1])), tall where
 * the rows are cheap and short where they are expensive.  A rank then pays one halo and one run of prepass rows instead of one per
 * stripe (at 8 GPUs the 54-row stripes cost 8 % more ray-steps than the frame holds).  grt_contiguous_bounds places the world + 1
 * boundaries at equal cumulative cost of a profile from grt_measure_stripe_costs (cost == NULL: equal heights); grt_refine_bounds moves
 * them by what every rank's block really cost in a frame rendered with them (the trial_steps of its grt_stats): the profile keeps its
 * shape inside a block and is scaled block by block to the measured totals (cost == NULL: uniform inside a block) - feedback an
 * interactive host applies every few frames; two rounds bring the ranks of the headline frame within a few percent.  Boundaries are
 * even, every block is at least two rows tall; both are pure functions of their arguments, so every rank computes the same partition. */
int grt_contiguous_bounds(int height, int world, int stripe_rows, const double* cost, int n_cost, int* bounds_out /* world + 1 */);
int grt_refine_bounds(int height, int world, int stripe_rows, const double* cost, int n_cost, const int* bounds, const double* measured_cost_per_rank, int* bounds_out /* world + 1 */);

/* A group: the contexts of several devices of one box driven as one renderer from one thread - what SURVEY.md 8(b2) writes as
 * grt_create(const int* devices, int ndev).  Rank 0 = devices[0] owns the assembled frame.  The setters broadcast to every context;
 * grt_group_render queues ONE frame on all devices (interleaved stripes of grt_interleaved_stripes with 54 rows, peer stores into
 * rank 0's frame, asynchronous) and, when host_out is given (pinned memory), the copy of the assembled frame behind it.
 * grt_group_ctx gives the context of a rank for everything else (statistics, introspection).  A device may appear more than once. */
typedef struct grt_group grt_group;
grt_group* grt_group_create(const int* devices, int ndev);
void grt_group_destroy(grt_group*);
int grt_group_size(const grt_group*);
grt_ctx* grt_group_ctx(grt_group*, int rank);
int grt_group_set_metric(grt_group*, grt_metric*, int is_static);
int grt_group_update_params(grt_group*);
int grt_group_set_features(grt_group*, const grt_features*);
int grt_group_set_background(grt_group*, const uint8_t* rgba, int width, int height, int which);
int grt_group_set_camera(grt_group*, const float pos_txyz[4], const float quat_xyzw[4], int flip_r);
int grt_group_set_frames_in_flight(grt_group*, int n);
int grt_group_render(grt_group*, int width, int height, int format, void* host_out);
/* Share the frame out by measured cost: one contiguous block of rows per device (grt_measure_stripe_costs on devices[0],
 * grt_contiguous_bounds, two rounds of grt_refine_bounds with every device's own ray-steps).  Call it after the metric, the features
 * and the camera are set, and again whenever the view has changed much; without it grt_group_render uses the interleaved stripes.
 * Waits for the frames in flight. */
int grt_group_rebalance(grt_group*, int width, int height);
void* grt_group_frame(grt_group*);          /* device pointer (on devices[0]) of the most recently queued frame's buffer */
int grt_group_synchronize(grt_group*);

/* ---- introspection for tests and tools --------------------------------------------------- */
typedef struct grt_camera_state {
    float pos_generic[4];
    float tetrad[16];
    int timelike_coordinate;
} grt_camera_state;
int grt_get_camera_state(grt_ctx*, grt_camera_state* out);   /* runs the camera kernel if needed */

/* ray record as the reference's `struct lightray` (render_state.hpp:8-19), 96 bytes */
typedef struct grt_lightray {
    float position[4], velocity[4], initial_quat[4], acceleration[4];
    float ku_uobsu, running_dlambda_dnew;
    int terminated, sx, sy;
    int pad[3];
} grt_lightray;
/* per-pixel shading record as `struct render_data` (cl.cl:5066-5074) */
typedef struct grt_render_data {
    float tex_coord[2];
    float z_shift;
    int sx, sy, terminated, side;
    int pad;
} grt_render_data;
/* state of the last grt_render call: first-pass rays then refinement rays (in buffer order) */
int grt_debug_ray_count(grt_ctx*, int* first_pass, int* extra);
int grt_debug_read_rays(grt_ctx*, grt_lightray* out, int max_count);
int grt_debug_read_render_data(grt_ctx*, grt_render_data* out, int max_count);
int grt_debug_read_prepass_mask(grt_ctx*, int* out, int max_count, int* pw, int* ph);
int grt_debug_read_mip_level(grt_ctx*, int which, int level, uint8_t* out, int* lw, int* lh, int* levels);
int grt_debug_read_geodesic(grt_ctx*, float* positions, float* velocities, float* ds, float* transported, int max_points);
/* evaluate generated functions on the device at `count` points (parity of g, dg, accel, maps) */
int grt_debug_eval_metric(grt_ctx*, const float* positions, const float* velocities, int count, float* g16, float* dg64, float* accel4, float* polar4,
                          float* radii2);
int grt_debug_tetrad_at(grt_ctx*, const float* positions, int count, const float speed[3], int orient, float* tetrads16, int* timelike);

/* Device time (ms) of every integration-kernel launch of the LAST frame of each frame slot in use (CUDA events on the slots' own
 * streams), i.e. of the most recent frames_in_flight frames: launch durations as they are while frames overlap. Waits for the
 * frames in flight. Returns the number of launches (writes at most max_count). */
int grt_debug_integrate_launch_ms(grt_ctx*, float* out, int max_count);
/* register count / spill info of the bound program's kernels: "name regs local_bytes;..." */
const char* grt_debug_kernel_info(grt_ctx*);
/* device time (ms) of each kernel launch of the last frame: "name:ms;..." */
const char* grt_debug_launch_log(grt_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* GRT_H_INCLUDED */
