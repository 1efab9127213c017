// Plain-data structures shared by the host runtime (grt_runtime.cu) and the sm_100a kernels.
// No functions, no host headers: the NVRTC-compiled metric module includes this text verbatim.
#pragma once

// ---- data layout in HBM -----------------------------------------------------------------------
// Ray state is SoA in float4 planes so every access is one coalesced 128-bit load/store per lane
// (the reference keeps a 96-byte AoS `struct lightray`, cl.cl:813-824 / render_state.hpp:8-19):
//   pos[i], vel[i], acc[i]           position, velocity, acceleration
//   quat[i]                          initial_quat, only allocated/touched for constant-theta metrics
//   aux[i] = {ku_uobsu, running_dlambda_dnew, terminated, sy<<16|sx}
// terminated: 0 live or dropped, 1 finished, 2 prepass-culled (as in the reference), 3 padding slot.
// A dropped ray's slot keeps its *initial* state, as in the reference (do_generic_rays only writes
// cleanly terminated rays, cl.cl:4088-4130): the adaptive-sampling stencil reads it (cl.cl:5244-5256).
struct GrtRays {
    float4* pos;
    float4* vel;
    float4* acc;
    float4* quat;
    int4* aux;
};

// Work list of one integration phase: the slots of the rays that really have to be traced (culled
// and padding slots never enter it), "hot" rays - the ones expected to run longest, next to the
// shadow - filled in from the front and the rest from the back, so the kernel starts the long rays
// first.  Item i is slots[i] for i < *n_hot, else slots[capacity - 1 - (i - *n_hot)].
struct GrtWorkList {
    int* slots;
    unsigned int* n_hot;
    unsigned int* n_cold;
    int capacity;
};

// per-pixel shading record, 16 B (the reference's `struct render_data` is 32 B, cl.cl:5066-5074;
// sx, sy are implied by the index): {u, v, z_shift, terminated | side << 8}
typedef float4 GrtRenderData;

// Runtime features (dynamic_feature_config, main.cpp:1123-1158) + the geometry of the rows this
// GPU holds.  A GPU holds one or more *windows* of the full frame: each window is a contiguous
// run of full-frame rows = an output row stripe plus the halo rows its stencils read
// (SURVEY.md 8(e)).  The windows are concatenated into `n_rows` local rows; ray slots and the
// shading records are indexed by local row, pixel directions and border rules by full-frame row.
struct GrtFrame {
    int full_w, full_h;     // full frame size: pixel directions and border rules use these
    int n_rows;             // local pixel rows held (all windows concatenated)
    int cell;               // 1: one first-pass ray per pixel, 2: one per 2x2 cell (adaptive sampling)
    int grid_w, grid_h;     // first-pass grid in cells: grid_w = full_w / cell, grid_h = n_rows / cell
    int tiles_x;            // 8x4-cell tiles per grid row
    int n_slots;            // tiles_x * tiles_y * 32 first-pass ray slots (padding slots are terminated=3)
    const int* row_global;      // [n_rows]  local row -> full-frame row
    const int* local_of_global; // [full_h]  full-frame row -> local row, -1 when not held
    const int* cellrow_refine;  // [grid_h]  adaptive sampling: 1 = this cell row feeds output pixels; in the 1/16 prepass launch:
                                //           1 = a cull test of this GPU reads this prepass row (others are not traced)
    int prepass_w, prepass_h;
    int use_prepass_mask;   // main pass consults the prepass termination mask
    int flip_geodesic_direction;
    int max_probes;
    float universe_size, max_acceleration_change, max_precision_radius, min_step;
    float sc_floor, sc_want, sc_singular;   // step-size control constants folded on the host (fill_frame_features)
    float adaptive_sampling_threshold, field_of_view;
    int redshift, reparameterisation, use_old_redshift, adaptive_sampling;
};

// One output stripe for the render kernel: output rows [out_global0, out_global0 + out_rows) of the
// full frame live at local rows starting at out_local0 inside a window that ends (exclusive) at
// local row win_local_end; pixels are written at out + out_offset_px.
struct GrtRenderStripe {
    int out_global0, out_rows;
    int out_local0;
    int win_local_end;
    unsigned long long out_offset_px;
};

struct GrtCamera {
    float4 pos_generic;     // camera position in the metric's native coordinates
    float4 e[4];            // tetrad, contravariant
    float4 quat;            // camera orientation (x, y, z, w)
    float4 basis_speed;     // observer 3-velocity in the tetrad frame
    float4 pos_cart;        // (t, x, y, z)
    float flip;
    int timelike_coordinate;
    int pad0, pad1;
};

// statistics accumulated by the integration kernel
struct GrtStats {
    unsigned long long trial_steps;     // Verlet bodies executed, rejected retries included
    unsigned long long accepted_steps;
    unsigned long long rays_integrated; // rays that entered the loop
    unsigned long long rays_finished;   // terminated == 1
};

// camera-on-a-geodesic path buffers (main.cpp:1230-1263)
struct GrtGeodesic {
    float4* positions;     // max_len
    float4* velocities;
    float* ds;
    float4* transported;   // 4 * max_len: transported tetrad legs
    int* count;
    int max_len;
};
