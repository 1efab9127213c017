// Metric plugin loader: runs the reference's metric scripts unchanged and turns them into
// generated CUDA `__device__` source (plus the identical text for the CPU oracle).
//
// Mirrors, on the host side, the reference's plugin surface:
//   /root/reference/content_manager.cpp:9-112   script + config discovery, one-level inheritance
//   /root/reference/js_interop.cpp:665-959      CShim / CMath / $cfg / $pin bindings, 4 -> 16 expansion
//   /root/reference/metric.hpp:82-274,330-435,557-708,725-959   metric_info, Christoffel/accel,
//        total differentials, metric_config, diagonal reduction, symmetry test, generated-code ABI
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "sym.hpp"

namespace grt {

enum class CoordinateSystem { X_Y_THETA_PHI, CARTESIAN, CYLINDRICAL, OTHER };

// /root/reference/metric.hpp:330-357 (same defaults)
struct MetricConfig {
    std::string name, description;
    bool use_prepass = false;
    float max_acceleration_change = 0.0000001f;
    bool singular = false;
    bool traversable_event_horizon = false;
    float singular_terminator = 1;
    bool adaptive_precision = true;
    bool detect_singularities = false;
    bool follow_geodesics_forward = false;
    bool has_cylindrical_singularity = false;
    float cylindrical_terminator = 0.005f;
    CoordinateSystem system = CoordinateSystem::X_Y_THETA_PHI;
    std::string to_polar, from_polar, origin_distance, coordinate_periodicity, inherit_settings;
    bool unconditionally_nonsingular = false;
    std::vector<std::string> warnings;  // unknown keys (the reference prints a warning)
    void load_json(const std::string& text);
};

// /root/reference/js_interop.hpp:12-23, js_interop.cpp:38-68
struct ConfigVariables {
    std::vector<std::string> names;
    std::vector<float> default_values, current_values;
    void add(const std::string& name, float val);
    void set_default(const std::string& name, float val);
};

struct ContentFolder {
    std::string folder;
    std::vector<std::string> metrics, configs, coordinates, origins;
    std::vector<MetricConfig> base_configs;  // parallel to configs, loaded without inheritance
    std::vector<std::string> sorting;
};

// /root/reference/content_manager.hpp:47-58
struct ContentManager {
    std::vector<ContentFolder> dirs;
    void add_content_folder(const std::string& folder);
    bool lookup_config_file(const std::string& name, std::string& out) const;       // by JSON "name"
    bool lookup_metric_file(const std::string& name, std::string& out) const;       // by JSON "name"
    bool lookup_coordinates_file(const std::string& name, std::string& out) const;  // by file stem
    bool lookup_origins_file(const std::string& name, std::string& out) const;
    std::vector<std::string> metric_names() const;                                  // sorted as the UI lists them
};

MetricConfig load_config(const ContentManager& cm, const std::string& filename, bool inherit);

struct MetricDescriptor {
    sym::Graph graph;
    bool big = false;             // GENERIC_BIG_METRIC: off-diagonal terms present
    bool is_spherical = false;    // is_polar_spherically_symmetric (metric.hpp:557-622)
    bool has_period = false;
    int g[16];                    // row-major g_ij (diagonal metrics: off-diagonals are const 0)
    int dg[64];                   // dg[k*16 + i*4 + j] = d g_ij / d v_(k+1)
    int to_polar[4], dto_polar[4], from_polar[4], dfrom_polar[4];
    int distance = 0;
    int period[4];
    int accel_ref[4];             // -Gamma^mu_ab v^a v^b with symbolic inverse (metric.hpp:184-244)
    int accel[4];                 // same acceleration, lowered-index formulation (product kernels)
    int radii[2];                 // {polar r, distance-to-object} composed on native coordinates
    // Derived parameters: maximal sub-expressions that depend on the `$cfg` parameters only (the double-Kerr family spends a
    // quarter of its operations on them).  They are evaluated once on the host (float64) whenever a parameter changes and
    // travel in the parameter block behind the named ones, instead of being recomputed - or kept live in registers - by every
    // ray at every step.  Empty unless the metric has at least 16 such operations.
    std::vector<int> derived;
    size_t flops_expr = 0;        // ops in the emitted g + dg DAG (SURVEY.md 8(d) F_expr)
    size_t flops_accel = 0;       // ops in the emitted lowered acceleration DAG
    size_t flops_accel_ref = 0;
};

struct Metric {
    std::string js_path;
    MetricConfig cfg;
    ConfigVariables params;       // sandbox.cfg: first-access order
    MetricDescriptor desc;
    bool constant_theta() const { return cfg.system == CoordinateSystem::X_Y_THETA_PHI && desc.is_spherical; }
};

// Throws std::runtime_error with the script/JSON error message (the reference catches these in
// content_manager.cpp:131-140 and keeps the previous metric).
std::unique_ptr<Metric> load_metric_from_script(const ContentManager& cm, const std::string& js_path);

// Generated-code ABI (replaces build_argument_string, metric.hpp:725-959).  `is_static`
// substitutes the current parameter values as literals (metric.hpp:494-547), otherwise the
// functions read `cfg.<name>` from `struct grt_params`.
std::string emit_metric_source(const Metric& m, bool is_static);

// W_V1..4 of the step-size control (metric.hpp:849-869); the generated program carries them as GRT_W_V1..4, the host needs
// their maximum for the step-control constants of GrtFrame.
void acceleration_weights(const Metric& m, double w[4]);

// The parameter block the generated functions read (`struct grt_params`): the named parameters' current values followed by the
// derived parameters evaluated from them.
std::vector<float> pack_params(const Metric& m);

}  // namespace grt
