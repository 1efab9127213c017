// libgrt.so runtime: the C ABI of include/grt.h.
//
// (b1) plugin half: thin wrappers over metric_loader (content folders, script loading, generated
//      source).  (b2) device half: per-metric program build (NVRTC -> sm_100a cubin, cached on
//      disk), device buffers (replaces render_state.hpp), camera / geodesic-camera state and the
//      per-frame kernel sequence of /root/reference/main.cpp:2244-2526 on one CUDA stream.
// There is no CPU path in this file: every device entry point fails with GRT_E_CUDA when no
// usable GPU is present.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/grt.h"
#include "kernels/grt_device_structs.h"
#include "kernels/grt_static_kernels.h"
#include "metric_loader.hpp"

extern "C" const char GRT_SRC_STRUCTS[];
extern "C" const char GRT_SRC_COMMON[];
extern "C" const char GRT_SRC_METRIC_KERNELS[];

namespace {

thread_local std::string g_err;
constexpr int kCounterWords = 8 + 3 * 64;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(GRT_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
// The driver API (module load + launch of the NVRTC-built cubins) is reached through
// cudaGetDriverEntryPoint, so libgrt.so has no link-time dependency on libcuda.so.1 and still loads
// (and exports every symbol) on a machine without a GPU driver.
struct Driver {
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned) = nullptr;    // stream memory operations: no SM involved
    CUresult (*StreamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned) = nullptr;
    bool loaded = false;
} g_drv;

void load_driver() {
    if (g_drv.loaded) return;
    auto get = [](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) throw Error(GRT_E_CUDA, std::string("CUDA driver entry point not available: ") + name);
    };
    get("cuGetErrorString", (void**)&g_drv.GetErrorString);
    get("cuModuleLoadData", (void**)&g_drv.ModuleLoadData);
    get("cuModuleUnload", (void**)&g_drv.ModuleUnload);
    get("cuModuleGetFunction", (void**)&g_drv.ModuleGetFunction);
    get("cuFuncGetAttribute", (void**)&g_drv.FuncGetAttribute);
    get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&g_drv.OccupancyMaxActiveBlocksPerMultiprocessor);
    get("cuLaunchKernel", (void**)&g_drv.LaunchKernel);
    get("cuStreamWaitValue32", (void**)&g_drv.StreamWaitValue32);
    get("cuStreamWriteValue32", (void**)&g_drv.StreamWriteValue32);
    g_drv.loaded = true;
}

void cu(CUresult r, const char* what) {
    if (r != CUDA_SUCCESS) {
        const char* s = nullptr;
        if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
        throw Error(GRT_E_CUDA, std::string(what) + ": " + (s ? s : "unknown CUDA driver error"));
    }
}
#define CK(x) ck((x), #x)
#define CU(x) cu((x), #x)

template <class F>
int guard(F&& f) {
    try {
        return f();
    } catch (Error& e) {
        g_err = e.what();
        return e.code;
    } catch (std::exception& e) {
        g_err = e.what();
        return GRT_E_INVALID;
    }
}

uint64_t fnv1a(const std::string& s, uint64_t h) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

std::string library_dir() {
    Dl_info info;
    if (dladdr((void*)&library_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.find_last_of('/');
        if (k != std::string::npos) return p.substr(0, k);
    }
    return ".";
}
std::string cache_dir() {
    const char* e = getenv("GRT_CACHE_DIR");
    std::string d = e && *e ? e : library_dir() + "/cubin_cache";
    mkdir(d.c_str(), 0755);
    return d;
}

// ---- per-metric program ---------------------------------------------------------------------------
// Fast math is confined to what the integration loop uses: approximate division and square root, flush-to-zero and FMA
// contraction come from the NVRTC options below, and the transcendental functions of the GENERATED metric code map to the
// hardware intrinsics here.  The hand-written setup and shading kernels (ray construction, tetrads, hit angles, adaptive-sampling
// interpolation) keep the precise sincosf / acosf / atan2f / tanf: they run once per ray, and their degenerate inputs (poles,
// the equatorial row of a constant-theta metric, signed zeros) must round like the CPU oracle's libm does.
const char* kPreludeCommon =
    "#define GRT_FN static __device__ __forceinline__\n"
    "#define GRT_SQRT(a) sqrtf(a)\n"
    "#define GRT_TANH(a) tanhf(a)\n"
    "#define GRT_SIGN(a) ((a) > 0.0f ? 1.0f : ((a) < 0.0f ? -1.0f : 0.0f))\n"
    "#define GRT_INF __int_as_float(0x7f800000)\n"
    "#define GRT_NAN __int_as_float(0x7fffffff)\n";
const char* kPreludeFast =
    "#define GRT_FAST_STEP_CONTROL 1\n"     // the step-size chain of the loop in its shortest form (rsqrt, no division): see accel_to_precision
    "#define GRT_DIV(a, b) __fdividef((a), (b))\n"      // div.approx (reciprocal + multiply); a plain `/` is div.full under --prec-div=false: three more instructions
    "#define GRT_SIN(a) __sinf(a)\n"
    "#define GRT_COS(a) __cosf(a)\n"
    "#define GRT_TAN(a) __tanf(a)\n"
    "#define GRT_LOG(a) __logf(a)\n"
    "#define GRT_EXP(a) __expf(a)\n"
    "#define GRT_POW(a, b) __powf(a, b)\n";
const char* kPreludePrecise =
    "#define GRT_FAST_STEP_CONTROL 0\n"
    "#define GRT_DIV(a, b) ((a) / (b))\n"
    "#define GRT_SIN(a) sinf(a)\n"
    "#define GRT_COS(a) cosf(a)\n"
    "#define GRT_TAN(a) tanf(a)\n"
    "#define GRT_LOG(a) logf(a)\n"
    "#define GRT_EXP(a) expf(a)\n"
    "#define GRT_POW(a, b) powf(a, b)\n";

// `reparameterisation` is the one runtime feature the integration loop branches on every step; like the
// reference's static program (-DFEATURE_name=value, dynamic_feature_config.cpp:122-180) it is baked in.
std::string program_source(const grt::Metric& m, bool is_static, bool reparameterisation, bool fast_math) {
    std::string s = kPreludeCommon;
    s += fast_math ? kPreludeFast : kPreludePrecise;
    s += reparameterisation ? "#define GRT_FEATURE_REPARAMETERISATION 1\n" : "#define GRT_FEATURE_REPARAMETERISATION 0\n";
    if (const char* e = getenv("GRT_INTEGRATE_MIN_BLOCKS")) s += std::string("// GRT_TIER_FIXED\n#define GRT_INTEGRATE_MIN_BLOCKS ") + std::to_string(atoi(e)) + "\n";
    s += GRT_SRC_STRUCTS;
    s += GRT_SRC_COMMON;
    s += grt::emit_metric_source(m, is_static);
    s += GRT_SRC_METRIC_KERNELS;
    return s;
}

std::vector<std::string> nvrtc_options(bool fast_math) {
    std::vector<std::string> o = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17", "-DGRT_NVRTC=1", "-diag-suppress=177,550", "--ptxas-options=-v"};
    if (fast_math)   // --use_fast_math minus its blanket sinf -> __sinf mapping (see the prelude)
        for (const char* f : {"--ftz=true", "--prec-div=false", "--prec-sqrt=false", "--fmad=true"}) o.push_back(f);
    return o;
}

// NVRTC -> cubin, with an on-disk cache keyed by hash(source + options + NVRTC version)
// (the reference caches built programs through cl::build_program_with_cache, metric_manager.hpp:106)
std::vector<char> compile_cubin_raw(const std::string& src, bool fast_math, const std::string& name, std::string* log_out, bool* cache_hit, std::string* path_out = nullptr) {
    std::vector<std::string> opts = nvrtc_options(fast_math);
    int major = 0, minor = 0;
    nvrtcVersion(&major, &minor);
    std::string key = src;
    for (auto& o : opts) key += "\n" + o;
    key += "\nnvrtc " + std::to_string(major) + "." + std::to_string(minor);
    char hex[40];
    snprintf(hex, sizeof hex, "%016llx%016llx", (unsigned long long)fnv1a(key, 14695981039346656037ull), (unsigned long long)fnv1a(key, 0x9e3779b97f4a7c15ull));
    std::string path = cache_dir() + "/" + name + "_" + hex + ".cubin";
    if (path_out) *path_out = path;
    {
        std::ifstream f(path, std::ios::binary);
        if (f) {
            std::vector<char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            if (!data.empty()) {
                if (cache_hit) *cache_hit = true;
                if (log_out) {
                    std::ifstream lf(path + ".log", std::ios::binary);   // the ptxas resource report of the build that filled the cache
                    *log_out = "cubin cache hit: " + path + "\n" + std::string((std::istreambuf_iterator<char>(lf)), std::istreambuf_iterator<char>());
                }
                return data;
            }
        }
    }
    if (cache_hit) *cache_hit = false;
    // the program text is kept next to the cubin under the name NVRTC records in the line info,
    // so `ncu --import-source on` and cuobjdump can map SASS back to the generated code
    const std::string src_path = path.substr(0, path.size() - 6) + ".cu";
    {
        std::ofstream f(src_path, std::ios::binary);
        f.write(src.data(), (std::streamsize)src.size());
    }
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), src_path.c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS) throw Error(GRT_E_COMPILE, "nvrtcCreateProgram failed");
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());
    nvrtcResult res = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    std::string log(log_size, '\0');
    if (log_size) nvrtcGetProgramLog(prog, log.data());
    if (log_out) *log_out = log;
    if (res != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        throw Error(GRT_E_COMPILE, std::string("NVRTC: ") + nvrtcGetErrorString(res) + "\n" + log);
    }
    size_t size = 0;
    if (nvrtcGetCUBINSize(prog, &size) != NVRTC_SUCCESS || size == 0) {
        nvrtcDestroyProgram(&prog);
        throw Error(GRT_E_COMPILE, "NVRTC produced no cubin");
    }
    std::vector<char> cubin(size);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    {
        std::ofstream f(tmp, std::ios::binary);
        f.write(cubin.data(), (std::streamsize)cubin.size());
    }
    {
        std::ofstream f(path + ".log", std::ios::binary);
        f.write(log.data(), (std::streamsize)log.size());
    }
    rename(tmp.c_str(), path.c_str());
    return cubin;
}

// registers and spill bytes of k_integrate from the `ptxas -v` report in an NVRTC log; false if absent
bool integrate_resources(const std::string& log, int* regs, int* spill) {
    size_t at = log.find("Function properties for k_integrate");
    if (at == std::string::npos) return false;
    int stack = 0, st = 0, ld = 0, r = 0;
    size_t a = log.find("bytes stack frame", at), u = log.find("Used ", at);
    if (a == std::string::npos || u == std::string::npos) return false;
    size_t line = log.rfind('\n', a);
    line = log.find_first_of("0123456789", line == std::string::npos ? 0 : line + 1);   // "ptxas         .     0 bytes stack frame, ..."
    if (line == std::string::npos || sscanf(log.c_str() + line, "%d bytes stack frame, %d bytes spill stores, %d bytes spill loads", &stack, &st, &ld) != 3) return false;
    if (sscanf(log.c_str() + u, "Used %d registers", &r) != 1) return false;
    *regs = r;
    *spill = stack + st + ld;
    return true;
}

// Occupancy tier of the integration kernel.  The loop is bound by FP32 issue with long dependent chains, so more resident
// warps per scheduler pay as long as nothing spills: after the free compile, the kernel is rebuilt with
// __launch_bounds__(128, n) for the next one or two block-per-SM tiers and the highest tier that needs no stack and at most
// three tenths fewer registers is kept (kerr_schild: 87 registers / 5 blocks -> 72 / 7, +5 % frames per second; double_unequal_kerr: 179 / 2 -> 128 / 4, +4 % over 150 / 3).  The choice is
// remembered next to the cubins.  GRT_INTEGRATE_MIN_BLOCKS (environment) overrides, GRT_OCCUPANCY_TIER=0 disables.
std::vector<char> compile_cubin(const std::string& src, bool fast_math, const std::string& name, std::string* log_out, bool* cache_hit, std::vector<std::string>* files_out = nullptr) {
    const char* off = getenv("GRT_OCCUPANCY_TIER");
    auto note_files = [&](const std::string& cubin_path, const std::string& tier) {      // what a superseded static program leaves in the cache
        if (!files_out) return;
        *files_out = {cubin_path, cubin_path + ".log", cubin_path.substr(0, cubin_path.size() - 6) + ".cu"};
        if (!tier.empty()) files_out->push_back(tier);
    };
    if (src.find("// GRT_TIER_FIXED") != std::string::npos || (off && atoi(off) == 0)) {
        std::string path;
        auto r = compile_cubin_raw(src, fast_math, name, log_out, cache_hit, &path);
        note_files(path, "");
        return r;
    }
    auto with_tier = [&](int n) { return "// GRT_TIER_FIXED\n#define GRT_INTEGRATE_MIN_BLOCKS " + std::to_string(n) + "\n" + src; };
    char hex[24];
    snprintf(hex, sizeof hex, "%016llx", (unsigned long long)fnv1a(src + (fast_math ? "f" : "p") + "|rule:0.7", 14695981039346656037ull));   // the rule is part of the key: a remembered choice made under another rule is not reused
    const std::string tier_path = cache_dir() + "/" + name + "_" + hex + ".tier";
    {
        std::ifstream f(tier_path);
        int n = 0;
        if (f >> n) {
            std::string path;
            auto r = n > 0 ? compile_cubin_raw(with_tier(n), fast_math, name, log_out, cache_hit, &path) : compile_cubin_raw(src, fast_math, name, log_out, cache_hit, &path);
            note_files(path, tier_path);
            return r;
        }
    }
    std::string log, best_path;
    std::vector<char> best = compile_cubin_raw(src, fast_math, name, &log, cache_hit, &best_path);
    int regs = 0, spill = 0, chosen = 0;
    std::string best_log = log;
    if (integrate_resources(log, &regs, &spill) && regs > 0) {
        auto blocks_for = [](int r) { return std::min(16, 65536 / (((r + 7) / 8 * 8) * 128)); };
        const int base_blocks = blocks_for(regs);
        for (int n = std::min(16, base_blocks + 2); n > base_blocks; n--) {
            std::string tlog, tpath;
            std::vector<char> cand;
            try { cand = compile_cubin_raw(with_tier(n), fast_math, name, &tlog, nullptr, &tpath); } catch (const Error&) { continue; }
            int r2 = 0, s2 = 0;
            if (integrate_resources(tlog, &r2, &s2) && s2 == 0 && r2 * 10 >= regs * 7 && blocks_for(r2) >= n) {
                if (files_out) for (const char* ext : {"", ".log"}) remove((best_path + ext).c_str());      // the free compile is not the one kept
                if (files_out) remove((best_path.substr(0, best_path.size() - 6) + ".cu").c_str());
                best = cand; best_log = tlog; best_path = tpath; chosen = n;
                break;
            }
            for (const char* ext : {"", ".log"}) remove((tpath + ext).c_str());                      // a rejected tier leaves nothing behind
            remove((tpath.substr(0, tpath.size() - 6) + ".cu").c_str());
        }
    }
    {
        std::ofstream f(tier_path);
        f << chosen << "\n";
    }
    if (log_out) *log_out = best_log;
    note_files(best_path, tier_path);
    return best;
}

struct Program {
    CUmodule mod = nullptr;
    CUfunction k_camera, k_tetrad_at, k_eval_metric, k_init_rays, k_prepass_mask, k_integrate, k_render_data, k_adaptive, k_geodesic_snapshot, k_geodesic_transport,
        k_geodesic_interpolate;
    int integrate_blocks_per_sm = 1;
    std::string info;
    std::string key;
    std::vector<std::string> cache_files;   // static programs: their files in the cubin cache (removed when the program is superseded)
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8;
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Texture {
    DevBuf data;
    GrtTexture desc{};
    bool set = false;
};

}  // namespace

struct grt_content {
    grt::ContentManager cm;
    std::vector<std::string> names;
};

struct grt_metric {
    std::unique_ptr<grt::Metric> m;
    std::string source[2];
    std::string compile_log;
    std::mutex text_mutex;   // guards source[] / compile_log: the returned pointers stay valid while the text is unchanged
};

// Everything one frame in flight owns.  Two slots alternate (the reference keeps two render_state
// objects in flight, main.cpp:1463-1469): consecutive asynchronous frames run on different streams,
// so the latency-bound stretches of one frame (the 1/16 prepass, the tails of the integration
// launches) are filled with the next frame's rays.
struct FrameSlot {
    cudaStream_t stream = nullptr;
    DevBuf rays_pos, rays_vel, rays_acc, rays_quat, rays_aux, worklist, rdata, termination, out;
    // counters (zeroed every frame): [4] extra ray count; [8 + phase] work counter and [16 + 2 * phase], [17 + 2 * phase]
    // hot / cold work-list lengths of phase 0 (prepass), 1 (first pass), 2 (refinement); then GrtStats
    unsigned int* d_counters = nullptr;
    GrtStats* d_stats = nullptr;
    GrtCamera* d_cam = nullptr;
    std::vector<cudaEvent_t> events;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> integrate_events;   // around every k_integrate launch of the slot's frames (created once, re-recorded)
    int integrate_launches_last = 0;                                      // how many of them the slot's last frame used
    GrtFrame frame{};
    GrtCamera cam_staging{};   // pageable source of this slot's camera upload (copied to the driver's staging at the call)
    int last_first_pass = 0, last_extra = 0, last_pw = 0, last_ph = 0;
    bool last_prepass = false;
    bool used = false;
};

struct grt_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;   // cold-path work: textures, tables, geodesic snapshot, introspection
    FrameSlot slots[16];
    int n_slots = 2;             // frames in flight (GRT_FRAME_SLOTS, 1..16)
    unsigned frame_index = 0;
    FrameSlot* last = nullptr;       // slot of the most recently submitted frame
    bool fast_math = true;
    grt_features feat{};
    grt_metric* metric = nullptr;          // only dereferenced by grt_update_params / feature-driven rebuilds
    bool is_static = false;
    bool bound_use_prepass = false, bound_constant_theta = false;   // what the frame path needs of the bound metric
    float bound_wmax = 1.0f;     // largest acceleration weight of the bound metric (GRT_W_V1..4): step-control constants of GrtFrame
    std::map<std::string, std::unique_ptr<Program>> programs;
    Program* prog = nullptr;
    std::string static_key;      // key of the one static-parameter program kept loaded (a parameter slider in static mode must not grow memory)
    float params[256] = {0};   // struct grt_params of the bound program: named $cfg values, then the derived parameters (pack_params)

    GrtCamera cam{};
    bool camera_dirty = true, camera_on_geodesic = false;
    GrtCamera* d_cam = nullptr;

    Texture tex[2];
    DevBuf tables, identity, scratch;
    int identity_len = 0;
    int blocks_per_sm = 3;           // 128-thread blocks of the integration kernel per SM and frame slot; 0 = as many as fit (GRT_BLOCKS_PER_SM)

    // geometry cache
    std::vector<int> geom_key;
    GrtFrame frame{};
    std::vector<GrtRenderStripe> stripes;
    const int* prepass_needed = nullptr;   // device table [h / 16]: prepass rows this geometry's cull tests read
    size_t out_pixels = 0;

    // geodesic camera
    GrtGeodesic geo{};
    DevBuf geo_pos, geo_vel, geo_ds, geo_tr, geo_misc;
    int geo_count = 0;
    float4 geo_speed = make_float4(0, 0, 0, 0);
    int geo_parallel_transport = 1;
    GrtCamera geo_cam_host{};    // the interpolated camera, read back by grt_set_camera_on_geodesic

    std::string launch_log, kernel_info;

    // Multi-GPU frame delivery: rank 0 owns the frames; every rank's k_render stores its stripes straight into
    // them over NVLink (peer-mapped pointer), and the hand-shake is two words per frame buffer moved by stream memory
    // operations (write-value straight to the peer-mapped word, wait-value on local words) - no kernel, no SM.  Every rank allocates
    // one region with the same layout:
    //   [0, 4096)   ready[n_frames][world]   written by rank r into the OWNER's region: seq + 1 of the last frame r delivered
    //   [4096, 8192) freed[n_frames]          written by the owner into EVERY rank's region: seq + 1 of the last frame consumed
    //   [8192, 12288) ring[1024]              local scratch: source words of the copy-engine writes
    //   [12288, ...)  frames (owner only)
    struct Comm {
        bool active = false, connected = false, ipc = false;
        int rank = 0, world = 1, n_frames = 0, w = 0, h = 0, format = 0;
        size_t frame_bytes = 0;
        DevBuf local;
        std::vector<char*> base;      // every rank's region as mapped in this process (base[rank] == local.p)
        unsigned ring_pos = 0;
    } comm;

    // background build (metric_manager's soft recompile, metric_manager.hpp:153-219)
    struct PendingBuild {
        std::future<std::vector<char>> cubin;
        std::shared_ptr<std::string> log;
        std::string key;
        grt_metric* metric = nullptr;
        bool is_static = false;
        bool active = false;
    } build;
};

namespace grt {   // png_loader.cpp
std::vector<uint8_t> load_png_rgba(const std::string& path, int* width, int* height);
void save_png_rgba(const std::string& path, const uint8_t* rgba, int width, int height);
}

namespace {

void features_default(grt_features* f) {
    f->universe_size = 20.0f;
    f->max_acceleration_change = 0.01f;
    f->max_precision_radius = 10.0f;
    f->min_step = 0.000001f;
    f->adaptive_sampling_threshold = 32.0f;
    f->field_of_view = 90.0f;
    f->redshift = 0;
    f->reparameterisation = 0;
    f->use_old_redshift = 0;
    f->adaptive_sampling = 1;
    f->max_probes = 8;
    f->use_prepass = -1;
}

void bind(grt_ctx* c) { CK(cudaSetDevice(c->device)); }

std::string program_key(grt_ctx* c, const std::string& src) { return std::to_string(fnv1a(src, 14695981039346656037ull)) + (c->fast_math ? "f" : "p"); }
std::string program_stem(const grt::Metric& m, bool is_static) {
    std::string stem = m.js_path;
    size_t k = stem.find_last_of('/');
    if (k != std::string::npos) stem = stem.substr(k + 1);
    k = stem.find_last_of('.');
    if (k != std::string::npos) stem = stem.substr(0, k);
    return stem + (is_static ? "_static" : "_dynamic");
}

Program* load_program(grt_ctx* c, const std::string& key, const std::vector<char>& cubin);

Program* get_program(grt_ctx* c, grt_metric* gm, bool is_static) {
    const grt::Metric& m = *gm->m;
    std::string src = program_source(m, is_static, c->feat.reparameterisation != 0, c->fast_math);
    std::string key = program_key(c, src);
    auto it = c->programs.find(key);
    if (it != c->programs.end()) return it->second.get();
    std::vector<std::string> files;
    std::vector<char> cubin = compile_cubin(src, c->fast_math, program_stem(m, is_static), &gm->compile_log, nullptr, is_static ? &files : nullptr);
    Program* p = load_program(c, key, cubin);
    p->cache_files = std::move(files);
    return p;
}

// cubin -> module + kernel handles (needs the device; the NVRTC step above does not)
Program* load_program(grt_ctx* c, const std::string& key, const std::vector<char>& cubin) {
    auto p = std::make_unique<Program>();
    p->key = key;
    CU(g_drv.ModuleLoadData(&p->mod, cubin.data()));
    struct { const char* name; CUfunction* f; } fns[] = {
        {"k_camera", &p->k_camera}, {"k_tetrad_at", &p->k_tetrad_at}, {"k_eval_metric", &p->k_eval_metric}, {"k_init_rays", &p->k_init_rays},
        {"k_prepass_mask", &p->k_prepass_mask}, {"k_integrate", &p->k_integrate}, {"k_render_data", &p->k_render_data}, {"k_adaptive", &p->k_adaptive},
        {"k_geodesic_snapshot", &p->k_geodesic_snapshot}, {"k_geodesic_transport", &p->k_geodesic_transport}, {"k_geodesic_interpolate", &p->k_geodesic_interpolate}};
    std::ostringstream info;
    for (auto& f : fns) {
        CU(g_drv.ModuleGetFunction(f.f, p->mod, f.name));
        int regs = 0, local = 0;
        g_drv.FuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, *f.f);
        g_drv.FuncGetAttribute(&local, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, *f.f);
        info << f.name << " " << regs << " " << local << ";";
    }
    int blocks = 1;
    CU(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&blocks, p->k_integrate, 128, 0));
    p->integrate_blocks_per_sm = std::max(1, blocks);
    info << "integrate_blocks_per_sm " << p->integrate_blocks_per_sm << ";";
    p->info = info.str();
    Program* raw = p.get();
    c->programs[key] = std::move(p);
    return raw;
}

void sync_all(grt_ctx* c);

// A context keeps ONE static-parameter program: every new set of literals is a new module, and the superseded one is unloaded
// once the frames in flight have drained (the reference replaces its single static program, metric_manager.hpp:153-219).  Its
// files in the on-disk cache are removed too, so a parameter slider in static mode grows neither device memory nor the disk.
void retire_static_program(grt_ctx* c, Program* incoming, bool incoming_is_static) {
    if (!incoming_is_static || incoming->key == c->static_key) return;
    if (!c->static_key.empty()) {
        auto it = c->programs.find(c->static_key);
        if (it != c->programs.end() && it->second.get() != incoming) {
            sync_all(c);
            if (c->prog == it->second.get()) c->prog = nullptr;
            if (it->second->mod && g_drv.ModuleUnload) g_drv.ModuleUnload(it->second->mod);
            for (const std::string& f : it->second->cache_files) remove(f.c_str());
            c->programs.erase(it);
        }
    }
    c->static_key = incoming->key;
}

struct Launcher {
    cudaStream_t stream;
    int launches = 0, integrate_launches = 0;
    void go(CUfunction f, unsigned grid, unsigned block, void** args) {
        if (grid == 0) return;
        CU(g_drv.LaunchKernel(f, grid, 1, 1, block, 1, 1, 0, (CUstream)stream, args, nullptr));
        launches++;
    }
};

void sync_all(grt_ctx* c) {
    for (auto& S : c->slots) if (S.stream) CK(cudaStreamSynchronize(S.stream));
    CK(cudaStreamSynchronize(c->stream));
}

void fill_frame_features(grt_ctx* c, GrtFrame& fr) {
    const grt_features& f = c->feat;
    fr.universe_size = f.universe_size;
    fr.max_acceleration_change = f.max_acceleration_change;
    fr.max_precision_radius = f.max_precision_radius;
    fr.min_step = f.min_step;
    {
        // step-size control of the integration loop in its folded form (k_integrate, GRT_FAST_STEP_CONTROL), from
        // acceleration_to_precision / calculate_ds_error (cl.cl:3400-3456) with S = sum (a_i w_i)^2:
        //   diff = 655.36 / wmax * sqrt(S), floored at K / 1e10 with K = 65536 max_acceleration_change
        //   0.99 * sqrt(K / diff) = sc_want * S^(-1/4);  singular when diff / 65536 > 1e4 max_acceleration_change
        const double K = 65536.0 * (double)f.max_acceleration_change, k = 655.36 / (double)c->bound_wmax;
        const double floor_diff = K / 1e10, sing = 1e4 * 65536.0 * (double)f.max_acceleration_change;
        auto clampf = [](double v) { return (float)std::min(v, 3.0e38); };
        fr.sc_floor = clampf((floor_diff / k) * (floor_diff / k));
        fr.sc_want = clampf(0.99 * std::sqrt(K / k));
        fr.sc_singular = clampf((sing / k) * (sing / k));
    }
    fr.adaptive_sampling_threshold = f.adaptive_sampling_threshold;
    fr.field_of_view = f.field_of_view;
    fr.redshift = f.redshift;
    fr.reparameterisation = f.reparameterisation;
    fr.use_old_redshift = f.use_old_redshift;
    fr.adaptive_sampling = f.adaptive_sampling;
    fr.max_probes = f.max_probes;
    fr.flip_geodesic_direction = 0;
}

void set_grid(GrtFrame& fr) {
    fr.grid_w = fr.full_w / fr.cell;
    fr.grid_h = fr.n_rows / fr.cell;
    fr.tiles_x = (fr.grid_w + 7) / 8;
    int tiles_y = (fr.grid_h + 3) / 4;
    fr.n_slots = fr.tiles_x * tiles_y * 32;
}

const int* identity_table(grt_ctx* c, int n) {
    if (n > c->identity_len) {
        int len = std::max(n, 4096);
        std::vector<int> h(len);
        for (int i = 0; i < len; i++) h[i] = i;
        sync_all(c);
        c->identity.ensure(sizeof(int) * (size_t)len);
        CK(cudaMemcpyAsync(c->identity.p, h.data(), sizeof(int) * (size_t)len, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->identity_len = len;
    }
    return (const int*)c->identity.p;
}

// Row geometry of the stripes this GPU renders: each stripe [y0, y1) gets the
// halo its stencils read - with adaptive sampling one 2x2-cell row above (neighbour of the first
// refine decision) and two below (the Jacobian's row y1 needs cell row y1/2 decided, which reads
// cell row y1/2 + 1); without it only the Jacobian's one row below.  Overlapping windows merge.
void build_geometry(grt_ctx* c, int w, int h, const int* y0s, const int* y1s, int n) {
    const bool adaptive = c->feat.adaptive_sampling != 0;
    std::vector<int> key = {w, h, adaptive ? 1 : 0, n};
    for (int i = 0; i < n; i++) { key.push_back(y0s[i]); key.push_back(y1s[i]); }
    if (key == c->geom_key) return;
    if (adaptive && ((w & 1) || (h & 1))) throw Error(GRT_E_INVALID, "adaptive sampling needs even width and height");
    int prev_end = 0;
    for (int i = 0; i < n; i++) {
        if (y0s[i] < prev_end || y1s[i] <= y0s[i] || y1s[i] > h) throw Error(GRT_E_INVALID, "row stripes must be sorted, disjoint, non-empty and inside the frame");
        if (adaptive && ((y0s[i] & 1) || (y1s[i] & 1))) throw Error(GRT_E_INVALID, "row stripes must start and end on even rows when adaptive sampling is on");
        prev_end = y1s[i];
    }
    std::vector<int> held(h, 0), refine(h, 0);   // per full-frame row
    for (int i = 0; i < n; i++) {
        int top = adaptive ? 2 : 0, bot = adaptive ? 4 : 1;
        if (y1s[i] == h) top = std::max(top, adaptive ? 2 : 1);   // last frame row differences upwards (cl.cl:5509-5520)
        for (int y = std::max(0, y0s[i] - top); y < std::min(h, y1s[i] + bot); y++) held[y] = 1;
        // cell rows intersecting [y0, y1] decide refine-or-interpolate
        for (int y = y0s[i]; y <= std::min(h - 1, y1s[i]); y++) refine[y] = 1;
    }
    std::vector<int> row_global, local_of_global(h, -1);
    for (int y = 0; y < h; y++)
        if (held[y]) { local_of_global[y] = (int)row_global.size(); row_global.push_back(y); }
    GrtFrame fr{};
    fr.full_w = w;
    fr.full_h = h;
    fr.n_rows = (int)row_global.size();
    fr.cell = adaptive ? 2 : 1;
    set_grid(fr);
    std::vector<int> cellrow_refine(std::max(1, fr.grid_h), 0);
    if (adaptive)
        for (int gy = 0; gy < fr.grid_h; gy++) {
            int y = row_global[gy * 2];
            cellrow_refine[gy] = (refine[y] || refine[y + 1]) ? 1 : 0;
        }
    // Rows of the 1/16 prepass this GPU needs: the cull test of a first-pass ray at frame row y reads the prepass rows
    // round(y / h * ph) - 1 .. + 1 (cl.cl:3213-3232).  A GPU that renders a few stripes traces only those prepass rows - at
    // 8 GPUs about a fifth of the 135 - instead of repeating the whole prepass on every rank.  All rows for a whole frame.
    const int ph = h / 16;
    std::vector<int> prepass_needed(std::max(1, ph), 0);
    for (int y = 0; y < h && ph > 0; y++) {
        if (!held[y]) continue;
        const int ly = (int)roundf(((float)y / (float)h) * (float)ph);
        for (int d = -1; d <= 1; d++)
            if (ly + d >= 0 && ly + d < ph) prepass_needed[ly + d] = 1;
    }
    sync_all(c);   // frames in flight still read the old tables
    // tables: row_global | local_of_global | cellrow_refine | prepass_needed
    std::vector<int> all;
    all.insert(all.end(), row_global.begin(), row_global.end());
    all.insert(all.end(), local_of_global.begin(), local_of_global.end());
    all.insert(all.end(), cellrow_refine.begin(), cellrow_refine.end());
    all.insert(all.end(), prepass_needed.begin(), prepass_needed.end());
    c->tables.ensure(all.size() * sizeof(int));
    CK(cudaMemcpyAsync(c->tables.p, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    fr.row_global = (const int*)c->tables.p;
    fr.local_of_global = fr.row_global + row_global.size();
    fr.cellrow_refine = fr.local_of_global + local_of_global.size();
    c->prepass_needed = fr.cellrow_refine + cellrow_refine.size();
    c->frame = fr;
    c->stripes.clear();
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        GrtRenderStripe st{};
        st.out_global0 = y0s[i];
        st.out_rows = y1s[i] - y0s[i];
        st.out_local0 = local_of_global[y0s[i]];
        int y = y1s[i];
        while (y < h && held[y] && local_of_global[y] == st.out_local0 + (y - y0s[i])) y++;   // contiguous run
        st.win_local_end = st.out_local0 + (y - y0s[i]);
        st.out_offset_px = off;
        off += (size_t)st.out_rows * w;
        c->stripes.push_back(st);
    }
    c->out_pixels = off;
    c->geom_key = key;
}

// master camera (geodesic snapshot, introspection) on the cold-path stream
void upload_camera(grt_ctx* c, Launcher& L) {
    if (c->camera_on_geodesic || !c->camera_dirty) return;
    CK(cudaMemcpyAsync(c->d_cam, &c->cam, sizeof(GrtCamera), cudaMemcpyHostToDevice, c->stream));
    void* args[] = {&c->d_cam, c->params};
    L.go(c->prog->k_camera, 1, 32, args);
    c->camera_dirty = false;
}
// per-frame camera: uploaded and converted on the frame's own stream (main.cpp:2244-2331); a camera
// attached to a geodesic is copied from the master, which handle_interpolating_geodesic wrote
void frame_camera(grt_ctx* c, FrameSlot& S, Launcher& L) {
    if (c->camera_on_geodesic) {
        // the camera state handle_interpolating_geodesic produced, snapshotted on the host when it was computed: a frame queued
        // now keeps this camera even if the master camera moves along the path before the frame runs (the reference keeps a
        // camera buffer per render_state, render_state.hpp)
        S.cam_staging = c->geo_cam_host;
        CK(cudaMemcpyAsync(S.d_cam, &S.cam_staging, sizeof(GrtCamera), cudaMemcpyHostToDevice, S.stream));
        return;
    }
    CK(cudaMemcpyAsync(S.d_cam, &c->cam, sizeof(GrtCamera), cudaMemcpyHostToDevice, S.stream));
    void* args[] = {&S.d_cam, c->params};
    L.go(c->prog->k_camera, 1, 32, args);
}

void need_program(grt_ctx* c) {
    if (!c->prog || !c->metric) throw Error(GRT_E_STATE, "no metric bound: call grt_set_metric first");
}

GrtRays rays_of(FrameSlot& S, size_t slots, bool need_quat) {
    S.rays_pos.ensure(slots * 16);
    S.rays_vel.ensure(slots * 16);
    S.rays_acc.ensure(slots * 16);
    S.rays_aux.ensure(slots * 16);
    if (need_quat) S.rays_quat.ensure(slots * 16);
    GrtRays r;
    r.pos = (float4*)S.rays_pos.p;
    r.vel = (float4*)S.rays_vel.p;
    r.acc = (float4*)S.rays_acc.p;
    r.quat = (float4*)S.rays_quat.p;
    r.aux = (int4*)S.rays_aux.p;
    return r;
}

cudaEvent_t event_at(FrameSlot& S, size_t i) {
    while (S.events.size() <= i) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        S.events.push_back(e);
    }
    CK(cudaEventRecord(S.events[i], S.stream));
    return S.events[i];
}

// what grt_comm_render asks of a frame: stripes are written at their own rows of a whole frame (`out` is the frame), the stream waits
// for `wait_addr` >= `wait_value` before the first pixel is stored, and `after` runs on the frame's stream behind the render kernels
struct FrameDelivery {
    bool frame_addressed = false;
    unsigned* wait_addr = nullptr;
    unsigned wait_value = 0;
    std::function<void(cudaStream_t)> after;
};

int render_stripes(grt_ctx* c, int w, int h, const int* y0s, const int* y1s, int n, void* out, int format, int out_is_device, grt_stats* stats,
                   const FrameDelivery* delivery = nullptr) {
    bind(c);
    need_program(c);
    if (w <= 0 || h <= 0 || w > 65535 || h > 65535 || n <= 0 || !out) throw Error(GRT_E_INVALID, "bad frame size, stripe list or output pointer");
    if (format != GRT_FORMAT_RGBA32F && format != GRT_FORMAT_RGBA8) throw Error(GRT_E_INVALID, "unknown output format");
    if (!c->tex[0].set) throw Error(GRT_E_STATE, "no background set: call grt_set_background first");
    Program* P = c->prog;
    build_geometry(c, w, h, y0s, y1s, n);
    FrameSlot& S = c->slots[c->frame_index++ % (unsigned)c->n_slots];
    c->last = &S;
    S.used = true;
    Launcher L{S.stream};
    GrtFrame fr = c->frame;
    fill_frame_features(c, fr);
    const bool adaptive = fr.cell == 2;
    const bool use_prepass = (c->feat.use_prepass < 0 ? c->bound_use_prepass : c->feat.use_prepass != 0) && (w / 16) > 0 && (h / 16) > 0;
    const int pw = w / 16, ph = h / 16;
    fr.prepass_w = pw;
    fr.prepass_h = ph;
    fr.use_prepass_mask = use_prepass ? 1 : 0;

    GrtFrame pf = fr;
    if (use_prepass) {
        pf.full_w = pw;
        pf.full_h = ph;
        pf.n_rows = ph;
        pf.cell = 1;
        set_grid(pf);
        const int* ident = identity_table(c, ph);
        pf.row_global = ident;
        pf.local_of_global = ident;
        pf.cellrow_refine = c->prepass_needed;      // the prepass launch reads this table as "row needed" (k_init_rays)
        pf.use_prepass_mask = 0;
        pf.adaptive_sampling = 0;
    }
    const size_t extra_cap = adaptive ? (size_t)3 * fr.grid_w * fr.grid_h : 0;
    const size_t slots = std::max((size_t)fr.n_slots + extra_cap, use_prepass ? (size_t)pf.n_slots : (size_t)0);
    const bool const_theta = c->bound_constant_theta;
    GrtRays rays = rays_of(S, slots, const_theta);
    S.rdata.ensure((size_t)fr.n_rows * w * 16);
    S.termination.ensure(sizeof(int) * (size_t)std::max(1, pw * ph));
    int* d_term = (int*)S.termination.p;
    const size_t px_bytes = format == GRT_FORMAT_RGBA32F ? 16 : 4;
    void* d_out = out;
    if (out_is_device != 1) {
        S.out.ensure(c->out_pixels * px_bytes);
        d_out = S.out.p;
    }

    size_t ev = 0;
    cudaEvent_t e_start = event_at(S, ev++);
    CK(cudaMemsetAsync(S.d_counters, 0, kCounterWords * sizeof(unsigned int) + sizeof(GrtStats), S.stream));
    frame_camera(c, S, L);
    cudaEvent_t e_camera = event_at(S, ev++);

    int* d_extra = (int*)(S.d_counters + 4);
    const int* null_count = nullptr;
    float max_accel = c->feat.max_acceleration_change;
    const unsigned persistent_grid = (unsigned)(c->num_sms * (c->blocks_per_sm > 0 ? std::min(c->blocks_per_sm, P->integrate_blocks_per_sm) : P->integrate_blocks_per_sm));
    // work lists (see GrtWorkList): one int buffer reused by the three phases, which run back to back
    S.worklist.ensure(sizeof(int) * std::max<size_t>({(size_t)fr.n_slots, extra_cap, use_prepass ? (size_t)pf.n_slots : (size_t)1}));
    auto worklist = [&](int phase, size_t capacity) {
        GrtWorkList wl;
        wl.slots = (int*)S.worklist.p;
        wl.n_hot = S.d_counters + 16 + 2 * phase;
        wl.n_cold = S.d_counters + 17 + 2 * phase;
        wl.capacity = (int)capacity;
        return wl;
    };
    GrtWorkList wl_p = worklist(0, use_prepass ? pf.n_slots : 1), wl_f = worklist(1, fr.n_slots), wl_x = worklist(2, std::max<size_t>(extra_cap, 1));
    auto integrate = [&](GrtFrame& f, GrtWorkList& wl, int phase, size_t max_items) {
        unsigned grid = std::min(persistent_grid, (unsigned)((max_items + 127) / 128));
        unsigned int* wc = S.d_counters + 8 + phase;
        void* args[] = {&rays, &f, &wl, &wc, &S.d_stats, &max_accel, c->params};
        if ((int)S.integrate_events.size() <= L.integrate_launches) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            S.integrate_events.push_back({a, b});
        }
        CK(cudaEventRecord(S.integrate_events[L.integrate_launches].first, S.stream));
        L.go(P->k_integrate, grid, 128, args);
        CK(cudaEventRecord(S.integrate_events[L.integrate_launches].second, S.stream));
        L.integrate_launches++;
    };

    if (use_prepass) {
        int one = 1;
        void* a0[] = {&rays, &pf, &S.d_cam, &d_term, &one, &wl_p, c->params};
        L.go(P->k_init_rays, (unsigned)((pf.n_slots + 127) / 128), 128, a0);
        integrate(pf, wl_p, 0, pf.n_slots);
        void* a1[] = {&rays, &pf, &d_term};
        L.go(P->k_prepass_mask, (unsigned)((pf.n_slots + 127) / 128), 128, a1);
    }
    cudaEvent_t e_prepass = event_at(S, ev++);
    {
        int zero = 0;
        void* a0[] = {&rays, &fr, &S.d_cam, &d_term, &zero, &wl_f, c->params};
        L.go(P->k_init_rays, (unsigned)((fr.n_slots + 127) / 128), 128, a0);
    }
    cudaEvent_t e_init = event_at(S, ev++);
    integrate(fr, wl_f, 1, fr.n_slots);
    cudaEvent_t e_int1 = event_at(S, ev++);
    GrtRenderData* d_rdata = (GrtRenderData*)S.rdata.p;
    {
        int begin = 0, fixed = fr.n_slots;
        void* a[] = {&rays, &fr, &begin, &null_count, &fixed, &d_rdata, c->params};
        L.go(P->k_render_data, (unsigned)std::min((fr.n_slots + 127) / 128, c->num_sms * 16), 128, a);
    }
    cudaEvent_t e_rd1 = event_at(S, ev++);
    cudaEvent_t e_adapt = e_rd1, e_int2 = e_rd1, e_rd2 = e_rd1;
    if (adaptive) {
        void* a[] = {&rays, &fr, &S.d_cam, &d_rdata, &d_extra, &wl_x, c->params};
        L.go(P->k_adaptive, (unsigned)((fr.n_slots + 127) / 128), 128, a);
        e_adapt = event_at(S, ev++);
        const int* cnt = d_extra;
        integrate(fr, wl_x, 2, extra_cap);
        e_int2 = event_at(S, ev++);
        int begin = fr.n_slots, fixed = 0;
        void* a2[] = {&rays, &fr, &begin, &cnt, &fixed, &d_rdata, c->params};
        L.go(P->k_render_data, (unsigned)(c->num_sms * 16), 128, a2);
        e_rd2 = event_at(S, ev++);
    }
    const GrtTexture* t1 = &c->tex[0].desc;
    const GrtTexture* t2 = c->tex[1].set ? &c->tex[1].desc : &c->tex[0].desc;
    if (delivery && delivery->wait_addr) CU(g_drv.StreamWaitValue32((CUstream)S.stream, (CUdeviceptr)delivery->wait_addr, delivery->wait_value, CU_STREAM_WAIT_VALUE_GEQ));
    for (auto& st : c->stripes) {
        GrtRenderStripe target = st;
        if (delivery && delivery->frame_addressed) target.out_offset_px = (unsigned long long)st.out_global0 * (unsigned long long)w;
        grt_launch_render(d_rdata, &fr, &target, t1, t2, d_out, format, S.stream);
        L.launches++;
    }
    CK(cudaGetLastError());
    if (delivery && delivery->after) delivery->after(S.stream);
    cudaEvent_t e_render = event_at(S, ev++);
    if (out_is_device != 1) CK(cudaMemcpyAsync(out, d_out, c->out_pixels * px_bytes, cudaMemcpyDeviceToHost, S.stream));
    cudaEvent_t e_end = event_at(S, ev++);

    S.integrate_launches_last = L.integrate_launches;
    S.last_first_pass = fr.n_slots;
    S.last_pw = pw;
    S.last_ph = ph;
    S.last_prepass = use_prepass;
    S.frame = fr;
    if (stats || out_is_device == 0) CK(cudaStreamSynchronize(S.stream));
    if (stats) {
        struct { unsigned int counters[kCounterWords]; GrtStats s; } host;
        CK(cudaMemcpy(&host, S.d_counters, sizeof host, cudaMemcpyDeviceToHost));
        memset(stats, 0, sizeof *stats);
        stats->trial_steps = host.s.trial_steps;
        stats->accepted_steps = host.s.accepted_steps;
        stats->rays_integrated = host.s.rays_integrated;
        stats->rays_finished = host.s.rays_finished;
        stats->extra_rays = host.counters[4];
        S.last_extra = (int)host.counters[4];
        auto ms = [](cudaEvent_t a, cudaEvent_t b) { float t = 0; cudaEventElapsedTime(&t, a, b); return t; };
        stats->ms_camera = ms(e_start, e_camera);
        stats->ms_prepass = ms(e_camera, e_prepass);
        stats->ms_init = ms(e_prepass, e_init);
        stats->ms_integrate = ms(e_init, e_int1) + (adaptive ? ms(e_adapt, e_int2) : 0.0f);
        stats->ms_render_data = ms(e_int1, e_rd1) + (adaptive ? ms(e_int2, e_rd2) : 0.0f);
        stats->ms_adaptive = adaptive ? ms(e_rd1, e_adapt) : 0.0f;
        stats->ms_render = ms(e_rd2, e_render);
        stats->ms_total = ms(e_start, e_end);
        stats->integrate_launches = L.integrate_launches;
        stats->total_launches = L.launches;
        char buf[512];
        snprintf(buf, sizeof buf, "camera:%.4f;prepass:%.4f;init_rays:%.4f;integrate:%.4f;render_data:%.4f;adaptive:%.4f;integrate_extra:%.4f;render_data_extra:%.4f;render:%.4f;copy_out:%.4f;",
                 stats->ms_camera, stats->ms_prepass, stats->ms_init, ms(e_init, e_int1), ms(e_int1, e_rd1), stats->ms_adaptive, adaptive ? ms(e_adapt, e_int2) : 0.0f,
                 adaptive ? ms(e_int2, e_rd2) : 0.0f, stats->ms_render, ms(e_render, e_end));
        c->launch_log = buf;
    } else {
        S.last_extra = -1;   // unknown until read back
    }
    return GRT_OK;
}

void build_texture(grt_ctx* c, Texture& t, const void* src, bool src_is_device, int w, int h) {
    if (w <= 0 || h <= 0 || !src) throw Error(GRT_E_INVALID, "bad background image");
    int max_mips = (int)floorf(log2f((float)std::min(w, h))) + 1;
    int levels = std::min(max_mips, GRT_MAX_MIP_LEVELS);
    GrtTexture d{};
    d.w = w;
    d.h = h;
    d.levels = levels;
    size_t total = 0;
    int lw = w, lh = h;
    for (int l = 0; l < levels; l++) {
        d.lw[l] = lw;
        d.lh[l] = lh;
        d.offset[l] = total;
        total += (size_t)lw * lh;
        lw = std::max(1, lw / 2);
        lh = std::max(1, lh / 2);
    }
    sync_all(c);   // frames in flight still sample the old pyramid
    // GL chain (box filter, round to nearest) in scratch, then the float-readback requantisation
    c->scratch.ensure(total * 4);
    t.data.ensure(total * 4);
    uchar4* gl = (uchar4*)c->scratch.p;
    CK(cudaMemcpyAsync(gl, src, (size_t)w * h * 4, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    for (int l = 1; l < levels; l++) grt_launch_mip_downsample(gl + d.offset[l - 1], d.lw[l - 1], d.lh[l - 1], gl + d.offset[l], d.lw[l], d.lh[l], c->stream);
    grt_launch_mip_requantise(gl, t.data.p, total, c->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    d.data = (const uchar4*)t.data.p;
    t.desc = d;
    t.set = true;
}

void ensure_geodesic(grt_ctx* c) {
    if (c->geo.positions) return;
    const int max_len = 64000;   // main.cpp:1230
    c->geo_pos.ensure((size_t)max_len * 16);
    c->geo_vel.ensure((size_t)max_len * 16);
    c->geo_ds.ensure((size_t)max_len * 4);
    c->geo_tr.ensure((size_t)max_len * 16 * 4);
    c->geo_misc.ensure(64);
    c->geo.positions = (float4*)c->geo_pos.p;
    c->geo.velocities = (float4*)c->geo_vel.p;
    c->geo.ds = (float*)c->geo_ds.p;
    c->geo.transported = (float4*)c->geo_tr.p;
    c->geo.count = (int*)c->geo_misc.p;
    c->geo.max_len = max_len;
    CK(cudaMemsetAsync(c->geo_misc.p, 0, 64, c->stream));
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* grt_last_error(void) { return g_err.c_str(); }
int grt_comm_disconnect(grt_ctx* c);
const char* grt_version(void) { return "geodesic_raytracing_b200 0.1 (sm_100a)"; }

grt_content* grt_content_create(void) { return new grt_content(); }
void grt_content_destroy(grt_content* c) { delete c; }
int grt_content_add_folder(grt_content* c, const char* folder) {
    return guard([&] {
        if (!c || !folder) throw Error(GRT_E_INVALID, "null argument");
        c->cm.add_content_folder(folder);
        c->names = c->cm.metric_names();
        return GRT_OK;
    });
}
int grt_content_metric_count(const grt_content* c) { return c ? (int)c->names.size() : 0; }
const char* grt_content_metric_name(const grt_content* c, int i) { return (c && i >= 0 && i < (int)c->names.size()) ? c->names[i].c_str() : nullptr; }

grt_metric* grt_metric_load(grt_content* c, const char* js_path) {
    grt_metric* out = nullptr;
    int rc = guard([&] {
        if (!c || !js_path) throw Error(GRT_E_INVALID, "null argument");
        try {
            auto m = grt::load_metric_from_script(c->cm, js_path);
            out = new grt_metric();
            out->m = std::move(m);
        } catch (Error&) {
            throw;
        } catch (std::exception& e) {
            throw Error(GRT_E_SCRIPT, e.what());
        }
        return GRT_OK;
    });
    return rc == GRT_OK ? out : nullptr;
}
grt_metric* grt_metric_load_by_name(grt_content* c, const char* name) {
    if (!c || !name) { g_err = "null argument"; return nullptr; }
    std::string path;
    if (!c->cm.lookup_metric_file(name, path)) { g_err = std::string("no metric named ") + name; return nullptr; }
    return grt_metric_load(c, path.c_str());
}
void grt_metric_destroy(grt_metric* m) { delete m; }

int grt_metric_param_count(const grt_metric* m) { return m ? (int)m->m->params.names.size() : 0; }
const char* grt_metric_param_name(const grt_metric* m, int i) { return (m && i >= 0 && i < (int)m->m->params.names.size()) ? m->m->params.names[i].c_str() : nullptr; }
float grt_metric_param_default(const grt_metric* m, int i) { return (m && i >= 0 && i < (int)m->m->params.names.size()) ? m->m->params.default_values[i] : 0.0f; }
float grt_metric_param_get(const grt_metric* m, int i) { return (m && i >= 0 && i < (int)m->m->params.names.size()) ? m->m->params.current_values[i] : 0.0f; }
int grt_metric_param_set(grt_metric* m, int i, float v) {
    if (!m || i < 0 || i >= (int)m->m->params.names.size()) { g_err = "parameter index out of range"; return GRT_E_INVALID; }
    m->m->params.current_values[i] = v;
    return GRT_OK;
}
int grt_metric_get_info(const grt_metric* gm, grt_metric_info* o) {
    if (!gm || !o) { g_err = "null argument"; return GRT_E_INVALID; }
    const grt::Metric& m = *gm->m;
    memset(o, 0, sizeof *o);
    o->is_big = m.desc.big;
    o->is_constant_theta = m.constant_theta();
    o->use_prepass = m.cfg.use_prepass;
    o->singular = m.cfg.singular;
    o->traversable_event_horizon = m.cfg.traversable_event_horizon;
    o->adaptive_precision = m.cfg.adaptive_precision;
    o->detect_singularities = m.cfg.detect_singularities;
    o->coordinate_system = (int)m.cfg.system;
    o->has_periodicity = m.desc.has_period;
    o->unconditionally_nonsingular = m.cfg.unconditionally_nonsingular;
    o->max_acceleration_change = m.cfg.max_acceleration_change;
    o->singular_terminator = m.cfg.singular_terminator;
    o->w_v[0] = 1; o->w_v[1] = 1; o->w_v[2] = 1; o->w_v[3] = 1;
    if (m.cfg.system == grt::CoordinateSystem::X_Y_THETA_PHI) { o->w_v[2] = 8; o->w_v[3] = m.constant_theta() ? 8.0f : 32.0f; }
    else if (m.cfg.system == grt::CoordinateSystem::CYLINDRICAL) o->w_v[2] = 8;
    o->flops_expr = (int)m.desc.flops_expr;
    o->flops_accel = (int)m.desc.flops_accel;
    o->flops_accel_reference = (int)m.desc.flops_accel_ref;
    return GRT_OK;
}
const char* grt_metric_name(const grt_metric* m) { return m ? m->m->cfg.name.c_str() : nullptr; }
const char* grt_metric_description(const grt_metric* m) { return m ? m->m->cfg.description.c_str() : nullptr; }
const char* grt_metric_source(grt_metric* m, int is_static) {
    if (!m) return nullptr;
    // concurrent callers (build() compiles the program and the checker of one metric in parallel) must not see a string being
    // reassigned: emit outside the lock, replace the stored text only when it really changed
    int rc = guard([&] {
        std::string text = grt::emit_metric_source(*m->m, is_static != 0);
        std::lock_guard<std::mutex> lock(m->text_mutex);
        if (m->source[is_static ? 1 : 0] != text) m->source[is_static ? 1 : 0] = std::move(text);
        return GRT_OK;
    });
    return rc == GRT_OK ? m->source[is_static ? 1 : 0].c_str() : nullptr;
}

int grt_metric_packed_params(const grt_metric* m, float* out, int max_count) {
    return guard([&] {
        if (!m) throw Error(GRT_E_INVALID, "null metric");
        std::vector<float> packed = grt::pack_params(*m->m);
        if (out)
            for (int i = 0; i < (int)packed.size() && i < max_count; i++) out[i] = packed[i];
        return (int)packed.size();
    });
}

uint8_t* grt_image_load_png(const char* path, int* width, int* height) {
    uint8_t* out = nullptr;
    int rc = guard([&] {
        if (!path || !width || !height) throw Error(GRT_E_INVALID, "null argument");
        std::vector<uint8_t> px;
        try { px = grt::load_png_rgba(path, width, height); } catch (const std::runtime_error& e) { throw Error(GRT_E_INVALID, e.what()); }
        out = (uint8_t*)malloc(px.size());
        if (!out) throw Error(GRT_E_INVALID, "out of memory");
        memcpy(out, px.data(), px.size());
        return GRT_OK;
    });
    return rc == GRT_OK ? out : nullptr;
}
void grt_image_free(uint8_t* pixels) { free(pixels); }
int grt_image_save_png(const char* path, const uint8_t* rgba, int width, int height) {
    return guard([&] {
        if (!path || !rgba) throw Error(GRT_E_INVALID, "null argument");
        try { grt::save_png_rgba(path, rgba, width, height); } catch (const std::runtime_error& e) { throw Error(GRT_E_INVALID, e.what()); }
        return GRT_OK;
    });
}

int grt_metric_eval_symbolic(const grt_metric* m, const double pos[4], const double vel[4], double* g16, double* dg64, double* accel4, double* accel_ref4,
                             double* to_polar4, double* from_polar4, double* distance1) {
    return guard([&] {
        if (!m || !pos) throw Error(GRT_E_INVALID, "null metric or position");
        const grt::Metric& M = *m->m;
        const grt::MetricDescriptor& d = M.desc;
        std::map<std::string, double> env;
        const char* vn[4] = {"v1", "v2", "v3", "v4"};
        const char* ivn[4] = {"iv1", "iv2", "iv3", "iv4"};
        const char* dvn[4] = {"dv1", "dv2", "dv3", "dv4"};
        for (int i = 0; i < 4; i++) { env[vn[i]] = pos[i]; env[ivn[i]] = vel ? vel[i] : 0.0; env[dvn[i]] = vel ? vel[i] : 0.0; }
        for (size_t i = 0; i < M.params.names.size(); i++) env[M.params.names[i]] = M.params.current_values[i];
        auto put = [&](double* out, const int* ids, int n) { if (out) for (int i = 0; i < n; i++) out[i] = d.graph.eval(ids[i], env); };
        put(g16, d.g, 16);
        put(dg64, d.dg, 64);
        put(accel4, d.accel, 4);
        put(accel_ref4, d.accel_ref, 4);
        put(to_polar4, d.to_polar, 4);
        put(from_polar4, d.from_polar, 4);
        put(distance1, &d.distance, 1);
        return GRT_OK;
    });
}

int grt_compile_metric(grt_metric* m, int is_static, int fast_math, const char** log_out) {
    return guard([&] {
        if (!m) throw Error(GRT_E_INVALID, "null metric");
        std::string stem = m->m->js_path;
        size_t k = stem.find_last_of('/');
        if (k != std::string::npos) stem = stem.substr(k + 1);
        k = stem.find_last_of('.');
        if (k != std::string::npos) stem = stem.substr(0, k);
        std::string log;
        compile_cubin(program_source(*m->m, is_static != 0, false, fast_math != 0), fast_math != 0, stem + (is_static ? "_static" : "_dynamic"), &log, nullptr);
        std::lock_guard<std::mutex> lock(m->text_mutex);
        if (m->compile_log != log) m->compile_log = std::move(log);
        if (log_out) *log_out = m->compile_log.c_str();
        return GRT_OK;
    });
}

void grt_features_default(grt_features* f) { if (f) features_default(f); }

grt_ctx* grt_create(int device) {
    grt_ctx* c = nullptr;
    int rc = guard([&] {
        // Frames in flight live on one stream each, and streams beyond the device's hardware work queues (8 by default) are
        // multiplexed onto shared queues, which serialises frames that should overlap: 12 - 16 slots only pay with more queues
        // (one-GPU emulation of rank 0 of 8: 3.89 -> 3.44 ms per frame).  Read by the driver when it creates the device's context,
        // so it has to be in the environment before the first CUDA call of the process (a host that initialises CUDA earlier, e.g.
        // through PyTorch, sets it itself: bench.py does).  Never overrides the user's value.
        setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) throw Error(GRT_E_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
        if (device < 0 || device >= count) throw Error(GRT_E_INVALID, "device index out of range");
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) throw Error(GRT_E_CUDA, std::string("device is not sm_100 (Blackwell B200): ") + prop.name);
        CK(cudaSetDevice(device));
        CK(cudaFree(0));
        load_driver();
        c = new grt_ctx();
        c->device = device;
        c->num_sms = prop.multiProcessorCount;
        features_default(&c->feat);
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CK(cudaMalloc((void**)&c->d_cam, sizeof(GrtCamera)));
        if (const char* e = getenv("GRT_FRAME_SLOTS")) c->n_slots = std::max(1, std::min(16, atoi(e)));
        for (auto& S : c->slots) {
            CK(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
            CK(cudaMalloc((void**)&S.d_cam, sizeof(GrtCamera)));
            void* counters = nullptr;
            CK(cudaMalloc(&counters, kCounterWords * sizeof(unsigned int) + sizeof(GrtStats)));
            S.d_counters = (unsigned int*)counters;
            S.d_stats = (GrtStats*)(S.d_counters + kCounterWords);
        }
        c->last = &c->slots[0];
        if (const char* e = getenv("GRT_BLOCKS_PER_SM")) c->blocks_per_sm = atoi(e);
        // default camera (main.cpp:664-673)
        c->cam.pos_cart = make_float4(0, 0, -4, 0);
        c->cam.quat = make_float4(-0.70710678f, 0, 0, 0.70710678f);
        c->cam.basis_speed = make_float4(0, 0, 0, 0);
        return GRT_OK;
    });
    if (rc != GRT_OK) { delete c; return nullptr; }
    return c;
}

void grt_destroy(grt_ctx* c) {
    if (!c) return;
    if (c->build.active) c->build.cubin.wait();
    cudaSetDevice(c->device);
    if (c->comm.active) {      // while the slot streams still exist: disconnect waits for the frames in flight
        grt_comm_disconnect(c);
        c->comm.local.release();
    }
    for (auto& S : c->slots) if (S.stream) cudaStreamSynchronize(S.stream);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->programs) if (kv.second->mod && g_drv.ModuleUnload) g_drv.ModuleUnload(kv.second->mod);
    for (auto& S : c->slots) {
        for (DevBuf* b : {&S.rays_pos, &S.rays_vel, &S.rays_acc, &S.rays_quat, &S.rays_aux, &S.worklist, &S.rdata, &S.termination, &S.out}) b->release();
        for (auto e : S.events) cudaEventDestroy(e);
        for (auto& e : S.integrate_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        cudaFree(S.d_cam);
        cudaFree(S.d_counters);
        if (S.stream) cudaStreamDestroy(S.stream);
    }
    for (DevBuf* b : {&c->tables, &c->identity, &c->scratch, &c->tex[0].data, &c->tex[1].data, &c->geo_pos, &c->geo_vel, &c->geo_ds, &c->geo_tr, &c->geo_misc}) b->release();
    cudaFree(c->d_cam);
    cudaStreamDestroy(c->stream);
    delete c;
}

int grt_set_fast_math(grt_ctx* c, int enabled) {
    if (!c) { g_err = "null context"; return GRT_E_INVALID; }
    c->fast_math = enabled != 0;
    return GRT_OK;
}

int grt_set_tuning(grt_ctx* c, int blocks_per_sm, int donate_threshold) {
    if (!c || blocks_per_sm < 0 || blocks_per_sm > 32 || donate_threshold < 0 || donate_threshold > 31) { g_err = "tuning out of range (blocks_per_sm 0..32, donate_threshold 0..31)"; return GRT_E_INVALID; }
    c->blocks_per_sm = blocks_per_sm;
    return GRT_OK;
}

int grt_set_frames_in_flight(grt_ctx* c, int n) {
    return guard([&] {
        if (!c || n < 1 || n > 16) throw Error(GRT_E_INVALID, "frames in flight must be 1..16");
        bind(c);
        sync_all(c);
        c->n_slots = n;
        c->frame_index = 0;
        return GRT_OK;
    });
}

int grt_set_metric_async(grt_ctx* c, grt_metric* m, int is_static) {
    return guard([&] {
        if (!c || !m) throw Error(GRT_E_INVALID, "null argument");
        if (m->m->params.names.size() > 64) throw Error(GRT_E_INVALID, "too many $cfg parameters (max 64)");
        if (c->build.active) {   // a newer request supersedes the one in flight (metric_manager.hpp:172-219 cancels it)
            c->build.cubin.wait();
            c->build.active = false;
        }
        std::string src = program_source(*m->m, is_static != 0, c->feat.reparameterisation != 0, c->fast_math);
        c->build.key = program_key(c, src);
        c->build.metric = m;
        c->build.is_static = is_static != 0;
        c->build.log = std::make_shared<std::string>();
        const bool fast = c->fast_math;
        const std::string stem = program_stem(*m->m, is_static != 0);
        auto log = c->build.log;
        c->build.cubin = std::async(std::launch::async, [src, fast, stem, log]() { return compile_cubin(src, fast, stem, log.get(), nullptr); });
        c->build.active = true;
        return GRT_OK;
    });
}

static void fill_params(grt_ctx* c, const grt::Metric& m) {
    std::vector<float> packed = grt::pack_params(m);
    if (packed.size() > sizeof(c->params) / sizeof(float)) throw Error(GRT_E_INVALID, "parameter block too large (named + derived > 256)");
    memset(c->params, 0, sizeof c->params);
    if (!packed.empty()) memcpy(c->params, packed.data(), packed.size() * sizeof(float));
}

int grt_build_poll(grt_ctx* c) {
    return guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        if (!c->build.active) return 1;
        if (c->build.cubin.wait_for(std::chrono::seconds(0)) != std::future_status::ready) return 0;
        c->build.active = false;
        std::vector<char> cubin = c->build.cubin.get();   // rethrows a compile error as GRT_E_COMPILE
        bind(c);
        grt_metric* m = c->build.metric;
        {
            std::lock_guard<std::mutex> lock(m->text_mutex);
            m->compile_log = *c->build.log;
        }
        // the program was specialised on the features and the fast-math setting of the REQUEST: if they changed meanwhile
        // (grt_set_features / grt_set_fast_math between request and poll) the finished cubin is not the program to bind
        Program* p = nullptr;
        if (program_key(c, program_source(*m->m, c->build.is_static, c->feat.reparameterisation != 0, c->fast_math)) != c->build.key) p = get_program(c, m, c->build.is_static);
        else {
            auto it = c->programs.find(c->build.key);
            p = it != c->programs.end() ? it->second.get() : load_program(c, c->build.key, cubin);
        }
        retire_static_program(c, p, c->build.is_static);
        // swap: frames already queued keep the module they were launched with
        c->metric = m;
        c->is_static = c->build.is_static;
        c->bound_use_prepass = m->m->cfg.use_prepass;
        c->bound_constant_theta = m->m->constant_theta();
        { double w[4]; grt::acceleration_weights(*m->m, w); c->bound_wmax = (float)std::max(std::max(w[0], w[1]), std::max(w[2], w[3])); }
        c->prog = p;
        c->kernel_info = p->info;
        fill_params(c, *m->m);
        c->feat.max_acceleration_change = m->m->cfg.max_acceleration_change;
        c->camera_dirty = true;
        return 1;
    });
}

int grt_update_params(grt_ctx* c) {
    return guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        need_program(c);
        const auto& p = c->metric->m->params;
        if (p.names.size() > 64) throw Error(GRT_E_INVALID, "too many $cfg parameters (max 64)");
        fill_params(c, *c->metric->m);
        if (c->is_static) {   // literals changed: rebuild (metric_manager.hpp:153-167); the reference replaces its single static program
            Program* p = get_program(c, c->metric, true);
            retire_static_program(c, p, true);
            c->prog = p;
        }
        c->kernel_info = c->prog->info;
        c->camera_dirty = true;
        return GRT_OK;
    });
}

int grt_set_metric(grt_ctx* c, grt_metric* m, int is_static) {
    return guard([&] {
        if (!c || !m) throw Error(GRT_E_INVALID, "null argument");
        bind(c);
        if (m->m->params.names.size() > 64) throw Error(GRT_E_INVALID, "too many $cfg parameters (max 64)");
        Program* p = get_program(c, m, is_static != 0);
        retire_static_program(c, p, is_static != 0);
        c->metric = m;
        c->is_static = is_static != 0;
        c->bound_use_prepass = m->m->cfg.use_prepass;
        c->bound_constant_theta = m->m->constant_theta();
        { double w[4]; grt::acceleration_weights(*m->m, w); c->bound_wmax = (float)std::max(std::max(w[0], w[1]), std::max(w[2], w[3])); }
        c->prog = p;
        c->kernel_info = p->info;
        fill_params(c, *m->m);
        c->feat.max_acceleration_change = m->m->cfg.max_acceleration_change;   // metric_manager.hpp:50
        c->camera_dirty = true;
        c->camera_on_geodesic = false;
        c->geo_count = 0;
        return GRT_OK;
    });
}

int grt_set_features(grt_ctx* c, const grt_features* f) {
    return guard([&] {
        if (!c || !f) throw Error(GRT_E_INVALID, "null argument");
        const bool rebuild = c->prog && (c->feat.reparameterisation != 0) != (f->reparameterisation != 0);
        c->feat = *f;
        if (rebuild) {   // a feature the program is specialised on changed (soft recompile, metric_manager.hpp:153-167)
            bind(c);
            c->prog = get_program(c, c->metric, c->is_static);
            c->kernel_info = c->prog->info;
        }
        return GRT_OK;
    });
}
int grt_get_features(const grt_ctx* c, grt_features* f) {
    if (!c || !f) { g_err = "null argument"; return GRT_E_INVALID; }
    *f = c->feat;
    return GRT_OK;
}

int grt_set_background(grt_ctx* c, const uint8_t* rgba, int w, int h, int which) {
    return guard([&] {
        if (!c || which < 0 || which > 1) throw Error(GRT_E_INVALID, "bad argument");
        bind(c);
        build_texture(c, c->tex[which], rgba, false, w, h);
        return GRT_OK;
    });
}
int grt_set_background_device(grt_ctx* c, const void* rgba, int w, int h, int which) {
    return guard([&] {
        if (!c || which < 0 || which > 1) throw Error(GRT_E_INVALID, "bad argument");
        bind(c);
        build_texture(c, c->tex[which], rgba, true, w, h);
        return GRT_OK;
    });
}

int grt_set_camera(grt_ctx* c, const float pos[4], const float quat[4], int flip_r) {
    if (!c || !pos || !quat) { g_err = "null argument"; return GRT_E_INVALID; }
    c->cam.pos_cart = make_float4(pos[0], pos[1], pos[2], pos[3]);
    c->cam.quat = make_float4(quat[0], quat[1], quat[2], quat[3]);
    c->cam.flip = flip_r ? 1.0f : 0.0f;
    c->camera_dirty = true;
    return GRT_OK;
}
int grt_set_observer_speed(grt_ctx* c, const float s[3]) {
    if (!c || !s) { g_err = "null argument"; return GRT_E_INVALID; }
    c->cam.basis_speed = make_float4(s[0], s[1], s[2], 0);
    c->camera_dirty = true;
    return GRT_OK;
}

int grt_snapshot_geodesic(grt_ctx* c, float linear_basis_speed, int parallel_transport_observer) {
    int count = 0;
    int rc = guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        need_program(c);
        Launcher L{c->stream};
        // the camera tetrad at the snapshot point (init_basis_vectors, main.cpp:2314-2331)
        c->camera_on_geodesic = false;
        c->camera_dirty = true;
        upload_camera(c, L);
        c->geo_parallel_transport = parallel_transport_observer != 0;
        // observer 3-velocity = speed * camera forward (main.cpp:2253-2262): forward = rotate (0,0,1) by the camera quaternion
        float4 q = c->cam.quat;
        float qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        q.x /= qn; q.y /= qn; q.z /= qn; q.w /= qn;
        float3 fwd = make_float3(2 * (q.x * q.z + q.w * q.y), 2 * (q.y * q.z - q.w * q.x), 1 - 2 * (q.x * q.x + q.y * q.y));
        float4 speed = make_float4(fwd.x * linear_basis_speed, fwd.y * linear_basis_speed, fwd.z * linear_basis_speed, 0);
        c->geo_speed = speed;
        CK(cudaMemcpyAsync((char*)c->d_cam + offsetof(GrtCamera, basis_speed), &c->geo_speed, sizeof speed, cudaMemcpyHostToDevice, c->stream));
        ensure_geodesic(c);
        GrtFrame fr{};
        fill_frame_features(c, fr);
        float max_accel = c->feat.max_acceleration_change;
        void* a0[] = {&c->d_cam, &c->geo, &fr, &max_accel, c->params};
        L.go(c->prog->k_geodesic_snapshot, 1, 32, a0);
        void* a1[] = {&c->d_cam, &c->geo, c->params};
        L.go(c->prog->k_geodesic_transport, 1, 32, a1);
        CK(cudaMemcpyAsync(&count, c->geo.count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->geo_count = count;
        c->camera_dirty = true;
        return GRT_OK;
    });
    return rc == GRT_OK ? count : rc;
}

int grt_set_camera_on_geodesic(grt_ctx* c, float proper_time) {
    return guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        need_program(c);
        if (c->geo_count <= 0) throw Error(GRT_E_STATE, "no geodesic snapshot: call grt_snapshot_geodesic first");
        Launcher L{c->stream};
        // quat / speeds come from the host camera, position + tetrad from the path
        CK(cudaMemcpyAsync((char*)c->d_cam + offsetof(GrtCamera, quat), &c->cam.quat, sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync((char*)c->d_cam + offsetof(GrtCamera, basis_speed), &c->geo_speed, sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        int pt = c->geo_parallel_transport;
        float4* vel_out = nullptr;
        void* a[] = {&c->d_cam, &c->geo, &proper_time, &pt, &vel_out, c->params};
        L.go(c->prog->k_geodesic_interpolate, 1, 32, a);
        CK(cudaMemcpyAsync(&c->geo_cam_host, c->d_cam, sizeof(GrtCamera), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));   // frames upload this host copy on their own streams (frame_camera)
        c->camera_on_geodesic = true;
        return GRT_OK;
    });
}
int grt_clear_camera_geodesic(grt_ctx* c) {
    if (!c) { g_err = "null context"; return GRT_E_INVALID; }
    c->camera_on_geodesic = false;
    c->camera_dirty = true;
    return GRT_OK;
}

int grt_render_stripes(grt_ctx* c, int w, int h, const int* y0s, const int* y1s, int n, void* out, int format, int out_is_device, grt_stats* stats) {
    return guard([&] {
        if (!c || !y0s || !y1s) throw Error(GRT_E_INVALID, "null argument");
        return render_stripes(c, w, h, y0s, y1s, n, out, format, out_is_device, stats);
    });
}
int grt_render_rows(grt_ctx* c, int w, int h, int y0, int y1, void* out, int format, int out_is_device, grt_stats* stats) {
    return grt_render_stripes(c, w, h, &y0, &y1, 1, out, format, out_is_device, stats);
}
int grt_render(grt_ctx* c, int w, int h, void* out, int format, int out_is_device, grt_stats* stats) {
    int y0 = 0, y1 = h;
    return grt_render_stripes(c, w, h, &y0, &y1, 1, out, format, out_is_device, stats);
}
int grt_synchronize(grt_ctx* c) {
    return guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        sync_all(c);
        return GRT_OK;
    });
}
void* grt_stream(grt_ctx* c) { return c ? (void*)c->last->stream : nullptr; }


// ---- multi-GPU frame delivery ---------------------------------------------------------------------------
namespace {
constexpr size_t kCommReady = 0, kCommFreed = 4096, kCommRing = 8192, kCommFrames = 12288;

unsigned* comm_ready(grt_ctx* c, int of_rank, int frame, int from_rank) { return (unsigned*)(c->comm.base[of_rank] + kCommReady) + (size_t)frame * c->comm.world + from_rank; }
unsigned* comm_freed(grt_ctx* c, int of_rank, int frame) { return (unsigned*)(c->comm.base[of_rank] + kCommFreed) + frame; }

// `value` -> *dst on `stream` without a kernel: a stream write-value (preceded by a system-wide fence, so the render kernel's peer
// stores are visible before the word changes).  The word may live in another rank's region: the peer-mapped address is written
// directly.  Measured on 2 B200s: 0.9 ms per frame faster than staging the word locally and moving it with a 4-byte
// cudaMemcpyAsync (GRT_COMM_DMA_SIGNAL=1 selects that path; profiles/round2_scaling.md).
void comm_post(grt_ctx* c, cudaStream_t stream, unsigned* dst, bool remote, unsigned value) {
    static const bool dma = getenv("GRT_COMM_DMA_SIGNAL") && atoi(getenv("GRT_COMM_DMA_SIGNAL")) != 0;
    if (!remote || !dma) {
        CU(g_drv.StreamWriteValue32((CUstream)stream, (CUdeviceptr)dst, value, CU_STREAM_WRITE_VALUE_DEFAULT));
        return;
    }
    unsigned* src = (unsigned*)(c->comm.base[c->comm.rank] + kCommRing) + (c->comm.ring_pos++ & 1023u);
    CU(g_drv.StreamWriteValue32((CUstream)stream, (CUdeviceptr)src, value, CU_STREAM_WRITE_VALUE_DEFAULT));
    CK(cudaMemcpyAsync(dst, src, sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
}
}  // namespace

int grt_comm_init(grt_ctx* c, int rank, int world, int width, int height, int format, int n_frames, uint8_t* handle_out) {
    return guard([&] {
        if (!c || world < 1 || rank < 0 || rank >= world || width <= 0 || height <= 0 || n_frames < 1) throw Error(GRT_E_INVALID, "bad communicator shape");
        if (format != GRT_FORMAT_RGBA32F && format != GRT_FORMAT_RGBA8) throw Error(GRT_E_INVALID, "unknown output format");
        if ((size_t)n_frames * world * sizeof(unsigned) > 4096 || (size_t)n_frames * sizeof(unsigned) > 4096) throw Error(GRT_E_INVALID, "too many frame buffers x ranks (n_frames * world <= 1024)");
        if (c->comm.active) throw Error(GRT_E_STATE, "communicator already initialised");
        bind(c);
        auto& m = c->comm;
        m.rank = rank; m.world = world; m.n_frames = n_frames; m.w = width; m.h = height; m.format = format;
        m.frame_bytes = ((size_t)width * height * (format == GRT_FORMAT_RGBA32F ? 16 : 4) + 255) / 256 * 256;
        const size_t bytes = kCommFrames + (rank == 0 ? m.frame_bytes * (size_t)n_frames : 0);
        CK(cudaMalloc(&m.local.p, bytes));      // exactly this allocation is exported: not through DevBuf::ensure's rounding
        m.local.cap = bytes;
        CK(cudaMemset(m.local.p, 0, bytes));
        m.base.assign(world, nullptr);
        m.base[rank] = (char*)m.local.p;
        if (handle_out) {
            cudaIpcMemHandle_t hnd;
            CK(cudaIpcGetMemHandle(&hnd, m.local.p));
            static_assert(sizeof(hnd) == GRT_COMM_HANDLE_BYTES, "IPC handle size");
            memcpy(handle_out, &hnd, sizeof hnd);
        }
        m.active = true;
        m.connected = world == 1;
        return GRT_OK;
    });
}

int grt_comm_connect(grt_ctx* c, const uint8_t* handles) {
    return guard([&] {
        if (!c || !handles) throw Error(GRT_E_INVALID, "null argument");
        auto& m = c->comm;
        if (!m.active) throw Error(GRT_E_STATE, "grt_comm_init first");
        bind(c);
        for (int r = 0; r < m.world; r++) {
            if (r == m.rank || m.base[r]) continue;
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, handles + (size_t)r * GRT_COMM_HANDLE_BYTES, sizeof hnd);
            void* p = nullptr;
            CK(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
            m.base[r] = (char*)p;
        }
        m.ipc = true;
        m.connected = true;
        return GRT_OK;
    });
}

int grt_comm_connect_local(grt_ctx** ctxs, int n) {
    return guard([&] {
        if (!ctxs || n < 1) throw Error(GRT_E_INVALID, "null argument");
        for (int i = 0; i < n; i++)
            if (!ctxs[i] || !ctxs[i]->comm.active || ctxs[i]->comm.world != n || ctxs[i]->comm.rank != i) throw Error(GRT_E_STATE, "contexts must be initialised with grt_comm_init(rank = index, world = n)");
        for (int i = 0; i < n; i++) {
            bind(ctxs[i]);
            for (int j = 0; j < n; j++) {
                if (i == j) continue;
                if (ctxs[i]->device != ctxs[j]->device) {
                    int can = 0;
                    CK(cudaDeviceCanAccessPeer(&can, ctxs[i]->device, ctxs[j]->device));
                    if (!can) throw Error(GRT_E_CUDA, "devices cannot access each other's memory");
                    cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else CK(e);
                }
                ctxs[i]->comm.base[j] = (char*)ctxs[j]->comm.local.p;
            }
            ctxs[i]->comm.connected = true;
        }
        return GRT_OK;
    });
}

int grt_comm_disconnect(grt_ctx* c) {
    if (!c) { g_err = "null context"; return GRT_E_INVALID; }
    auto& m = c->comm;
    if (!m.active) return GRT_OK;
    cudaSetDevice(c->device);
    for (auto& S : c->slots) if (S.stream) cudaStreamSynchronize(S.stream);
    for (int r = 0; r < (int)m.base.size(); r++) {
        if (r == m.rank || !m.base[r]) continue;
        if (m.ipc) cudaIpcCloseMemHandle(m.base[r]);
        m.base[r] = nullptr;
    }
    m.connected = m.world == 1;
    return GRT_OK;
}

int grt_interleaved_stripes(int height, int rank, int world, int stripe_rows, int* y0s, int* y1s, int max_stripes) {
    if (height <= 0 || world < 1 || rank < 0 || rank >= world || !y0s || !y1s || max_stripes < 1) { g_err = "bad stripe request"; return GRT_E_INVALID; }
    if (world == 1) { y0s[0] = 0; y1s[0] = height; return 1; }
    stripe_rows = std::max(2, stripe_rows - (stripe_rows & 1));      // even: adaptive sampling works on 2x2 cells
    // Stripe i belongs to rank (i + 3 * (i / world)) mod world: within every group of `world` consecutive stripes each rank gets one,
    // and the position it gets rotates by three from group to group, so no rank collects the same kind of rows (all-sky stripes at the
    // top of every group, say) - measured at 8 GPUs, plain round-robin left ranks 0 and 7 with 10 % more work than the rest.
    int n = 0, i = 0;
    for (int y = 0; y < height; y += stripe_rows, i++) {
        if ((i + 3 * (i / world)) % world != rank) continue;
        if (n == max_stripes) { g_err = "more stripes than max_stripes"; return GRT_E_INVALID; }
        y0s[n] = y;
        y1s[n] = std::min(height, y + stripe_rows);
        n++;
    }
    return n;
}

namespace {
// rows [y0, y1) of stripe i and of its two halves (even heights: adaptive sampling works on 2x2 cells)
struct StripeRows { int y0, mid, y1; };
StripeRows stripe_rows_of(int i, int stripe_rows, int height) {
    const int y0 = i * stripe_rows, y1 = std::min(height, y0 + stripe_rows);
    int mid = y0 + ((y1 - y0) / 2 & ~1);
    if (mid <= y0 || mid >= y1) mid = y1;          // too short to split
    return {y0, mid, y1};
}
}  // namespace

int grt_measure_stripe_costs(grt_ctx* c, int width, int height, int stripe_rows, double* cost_out, int max_costs) {
    if (!c || !cost_out || width <= 0 || height <= 0 || max_costs < 1) { g_err = "bad stripe cost request"; return GRT_E_INVALID; }
    stripe_rows = std::max(2, stripe_rows - (stripe_rows & 1));
    const int n = (height + stripe_rows - 1) / stripe_rows;
    if (2 * n > max_costs) { g_err = "more half stripes than max_costs"; return GRT_E_INVALID; }
    DevBuf scratch;
    int rc = guard([&] {
        bind(c);
        scratch.ensure((size_t)stripe_rows * (size_t)width * 4);
        return GRT_OK;
    });
    if (rc != GRT_OK) return rc;
    for (int i = 0; i < n && rc == GRT_OK; i++) {
        const StripeRows r = stripe_rows_of(i, stripe_rows, height);
        const int y0s[2] = {r.y0, r.mid}, y1s[2] = {r.mid, r.y1};
        for (int half = 0; half < 2 && rc == GRT_OK; half++) {
            cost_out[2 * i + half] = 0.0;
            if (y1s[half] <= y0s[half]) continue;
            grt_stats st{};
            rc = grt_render_stripes(c, width, height, &y0s[half], &y1s[half], 1, scratch.p, GRT_FORMAT_RGBA8, GRT_OUT_DEVICE, &st);
            cost_out[2 * i + half] = (double)st.trial_steps;
        }
    }
    scratch.release();
    return rc == GRT_OK ? 2 * n : rc;
}

int grt_balanced_stripes(int height, int rank, int world, int stripe_rows, const double* cost, int n_cost, int* y0s, int* y1s, int max_stripes) {
    if (!cost) return grt_interleaved_stripes(height, rank, world, stripe_rows, y0s, y1s, max_stripes);
    if (height <= 0 || world < 1 || rank < 0 || rank >= world || !y0s || !y1s || max_stripes < 1) { g_err = "bad stripe request"; return GRT_E_INVALID; }
    if (world == 1) { y0s[0] = 0; y1s[0] = height; return 1; }
    stripe_rows = std::max(2, stripe_rows - (stripe_rows & 1));
    const int n = (height + stripe_rows - 1) / stripe_rows;
    if (n_cost != 2 * n) { g_err = "stripe cost vector does not match the stripe count of this height (two halves per stripe)"; return GRT_E_INVALID; }
    // Items: whole stripes, except that a stripe heavier than 3/10 of a rank's fair share goes in as its two halves (the rows tangent to
    // the shadow's edge: two such stripes of a 4K kerr_schild frame carry 150 M of the 240 M ray-steps a rank of eight should get).  A
    // half rendered alone pays its own halo; a whole stripe pays one halo for both, hence the discount on the sum.
    struct Item { int y0, y1; double cost; };
    std::vector<Item> items;
    double total = 0.0;
    for (int i = 0; i < 2 * n; i++) total += std::max(0.0, cost[i]);
    const double fair = total / world, halo = 6.0;
    for (int i = 0; i < n; i++) {
        const StripeRows r = stripe_rows_of(i, stripe_rows, height);
        const double top = std::max(0.0, cost[2 * i]), bottom = std::max(0.0, cost[2 * i + 1]);
        if (r.mid < r.y1 && top + bottom > 0.3 * fair) {
            items.push_back({r.y0, r.mid, top});
            items.push_back({r.mid, r.y1, bottom});
        } else {
            const double rows = (double)(r.y1 - r.y0);
            items.push_back({r.y0, r.y1, (top + bottom) * (r.mid < r.y1 ? (rows + halo) / (rows + 2.0 * halo) : 1.0)});
        }
    }
    const int m = (int)items.size();
    std::vector<int> order(m), owner(m, 0), held(world, 0);
    for (int i = 0; i < m; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return items[a].cost > items[b].cost; });
    std::vector<double> load(world, 0.0);
    // Nearly the same number of items on every rank: what an item costs on top of its rays - halo rows, its share of the prepass,
    // launches - is then about the same everywhere and drops out of the comparison.
    const int cap = (m + world - 1) / world + 1;
    for (int i : order) {
        int best = -1;
        for (int r = 0; r < world; r++)
            if (held[r] < cap && (best < 0 || load[r] < load[best])) best = r;
        owner[i] = best;
        load[best] += items[i].cost;
        held[best]++;
    }
    // refinement: exchange one item of the heaviest rank against one item of another rank while that lowers the larger of the two
    // loads (the counts stay as they are); deterministic, a few dozen exchanges at most for the stripes of a 4K frame
    for (int round = 0; round < 4 * m; round++) {
        int p = 0;
        for (int r = 1; r < world; r++) if (load[r] > load[p]) p = r;
        double best_peak = load[p];
        int best_a = -1, best_b = -1;
        for (int a = 0; a < m; a++) {
            if (owner[a] != p) continue;
            for (int b = 0; b < m; b++) {
                const int q = owner[b];
                if (q == p) continue;
                const double peak = std::max(load[p] - items[a].cost + items[b].cost, load[q] - items[b].cost + items[a].cost);
                if (peak < best_peak - 1e-9 * std::max(1.0, load[p])) { best_peak = peak; best_a = a; best_b = b; }
            }
        }
        if (best_a < 0) break;
        const int q = owner[best_b];
        load[p] += items[best_b].cost - items[best_a].cost;
        load[q] += items[best_a].cost - items[best_b].cost;
        owner[best_a] = q;
        owner[best_b] = p;
    }
    // this rank's items in ascending order, neighbours merged into one stripe (one halo less)
    int k = 0;
    for (int i = 0; i < m; i++) {
        if (owner[i] != rank) continue;
        if (k > 0 && y1s[k - 1] == items[i].y0) { y1s[k - 1] = items[i].y1; continue; }
        if (k == max_stripes) { g_err = "more stripes than max_stripes"; return GRT_E_INVALID; }
        y0s[k] = items[i].y0;
        y1s[k] = items[i].y1;
        k++;
    }
    return k;
}

// Contiguous blocks of rows with equal cost.  `density` is a piecewise-constant cost per row: piece k covers rows [edge[k], edge[k + 1])
// and costs cost[k] in all.  Boundaries go where the cumulative cost crosses r / world of the total, rounded to even rows, every
// block at least two rows tall.
namespace {
void equal_cost_bounds(int height, int world, const std::vector<int>& edge, const std::vector<double>& cost, int* bounds) {
    const int pieces = (int)cost.size();
    double total = 0.0;
    for (double c : cost) total += std::max(0.0, c);
    bounds[0] = 0;
    bounds[world] = height;
    int k = 0;
    double before = 0.0;       // cost of the pieces in front of piece k
    for (int r = 1; r < world; r++) {
        const double want = total * r / world;
        while (k < pieces - 1 && before + std::max(0.0, cost[k]) < want) before += std::max(0.0, cost[k++]);
        const double ck = std::max(0.0, cost[k]);
        const double frac = ck > 0.0 ? std::min(1.0, std::max(0.0, (want - before) / ck)) : 0.0;
        int y = (int)std::lround(edge[k] + frac * (edge[k + 1] - edge[k]));
        y &= ~1;
        y = std::max(y, bounds[r - 1] + 2);
        y = std::min(y, height - 2 * (world - r) - (height & 1));
        bounds[r] = y;
    }
}
}  // namespace

int grt_contiguous_bounds(int height, int world, int stripe_rows, const double* cost, int n_cost, int* bounds_out) {
    if (height < 2 * world || world < 1 || !bounds_out) { g_err = "bad partition request"; return GRT_E_INVALID; }
    stripe_rows = std::max(2, stripe_rows - (stripe_rows & 1));
    const int n = (height + stripe_rows - 1) / stripe_rows;
    std::vector<int> edge;
    std::vector<double> c;
    if (!cost) {                                 // no profile: equal heights
        edge = {0, height};
        c = {1.0};
    } else {
        if (n_cost != 2 * n) { g_err = "stripe cost vector does not match the stripe count of this height (two halves per stripe)"; return GRT_E_INVALID; }
        for (int i = 0; i < n; i++) {
            const StripeRows r = stripe_rows_of(i, stripe_rows, height);
            edge.push_back(r.y0);
            c.push_back(cost[2 * i]);
            if (r.mid < r.y1) { edge.push_back(r.mid); c.push_back(cost[2 * i + 1]); }
        }
        edge.push_back(height);
    }
    equal_cost_bounds(height, world, edge, c, bounds_out);
    return world + 1;
}

int grt_refine_bounds(int height, int world, int stripe_rows, const double* cost, int n_cost, const int* bounds, const double* measured, int* bounds_out) {
    if (height < 2 * world || world < 1 || !bounds || !measured || !bounds_out) { g_err = "bad partition request"; return GRT_E_INVALID; }
    if (bounds[0] != 0 || bounds[world] != height) { g_err = "bounds must run from 0 to the frame height"; return GRT_E_INVALID; }
    for (int r = 0; r < world; r++)
        if (bounds[r + 1] <= bounds[r]) { g_err = "bounds must ascend"; return GRT_E_INVALID; }
    // the profile's pieces (its shape inside a block: the rows tangent to the shadow's edge cost ten times the sky rows) ...
    std::vector<int> pe;
    std::vector<double> pc;
    if (cost) {
        stripe_rows = std::max(2, stripe_rows - (stripe_rows & 1));
        const int n = (height + stripe_rows - 1) / stripe_rows;
        if (n_cost != 2 * n) { g_err = "stripe cost vector does not match the stripe count of this height (two halves per stripe)"; return GRT_E_INVALID; }
        for (int i = 0; i < n; i++) {
            const StripeRows r = stripe_rows_of(i, stripe_rows, height);
            pe.push_back(r.y0);
            pc.push_back(std::max(0.0, cost[2 * i]));
            if (r.mid < r.y1) { pe.push_back(r.mid); pc.push_back(std::max(0.0, cost[2 * i + 1])); }
        }
        pe.push_back(height);
    } else {
        pe = {0, height};
        pc = {1.0};
    }
    // ... cut at the block boundaries, and every block's pieces scaled so that they add up to what the block really cost
    std::vector<int> edge;
    std::vector<double> c;
    std::vector<int> block_of;
    std::vector<double> predicted(world, 0.0);
    int k = 0;
    for (int r = 0; r < world; r++) {
        int y = bounds[r];
        while (y < bounds[r + 1]) {
            while (pe[k + 1] <= y) k++;
            const int y_end = std::min(bounds[r + 1], pe[k + 1]);
            const double part = pc[k] * (double)(y_end - y) / (double)(pe[k + 1] - pe[k]);
            edge.push_back(y);
            c.push_back(part);
            block_of.push_back(r);
            predicted[r] += part;
            y = y_end;
        }
    }
    edge.push_back(height);
    for (size_t i = 0; i < c.size(); i++) {
        const int r = block_of[i];
        const double rows = (double)(edge[i + 1] - edge[i]), block_rows = (double)(bounds[r + 1] - bounds[r]);
        c[i] = predicted[r] > 0.0 ? c[i] * std::max(0.0, measured[r]) / predicted[r] : std::max(0.0, measured[r]) * rows / block_rows;
    }
    equal_cost_bounds(height, world, edge, c, bounds_out);
    return world + 1;
}

void* grt_comm_frame(grt_ctx* c, unsigned seq) {
    if (!c || !c->comm.active || c->comm.rank != 0) { g_err = "only the owner (rank 0) holds the frames"; return nullptr; }
    return c->comm.base[0] + kCommFrames + (size_t)(seq % (unsigned)c->comm.n_frames) * c->comm.frame_bytes;
}

int grt_comm_render(grt_ctx* c, const int* y0s, const int* y1s, int n_stripes, unsigned seq, void* host_out) {
    return guard([&] {
        if (!c || !y0s || !y1s) throw Error(GRT_E_INVALID, "null argument");
        auto& m = c->comm;
        if (!m.active || !m.connected) throw Error(GRT_E_STATE, "communicator not connected: grt_comm_init, then grt_comm_connect / grt_comm_connect_local");
        if (host_out && m.rank != 0) throw Error(GRT_E_INVALID, "only the owner (rank 0) delivers the frame to the host");
        bind(c);
        const int f = (int)(seq % (unsigned)m.n_frames);
        char* frame = m.base[0] + kCommFrames + (size_t)f * m.frame_bytes;     // the owner's buffer, local or peer-mapped
        FrameDelivery d;
        d.frame_addressed = true;
        if (seq >= (unsigned)m.n_frames) {        // the buffer's previous frame must have been consumed by the owner
            d.wait_addr = comm_freed(c, m.rank, f);
            d.wait_value = seq - (unsigned)m.n_frames + 1u;
        }
        d.after = [&, f](cudaStream_t stream) {
            if (m.rank != 0) {
                comm_post(c, stream, comm_ready(c, 0, f, m.rank), true, seq + 1u);          // my stripes of frame `seq` are in the owner's buffer
                return;
            }
            for (int r = 1; r < m.world; r++)
                CU(g_drv.StreamWaitValue32((CUstream)stream, (CUdeviceptr)comm_ready(c, 0, f, r), seq + 1u, CU_STREAM_WAIT_VALUE_GEQ));
            if (host_out) CK(cudaMemcpyAsync(host_out, frame, (size_t)m.w * m.h * (m.format == GRT_FORMAT_RGBA32F ? 16 : 4), cudaMemcpyDeviceToHost, stream));
            for (int r = 0; r < m.world; r++) comm_post(c, stream, comm_freed(c, r, f), r != 0, seq + 1u);   // the buffer may be written again
        };
        return render_stripes(c, m.w, m.h, y0s, y1s, n_stripes, frame, m.format, 1, nullptr, &d);
    });
}

// ---- a group of contexts driven as one renderer -------------------------------------------------------------
struct grt_group {
    std::vector<grt_ctx*> ctx;
    std::vector<std::vector<int>> y0s, y1s;
    int w = 0, h = 0, format = -1, frames = 2;
    unsigned seq = 0;
    bool connected = false;
    std::vector<int> pending_bounds;       // grt_group_rebalance: the contiguous blocks (world + 1 boundaries) for a frame of height pending_bounds.back()
};

namespace {
int for_each_ctx(grt_group* g, const std::function<int(grt_ctx*)>& f) {
    if (!g) { g_err = "null group"; return GRT_E_INVALID; }
    for (grt_ctx* c : g->ctx) { int rc = f(c); if (rc < 0) return rc; }
    return GRT_OK;
}
void group_disconnect(grt_group* g) {
    for (grt_ctx* c : g->ctx) grt_comm_disconnect(c);
    for (grt_ctx* c : g->ctx)
        if (c->comm.active) { cudaSetDevice(c->device); c->comm.local.release(); c->comm = grt_ctx::Comm(); }
    g->connected = false;
}
}  // namespace

grt_group* grt_group_create(const int* devices, int ndev) {
    if (!devices || ndev < 1 || ndev > 64) { g_err = "bad device list"; return nullptr; }
    auto g = new grt_group();
    for (int i = 0; i < ndev; i++) {
        grt_ctx* c = grt_create(devices[i]);
        if (!c) { grt_group_destroy(g); return nullptr; }
        g->ctx.push_back(c);
    }
    g->y0s.resize(ndev);
    g->y1s.resize(ndev);
    return g;
}
void grt_group_destroy(grt_group* g) {
    if (!g) return;
    group_disconnect(g);
    for (int r = (int)g->ctx.size() - 1; r >= 0; r--) grt_destroy(g->ctx[r]);
    delete g;
}
int grt_group_size(const grt_group* g) { return g ? (int)g->ctx.size() : 0; }
grt_ctx* grt_group_ctx(grt_group* g, int rank) { return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr; }
int grt_group_set_metric(grt_group* g, grt_metric* m, int is_static) { return for_each_ctx(g, [&](grt_ctx* c) { return grt_set_metric(c, m, is_static); }); }
int grt_group_update_params(grt_group* g) { return for_each_ctx(g, [&](grt_ctx* c) { return grt_update_params(c); }); }
int grt_group_set_features(grt_group* g, const grt_features* f) { return for_each_ctx(g, [&](grt_ctx* c) { return grt_set_features(c, f); }); }
int grt_group_set_background(grt_group* g, const uint8_t* rgba, int w, int h, int which) { return for_each_ctx(g, [&](grt_ctx* c) { return grt_set_background(c, rgba, w, h, which); }); }
int grt_group_set_camera(grt_group* g, const float pos[4], const float quat[4], int flip_r) { return for_each_ctx(g, [&](grt_ctx* c) { return grt_set_camera(c, pos, quat, flip_r); }); }
int grt_group_set_frames_in_flight(grt_group* g, int n) {
    int rc = for_each_ctx(g, [&](grt_ctx* c) { return grt_set_frames_in_flight(c, n); });
    if (rc == GRT_OK && n != g->frames) { g->frames = n; if (g->connected) group_disconnect(g); }     // the frame buffers follow the depth
    return rc;
}
int grt_group_render(grt_group* g, int w, int h, int format, void* host_out) {
    if (!g) { g_err = "null group"; return GRT_E_INVALID; }
    const int n = (int)g->ctx.size();
    if (!g->connected || w != g->w || h != g->h || format != g->format) {
        // (re)build the communicator for this frame shape: every rank's region, then the peer pointers
        if (g->connected) group_disconnect(g);
        for (int r = 0; r < n; r++) {
            g->y0s[r].assign(h / 2 + 2, 0);
            g->y1s[r].assign(h / 2 + 2, 0);
            int k;
            if ((int)g->pending_bounds.size() == n + 1 && g->pending_bounds.back() == h) {       // shared out by cost (grt_group_rebalance)
                g->y0s[r][0] = g->pending_bounds[r];
                g->y1s[r][0] = g->pending_bounds[r + 1];
                k = 1;
            } else {
                k = grt_interleaved_stripes(h, r, n, 54, g->y0s[r].data(), g->y1s[r].data(), (int)g->y0s[r].size());
            }
            if (k < 0) return k;
            g->y0s[r].resize(k);
            g->y1s[r].resize(k);
            int rc = grt_comm_init(g->ctx[r], r, n, w, h, format, std::max(2, g->frames), nullptr);
            if (rc < 0) return rc;
        }
        int rc = grt_comm_connect_local(g->ctx.data(), n);
        if (rc < 0) return rc;
        g->w = w; g->h = h; g->format = format; g->seq = 0; g->connected = true;
    }
    for (int r = n - 1; r >= 0; r--) {       // the owner last: its stream then waits for stores that are already queued
        int rc = grt_comm_render(g->ctx[r], g->y0s[r].data(), g->y1s[r].data(), (int)g->y0s[r].size(), g->seq, r == 0 ? host_out : nullptr);
        if (rc < 0) return rc;
    }
    g->seq++;
    return GRT_OK;
}
// Share the frame out again by what it costs as it is now (metric, features and camera as set): rank 0 measures the stripe-cost profile,
// every rank gets one contiguous block of rows with boundaries at equal cost, and two rounds of feedback move the boundaries by what
// every context's block really cost (its trial ray-steps).  Takes effect with the next grt_group_render.  Waits for the frames in flight.
int grt_group_rebalance(grt_group* g, int w, int h) {
    if (!g) { g_err = "null group"; return GRT_E_INVALID; }
    const int n = (int)g->ctx.size();
    if (n == 1) return GRT_OK;
    if (h < 2 * n) { g_err = "frame too short to share out"; return GRT_E_INVALID; }
    int rc = grt_group_synchronize(g);
    if (rc < 0) return rc;
    std::vector<double> cost(h + 4);
    const int k = grt_measure_stripe_costs(g->ctx[0], w, h, 54, cost.data(), (int)cost.size());
    if (k < 0) return k;
    std::vector<int> bounds(n + 1), next(n + 1);
    rc = grt_contiguous_bounds(h, n, 54, cost.data(), k, bounds.data());
    if (rc < 0) return rc;
    for (int round = 0; round < 2; round++) {
        std::vector<double> measured(n, 0.0);
        for (int r = 0; r < n; r++) {
            grt_ctx* c = g->ctx[r];
            DevBuf scratch;
            rc = guard([&] {
                bind(c);
                scratch.ensure((size_t)(bounds[r + 1] - bounds[r]) * (size_t)w * 4);
                return GRT_OK;
            });
            if (rc < 0) return rc;
            grt_stats st{};
            rc = grt_render_stripes(c, w, h, &bounds[r], &bounds[r + 1], 1, scratch.p, GRT_FORMAT_RGBA8, GRT_OUT_DEVICE, &st);
            scratch.release();
            if (rc < 0) return rc;
            measured[r] = (double)st.trial_steps;
        }
        rc = grt_refine_bounds(h, n, 54, cost.data(), k, bounds.data(), measured.data(), next.data());
        if (rc < 0) return rc;
        bounds = next;
    }
    g->pending_bounds = bounds;
    if (g->connected && w == g->w && h == g->h)
        for (int r = 0; r < n; r++) { g->y0s[r] = {bounds[r]}; g->y1s[r] = {bounds[r + 1]}; }
    return GRT_OK;
}
void* grt_group_frame(grt_group* g) { return (g && g->connected && g->seq > 0) ? grt_comm_frame(g->ctx[0], g->seq - 1) : nullptr; }
int grt_group_synchronize(grt_group* g) { return for_each_ctx(g, [&](grt_ctx* c) { return grt_synchronize(c); }); }

// ---- introspection ------------------------------------------------------------------------------------
int grt_get_camera_state(grt_ctx* c, grt_camera_state* o) {
    return guard([&] {
        if (!c || !o) throw Error(GRT_E_INVALID, "null argument");
        bind(c);
        need_program(c);
        Launcher L{c->stream};
        upload_camera(c, L);
        GrtCamera h;
        CK(cudaMemcpyAsync(&h, c->d_cam, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        memcpy(o->pos_generic, &h.pos_generic, 16);
        memcpy(o->tetrad, h.e, 64);
        o->timelike_coordinate = h.timelike_coordinate;
        return GRT_OK;
    });
}

int grt_debug_ray_count(grt_ctx* c, int* first, int* extra) {
    return guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        if (c->last->last_extra < 0) {
            sync_all(c);
            unsigned int v = 0;
            CK(cudaMemcpy(&v, c->last->d_counters + 4, 4, cudaMemcpyDeviceToHost));
            c->last->last_extra = (int)v;
        }
        if (first) *first = c->last->last_first_pass;
        if (extra) *extra = c->last->last_extra;
        return GRT_OK;
    });
}

int grt_debug_read_rays(grt_ctx* c, grt_lightray* out, int max_count) {
    int n = 0;
    int rc = guard([&] {
        if (!c || !out) throw Error(GRT_E_INVALID, "null argument");
        int first = 0, extra = 0;
        if (grt_debug_ray_count(c, &first, &extra) != GRT_OK) throw Error(GRT_E_CUDA, g_err);
        sync_all(c);
        n = std::min(max_count, first + extra);
        if (n <= 0) return GRT_OK;
        std::vector<float4> pos(n), vel(n), acc(n), quat(n, make_float4(0, 0, 0, 1));
        std::vector<int4> aux(n);
        CK(cudaMemcpy(pos.data(), c->last->rays_pos.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(vel.data(), c->last->rays_vel.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(acc.data(), c->last->rays_acc.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(aux.data(), c->last->rays_aux.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
        if (c->bound_constant_theta && c->last->rays_quat.p) CK(cudaMemcpy(quat.data(), c->last->rays_quat.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) {
            grt_lightray& r = out[i];
            memset(&r, 0, sizeof r);
            memcpy(r.position, &pos[i], 16);
            memcpy(r.velocity, &vel[i], 16);
            memcpy(r.acceleration, &acc[i], 16);
            memcpy(r.initial_quat, &quat[i], 16);
            memcpy(&r.ku_uobsu, &aux[i].x, 4);
            memcpy(&r.running_dlambda_dnew, &aux[i].y, 4);
            r.terminated = aux[i].z;
            r.sx = aux[i].w & 0xffff;
            r.sy = (aux[i].w >> 16) & 0xffff;
        }
        return GRT_OK;
    });
    return rc == GRT_OK ? n : rc;
}

int grt_debug_read_render_data(grt_ctx* c, grt_render_data* out, int max_count) {
    int n = 0;
    int rc = guard([&] {
        if (!c || !out) throw Error(GRT_E_INVALID, "null argument");
        bind(c);
        sync_all(c);
        const GrtFrame& fr = c->last->frame;
        n = std::min(max_count, fr.n_rows * fr.full_w);
        if (n <= 0) return GRT_OK;
        std::vector<float4> rd(n);
        std::vector<int> row_global(fr.n_rows);
        CK(cudaMemcpy(rd.data(), c->last->rdata.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(row_global.data(), fr.row_global, sizeof(int) * fr.n_rows, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) {
            grt_render_data& r = out[i];
            memset(&r, 0, sizeof r);
            r.tex_coord[0] = rd[i].x;
            r.tex_coord[1] = rd[i].y;
            r.z_shift = rd[i].z;
            int flags;
            memcpy(&flags, &rd[i].w, 4);
            r.terminated = flags & 0xff;
            r.side = flags >> 8;
            r.sx = i % fr.full_w;
            r.sy = row_global[i / fr.full_w];
        }
        return GRT_OK;
    });
    return rc == GRT_OK ? n : rc;
}

int grt_debug_read_prepass_mask(grt_ctx* c, int* out, int max_count, int* pw, int* ph) {
    int n = 0;
    int rc = guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        sync_all(c);
        if (pw) *pw = c->last->last_pw;
        if (ph) *ph = c->last->last_ph;
        if (!c->last->last_prepass) return GRT_OK;
        n = std::min(max_count, c->last->last_pw * c->last->last_ph);
        if (out && n > 0) CK(cudaMemcpy(out, c->last->termination.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
        return GRT_OK;
    });
    return rc == GRT_OK ? n : rc;
}

int grt_debug_read_mip_level(grt_ctx* c, int which, int level, uint8_t* out, int* lw, int* lh, int* levels) {
    return guard([&] {
        if (!c || which < 0 || which > 1) throw Error(GRT_E_INVALID, "bad argument");
        bind(c);
        Texture& t = c->tex[which].set ? c->tex[which] : c->tex[0];
        if (!t.set) throw Error(GRT_E_STATE, "no background set");
        if (levels) *levels = t.desc.levels;
        if (level < 0 || level >= t.desc.levels) throw Error(GRT_E_INVALID, "mip level out of range");
        if (lw) *lw = t.desc.lw[level];
        if (lh) *lh = t.desc.lh[level];
        if (out) CK(cudaMemcpy(out, (const uchar4*)t.data.p + t.desc.offset[level], (size_t)t.desc.lw[level] * t.desc.lh[level] * 4, cudaMemcpyDeviceToHost));
        return GRT_OK;
    });
}

int grt_debug_read_geodesic(grt_ctx* c, float* positions, float* velocities, float* ds, float* transported, int max_points) {
    int n = 0;
    int rc = guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        CK(cudaStreamSynchronize(c->stream));
        n = std::min(max_points, c->geo_count);
        if (n <= 0) return GRT_OK;
        if (positions) CK(cudaMemcpy(positions, c->geo.positions, (size_t)n * 16, cudaMemcpyDeviceToHost));
        if (velocities) CK(cudaMemcpy(velocities, c->geo.velocities, (size_t)n * 16, cudaMemcpyDeviceToHost));
        if (ds) CK(cudaMemcpy(ds, c->geo.ds, (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (transported)
            for (int k = 0; k < 4; k++)
                CK(cudaMemcpy(transported + (size_t)k * max_points * 4, c->geo.transported + (size_t)k * c->geo.max_len, (size_t)n * 16, cudaMemcpyDeviceToHost));
        return GRT_OK;
    });
    return rc == GRT_OK ? n : rc;
}

int grt_debug_eval_metric(grt_ctx* c, const float* positions, const float* velocities, int count, float* g16, float* dg64, float* accel4, float* polar4, float* radii2) {
    return guard([&] {
        if (!c || !positions || !velocities || count <= 0) throw Error(GRT_E_INVALID, "bad argument");
        bind(c);
        need_program(c);
        size_t n = (size_t)count;
        size_t bytes = n * (16 + 16 + 64 + 256 + 16 + 16 + 8);
        c->scratch.ensure(bytes);
        char* base = (char*)c->scratch.p;
        float4* d_pos = (float4*)base;
        float4* d_vel = d_pos + n;
        float* d_g = (float*)(d_vel + n);
        float* d_dg = d_g + n * 16;
        float4* d_acc = (float4*)(d_dg + n * 64);
        float4* d_pol = d_acc + n;
        float2* d_rad = (float2*)(d_pol + n);
        CK(cudaMemcpyAsync(d_pos, positions, n * 16, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_vel, velocities, n * 16, cudaMemcpyHostToDevice, c->stream));
        Launcher L{c->stream};
        void* a[] = {&d_pos, &d_vel, &count, &d_g, &d_dg, &d_acc, &d_pol, &d_rad, c->params};
        L.go(c->prog->k_eval_metric, (unsigned)((count + 63) / 64), 64, a);
        if (g16) CK(cudaMemcpyAsync(g16, d_g, n * 64, cudaMemcpyDeviceToHost, c->stream));
        if (dg64) CK(cudaMemcpyAsync(dg64, d_dg, n * 256, cudaMemcpyDeviceToHost, c->stream));
        if (accel4) CK(cudaMemcpyAsync(accel4, d_acc, n * 16, cudaMemcpyDeviceToHost, c->stream));
        if (polar4) CK(cudaMemcpyAsync(polar4, d_pol, n * 16, cudaMemcpyDeviceToHost, c->stream));
        if (radii2) CK(cudaMemcpyAsync(radii2, d_rad, n * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return GRT_OK;
    });
}

int grt_debug_tetrad_at(grt_ctx* c, const float* positions, int count, const float speed[3], int orient, float* tetrads16, int* timelike) {
    return guard([&] {
        if (!c || !positions || count <= 0 || !tetrads16) throw Error(GRT_E_INVALID, "bad argument");
        bind(c);
        need_program(c);
        size_t n = (size_t)count;
        c->scratch.ensure(n * (16 + 64 + 4));
        float4* d_pos = (float4*)c->scratch.p;
        float4* d_tet = d_pos + n;
        int* d_tl = (int*)(d_tet + n * 4);
        CK(cudaMemcpyAsync(d_pos, positions, n * 16, cudaMemcpyHostToDevice, c->stream));
        float3 sp = make_float3(speed ? speed[0] : 0.f, speed ? speed[1] : 0.f, speed ? speed[2] : 0.f);
        Launcher L{c->stream};
        void* a[] = {&d_pos, &count, &sp, &orient, &d_tet, &d_tl, c->params};
        L.go(c->prog->k_tetrad_at, (unsigned)((count + 63) / 64), 64, a);
        CK(cudaMemcpyAsync(tetrads16, d_tet, n * 64, cudaMemcpyDeviceToHost, c->stream));
        if (timelike) CK(cudaMemcpyAsync(timelike, d_tl, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return GRT_OK;
    });
}

int grt_debug_integrate_launch_ms(grt_ctx* c, float* out, int max_count) {
    int n = 0;
    int rc = guard([&] {
        if (!c) throw Error(GRT_E_INVALID, "null context");
        bind(c);
        sync_all(c);
        for (int i = 0; i < c->n_slots; i++) {
            FrameSlot& S = c->slots[i];
            if (!S.used) continue;
            for (int k = 0; k < S.integrate_launches_last; k++) {
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, S.integrate_events[k].first, S.integrate_events[k].second));
                if (out && n < max_count) out[n] = ms;
                n++;
            }
        }
        return GRT_OK;
    });
    return rc == GRT_OK ? n : rc;
}

const char* grt_debug_kernel_info(grt_ctx* c) { return c ? c->kernel_info.c_str() : nullptr; }
const char* grt_debug_launch_log(grt_ctx* c) { return c ? c->launch_log.c_str() : nullptr; }

}  // extern "C"
