// oracle/ref_host_harness.cu -- TEST INFRASTRUCTURE, not product code.
//
// The reference's WHOLE program -- /root/reference/src/main.cpp, PatchMatch.cpp, utility.cpp and PatchMatch.cu, each
// compiled from where it lies, unmodified -- as oracle/_ref/libmpmvs_ref_host.so, so that tests can run "the reference
// itself" over a dense folder and compare the files it writes (depths/normals/costs .dmb, MPMVS_model.ply) and the
// planar prior it builds with the product's (SURVEY.md 8(c), rows f2 / f3 of 8(f)).
//
// What is the reference's and what is not:
//   * main(), ProcessProblem, the planar-prior host stage (GetTriangulateVertices, DelaunayTriangulation, the triangle
//     rasterisation, GetPriorPlaneParams, the depth-range check), RunFusion, the readers/writers, every kernel and Run():
//     the reference's own translation units, #included below;
//   * OpenCV: the C++ SDK is not in this image. Containers come from oracle/ref_shim_host/opencv2/opencv.hpp (ours);
//     cv::Subdiv2D, cv::SVD::solveZ, cv::resize and cv::imread are forwarded through callbacks to the REAL OpenCV of this
//     image (cv2 4.13, tests/ref_host.py). The reference pins no OpenCV version (README.md:5 ">= 2.4");
//   * three redirections, all by macro around the #include (no source edit; main() keeps its name, as a hidden symbol):
//       curand_init   -> fixed seed instead of clock64()            (as oracle/ref_harness.cu)
//       cudaSetDevice -> called once at the top of every ProcessProblem (PatchMatch.cpp:509): picks that call's seed
//       cudaMemcpy    -> passes through; the two uploads of CudaPlanarPriorInitialization (PatchMatch.cpp:994-995) and the
//                        downloads at the end of Run() (PatchMatch.cu:1246-1250) are also copied aside, so that tests can
//                        look at the prior the reference built and at the state it built it from
//   * the preview writers (cv::imwrite & co.) write nothing; the imwrite of triangulation.png (PatchMatch.cpp:603) is
//     used as the marker "the planar-prior Run() of this ProcessProblem comes next".
//
// Only tests/ may load the resulting library.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_runtime_api.h>
#include <cuda_texture_types.h>
#include <curand_kernel.h>
#include <vector_types.h>

#include <opencv2/opencv.hpp>  // oracle/ref_shim_host

// ---------------------------------------------------------------- fixed-seed redirection (same mixing as ref_harness.cu)
__device__ unsigned long long g_ref_seed = 0ULL;

__host__ __device__ inline unsigned long long ref_mix_seed(unsigned long long seed, unsigned int x, unsigned int y) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * ((((unsigned long long)y) << 32) | (unsigned long long)x) +
                           0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ inline void ref_curand_init(unsigned long long /*clock*/, unsigned long long row, unsigned long long col,
                                       curandState* state) {
    curand_init(ref_mix_seed(g_ref_seed, (unsigned int)col, (unsigned int)row), 0ULL, 0ULL, state);
}

// ---------------------------------------------------------------- host-side hooks
namespace ref_host {
struct SeedPlan { unsigned long long base = 0, per_image = 0, per_stage = 0, second_run_xor = 0; };
struct CapturedPrior {
    int scene_index, stage;
    size_t pixels;
    std::vector<float> planes;            // the prior the reference built (float4 per pixel; defined where mask > 0)
    std::vector<unsigned int> mask;
    std::vector<float> in_planes, in_costs, in_geom;   // what the Run() before it downloaded: the host stage's input
};
struct LastRun { size_t pixels = 0; int next = 0; std::vector<float> planes, costs, geom; };
LastRun g_last;
SeedPlan g_plan;
unsigned long long g_current_seed = 0;
int g_calls = 0, g_scene_index = -1, g_stage = -1;
int g_expect_prior_uploads = 0;
bool g_capture = false;
std::vector<CapturedPrior> g_priors;

void set_device_seed(unsigned long long s) {
    g_current_seed = s;
    cudaMemcpyToSymbol(g_ref_seed, &s, sizeof(s));
}
void begin_process_problem();   // needs the reference's global `Scenes`; defined after main.cpp is included

cudaError_t hook_cudaSetDevice(int device) {
    const cudaError_t rc = cudaSetDevice(device);
    begin_process_problem();
    return rc;
}
cudaError_t hook_cudaMemcpy(void* dst, const void* src, size_t count, cudaMemcpyKind kind) {
    if (kind == cudaMemcpyDeviceToHost) {                 // the downloads at the end of Run(), PatchMatch.cu:1246-1250: planes, costs[, geometric costs]
        const cudaError_t rc = cudaMemcpy(dst, src, count, kind);
        if (g_capture && rc == cudaSuccess) {
            const float* f = (const float*)dst;
            if (g_last.next == 0 || count == g_last.pixels * sizeof(float4)) {
                g_last.pixels = count / sizeof(float4);
                g_last.planes.assign(f, f + count / sizeof(float));
                g_last.costs.clear(); g_last.geom.clear();
                g_last.next = 1;
            } else if (g_last.next == 1 && count == g_last.pixels * sizeof(float)) {
                g_last.costs.assign(f, f + g_last.pixels);
                g_last.next = 2;
            } else if (g_last.next == 2 && count == g_last.pixels * sizeof(float)) {
                g_last.geom.assign(f, f + g_last.pixels);
                g_last.next = 0;
            }
        }
        return rc;
    }
    if (g_expect_prior_uploads > 0 && kind == cudaMemcpyHostToDevice) {
        if (g_capture) {
            if (g_expect_prior_uploads == 2) {         // hostPriorPlanes (float4 per pixel), PatchMatch.cpp:994
                CapturedPrior c;
                c.scene_index = g_scene_index; c.stage = g_stage; c.pixels = count / sizeof(float4);
                c.planes.assign((const float*)src, (const float*)src + count / sizeof(float));
                if (g_last.pixels == c.pixels) { c.in_planes = g_last.planes; c.in_costs = g_last.costs; c.in_geom = g_last.geom; }
                g_priors.push_back(std::move(c));
            } else if (!g_priors.empty()) {              // hostPlaneMask (unsigned per pixel), PatchMatch.cpp:995
                g_priors.back().mask.assign((const unsigned int*)src, (const unsigned int*)src + count / sizeof(unsigned int));
            }
        }
        --g_expect_prior_uploads;
    }
    return cudaMemcpy(dst, src, count, kind);
}
void on_imwrite(const char* path) {
    const char* marker = "/triangulation.png";
    const size_t n = strlen(path), m = strlen(marker);
    if (n >= m && strcmp(path + n - m, marker) == 0) {   // PatchMatch.cpp:603: CudaPlanarPriorInitialization and the prior Run() follow
        g_expect_prior_uploads = 2;
        set_device_seed(g_current_seed ^ g_plan.second_run_xor);
    }
}
}  // namespace ref_host

// ---------------------------------------------------------------- the reference's translation units, in place
#define cudaMemcpy ref_host::hook_cudaMemcpy
#define curand_init ref_curand_init
#include <src/PatchMatch.cu>   // -I /root/reference
#undef curand_init

#define cudaSetDevice ref_host::hook_cudaSetDevice
#include <src/PatchMatch.cpp>
#undef cudaMemcpy
#undef cudaSetDevice

#include <src/utility.cpp>

// main() has no return statement (main.cpp:7-54): legal only for a function called `main` (any other name runs into the
// trap gcc puts at the end of a non-void function). So it keeps its name, with hidden visibility: a local symbol of this
// library that the harness calls directly and that nobody can confuse with the process's own main.
__attribute__((visibility("hidden"))) int main(int argc, char* argv[]);
#include <src/main.cpp>

// ---------------------------------------------------------------- hooks that need the program's globals
namespace ref_host {
void begin_process_problem() {
    std::vector<int> estimated;
    for (size_t i = 0; i < Scenes.size(); ++i)
        if (Scenes[i].estimate) estimated.push_back((int)i);
    const int n = std::max<int>(1, (int)estimated.size());
    g_stage = g_calls / n;                                // main() walks the estimated images once per stage (main.cpp:20-41)
    g_scene_index = estimated.empty() ? 0 : estimated[g_calls % n];
    ++g_calls;
    g_expect_prior_uploads = 0;
    set_device_seed(g_plan.base + g_plan.per_image * (unsigned long long)g_scene_index + g_plan.per_stage * (unsigned long long)g_stage);
}
}  // namespace ref_host

// ---------------------------------------------------------------- C API for tests/ref_host.py
extern "C" {

void ref_host_set_callbacks(int (*imread)(const char*, int, int*, int*, int*, const uchar**),
                            int (*resize)(const void*, int, int, int, int, int, void*),
                            int (*subdiv)(int, int, int, int, const float*, int, const float**, int*),
                            int (*solvez)(const float*, int, int, float*)) {
    ref_shim::Callbacks& cb = ref_shim::callbacks();
    cb.imread = imread; cb.resize = resize; cb.subdiv = subdiv; cb.solvez = solvez;
    cb.on_imwrite = ref_host::on_imwrite;
}

// seed of ProcessProblem(image i, stage s) = base + per_image * i + per_stage * s; its planar-prior Run() uses that ^ second_run_xor
void ref_host_set_seed_plan(unsigned long long base, unsigned long long per_image, unsigned long long per_stage, unsigned long long second_run_xor) {
    ref_host::g_plan.base = base; ref_host::g_plan.per_image = per_image; ref_host::g_plan.per_stage = per_stage;
    ref_host::g_plan.second_run_xor = second_run_xor;
}

void ref_host_capture_priors(int enable) { ref_host::g_capture = enable != 0; ref_host::g_priors.clear(); }
int ref_host_num_priors() { return (int)ref_host::g_priors.size(); }
int ref_host_prior_info(int k, int* scene_index, int* stage, long long* pixels) {
    if (k < 0 || k >= (int)ref_host::g_priors.size()) return 1;
    *scene_index = ref_host::g_priors[k].scene_index; *stage = ref_host::g_priors[k].stage; *pixels = (long long)ref_host::g_priors[k].pixels;
    return 0;
}
int ref_host_get_prior(int k, float* planes, unsigned int* mask) {
    if (k < 0 || k >= (int)ref_host::g_priors.size()) return 1;
    const ref_host::CapturedPrior& c = ref_host::g_priors[k];
    if (c.mask.size() != c.pixels) return 2;
    memcpy(planes, c.planes.data(), c.planes.size() * sizeof(float));
    memcpy(mask, c.mask.data(), c.mask.size() * sizeof(unsigned int));
    return 0;
}
// the state the prior k was built from; geom may be absent (returns 3 then, planes and costs still filled)
int ref_host_get_prior_input(int k, float* planes, float* costs, float* geom) {
    if (k < 0 || k >= (int)ref_host::g_priors.size()) return 1;
    const ref_host::CapturedPrior& c = ref_host::g_priors[k];
    if (c.in_planes.size() != c.pixels * 4 || c.in_costs.size() != c.pixels) return 2;
    memcpy(planes, c.in_planes.data(), c.in_planes.size() * sizeof(float));
    memcpy(costs, c.in_costs.data(), c.in_costs.size() * sizeof(float));
    if (c.in_geom.size() != c.pixels) return 3;
    memcpy(geom, c.in_geom.data(), c.in_geom.size() * sizeof(float));
    return 0;
}

// Runs the reference's main() (src/main.cpp:7-54) with <project>/config/config.yaml as its configuration. Errors inside
// the reference call exit(): run this in a child process.
int ref_host_main(const char* project) {
    project_path = project;                               // the global of main.cpp:5 (PROJECT_PATH is the author's home directory)
    ref_host::g_calls = 0;
    char name[] = "MPMVS";
    char* argv[] = {name, nullptr};
    return main(1, argv);
}

// ---- single functions of the reference's host stage on caller-supplied state (unit comparisons) ---------------------
// RunFusion alone over a folder that already holds the depth maps (src/PatchMatch.cpp:287-504).
int ref_host_fusion(const char* project) {
    project_path = project;
    std::string yaml_path = project_path + "/config/config.yaml";
    ConfigParams config = readConfig(yaml_path);
    GenerateSampleList(config, Scenes);
    RunFusion(config, Scenes);
    return 0;
}

}  // extern "C"
