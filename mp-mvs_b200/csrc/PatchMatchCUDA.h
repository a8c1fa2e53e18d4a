// PatchMatchCUDA.h -- OpenCV-free C++ mirror of the reference's host surface for the PatchMatch path, on top of the
// C ABI (include/mpmvs_b200.h). Same class name, method names, argument meaning and call order as
// /root/reference/include/PatchMatch.h:87-154, so a ProcessProblem written against the reference reads the same here.
//
// What differs, deliberately:
//   * no cv::Mat / cv::Point in signatures (OpenCV is not a dependency): GrayImage, Point, Rect, Triangle below;
//   * the object owns ONE mpmvs_problem handle; AllocatePatchMatch / CudaMemInit / Release forward to it instead of doing
//     cudaMalloc / cudaMallocArray / texture creation per call (PatchMatch.cpp:960-1089);
//   * errors throw std::runtime_error (the reference prints and exit()s, PatchMatch.cpp:60-65);
//   * Run(seed): the RNG seed is explicit (the reference seeds from clock64(), PatchMatch.cu:546);
//   * BuildPlanarPrior() runs the whole host prior stage of ProcessProblem (PatchMatch.cpp:536-600) on the GPU state;
//     the piecewise methods (GetTriangulateVertices, DelaunayTriangulation, GetPriorPlaneParams, ...) are kept too.
#ifndef MPMVS_PATCHMATCHCUDA_H
#define MPMVS_PATCHMATCHCUDA_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mpmvs_b200.h"

namespace mpmvs {

using Camera = mpmvs_camera;                       // struct Camera, PatchMatch.h:35-46
struct float4 { float x, y, z, w; };
struct Point { int x = 0, y = 0; };
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    bool contains(const Point& p) const { return p.x >= x && p.x < x + width && p.y >= y && p.y < y + height; }
};
struct Triangle { Point pt1, pt2, pt3; };          // PatchMatch.h:69-72
struct GrayImage {                                 // float grey levels 0..255, row-major (cv::Mat CV_32FC1 of PatchMatchInit)
    int width = 0, height = 0;
    std::vector<float> px;
    std::vector<unsigned char> u8;                  // the same pixels as decoded, when the image was not resized (8-bit upload path)
    bool empty() const { return px.empty(); }
};
struct Scene {                                     // utility.h:17-26
    bool estimate = false;
    int refID = -1;
    std::vector<int> srcID;
    GrayImage image;                                // decoded (and, above max_image_size, resized) grey image, cached
    int orig_width = 0, orig_height = 0;            // size of the file on disk (K is scaled by image size / this)
    // latest results of this image (what its .dmb files hold), kept in memory; shared so that the background .dmb writer
    // and a later pass can hold the same buffers
    std::shared_ptr<const std::vector<float>> depth, normal, cost;
    int max_image_size = 3200;
};
struct ConfigParams {                              // utility.h:28-46
    std::string input_folder, output_folder;
    int geom_iterations = 2;
    bool geom_consistency = false, planar_prior = true, geomPlanarPrior = true, sky_seg = false, use_dynamic_consistency = true;
    bool saveDmb = false, saveProirDmb = false, saveCostDmb = false, saveNormalDmb = false;
    int MaxSourceImageNum = 20, MaxImageSize = 3200;
};

inline void check(int rc, const char* what) {
    if (rc != MPMVS_OK) throw std::runtime_error(std::string(what) + ": " + mpmvs_error_string(rc));
}

// ------------------------------------------------------------------------------------------------ formats
std::string id8(int id);                                                        // "%08d"
ConfigParams readConfig(const std::string& yaml_path);                          // utility.cpp:8-35 (13 keys)
void GenerateSampleList(const ConfigParams& config, std::vector<Scene>& Scenes);  // PatchMatch.cpp:67-109 (pair.txt)
Camera ReadCamera(const std::string& cam_path);                                 // PatchMatch.cpp:111-143
bool readDmb(const std::string& path, int& h, int& w, int& nb, std::vector<float>& data);   // utility.cpp:193-217,251-280
bool writeDmb(const std::string& path, int h, int w, int nb, const float* data);            // utility.cpp:219-248,282-308
bool readGrayImage(const std::string& image_folder, int id, GrayImage& out);    // %08d.pgm (decoded sidecar) or %08d.jpg (nvJPEG)
bool readGrayFile(const std::string& path_without_ext, GrayImage& out);        // <path>.pgm or <path>.jpg (nvJPEG luma)
bool readColorFile(const std::string& path_without_ext, int& w, int& h, std::vector<unsigned char>& bgr);   // .ppm or .jpg, B G R interleaved
std::vector<unsigned char> resizeLinearBGR(const std::vector<unsigned char>& src, int width, int height, int new_cols, int new_rows);   // cv::resize, CV_8UC3, INTER_LINEAR
bool writePgm(const std::string& path, int w, int h, const unsigned char* px);
GrayImage resizeLinear(const GrayImage& src, int new_cols, int new_rows);       // cv::resize(..., INTER_LINEAR) on float
// Background .dmb writers: ProcessProblem hands its three result maps over and returns; Flush() waits for the files
// (the reference writes them synchronously, PatchMatch.cpp:620-633).
void SubmitDmb(const std::string& path, int h, int w, int nb, std::shared_ptr<const std::vector<float>> data);
void FlushDmbWriters();
// seconds spent per host phase of ProcessProblem since the start of the process (printed by mpmvs_main --profile)
struct HostPhaseTimes { double init = 0, upload = 0, run = 0, prior = 0, collect = 0, write = 0; };
HostPhaseTimes& PhaseTimes();
int GenerateSkyRegionMask(std::vector<Scene>& Scenes, const ConfigParams& config);   // PatchMatch.cpp:4-57 minus the network
std::vector<unsigned char> readSkyMask(const std::string& result_folder, int w, int h);   // PatchMatch.cpp:358-373

// ------------------------------------------------------------------------------------------------ the class
class PatchMatchCUDA {
  public:
    // The reference constructs one object per ProcessProblem call and Release()s it (PatchMatch.cpp:516,637). Handles are
    // pooled per device here, so that costs one mpmvs_reset_params instead of a round of cudaMalloc/cudaFree, texture
    // creation and pinned allocations.
    explicit PatchMatchCUDA(int device = 0) : device_(device) {
        std::vector<mpmvs_problem*>& pool = handle_pool(device);
        if (!pool.empty()) { h_ = pool.back(); pool.pop_back(); check(mpmvs_reset_params(h_), "mpmvs_reset_params"); }
        else check(mpmvs_create(device, nullptr, &h_), "mpmvs_create");
    }
    ~PatchMatchCUDA() { if (h_) handle_pool(device_).push_back(h_); }
    static void ReleasePool() {
        for (auto& kv : pools()) { for (mpmvs_problem* h : kv.second) mpmvs_destroy(h); kv.second.clear(); }
    }
    PatchMatchCUDA(const PatchMatchCUDA&) = delete;
    PatchMatchCUDA& operator=(const PatchMatchCUDA&) = delete;

    void SetGeomConsistencyParams(bool geom_consistency, bool planar_prior) {          // PatchMatch.cpp:655-665
        geom_ = geom_consistency;
        if (geom_consistency) geomPlanarPrior_ = planar_prior;
        check(mpmvs_set_geom_consistency_params(h_, geom_consistency, planar_prior), "SetGeomConsistencyParams");
    }
    void SetPlanarPriorParams() { planar_ = true; check(mpmvs_set_planar_prior_params(h_), "SetPlanarPriorParams"); }
    void SetFolder(const std::string& in, const std::string& out) { input_folder_ = in; output_folder_ = out; }
    void SetTexFormat(int fmt) { tex_format_ = fmt; check(mpmvs_set_tex_format(h_, fmt), "SetTexFormat"); }

    // PatchMatch.cpp:863-958: load images + cameras of Scenes[ID].srcID, resize above max_image_size, depth range;
    // in geom mode load the sources' depths.dmb. Scenes is taken by reference: decoded images stay cached in it.
    void PatchMatchInit(std::vector<Scene>& Scenes, int ID);
    void AllocatePatchMatch() {}                                                        // buffers live in the handle
    void CudaMemInit(Scene& scene);                                                     // PatchMatch.cpp:998-1089
    void CudaPlanarPriorInitialization(const std::vector<float4>& PlaneParams, const std::vector<float>& masks);  // :978-996
    void BuildPlanarPrior(mpmvs_prior_stats* stats = nullptr) { check(mpmvs_build_prior(h_, stats), "mpmvs_build_prior"); }
    void Run(uint64_t seed);                                                            // PatchMatch.cu:1188-1254
    void Release(std::vector<Scene>&, const int&) {}                                    // the destructor releases the handle

    float GetDepthFromPlaneParam(const float4 pl, int x, int y) const {                 // PatchMatch.cpp:650-653
        const Camera& c = cameras_[0];
        return -pl.w * c.K[0] / ((x - c.K[2]) * pl.x + (c.K[0] / c.K[4]) * (y - c.K[5]) * pl.y + c.K[0] * pl.z);
    }
    float GetMinDepth() const { return depth_min_; }
    float GetMaxDepth() const { return depth_max_; }
    int GetReferenceImageWidth() const { return cameras_[0].width; }
    int GetReferenceImageHeight() const { return cameras_[0].height; }
    const GrayImage& GetReferenceImage() const { return *images_[0]; }
    float4 GetPlaneHypothesis(int index) const { return planes_[index]; }
    float GetCost(int index) const { return costs_[index]; }
    float GetGeomCost(int index) const { return geom_costs_[index]; }
    const std::vector<float4>& planes() const { return planes_; }
    const std::vector<float>& costs() const { return costs_; }

    float4 GetPriorPlaneParams(const Triangle& t, int width) const;                     // PatchMatch.cpp:723-755
    std::vector<Triangle> DelaunayTriangulation(const Rect& boundRC, const std::vector<Point>& points) const;  // :757-780
    void GetTriangulateVertices(std::vector<Point>& Vertices);                          // :782-853

  private:
    static std::map<int, std::vector<mpmvs_problem*>>& pools() { static std::map<int, std::vector<mpmvs_problem*>> p; return p; }
    static std::vector<mpmvs_problem*>& handle_pool(int device) { return pools()[device]; }
    int device_ = 0;
    int tex_format_ = MPMVS_TEX_F32;
    mpmvs_problem* h_ = nullptr;
    std::vector<const GrayImage*> images_;         // the decoded images cached in Scenes (not copied per call)
    std::vector<std::shared_ptr<const std::vector<float>>> depths_;   // the sources' depth maps of the previous pass (not copied)
    std::vector<Camera> cameras_;
    std::vector<float4> planes_;
    std::vector<float> costs_, geom_costs_;
    bool geom_ = false, geomPlanarPrior_ = false, planar_ = false, restart_loaded_ = false;
    float depth_min_ = 0.f, depth_max_ = 1.f;
    std::string input_folder_, output_folder_;
    int ref_id_ = -1;
};

// PatchMatch.cpp:506-638. seed: base seed of this (image, pass).
void ProcessProblem(const std::string& input_folder, const std::string& output_folder, std::vector<Scene>& Scenes, int ID,
                    bool geom_consistency, bool planar_prior, uint64_t seed, int tex_format = MPMVS_TEX_F32);

}  // namespace mpmvs
#endif
