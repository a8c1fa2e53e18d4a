// mpmvs_main.cpp -- the reference's entry point (/root/reference/src/main.cpp:6-55) over the C++ host mirror:
// config.yaml -> pair.txt -> stage 1 (multi-scale photometric [+ planar prior]) -> geometric-consistency iterations
// [+ planar prior] -> fusion -> MPMVS_model.ply, on the same dense-folder layout and with the same output files.
//
//   mpmvs_main [config.yaml] [--seed N] [--tex f32|u8] [--no-fusion] [--gpu-fusion] [--resident [--in-flight N] [--device D] [--gpus G]] [--profile] [--check-inputs DIR] [--fusion-only] [--arithmetic exact|fast]
//
// Default: the reference's strictly sequential order (one ProcessProblem at a time, results exchanged in place).
// --resident: every view uploaded once into a per-GPU image cache, one resident handle per reference image, depth maps
// exchanged between passes in device memory (Jacobi order: every image reads the previous pass), N images in flight on
// host threads; with --gpus G the reference images are sharded over G GPUs and the depth maps all-gathered with peer copies
// over NVLink -- mp-mvs_b200/pipeline.py's schedule in C++, one process driving all GPUs.
//
// The reference bakes the config path in at cmake time (include/ProjectPath.h.in) and takes no arguments.
#include <sys/stat.h>

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "PatchMatchCUDA.h"

namespace mpmvs {

struct PointList { float coord[3], normal[3], color[3]; };      // PatchMatch.h:29-33

static void world_point(int x, int y, float depth, const Camera& c, float* X) {     // Get3DPointonWorld, PatchMatch.cpp:211-230
    const float px = depth * (x - c.K[2]) / c.K[0], py = depth * (y - c.K[5]) / c.K[4], pz = depth;
    X[0] = c.R[0] * px + c.R[3] * py + c.R[6] * pz + c.C[0];
    X[1] = c.R[1] * px + c.R[4] * py + c.R[7] * pz + c.C[1];
    X[2] = c.R[2] * px + c.R[5] * py + c.R[8] * pz + c.C[2];
}
static void project(const float* X, const Camera& c, float& u, float& v, float& depth) {   // ProjectonCamera, :252-262
    const float tx = c.R[0] * X[0] + c.R[1] * X[1] + c.R[2] * X[2] + c.t[0];
    const float ty = c.R[3] * X[0] + c.R[4] * X[1] + c.R[5] * X[2] + c.t[1];
    const float tz = c.R[6] * X[0] + c.R[7] * X[1] + c.R[8] * X[2] + c.t[2];
    depth = c.K[6] * tx + c.K[7] * ty + c.K[8] * tz;
    u = (c.K[0] * tx + c.K[1] * ty + c.K[2] * tz) / depth;
    v = (c.K[3] * tx + c.K[4] * ty + c.K[5] * tz) / depth;
}

static void StoreColorPlyFileBinaryPointCloud(const std::string& path, const std::vector<PointList>& pc) {   // PatchMatch.cpp:145-198
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("can not write " + path);
    fprintf(f, "ply\nformat binary_little_endian 1.0\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\n"
               "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n",
            pc.size());
    std::vector<unsigned char> buf;                 // one write per megabyte instead of nine per point
    buf.reserve(27u << 16);
    for (const PointList& p : pc) {
        float X[3] = {p.coord[0], p.coord[1], p.coord[2]};
        const bool finite = X[0] < FLT_MAX && X[0] > -FLT_MAX && X[1] < FLT_MAX && X[1] > -FLT_MAX && X[2] < FLT_MAX && X[2] >= -FLT_MAX;
        if (!finite) X[0] = X[1] = X[2] = 0.f;
        const unsigned char rgb[3] = {(unsigned char)(char)(int)p.color[2], (unsigned char)(char)(int)p.color[1], (unsigned char)(char)(int)p.color[0]};
        const unsigned char* a = (const unsigned char*)X;
        const unsigned char* n = (const unsigned char*)p.normal;
        buf.insert(buf.end(), a, a + 12);
        buf.insert(buf.end(), n, n + 12);
        buf.insert(buf.end(), rgb, rgb + 3);
        if (buf.size() >= (27u << 16)) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    fclose(f);
}

// RunFusion, PatchMatch.cpp:287-504 (host, single thread, pixel order and mask side effects as in the reference, the sky
// gate of its BUILD_NCNN branch included when `Sky segment` is set). Colour: B, G, R of the colour image as the reference
// averages them (cv::imread(IMREAD_COLOR), :322,399,443-445) from images/%08d.{ppm,jpg}, resized to the depth map's size like
// RescaleImageAndCamera does (resizeLinearBGR = cv::resize's 8-bit fixed-point path); the grey level in all three channels
// when there is no colour file.
size_t RunFusion(const ConfigParams& config, std::vector<Scene>& Scenes) {
    const size_t n = Scenes.size();
    std::vector<Camera> cams(n);
    std::vector<std::vector<float>> depths(n), normals(n);
    std::vector<std::vector<unsigned char>> masks(n), sky(n), bgr(n);
    std::vector<int> W(n, 0), H(n, 0);
    const std::string image_folder = config.input_folder + "/images", cam_folder = config.input_folder + "/cams";
    for (size_t i = 0; i < n; ++i) {
        if (!Scenes[i].estimate) continue;
        const int id = Scenes[i].refID;
        cams[i] = ReadCamera(cam_folder + "/" + id8(id) + "_cam.txt");
        const std::string folder = config.input_folder + "/MPMVS/2333_" + id8(id);
        int h, w, nb, hn, wn, nbn;
        if (!readDmb(folder + "/depths.dmb", h, w, nb, depths[i]) || !readDmb(folder + "/normals.dmb", hn, wn, nbn, normals[i]))
            throw std::runtime_error("fusion: missing results in " + folder);
        if (nb != 1 || nbn != 3 || hn != h || wn != w) throw std::runtime_error("fusion: depths.dmb and normals.dmb do not belong together in " + folder);
        Scenes[i].depth.reset(); Scenes[i].normal.reset(); Scenes[i].cost.reset();     // the files are the interface from here on
        W[i] = w; H[i] = h;
        masks[i].assign((size_t)w * h, 0);
        if (config.sky_seg) sky[i] = readSkyMask(folder, w, h);                    // :358-373
        if (Scenes[i].image.empty() && !readGrayImage(image_folder, id, Scenes[i].image)) throw std::runtime_error("fusion: missing image " + id8(id));
        const GrayImage& im = Scenes[i].image;      // RescaleImageAndCamera, :264-285
        if (im.width != w || im.height != h) {
            const float sx = w / (float)im.width, sy = h / (float)im.height;
            Scenes[i].image = resizeLinear(im, w, h);
            cams[i].K[0] *= sx; cams[i].K[2] *= sx; cams[i].K[4] *= sy; cams[i].K[5] *= sy;
        } else if (Scenes[i].orig_width && (Scenes[i].orig_width != w || Scenes[i].orig_height != h)) {
            const float sx = w / (float)Scenes[i].orig_width, sy = h / (float)Scenes[i].orig_height;
            cams[i].K[0] *= sx; cams[i].K[2] *= sx; cams[i].K[4] *= sy; cams[i].K[5] *= sy;
        }
        cams[i].width = w; cams[i].height = h;
        int cw = 0, ch = 0;
        std::vector<unsigned char> col;
        if (readColorFile(image_folder + "/" + id8(id), cw, ch, col))              // RescaleImageAndCamera, :264-285: the colour image at the depth map's size
            bgr[i] = (cw == w && ch == h) ? std::move(col) : resizeLinearBGR(col, cw, ch, w, h);
    }
    std::map<int, int> id2index;
    for (size_t i = 0; i < n; ++i) if (Scenes[i].estimate) id2index[Scenes[i].refID] = (int)i;
    std::vector<PointList> cloud;
    for (size_t i = 0; i < n; ++i) {
        if (!Scenes[i].estimate) continue;
        std::cout << "Fusing image " << id8((int)i) << "..." << std::endl;
        const int cols = W[i], rows = H[i], num_ngb = (int)Scenes[i].srcID.size();
        std::vector<int> used_x(num_ngb, -1), used_y(num_ngb, -1);   // QUIRK :371: not reset per pixel
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const size_t idx = (size_t)r * cols + c;
                if (masks[i][idx] == 1) continue;
                if (!sky[i].empty() && sky[i][idx] > 0) { masks[i][idx] = 1; continue; }   // :385-388
                const float ref_depth = depths[i][idx];
                if (ref_depth <= 0.0f) continue;
                const float* rn = &normals[i][3 * idx];
                float PX[3];
                world_point(c, r, ref_depth, cams[i], PX);
                float sumP[3] = {PX[0], PX[1], PX[2]}, sumN[3] = {rn[0], rn[1], rn[2]};
                const float g = Scenes[i].image.px[idx];
                float sumC[3] = {g, g, g};
                if (!bgr[i].empty()) for (int k = 0; k < 3; ++k) sumC[k] = bgr[i][3 * idx + k];
                int num_consistent = 0;
                float dynamic_consistency = 0.f;
                for (int j = 1; j < num_ngb; ++j) {
                    if (j == num_ngb - 1 && num_consistent == 0) break;
                    auto it = id2index.find(Scenes[i].srcID[j]);
                    if (it == id2index.end()) continue;
                    const int s = it->second;
                    float u, v, pd;
                    project(PX, cams[s], u, v, pd);
                    const int sr = (int)(v + 0.5f), sc = (int)(u + 0.5f);
                    if (!(sc >= 0 && sc < W[s] && sr >= 0 && sr < H[s])) continue;
                    const size_t sidx = (size_t)sr * W[s] + sc;
                    if (masks[s][sidx] == 1) continue;
                    const float src_depth = depths[s][sidx];
                    if (src_depth <= 0.0f) continue;
                    const float* sn = &normals[s][3 * sidx];
                    float TX[3], bu, bv;
                    world_point(sc, sr, src_depth, cams[s], TX);
                    project(TX, cams[i], bu, bv, pd);
                    const float reproj = std::sqrt((c - bu) * (c - bu) + (r - bv) * (r - bv));
                    if (reproj >= 2.0f) continue;
                    const float rel = std::fabs(pd - ref_depth) / ref_depth;
                    if (rel >= 0.01f) continue;
                    float angle = std::acos(rn[0] * sn[0] + rn[1] * sn[1] + rn[2] * sn[2]);   // GetAngle, :241-250
                    if (angle != angle) angle = 0.f;
                    if (angle < 0.174533f) {
                        used_x[j] = sc; used_y[j] = sr;
                        for (int k = 0; k < 3; ++k) { sumP[k] += TX[k]; sumN[k] += sn[k]; }
                        const float sg = Scenes[s].image.px[sidx];
                        for (int k = 0; k < 3; ++k) sumC[k] += bgr[s].empty() ? sg : (float)bgr[s][3 * sidx + k];
                        dynamic_consistency += std::exp(-(reproj + 200 * rel + angle * 10));
                        ++num_consistent;
                    }
                }
                const bool keep = config.use_dynamic_consistency ? (num_consistent >= 1 && dynamic_consistency > 0.3 * num_consistent)
                                                                 : (num_consistent >= 2);
                if (!keep) continue;
                PointList p;
                // :454-460. consistent_normal is a cv::Vec3f: OpenCV's `operator/=(Vec&, float alpha)` (core/matx.hpp) multiplies by
                // the float reciprocal `1.f / alpha`; the point and the colour are plain float divisions. Checked against the
                // reference's own RunFusion compiled in place (tests/test_reference_program.py): the .ply is byte-identical.
                const float ialpha = 1.f / (num_consistent + 1.0f);
                for (int k = 0; k < 3; ++k) {
                    p.coord[k] = sumP[k] / (num_consistent + 1.0f);
                    p.normal[k] = sumN[k] * ialpha;
                    p.color[k] = sumC[k] / (num_consistent + 1.0f);
                }
                cloud.push_back(p);
                for (int j = 1; j < num_ngb; ++j) {
                    if (used_x[j] == -1) continue;
                    auto it = id2index.find(Scenes[i].srcID[j]);
                    if (it != id2index.end()) masks[it->second][(size_t)used_y[j] * W[it->second] + used_x[j]] = 1;
                }
            }
    }
    StoreColorPlyFileBinaryPointCloud(config.output_folder + "/MPMVS_model.ply", cloud);
    return cloud.size();
}

// The same stage on the GPU (mpmvs_fusion_*, pm_fusion.cu): images in the reference's order, the pixels of one image in
// parallel. Selected with --gpu-fusion; the sequential host loop above stays the default because it is the reference's
// exact order.
size_t RunFusionGPU(const ConfigParams& config, std::vector<Scene>& Scenes) {
    const int n = (int)Scenes.size();
    mpmvs_fusion* f = nullptr;
    check(mpmvs_fusion_create(0, n, &f), "mpmvs_fusion_create");
    const std::string image_folder = config.input_folder + "/images", cam_folder = config.input_folder + "/cams";
    int max_list = 2;
    for (const Scene& s : Scenes) max_list = std::max(max_list, (int)s.srcID.size() + 1);
    std::vector<int> lists((size_t)n * max_list, -2);
    for (int i = 0; i < n; ++i) {
        if (!Scenes[i].estimate) { lists[(size_t)i * max_list] = -1; continue; }
        const int id = Scenes[i].refID;
        Camera cam = ReadCamera(cam_folder + "/" + id8(id) + "_cam.txt");
        const std::string folder = config.input_folder + "/MPMVS/2333_" + id8(id);
        int h = 0, w = 0, nb;
        std::vector<float> fdepth, fnormal;
        const float *depth_px, *normal_px;
        const Scene& S = Scenes[i];
        if (S.depth && S.normal && !S.image.empty() && S.depth->size() == (size_t)S.image.width * S.image.height &&
            S.normal->size() == 3 * S.depth->size()) {          // this process produced the maps: no need to read its own files back
            w = S.image.width; h = S.image.height;
            depth_px = S.depth->data(); normal_px = S.normal->data();
        } else {
            int hn = 0, wn = 0, nbn = 0;
            if (!readDmb(folder + "/depths.dmb", h, w, nb, fdepth) || !readDmb(folder + "/normals.dmb", hn, wn, nbn, fnormal))
                throw std::runtime_error("fusion: missing results in " + folder);
            if (nb != 1 || nbn != 3 || hn != h || wn != w) throw std::runtime_error("fusion: depths.dmb and normals.dmb do not belong together in " + folder);
            depth_px = fdepth.data(); normal_px = fnormal.data();
        }
        if (Scenes[i].image.empty() && !readGrayImage(image_folder, id, Scenes[i].image)) throw std::runtime_error("fusion: missing image " + id8(id));
        const GrayImage* im = &Scenes[i].image;
        GrayImage resized;
        const int ow = Scenes[i].orig_width ? Scenes[i].orig_width : im->width, oh = Scenes[i].orig_height ? Scenes[i].orig_height : im->height;
        if (im->width != w || im->height != h) { resized = resizeLinear(*im, w, h); im = &resized; }
        if (ow != w || oh != h) { cam.K[0] *= w / (float)ow; cam.K[2] *= w / (float)ow; cam.K[4] *= h / (float)oh; cam.K[5] *= h / (float)oh; }
        cam.width = w; cam.height = h;
        std::vector<unsigned char> gray;
        const unsigned char* gray_px = im->u8.data();
        if (im->u8.size() != im->px.size()) {            // a resized image has no 8-bit original
            gray.resize(im->px.size());
            for (size_t k = 0; k < gray.size(); ++k) gray[k] = (unsigned char)std::min(255.f, std::max(0.f, std::round(im->px[k])));
            gray_px = gray.data();
        }
        check(mpmvs_fusion_set_view(f, i, &cam, depth_px, normal_px, gray_px), "mpmvs_fusion_set_view");
        {   // the colour image RunFusion averages (PatchMatch.cpp:322), when it decodes at the depth map's size
            int cw = 0, ch = 0;
            std::vector<unsigned char> col;
            if (readColorFile(image_folder + "/" + id8(id), cw, ch, col)) {
                if (cw != w || ch != h) col = resizeLinearBGR(col, cw, ch, w, h);
                check(mpmvs_fusion_set_color(f, i, col.data()), "mpmvs_fusion_set_color");
            }
        }
        if (config.sky_seg) {
            const std::vector<unsigned char> sky = readSkyMask(folder, w, h);
            if (!sky.empty()) check(mpmvs_fusion_set_sky_mask(f, i, sky.data()), "mpmvs_fusion_set_sky_mask");
        }
        for (size_t j = 0; j < Scenes[i].srcID.size(); ++j) {
            const int s = Scenes[i].srcID[j];
            lists[(size_t)i * max_list + j] = (s >= 0 && s < n && Scenes[s].estimate) ? s : -1;
        }
    }
    uint64_t npts = 0;
    float ms = 0.f;
    check(mpmvs_fusion_run(f, lists.data(), max_list, config.use_dynamic_consistency ? 1 : 0, &npts, &ms), "mpmvs_fusion_run");
    std::vector<PointList> cloud(npts);
    if (npts) check(mpmvs_fusion_get_points(f, (float*)cloud.data(), npts), "mpmvs_fusion_get_points");
    mpmvs_fusion_destroy(f);
    printf("GPU fusion kernels: %.3f ms\n", ms);
    StoreColorPlyFileBinaryPointCloud(config.output_folder + "/MPMVS_model.ply", cloud);
    return cloud.size();
}

// ------------------------------------------------------------------------------------------------ resident pipeline
// FIFO ticket for whole Run()s on one GPU. Several reference images are in flight so that the device never waits for the
// host (triangulation, copies); but Run()s that share the SMs finish together, send their images into the host stage
// together and leave the device idle together, and that lock-step is stable. Taking turns keeps the images out of phase:
// while one image's Run() has the device, the others do their host work (pipeline.GpuTurn).
class GpuTurn {
  public:
    void acquire() {
        std::unique_lock<std::mutex> l(m_);
        const uint64_t t = next_++;
        cv_.wait(l, [&] { return serving_ == t; });
    }
    void release() {
        { std::lock_guard<std::mutex> l(m_); ++serving_; }
        cv_.notify_all();
    }
  private:
    std::mutex m_;
    std::condition_variable cv_;
    uint64_t next_ = 0, serving_ = 0;
};
struct TurnGuard {
    explicit TurnGuard(GpuTurn& t) : t_(t) { t_.acquire(); }
    ~TurnGuard() { t_.release(); }
    GpuTurn& t_;
};

static void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// The stage schedule of main() (main.cpp:19-41) with everything resident on the GPUs of `devices`: reference images in
// contiguous blocks, one resident handle each, the depth maps of a pass all-gathered with peer-to-peer copies (one process
// drives every GPU with host threads; pipeline.py is the one-process-per-GPU / NCCL form of the same schedule).
// Results do not depend on the number of GPUs or of images in flight.
static void RunResident(const ConfigParams& config, std::vector<Scene>& Scenes, uint64_t seed, int tex, int in_flight,
                        const std::vector<int>& devices) {
    const int n = (int)Scenes.size();
    const auto t_setup = std::chrono::steady_clock::now();
    cuda_ok(cudaSetDevice(devices[0]), "cudaSetDevice");
    const std::string image_folder = config.input_folder + "/images", cam_folder = config.input_folder + "/cams";
    // 1. every view that some problem uses: decoded once (host threads), resized by PatchMatchInit's rule
    std::vector<char> needed(n, 0);
    for (const Scene& sc : Scenes)
        if (sc.estimate) for (int id : sc.srcID) if (id >= 0 && id < n) needed[id] = 1;
    std::vector<Camera> cams(n);
    {
        std::atomic<int> next{0};
        std::atomic<bool> failed{false};
        auto work = [&] {
            for (int i; (i = next++) < n;) {
                if (!needed[i]) continue;
                Scene& sc = Scenes[i];
                try {
                    cams[i] = ReadCamera(cam_folder + "/" + id8(i) + "_cam.txt");
                    if (sc.image.empty()) {
                        if (!readGrayImage(image_folder, i, sc.image)) throw std::runtime_error("Can not read this image ! " + id8(i));
                        sc.orig_width = sc.image.width; sc.orig_height = sc.image.height;
                        const int m = config.MaxImageSize;        // PatchMatch.cpp:893-925
                        if (sc.image.width > m || sc.image.height > m) {
                            const float factor = std::min((float)m / sc.image.width, (float)m / sc.image.height);
                            sc.image = resizeLinear(sc.image, (int)std::round(sc.image.width * factor), (int)std::round(sc.image.height * factor));
                        }
                    }
                    const float sx = sc.image.width / (float)sc.orig_width, sy = sc.image.height / (float)sc.orig_height;
                    if (sc.image.width != sc.orig_width || sc.image.height != sc.orig_height) {
                        cams[i].K[0] *= sx; cams[i].K[2] *= sx; cams[i].K[4] *= sy; cams[i].K[5] *= sy;
                    }
                    cams[i].width = sc.image.width; cams[i].height = sc.image.height;
                } catch (const std::exception& e) {
                    std::cout << e.what() << std::endl;
                    failed = true;
                }
            }
        };
        std::vector<std::thread> pool;
        const int n_dec = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        for (int t = 0; t < n_dec; ++t) pool.emplace_back(work);
        for (std::thread& t : pool) t.join();
        if (failed) throw std::runtime_error("resident pipeline: missing inputs");
    }
    int W = 0, H = 0;
    bool all_u8 = tex == MPMVS_TEX_U8;
    for (int i = 0; i < n; ++i) {
        if (!needed[i]) continue;
        const GrayImage& im = Scenes[i].image;
        if (!W) { W = im.width; H = im.height; }
        if (im.width != W || im.height != H) throw std::runtime_error("--resident needs equally sized views (one exchange buffer); use the default mode");
        all_u8 = all_u8 && im.u8.size() == im.px.size();
    }
    if (!W) return;
    // 2. reference images in contiguous blocks over the GPUs (as pipeline.shard_refs); per GPU: an image cache with the
    //    views its problems use, one resident handle per reference image, and the depth maps of ALL images of the previous
    //    pass (double-buffered: Jacobi)
    const int G = (int)devices.size();
    std::vector<int> refs, owner(n, -1);
    for (int i = 0; i < n; ++i) if (Scenes[i].estimate) refs.push_back(i);
    const size_t block = (refs.size() + G - 1) / G;
    for (size_t k = 0; k < refs.size(); ++k) owner[refs[k]] = (int)(k / block);
    const size_t wh = (size_t)W * H;
    struct Gpu { int dev = 0; mpmvs_image_cache* cache = nullptr; float* depth[2] = {nullptr, nullptr}; cudaStream_t copy = nullptr; std::vector<int> refs; };
    std::vector<Gpu> gpus(G);
    std::vector<mpmvs_problem*> handles(n, nullptr);
    auto setup_gpu = [&](int g) {
        Gpu& U = gpus[g];
        U.dev = devices[g];
        cuda_ok(cudaSetDevice(U.dev), "cudaSetDevice");
        for (int g2 = 0; g2 < G; ++g2) {                    // depth maps travel GPU to GPU (NVLink) when peer access is available
            int can = 0;
            if (g2 != g && cudaDeviceCanAccessPeer(&can, U.dev, devices[g2]) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[g2], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cuda_ok(e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
        }
        cuda_ok(cudaStreamCreateWithFlags(&U.copy, cudaStreamNonBlocking), "cudaStreamCreate");
        std::vector<char> need_g(n, 0);
        int count = 0;
        for (int i : refs) if (owner[i] == g) { U.refs.push_back(i); for (int id : Scenes[i].srcID) if (!need_g[id]) { need_g[id] = 1; ++count; } }
        if (U.refs.empty()) return;
        check(mpmvs_cache_create_fmt(U.dev, W, H, std::max(1, count), all_u8 ? MPMVS_TEX_U8 : MPMVS_TEX_F32, &U.cache), "mpmvs_cache_create_fmt");
        for (int i = 0; i < n; ++i) {
            if (!need_g[i]) continue;
            const GrayImage& im = Scenes[i].image;
            check(all_u8 ? mpmvs_cache_put_u8(U.cache, i, im.u8.data(), W, H) : mpmvs_cache_put(U.cache, i, im.px.data(), W, H), "mpmvs_cache_put");
        }
        for (int i : U.refs) {
            check(mpmvs_create(U.dev, nullptr, &handles[i]), "mpmvs_create");
            std::vector<Camera> pc;
            for (int id : Scenes[i].srcID) pc.push_back(cams[id]);
            check(mpmvs_set_views_cached(handles[i], U.cache, (int)pc.size(), Scenes[i].srcID.data(), pc.data()), "mpmvs_set_views_cached");
        }
        if (config.geom_iterations > 0)
            for (float*& d : U.depth) { cuda_ok(cudaMalloc((void**)&d, wh * 4 * n), "cudaMalloc"); cuda_ok(cudaMemset(d, 0, wh * 4 * n), "cudaMemset"); }
    };
    {   // one host thread per GPU: context creation, uploads and allocations of different devices overlap
        std::vector<std::thread> pool;
        std::vector<std::string> errs(G);
        for (int g = 0; g < G; ++g)
            pool.emplace_back([&, g] { try { setup_gpu(g); } catch (const std::exception& e) { errs[g] = e.what(); } });
        for (std::thread& t : pool) t.join();
        for (const std::string& e : errs) if (!e.empty()) throw std::runtime_error(e);
    }
    for (const Gpu& U : gpus) { cuda_ok(cudaSetDevice(U.dev), "cudaSetDevice"); cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); }
    printf("resident set-up on %d GPU(s) (decode, upload, handles): %.3f s\n", G, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_setup).count());

    const auto t0 = std::chrono::steady_clock::now();
    int cur = 0;
    const bool exchange = config.geom_iterations > 0;
    // `per_gpu` worker threads per GPU walk that GPU's images; fn(gpu, image) throws on error
    auto for_all_images = [&](int per_gpu, const std::function<void(const Gpu&, int, int)>& fn) {
        std::vector<std::atomic<size_t>> next(G);
        for (auto& x : next) x = 0;
        std::atomic<bool> failed{false};
        std::string err;
        std::mutex err_m;
        std::vector<std::thread> pool;
        for (int g = 0; g < G; ++g)
            for (int t = 0; t < std::max(1, per_gpu); ++t)
                pool.emplace_back([&, g, t, per_gpu] {
                    const Gpu& U = gpus[g];
                    cudaSetDevice(U.dev);
                    for (size_t k; (k = next[g]++) < U.refs.size();) {
                        try { fn(U, U.refs[k], g * std::max(1, per_gpu) + t); }
                        catch (const std::exception& e) { std::lock_guard<std::mutex> l(err_m); err = e.what(); failed = true; }
                    }
                });
        for (std::thread& t : pool) t.join();
        if (failed) throw std::runtime_error(err);
    };
    std::vector<GpuTurn> turns(G);
    auto run_pass = [&](int stage, bool geom, bool planar) {
        for_all_images(in_flight, [&](const Gpu& U, int i, int) {
            mpmvs_problem* h = handles[i];
            const uint64_t sd = seed + 1000003ULL * i + 7919ULL * stage;
            check(mpmvs_reset_params(h), "mpmvs_reset_params");                               // a fresh PatchMatchCUDA per call (cpp:516)
            check(mpmvs_set_geom_consistency_params(h, geom, planar), "mpmvs_set_geom_consistency_params");
            if (geom) {
                std::vector<const float*> dep;
                for (size_t j = 1; j < Scenes[i].srcID.size(); ++j) dep.push_back(U.depth[cur] + wh * Scenes[i].srcID[j]);
                check(mpmvs_set_src_depths_device(h, dep.data(), nullptr), "mpmvs_set_src_depths_device");
            }
            GpuTurn& turn = turns[&U - gpus.data()];
            if (planar) {           // a host stage follows: whole Run()s take turns (see GpuTurn)
                TurnGuard g(turn);
                check(mpmvs_run_async(h, sd), "mpmvs_run_async");
                check(mpmvs_synchronize(h), "mpmvs_synchronize");
            } else {                // no host stage: runs of different images share the device and fill each other's kernel tails
                check(mpmvs_run_async(h, sd), "mpmvs_run_async");
            }
            if (planar) {
                check(mpmvs_set_planar_prior_params(h), "mpmvs_set_planar_prior_params");
                check(mpmvs_set_geom_consistency_params(h, 0, 1), "mpmvs_set_geom_consistency_params");
                check(mpmvs_build_prior(h, nullptr), "mpmvs_build_prior");                   // host triangulation: blocks this thread only
                TurnGuard g(turn);
                check(mpmvs_run_async(h, sd ^ 0x5DEECE66DULL), "mpmvs_run_async");
                check(mpmvs_synchronize(h), "mpmvs_synchronize");
            }
            if (exchange) check(mpmvs_export_depth_device(h, U.depth[cur ^ 1] + wh * i, 0), "mpmvs_export_depth_device");
        });
        for (int i : refs) check(mpmvs_synchronize(handles[i]), "mpmvs_synchronize");
        // all-gather by peer copies: every image's new depth map goes from its owner to the other GPUs' buffers
        if (exchange && G > 1) {
            for (const Gpu& U : gpus) {
                cuda_ok(cudaSetDevice(U.dev), "cudaSetDevice");
                for (int i : U.refs)
                    for (const Gpu& V : gpus)
                        if (V.dev != U.dev && V.depth[0])
                            cuda_ok(cudaMemcpyPeerAsync(V.depth[cur ^ 1] + wh * i, V.dev, U.depth[cur ^ 1] + wh * i, U.dev, wh * 4, U.copy), "cudaMemcpyPeerAsync");
            }
            for (const Gpu& U : gpus) { cuda_ok(cudaSetDevice(U.dev), "cudaSetDevice"); cuda_ok(cudaStreamSynchronize(U.copy), "cudaStreamSynchronize"); }
        }
        cur ^= 1;
    };
    const auto stamp = [&](const char* what) {
        printf("%s done after %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    run_pass(0, false, !config.geomPlanarPrior && config.planar_prior);                       // main.cpp:19-26
    stamp("stage 1");
    for (int g = 0; g < config.geom_iterations; ++g) {                                         // main.cpp:28-41
        run_pass(1 + g, true, config.geomPlanarPrior && g != config.geom_iterations - 1);
        stamp("geometric consistency pass");
    }
    printf("cost time is %.10f us\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    // 3. results: pinned staging, split into the three maps of a result folder, files written in the background
    const int collectors = G > 1 ? 2 : 4;
    std::vector<float*> pin_planes(G * collectors, nullptr), pin_costs(G * collectors, nullptr);
    for (int g = 0; g < G; ++g) {
        cuda_ok(cudaSetDevice(gpus[g].dev), "cudaSetDevice");
        for (int t = 0; t < collectors; ++t) {
            cuda_ok(cudaMallocHost((void**)&pin_planes[g * collectors + t], wh * 16), "cudaMallocHost");
            cuda_ok(cudaMallocHost((void**)&pin_costs[g * collectors + t], wh * 4), "cudaMallocHost");
        }
    }
    for_all_images(collectors, [&](const Gpu&, int i, int slot) {
        float *pl = pin_planes[slot], *co = pin_costs[slot];
        check(mpmvs_get_results_async(handles[i], pl, co, nullptr), "mpmvs_get_results_async");
        check(mpmvs_synchronize(handles[i]), "mpmvs_synchronize");
        auto depths = std::make_shared<std::vector<float>>(wh), normals = std::make_shared<std::vector<float>>(wh * 3);
        auto costs = std::make_shared<std::vector<float>>(co, co + wh);
        float *dd = depths->data(), *nn = normals->data();
        for (size_t q = 0; q < wh; ++q) { nn[3 * q] = pl[4 * q]; nn[3 * q + 1] = pl[4 * q + 1]; nn[3 * q + 2] = pl[4 * q + 2]; dd[q] = pl[4 * q + 3]; }
        const std::string folder = config.output_folder + "/2333_" + id8(Scenes[i].refID);
        mkdir(folder.c_str(), 0777);
        SubmitDmb(folder + "/depths.dmb", H, W, 1, depths);
        SubmitDmb(folder + "/normals.dmb", H, W, 3, normals);
        SubmitDmb(folder + "/costs.dmb", H, W, 1, costs);
        Scenes[i].depth = depths; Scenes[i].normal = normals; Scenes[i].cost = costs;
    });
    printf("results collected after %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    for (float* b : pin_planes) cudaFreeHost(b);
    for (float* b : pin_costs) cudaFreeHost(b);
    for (mpmvs_problem* h : handles) if (h) mpmvs_destroy(h);
    for (Gpu& U : gpus) {
        cudaSetDevice(U.dev);
        for (float* d : U.depth) cudaFree(d);
        if (U.copy) cudaStreamDestroy(U.copy);
        if (U.cache) mpmvs_cache_destroy(U.cache);
    }
    cudaSetDevice(devices[0]);
}

// --check-inputs DIR: parse everything the run would parse (config.yaml, pair.txt, cameras, images incl. PatchMatchInit's
// resize rule) WITHOUT touching the GPU, print it as JSON and write every (resized) image as DIR/%08d.dmb. The CPU test
// suite compares this with the Python readers and with cv2.resize (tests/test_cpp_host.py).
static void CheckInputs(const ConfigParams& c, std::vector<Scene>& Scenes, const std::string& dump_dir) {
    auto arr = [](const float* v, int n) {
        std::ostringstream o;
        o << std::setprecision(9) << "[";
        for (int i = 0; i < n; ++i) o << (i ? ", " : "") << v[i];
        o << "]";
        return o.str();
    };
    std::cout << "{\"config\": {\"input_folder\": \"" << c.input_folder << "\", \"output_folder\": \"" << c.output_folder << "\", \"geom_iterations\": "
              << c.geom_iterations << ", \"planar_prior\": " << c.planar_prior << ", \"geomPlanarPrior\": " << c.geomPlanarPrior << ", \"sky_seg\": " << c.sky_seg
              << ", \"use_dynamic_consistency\": " << c.use_dynamic_consistency << ", \"saveDmb\": " << c.saveDmb << ", \"saveProirDmb\": " << c.saveProirDmb
              << ", \"saveCostDmb\": " << c.saveCostDmb << ", \"saveNormalDmb\": " << c.saveNormalDmb << ", \"MaxSourceImageNum\": " << c.MaxSourceImageNum
              << ", \"MaxImageSize\": " << c.MaxImageSize << "},\n \"scenes\": [";
    bool first = true;
    for (size_t i = 0; i < Scenes.size(); ++i) {
        Scene& sc = Scenes[i];
        std::cout << (first ? "" : ",") << "\n  {\"index\": " << i << ", \"estimate\": " << sc.estimate << ", \"refID\": " << sc.refID << ", \"srcID\": [";
        first = false;
        for (size_t j = 0; j < sc.srcID.size(); ++j) std::cout << (j ? ", " : "") << sc.srcID[j];
        std::cout << "]";
        if (sc.estimate) {
            Camera cam = ReadCamera(c.input_folder + "/cams/" + id8(sc.refID) + "_cam.txt");
            GrayImage im;
            if (!readGrayFile(c.input_folder + "/images/" + id8(sc.refID), im)) throw std::runtime_error("Can not read this image ! " + id8(sc.refID));
            const int ow = im.width, oh = im.height;
            if (im.width > c.MaxImageSize || im.height > c.MaxImageSize) {       // PatchMatch.cpp:893-925
                const float factor = std::min((float)c.MaxImageSize / im.width, (float)c.MaxImageSize / im.height);
                im = resizeLinear(im, (int)std::round(im.width * factor), (int)std::round(im.height * factor));
                const float sx = im.width / (float)ow, sy = im.height / (float)oh;
                cam.K[0] *= sx; cam.K[2] *= sx; cam.K[4] *= sy; cam.K[5] *= sy;
            }
            writeDmb(dump_dir + "/" + id8(sc.refID) + ".dmb", im.height, im.width, 1, im.px.data());
            std::cout << ", \"K\": " << arr(cam.K, 9) << ", \"R\": " << arr(cam.R, 9) << ", \"t\": " << arr(cam.t, 3) << ", \"C\": " << arr(cam.C, 3)
                      << ", \"depth_min\": " << std::setprecision(9) << cam.depth_min << ", \"depth_max\": " << cam.depth_max << ", \"width\": " << im.width
                      << ", \"height\": " << im.height << ", \"orig_width\": " << ow << ", \"orig_height\": " << oh;
        }
        std::cout << "}";
    }
    std::cout << "\n ]}" << std::endl;
}

}  // namespace mpmvs

int main(int argc, char* argv[]) {
    using namespace mpmvs;
    std::string yaml = "config/config.yaml";
    uint64_t seed = 0x2333;
    int tex = MPMVS_TEX_F32;
    bool fusion = true, gpu_fusion = false, profile = false, resident = false, fusion_only = false;
    std::string check_dir;
    int in_flight = 8, device = 0, n_gpus = 1;   // host threads per GPU: the triangulation of the planar prior (up to 0.4 s per image) is host work
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 0);
        else if (!strcmp(argv[i], "--tex") && i + 1 < argc) tex = !strcmp(argv[++i], "u8") ? MPMVS_TEX_U8 : MPMVS_TEX_F32;
        else if (!strcmp(argv[i], "--no-fusion")) fusion = false;
        else if (!strcmp(argv[i], "--gpu-fusion")) gpu_fusion = true;
        else if (!strcmp(argv[i], "--profile")) profile = true;
        else if (!strcmp(argv[i], "--fusion-only")) fusion_only = true;     // fuse the .dmb files already in the result folders
        else if (!strcmp(argv[i], "--check-inputs") && i + 1 < argc) check_dir = argv[++i];
        else if (!strcmp(argv[i], "--resident")) resident = true;
        else if (!strcmp(argv[i], "--in-flight") && i + 1 < argc) in_flight = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) n_gpus = atoi(argv[++i]);
        // both arithmetics are in the one library; every handle created from here on starts with this one (mpmvs_default_arithmetic)
        else if (!strcmp(argv[i], "--arithmetic") && i + 1 < argc) setenv("MPMVS_ARITHMETIC", argv[++i], 1);
        else yaml = argv[i];
    }
    try {
        ConfigParams config = readConfig(yaml);
        if (!check_dir.empty()) {
            std::vector<Scene> Scenes;
            GenerateSampleList(config, Scenes);
            CheckInputs(config, Scenes, check_dir);
            return 0;
        }
        std::cout << "Input data path:" << config.input_folder << "\nOutput data path:" << config.output_folder << std::endl;
        std::cout << "Kernels: " << mpmvs_arithmetic_name(mpmvs_default_arithmetic()) << " arithmetic"
                  << (mpmvs_default_arithmetic() == MPMVS_ARITH_EXACT
                          ? (tex == MPMVS_TEX_F32 ? " (bit-identical to the reference's kernels; --arithmetic fast --tex u8 is faster, statistically equal)"
                                                  : " on 8-bit views (the bilinear filter differs in the last bits: use --tex f32 for bit-identity)")
                          : " (statistically equal to the reference's kernels; --arithmetic exact is bit-identical)")
                  << std::endl;
        mkdir(config.output_folder.c_str(), 0777);
        std::vector<Scene> Scenes;
        GenerateSampleList(config, Scenes);
        const int num_img = (int)Scenes.size();
        std::cout << "There are " << num_img << " depthmaps need to be computed!\n" << std::endl;
        if (!resident && !fusion_only)
            std::cout << "(reference order: one image at a time, results exchanged in place; --resident [--gpus G] keeps every view, state "
                         "and depth map on the GPU(s) and is 2-3x faster)\n" << std::endl;
        const auto t0 = std::chrono::steady_clock::now();
        if (fusion_only) {
            // nothing to estimate: RunFusion below reads depths.dmb / normals.dmb of every image (host fusion needs no GPU)
        } else if (resident) {
            mkdir(config.output_folder.c_str(), 0777);
            std::vector<int> devices;
            for (int g = 0; g < std::max(1, n_gpus); ++g) devices.push_back(device + g);
            RunResident(config, Scenes, seed, tex, in_flight, devices);
        } else {
            // stage 1: multi-scale-window PatchMatch (main.cpp:19-26)
            bool planar_prior = !config.geomPlanarPrior && config.planar_prior;
            for (int i = 0; i < num_img; ++i)
                if (Scenes[i].estimate) ProcessProblem(config.input_folder, config.output_folder, Scenes, i, false, planar_prior, seed + 1000003ULL * i, tex);
            // stage 2: geometric consistency [+ planar prior] (main.cpp:28-41)
            for (int g = 0; g < config.geom_iterations; ++g) {
                planar_prior = config.geomPlanarPrior && g != config.geom_iterations - 1;
                for (int i = 0; i < num_img; ++i)
                    if (Scenes[i].estimate)
                        ProcessProblem(config.input_folder, config.output_folder, Scenes, i, true, planar_prior, seed + 1000003ULL * i + 7919ULL * (g + 1), tex);
            }
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            printf("cost time is %.10f us\n", us);
        }
        if (config.sky_seg && !fusion_only) {               // main.cpp:44-46
            const int n_sky = GenerateSkyRegionMask(Scenes, config);
            std::cout << "refined " << n_sky << " sky masks" << std::endl;
        }
        FlushDmbWriters();
        const double us_files = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        printf("results on disk after %.10f us\n", us_files);
        if (profile) {
            const HostPhaseTimes& T = PhaseTimes();
            printf("host phases (s): init %.3f upload %.3f run %.3f prior %.3f collect %.3f write %.3f\n", T.init, T.upload, T.run, T.prior, T.collect, T.write);
        }
        if (fusion) {
            const auto f0 = std::chrono::steady_clock::now();
            const size_t npts = gpu_fusion ? RunFusionGPU(config, Scenes) : RunFusion(config, Scenes);
            const double fus = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - f0).count();
            printf("fusion time is %.10f us (%s)\n", fus, gpu_fusion ? "GPU kernels + file I/O" : "host, 1 thread");
            std::cout << "store 3D points to ply file: " << npts << " points" << std::endl;
        }
        PatchMatchCUDA::ReleasePool();
    } catch (const std::exception& e) {
        std::cout << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return 0;
}
