// mpmvs_main.cpp -- the reference's entry point (/root/reference/src/main.cpp:6-55) over the C++ host mirror:
// config.yaml -> pair.txt -> stage 1 (multi-scale photometric [+ planar prior]) -> geometric-consistency iterations
// [+ planar prior] -> fusion -> MPMVS_model.ply, on the same dense-folder layout and with the same output files.
//
//   mpmvs_main [config.yaml] [--seed N] [--tex f32|u8] [--no-fusion] [--gpu-fusion]
//
// The reference bakes the config path in at cmake time (include/ProjectPath.h.in) and takes no arguments.
#include <sys/stat.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdlib>
#include <cstring>

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
// gate of its BUILD_NCNN branch included when `Sky segment` is set). Colour: the grey image replicated to three channels.
size_t RunFusion(const ConfigParams& config, std::vector<Scene>& Scenes) {
    const size_t n = Scenes.size();
    std::vector<Camera> cams(n);
    std::vector<std::vector<float>> depths(n), normals(n);
    std::vector<std::vector<unsigned char>> masks(n), sky(n);
    std::vector<int> W(n, 0), H(n, 0);
    const std::string image_folder = config.input_folder + "/images", cam_folder = config.input_folder + "/cams";
    for (size_t i = 0; i < n; ++i) {
        if (!Scenes[i].estimate) continue;
        const int id = Scenes[i].refID;
        cams[i] = ReadCamera(cam_folder + "/" + id8(id) + "_cam.txt");
        const std::string folder = config.input_folder + "/MPMVS/2333_" + id8(id);
        int h, w, nb;
        if (!readDmb(folder + "/depths.dmb", h, w, nb, depths[i]) || !readDmb(folder + "/normals.dmb", h, w, nb, normals[i]))
            throw std::runtime_error("fusion: missing results in " + folder);
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
                        sumC[0] += sg; sumC[1] += sg; sumC[2] += sg;
                        dynamic_consistency += std::exp(-(reproj + 200 * rel + angle * 10));
                        ++num_consistent;
                    }
                }
                const bool keep = config.use_dynamic_consistency ? (num_consistent >= 1 && dynamic_consistency > 0.3 * num_consistent)
                                                                 : (num_consistent >= 2);
                if (!keep) continue;
                PointList p;
                for (int k = 0; k < 3; ++k) {
                    p.coord[k] = sumP[k] / (num_consistent + 1.0f);
                    p.normal[k] = sumN[k] / (num_consistent + 1.0f);
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
        int h, w, nb;
        std::vector<float> depth, normal;
        if (!readDmb(folder + "/depths.dmb", h, w, nb, depth) || !readDmb(folder + "/normals.dmb", h, w, nb, normal))
            throw std::runtime_error("fusion: missing results in " + folder);
        if (Scenes[i].image.empty() && !readGrayImage(image_folder, id, Scenes[i].image)) throw std::runtime_error("fusion: missing image " + id8(id));
        GrayImage im = Scenes[i].image;
        const int ow = Scenes[i].orig_width ? Scenes[i].orig_width : im.width, oh = Scenes[i].orig_height ? Scenes[i].orig_height : im.height;
        if (im.width != w || im.height != h) im = resizeLinear(im, w, h);
        if (ow != w || oh != h) { cam.K[0] *= w / (float)ow; cam.K[2] *= w / (float)ow; cam.K[4] *= h / (float)oh; cam.K[5] *= h / (float)oh; }
        cam.width = w; cam.height = h;
        std::vector<unsigned char> gray(im.px.size());
        for (size_t k = 0; k < gray.size(); ++k) gray[k] = (unsigned char)std::min(255.f, std::max(0.f, std::round(im.px[k])));
        check(mpmvs_fusion_set_view(f, i, &cam, depth.data(), normal.data(), gray.data()), "mpmvs_fusion_set_view");
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

}  // namespace mpmvs

int main(int argc, char* argv[]) {
    using namespace mpmvs;
    std::string yaml = "config/config.yaml";
    uint64_t seed = 0x2333;
    int tex = MPMVS_TEX_F32;
    bool fusion = true, gpu_fusion = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 0);
        else if (!strcmp(argv[i], "--tex") && i + 1 < argc) tex = !strcmp(argv[++i], "u8") ? MPMVS_TEX_U8 : MPMVS_TEX_F32;
        else if (!strcmp(argv[i], "--no-fusion")) fusion = false;
        else if (!strcmp(argv[i], "--gpu-fusion")) gpu_fusion = true;
        else yaml = argv[i];
    }
    try {
        ConfigParams config = readConfig(yaml);
        std::cout << "Input data path:" << config.input_folder << "\nOutput data path:" << config.output_folder << std::endl;
        mkdir(config.output_folder.c_str(), 0777);
        std::vector<Scene> Scenes;
        GenerateSampleList(config, Scenes);
        const int num_img = (int)Scenes.size();
        std::cout << "There are " << num_img << " depthmaps need to be computed!\n" << std::endl;
        const auto t0 = std::chrono::steady_clock::now();
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
        if (config.sky_seg) {                               // main.cpp:44-46
            const int n_sky = GenerateSkyRegionMask(Scenes, config);
            std::cout << "refined " << n_sky << " sky masks" << std::endl;
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
