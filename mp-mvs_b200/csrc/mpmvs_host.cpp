// mpmvs_host.cpp -- implementation of the C++ host mirror (PatchMatchCUDA.h): file formats, the class members that
// forward to the C ABI, and ProcessProblem. Plain C++17; JPEG decoding (only when no decoded .pgm sidecar exists) uses
// nvJPEG from the CUDA toolkit -- there is no OpenCV here.
#include "PatchMatchCUDA.h"

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

#ifdef MPMVS_WITH_NVJPEG
#include <cuda_runtime.h>
#include <nvjpeg.h>
#endif

namespace mpmvs {

std::string id8(int id) {
    std::ostringstream s;
    s << std::setw(8) << std::setfill('0') << id;
    return s.str();
}

// ---------------------------------------------------------------------------------------------- config.yaml
// The reference reads the file with cv::FileStorage (utility.cpp:10-26). The keys contain spaces, the values are
// scalars or quoted strings: a line-based `key: value` reader covers what FileStorage accepts for this file.
ConfigParams readConfig(const std::string& yaml_path) {
    std::ifstream f(yaml_path);
    if (!f.is_open()) throw std::runtime_error("can not open config file: " + yaml_path);
    std::map<std::string, std::string> kv;
    std::string line;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (line.rfind("%YAML", 0) == 0 || line.rfind("---", 0) == 0) continue;
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        auto trim = [](std::string s) {
            const char* ws = " \t\r\n\"";
            const size_t a = s.find_first_not_of(ws), b = s.find_last_not_of(ws);
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        };
        kv[trim(line.substr(0, colon))] = trim(line.substr(colon + 1));
    }
    auto geti = [&](const char* k, int d) { auto it = kv.find(k); return it == kv.end() || it->second.empty() ? d : std::stoi(it->second); };
    ConfigParams c;
    c.input_folder = kv["Input-folder"];
    c.output_folder = kv["Output-folder"];
    c.geom_iterations = geti("Geometric consistency iterations", 2);
    c.planar_prior = geti("Planer prior", 1) != 0;
    c.geomPlanarPrior = geti("Geometric consistency planer prior", 1) != 0;
    c.sky_seg = geti("Sky segment", 0) != 0;
    c.use_dynamic_consistency = geti("Use dynamic_consistency to fuse", 1) != 0;
    c.saveDmb = geti("Save Dmb as JPG", 0) != 0;
    c.saveProirDmb = geti("Save Prior Dmb as JPG", 0) != 0;
    c.saveCostDmb = geti("Save Cost Map", 0) != 0;
    c.saveNormalDmb = geti("Save Normal Map", 0) != 0;
    c.MaxSourceImageNum = geti("Max source images num", 20);
    c.MaxImageSize = geti("Max image size", 3200);
    auto checkpath = [](std::string& p) { if (!p.empty() && p.back() == '/') p.pop_back(); };   // utility.cpp:3-6
    checkpath(c.input_folder);
    checkpath(c.output_folder);
    c.output_folder += "/MPMVS";                                                                 // utility.cpp:30
    return c;
}

// ---------------------------------------------------------------------------------------------- pair.txt, cams
void GenerateSampleList(const ConfigParams& config, std::vector<Scene>& Scenes) {
    Scenes.clear();
    const std::string path = config.input_folder + "/pair.txt";
    std::ifstream file(path);
    if (!file.is_open()) throw std::runtime_error("can not open file in path: " + path);
    int num_images = 0;
    file >> num_images;
    if (!file || num_images < 0) throw std::runtime_error("malformed pair.txt (image count): " + path);
    for (int i = 0; i < num_images; ++i) {
        Scene scene;
        scene.max_image_size = config.MaxImageSize;
        file >> scene.refID;
        // ids index `Scenes` (PatchMatch.cpp:871-890 does Scenes[srcID[i]] unchecked): a truncated or out-of-order file
        // must end in a message, not in an out-of-bounds access
        if (!file || scene.refID < (int)Scenes.size() || scene.refID > (1 << 24)) throw std::runtime_error("malformed pair.txt (reference id of entry " + std::to_string(i) + "): " + path);
        scene.srcID.push_back(scene.refID);
        while (scene.refID > (int)Scenes.size()) Scenes.emplace_back();     // gaps in the ids: estimate == false
        int num_src = 0;
        file >> num_src;
        if (!file || num_src < 0) throw std::runtime_error("malformed pair.txt (source count of image " + std::to_string(scene.refID) + "): " + path);
        for (int j = 0; j < num_src; ++j) {
            int id = -1; float score = 0.f;
            file >> id >> score;
            if (!file || id < 0) throw std::runtime_error("malformed pair.txt (sources of image " + std::to_string(scene.refID) + "): " + path);
            if (score <= 0.0f) continue;
            if (j < config.MaxSourceImageNum) scene.srcID.push_back(id);
        }
        scene.estimate = num_src != 0;
        Scenes.push_back(std::move(scene));
    }
    for (const Scene& sc : Scenes)
        for (int id : sc.srcID)
            if (id < 0 || id >= (int)Scenes.size())
                throw std::runtime_error("pair.txt names image " + std::to_string(id) + ", which has no entry of its own: " + path);
}

Camera ReadCamera(const std::string& cam_path) {
    std::ifstream file(cam_path);
    if (!file.is_open()) throw std::runtime_error("can not open file in path: " + cam_path);
    Camera cam{};
    std::string word;
    file >> word;                                   // "extrinsic"
    for (int i = 0; i < 3; ++i) file >> cam.R[3 * i] >> cam.R[3 * i + 1] >> cam.R[3 * i + 2] >> cam.t[i];
    float last_row[4];
    file >> last_row[0] >> last_row[1] >> last_row[2] >> last_row[3];
    file >> word;                                   // "intrinsic"
    for (int i = 0; i < 9; ++i) file >> cam.K[i];
    for (int i = 0; i < 3; ++i) cam.C[i] = -(cam.R[i] * cam.t[0] + cam.R[3 + i] * cam.t[1] + cam.R[6 + i] * cam.t[2]);   // C = -R^T t
    float interval, depth_num;
    file >> cam.depth_min >> interval >> depth_num >> cam.depth_max;
    if (!file) throw std::runtime_error("malformed camera file: " + cam_path);
    return cam;
}

// ---------------------------------------------------------------------------------------------- .dmb
bool readDmb(const std::string& path, int& h, int& w, int& nb, std::vector<float>& data) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { std::cout << "Error opening file " << path << std::endl; return false; }
    int32_t hdr[4] = {-1, 0, 0, 0};
    const bool ok = fread(hdr, sizeof(int32_t), 4, f) == 4 && hdr[0] == 1 && hdr[1] > 0 && hdr[2] > 0 && hdr[3] > 0 &&
                    hdr[3] <= 4 && (int64_t)hdr[1] * hdr[2] <= (int64_t)1 << 28;     // a depth / normal / cost map, not garbage
    if (ok) {
        h = hdr[1]; w = hdr[2]; nb = hdr[3];
        data.resize((size_t)h * w * nb);
        if (fread(data.data(), sizeof(float), data.size(), f) != data.size()) { fclose(f); return false; }
    }
    fclose(f);
    return ok;
}

bool writeDmb(const std::string& path, int h, int w, int nb, const float* data) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { std::cout << "Error opening file " << path << std::endl; return false; }
    const int32_t hdr[4] = {1, h, w, nb};
    const size_t count = (size_t)h * w * nb;
    bool ok = fwrite(hdr, sizeof(int32_t), 4, f) == 4 && fwrite(data, sizeof(float), count, f) == count;
    ok = fclose(f) == 0 && ok;          // a full disk shows up here at the latest
    if (!ok) std::cout << "Error writing file " << path << std::endl;
    return ok;
}

// ---------------------------------------------------------------------------------------------- background writers
namespace {
struct DmbJob { std::string path; int h, w, nb; std::shared_ptr<const std::vector<float>> data; };
class DmbWriterPool {
  public:
    explicit DmbWriterPool(int n) { for (int i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); }); }
    ~DmbWriterPool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (std::thread& t : threads_) t.join();
    }
    void submit(DmbJob j) {
        { std::lock_guard<std::mutex> l(m_); q_.push_back(std::move(j)); ++pending_; }
        cv_.notify_one();
    }
    void flush() {
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
        if (failed_) { failed_ = false; throw std::runtime_error("writing a .dmb file failed"); }
    }
  private:
    void loop() {
        for (;;) {
            DmbJob j;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                j = std::move(q_.front());
                q_.pop_front();
            }
            const bool ok = writeDmb(j.path, j.h, j.w, j.nb, j.data->data());
            std::lock_guard<std::mutex> l(m_);
            if (!ok) failed_ = true;
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> threads_;
    std::deque<DmbJob> q_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    int pending_ = 0;
    bool stop_ = false, failed_ = false;
};
DmbWriterPool& writers() { static DmbWriterPool pool(4); return pool; }
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

void SubmitDmb(const std::string& path, int h, int w, int nb, std::shared_ptr<const std::vector<float>> data) {
    writers().submit(DmbJob{path, h, w, nb, std::move(data)});
}
void FlushDmbWriters() { writers().flush(); }
HostPhaseTimes& PhaseTimes() { static HostPhaseTimes t; return t; }

// ---------------------------------------------------------------------------------------------- images
static bool readPgm(const std::string& path, GrayImage& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    f >> magic >> w >> h >> maxv;
    if (magic != "P5" || maxv != 255 || w <= 0 || h <= 0) return false;
    f.get();
    std::vector<unsigned char> buf((size_t)w * h);
    f.read((char*)buf.data(), buf.size());
    if ((size_t)f.gcount() != buf.size()) return false;
    out.width = w; out.height = h;
    out.px.resize(buf.size());
    for (size_t i = 0; i < buf.size(); ++i) out.px[i] = (float)buf[i];      // convertTo(CV_32FC1), PatchMatch.cpp:882
    out.u8 = std::move(buf);
    return true;
}

#ifdef MPMVS_WITH_NVJPEG
static bool readJpegLuma(const std::string& path, GrayImage& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static nvjpegHandle_t handle = nullptr;
    static nvjpegJpegState_t state = nullptr;
    static std::mutex decoder_mutex;               // one decoder state: callers on several threads take turns
    std::lock_guard<std::mutex> lock(decoder_mutex);
    if (!handle) {
        nvjpegHandle_t hnd = nullptr;
        if (nvjpegCreateSimple(&hnd) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegJpegStateCreate(hnd, &state) != NVJPEG_STATUS_SUCCESS) { nvjpegDestroy(hnd); return false; }
        handle = hnd;                              // published only together with its state
    }
    int comps = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (nvjpegGetImageInfo(handle, bytes.data(), bytes.size(), &comps, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return false;
    const int w = ws[0], h = hs[0];
    nvjpegImage_t img{};
    if (cudaMalloc((void**)&img.channel[0], (size_t)w * h) != cudaSuccess) return false;
    img.pitch[0] = (size_t)w;
    // luma plane = what cv::imread(IMREAD_GRAYSCALE) returns for a JPEG (libjpeg JCS_GRAYSCALE), up to IDCT rounding
    const bool ok = nvjpegDecode(handle, state, bytes.data(), bytes.size(), NVJPEG_OUTPUT_Y, &img, 0) == NVJPEG_STATUS_SUCCESS;
    std::vector<unsigned char> buf((size_t)w * h);
    if (ok) cudaMemcpy(buf.data(), img.channel[0], buf.size(), cudaMemcpyDeviceToHost);
    cudaFree(img.channel[0]);
    if (!ok) return false;
    out.width = w; out.height = h;
    out.px.resize(buf.size());
    for (size_t i = 0; i < buf.size(); ++i) out.px[i] = (float)buf[i];
    out.u8 = std::move(buf);
    return true;
}
#endif

bool readGrayFile(const std::string& path_without_ext, GrayImage& out) {
    if (readPgm(path_without_ext + ".pgm", out)) return true;
#ifdef MPMVS_WITH_NVJPEG
    if (readJpegLuma(path_without_ext + ".jpg", out)) return true;
#endif
    return false;
}

bool readGrayImage(const std::string& image_folder, int id, GrayImage& out) { return readGrayFile(image_folder + "/" + id8(id), out); }

bool writePgm(const std::string& path, int w, int h, const unsigned char* px) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P5\n%d %d\n255\n", w, h);
    const bool ok = fwrite(px, 1, (size_t)w * h, f) == (size_t)w * h;
    fclose(f);
    return ok;
}

// cv::imread(path, 1): interleaved B, G, R. Sources: a binary PPM sidecar (R, G, B on disk) or the JPEG through nvJPEG.
bool readColorFile(const std::string& path_without_ext, int& w, int& h, std::vector<unsigned char>& bgr) {
    {
        std::ifstream f(path_without_ext + ".ppm", std::ios::binary);
        std::string magic;
        int maxv = 0;
        if (f.is_open() && (f >> magic >> w >> h >> maxv) && magic == "P6" && maxv == 255 && w > 0 && h > 0) {
            f.get();
            bgr.resize((size_t)w * h * 3);
            f.read((char*)bgr.data(), bgr.size());
            if ((size_t)f.gcount() == bgr.size()) {
                for (size_t i = 0; i < bgr.size(); i += 3) std::swap(bgr[i], bgr[i + 2]);
                return true;
            }
        }
    }
#ifdef MPMVS_WITH_NVJPEG
    std::ifstream f(path_without_ext + ".jpg", std::ios::binary);
    if (!f.is_open()) return false;
    std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    if (nvjpegCreateSimple(&handle) != NVJPEG_STATUS_SUCCESS) return false;
    if (nvjpegJpegStateCreate(handle, &state) != NVJPEG_STATUS_SUCCESS) { nvjpegDestroy(handle); return false; }
    int comps = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    bool ok = nvjpegGetImageInfo(handle, bytes.data(), bytes.size(), &comps, &ss, ws, hs) == NVJPEG_STATUS_SUCCESS;
    nvjpegImage_t img{};
    const bool grey = ok && comps == 1;            // a single-component JPEG: cv::imread(path, 1) replicates the luma plane
    if (ok) {
        w = ws[0]; h = hs[0];
        ok = cudaMalloc((void**)&img.channel[0], (size_t)w * h * (grey ? 1 : 3)) == cudaSuccess;
        img.pitch[0] = (size_t)w * (grey ? 1 : 3);
    }
    if (ok) ok = nvjpegDecode(handle, state, bytes.data(), bytes.size(), grey ? NVJPEG_OUTPUT_Y : NVJPEG_OUTPUT_BGRI, &img, 0) == NVJPEG_STATUS_SUCCESS;
    if (ok && grey) {
        std::vector<unsigned char> y((size_t)w * h);
        ok = cudaMemcpy(y.data(), img.channel[0], y.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
        bgr.resize(y.size() * 3);
        for (size_t i = 0; i < y.size(); ++i) bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = y[i];
    } else if (ok) {
        bgr.resize((size_t)w * h * 3);
        ok = cudaMemcpy(bgr.data(), img.channel[0], bgr.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    cudaFree(img.channel[0]);
    nvjpegJpegStateDestroy(state);
    nvjpegDestroy(handle);
    return ok;
#else
    return false;
#endif
}

// GenerateSkyRegionMask, PatchMatch.cpp:4-57, without the segmentation network (ncnn): the coarse map it would produce is
// read from <out>/MPMVS/2333_%08d/skymask.{pgm,jpg} (255 x probability, any size -- the file the reference itself writes at
// cpp:42-44), refined by joint-bilateral upsampling on the GPU (mpmvs_sky_mask_refine; bilateral_filter, cpp:46-47) and
// stored as skymask_refine.pgm for RunFusion. Images above `Max image size` would need cv::resize on 8-bit colour
// (fixed-point arithmetic, not restated here): they are skipped with a message -- mp-mvs_b200/run.py handles them via cv2.
int GenerateSkyRegionMask(std::vector<Scene>& Scenes, const ConfigParams& config) {
    int done = 0;
    for (Scene& sc : Scenes) {
        if (!sc.estimate) continue;
        const std::string folder = config.output_folder + "/2333_" + id8(sc.refID);   // output_folder already ends in /MPMVS (readConfig)
        GrayImage coarse;
        if (!readGrayFile(folder + "/skymask", coarse)) continue;
        int w = 0, h = 0;
        std::vector<unsigned char> bgr;
        if (!readColorFile(config.input_folder + "/images/" + id8(sc.refID), w, h, bgr)) {
            std::cout << "Can not read this image ! " << id8(sc.refID) << std::endl;
            continue;
        }
        if (w > config.MaxImageSize || h > config.MaxImageSize) {
            std::cout << "sky mask of image " << id8(sc.refID) << " skipped: image above Max image size (use mp-mvs_b200/run.py)" << std::endl;
            continue;
        }
        std::vector<float> prob(coarse.px.size()), refined((size_t)w * h);
        for (size_t i = 0; i < prob.size(); ++i) prob[i] = coarse.px[i] / 255.0f;
        check(mpmvs_sky_mask_refine(0, nullptr, bgr.data(), w, h, prob.data(), coarse.width, coarse.height, refined.data(), nullptr, nullptr),
              "mpmvs_sky_mask_refine");
        std::vector<unsigned char> u8(refined.size());
        for (size_t i = 0; i < u8.size(); ++i) u8[i] = refined[i] > 0.f ? 255 : 0;       // convertTo(CV_8U), cpp:48-50
        if (!writePgm(folder + "/skymask_refine.pgm", w, h, u8.data())) throw std::runtime_error("can not write " + folder + "/skymask_refine.pgm");
        ++done;
    }
    return done;
}

// The sky gate RunFusion reads (PatchMatch.cpp:358-373): skymask_refine brought to the depth map's size; empty = none.
std::vector<unsigned char> readSkyMask(const std::string& folder, int w, int h) {
    GrayImage m;
    std::vector<unsigned char> out;
    if (!readGrayFile(folder + "/skymask_refine", m)) return out;
    if (m.width != w || m.height != h) m = resizeLinear(m, w, h);
    out.resize((size_t)w * h);
    for (size_t i = 0; i < out.size(); ++i) out[i] = m.px[i] > 0.f ? 255 : 0;
    return out;
}

// cv::resize(src, dst, Size(new_cols, new_rows), 0, 0, INTER_LINEAR) for CV_32FC1 (PatchMatch.cpp:915): pixel-centre
// mapping sx = (dx + 0.5) * scale - 0.5, source index clamped to the image, separable linear weights.
GrayImage resizeLinear(const GrayImage& src, int new_cols, int new_rows) {
    GrayImage dst;                                  // (no 8-bit copy: resized pixels are not integers)
    dst.width = new_cols; dst.height = new_rows;
    dst.px.resize((size_t)new_cols * new_rows);
    const double sx = (double)src.width / new_cols, sy = (double)src.height / new_rows;
    std::vector<int> x0(new_cols);
    std::vector<float> ax(new_cols);
    for (int dx = 0; dx < new_cols; ++dx) {
        float fx = (float)((dx + 0.5) * sx - 0.5);
        int ix = (int)std::floor(fx);
        fx -= ix;
        if (ix < 0) { ix = 0; fx = 0.f; }
        if (ix >= src.width - 1) { ix = src.width - 1; fx = 0.f; }
        x0[dx] = ix; ax[dx] = fx;
    }
    for (int dy = 0; dy < new_rows; ++dy) {
        float fy = (float)((dy + 0.5) * sy - 0.5);
        int iy = (int)std::floor(fy);
        fy -= iy;
        if (iy < 0) { iy = 0; fy = 0.f; }
        if (iy >= src.height - 1) { iy = src.height - 1; fy = 0.f; }
        const float* r0 = &src.px[(size_t)iy * src.width];
        const float* r1 = &src.px[(size_t)std::min(iy + 1, src.height - 1) * src.width];
        for (int dx = 0; dx < new_cols; ++dx) {
            const int i0 = x0[dx], i1 = std::min(i0 + 1, src.width - 1);
            const float a = ax[dx];
            const float top = r0[i0] * (1.f - a) + r0[i1] * a, bot = r1[i0] * (1.f - a) + r1[i1] * a;
            dst.px[(size_t)dy * new_cols + dx] = top * (1.f - fy) + bot * fy;
        }
    }
    return dst;
}

// cv::resize(src, dst, Size(new_cols, new_rows), 0, 0, INTER_LINEAR) for CV_8UC3 (RescaleImageAndCamera, PatchMatch.cpp:277):
// OpenCV's 8-bit path is fixed point -- the two weights of a sample rounded to 11 bits, the horizontal pass kept as a 32-bit
// integer, the vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16), rounded by (+2) >> 2 -- with the same pixel-centre
// mapping and border rule as the float path above. Byte-identical to cv2.resize of an OpenCV build without IPP
// (tests/test_cpp_host.py); Intel IPP, where OpenCV is built with it, answers up to 1 grey level differently.
std::vector<unsigned char> resizeLinearBGR(const std::vector<unsigned char>& src, int width, int height, int new_cols, int new_rows) {
    std::vector<unsigned char> dst((size_t)new_cols * new_rows * 3);
    const double sx = (double)width / new_cols, sy = (double)height / new_rows;
    std::vector<int> x0(new_cols), a0(new_cols), a1(new_cols);
    for (int dx = 0; dx < new_cols; ++dx) {
        float fx = (float)((dx + 0.5) * sx - 0.5);
        int ix = (int)std::floor(fx);
        fx -= ix;
        if (ix < 0) { ix = 0; fx = 0.f; }
        if (ix >= width - 1) { ix = width - 1; fx = 0.f; }
        x0[dx] = ix;
        a0[dx] = (int)std::lrint((1.f - fx) * 2048.f);
        a1[dx] = (int)std::lrint(fx * 2048.f);
    }
    std::vector<int> row0((size_t)new_cols * 3), row1((size_t)new_cols * 3);
    auto hpass = [&](int iy, std::vector<int>& out) {
        const unsigned char* r = &src[(size_t)iy * width * 3];
        for (int dx = 0; dx < new_cols; ++dx) {
            const int i0 = x0[dx], i1 = std::min(i0 + 1, width - 1);
            for (int k = 0; k < 3; ++k) out[(size_t)dx * 3 + k] = r[i0 * 3 + k] * a0[dx] + r[i1 * 3 + k] * a1[dx];
        }
    };
    for (int dy = 0; dy < new_rows; ++dy) {
        float fy = (float)((dy + 0.5) * sy - 0.5);
        int iy = (int)std::floor(fy);
        fy -= iy;
        if (iy < 0) { iy = 0; fy = 0.f; }
        if (iy >= height - 1) { iy = height - 1; fy = 0.f; }
        const int b0 = (int)std::lrint((1.f - fy) * 2048.f), b1 = (int)std::lrint(fy * 2048.f);
        hpass(iy, row0);
        hpass(std::min(iy + 1, height - 1), row1);
        unsigned char* d = &dst[(size_t)dy * new_cols * 3];
        for (int k = 0; k < new_cols * 3; ++k) {
            const int v = (((b0 * (row0[k] >> 4)) >> 16) + ((b1 * (row1[k] >> 4)) >> 16) + 2) >> 2;
            d[k] = (unsigned char)std::min(255, std::max(0, v));
        }
    }
    return dst;
}

// ---------------------------------------------------------------------------------------------- class members
void PatchMatchCUDA::PatchMatchInit(std::vector<Scene>& Scenes, int ID) {
    images_.clear(); depths_.clear(); cameras_.clear();
    const std::vector<int>& srcID = Scenes[ID].srcID;
    for (int id : srcID)
        if (id < 0 || id >= (int)Scenes.size()) throw std::runtime_error("pair.txt names image " + std::to_string(id) + ", which has no entry of its own");
    ref_id_ = srcID[0];
    const std::string image_folder = input_folder_ + "/images", cam_folder = input_folder_ + "/cams";
    for (size_t i = 0; i < srcID.size(); ++i) {
        Scene& sc = Scenes[srcID[i]];
        Camera cam = ReadCamera(cam_folder + "/" + id8(srcID[i]) + "_cam.txt");
        if (sc.image.empty()) {                     // decoded once per process (the reference re-decodes for every call)
            if (!readGrayImage(image_folder, srcID[i], sc.image)) throw std::runtime_error("Can not read this image ! " + id8(srcID[i]));
            sc.orig_width = sc.image.width; sc.orig_height = sc.image.height;
            const int m = sc.max_image_size;        // resize rule, PatchMatch.cpp:893-925
            if (sc.image.width > m || sc.image.height > m) {
                const float factor = std::min((float)m / sc.image.width, (float)m / sc.image.height);
                sc.image = resizeLinear(sc.image, (int)std::round(sc.image.width * factor), (int)std::round(sc.image.height * factor));
            }
        }
        if (sc.image.width != sc.orig_width || sc.image.height != sc.orig_height) {   // PatchMatch.cpp:905-923
            const float scale_x = sc.image.width / (float)sc.orig_width, scale_y = sc.image.height / (float)sc.orig_height;
            cam.K[0] *= scale_x; cam.K[2] *= scale_x;
            cam.K[4] *= scale_y; cam.K[5] *= scale_y;
        }
        cam.width = sc.image.width; cam.height = sc.image.height;
        images_.push_back(&sc.image);
        cameras_.push_back(cam);
    }
    depth_min_ = cameras_[0].depth_min * 0.6f;      // PatchMatch.cpp:929-930
    depth_max_ = cameras_[0].depth_max * 1.2f;
    if (geom_) {                                    // PatchMatch.cpp:933-950: the sources' depth maps of the previous pass
        for (size_t i = 1; i < srcID.size(); ++i) {
            const size_t wh = (size_t)cameras_[i].width * cameras_[i].height;
            const Scene& src = Scenes[srcID[i]];
            if (src.depth && src.depth->size() == wh) { depths_.push_back(src.depth); continue; }   // this process wrote that file itself
            int h, w, nb;
            auto d = std::make_shared<std::vector<float>>();
            if (!readDmb(input_folder_ + "/MPMVS/2333_" + id8(srcID[i]) + "/depths.dmb", h, w, nb, *d)) d->assign(wh, 0.f);
            depths_.push_back(std::move(d));
        }
    }
}

void PatchMatchCUDA::CudaMemInit(Scene& scene) {
    bool all_u8 = tex_format_ == MPMVS_TEX_U8;
    for (const GrayImage* im : images_) all_u8 = all_u8 && im->u8.size() == im->px.size();
    if (all_u8) {                                   // un-resized 8-bit images: a quarter of the upload, no conversion
        std::vector<const uint8_t*> img(images_.size());
        for (size_t i = 0; i < images_.size(); ++i) img[i] = images_[i]->u8.data();
        check(mpmvs_set_views_u8(h_, (int)images_.size(), img.data(), cameras_.data()), "mpmvs_set_views_u8");
    } else {
        std::vector<const float*> img(images_.size());
        for (size_t i = 0; i < images_.size(); ++i) img[i] = images_[i]->px.data();
        check(mpmvs_set_views(h_, (int)images_.size(), img.data(), cameras_.data()), "mpmvs_set_views");
    }
    const size_t wh = (size_t)cameras_[0].width * cameras_[0].height;
    planes_.resize(wh); costs_.resize(wh); geom_costs_.assign(wh, 0.f);
    if (geom_) {
        std::vector<const float*> dep(depths_.size());
        for (size_t i = 0; i < depths_.size(); ++i) dep[i] = depths_[i]->data();
        check(mpmvs_set_src_depths(h_, dep.data()), "mpmvs_set_src_depths");
        // own result of the previous pass (PatchMatch.cpp:1051-1087)
        const std::string folder = input_folder_ + "/MPMVS/2333_" + id8(ref_id_);
        int h, w, nb;
        std::vector<float> fd, fn, fc;
        const float *d, *n, *c;
        if (scene.depth && scene.normal && scene.cost && scene.depth->size() == wh && scene.normal->size() == 3 * wh && scene.cost->size() == wh) {
            d = scene.depth->data(); n = scene.normal->data(); c = scene.cost->data();      // written by this process in the previous pass
        } else if (!readDmb(folder + "/depths.dmb", h, w, nb, fd) || !readDmb(folder + "/normals.dmb", h, w, nb, fn) ||
                   !readDmb(folder + "/costs.dmb", h, w, nb, fc) || fd.size() != wh) {
            throw std::runtime_error("geometric consistency pass needs the previous results in " + folder);
        } else {
            d = fd.data(); n = fn.data(); c = fc.data();
        }
        for (size_t i = 0; i < wh; ++i) planes_[i] = {n[3 * i], n[3 * i + 1], n[3 * i + 2], d[i]};
        check(mpmvs_set_state(h_, (const float*)planes_.data(), c), "mpmvs_set_state");
    }
}

void PatchMatchCUDA::CudaPlanarPriorInitialization(const std::vector<float4>& PlaneParams, const std::vector<float>& masks) {
    const size_t wh = (size_t)cameras_[0].width * cameras_[0].height;
    std::vector<float4> prior(wh, float4{0, 0, 0, 0});
    std::vector<uint32_t> mask(wh, 0u);
    for (size_t i = 0; i < wh; ++i) {
        mask[i] = (uint32_t)masks[i];
        if (masks[i] > 0) prior[i] = PlaneParams[(size_t)masks[i] - 1];
    }
    check(mpmvs_set_prior(h_, (const float*)prior.data(), mask.data()), "mpmvs_set_prior");
}

void PatchMatchCUDA::Run(uint64_t seed) {
    check(mpmvs_run_into(h_, seed, (float*)planes_.data(), costs_.data(), geomPlanarPrior_ && geom_ ? geom_costs_.data() : nullptr), "mpmvs_run");
    if (geom_ && !geomPlanarPrior_) check(mpmvs_get_geom_costs(h_, geom_costs_.data()), "mpmvs_get_geom_costs");
}

float4 PatchMatchCUDA::GetPriorPlaneParams(const Triangle& t, int width) const {
    // The reference solves the 3x4 homogeneous system [X 1] n4 = 0 with cv::SVD::solveZ: its null vector is the plane
    // through the three back-projected vertices.
    const Camera& c = cameras_[0];
    const Point p[3] = {t.pt1, t.pt2, t.pt3};
    double X[3][3];
    for (int k = 0; k < 3; ++k) {
        const float depth = planes_[(size_t)p[k].y * width + p[k].x].w;
        X[k][0] = depth * (p[k].x - c.K[2]) / c.K[0];
        X[k][1] = depth * (p[k].y - c.K[5]) / c.K[4];
        X[k][2] = depth;
    }
    const double u[3] = {X[1][0] - X[0][0], X[1][1] - X[0][1], X[1][2] - X[0][2]};
    const double v[3] = {X[2][0] - X[0][0], X[2][1] - X[0][1], X[2][2] - X[0][2]};
    const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double d = -(n[0] * X[0][0] + n[1] * X[0][1] + n[2] * X[0][2]);
    if (d < 0) nn = -nn;
    return float4{(float)(n[0] / nn), (float)(n[1] / nn), (float)(n[2] / nn), (float)(d / nn)};
}

std::vector<Triangle> PatchMatchCUDA::DelaunayTriangulation(const Rect& rc, const std::vector<Point>& points) const {
    if (points.empty()) throw std::runtime_error("No Point to Triangulate!");
    std::vector<int> xy(2 * points.size());
    for (size_t i = 0; i < points.size(); ++i) { xy[2 * i] = points[i].x - rc.x; xy[2 * i + 1] = points[i].y - rc.y; }
    int n = 0;
    check(mpmvs_delaunay(xy.data(), (int)points.size(), rc.width, rc.height, nullptr, 0, &n), "mpmvs_delaunay");
    std::vector<int> tri((size_t)3 * n);
    check(mpmvs_delaunay(xy.data(), (int)points.size(), rc.width, rc.height, tri.data(), n, &n), "mpmvs_delaunay");
    std::vector<Triangle> out((size_t)n);
    for (int i = 0; i < n; ++i) out[i] = Triangle{points[tri[3 * i]], points[tri[3 * i + 1]], points[tri[3 * i + 2]]};
    return out;
}

void PatchMatchCUDA::GetTriangulateVertices(std::vector<Point>& Vertices) {
    const int cap = 3 * ((GetReferenceImageWidth() + 4) / 5) * ((GetReferenceImageHeight() + 4) / 5);
    std::vector<int> xy((size_t)2 * cap);
    int n = 0;
    check(mpmvs_pick_vertices(h_, geomPlanarPrior_ ? 1 : 0, xy.data(), cap, &n), "mpmvs_pick_vertices");
    Vertices.resize(n);
    for (int i = 0; i < n; ++i) Vertices[i] = Point{xy[2 * i], xy[2 * i + 1]};
}

// ---------------------------------------------------------------------------------------------- ProcessProblem
void ProcessProblem(const std::string& input_folder, const std::string& output_folder, std::vector<Scene>& Scenes, int ID,
                    bool geom_consistency, bool planar_prior, uint64_t seed, int tex_format) {
    Scene& scene = Scenes[ID];
    std::cout << "Processing image " << id8(scene.refID) << " ..." << std::endl;
    const std::string result_folder = output_folder + "/2333_" + id8(scene.refID);
    mkdir(result_folder.c_str(), 0777);

    HostPhaseTimes& T = PhaseTimes();
    double t0 = now_s();
    PatchMatchCUDA MP(0);
    MP.SetFolder(input_folder, output_folder);
    MP.SetTexFormat(tex_format);
    MP.SetGeomConsistencyParams(geom_consistency, planar_prior);
    MP.PatchMatchInit(Scenes, ID);
    T.init += now_s() - t0; t0 = now_s();
    MP.AllocatePatchMatch();
    MP.CudaMemInit(Scenes[ID]);
    T.upload += now_s() - t0; t0 = now_s();
    MP.Run(seed);
    T.run += now_s() - t0; t0 = now_s();
    if (planar_prior) {
        std::cout << "Run Planar Prior PatchMatch MVS ..." << std::endl;
        MP.SetPlanarPriorParams();
        MP.SetGeomConsistencyParams(false, true);
        mpmvs_prior_stats st{};
        MP.BuildPlanarPrior(&st);       // vertices, Delaunay, rasterisation, plane fit, range check, upload (cpp:536-600)
        MP.Run(seed ^ 0x5DEECE66DULL);
        MP.SetGeomConsistencyParams(geom_consistency, planar_prior);
        T.prior += now_s() - t0; t0 = now_s();
    }
    const int width = MP.GetReferenceImageWidth(), height = MP.GetReferenceImageHeight();
    auto depths = std::make_shared<std::vector<float>>((size_t)width * height);
    auto normals = std::make_shared<std::vector<float>>((size_t)width * height * 3);
    const float4* pl = MP.planes().data();
    float *dd = depths->data(), *nn = normals->data();
    for (size_t i = 0; i < depths->size(); ++i) {
        dd[i] = pl[i].w;
        nn[3 * i] = pl[i].x; nn[3 * i + 1] = pl[i].y; nn[3 * i + 2] = pl[i].z;
    }
    auto costs = std::make_shared<std::vector<float>>(MP.costs());
    T.collect += now_s() - t0; t0 = now_s();
    // the three result files (PatchMatch.cpp:620-633) are written in the background; later passes (and other images of
    // this pass: the reference's in-place, Gauss-Seidel order) read the buffers, not the files
    SubmitDmb(result_folder + "/depths.dmb", height, width, 1, depths);
    SubmitDmb(result_folder + "/normals.dmb", height, width, 3, normals);
    SubmitDmb(result_folder + "/costs.dmb", height, width, 1, costs);
    scene.depth = depths; scene.normal = normals; scene.cost = costs;
    T.write += now_s() - t0;
    std::cout << "Processing image " << id8(scene.refID) << " done!" << std::endl;
    MP.Release(Scenes, ID);
}

}  // namespace mpmvs
