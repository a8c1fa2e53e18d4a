// oracle/ref_shim_host/opencv2/opencv.hpp -- TEST INFRASTRUCTURE, ours, not reference code and not OpenCV code.
//
// A small functional stand-in for the part of OpenCV's C++ API that the reference's HOST program uses
// (/root/reference/src/PatchMatch.cpp, utility.cpp, main.cpp), so that those files can be compiled WHERE THEY LIE into
// oracle/_ref/libmpmvs_ref_host.so (oracle/ref_host_harness.cu). OpenCV's C++ SDK is not in this image; its Python
// module (cv2, the same library) is. Therefore:
//   * containers (Mat, Mat_, Vec, Point, Rect, Size), convertTo, merge, FileStorage's "key: value" reading are
//     implemented here -- they hold data, they have no numerics of their own;
//   * everything that IS numerics or a codec -- cv::Subdiv2D, cv::SVD::solveZ, cv::resize, cv::imread -- is forwarded
//     through function pointers (ref_shim::callbacks) that the Python side fills with the real OpenCV (cv2.Subdiv2D,
//     cv2.SVDecomp, cv2.resize, cv2.imread). Without callbacks those calls abort with a message;
//   * debug output (imwrite, line, imshow, applyColorMap, calcHist, Mat arithmetic of the .jpg previews) are no-ops or
//     abort: the reference's results (the .dmb maps and the .ply) do not depend on them.
#ifndef MPMVS_REF_SHIM_HOST_OPENCV_HPP
#define MPMVS_REF_SHIM_HOST_OPENCV_HPP
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

namespace ref_shim {
// Filled by ref_host_set_callbacks() (oracle/ref_host_harness.cu) from Python (tests/ref_host.py) with the real OpenCV.
struct Callbacks {
    // cv2.imread(path, flags): returns 0 and a buffer (rows x cols x channels, 8 bit) that stays valid until the next call
    int (*imread)(const char* path, int flags, int* rows, int* cols, int* channels, const uchar** data);
    // cv2.resize(src, (new_cols, new_rows), interpolation=INTER_LINEAR) of a rows x cols image of `type` into dst
    int (*resize)(const void* src, int rows, int cols, int type, int new_rows, int new_cols, void* dst);
    // cv2.Subdiv2D(rect), insert(points in order), getTriangleList(): returns n triangles, 6 floats each, valid until the next call
    int (*subdiv)(int rx, int ry, int rw, int rh, const float* pts, int npts, const float** tris, int* ntris);
    // cv::SVD::solveZ(A (rows x cols, float32), z (cols))
    int (*solvez)(const float* A, int rows, int cols, float* z);
    // told about every cv::imwrite (the harness uses the triangulation.png of ProcessProblem as a marker)
    void (*on_imwrite)(const char* path);
};
inline Callbacks& callbacks() {
    static Callbacks cb = {nullptr, nullptr, nullptr, nullptr, nullptr};
    return cb;
}
[[noreturn]] inline void unsupported(const char* what) {
    fprintf(stderr, "ref_shim_host: %s is not provided by the OpenCV stand-in (only the reference's result path is)\n", what);
    abort();
}
}  // namespace ref_shim

namespace cv {

inline int cvRoundImpl(double v) { return (int)lrint(v); }

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int _w, int _h) : x(_x), y(_y), width(_w), height(_h) {}
    bool contains(const Point& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "three-component constructor"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    // OpenCV's header-inline `operator /= (Vec<_Tp, cn>&, float alpha)` (modules/core/include/opencv2/core/matx.hpp) multiplies
    // by the float reciprocal, it does not divide: `float ialpha = 1.f/alpha; a[i] = saturate_cast<_Tp>(a[i]*ialpha);`.
    // RunFusion's averaged normal (src/PatchMatch.cpp:460,483) goes through it, so the last bit depends on this.
    Vec& operator/=(float s) { const float is = 1.f / s; for (int i = 0; i < N; ++i) val[i] = (T)(val[i] * is); return *this; }
};
template <typename T, int N>
inline Vec<T, N> operator+(const Vec<T, N>& a, const Vec<T, N>& b) {
    Vec<T, N> r;
    for (int i = 0; i < N; ++i) r.val[i] = (T)(a.val[i] + b.val[i]);
    return r;
}
typedef Vec<float, 3> Vec3f;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 6> Vec6f;

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<Vec3b> { enum { type = CV_8UC3 }; };
template <> struct DataType<Vec3f> { enum { type = CV_32FC3 }; };

enum { INTER_LINEAR = 1 };
enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { COLORMAP_JET = 2 };
enum { WINDOW_AUTOSIZE = 1 };

// Reference-counted dense 2-D array; copies share the pixels, as cv::Mat's do (PatchMatchInit and Release take the
// scene list BY VALUE and rely on that, src/PatchMatch.cpp:863,1091).
struct Mat {
    int rows, cols;
    uchar* data;
    size_t step[2];
    Mat() : rows(0), cols(0), data(nullptr), type_(0) { step[0] = step[1] = 0; }
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        step[1] = elemSize();
        step[0] = step[1] * (size_t)c;
        buf_ = std::make_shared<std::vector<uchar>>(step[0] * (size_t)r);
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }     // value-initialised storage
    static Mat zeros(Size s, int type) { return Mat(s, type); }
    static Mat ones(int r, int c, int type) {
        Mat m(r, c, type);
        const int cn = m.channels();
        for (size_t i = 0; i < (size_t)r * c; ++i) {                           // cv::Mat::ones sets the first channel
            if (m.depth() == CV_8U) m.data[i * cn] = 1;
            else ((float*)m.data)[i * cn] = 1.f;
        }
        return m;
    }
    static Mat ones(Size s, int type) { return ones(s.height, s.width, type); }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : 4); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step[0] = step[1] = 0; }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        memcpy(m.data, data, step[0] * (size_t)rows);
        return m;
    }
    template <typename T> T& at(int r, int c) { return *(T*)(data + step[0] * (size_t)r + sizeof(T) * (size_t)c); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + step[0] * (size_t)r + sizeof(T) * (size_t)c); }
    template <typename T> T& at(int i) { return ((T*)data)[i]; }               // single-row / single-column arrays
    template <typename T> const T& at(int i) const { return ((const T*)data)[i]; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + step[0] * (size_t)r); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + step[0] * (size_t)r); }
    // dst = saturate_cast<rtype>(src * alpha + beta), channel count kept (cv::Mat::convertTo)
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
        const int ddepth = rtype & 7, cn = channels();
        Mat out;
        if (!empty()) {
            out.create(rows, cols, CV_MAKETYPE(ddepth, cn));
            const size_t n = (size_t)rows * cols * cn;
            for (size_t i = 0; i < n; ++i) {
                const double v = (depth() == CV_8U ? (double)data[i] : (double)((const float*)data)[i]) * alpha + beta;
                if (ddepth == CV_8U) {
                    const int q = cvRoundImpl(v);
                    out.data[i] = (uchar)(q < 0 ? 0 : q > 255 ? 255 : q);
                } else {
                    ((float*)out.data)[i] = (float)v;
                }
            }
        }
        dst = out;
    }
  protected:
    int type_;
    std::shared_ptr<std::vector<uchar>> buf_;
};

template <typename T>
struct Mat_ : public Mat {
    Mat_() : Mat() { type_ = DataType<T>::type; }
    Mat_(const Mat& m) : Mat() { assign(m); }
    Mat_& operator=(const Mat& m) { assign(m); return *this; }
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
    Mat_ clone() const { return Mat_(Mat::clone()); }
  private:
    void assign(const Mat& m) {                                                // shares the pixels when the type matches, converts otherwise
        if (m.empty() || m.type() == (int)DataType<T>::type) { Mat::operator=(m); type_ = DataType<T>::type; }
        else if (m.channels() == ((int)DataType<T>::type >> 3) + 1) { Mat c; m.convertTo(c, (int)DataType<T>::type & 7); Mat::operator=(c); }
        else ref_shim::unsupported("Mat_<T>(Mat) with a different channel count");
    }
};

// ---- what the reference's result path calls into OpenCV for: forwarded to the real library ---------------------------
inline Mat imread(const std::string& path, int flags = IMREAD_COLOR) {
    Mat m;
    if (!ref_shim::callbacks().imread) ref_shim::unsupported("cv::imread without callbacks");
    int rows = 0, cols = 0, cn = 0;
    const uchar* px = nullptr;
    if (ref_shim::callbacks().imread(path.c_str(), flags, &rows, &cols, &cn, &px) != 0 || !px) return m;   // empty, like cv::imread
    m.create(rows, cols, CV_MAKETYPE(CV_8U, cn));
    memcpy(m.data, px, (size_t)rows * cols * cn);
    return m;
}
inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR) {
    if (!ref_shim::callbacks().resize || interpolation != INTER_LINEAR) ref_shim::unsupported("cv::resize without callbacks");
    Mat out(dsize.height, dsize.width, src.type());
    if (ref_shim::callbacks().resize(src.data, src.rows, src.cols, src.type(), dsize.height, dsize.width, out.data) != 0)
        ref_shim::unsupported("cv::resize callback failed");
    dst = out;
}
template <typename T>
inline void resize(const Mat& src, Mat_<T>& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    Mat out;
    resize(src, out, dsize, fx, fy, interpolation);
    dst = out;
}
class Subdiv2D {
  public:
    explicit Subdiv2D(Rect r) : rect_(r) {}
    int insert(Point2f p) { pts_.push_back(p.x); pts_.push_back(p.y); return (int)pts_.size() / 2; }
    void getTriangleList(std::vector<Vec6f>& out) const {
        if (!ref_shim::callbacks().subdiv) ref_shim::unsupported("cv::Subdiv2D without callbacks");
        const float* t = nullptr;
        int n = 0;
        if (ref_shim::callbacks().subdiv(rect_.x, rect_.y, rect_.width, rect_.height, pts_.data(), (int)pts_.size() / 2, &t, &n) != 0)
            ref_shim::unsupported("cv::Subdiv2D callback failed");
        out.resize(n);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 6; ++k) out[i][k] = t[6 * i + k];
    }
  private:
    Rect rect_;
    std::vector<float> pts_;
};
struct SVD {
    static void solveZ(const Mat& A, Mat& z) {
        if (!ref_shim::callbacks().solvez || A.type() != CV_32FC1) ref_shim::unsupported("cv::SVD::solveZ without callbacks");
        if (z.rows != A.cols || z.cols != 1 || z.type() != CV_32FC1) z.create(A.cols, 1, CV_32FC1);
        if (ref_shim::callbacks().solvez((const float*)A.data, A.rows, A.cols, (float*)z.data) != 0) ref_shim::unsupported("cv::SVD::solveZ callback failed");
    }
};

// ---- containers / debug output --------------------------------------------------------------------------------------
inline void merge(const std::vector<Mat>& planes, Mat& dst) {
    const Mat& p0 = planes.at(0);
    const int cn = (int)planes.size();
    Mat out(p0.rows, p0.cols, CV_MAKETYPE(p0.depth(), cn));
    const size_t es = p0.depth() == CV_8U ? 1 : 4, n = (size_t)p0.rows * p0.cols;
    for (int k = 0; k < cn; ++k)
        for (size_t i = 0; i < n; ++i) memcpy(out.data + (i * cn + k) * es, planes[k].data + i * es, es);
    dst = out;
}
inline bool imwrite(const std::string& path, const Mat&) {                    // previews are not results: nothing is written
    if (ref_shim::callbacks().on_imwrite) ref_shim::callbacks().on_imwrite(path.c_str());
    return true;
}
inline void line(Mat&, Point, Point, const Scalar&, int = 1) {}
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1) {}
inline void namedWindow(const std::string&, int = 0) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return 0; }
inline void applyColorMap(const Mat&, Mat&, int) { ref_shim::unsupported("cv::applyColorMap (the .jpg previews)"); }
inline void calcHist(const Mat*, int, const int*, const Mat&, Mat&, int, const int*, const float**) { ref_shim::unsupported("cv::calcHist (the .jpg previews)"); }
inline void minMaxLoc(const Mat&, double*, double*, Point* = nullptr, Point* = nullptr) { ref_shim::unsupported("cv::minMaxLoc (the .jpg previews)"); }
inline Mat operator-(const Mat&, double) { ref_shim::unsupported("Mat arithmetic (the .jpg previews)"); }
inline Mat operator+(const Mat&, double) { ref_shim::unsupported("Mat arithmetic (the .jpg previews)"); }
inline Mat operator*(const Mat&, double) { ref_shim::unsupported("Mat arithmetic (the .jpg previews)"); }

// "key: value" lines of the reference's config.yaml (utility.cpp:8-35)
class FileNode {
  public:
    FileNode() : found_(false) {}
    explicit FileNode(const std::string& v) : found_(true), value_(v) {}
    void operator>>(std::string& s) const { if (found_) s = value_; }
    void operator>>(int& v) const { if (found_) v = (int)strtol(value_.c_str(), nullptr, 10); }
    void operator>>(bool& v) const { if (found_) v = strtol(value_.c_str(), nullptr, 10) != 0; }
    void operator>>(float& v) const { if (found_) v = strtof(value_.c_str(), nullptr); }
  private:
    bool found_;
    std::string value_;
};
class FileStorage {
  public:
    enum { READ = 0 };
    FileStorage(const std::string& path, int) {
        std::ifstream f(path);
        std::string ln;
        while (std::getline(f, ln)) {
            if (ln.empty() || ln[0] == '#' || ln[0] == '%' || ln.compare(0, 3, "---") == 0) continue;
            const size_t colon = ln.find(':');
            if (colon == std::string::npos) continue;
            std::string key = trim(ln.substr(0, colon)), val = trim(ln.substr(colon + 1));
            if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
            kv_[key] = val;
        }
    }
    FileNode operator[](const char* key) const {
        auto it = kv_.find(key);
        return it == kv_.end() ? FileNode() : FileNode(it->second);
    }
  private:
    static std::string trim(const std::string& s) {
        const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    std::map<std::string, std::string> kv_;
};

}  // namespace cv

inline int cvRound(double v) { return cv::cvRoundImpl(v); }
#endif
