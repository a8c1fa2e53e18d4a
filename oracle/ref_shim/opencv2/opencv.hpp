// Declaration-only stand-in for the OpenCV headers that /root/reference/include/PatchMatch.h
// and utility.h include (PatchMatch.h:15-20, utility.h:6-11). OpenCV's C++ SDK is not in this
// image; the reference's CUDA translation unit only needs these *types to exist* -- none of
// the kernels touch them. This file is ours (test infrastructure), not reference code.
#ifndef MPMVS_REF_SHIM_OPENCV_HPP
#define MPMVS_REF_SHIM_OPENCV_HPP
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

typedef unsigned char uchar;

namespace cv {
struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int _x, int _y) : x(_x), y(_y) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int _w, int _h) : x(_x), y(_y), width(_w), height(_h) {}
};
template <typename T, int N>
struct Vec {
    T val[N];
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<uchar, 3> Vec3b;
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
enum { INTER_LINEAR = 1 };
#ifndef CV_32F
#define CV_32F 5
#endif
// A non-owning view of a float image; enough for the harness to hand raw rows to cudaMemcpy2DToArray.
struct Mat {
    int rows, cols;
    void* data;
    size_t step[2];
    Mat() : rows(0), cols(0), data(nullptr) { step[0] = step[1] = 0; }
    Mat(Size s, int) : rows(s.height), cols(s.width), data(nullptr) { step[0] = step[1] = 0; }
    Mat clone() const { return *this; }
    bool empty() const { return data == nullptr; }
    template <typename T> T* ptr(int r = 0) { return (T*)((char*)data + r * step[0]); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)((const char*)data + r * step[0]); }
};
template <typename T>
struct Mat_ : public Mat {};
// host helper of SkySegment/src/SkyRegionDetect.cu:37-66 only needs this to exist; the harness never calls that helper
inline void resize(const Mat&, Mat&, Size, double = 0, double = 0, int = INTER_LINEAR) {}
}  // namespace cv
#endif
