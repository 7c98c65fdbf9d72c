/*
 * cvshim.h -- a minimal stand-in for the OpenCV C++ API used by the reference's
 * src/ORBextractor.cpp, so that file can be compiled UNMODIFIED from /root/reference (OpenCV C++
 * headers/libs do not exist in this image).  TEST INFRASTRUCTURE ONLY (oracle/_ref).
 *
 * Containers are original minimal code; the arithmetic primitives (resize, GaussianBlur, FAST,
 * fastAtan2, copyMakeBorder) forward to the C oracle's restatements, which are pinned bit-exact
 * against real OpenCV 4.13.0 by oracle/pin_cv2.py.
 */
#ifndef ORB_CVSHIM_H
#define ORB_CVSHIM_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "orb_oracle.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

static inline int cvRound(double v) { return orbo_cv_round(v); }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_ &operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct ZerosExpr { int rows, cols, type; };

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar *data;
    Mat() : rows(0), cols(0), step(0), data(0) {}
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(0) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(0) { create(r, c, type); }
    void create(int r, int c, int /*type*/)
    {
        if (data && r == rows && c == cols) return;
        rows = r; cols = c; step = (size_t)c;
        buf_.reset((uchar *)std::malloc((size_t)r * (size_t)c + 1), std::free);
        data = buf_.get();
    }
    void release() { buf_.reset(); data = 0; rows = cols = 0; step = 0; }
    static ZerosExpr zeros(int r, int c, int type) { ZerosExpr z = { r, c, type }; return z; }
    Mat &operator=(const ZerosExpr &z)
    { /* cv::MatExpr assignment writes into an existing header of the same size */
        create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols);
        return *this;
    }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    Mat operator()(const Rect &r) const { Mat m(*this); m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
    Mat clone() const
    {
        Mat m; m.create(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    Mat row(int r) const { Mat m(*this); m.data = data + (size_t)r * step; m.rows = 1; return m; }
private:
    std::shared_ptr<uchar> buf_;
};

/* InputArray / OutputArray as thin Mat references (the reference only calls these members) */
class _InputArray {
public:
    _InputArray(const Mat &m) : m_(const_cast<Mat *>(&m)) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }
protected:
    Mat *m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat &m) : _InputArray(m) {}
    void release() const { m_->release(); }
    void create(int r, int c, int t) const { m_->create(r, c, t); }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

static inline float fastAtan2(float y, float x) { return orbo_fast_atan2(y, x); }

static inline void resize(const Mat &src, Mat &dst, Size sz, double, double, int)
{
    if (dst.empty() || dst.cols != sz.width || dst.rows != sz.height) dst.create(sz.height, sz.width, CV_8UC1);
    orbo_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, sz.width, sz.height, dst.step);
}
static inline void copyMakeBorder(const Mat &src, Mat &dst, int t, int b, int l, int r, int)
{
    assert(t == b && l == r && t == l);
    if (dst.empty() || dst.cols != src.cols + 2 * l || dst.rows != src.rows + 2 * t) dst.create(src.rows + 2 * t, src.cols + 2 * l, CV_8UC1);
    /* src may alias the interior of dst (ORBextractor.cpp:1086): read through a copy */
    Mat s = src.clone();
    orbo_reflect101_border(s.data, s.cols, s.rows, s.step, dst.data, t, dst.step);
}
static inline void GaussianBlur(const Mat &src, Mat &dst, Size k, double sx, double sy, int)
{
    assert(k.width == 7 && k.height == 7 && sx == 2 && sy == 2);
    Mat s = src.clone();
    if (dst.empty() || dst.cols != src.cols || dst.rows != src.rows) dst.create(src.rows, src.cols, CV_8UC1);
    orbo_gaussian7_s2_u8(s.data, s.cols, s.rows, s.step, dst.data, dst.step);
}
static inline void FAST(const Mat &img, std::vector<KeyPoint> &kps, int threshold, bool nms)
{
    assert(nms);
    kps.clear();
    int cap = img.rows * img.cols / 4 + 16;
    std::vector<orbo_cand> c((size_t)cap);
    int n = orbo_fast9_nms(img.data, img.cols, img.rows, img.step, threshold, c.data(), cap);
    for (int i = 0; i < n; ++i) kps.push_back(KeyPoint((float)c[i].x, (float)c[i].y, 7.f, -1, (float)c[i].score));
}
struct KeyPointsFilter { /* only reached from the dead ComputeKeyPointsOld path */
    static void retainBest(std::vector<KeyPoint> &k, int n)
    {
        if (n >= 0 && (int)k.size() > n) {
            std::stable_sort(k.begin(), k.end(), [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
            k.resize((size_t)n);
        }
    }
};

} // namespace cv
#endif
