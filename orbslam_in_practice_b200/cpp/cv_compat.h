// cv_compat.h -- the few OpenCV types that appear in the reference's ORBextractor / ORBmatcher
// signatures (include/ORBextractor.h:43-45,71; include/ORBmatcher.h:19), for builds where the OpenCV
// C++ headers are not installed (they are absent in this image).  With OpenCV present the real
// headers are used instead and this file adds nothing.
#pragma once

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(ORBX_FORCE_CV_COMPAT)
#include <opencv2/core.hpp>
#define ORBX_HAVE_OPENCV 1
#endif
#endif

#ifndef ORBX_HAVE_OPENCV
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#endif

namespace cv {

struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };

// same 28-byte layout as cv::KeyPoint (and as orbx_keypoint)
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

// 8-bit single-channel matrix header over a shared buffer
class Mat {
public:
    int rows, cols; size_t step; uchar *data;
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr) { create(r, c, type); }
    Mat(int r, int c, int, void *ext, size_t stp = 0) : rows(r), cols(c), step(stp ? stp : (size_t)c), data((uchar *)ext) {}
    void create(int r, int c, int)
    {
        if (data && r == rows && c == cols) return;
        rows = r; cols = c; step = (size_t)c;
        hold_.reset((uchar *)std::malloc((size_t)r * c + 1), std::free);
        data = hold_.get();
    }
    void release() { hold_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    bool isContinuous() const { return step == (size_t)cols; }
    Mat operator()(const Rect &r) const { Mat m(*this); m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat row(int r) const { Mat m(*this); m.data = data + (size_t)r * step; m.rows = 1; return m; }
    Mat clone() const { Mat m(rows, cols, CV_8UC1); for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols); return m; }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
private:
    std::shared_ptr<uchar> hold_;
};

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
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    void release() const { m_->release(); }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

} // namespace cv
#endif // !ORBX_HAVE_OPENCV
