// Minimal stand-in for the subset of OpenCV's core types that the ORB front-end hot path
// touches (ORBextractor.h:26, ORBextractor.cc:57-60, ORBmatcher.h).  There is no OpenCV C++ in
// the build image, so the oracle build (oracle/Makefile) and the drop-in host shells
// (orb_slam2_with_comment_b200/csrc/host) compile against this directory instead; a real deployment puts the
// genuine OpenCV include directory first on the include path and this file is never seen.
//
// Written from the public OpenCV API documentation; only the members used on the hot path exist.
#ifndef ORBGPU_SHIM_OPENCV2_CORE_HPP
#define ORBGPU_SHIM_OPENCV2_CORE_HPP

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

static inline int cvRound(double v) { return (int)std::lrint(v); }
static inline int cvRound(float v) { return (int)std::lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(T s) { x *= s; y *= s; return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct KeyPoint {
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

class MatExpr;
class _OutputArray;

// Byte-addressed 2-D array header with shared ownership of the pixel buffer (like cv::Mat,
// one channel only: CV_8UC1 or CV_32FC1).
class Mat {
public:
    struct Step {
        size_t v;
        Step() : v(0) {}
        operator size_t() const { return v; }
        Step& operator=(size_t s) { v = s; return *this; }
    };

    int rows, cols;
    uchar* data;
    Step step;

    Mat() : rows(0), cols(0), data(0), type_(CV_8UC1) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(0), type_(type) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), data(0), type_(type) { create(sz.height, sz.width, type); }
    // header over caller-owned memory (no copy, no ownership)
    Mat(int r, int c, int type, void* ext, size_t stepBytes = 0) : rows(r), cols(c), data((uchar*)ext), type_(type) {
        step = stepBytes ? stepBytes : (size_t)c * elemSize();
    }

    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && type_ == type) return;   // like cv::Mat::create: same geometry => keep (also a view)
        type_ = type;
        rows = r;
        cols = c;
        step = (size_t)c * elemSize();
        size_t n = (size_t)r * step.v;
        buf_.reset(new uchar[n ? n : 1], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = 0; rows = cols = 0; step = 0; }

    // Like cv::MatExpr for Mat::zeros: assigning it to an existing Mat of the same geometry zero-fills
    // that Mat IN PLACE (no reallocation) — computeDescriptors (ORBextractor.cc:1037) relies on this to
    // write into a rowRange of the caller's descriptor matrix.
    struct ZerosExpr { int r, c, type; };
    static ZerosExpr zeros(int r, int c, int type) { ZerosExpr e = {r, c, type}; return e; }
    Mat(const ZerosExpr& e) : rows(0), cols(0), data(0), type_(e.type) { *this = e; }
    Mat& operator=(const ZerosExpr& e) {
        if (!(data && rows == e.r && cols == e.c && type_ == e.type)) create(e.r, e.c, e.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step.v, 0, (size_t)cols * elemSize());
        return *this;
    }

    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32FC1 ? 4 : 1; }
    size_t step1() const { return step.v / elemSize(); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    bool isContinuous() const { return step.v == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }

    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step.v + (size_t)r.x * elemSize();
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }

    // small float algebra (mat_algebra.hpp)
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);
    MatExpr t() const;
    static MatExpr eye(int r, int c, int type);
    static MatExpr ones(int r, int c, int type);
    // Mat::dot of two float vectors / matrices of equal size: products and sum in double
    double dot(const Mat& m) const {
        double s = 0.0;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) s += (double)at<float>(r, c) * (double)m.at<float>(r, c);
        return s;
    }
    void copyTo(const _OutputArray& dst) const;
    // convertTo without scaling: 8U -> 32F (Frame.cc:599, :616) or same type
    void convertTo(Mat& dst, int rtype) const {
        Mat out(rows, cols, rtype);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const float v = type_ == CV_32FC1 ? at<float>(r, c) : (float)at<uchar>(r, c);
                if (rtype == CV_32FC1) out.at<float>(r, c) = v; else out.at<uchar>(r, c) = (uchar)v;
            }
        dst = out;
    }
    // multi-channel views exist only on the distorted-camera branch (Frame.cc:454-457, :482-484), which ends in
    // cv::undistortPoints (not provided): reshape is the identity here and that call aborts
    Mat reshape(int) const { return *this; }

    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step.v, data + (size_t)r * step.v, (size_t)cols * elemSize());
        return m;
    }

    // single index: element i of a row or column vector (cv::Mat::at(int i0))
    template <typename T> T& at(int i) { return (rows == 1 || isContinuous()) ? ((T*)data)[i] : *(T*)(data + (size_t)i * step.v); }
    template <typename T> const T& at(int i) const { return (rows == 1 || isContinuous()) ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step.v); }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step.v + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.v + (ptrdiff_t)c * (ptrdiff_t)sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.v); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.v); }

private:
    int type_;
    std::shared_ptr<uchar> buf_;
};

// InputArray / OutputArray as used by ORBextractor::operator() (ORBextractor.h:59-61): the
// caller passes a Mat (or a temporary cv::Mat()); the callee uses empty()/getMat()/create()/release().
class _InputArray {
public:
    _InputArray() : m_(0) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) { m_ = &m; }
    _OutputArray(const Mat& m) { m_ = const_cast<Mat*>(&m); }   // a temporary view as destination (KeyFrame.cc:82-83)
    void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
    void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

float fastAtan2(float y, float x);

inline void Mat::copyTo(const _OutputArray& dst) const {
    dst.create(rows, cols, type_);
    Mat d = dst.getMat();
    for (int r = 0; r < rows; ++r) std::memcpy(d.data + (size_t)r * d.step.v, data + (size_t)r * step.v, (size_t)cols * elemSize());
}

typedef Point_<float> Point2f;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;

}  // namespace cv

#include "mat_algebra.hpp"

#endif
