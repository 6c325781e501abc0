// Small float cv::Mat algebra for the cv:: stand-in (included at the end of core.hpp): the lazy MatExpr forms the
// ORB-SLAM2 sources use on and around the hot path (ORBmatcher.cc:368-374, :396, :418-428, :794, :1304-1308, :1553-1558;
// Frame.cc:265-268, :285, :311-321, :599-619; KeyFrame.cc:75-86, :641; MapPoint.cc:55-58, :369-378), evaluated the way
// OpenCV evaluates them.  Written from OpenCV's documented expression rules (matop: transposes and scale factors fold
// into one cv::gemm call; A*B+C is one gemm) and its arithmetic: cv::gemm's small-matrix path (no transpose flags, inner
// dimension 2..4) forms the dot products in float and applies alpha / beta*C in double; every other gemm accumulates
// in double; cv::norm and Mat::dot of short float vectors accumulate in double; A/s is convertTo(alpha = 1/s), a float
// multiply.  gemm (both paths) and norm are pinned to cv2 4.13 golden vectors (tests/golden/cvsmall_golden.npz); the
// forms that have no Python binding (Mat::dot, A/s, s*A) follow from reading the expression rules and are only ever
// compared shell-vs-reference through this same header.  Anything outside this inventory aborts instead of guessing.
#ifndef ORBGPU_SHIM_MAT_ALGEBRA_HPP
#define ORBGPU_SHIM_MAT_ALGEBRA_HPP

#include <cstdio>
#include <cstdlib>

namespace cv {

enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

namespace shim_detail {
inline void unsupported(const char* what) {
    std::fprintf(stderr, "cv shim: %s is outside the pinned inventory of shim/opencv2/core/mat_algebra.hpp\n", what);
    std::abort();
}
inline const float& f(const Mat& m, int r, int c) { return m.at<float>(r, c); }
}  // namespace shim_detail

// D = alpha * op(A) * op(B) + beta * op(C), CV_32F.
inline void gemm(const Mat& A, const Mat& B, double alpha, const Mat& C, double beta, Mat& D, int flags = 0) {
    using shim_detail::f;
    if (A.type() != CV_32F || B.type() != CV_32F) shim_detail::unsupported("gemm on a non-float matrix");
    const bool at = (flags & GEMM_1_T) != 0, bt = (flags & GEMM_2_T) != 0, ct = (flags & GEMM_3_T) != 0;
    const int M = at ? A.cols : A.rows, len = at ? A.rows : A.cols, N = bt ? B.rows : B.cols;
    if ((bt ? B.cols : B.rows) != len) shim_detail::unsupported("gemm with mismatched inner dimensions");
    const bool haveC = !C.empty() && beta != 0;
    Mat out(M, N, CV_32F);
    if (flags == 0 && 2 <= len && len <= 4 && (len == N || len == M)) {
        // small-matrix path: float dot product left to right, then alpha / beta*C in double, rounded once
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                float t = f(A, i, 0) * f(B, 0, j);
                for (int k = 1; k < len; ++k) t = t + f(A, i, k) * f(B, k, j);
                out.at<float>(i, j) = (float)((double)t * alpha + (haveC ? (double)f(C, i, j) * beta : 0.0));
            }
    } else {
        if (len >= 4) shim_detail::unsupported("general gemm with an inner dimension >= 4 (unrolled partial sums)");
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                double s = 0.0;
                for (int k = 0; k < len; ++k) s += (double)(at ? f(A, k, i) : f(A, i, k)) * (double)(bt ? f(B, j, k) : f(B, k, j));
                s *= alpha;
                out.at<float>(i, j) = haveC ? (float)(s + (double)(ct ? f(C, j, i) : f(C, i, j)) * beta) : (float)s;
            }
    }
    D = out;
}

// cv::norm(m), NORM_L2: squares and sum in double
inline double norm(const Mat& m, int normType = NORM_L2) {
    if (m.type() != CV_32F || normType != NORM_L2) shim_detail::unsupported("norm other than L2 of a float matrix");
    double s = 0.0;
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c) { const double v = m.at<float>(r, c); s += v * v; }
    return std::sqrt(s);
}
// cv::norm(a, b, NORM_L1) of two float patches (Frame.cc:619: integer-valued entries, the sum is exact in any order)
inline double norm(const Mat& a, const Mat& b, int normType) {
    if (a.type() != CV_32F || b.type() != CV_32F || normType != NORM_L1 || a.rows != b.rows || a.cols != b.cols)
        shim_detail::unsupported("norm(a, b) other than L1 of two equal-size float matrices");
    double s = 0.0;
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) s += std::fabs((double)(a.at<float>(r, c) - b.at<float>(r, c)));
    return s;
}

// Lazy expression with OpenCV's folding rules.  kind: 0 plain matrix a; 1 alpha*a [+ beta*b] (AddEx);
// 2 alpha*a^T; 3 gemm(a, b, alpha, c, beta, flags).
class MatExpr {
public:
    int kind, flags;
    Mat a, b, c;
    double alpha, beta;
    MatExpr() : kind(0), flags(0), alpha(1), beta(0) {}
    MatExpr(const Mat& m) : kind(0), flags(0), a(m), alpha(1), beta(0) {}
    static MatExpr addex(const Mat& a, const Mat& b, double alpha, double beta) { MatExpr e; e.kind = 1; e.a = a; e.b = b; e.alpha = alpha; e.beta = beta; return e; }
    static MatExpr transposed(const Mat& a, double alpha) { MatExpr e; e.kind = 2; e.a = a; e.alpha = alpha; return e; }
    static MatExpr product(int flags, const Mat& a, const Mat& b, double alpha, const Mat& c = Mat(), double beta = 0) {
        MatExpr e; e.kind = 3; e.flags = flags; e.a = a; e.b = b; e.c = c; e.alpha = alpha; e.beta = beta; return e;
    }
    bool isScaled() const { return kind == 1 && (b.empty() || beta == 0); }

    Mat eval() const {
        switch (kind) {
            case 0: return a;
            case 1: {
                if (b.empty() || beta == 0) return scaled(a, alpha);
                if (a.rows != b.rows || a.cols != b.cols || a.type() != CV_32F || b.type() != CV_32F) shim_detail::unsupported("a +- b of unequal matrices");
                Mat m(a.rows, a.cols, CV_32F);
                for (int r = 0; r < a.rows; ++r)
                    for (int col = 0; col < a.cols; ++col) {
                        const float x = a.at<float>(r, col), y = b.at<float>(r, col);
                        float v;
                        if (alpha == 1 && beta == 1) v = x + y;
                        else if (alpha == 1 && beta == -1) v = x - y;
                        else if (alpha == 1) v = y * (float)beta + x;        // cv::scaleAdd(b, beta, a)
                        else if (beta == 1) v = alpha == -1 ? y - x : x * (float)alpha + y;
                        else { shim_detail::unsupported("addWeighted"); v = 0; }
                        m.at<float>(r, col) = v;
                    }
                return m;
            }
            case 2: {
                Mat m(a.cols, a.rows, a.type());
                if (a.type() != CV_32F) shim_detail::unsupported("transpose of a non-float matrix");
                for (int r = 0; r < a.rows; ++r)
                    for (int col = 0; col < a.cols; ++col) m.at<float>(col, r) = a.at<float>(r, col);
                return alpha == 1 ? m : scaled(m, alpha);
            }
            default: { Mat d; gemm(a, b, alpha, c, beta, d, flags); return d; }
        }
    }
    operator Mat() const { return eval(); }

    // members the sources call on an expression
    MatExpr t() const { return kind == 0 ? transposed(a, 1) : transposed(eval(), 1); }
    double dot(const Mat& m) const { return eval().dot(m); }
    Mat clone() const { return eval().clone(); }
    Mat col(int i) const { return eval().col(i); }
    Mat row(int i) const { return eval().row(i); }
    template <typename T> T& at(int i) { tmp_ = eval(); return tmp_.at<T>(i); }
    template <typename T> T& at(int r, int col) { tmp_ = eval(); return tmp_.at<T>(r, col); }

    // alpha * m as Mat::convertTo does it: a float multiply by (float)alpha
    static Mat scaled(const Mat& m, double alpha) {
        if (alpha == 1) return m;
        if (m.type() != CV_32F) shim_detail::unsupported("scaling a non-float matrix");
        Mat o(m.rows, m.cols, CV_32F);
        const float s = (float)alpha;
        for (int r = 0; r < m.rows; ++r)
            for (int col = 0; col < m.cols; ++col) o.at<float>(r, col) = m.at<float>(r, col) * s;
        return o;
    }
private:
    Mat tmp_;
};

inline MatExpr Mat::t() const { return MatExpr::transposed(*this, 1); }
inline Mat::Mat(const MatExpr& e) : rows(0), cols(0), data(0), type_(CV_8UC1) { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e) { *this = e.eval(); return *this; }

namespace shim_detail {
// e1 * e2: transposes and scale factors fold into the gemm call
inline MatExpr matmul(const MatExpr& e1, const MatExpr& e2) {
    int flags = 0;
    double scale = 1;
    Mat m1, m2;
    if (e1.kind == 2) { flags |= GEMM_1_T; scale = e1.alpha; m1 = e1.a; }
    else if (e1.isScaled()) { scale = e1.alpha; m1 = e1.a; }
    else m1 = e1.eval();
    if (e2.kind == 2) { flags |= GEMM_2_T; scale *= e2.alpha; m2 = e2.a; }
    else if (e2.isScaled()) { scale *= e2.alpha; m2 = e2.a; }
    else m2 = e2.eval();
    return MatExpr::product(flags, m1, m2, scale);
}
inline MatExpr mulscalar(const MatExpr& e, double s) {
    MatExpr r = e;
    switch (e.kind) {
        case 0: return MatExpr::addex(e.a, Mat(), s, 0);
        case 1: r.alpha *= s; r.beta *= s; return r;
        case 2: r.alpha *= s; return r;
        default: r.alpha *= s; r.beta *= s; return r;
    }
}
// e1 + sign*e2
inline MatExpr addsub(const MatExpr& e1, const MatExpr& e2, double sign) {
    if (e1.kind == 3 && e1.c.empty() && (e2.kind == 0 || e2.isScaled() || e2.kind == 2)) {   // A*B + C: one gemm
        const int fl = e1.flags | (e2.kind == 2 ? GEMM_3_T : 0);
        return MatExpr::product(fl, e1.a, e1.b, e1.alpha, e2.a, sign * (e2.kind == 0 ? 1.0 : e2.alpha));
    }
    if (sign > 0 && e2.kind == 3 && e2.c.empty() && (e1.kind == 0 || e1.isScaled() || e1.kind == 2)) return addsub(e2, e1, 1);
    double alpha = 1, beta = sign;
    Mat m1, m2;
    if (e1.isScaled()) { m1 = e1.a; alpha = e1.alpha; } else m1 = e1.eval();
    if (e2.isScaled()) { m2 = e2.a; beta = sign * e2.alpha; } else m2 = e2.eval();
    return MatExpr::addex(m1, m2, alpha, beta);
}
}  // namespace shim_detail

inline MatExpr operator*(const Mat& a, const Mat& b) { return MatExpr::product(0, a, b, 1); }
inline MatExpr operator*(const MatExpr& a, const Mat& b) { return shim_detail::matmul(a, MatExpr(b)); }
inline MatExpr operator*(const Mat& a, const MatExpr& b) { return shim_detail::matmul(MatExpr(a), b); }
inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return shim_detail::matmul(a, b); }
inline MatExpr operator*(double s, const Mat& a) { return MatExpr::addex(a, Mat(), s, 0); }
inline MatExpr operator*(const Mat& a, double s) { return MatExpr::addex(a, Mat(), s, 0); }
inline MatExpr operator*(double s, const MatExpr& e) { return shim_detail::mulscalar(e, s); }
inline MatExpr operator*(const MatExpr& e, double s) { return shim_detail::mulscalar(e, s); }
inline MatExpr operator/(const Mat& a, double s) { return MatExpr::addex(a, Mat(), 1.0 / s, 0); }
inline MatExpr operator/(const MatExpr& e, double s) { return shim_detail::mulscalar(e, 1.0 / s); }
inline MatExpr operator-(const Mat& a) { return MatExpr::addex(a, Mat(), -1, 0); }
inline MatExpr operator-(const MatExpr& e) { return shim_detail::mulscalar(e, -1); }
inline MatExpr operator+(const Mat& a, const Mat& b) { return MatExpr::addex(a, b, 1, 1); }
inline MatExpr operator+(const MatExpr& a, const Mat& b) { return shim_detail::addsub(a, MatExpr(b), 1); }
inline MatExpr operator+(const Mat& a, const MatExpr& b) { return shim_detail::addsub(MatExpr(a), b, 1); }
inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return shim_detail::addsub(a, b, 1); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return MatExpr::addex(a, b, 1, -1); }
inline MatExpr operator-(const MatExpr& a, const Mat& b) { return shim_detail::addsub(a, MatExpr(b), -1); }
inline MatExpr operator-(const Mat& a, const MatExpr& b) { return shim_detail::addsub(MatExpr(a), b, -1); }
inline MatExpr operator-(const MatExpr& a, const MatExpr& b) { return shim_detail::addsub(a, b, -1); }

inline MatExpr Mat::eye(int r, int c, int type) {
    if (type != CV_32F) shim_detail::unsupported("Mat::eye of a non-float type");
    Mat m(r, c, type);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = i == j ? 1.f : 0.f;
    return MatExpr(m);
}
inline MatExpr Mat::ones(int r, int c, int type) {
    if (type != CV_32F) shim_detail::unsupported("Mat::ones of a non-float type");
    Mat m(r, c, type);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = 1.f;
    return MatExpr(m);
}

// cv::Mat_<float>(r, c) << a, b, c  (KeyFrame.cc:84, :638; Frame.cc:710)
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(Mat_<T>* m) : m_(m), i_(0) {}
    template <typename U> MatCommaInitializer_& operator,(U v) { put((T)v); return *this; }
    void put(T v);
    operator Mat() const;
    operator MatExpr() const { return MatExpr((Mat)*this); }
private:
    Mat_<T>* m_;
    int i_;
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_8U) {}
};
template <typename T> inline void MatCommaInitializer_<T>::put(T v) {
    ((T*)m_->data)[i_++] = v;
}
template <typename T> inline MatCommaInitializer_<T>::operator Mat() const { return *m_; }
template <typename T, typename U> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) {
    MatCommaInitializer_<T> ci(const_cast<Mat_<T>*>(&m));
    ci.put((T)v);
    return ci;
}
// x3Dc-style temporaries are used in products right away (KeyFrame.cc:641)
template <typename T> inline MatExpr operator*(const Mat& a, const MatCommaInitializer_<T>& b) { return a * (Mat)b; }

// declared so that Frame.cc's distorted-camera branch compiles (Frame.cc:456, :483); the rectified / undistorted
// configurations of the hot path never reach it
inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat& = Mat(), const Mat& = Mat()) {
    shim_detail::unsupported("undistortPoints (distorted camera model)");
}

}  // namespace cv
#endif
