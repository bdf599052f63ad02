// mc_types.h — the handful of Eigen/Boost types that appear in the reference's API signatures
// (Vector3d, Matrix3d, Vector3l, VectorXd, VectorXl, ArrayXXd, Matrix3Xd, Rng), restated without
// Eigen or Boost (neither exists in this build environment).  Only the members the hot-path API
// uses are provided.  Layouts follow Eigen: matrices and arrays are column-major.
#ifndef MCB_HOST_MC_TYPES_H
#define MCB_HOST_MC_TYPES_H

#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

// Contract violations: the reference uses BOOST_ASSERT_MSG (abort in debug builds, unchecked under
// NDEBUG).  The mirror always checks and throws, so callers get the same message without UB.
struct McAssertion : std::logic_error { using std::logic_error::logic_error; };
#define MC_ASSERT_MSG(cond, msg) do { if (!(cond)) throw McAssertion(msg); } while (0)

struct Vector3d {
    double v[3];
    Vector3d() : v{0., 0., 0.} {}
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    static Vector3d Zero() { return Vector3d(); }
    static Vector3d UnitZ() { return Vector3d(0., 0., 1.); }
    double  operator()(int i) const { return v[i]; }
    double& operator()(int i) { return v[i]; }
    double dot(const Vector3d& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    Vector3d cross(const Vector3d& o) const {
        return Vector3d(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    Vector3d normalized() const { double n = norm(); return Vector3d(v[0] / n, v[1] / n, v[2] / n); }
};
inline Vector3d operator+(const Vector3d& a, const Vector3d& b) { return Vector3d(a(0) + b(0), a(1) + b(1), a(2) + b(2)); }
inline Vector3d operator-(const Vector3d& a, const Vector3d& b) { return Vector3d(a(0) - b(0), a(1) - b(1), a(2) - b(2)); }
inline Vector3d operator-(const Vector3d& a) { return Vector3d(-a(0), -a(1), -a(2)); }
inline Vector3d operator*(const Vector3d& a, double s) { return Vector3d(a(0) * s, a(1) * s, a(2) * s); }
inline Vector3d operator*(double s, const Vector3d& a) { return Vector3d(s * a(0), s * a(1), s * a(2)); }
inline Vector3d operator/(const Vector3d& a, double s) { return Vector3d(a(0) / s, a(1) / s, a(2) / s); }

struct Vector3l {
    long v[3];
    Vector3l() : v{0, 0, 0} {}
    Vector3l(long a, long b, long c) : v{a, b, c} {}
    long  operator()(int i) const { return v[i]; }
    long& operator()(int i) { return v[i]; }
    long prod() const { return v[0] * v[1] * v[2]; }
};

struct Matrix3d {                       // column-major: m[r + 3*c]
    double m[9];
    Matrix3d() : m{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
    static Matrix3d Zero() { return Matrix3d(); }
    static Matrix3d Identity() { Matrix3d r; r(0, 0) = r(1, 1) = r(2, 2) = 1.; return r; }
    static Matrix3d Diagonal(double a, double b, double c) { Matrix3d r; r(0, 0) = a; r(1, 1) = b; r(2, 2) = c; return r; }
    static Matrix3d Diagonal(const Vector3d& d) { return Diagonal(d(0), d(1), d(2)); }
    static Matrix3d Columns(const Vector3d& a, const Vector3d& b, const Vector3d& c) {
        Matrix3d r; for (int i = 0; i < 3; ++i) { r(i, 0) = a(i); r(i, 1) = b(i); r(i, 2) = c(i); } return r;
    }
    double  operator()(int r, int c) const { return m[r + 3 * c]; }
    double& operator()(int r, int c) { return m[r + 3 * c]; }
    Vector3d col(int c) const { return Vector3d(m[3 * c], m[3 * c + 1], m[3 * c + 2]); }
    Matrix3d transpose() const { Matrix3d r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i); return r; }
    double determinant() const {
        const Matrix3d& A = *this;
        return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
               A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    }
    Matrix3d inverse() const;           // cofactors / determinant (Eigen's fixed-size 3x3 path)
};
inline Vector3d operator*(const Matrix3d& A, const Vector3d& x) {
    return Vector3d(A(0, 0) * x(0) + A(0, 1) * x(1) + A(0, 2) * x(2), A(1, 0) * x(0) + A(1, 1) * x(1) + A(1, 2) * x(2),
                    A(2, 0) * x(0) + A(2, 1) * x(1) + A(2, 2) * x(2));
}
inline Matrix3d operator*(const Matrix3d& A, const Matrix3d& B) {
    Matrix3d r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0.; for (int k = 0; k < 3; ++k) s += A(i, k) * B(k, j); r(i, j) = s; }
    return r;
}
inline Matrix3d Matrix3d::inverse() const {
    const Matrix3d& A = *this; Matrix3d c;
    c(0, 0) =  (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)); c(1, 0) = -(A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)); c(2, 0) =  (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    c(0, 1) = -(A(0, 1) * A(2, 2) - A(0, 2) * A(2, 1)); c(1, 1) =  (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)); c(2, 1) = -(A(0, 0) * A(2, 1) - A(0, 1) * A(2, 0));
    c(0, 2) =  (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)); c(1, 2) = -(A(0, 0) * A(1, 2) - A(0, 2) * A(1, 0)); c(2, 2) =  (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0));
    const double det = A(0, 0) * c(0, 0) + A(0, 1) * c(1, 0) + A(0, 2) * c(2, 0);
    const double inv = 1. / det;
    Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = c.m[i] * inv;
    return r;
}

typedef std::vector<double> VectorXd;
typedef std::vector<long>   VectorXl;

// 3 x N matrix of points (Domain::checkpoints, trajectories): column j = point j
struct Matrix3Xd {
    std::vector<Vector3d> c;
    Matrix3Xd() {}
    explicit Matrix3Xd(long n) : c((size_t)n) {}
    long cols() const { return (long)c.size(); }
    const Vector3d& col(long j) const { return c[(size_t)j]; }
    Vector3d& col(long j) { return c[(size_t)j]; }
};

// Eigen::ArrayXXd (rows, cols), column-major, coefficient-wise arithmetic
class ArrayXXd {
    long r_, c_; std::vector<double> d_;
public:
    ArrayXXd() : r_(0), c_(0) {}
    ArrayXXd(long rows, long cols) : r_(rows), c_(cols), d_((size_t)(rows * cols), 0.) {}
    static ArrayXXd Zero(long rows, long cols) { return ArrayXXd(rows, cols); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    long size() const { return r_ * c_; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    double  operator()(long i, long j) const { return d_[(size_t)(i + r_ * j)]; }
    double& operator()(long i, long j) { return d_[(size_t)(i + r_ * j)]; }
    ArrayXXd& operator+=(const ArrayXXd& o) { same(o); for (size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i]; return *this; }
    ArrayXXd operator-(const ArrayXXd& o) const { same(o); ArrayXXd r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] -= o.d_[i]; return r; }
    ArrayXXd operator*(const ArrayXXd& o) const { same(o); ArrayXXd r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] *= o.d_[i]; return r; }
    ArrayXXd operator/(double s) const { ArrayXXd r(*this); for (double& x : r.d_) x /= s; return r; }
    ArrayXXd sqrt() const { ArrayXXd r(*this); for (double& x : r.d_) x = std::sqrt(x); return r; }
    bool operator!=(const ArrayXXd& o) const { return r_ != o.r_ || c_ != o.c_ || d_ != o.d_; }
private:
    void same(const ArrayXXd& o) const { MC_ASSERT_MSG(r_ == o.r_ && c_ == o.c_, "ArrayXXd shape mismatch"); }
};

// random.h:22 — boost::random::mt19937 and std::mt19937 are the same generator (same parameters, same
// default seeding), so the signature `solve(Rng& gen, ...)` keeps its meaning.
typedef std::mt19937 Rng;

// identity of a table-bearing object (Material, Domain) for the device-side table cache: heap addresses are reused after
// delete, a counter is not
#include <atomic>
inline unsigned long mcNextUid() { static std::atomic<unsigned long> next(1); return next.fetch_add(1); }

#endif
