// boundary.h — host mirror of the reference's Boundary family (boundary.h:34-282).
// These objects carry the geometry and the pairing; Boundary::scatter itself
// (boundary.cpp:283-287, 308-312, 349-359, 455-460, 516-522) runs inside the CUDA step kernel,
// which reads the flattened mcb_plane_desc each object produces through describe().
#ifndef MCB_HOST_BOUNDARY_H
#define MCB_HOST_BOUNDARY_H
#include <string>
#include <vector>
#include "../../include/mcb.h"
#include "constants.h"
#include "mc_types.h"

class Subdomain;

Matrix3d rotMatrix(const Vector3d& n);     // rotation taking +z onto n (boundary.cpp:33-37)

class Boundary {
public:
    class Shape;
    typedef std::vector<const Boundary*> Pointers;

private:
    const Subdomain* sdom_;
    Vector3d n_;
    double off_;

public:
    Boundary();
    Boundary(const Vector3d& o, const Vector3d& n);
    Boundary(const Vector3d& o, const Shape& s);
    virtual ~Boundary();

    virtual std::string type() const = 0;
    virtual int kind() const = 0;                        // MCB_BDRY_*
    virtual bool isInit() const;
    const Subdomain* sdom() const { return sdom_; }
    void sdom(const Subdomain* s) { sdom_ = s; }

    Vector3d normal() const { return n_; }
    double offset() const { return off_; }
    Vector3d projection(const Vector3d& pos) const { return pos - distance(pos) * n_; }
    double distance(const Vector3d& pos) const { return n_.dot(pos) + off_; }

    virtual Pointers partners() const { return Pointers(); }
    virtual void describe(mcb_plane_desc& d) const;       // fills everything but sdom / pair ids
};

class Boundary::Shape {
public:
    virtual ~Shape() {}
    virtual std::string type() const = 0;
    virtual int kind() const = 0;                         // MCB_SHAPE_*
    virtual bool isInit() const { return area() != 0.; }
    virtual Vector3d normal() const = 0;
    virtual double area() const = 0;
    virtual std::vector<Vector3d> verts() const = 0;
};

class Parallelogram : public Boundary::Shape {
    Vector3d i_, j_;
public:
    Parallelogram() {}
    Parallelogram(const Vector3d& i, const Vector3d& j) : i_(i), j_(j) {}
    std::string type() const { return "P"; }
    int kind() const { return MCB_SHAPE_PARALLELOGRAM; }
    Vector3d normal() const { return i_.cross(j_).normalized(); }
    double area() const { return i_.cross(j_).norm(); }
    std::vector<Vector3d> verts() const { return {i_, j_}; }
};

class Triangle : public Boundary::Shape {
    Vector3d i_, j_;
public:
    Triangle() {}
    Triangle(const Vector3d& i, const Vector3d& j) : i_(i), j_(j) {}
    std::string type() const { return "T"; }
    int kind() const { return MCB_SHAPE_TRIANGLE; }
    Vector3d normal() const { return i_.cross(j_).normalized(); }
    double area() const { return i_.cross(j_).norm() / 2.; }
    std::vector<Vector3d> verts() const { return {i_, j_}; }
};

// Polygon<N>: planar fan of N-1 edge vectors from a common corner, N-2 triangles (boundary.h:112-131, boundary.cpp:189-258)
template <int N>
class Polygon : public Boundary::Shape {
    static_assert(N > 3, "Polygon must have more than 3 sides");
    std::vector<Vector3d> verts_;
    std::vector<double> areas_;
public:
    Polygon() : verts_((size_t)(N - 1)), areas_((size_t)(N - 2), 0.) {}
    explicit Polygon(const std::vector<Vector3d>& v) : verts_(v), areas_((size_t)(N - 2)) {
        MC_ASSERT_MSG((int)verts_.size() == N - 1, "Incorrect number of vertices");
        for (int n = 0; n < N - 2; ++n) {
            const Vector3d c = verts_[(size_t)n].cross(verts_[(size_t)n + 1]);
            areas_[(size_t)n] = c.norm() / 2.;
            MC_ASSERT_MSG(areas_[(size_t)n] > Dbl::min(), "Area too small");
            const Vector3d a = c.normalized(), b = normal();
            for (int k = 0; k < 3; ++k) MC_ASSERT_MSG(std::abs(a(k) - b(k)) <= 1e-9, "Normals are inconsistent");
        }
    }
    std::string type() const { return std::to_string(N); }
    int kind() const { return MCB_SHAPE_POLYGON; }
    Vector3d normal() const { return verts_[0].cross(verts_[1]).normalized(); }
    double area() const { double s = 0.; for (double a : areas_) s += a; return s; }
    std::vector<Vector3d> verts() const { return verts_; }
};

//---------------------------------------- non-emitting boundaries
class SpecBoundary : public Boundary {
public:
    SpecBoundary() {}
    SpecBoundary(const Vector3d& o, const Vector3d& n) : Boundary(o, n) {}
    SpecBoundary(const Vector3d& o, const Shape& s, const double T) : Boundary(o, s) {
        MC_ASSERT_MSG(T == 0., "Cannot specify temperature");
    }
    std::string type() const { return "Spec"; }
    int kind() const { return MCB_BDRY_SPEC; }
};

class DiffBoundary : public Boundary {
public:
    DiffBoundary() {}
    DiffBoundary(const Vector3d& o, const Vector3d& n) : Boundary(o, n) {}
    DiffBoundary(const Vector3d& o, const Shape& s, const double T) : Boundary(o, s) {
        MC_ASSERT_MSG(T == 0., "Cannot specify temperature");
    }
    std::string type() const { return "Diff"; }
    int kind() const { return MCB_BDRY_DIFF; }
};

class InterBoundary : public Boundary {
    Pointers pairs_;
public:
    InterBoundary() {}
    InterBoundary(const Vector3d& o, const Vector3d& n) : Boundary(o, n) {}
    InterBoundary(const Vector3d& o, const Shape& s, const double T) : Boundary(o, s) {
        MC_ASSERT_MSG(T == 0., "Cannot specify temperature");
    }
    std::string type() const { return "Inter"; }
    int kind() const { return MCB_BDRY_INTER; }
    bool isInit() const { return Boundary::isInit() && !pairs_.empty(); }
    Pointers partners() const { return pairs_; }
    friend void makePair(InterBoundary& bdry1, InterBoundary& bdry2);
};
void makePair(InterBoundary& bdry1, InterBoundary& bdry2);      // boundary.cpp:361-369

//---------------------------------------- emitting boundaries
class Emitter {
public:
    typedef std::vector<const Emitter*> Pointers;
    virtual ~Emitter() {}
    virtual const Subdomain* emitSdom() const = 0;
    virtual const Boundary* emitBdry() const = 0;
    virtual double emitWeight() const = 0;
};

class EmitBoundary : public Boundary, public Emitter {
    Vector3d o_;
protected:
    double T_;
public:
    EmitBoundary() : T_(0.) {}
    EmitBoundary(const Vector3d& o, const Shape& s, const double T) : Boundary(o, s), o_(o), T_(T) {}
    const Subdomain* emitSdom() const { return sdom(); }
    const Boundary* emitBdry() const { return this; }
    double emitWeight() const { return shape().area() * std::abs(T_); }   // boundary.cpp:413-416
    double temperature() const { return T_; }
    void describe(mcb_plane_desc& d) const;
    virtual const Shape& shape() const = 0;
};

template <typename S>
class IsotBoundary : public EmitBoundary {
    S shape_;
public:
    IsotBoundary() {}
    IsotBoundary(const Vector3d& o, const S& s, const double T) : EmitBoundary(o, s, T), shape_(s) {}
    const Shape& shape() const { return shape_; }
    std::string type() const { return std::string("Isot") + shape_.type(); }
    int kind() const { return MCB_BDRY_ISOT; }
};

template <typename S> class PeriBoundary;
template <typename S>
void makePair(PeriBoundary<S>& bdry1, PeriBoundary<S>& bdry2, const Vector3d& t, const Matrix3d& r = Matrix3d::Identity());

template <typename S>
class PeriBoundary : public EmitBoundary {
    S shape_;
    const PeriBoundary* pair_;
    Matrix3d rot_;
    Vector3d transl_;
public:
    PeriBoundary() : pair_(0) {}
    PeriBoundary(const Vector3d& o, const S& s, const double T) : EmitBoundary(o, s, T), shape_(s), pair_(0) {}
    const Shape& shape() const { return shape_; }
    std::string type() const { return std::string("Peri") + shape_.type(); }
    int kind() const { return MCB_BDRY_PERI; }
    bool isInit() const { return Boundary::isInit() && pair_ != 0; }
    Boundary::Pointers partners() const { return pair_ ? Boundary::Pointers(1, pair_) : Boundary::Pointers(); }
    void describe(mcb_plane_desc& d) const {
        EmitBoundary::describe(d);
        for (int k = 0; k < 9; ++k) d.peri_rot[k] = rot_.m[k];
        for (int k = 0; k < 3; ++k) d.peri_transl[k] = transl_(k);
    }
    friend void makePair<>(PeriBoundary& bdry1, PeriBoundary& bdry2, const Vector3d& t, const Matrix3d& r);
};

// boundary.cpp:524-550: the partner maps back with R^T and -R^T t; wall temperatures become differences
template <typename S>
void makePair(PeriBoundary<S>& bdry1, PeriBoundary<S>& bdry2, const Vector3d& transl, const Matrix3d& rot) {
    MC_ASSERT_MSG(bdry1.pair_ == 0 && bdry2.pair_ == 0, "Boundary already paired");
    const Matrix3d rrt = rot * rot.transpose();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        MC_ASSERT_MSG(std::abs(rrt(i, j) - (i == j ? 1. : 0.)) <= 1e-12, "Rotation matrix must be orthogonal");
    const Vector3d rn = rot * bdry1.normal();
    for (int i = 0; i < 3; ++i) MC_ASSERT_MSG(std::abs(rn(i) + bdry2.normal()(i)) <= 1e-12, "Rotation matrix incorrect");
    const Vector3d transform = rn * (-bdry1.offset()) + transl;
    MC_ASSERT_MSG(isApprox(bdry2.normal().dot(transform), -bdry2.offset()), "Translation vector incorrect");

    const double T1 = bdry1.T_, T2 = bdry2.T_;
    bdry1.T_ = T1 - T2; bdry2.T_ = T2 - T1;
    bdry1.rot_ = rot; bdry2.rot_ = rot.transpose();
    bdry1.transl_ = transl; bdry2.transl_ = -(rot.transpose() * transl);
    bdry1.pair_ = &bdry2; bdry2.pair_ = &bdry1;
}
#endif
