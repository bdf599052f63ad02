// geometry.cpp — host mirror implementations for boundary.h / subdomain.h / domain.h.
// Reference: boundary.cpp:28-110,361-369,387-431 ; subdomain.cpp:41-71,108-116,148-159,199-236 ;
// domain.cpp:28-116,137-148,188-199,352-385,429-468,509-540.
#include <cstring>
#include <iostream>
#include <sstream>
#include "domain.h"

//---------------------------------------- helpers
// Rotation taking +z onto n, i.e. Quaternion::FromTwoVectors(UnitZ, n).matrix() of Eigen.  For n = -z
// Eigen picks an SVD-dependent axis; any axis perpendicular to z is statistically equivalent because
// the rotated distribution (drawAniso) is azimuthally uniform.  We turn about x.
Matrix3d rotMatrix(const Vector3d& n) {
    const double len = n.norm();
    if (!(len > 0.)) return Matrix3d::Identity();            // zero gradient: never used for emission
    const Vector3d z = Vector3d::UnitZ(), u = n / len;
    const double c = u.dot(z);
    double qx, qy, qz, qw;
    if (c < -1. + 1e-12) { qx = 1.; qy = 0.; qz = 0.; qw = 0.; }
    else {
        const Vector3d axis = z.cross(u);
        const double s = std::sqrt((1. + c) * 2.), invs = 1. / s;
        qx = axis(0) * invs; qy = axis(1) * invs; qz = axis(2) * invs; qw = s * 0.5;
    }
    const double tx = 2. * qx, ty = 2. * qy, tz = 2. * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    Matrix3d r;
    r(0, 0) = 1. - (tyy + tzz); r(0, 1) = txy - twz;        r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;        r(1, 1) = 1. - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;        r(2, 1) = tyz + twx;        r(2, 2) = 1. - (txx + tyy);
    return r;
}

//---------------------------------------- Boundary
Boundary::Boundary() : sdom_(0), off_(0.) {}
Boundary::Boundary(const Vector3d& o, const Vector3d& n) : sdom_(0), n_(n.normalized()) { off_ = -n_.dot(o); }
Boundary::Boundary(const Vector3d& o, const Shape& s) : sdom_(0), n_(s.normal()) {
    MC_ASSERT_MSG(s.isInit(), "Shape not initialized");
    off_ = -n_.dot(o);                                       // Hyperplane(n, e): offset = -n.e
}
Boundary::~Boundary() {}
bool Boundary::isInit() const { return sdom_ != 0; }

void Boundary::describe(mcb_plane_desc& d) const {
    std::memset(&d, 0, sizeof d);
    for (int k = 0; k < 3; ++k) d.normal[k] = n_(k);
    d.offset = off_; d.kind = kind();
    const Matrix3d r = rotMatrix(n_);                        // DiffBoundary::rot_ / EmitBoundary::rot_
    for (int k = 0; k < 9; ++k) d.rot[k] = r.m[k];
    d.shape = MCB_SHAPE_NONE;
}

void EmitBoundary::describe(mcb_plane_desc& d) const {
    Boundary::describe(d);
    d.T = T_;
    for (int k = 0; k < 3; ++k) d.origin[k] = o_(k);
    const std::vector<Vector3d> v = shape().verts();
    MC_ASSERT_MSG(v.size() <= MCB_MAX_VERTS, "Too many shape vertices");
    d.shape = shape().kind(); d.nvert = (int32_t)v.size();
    for (size_t i = 0; i < v.size(); ++i) for (int k = 0; k < 3; ++k) d.verts[3 * i + k] = v[i](k);
}

void makePair(InterBoundary& bdry1, InterBoundary& bdry2) {
    const Vector3d a = bdry1.normal(), b = bdry2.normal();
    for (int i = 0; i < 3; ++i) MC_ASSERT_MSG(std::abs(a(i) + b(i)) <= 1e-12, "Boundary normals not antiparallel");
    MC_ASSERT_MSG(isApprox(bdry1.offset(), -bdry2.offset()), "Boundaries planes not the same");
    bdry1.pairs_.push_back(&bdry2);
    bdry2.pairs_.push_back(&bdry1);
}

//---------------------------------------- Subdomain
Subdomain::Subdomain() : vol_(0.), accum_(-1), eps_(0.) {}

Subdomain::Subdomain(double vol, const Vector3d& o, const Matrix3d& mat, const Vector3l& div)
    : vol_(vol), o_(o), mat_(mat), inv_(mat.inverse()), div_(div) {
    double longest = 0.;
    for (int c = 0; c < 3; ++c) longest = std::max(longest, mat.col(c).norm());
    eps_ = 100. * Dbl::epsilon() * longest;
    int npos = 0; bool neg = false;
    for (int d = 0; d < 3; ++d) {
        shape_(d) = std::max(div(d), 1l); max_(d) = shape_(d) - 1;
        if (div(d) < 0) neg = true;
        if (div(d) > 0) ++npos;
    }
    if (neg) { shape_ = Vector3l(); max_ = Vector3l(); accum_ = -2; }      // not tallied
    else if (npos == 0) accum_ = -1;                                       // one cell
    else if (npos == 1) { int dir = 0; for (int d = 1; d < 3; ++d) if (div(d) > div(dir)) dir = d; accum_ = dir; }
    else accum_ = npos + 1;                                                // 3 (2-D) or 4 (3-D)
}
Subdomain::~Subdomain() {}

bool Subdomain::isInit() const {
    if (vol_ == 0 || bdryPtrs_.empty()) return false;
    for (const Boundary* b : bdryPtrs_) if (!b->isInit()) return false;
    return true;
}
bool Subdomain::isInside(const Vector3d& pos) const {
    for (const Boundary* b : bdryPtrs_) if (b->distance(pos) < -eps_) return false;
    return true;
}
Vector3d Subdomain::coord(const Vector3d& pos) const {
    const Vector3d t = inv_ * (pos - o_);
    return Vector3d((double)div_(0) * t(0), (double)div_(1) * t(1), (double)div_(2) * t(2));
}
Vector3l Subdomain::coord2index(const Vector3d& c) const {
    Vector3l idx;
    for (int d = 0; d < 3; ++d) idx(d) = std::min(std::max(static_cast<long>(std::floor(c(d))), 0l), max_(d));
    return idx;
}
void Subdomain::addBdry(Boundary* bdry) { bdry->sdom(this); bdryPtrs_.push_back(bdry); }
void Subdomain::addBdry(EmitBoundary* bdry) {
    addBdry(static_cast<Boundary*>(bdry));
    if (bdry->emitWeight() != 0.) emitPtrs_.push_back(static_cast<Emitter*>(bdry));   // subdomain.cpp:205-212
}
void Subdomain::describe(mcb_sdom_desc& d) const {
    std::memset(&d, 0, sizeof d);
    for (int k = 0; k < 3; ++k) { d.origin[k] = o_(k); d.div[k] = div_(k); d.shape[k] = shape_(k); d.max[k] = max_(k); }
    for (int k = 0; k < 9; ++k) { d.mat[k] = mat_.m[k]; d.inv[k] = inv_.m[k]; }
    d.accum = accum_; d.cell = cellKind(); d.eps = eps_; d.vol = vol_;
    const Matrix3d id = Matrix3d::Identity();
    for (int k = 0; k < 9; ++k) d.emit_rot[k] = id.m[k];
}

EmitSubdomain::EmitSubdomain(double vol, const Vector3d& o, const Matrix3d& mat, const Vector3l& div, const Vector3d& gradT)
    : Subdomain(vol, o, mat, div), gradT_(gradT), rot_(rotMatrix(gradT)) {}
void EmitSubdomain::describe(mcb_sdom_desc& d) const {
    Subdomain::describe(d);
    for (int k = 0; k < 3; ++k) d.grad_t[k] = gradT_(k);
    for (int k = 0; k < 9; ++k) d.emit_rot[k] = rot_.m[k];
}

//---------------------------------------- simplex / prism helpers
// Volume of grid cell `index` of a triangular prism: the part of the cell's base rectangle under the diagonal, by
// inclusion-exclusion on the normalised distance f to the diagonal.  (The normalisation treats a full cell as
// vol / prod(shape); kept as in the reference, subdomain.cpp:283-307.)
double TriangularPrismImpl::cellVol(const Vector3l& index, const Vector3l& shape, double vol) {
    const double s0 = (double)shape(0), s1 = (double)shape(1);
    const double f0 = 1. - ((double)index(0) / s0 + (double)index(1) / s1);
    if (f0 <= 0.) return 0.;
    const double f1 = f0 - (1. / s0 + 1. / s1);
    if (f1 >= 0.) return vol / shape.prod();
    double frac = std::pow(f0, 2);
    const double step[2] = {1. / s0, 1. / s1};
    for (int i = 0; i < 2; i++) {
        const double f = f0 - step[i];
        if (f > 0.) frac += -1 * std::pow(f, 2);
    }
    return vol * frac / (2. * shape(2));
}
// Same for a tetrahedron (subdomain.cpp:322-349): corners one step (sign -) and two steps (sign +) away
double TetrahedronImpl::cellVol(const Vector3l& index, const Vector3l& shape, double vol) {
    const double s[3] = {(double)shape(0), (double)shape(1), (double)shape(2)};
    const double f0 = 1. - ((double)index(0) / s[0] + (double)index(1) / s[1] + (double)index(2) / s[2]);
    if (f0 <= 0.) return 0.;
    const double f1 = f0 - (1. / s[0] + 1. / s[1] + 1. / s[2]);
    if (f1 >= 0.) return vol / shape.prod();
    static const int corner[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
    double frac = std::pow(f0, 3);
    for (int i = 0; i < 6; i++) {
        const int n = corner[i][0] + corner[i][1] + corner[i][2];
        const int sign = (n % 2) ? -1 : 1;
        const double f = f0 - ((double)corner[i][0] / s[0] + (double)corner[i][1] / s[1] + (double)corner[i][2] / s[2]);
        if (f > 0.) frac += sign * std::pow(f, 3);
    }
    return vol * frac / 6.;
}
Matrix3d PrismImpl::matBase(const std::vector<Vector3d>& mat) {
    MC_ASSERT_MSG(mat.size() >= 4, "Incorrect number of matrix columns");
    return Matrix3d::Columns(mat[1], mat[mat.size() - 1], mat[0]);
}
// Sub-wedge volumes as the reference computes them (subdomain.cpp:385-394, 417-426): columns i and i+1 for i < N-2.
std::vector<double> PrismImpl::volume(const std::vector<Vector3d>& mat, double div) {
    const size_t N = mat.size();
    std::vector<double> vol(N - 2);
    for (size_t i = 0; i + 2 < N; ++i) vol[i] = mat[i].cross(mat[i + 1]).dot(mat[0]) / div;
    return vol;
}

//---------------------------------------- Domain
bool Domain::isInit() const {
    if (sdomPtrs_.empty()) return false;
    for (const Subdomain* s : sdomPtrs_) if (!s->isInit()) return false;
    return true;
}
const Subdomain* Domain::locate(const Vector3d& pos) const {
    for (const Subdomain* s : sdomPtrs_) if (s->isInside(pos)) return s;
    return 0;
}
void Domain::addSdom(const Subdomain* sdom) {
    sdomPtrs_.push_back(sdom);
    emitPtrs_.insert(emitPtrs_.end(), sdom->emitPtrs().begin(), sdom->emitPtrs().end());
}
void Domain::addSdom(const EmitSubdomain* sdom) {
    if (sdom->emitWeight() != 0.) emitPtrs_.push_back(static_cast<const Emitter*>(sdom));
    addSdom(static_cast<const Subdomain*>(sdom));
}
std::ostream& operator<<(std::ostream& os, const Domain& dom) { return os << dom.info(); }

std::string Domain::describe(const char* name, const void* self, const VectorXd& dim, const VectorXl& div, double dT) {
    std::ostringstream ss;
    ss << name << " " << self << std::endl;
    ss << "  dim: [";
    for (size_t i = 0; i < dim.size(); ++i) ss << (i ? " " : "") << dim[i];
    ss << "]" << std::endl << "  div: [";
    for (size_t i = 0; i < div.size(); ++i) ss << (i ? " " : "") << div[i];
    ss << "]" << std::endl << "  dT:  " << dT;
    return ss.str();
}

namespace {
VectorXd vec(const Vector3d& v) { return VectorXd{v(0), v(1), v(2)}; }
VectorXl vec(const Vector3l& v) { return VectorXl{v(0), v(1), v(2)}; }
Matrix3Xd points(std::initializer_list<Vector3d> p) { Matrix3Xd m; m.c.assign(p.begin(), p.end()); return m; }
VectorXd checked(const VectorXd& v, size_t n, const char* what) { MC_ASSERT_MSG(v.size() == n, what); return v; }
VectorXl checked(const VectorXl& v, size_t n, const char* what) { MC_ASSERT_MSG(v.size() == n, what); return v; }
}

// one periodic box, heat driven by a volumetric gradient -dT/dim0 along x
BulkDomain::BulkDomain(const Vector3d& dim, const Vector3l& div, double dT)
    : dim_(dim), div_(div), dT_(dT),
      sdom_(Vector3d::Zero(), Matrix3d::Diagonal(dim), div, Vector3d(-dT / dim(0), 0., 0.)) {
    makePair(sdom_.bdry<0>(), sdom_.bdry<3>(), Vector3d(dim_(0), 0., 0.));
    addSdom(&sdom_);
}
std::string BulkDomain::info() const { return describe("BulkDomain", static_cast<const Domain*>(this), vec(dim_), vec(div_), dT_); }
Matrix3Xd BulkDomain::checkpoints() const { return points({0.5 * dim_}); }

FilmDomain::FilmDomain(const Vector3d& dim, const Vector3l& div, double dT)
    : dim_(dim), div_(div), dT_(dT),
      sdom_(Vector3d::Zero(), Matrix3d::Diagonal(dim), div, Vector3d(-dT / dim(0), 0., 0.)) {
    makePair(sdom_.bdry<0>(), sdom_.bdry<3>(), Vector3d(dim_(0), 0., 0.));
    addSdom(&sdom_);
}
std::string FilmDomain::info() const { return describe("FilmDomain", static_cast<const Domain*>(this), vec(dim_), vec(div_), dT_); }
Matrix3Xd FilmDomain::checkpoints() const { return points({0.5 * dim_}); }

SlabDomain::SlabDomain(const Vector3d& dim, const Vector3l& div, double dT)
    : dim_(dim), div_(div), dT_(dT),
      sdom_(Vector3d::Zero(), Matrix3d::Diagonal(dim), div, Vector3d::Zero(), VectorXd{dT / 2., 0., 0., -dT / 2., 0., 0.}) {
    addSdom(&sdom_);
}
std::string SlabDomain::info() const { return describe("SlabDomain", static_cast<const Domain*>(this), vec(dim_), vec(div_), dT_); }
Matrix3Xd SlabDomain::checkpoints() const { return points({0.5 * dim_}); }

WireDomain::WireDomain(const Vector3d& dim, const Vector3l& div, double dT)
    : dim_(dim), div_(div), dT_(dT),
      sdom_(Vector3d::Zero(), Matrix3d::Diagonal(dim), div, Vector3d(-dT / dim(0), 0., 0.)) {
    makePair(sdom_.bdry<0>(), sdom_.bdry<3>(), Vector3d(dim_(0), 0., 0.));
    addSdom(&sdom_);
}
std::string WireDomain::info() const { return describe("WireDomain", static_cast<const Domain*>(this), vec(dim_), vec(div_), dT_); }
Matrix3Xd WireDomain::checkpoints() const { return points({0.5 * dim_}); }

// T-junction of a 2a-long bar with two stubs (three boxes)
JctDomain::JctDomain(const VectorXd& dim, const VectorXl& div, double dT)
    : dim_(checked(dim, 4, "JctDomain needs 4 dimensions")), div_(checked(div, 4, "JctDomain needs 4 divisions")), dT_(dT),
      s0_(Vector3d(0., 0., 0.), Matrix3d::Diagonal(2. * dim[0], dim[1], dim[3]), Vector3l(2 * div[0], div[1], div[3]),
          Vector3d(-dT / (2. * dim[0]), 0., 0.)),
      s1_(Vector3d(0., dim[1], 0.), Matrix3d::Diagonal(dim[0], dim[2], dim[3]), Vector3l(div[0], div[2], div[3]),
          Vector3d(-dT / (2. * dim[0]), 0., 0.)),
      s2_(Vector3d(dim[0], dim[1], 0.), Matrix3d::Diagonal(dim[0], dim[2], dim[3]), Vector3l(div[0], div[2], div[3]),
          Vector3d(-dT / (2. * dim[0]), 0., 0.)) {
    makePair(s0_.bdry<4>(), s1_.bdry<1>());
    makePair(s0_.bdry<4>(), s2_.bdry<1>());
    makePair(s1_.bdry<3>(), s2_.bdry<0>());
    const Vector3d transl(2. * dim_[0], 0., 0.);
    makePair(s0_.bdry<0>(), s0_.bdry<3>(), transl);
    makePair(s1_.bdry<0>(), s2_.bdry<3>(), transl);
    addSdom(&s0_); addSdom(&s1_); addSdom(&s2_);
}
std::string JctDomain::info() const { return describe("JctDomain", static_cast<const Domain*>(this), dim_, div_, dT_); }
Matrix3Xd JctDomain::checkpoints() const {
    return points({Vector3d(dim_[0], 0.5 * dim_[1], 0.5 * dim_[3]),
                   Vector3d(0.5 * dim_[0], dim_[1] + 0.5 * dim_[2], 0.5 * dim_[3]),
                   Vector3d(1.5 * dim_[0], dim_[1] + 0.5 * dim_[2], 0.5 * dim_[3])});
}

TeeDomain::TeeDomain(const VectorXd& dim, const VectorXl& div, double dT)
    : dim_(checked(dim, 5, "TeeDomain needs 5 dimensions")), div_(checked(div, 5, "TeeDomain needs 5 divisions")), dT_(dT),
      s0_(Vector3d::Zero(), Matrix3d::Diagonal(dim[0], dim[2], dim[4]), Vector3l(div[0], div[2], div[4]),
          Vector3d(-dT / (2. * dim[0] + dim[1]), 0., 0.)),
      s1_(Vector3d(dim[0], 0., 0.), Matrix3d::Diagonal(dim[1], dim[2], dim[4]), Vector3l(div[1], div[2], div[4]),
          Vector3d(-dT / (2. * dim[0] + dim[1]), 0., 0.)),
      s2_(Vector3d(dim[0], dim[2], 0.), Matrix3d::Diagonal(dim[1], dim[3], dim[4]), Vector3l(div[1], div[3], div[4]),
          Vector3d(-dT / (2. * dim[0] + dim[1]), 0., 0.)),
      s3_(Vector3d(dim[0] + dim[1], 0., 0.), Matrix3d::Diagonal(dim[0], dim[2], dim[4]), Vector3l(div[0], div[2], div[4]),
          Vector3d(-dT / (2. * dim[0] + dim[1]), 0., 0.)) {
    makePair(s0_.bdry<3>(), s1_.bdry<0>());
    makePair(s1_.bdry<4>(), s2_.bdry<1>());
    makePair(s1_.bdry<3>(), s3_.bdry<0>());
    makePair(s0_.bdry<0>(), s3_.bdry<3>(), Vector3d(2. * dim_[0] + dim_[1], 0., 0.));
    addSdom(&s0_); addSdom(&s1_); addSdom(&s2_); addSdom(&s3_);
}
std::string TeeDomain::info() const { return describe("TeeDomain", static_cast<const Domain*>(this), dim_, div_, dT_); }
Matrix3Xd TeeDomain::checkpoints() const {
    return points({Vector3d(0.5 * dim_[0], 0.5 * dim_[2], 0.5 * dim_[4]),
                   Vector3d(dim_[0] + 0.5 * dim_[1], 0.5 * dim_[2], 0.5 * dim_[4]),
                   Vector3d(dim_[0] + 0.5 * dim_[1], dim_[2] + 0.5 * dim_[3], 0.5 * dim_[4]),
                   Vector3d(1.5 * dim_[0] + dim_[1], 0.5 * dim_[2], 0.5 * dim_[4])});
}

// quarter cross-section of a hollow square tube, periodic along x (three boxes in an L)
TubeDomain::TubeDomain(const VectorXd& dim, const VectorXl& div, double dT)
    : dim_(checked(dim, 4, "TubeDomain needs 4 dimensions")), div_(checked(div, 4, "TubeDomain needs 4 divisions")), dT_(dT),
      s0_(Vector3d(0., dim[1], 0.), Matrix3d::Diagonal(dim[0], dim[3], dim[2]), Vector3l(div[0], div[3], div[2]),
          Vector3d(-dT / dim[0], 0., 0.)),
      s1_(Vector3d(0., dim[1], dim[2]), Matrix3d::Diagonal(dim[0], dim[3], dim[3]), Vector3l(div[0], div[3], div[3]),
          Vector3d(-dT / dim[0], 0., 0.)),
      s2_(Vector3d(0., 0., dim[2]), Matrix3d::Diagonal(dim[0], dim[1], dim[3]), Vector3l(div[0], div[1], div[3]),
          Vector3d(-dT / dim[0], 0., 0.)) {
    makePair(s1_.bdry<1>(), s2_.bdry<4>());
    makePair(s1_.bdry<2>(), s0_.bdry<5>());
    const Vector3d transl(dim_[0], 0., 0.);
    makePair(s0_.bdry<0>(), s0_.bdry<3>(), transl);
    makePair(s1_.bdry<0>(), s1_.bdry<3>(), transl);
    makePair(s2_.bdry<0>(), s2_.bdry<3>(), transl);
    addSdom(&s0_); addSdom(&s1_); addSdom(&s2_);
}
std::string TubeDomain::info() const { return describe("TubeDomain", static_cast<const Domain*>(this), dim_, div_, dT_); }
Matrix3Xd TubeDomain::checkpoints() const {
    return points({Vector3d(0.5 * dim_[0], dim_[1] + 0.5 * dim_[3], 0.5 * dim_[2]),
                   Vector3d(0.5 * dim_[0], dim_[1] + 0.5 * dim_[3], dim_[2] + 0.5 * dim_[3]),
                   Vector3d(0.5 * dim_[0], 0.5 * dim_[1], dim_[2] + 0.5 * dim_[3])});
}

// hexagonal prism, periodic along x (domain.h:142-143, domain.cpp:238-256).  The reference's (dim, dT) constructor
// never calls init(), which leaves the domain unusable; here it does.
static std::vector<Vector3d> hexColumns(const VectorXd& d) {
    return {Vector3d(d[0], 0., 0.), Vector3d(0., d[1], -d[3]), Vector3d(0., 2. * d[1], 0.), Vector3d(0., 2. * d[1], d[2]),
            Vector3d(0., d[1], d[2] + d[3]), Vector3d(0., 0., d[2])};
}
HexDomain::HexDomain(const VectorXd& dim, double dT)
    : dim_(checked(dim, 4, "HexDomain needs 4 dimensions")), dT_(dT),
      sdom_(Vector3d::Zero(), hexColumns(dim), 0, Vector3d(-dT / dim[0], 0., 0.)) {
    makePair(sdom_.bottom(), sdom_.top(), Vector3d(dim_[0], 0., 0.));
    addSdom(&sdom_);
}
std::string HexDomain::info() const { return describe("HexDomain", static_cast<const Domain*>(this), dim_, VectorXl(), dT_); }
Matrix3Xd HexDomain::checkpoints() const {
    return points({Vector3d(0.5 * dim_[0], 0.5 * dim_[1], 0.5 * dim_[2]), Vector3d(0.5 * dim_[0], 1.5 * dim_[1], 0.5 * dim_[2])});
}

// square pyramid with its apex above the centre of the y-z base, all faces specular (domain.h:165, domain.cpp:299-314)
PyrDomain::PyrDomain(const Vector3d& dim, double dT)
    : dim_(dim), dT_(dT),
      sdom_(Vector3d::Zero(), {Vector3d(dim(0), 0.5 * dim(1), 0.5 * dim(2)), Vector3d(0., dim(1), 0.), Vector3d(0., dim(1), dim(2)), Vector3d(0., 0., dim(2))},
            0, Vector3d(-dT / dim(0), 0., 0.)) {
    addSdom(&sdom_);
}
std::string PyrDomain::info() const { return describe("PyrDomain", static_cast<const Domain*>(this), vec(dim_), VectorXl(), dT_); }
Matrix3Xd PyrDomain::checkpoints() const { return points({0.5 * dim_}); }

//---------------------------------------- OctetDomain (domain.cpp:576-1280)
namespace {
#include "octet_table.inc"
typedef SpecBoundary Spec; typedef DiffBoundary Diff; typedef InterBoundary Inter;
typedef PeriBoundary<Parallelogram> PeriP; typedef PeriBoundary<Polygon<4> > Peri4;
typedef std::tuple<MCB_OCTET_CELL_TYPES> OctetTypes;
static_assert(std::tuple_size<OctetTypes>::value == 42, "octet table");

double evalForm(const OctetForm& f, double s, double a, double t) {
    const double r2 = std::sqrt(2.);
    return ((f.ps + f.qs * r2) * s + (f.pa + f.qa * r2) * a + (f.pt + f.qt * r2) * t) / 16.;
}
// the two constructor shapes: (o, 3 x 3 matrix, div vector, gradT) for parallelepipeds / triangular prisms,
// (o, N columns, one div, gradT) for prisms / pyramids
template <class T> std::unique_ptr<EmitSubdomain> makeCell(const Vector3d& o, const std::vector<Vector3d>& cols, const Vector3l& div, const Vector3d& grad) {
    if constexpr (std::is_constructible<T, const Vector3d&, const Matrix3d&, const Vector3l&, const Vector3d&>::value)
        return std::unique_ptr<EmitSubdomain>(new T(o, Matrix3d::Columns(cols.at(0), cols.at(1), cols.at(2)), div, grad));
    else
        return std::unique_ptr<EmitSubdomain>(new T(o, cols, div(0), grad));
}
template <size_t... I>
std::unique_ptr<EmitSubdomain> makeCellAt(size_t i, const Vector3d& o, const std::vector<Vector3d>& cols, const Vector3l& div, const Vector3d& grad,
                                           std::index_sequence<I...>) {
    typedef std::unique_ptr<EmitSubdomain> (*Fn)(const Vector3d&, const std::vector<Vector3d>&, const Vector3l&, const Vector3d&);
    static const Fn table[] = {&makeCell<typename std::tuple_element<I, OctetTypes>::type>...};
    return table[i](o, cols, div, grad);
}
Boundary* bdryOf(const std::vector<std::unique_ptr<EmitSubdomain> >& cells, int cell, int b) {
    return const_cast<Boundary*>(cells.at((size_t)cell)->bdryPtrs().at((size_t)b));
}
}

OctetDomain::OctetDomain(const VectorXd& dim, const VectorXl& div, double dT)
    : dim_(checked(dim, 4, "OctetDomain needs 4 dimensions")), div_(checked(div, 4, "OctetDomain needs 4 divisions")), dT_(dT) {
    const double r2 = std::sqrt(2.);
    const double s = dim_[0], t = dim_[3];
    const double a = PI / 5. * (dim_[1] + dim_[2]) - (4. - PI) / 5. * t, b = a / 4.;
    MC_ASSERT_MSG(b > t, "Invalid dimensions");                                  // domain.cpp:702-704
    MC_ASSERT_MSG(a > (r2 + 1.) * (b + t), "Invalid dimensions");
    MC_ASSERT_MSG(s / 2. > a + (r2 + 1.) * (b + t) + t, "Invalid dimensions");
    // the gradient runs along the strut, from the top node to the bottom node of the cell
    const double g = dT / (s - 2. * a - 2. * (r2 + 1.) * b - 2. * (2. + r2) * t);
    const Vector3d gradT(g, 0., -g), noGrad;
    for (size_t i = 0; i < 42; ++i) {
        const OctetCell& c = kOctetCells[i];
        const Vector3d o(evalForm(c.o[0], s, a, t), evalForm(c.o[1], s, a, t), evalForm(c.o[2], s, a, t));
        std::vector<Vector3d> cols;
        for (int k = 0; k < c.ncol; ++k) cols.push_back(Vector3d(evalForm(c.col[k][0], s, a, t), evalForm(c.col[k][1], s, a, t), evalForm(c.col[k][2], s, a, t)));
        Vector3l dv;
        for (int k = 0; k < 3; ++k) dv(k) = c.div[k] < 0 ? -1 : div_[(size_t)c.div[k]];
        cells_.push_back(makeCellAt(i, o, cols, dv, c.grad ? gradT : noGrad, std::make_index_sequence<42>()));
    }
    for (const auto& pr : kOctetInter) {
        InterBoundary* x = dynamic_cast<InterBoundary*>(bdryOf(cells_, pr[0], pr[1]));
        InterBoundary* y = dynamic_cast<InterBoundary*>(bdryOf(cells_, pr[2], pr[3]));
        MC_ASSERT_MSG(x && y, "octet table: Inter pair on a boundary of another class");
        makePair(*x, *y);
    }
    const Vector3d transl(s / 2., 0., -s / 2.);
    const Matrix3d rot = Matrix3d::Diagonal(-1., 1., 1.);
    for (const auto& pr : kOctetPeri) {
        Boundary* x = bdryOf(cells_, pr[0], pr[1]); Boundary* y = bdryOf(cells_, pr[2], pr[3]);
        if (Peri4* p4 = dynamic_cast<Peri4*>(x)) { Peri4* q4 = dynamic_cast<Peri4*>(y); MC_ASSERT_MSG(q4, "octet table: periodic pair of two shapes"); makePair(*p4, *q4, transl, rot); }
        else { PeriP* pp = dynamic_cast<PeriP*>(x); PeriP* qp = dynamic_cast<PeriP*>(y); MC_ASSERT_MSG(pp && qp, "octet table: periodic pair of two shapes"); makePair(*pp, *qp, transl, rot); }
    }
    for (const auto& c : cells_) addSdom(c.get());
}
std::string OctetDomain::info() const { return describe("OctetDomain", static_cast<const Domain*>(this), dim_, div_, dT_); }

Matrix3Xd OctetDomain::checkpoints() const {
    std::vector<Vector3d> pts;
    const double s = dim_[0], t = dim_[3], a = PI / 5. * (dim_[1] + dim_[2]) - (4. - PI) / 5. * t;
    for (size_t i = 0; i < 42; ++i) {
        const OctetCell& c = kOctetCells[i];
        Vector3d o(evalForm(c.o[0], s, a, t), evalForm(c.o[1], s, a, t), evalForm(c.o[2], s, a, t));
        std::vector<Vector3d> col;
        for (int k = 0; k < c.ncol; ++k) col.push_back(Vector3d(evalForm(c.col[k][0], s, a, t), evalForm(c.col[k][1], s, a, t), evalForm(c.col[k][2], s, a, t)));
        Vector3d p = o;
        if (c.kind == MCB_CELL_PARALLELEPIPED) p = o + 0.5 * (col[0] + col[1] + col[2]);
        else if (c.kind == MCB_CELL_TRIPRISM) p = o + (1. / 3.) * (col[0] + col[1]) + 0.5 * col[2];
        else {                                                      // base polygon o, o + col 1 .. o + col N-1; col 0 = axis / apex
            Vector3d base; for (int k = 1; k < c.ncol; ++k) base = base + col[(size_t)k];
            base = (1. / c.ncol) * base;
            p = c.kind == MCB_CELL_PRISM ? o + base + 0.5 * col[0] : o + 0.75 * base + 0.25 * col[0];
        }
        pts.push_back(p);
    }
    Matrix3Xd m; m.c = pts;
    return m;
}

std::vector<double> OctetDomain::averageWeights() const {
    std::vector<double> w;
    for (size_t i = 0; i < cells_.size(); ++i) {
        const Subdomain* sd = cells_[i].get();
        const Vector3l shp = sd->shape();
        for (long k = 0; k < shp(2); ++k) for (long j = 0; j < shp(1); ++j) for (long ii = 0; ii < shp(0); ++ii)
            w.push_back((i >= 16 && i < 26 && k == 0) ? sd->cellVol(Vector3l(ii, j, k)) : 0.);      // WeightF, domain.cpp:1259-1280
    }
    return w;
}
ArrayXXd OctetDomain::average(const ArrayXXd& data) const {
    const std::vector<double> w = averageWeights();
    MC_ASSERT_MSG((long)w.size() == data.cols(), "average: column count");
    double wsum = 0.; for (double x : w) wsum += x;
    ArrayXXd avg(data.rows(), 1);
    for (long r = 0; r < data.rows(); ++r) { double acc = 0.; for (long c = 0; c < data.cols(); ++c) acc += data(r, c) * w[(size_t)c]; avg(r, 0) = acc / wsum; }
    return avg;
}
