// domain.h — host mirror of Domain and the box-celled domains the reference ships (domain.h:38-296,
// domain.cpp:28-570): Bulk, Film, Jct, Tee, Tube.  SlabDomain and WireDomain are NOT in the reference;
// they are the two box variants BASELINE.json's configs need (isothermal walls; diffuse wire) and use
// only reference building blocks.  HexDomain / PyrDomain (non-box cells) and the 42-subdomain OctetDomain are built.
#ifndef MCB_HOST_DOMAIN_H
#define MCB_HOST_DOMAIN_H
#include <iosfwd>
#include <memory>
#include <string>
#include "subdomain.h"

class Domain {
protected:
    typedef SpecBoundary Spec;
    typedef DiffBoundary Diff;
    typedef InterBoundary Inter;
    typedef PeriBoundary<Parallelogram> PeriP;
    typedef IsotBoundary<Parallelogram> IsotP;
private:
    Subdomain::Pointers sdomPtrs_;
    Emitter::Pointers emitPtrs_;
    unsigned long uid_ = mcNextUid();
    virtual std::string info() const = 0;
public:
    Domain() {}
    Domain(const Domain&) = delete;
    Domain& operator=(const Domain&) = delete;
    virtual ~Domain() {}

    bool isInit() const;
    bool isInside(const Vector3d& pos) const { return locate(pos) != 0; }
    const Subdomain* locate(const Vector3d& pos) const;
    const Subdomain::Pointers& sdomPtrs() const { return sdomPtrs_; }
    const Emitter::Pointers& emitPtrs() const { return emitPtrs_; }
    unsigned long uid() const { return uid_; }

    virtual Matrix3Xd checkpoints() const = 0;
    virtual ArrayXXd average(const ArrayXXd& data) const { return data; }     // domain.cpp:78-81

    friend std::ostream& operator<<(std::ostream& os, const Domain& dom);
protected:
    void addSdom(const Subdomain* sdom);
    void addSdom(const EmitSubdomain* sdom);
    static std::string describe(const char* name, const void* self, const VectorXd& dim, const VectorXl& div, double dT);
};

class BulkDomain : public Domain {
public:
    typedef Parallelepiped<PeriP, Spec, Spec> Sdom;
private:
    Vector3d dim_; Vector3l div_; double dT_;
    Sdom sdom_;
    std::string info() const;
public:
    BulkDomain(const Vector3d& dim, const Vector3l& div, double dT);
    Matrix3Xd checkpoints() const;
};

class FilmDomain : public Domain {
public:
    typedef Parallelepiped<PeriP, Diff, Spec> Sdom;
private:
    Vector3d dim_; Vector3l div_; double dT_;
    Sdom sdom_;
    std::string info() const;
public:
    FilmDomain(const Vector3d& dim, const Vector3l& div, double dT);
    Matrix3Xd checkpoints() const;
};

// not in the reference: slab between two isothermal walls at +dT/2 (x = 0) and -dT/2 (x = dim0)
class SlabDomain : public Domain {
public:
    typedef Parallelepiped<IsotP, Spec, Spec> Sdom;
private:
    Vector3d dim_; Vector3l div_; double dT_;
    Sdom sdom_;
    std::string info() const;
public:
    SlabDomain(const Vector3d& dim, const Vector3l& div, double dT);
    Matrix3Xd checkpoints() const;
};

// not in the reference: wire periodic along x with four diffuse side walls
class WireDomain : public Domain {
public:
    typedef Parallelepiped<PeriP, Diff, Diff> Sdom;
private:
    Vector3d dim_; Vector3l div_; double dT_;
    Sdom sdom_;
    std::string info() const;
public:
    WireDomain(const Vector3d& dim, const Vector3l& div, double dT);
    Matrix3Xd checkpoints() const;
};

class HexDomain : public Domain {
public:
    typedef PeriBoundary<Polygon<6> > Peri6;
    typedef Prism<Peri6, Peri6, std::tuple<Spec, Spec, Spec, Spec, Spec, Spec> > Sdom;
private:
    VectorXd dim_; double dT_;
    Sdom sdom_;
    std::string info() const;
public:
    HexDomain(const VectorXd& dim, double dT);                      // 4 dims
    Matrix3Xd checkpoints() const;
};

class PyrDomain : public Domain {
public:
    typedef Pyramid<Spec, std::tuple<Spec, Spec, Spec, Spec> > Sdom;
private:
    Vector3d dim_; double dT_;
    Sdom sdom_;
    std::string info() const;
public:
    PyrDomain(const Vector3d& dim, double dT);
    Matrix3Xd checkpoints() const;
};

class JctDomain : public Domain {
public:
    typedef Parallelepiped<PeriP, Spec, Diff, PeriP, Inter, Diff> Sdom0;
    typedef Parallelepiped<PeriP, Inter, Diff, Inter, Spec, Diff> Sdom1;
    typedef Parallelepiped<Inter, Inter, Diff, PeriP, Spec, Diff> Sdom2;
private:
    VectorXd dim_; VectorXl div_; double dT_;
    Sdom0 s0_; Sdom1 s1_; Sdom2 s2_;
    std::string info() const;
public:
    JctDomain(const VectorXd& dim, const VectorXl& div, double dT);     // 4 dims, 4 divs
    Matrix3Xd checkpoints() const;
};

class TeeDomain : public Domain {
public:
    typedef Parallelepiped<PeriP, Diff, Spec, Inter, Diff, Spec> Sdom0;
    typedef Parallelepiped<Inter, Spec, Spec, Inter, Inter, Spec> Sdom1;
    typedef Parallelepiped<Diff, Inter, Spec, Diff, Spec, Spec> Sdom2;
    typedef Parallelepiped<Inter, Diff, Spec, PeriP, Diff, Spec> Sdom3;
private:
    VectorXd dim_; VectorXl div_; double dT_;
    Sdom0 s0_; Sdom1 s1_; Sdom2 s2_; Sdom3 s3_;
    std::string info() const;
public:
    TeeDomain(const VectorXd& dim, const VectorXl& div, double dT);     // 5 dims, 5 divs
    Matrix3Xd checkpoints() const;
};

class TubeDomain : public Domain {
public:
    typedef Parallelepiped<PeriP, Diff, Spec, PeriP, Diff, Inter> Sdom0;
    typedef Parallelepiped<PeriP, Inter, Inter, PeriP, Diff, Diff> Sdom1;
    typedef Parallelepiped<PeriP, Spec, Diff, PeriP, Inter, Diff> Sdom2;
private:
    VectorXd dim_; VectorXl div_; double dT_;
    Sdom0 s0_; Sdom1 s1_; Sdom2 s2_;
    std::string info() const;
public:
    TubeDomain(const VectorXd& dim, const VectorXl& div, double dT);    // 4 dims, 4 divs
    Matrix3Xd checkpoints() const;
};
// Octet-truss unit cell (domain.h:299-385, domain.cpp:576-1280): 42 convex subdomains -- prisms, pyramids, triangular prisms
// and parallelepipeds -- joined by 71 Inter pairs and 5 mirror-periodic pairs; the ten strut subdomains (16..25) carry the
// temperature gradient and the tally grids.  dim = (s, d1, d2, t), div = (div0..div3).  The geometry table (origins and edge
// vectors as linear forms of s, a, t; boundary classes; pair lists) is octet_table.inc, derived from outputs of the reference
// by tools/derive_octet_table.py and pinned against the reference's own objects in tests/test_reference_pin.py.
class OctetDomain : public Domain {
public:
    typedef PeriBoundary<Polygon<4> > Peri4;
private:
    VectorXd dim_; VectorXl div_; double dT_;
    std::vector<std::unique_ptr<EmitSubdomain> > cells_;
    std::string info() const;
public:
    OctetDomain(const VectorXd& dim, const VectorXl& div, double dT);
    Matrix3Xd checkpoints() const;                  // one interior point per subdomain (the reference hand-picks 52)
    ArrayXXd average(const ArrayXXd& data) const;   // domain.cpp:1252-1280: cell-volume weighted mean over the k = 0 layer of the strut grids
    std::vector<double> averageWeights() const;     // the per-column weights of that mean (OctetDomain::WeightF)
};
#endif
