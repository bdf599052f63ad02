// subdomain.h — host mirror of Subdomain / EmitSubdomain / Parallelepiped (subdomain.h:35-198,
// subdomain.cpp:41-71, 108-116, 148-159, 250-281).  advect() and the emission draws run on the device.
#ifndef MCB_HOST_SUBDOMAIN_H
#define MCB_HOST_SUBDOMAIN_H
#include <tuple>
#include <vector>
#include "boundary.h"

class Subdomain {
public:
    typedef std::vector<const Subdomain*> Pointers;
private:
    Boundary::Pointers bdryPtrs_;
    Emitter::Pointers emitPtrs_;
    double vol_;
    Vector3d o_;
    Matrix3d mat_, inv_;
    Vector3l div_, shape_, max_;
    int accum_;
    double eps_;
public:
    Subdomain();
    Subdomain(double vol, const Vector3d& o, const Matrix3d& mat, const Vector3l& div);
    Subdomain(const Subdomain&) = delete;              // boundary/emitter tables hold addresses
    Subdomain& operator=(const Subdomain&) = delete;
    virtual ~Subdomain();

    bool isInit() const;
    bool isInside(const Vector3d& pos) const;
    const Boundary::Pointers& bdryPtrs() const { return bdryPtrs_; }
    const Emitter::Pointers& emitPtrs() const { return emitPtrs_; }

    const Vector3d& origin() const { return o_; }
    const Matrix3d& matrix() const { return mat_; }
    const Matrix3d& inverse() const { return inv_; }
    const Vector3l& div() const { return div_; }
    const Vector3l& shape() const { return shape_; }
    const Vector3l& max() const { return max_; }
    double eps() const { return eps_; }

    int accumFlag() const { return accum_; }
    Vector3d coord(const Vector3d& pos) const;
    Vector3l coord2index(const Vector3d& coord) const;

    double vol() const { return vol_; }
    virtual double cellVol(const Vector3l& index) const = 0;
    virtual int cellKind() const = 0;                   // MCB_CELL_*
    virtual void describe(mcb_sdom_desc& d) const;      // everything but the plane range

protected:
    void addBdry(Boundary* bdry);
    void addBdry(EmitBoundary* bdry);
};

class EmitSubdomain : public Subdomain, public Emitter {
    Vector3d gradT_;
    Matrix3d rot_;
public:
    EmitSubdomain() {}
    EmitSubdomain(double vol, const Vector3d& o, const Matrix3d& mat, const Vector3l& div, const Vector3d& gradT);
    const Subdomain* emitSdom() const { return this; }
    const Boundary* emitBdry() const { return 0; }
    double emitWeight() const { return 2. * vol() * gradT_.norm(); }     // subdomain.cpp:250-253
    const Vector3d& gradT() const { return gradT_; }
    void describe(mcb_sdom_desc& d) const;
};

// Parallelepiped<Bac,Lef,Bot,Fro,Rig,Top>: faces in that order, inward normals (subdomain.h:128-198).
template <typename Bac, typename Lef, typename Bot, typename Fro = Bac, typename Rig = Lef, typename Top = Bot>
class Parallelepiped : public EmitSubdomain {
public:
    typedef std::tuple<Bac, Lef, Bot, Fro, Rig, Top> BdryCont;
private:
    typedef Parallelogram Par;
    BdryCont bdryCont_;
    void init() {
        MC_ASSERT_MSG(vol() >= Dbl::min(), "Volume too small, check vector order");
        addBdry(&std::get<0>(bdryCont_)); addBdry(&std::get<1>(bdryCont_)); addBdry(&std::get<2>(bdryCont_));
        addBdry(&std::get<3>(bdryCont_)); addBdry(&std::get<4>(bdryCont_)); addBdry(&std::get<5>(bdryCont_));
    }
public:
    Parallelepiped(const Vector3d& o, const Matrix3d& mat, const Vector3l& div,
                   const Vector3d& gradT = Vector3d::Zero(), const VectorXd& T = VectorXd(6, 0.))
        : EmitSubdomain(mat.determinant(), o, mat, div, gradT),
          bdryCont_(Bac(o,              Par(mat.col(1), mat.col(2)), T.at(0)),
                    Lef(o,              Par(mat.col(2), mat.col(0)), T.at(1)),
                    Bot(o,              Par(mat.col(0), mat.col(1)), T.at(2)),
                    Fro(o + mat.col(0), Par(mat.col(2), mat.col(1)), T.at(3)),
                    Rig(o + mat.col(1), Par(mat.col(0), mat.col(2)), T.at(4)),
                    Top(o + mat.col(2), Par(mat.col(1), mat.col(0)), T.at(5))) {
        MC_ASSERT_MSG(T.size() == 6, "Incorrect number of temperatures");
        init();
    }
    template <int I>
    typename std::tuple_element<I, BdryCont>::type& bdry() { return std::get<I>(bdryCont_); }
    double cellVol(const Vector3l&) const { return vol() / shape().prod(); }   // subdomain.cpp:269-273
    int cellKind() const { return MCB_CELL_PARALLELEPIPED; }
};
#endif
