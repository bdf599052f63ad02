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
// cell volumes of the simplex cells (subdomain.cpp:283-307, 322-349), formulas kept literally
namespace TriangularPrismImpl { double cellVol(const Vector3l& index, const Vector3l& shape, double vol); }
namespace TetrahedronImpl { double cellVol(const Vector3l& index, const Vector3l& shape, double vol); }

// TriangularPrism<Bac,Lef,Bot,Dia,Top> (subdomain.h:206-281): half of a parallelepiped cut along the diagonal of its base
template <typename Bac, typename Lef, typename Bot, typename Dia, typename Top = Bot>
class TriangularPrism : public EmitSubdomain {
public:
    typedef std::tuple<Bac, Lef, Bot, Dia, Top> BdryCont;
private:
    typedef Parallelogram Par;
    typedef Triangle Tri;
    BdryCont bdryCont_;
public:
    TriangularPrism(const Vector3d& o, const Matrix3d& mat, const Vector3l& div, const Vector3d& gradT = Vector3d::Zero(),
                    const VectorXd& T = VectorXd(5, 0.))
        : EmitSubdomain(mat.determinant() / 2., o, mat, div, gradT),
          bdryCont_(Bac(o,              Par(mat.col(1), mat.col(2)), T.at(0)),
                    Lef(o,              Par(mat.col(2), mat.col(0)), T.at(1)),
                    Bot(o,              Tri(mat.col(0), mat.col(1)), T.at(2)),
                    Dia(o + mat.col(0), Par(mat.col(2), mat.col(1) - mat.col(0)), T.at(3)),
                    Top(o + mat.col(2), Tri(mat.col(1), mat.col(0)), T.at(4))) {
        MC_ASSERT_MSG(T.size() == 5, "Incorrect number of temperatures");
        MC_ASSERT_MSG(vol() >= Dbl::min(), "Volume too small, check vector order");
        addBdry(&std::get<0>(bdryCont_)); addBdry(&std::get<1>(bdryCont_)); addBdry(&std::get<2>(bdryCont_));
        addBdry(&std::get<3>(bdryCont_)); addBdry(&std::get<4>(bdryCont_));
    }
    template <int I> typename std::tuple_element<I, BdryCont>::type& bdry() { return std::get<I>(bdryCont_); }
    double cellVol(const Vector3l& index) const { return TriangularPrismImpl::cellVol(index, shape(), vol()); }
    int cellKind() const { return MCB_CELL_TRIPRISM; }
};

// Tetrahedron<Bac,Lef,Bot,Dia> (subdomain.h:289-353)
template <typename Bac, typename Lef, typename Bot, typename Dia>
class Tetrahedron : public EmitSubdomain {
public:
    typedef std::tuple<Bac, Lef, Bot, Dia> BdryCont;
private:
    typedef Triangle Tri;
    BdryCont bdryCont_;
public:
    Tetrahedron(const Vector3d& o, const Matrix3d& mat, const Vector3l& div, const Vector3d& gradT = Vector3d::Zero(),
                const VectorXd& T = VectorXd(4, 0.))
        : EmitSubdomain(mat.determinant() / 6., o, mat, div, gradT),
          bdryCont_(Bac(o,              Tri(mat.col(1), mat.col(2)), T.at(0)),
                    Lef(o,              Tri(mat.col(2), mat.col(0)), T.at(1)),
                    Bot(o,              Tri(mat.col(0), mat.col(1)), T.at(2)),
                    Dia(o + mat.col(0), Tri(mat.col(2) - mat.col(0), mat.col(1) - mat.col(0)), T.at(3))) {
        MC_ASSERT_MSG(T.size() == 4, "Incorrect number of temperatures");
        MC_ASSERT_MSG(vol() >= Dbl::min(), "Volume too small, check vector order");
        addBdry(&std::get<0>(bdryCont_)); addBdry(&std::get<1>(bdryCont_)); addBdry(&std::get<2>(bdryCont_)); addBdry(&std::get<3>(bdryCont_));
    }
    template <int I> typename std::tuple_element<I, BdryCont>::type& bdry() { return std::get<I>(bdryCont_); }
    double cellVol(const Vector3l& index) const { return TetrahedronImpl::cellVol(index, shape(), vol()); }
    int cellKind() const { return MCB_CELL_TETRAHEDRON; }
};

// Prism<Bot, Top, std::tuple<Sides...>> (subdomain.h:363-470) and Pyramid<Bot, std::tuple<Sides...>> (:486-596).
// mat columns: 0 = axis (prism) / apex (pyramid) vector, 1..N-1 = fan of the N-gon base; single cell or untallied.
namespace PrismImpl {
Matrix3d matBase(const std::vector<Vector3d>& mat);                    // (col 1, col N-1, col 0)
std::vector<double> volume(const std::vector<Vector3d>& mat, double div);   // (col i x col i+1) . col 0 / div, i < N-2 (sic)
}

template <typename Bot, typename Top, typename Sid> class Prism;
template <typename Bot, typename Top, typename... S>
class Prism<Bot, Top, std::tuple<S...>> : public EmitSubdomain {
    static const int N = (int)sizeof...(S);
    Bot bot_; Top top_; std::tuple<S...> sides_;
    std::vector<Vector3d> mat_;
    template <typename Side> static Side side(const Vector3d& o, const std::vector<Vector3d>& m, const VectorXd& T, int ind) {
        Vector3d p; if (ind != 0) p = p + m[(size_t)ind];
        Vector3d j = -p; if (ind != N - 1) j = j + m[(size_t)ind + 1];
        return Side(o + p, Parallelogram(m[0], j), T.at((size_t)ind + 2));
    }
    template <size_t... I> static std::tuple<S...> sides(const Vector3d& o, const std::vector<Vector3d>& m, const VectorXd& T, std::index_sequence<I...>) {
        return std::tuple<S...>(side<S>(o, m, T, (int)I)...);
    }
    template <size_t... I> void addSides(std::index_sequence<I...>) { (addBdry(&std::get<I>(sides_)), ...); }
    static double total(const std::vector<double>& v) { double s = 0.; for (double x : v) s += x; return s; }
public:
    Prism(const Vector3d& o, const std::vector<Vector3d>& mat, long div, const Vector3d& gradT = Vector3d::Zero(),
          const VectorXd& T = VectorXd((size_t)N + 2, 0.))
        : EmitSubdomain(total(PrismImpl::volume(mat, 2.)), o, PrismImpl::matBase(mat), Vector3l(div < 0 ? -1 : 0, div < 0 ? -1 : 0, div < 0 ? -1 : 0), gradT),
          bot_(o, Polygon<N>(std::vector<Vector3d>(mat.begin() + 1, mat.end())), T.at(0)),
          top_(o + mat.at(0), Polygon<N>(std::vector<Vector3d>(mat.rbegin(), mat.rend() - 1)), T.at(1)),
          sides_(sides(o, mat, T, std::index_sequence_for<S...>())), mat_(mat) {
        MC_ASSERT_MSG((int)mat.size() == N, "Incorrect number of matrix columns");
        MC_ASSERT_MSG((int)T.size() == N + 2, "Incorrect number of temperatures");
        MC_ASSERT_MSG(vol() >= Dbl::min(), "Volume too small");
        addBdry(&bot_); addBdry(&top_); addSides(std::index_sequence_for<S...>());
    }
    Bot& bottom() { return bot_; }
    Top& top() { return top_; }
    template <int I> typename std::tuple_element<I, std::tuple<S...>>::type& sideBdry() { return std::get<I>(sides_); }
    double cellVol(const Vector3l&) const { return vol(); }                 // subdomain.cpp:396-399
    int cellKind() const { return MCB_CELL_PRISM; }
    void describe(mcb_sdom_desc& d) const {
        EmitSubdomain::describe(d);
        d.nbase = N;
        for (int v = 0; v < N; ++v) for (int k = 0; k < 3; ++k) d.base[3 * v + k] = mat_[(size_t)v](k);
    }
};

template <typename Bot, typename Sid> class Pyramid;
template <typename Bot, typename... S>
class Pyramid<Bot, std::tuple<S...>> : public EmitSubdomain {
    static const int N = (int)sizeof...(S);
    Bot bot_; std::tuple<S...> sides_;
    std::vector<Vector3d> mat_;
    template <typename Side> static Side side(const Vector3d& o, const std::vector<Vector3d>& m, const VectorXd& T, int ind) {
        Vector3d p; if (ind != 0) p = p + m[(size_t)ind];
        Vector3d i = m[0]; if (ind != 0) i = i - m[(size_t)ind];
        Vector3d j = -p; if (ind != N - 1) j = j + m[(size_t)ind + 1];
        return Side(o + p, Triangle(i, j), T.at((size_t)ind + 1));
    }
    template <size_t... I> static std::tuple<S...> sides(const Vector3d& o, const std::vector<Vector3d>& m, const VectorXd& T, std::index_sequence<I...>) {
        return std::tuple<S...>(side<S>(o, m, T, (int)I)...);
    }
    template <size_t... I> void addSides(std::index_sequence<I...>) { (addBdry(&std::get<I>(sides_)), ...); }
    static double total(const std::vector<double>& v) { double s = 0.; for (double x : v) s += x; return s; }
public:
    Pyramid(const Vector3d& o, const std::vector<Vector3d>& mat, long div, const Vector3d& gradT = Vector3d::Zero(),
            const VectorXd& T = VectorXd((size_t)N + 1, 0.))
        : EmitSubdomain(total(PrismImpl::volume(mat, 6.)), o, PrismImpl::matBase(mat), Vector3l(div < 0 ? -1 : 0, div < 0 ? -1 : 0, div < 0 ? -1 : 0), gradT),
          bot_(o, Polygon<N>(std::vector<Vector3d>(mat.begin() + 1, mat.end())), T.at(0)),
          sides_(sides(o, mat, T, std::index_sequence_for<S...>())), mat_(mat) {
        MC_ASSERT_MSG((int)mat.size() == N, "Incorrect number of matrix columns");
        MC_ASSERT_MSG((int)T.size() == N + 1, "Incorrect number of temperatures");
        MC_ASSERT_MSG(vol() >= Dbl::min(), "Volume too small");
        addBdry(&bot_); addSides(std::index_sequence_for<S...>());
    }
    Bot& bottom() { return bot_; }
    template <int I> typename std::tuple_element<I, std::tuple<S...>>::type& sideBdry() { return std::get<I>(sides_); }
    double cellVol(const Vector3l&) const { return vol(); }                 // subdomain.cpp:428-431
    int cellKind() const { return MCB_CELL_PYRAMID; }
    void describe(mcb_sdom_desc& d) const {
        EmitSubdomain::describe(d);
        d.nbase = N;
        for (int v = 0; v < N; ++v) for (int k = 0; k < 3; ++k) d.base[3 * v + k] = mat_[(size_t)v](k);
    }
};
#endif
