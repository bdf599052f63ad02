// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, not product code.
//
// A driver that LINKS THE REFERENCE'S OWN OBJECTS (everything oracle/Makefile compiles from /root/reference/montecarlo except
// main.o) and runs FieldProblem::solve the way main.cpp:141-181 does (one Rng per OpenMP thread, partial fields summed).
// It exists for two things the reference's main() cannot do:
//   * workloads C2 / C3 of BASELINE.json need domains the reference does not ship: a slab between isothermal walls and a
//     wire with diffuse side walls.  Both are composed here from the reference's own templates
//     (Parallelepiped<IsotBoundary<Parallelogram>, Spec, Spec>, Parallelepiped<PeriP, Diff, Diff>) exactly like BulkDomain /
//     FilmDomain compose theirs (domain.h:77-117, domain.cpp:97-195) -- what a user of the reference would write;
//   * it reports the number of loop trips (phonon-steps) the reference executed and the wall time of the solve alone, and
//     prints the solution with 17 significant digits.
// Seeds are deterministic: thread t uses mt19937(seed + t), the reference's -DDEBUG scheme (main.cpp:29-39).
//
// usage: ref_driver <matdir> <disp> <relax> <T> <domain> <ndim> <dim...> <ndiv> <div...> <dT>
//                   <problem> <nemit> <size> <maxscat> <maxloop> <seed>
//
// Third use: `flatten <file>` as the last two arguments writes the const state of the reference's own Domain / FieldProblem
// objects as the C-ABI descriptors of include/mcb.h (mcb_sdom_desc, mcb_plane_desc, pair lists, mcb_emitter_desc, cell
// volumes, mcb_problem_desc) instead of solving -- the reference-side half of the binding INTEGRATION.md describes, compiled
// against the real classes.  tests/ feed the file to mcb_upload_domain / mcb_solve, which is how the reference's OctetDomain
// (not restated anywhere in this repo) runs on the GPU.  The binding reads private members; a maintainer would add the
// getters listed in INTEGRATION.md, this test harness uses `#define private public` around the reference's headers instead
// (access specifiers do not change the Itanium-ABI layout, and the objects themselves are compiled normally).
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <boost/fusion/algorithm/iteration.hpp>
#include <boost/optional.hpp>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <typeinfo>
#include <vector>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#define private public
#define protected public
#include "problem.h"
#include "domain.h"
#include "field.h"
#include "material.h"
#include "random.h"
#undef private
#undef protected
#include "../include/mcb.h"

typedef Eigen::Matrix<double, Eigen::Dynamic, 1> VecXd;
typedef Eigen::Matrix<long, Eigen::Dynamic, 1> VecXl;

namespace {

// slab between two isothermal walls at +-dT/2 (x faces), specular sides; no temperature gradient in the volume
class SlabDomain : public Domain {
public:
    typedef Parallelepiped<IsotBoundary<Parallelogram>, Spec, Spec> Sdom;
public:
    Sdom sdom_;
private:
    std::string info() const { return "SlabDomain"; }
    static VectorXd temps(double dT) { VectorXd T = VectorXd::Zero(6); T(0) = dT / 2.; T(3) = -dT / 2.; return T; }
public:
    SlabDomain(const Vector3d& dim, const Vector3l& div, double dT)
        : Domain(), sdom_(Vector3d::Zero(), dim.asDiagonal(), div, Vector3d::Zero(), temps(dT)) {
        addSdom(&sdom_);
    }
    Matrix3Xd checkpoints() const { return Matrix3Xd(); }
};

// wire periodic along x (temperature gradient -dT/L like BulkDomain), four diffuse side walls
class WireDomain : public Domain {
public:
    typedef Parallelepiped<PeriP, Diff, Diff> Sdom;
public:
    Sdom sdom_;
private:
    std::string info() const { return "WireDomain"; }
public:
    WireDomain(const Vector3d& dim, const Vector3l& div, double dT)
        : Domain(), sdom_(Vector3d::Zero(), dim.asDiagonal(), div, Vector3d(-dT / dim(0), 0., 0.)) {
        makePair(sdom_.bdry<0>(), sdom_.bdry<3>(), Vector3d(dim(0), 0., 0.));
        addSdom(&sdom_);
    }
    Matrix3Xd checkpoints() const { return Matrix3Xd(); }
};

// one non-box cell of the reference's own templates as a domain (the reference only uses these inside OctetDomain):
// pins the tri-prism / tetrahedron / prism emission folds, cellVol and Triangle / Polygon emitters of N3
template <class Cell> class OneCellDomain : public Domain {
public:
    Cell sdom_;
private:
    std::string info() const { return "OneCellDomain"; }
public:
    explicit OneCellDomain(const Cell& c) : Domain(), sdom_(c) { addSdom(&sdom_); }      // Cell's copy constructor re-registers its boundaries
    Matrix3Xd checkpoints() const { return Matrix3Xd(); }
};
typedef TriangularPrism<DiffBoundary, SpecBoundary, DiffBoundary, SpecBoundary, DiffBoundary> TriPrismCell;
typedef Tetrahedron<IsotBoundary<Triangle>, SpecBoundary, DiffBoundary, DiffBoundary> TetCell;
typedef Prism<IsotBoundary<Polygon<5> >, IsotBoundary<Polygon<5> >,
              boost::fusion::vector5<DiffBoundary, SpecBoundary, DiffBoundary, SpecBoundary, DiffBoundary> > Prism5Cell;

void die(const char* msg) { std::fprintf(stderr, "ref_driver: %s\n", msg); std::exit(2); }

// ---------------------------------------------------------------------------------------------- flatten
void put3(double* d, const Vector3d& v) { for (int k = 0; k < 3; ++k) d[k] = v(k); }
void put9(double* d, const Matrix3d& m) { for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) d[3 * c + r] = m(r, c); }   // column-major
Matrix3d rotTo(const Vector3d& n) { return Eigen::Quaternion<double>::FromTwoVectors(Vector3d::UnitZ(), n).matrix(); }     // boundary.cpp:33-37

// cell kind and (prism / pyramid) base columns need the concrete template type: visit the domain's own containers
struct CellInfo { int cell; int nbase; double base[3 * MCB_MAX_BASE]; };
struct CellVisitor {
    std::vector<CellInfo>* out;
    void push(int cell, const Matrix3Xd* mat) const {
        CellInfo c; std::memset(&c, 0, sizeof c); c.cell = cell;
        if (mat) {
            if (mat->cols() > MCB_MAX_BASE) die("too many base columns");
            c.nbase = (int)mat->cols();
            for (long j = 0; j < mat->cols(); ++j) for (int k = 0; k < 3; ++k) c.base[3 * j + k] = (*mat)(k, j);
        }
        out->push_back(c);
    }
    template <class A, class B, class C, class D, class E, class F> void operator()(const Parallelepiped<A, B, C, D, E, F>&) const { push(MCB_CELL_PARALLELEPIPED, 0); }
    template <class A, class B, class C, class D, class E> void operator()(const TriangularPrism<A, B, C, D, E>&) const { push(MCB_CELL_TRIPRISM, 0); }
    template <class A, class B, class C, class D> void operator()(const Tetrahedron<A, B, C, D>&) const { push(MCB_CELL_TETRAHEDRON, 0); }
    template <class B, class T, class S> void operator()(const Prism<B, T, S>& p) const { push(MCB_CELL_PRISM, &p.mat_); }
    template <class B, class S> void operator()(const Pyramid<B, S>& p) const { push(MCB_CELL_PYRAMID, &p.mat_); }
};

template <class S> bool periOf(const Boundary* b, mcb_plane_desc& d, const Boundary** pair) {
    const PeriBoundary<S>* p = dynamic_cast<const PeriBoundary<S>*>(b);
    if (!p) return false;
    put9(d.peri_rot, p->rot_); put3(d.peri_transl, p->transl_); *pair = p->pair_;
    return true;
}
void shapeOf(const Boundary::Shape& sh, mcb_plane_desc& d) {
    if (const Parallelogram* q = dynamic_cast<const Parallelogram*>(&sh)) { d.shape = MCB_SHAPE_PARALLELOGRAM; d.nvert = 2; put3(d.verts, q->i_); put3(d.verts + 3, q->j_); return; }
    if (const Triangle* q = dynamic_cast<const Triangle*>(&sh)) { d.shape = MCB_SHAPE_TRIANGLE; d.nvert = 2; put3(d.verts, q->i_); put3(d.verts + 3, q->j_); return; }
    const Matrix3Xd* v = 0;
    if (const Polygon<4>* q = dynamic_cast<const Polygon<4>*>(&sh)) v = &q->verts_;
    else if (const Polygon<5>* q = dynamic_cast<const Polygon<5>*>(&sh)) v = &q->verts_;
    else if (const Polygon<6>* q = dynamic_cast<const Polygon<6>*>(&sh)) v = &q->verts_;
    else if (const Polygon<7>* q = dynamic_cast<const Polygon<7>*>(&sh)) v = &q->verts_;
    else if (const Polygon<8>* q = dynamic_cast<const Polygon<8>*>(&sh)) v = &q->verts_;
    else if (const Polygon<9>* q = dynamic_cast<const Polygon<9>*>(&sh)) v = &q->verts_;
    if (!v) die("unknown boundary shape");
    d.shape = MCB_SHAPE_POLYGON; d.nvert = (int32_t)v->cols();
    for (long j = 0; j < v->cols(); ++j) for (int k = 0; k < 3; ++k) d.verts[3 * j + k] = (*v)(k, j);
}

int flatten(const Domain* dom, const std::vector<CellInfo>& cells, const FieldProblem* prob, int probKind, long size, const char* path) {
    const Subdomain::Pointers& sp = dom->sdomPtrs();
    if (cells.size() != sp.size()) die("cell visitor / sdomPtrs mismatch");
    std::vector<mcb_sdom_desc> sdoms; std::vector<mcb_plane_desc> planes; std::vector<int32_t> pairs;
    std::vector<mcb_emitter_desc> emitters; std::vector<double> cellVol;
    std::map<const Boundary*, int32_t> planeId; std::map<const Subdomain*, int32_t> sdomId;
    std::vector<std::vector<const Boundary*> > partners;
    for (size_t s = 0; s < sp.size(); ++s) {
        const Subdomain* sd = sp[s]; sdomId[sd] = (int32_t)s;
        mcb_sdom_desc d; std::memset(&d, 0, sizeof d);
        put3(d.origin, sd->o_); put9(d.mat, sd->mat_); put9(d.inv, sd->inv_);
        for (int k = 0; k < 3; ++k) { d.div[k] = sd->div_(k); d.shape[k] = sd->shape_(k); d.max[k] = sd->max_(k); }
        d.accum = sd->accum_; d.cell = cells[s].cell; d.eps = sd->eps_; d.vol = sd->vol_;
        d.nbase = cells[s].nbase; std::memcpy(d.base, cells[s].base, sizeof d.base);
        const EmitSubdomain* es = dynamic_cast<const EmitSubdomain*>(sd);
        put9(d.emit_rot, Matrix3d::Identity());
        if (es) {
            put3(d.grad_t, es->gradT_);
            if (es->gradT_.norm() > 0.) put9(d.emit_rot, es->rot_);            // rotMatrix(0/0) is NaN in the reference and never used
        }
        d.plane_begin = (int32_t)planes.size(); d.plane_count = (int32_t)sd->bdryPtrs().size();
        for (size_t b = 0; b < sd->bdryPtrs().size(); ++b) {
            const Boundary* bd = sd->bdryPtrs()[b];
            planeId[bd] = (int32_t)planes.size();
            mcb_plane_desc p; std::memset(&p, 0, sizeof p);
            put3(p.normal, bd->normal()); p.offset = bd->offset(); p.sdom = (int32_t)s; p.shape = MCB_SHAPE_NONE;
            put9(p.rot, rotTo(bd->normal()));
            std::vector<const Boundary*> prt;
            const std::string ty = bd->type();
            if (ty == "Spec") p.kind = MCB_BDRY_SPEC;
            else if (ty == "Diff") { p.kind = MCB_BDRY_DIFF; put9(p.rot, dynamic_cast<const DiffBoundary*>(bd)->rot_); }
            else if (ty == "Inter") { p.kind = MCB_BDRY_INTER; const InterBoundary* ib = dynamic_cast<const InterBoundary*>(bd); prt.assign(ib->pairs_.begin(), ib->pairs_.end()); }
            else {
                const EmitBoundary* eb = dynamic_cast<const EmitBoundary*>(bd);
                if (!eb) die("unknown boundary type");
                p.kind = ty.compare(0, 4, "Isot") == 0 ? MCB_BDRY_ISOT : MCB_BDRY_PERI;
                put9(p.rot, eb->rot_); p.T = eb->T_; put3(p.origin, eb->o_);
                shapeOf(eb->shape(), p);
                if (p.kind == MCB_BDRY_PERI) {
                    const Boundary* pr = 0;
                    if (!(periOf<Parallelogram>(bd, p, &pr) || periOf<Triangle>(bd, p, &pr) || periOf<Polygon<4> >(bd, p, &pr) ||
                          periOf<Polygon<5> >(bd, p, &pr) || periOf<Polygon<6> >(bd, p, &pr) || periOf<Polygon<7> >(bd, p, &pr) ||
                          periOf<Polygon<8> >(bd, p, &pr) || periOf<Polygon<9> >(bd, p, &pr))) die("unknown periodic boundary");
                    if (pr) prt.push_back(pr);
                }
            }
            partners.push_back(prt);
            planes.push_back(p);
        }
        sdoms.push_back(d);
        const Vector3l shp = sd->shape();                                   // Field(rows, dom, fun) nesting: k, j, i (field.cpp:62-78)
        for (long k = 0; k < shp(2); ++k) for (long j = 0; j < shp(1); ++j) for (long i = 0; i < shp(0); ++i)
            cellVol.push_back(sd->cellVol(Vector3l(i, j, k)));
    }
    for (size_t q = 0; q < planes.size(); ++q) {
        planes[q].pair_begin = (int32_t)pairs.size(); planes[q].pair_count = (int32_t)partners[q].size();
        for (size_t k = 0; k < partners[q].size(); ++k) {
            if (!planeId.count(partners[q][k])) die("boundary paired with a boundary outside the domain");
            pairs.push_back(planeId[partners[q][k]]);
        }
    }
    for (size_t e = 0; e < dom->emitPtrs().size(); ++e) {
        const Emitter* em = dom->emitPtrs()[e];
        mcb_emitter_desc d; std::memset(&d, 0, sizeof d);
        if (em->emitBdry()) { d.kind = MCB_EMIT_BDRY; d.index = planeId.at(em->emitBdry()); }
        else { d.kind = MCB_EMIT_SDOM; d.index = sdomId.at(em->emitSdom()); }
        d.weight = em->emitWeight();
        emitters.push_back(d);
    }
    mcb_problem_desc pd; std::memset(&pd, 0, sizeof pd);
    pd.kind = probKind; pd.rows = (int32_t)prob->rows(); pd.size = size;
    pd.step = 0;
    if (const CumTempProblem* c = dynamic_cast<const CumTempProblem*>(prob)) pd.step = c->step_;
    if (const CumFluxProblem* c = dynamic_cast<const CumFluxProblem*>(prob)) pd.step = c->step_;
    pd.nemit = prob->nemit_; pd.maxscat = prob->maxscat_; pd.maxloop = prob->maxloop_; pd.power = prob->power_;
    std::vector<int64_t> counts(prob->emitPdf_.data(), prob->emitPdf_.data() + prob->emitPdf_.size());

    std::FILE* f = std::fopen(path, "wb");
    if (!f) die("cannot open the output file");
    const int32_t hdr[8] = {0x4642434d /* "MCBF" */, (int32_t)sdoms.size(), (int32_t)planes.size(), (int32_t)pairs.size(),
                            (int32_t)emitters.size(), (int32_t)cellVol.size(), (int32_t)sizeof(mcb_sdom_desc), (int32_t)sizeof(mcb_plane_desc)};
    std::fwrite(hdr, sizeof hdr, 1, f);
    std::fwrite(sdoms.data(), sizeof(mcb_sdom_desc), sdoms.size(), f);
    std::fwrite(planes.data(), sizeof(mcb_plane_desc), planes.size(), f);
    std::fwrite(pairs.data(), sizeof(int32_t), pairs.size(), f);
    std::fwrite(emitters.data(), sizeof(mcb_emitter_desc), emitters.size(), f);
    std::fwrite(cellVol.data(), sizeof(double), cellVol.size(), f);
    std::fwrite(&pd, sizeof pd, 1, f);
    std::fwrite(counts.data(), sizeof(int64_t), counts.size(), f);
    // Domain::average (domain.cpp:78-81, 1252-1280): the per-column weights of OctetDomain::WeightF (none for the other domains,
    // whose average() is the identity)
    std::vector<double> weights;
    if (const OctetDomain* oct = dynamic_cast<const OctetDomain*>(dom)) {
        Eigen::Array<double, 1, Eigen::Dynamic> w = Field(1, oct, OctetDomain::WeightF(*oct)).data().row(0);
        for (long j = 0; j < w.cols(); ++j) weights.push_back(w(0, j));
    }
    const int32_t nw = (int32_t)weights.size();
    std::fwrite(&nw, sizeof nw, 1, f);
    std::fwrite(weights.data(), sizeof(double), weights.size(), f);
    std::fclose(f);
    return 0;
}

} // namespace

int main(int argc, char** argv) {
    int a = 1;
    auto next = [&]() -> const char* { if (a >= argc) die("too few arguments"); return argv[a++]; };
    const std::string matdir = next(), disp = next(), relax = next();
    const double T = std::atof(next());
    const std::string domStr = next();
    const int ndim = std::atoi(next());
    VecXd dim(ndim); for (int i = 0; i < ndim; ++i) dim(i) = std::atof(next());
    const int ndiv = std::atoi(next());
    VecXl div(ndiv); for (int i = 0; i < ndiv; ++i) div(i) = std::atol(next());
    const double dT = std::atof(next());
    const std::string probStr = next();
    const long nemit = std::atol(next()), size = std::atol(next()), maxscat = std::atol(next()), maxloop = std::atol(next());
    const unsigned long seed = std::strtoul(next(), 0, 10);

    if (chdir(matdir.c_str()) != 0) die("invalid material directory");
    const Material* mat = new Material(disp, relax, T);

    const Domain* dom = 0;
    std::vector<CellInfo> cells;                       // cell kinds in sdomPtrs() order (for `flatten`)
    CellVisitor vis = {&cells};
    if (domStr == "bulk") { BulkDomain* d = new BulkDomain(dim, div, dT); vis(d->sdom_); dom = d; }
    else if (domStr == "film") { FilmDomain* d = new FilmDomain(dim, div, dT); vis(d->sdom_); dom = d; }
    else if (domStr == "jct") { JctDomain* d = new JctDomain(dim, div, dT); boost::fusion::for_each(d->sdomCont_, vis); dom = d; }
    else if (domStr == "tee") { TeeDomain* d = new TeeDomain(dim, div, dT); boost::fusion::for_each(d->sdomCont_, vis); dom = d; }
    else if (domStr == "tube") { TubeDomain* d = new TubeDomain(dim, div, dT); boost::fusion::for_each(d->sdomCont_, vis); dom = d; }
    else if (domStr == "octet") { OctetDomain* d = new OctetDomain(dim, div, dT); boost::fusion::for_each(d->sdomCont_, vis); dom = d; }
    else if (domStr == "slab") { SlabDomain* d = new SlabDomain(dim, div, dT); vis(d->sdom_); dom = d; }
    else if (domStr == "wire") { WireDomain* d = new WireDomain(dim, div, dT); vis(d->sdom_); dom = d; }
    // HexDomain(dim, dT) / PyrDomain(dim, dT) never call their own init() in the reference (the domain is unusable as shipped:
    // "Domain setup not complete"); calling it is the one-line fix a user would make (domain.cpp:224-336)
    else if (domStr == "hex") { HexDomain* d = new HexDomain(dim, dT); d->init(); vis(d->sdom_); dom = d; }
    else if (domStr == "pyr") { PyrDomain* d = new PyrDomain(dim, dT); d->init(); vis(d->sdom_); dom = d; }
    else if (domStr == "triprism" || domStr == "tet" || domStr == "prism5") {
        // dim = origin(3), edge columns (3 each), gradT(3); div as the cell's constructor takes it; dT = wall temperature scale
        const int ncol = domStr == "prism5" ? 5 : 3;
        if (ndim != 3 + 3 * ncol + 3) die("non-box cell: dim must hold origin, columns and gradT");
        const Vector3d o(dim(0), dim(1), dim(2)), gradT(dim(3 + 3 * ncol), dim(4 + 3 * ncol), dim(5 + 3 * ncol));
        Matrix3Xd cols(3, ncol);
        for (int j = 0; j < ncol; ++j) for (int k = 0; k < 3; ++k) cols(k, j) = dim(3 + 3 * j + k);
        if (domStr == "triprism") {
            OneCellDomain<TriPrismCell>* d = new OneCellDomain<TriPrismCell>(TriPrismCell(o, Matrix3d(cols), Vector3l(div(0), div(1), div(2)), gradT));
            vis(d->sdom_); dom = d;
        } else if (domStr == "tet") {
            VectorXd T = VectorXd::Zero(4); T(0) = dT;
            OneCellDomain<TetCell>* d = new OneCellDomain<TetCell>(TetCell(o, Matrix3d(cols), Vector3l(div(0), div(1), div(2)), gradT, T));
            vis(d->sdom_); dom = d;
        } else {
            VectorXd T = VectorXd::Zero(7); T(0) = dT / 2.; T(1) = -dT / 2.;
            OneCellDomain<Prism5Cell>* d = new OneCellDomain<Prism5Cell>(Prism5Cell(o, cols, div(0), gradT, T));
            vis(d->sdom_); dom = d;
        }
    }
    else die("invalid domain");

    const FieldProblem* prob = 0;
    if (probStr == "temp") prob = new TempProblem(mat, dom, nemit, maxscat, maxloop);
    else if (probStr == "flux") prob = new FluxProblem(mat, dom, nemit, maxscat, maxloop);
    else if (probStr == "multi") prob = new MultiProblem(mat, dom, nemit, maxscat, maxloop);
    else if (probStr == "cumtemp") prob = new CumTempProblem(mat, dom, nemit, size, maxscat, maxloop);
    else if (probStr == "cumflux") prob = new CumFluxProblem(mat, dom, nemit, size, maxscat, maxloop);
    else die("invalid problem");

    if (a + 1 < argc && std::string(argv[a]) == "flatten") {
        const int kind = probStr == "temp" ? MCB_PROB_TEMP : probStr == "flux" ? MCB_PROB_FLUX : probStr == "multi" ? MCB_PROB_MULTI
                       : probStr == "cumtemp" ? MCB_PROB_CUMTEMP : MCB_PROB_CUMFLUX;
        return flatten(dom, cells, prob, kind, (kind == MCB_PROB_CUMTEMP || kind == MCB_PROB_CUMFLUX) ? size : 0, argv[a + 1]);
    }

    // main.cpp:141-167 solveField: per-thread generator, orphaned `omp for` inside FieldProblem::solve, partials summed
    ArrayXXd sol = prob->initSolution();
    Progress prog = prob->initProgress();
    unsigned long long steps = 0;
    int threads = 1;
    std::FILE* keep = stdout;
    // Progress::incrCount prints its bar to std::cout from inside the timed region; silence it like `> /dev/null` would
    std::cout.setstate(std::ios_base::failbit);
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel reduction(+ : steps)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#pragma omp single
        threads = omp_get_num_threads();
#endif
        Rng gen((Rng::result_type)(seed + (unsigned long)tid));
        const unsigned long long before = Eigen::shim::loop_trips;
        ArrayXXd partial = prob->solve(gen, &prog);
        steps += Eigen::shim::loop_trips - before;
#pragma omp critical
        {
            sol += partial;
        }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout.clear();

    std::fprintf(keep, "threads %d\nsteps %llu\nseconds %.6f\nesc %ld\nrows %ld\ncols %ld\nOutput\n", threads, steps, secs,
                 prog.esc(), (long)sol.rows(), (long)sol.cols());
    for (long i = 0; i < sol.rows(); ++i) {
        for (long j = 0; j < sol.cols(); ++j) std::fprintf(keep, "%.17g ", sol(i, j));
        std::fprintf(keep, "\n");
    }
    delete prob; delete dom; delete mat;
    return 0;
}
