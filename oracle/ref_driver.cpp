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
#include "ref_flatten.h"

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

using namespace refflat;

int flatten(const Domain* dom, const std::vector<CellInfo>& cells, const FieldProblem* prob, int probKind, long size, const char* path) {
    FlatDomain fd; flattenDomain(dom, cells, fd);
    const std::vector<mcb_sdom_desc>& sdoms = fd.sdoms; const std::vector<mcb_plane_desc>& planes = fd.planes; const std::vector<int32_t>& pairs = fd.pairs;
    const std::vector<mcb_emitter_desc>& emitters = fd.emitters; const std::vector<double>& cellVol = fd.cellVol;
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
