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
#include "problem.h"
#include "domain.h"
#include "field.h"
#include "material.h"
#include "random.h"
#include <Eigen/Core>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef Eigen::Matrix<double, Eigen::Dynamic, 1> VecXd;
typedef Eigen::Matrix<long, Eigen::Dynamic, 1> VecXl;

namespace {

// slab between two isothermal walls at +-dT/2 (x faces), specular sides; no temperature gradient in the volume
class SlabDomain : public Domain {
public:
    typedef Parallelepiped<IsotBoundary<Parallelogram>, Spec, Spec> Sdom;
private:
    Sdom sdom_;
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
private:
    Sdom sdom_;
    std::string info() const { return "WireDomain"; }
public:
    WireDomain(const Vector3d& dim, const Vector3l& div, double dT)
        : Domain(), sdom_(Vector3d::Zero(), dim.asDiagonal(), div, Vector3d(-dT / dim(0), 0., 0.)) {
        makePair(sdom_.bdry<0>(), sdom_.bdry<3>(), Vector3d(dim(0), 0., 0.));
        addSdom(&sdom_);
    }
    Matrix3Xd checkpoints() const { return Matrix3Xd(); }
};

void die(const char* msg) { std::fprintf(stderr, "ref_driver: %s\n", msg); std::exit(2); }

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
    if (domStr == "bulk") dom = new BulkDomain(dim, div, dT);
    else if (domStr == "film") dom = new FilmDomain(dim, div, dT);
    else if (domStr == "jct") dom = new JctDomain(dim, div, dT);
    else if (domStr == "tee") dom = new TeeDomain(dim, div, dT);
    else if (domStr == "tube") dom = new TubeDomain(dim, div, dT);
    else if (domStr == "octet") dom = new OctetDomain(dim, div, dT);
    else if (domStr == "slab") dom = new SlabDomain(dim, div, dT);
    else if (domStr == "wire") dom = new WireDomain(dim, div, dT);
    else die("invalid domain");

    const FieldProblem* prob = 0;
    if (probStr == "temp") prob = new TempProblem(mat, dom, nemit, maxscat, maxloop);
    else if (probStr == "flux") prob = new FluxProblem(mat, dom, nemit, maxscat, maxloop);
    else if (probStr == "multi") prob = new MultiProblem(mat, dom, nemit, maxscat, maxloop);
    else if (probStr == "cumtemp") prob = new CumTempProblem(mat, dom, nemit, size, maxscat, maxloop);
    else if (probStr == "cumflux") prob = new CumFluxProblem(mat, dom, nemit, size, maxscat, maxloop);
    else die("invalid problem");

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
