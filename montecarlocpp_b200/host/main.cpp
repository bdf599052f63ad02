// main.cpp — command-line driver with the reference's positional grammar and stdout blocks
// (main.cpp:67-84 printSolution, :146-210 solveField / calcStats, :216-479 argument grammar).
//
//   <dir> <material> [T] <domain> <dims...> <divs...> <problem> <nemit> [size] <maxscat> <maxloop> <nsim>
//
//   grey|silicon [T] | custom [T] disp relax
//   bulk dim div0 | film dim0 dim1 div1 | hex dim0 dim1 | pyr dim0 dim1 | jct dim0 dim3 | tee dim0 dim4 div0 | tube dim0 dim1 dim3 div1 div3
//   octet dim0 dim1 dim2 dim3 div0 div1 div2 div3 dT
//   slab dim0 dim1 div0 dT | wire dim0 dim1 div1        (not in the reference; see domain.h)
//   temp|flux|multi nemit maxscat maxloop nsim | cumtemp|cumflux nemit size maxscat maxloop nsim
//   check r00 r01 r02 r10 r11 r12 r20 r21 r22 | traj px py pz dx dy dz maxscat maxloop
//
// Differences from the reference driver: FieldProblem::solve runs on the GPU(s).  It is still called the reference's way --
// from every thread of an OpenMP region, partial fields summed (main.cpp:155-166) -- and sums to exactly one solve (thread 0
// brings it, see problem.cpp).  The environment variable MCB_DEVICES selects the GPUs: "all" (default: every visible
// sm_100 device; the particle range is sharded over them and the tallies summed with NCCL) or a comma-separated list.
#include <unistd.h>
#include <cstdlib>
#include <memory>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <vector>
#include "domain.h"
#include "field.h"
#include "material.h"
#include "problem.h"
#include "random.h"

typedef Dev::result_type Seed;

#ifdef DEBUG
static Seed getSeed() { static Seed s(0); Seed r;
#pragma omp critical
    { r = s++; }
    return r; }
#else
static Seed getSeed() { static Dev urandom; Seed r;
#pragma omp critical(mc_seed)
    { r = urandom(); }
    return r; }
#endif
static void printSeed(Seed s) {                                                   // main.cpp:50-65
#pragma omp single
    { std::cout << "  seeds: "; }
#pragma omp critical
    { std::cout << s << ' '; }
#pragma omp barrier
#pragma omp single
    { std::cout << std::endl; }
}

static void printSolution(const ArrayXXd& sol) {
    std::ios_base::fmtflags mask = std::cout.flags();
    std::cout << std::scientific << std::setprecision(9);
    for (long i = 0; i < sol.rows(); ++i) {
        for (long j = 0; j < sol.cols(); ++j) std::cout << std::setw(16) << sol(i, j) << ' ';
        std::cout << std::endl;
    }
    std::cout << std::endl;
    std::cout.flags(mask);
}

static ArrayXXd solveTraj(const TrajProblem* prob, const Clock& clk) {          // main.cpp:86-102
    static long n = 0;
    std::cout << "Trajectory " << n++ << std::endl;
    Progress prog = prob->initProgress();
    prog.clock(clk);
    Seed s = getSeed();
    std::cout << "  seeds: " << s << ' ' << std::endl;
    Rng gen(s);
    ArrayXXd sol = prob->solve(gen, &prog);
    printSolution(sol);
    return sol;
}

// main.cpp:104-143: from every checkpoint of the domain fire +-rot.col(0..2) with maxscat = maxloop = 2 and join the
// polylines, NaN-separated, for plotting
static ArrayXXd checkDomain(const Material* mat, const Domain* dom, const Matrix3d& rot, const Clock& clk) {
    std::vector<ArrayXXd> traj;
    std::vector<long> ind(1, 0);
    const Matrix3Xd pts = dom->checkpoints();
    for (long p = 0; p < pts.cols(); ++p) {
        for (int i = 0; i < 6; ++i) {
            Vector3d dir = rot.col(i % 3);
            if (i >= 3) dir = -dir;
            TrajProblem prob(mat, dom, pts.col(p), dir, 2, 2);
            std::cout << prob << std::endl << std::endl;
            traj.push_back(solveTraj(&prob, clk));
            ind.push_back(ind.back() + traj.back().cols() + 1);
        }
    }
    ArrayXXd sol(3, ind.back() - 1);
    for (long j = 0; j < sol.cols(); ++j) for (int k = 0; k < 3; ++k) sol(k, j) = Dbl::quiet_NaN();
    for (size_t n = 0; n < traj.size(); ++n)
        for (long j = 0; j < traj[n].cols(); ++j) for (int k = 0; k < 3; ++k) sol(k, ind[n] + j) = traj[n](k, j);
    std::cout << "Combined Trajectory" << std::endl;
    printSolution(sol);
    return sol;
}

static ArrayXXd solveField(const FieldProblem* prob, const Clock& clk) {
    static long n = 0;
    std::cout << "Solution " << n++ << std::endl;
    ArrayXXd sol = prob->initSolution();
    Progress prog = prob->initProgress();
    prog.clock(clk);
#pragma omp parallel
    {
        Seed s = getSeed();
        printSeed(s);
        Rng gen(s);
        ArrayXXd partial = prob->solve(gen, &prog);
#pragma omp critical
        { sol += partial; }
    }
    std::cout << "Output" << std::endl;
    printSolution(sol);
    ArrayXXd avg = prob->dom()->average(sol);
    if (avg != sol) { std::cout << "Averaged" << std::endl; printSolution(avg); }
    return avg;
}

static Statistics<ArrayXXd> calcStats(const std::vector<ArrayXXd>& sol) {
    Statistics<ArrayXXd> stats(ArrayXXd::Zero(sol.front().rows(), sol.front().cols()));
    for (const ArrayXXd& s : sol) stats.add(s);
    std::cout << "Mean" << std::endl;
    printSolution(stats.mean());
    std::cout << "Standard Deviation" << std::endl;
    printSolution(stats.variance().sqrt());
    return stats;
}

int main(int argc, const char* argv[]) {
    std::stringstream argss;
    for (int i = 1; i < argc; ++i) argss << argv[i] << ' ';
    std::cout << std::string(40, '-') << std::endl;
    Clock clk;
    std::cout << clk.timestamp() << std::endl;
#ifdef DEBUG
    std::cout << "DEBUG" << std::endl;
#endif
    std::cout << argss.str() << std::endl << std::endl;
    try {
        {   // MCB_DEVICES: "all" (default) | "0,2,3"
            std::vector<int> devs;
            const char* e = std::getenv("MCB_DEVICES");
            if (e && std::string(e) != "all") { std::stringstream ds(e); std::string tok; while (std::getline(ds, tok, ',')) if (!tok.empty()) devs.push_back(std::atoi(tok.c_str())); }
            FieldProblem::devices(devs);
        }
        std::string prefix; argss >> prefix;
        MC_ASSERT_MSG(chdir(prefix.c_str()) == 0, "Invalid directory");

        std::string matStr; argss >> matStr;
        double T = 300.; argss >> T;
        std::unique_ptr<Material> mat;
        if (matStr == "grey") mat.reset(new Material("grey_disp.txt", "grey_relax2.txt", T));
        else if (matStr == "silicon") mat.reset(new Material("Si_disp.txt", "Si_relax2.txt", T));
        else if (matStr == "custom") { std::string disp, relax; argss >> disp >> relax; mat.reset(new Material(disp, relax, T)); }
        else MC_ASSERT_MSG(false, "Invalid material");
        MC_ASSERT_MSG(!argss.fail(), "Invalid material arguments");
        std::cout << *mat << std::endl << std::endl;

        std::string domStr; argss >> domStr;
        std::unique_ptr<Domain> dom;
        if (domStr == "bulk") {
            double d0; long v0; argss >> d0 >> v0;
            dom.reset(new BulkDomain(Vector3d(d0, d0, d0), Vector3l(v0, 0, 0), 1e6 * d0));
        } else if (domStr == "film") {
            double d0, d1; long v1; argss >> d0 >> d1 >> v1;
            dom.reset(new FilmDomain(Vector3d(d0, d1, d0), Vector3l(0, v1, 0), 1e6 * d0));
        } else if (domStr == "slab") {
            double d0, d1, dT; long v0; argss >> d0 >> d1 >> v0 >> dT;
            dom.reset(new SlabDomain(Vector3d(d0, d1, d1), Vector3l(v0, 0, 0), dT));
        } else if (domStr == "wire") {
            double d0, d1; long v1; argss >> d0 >> d1 >> v1;
            dom.reset(new WireDomain(Vector3d(d0, d1, d1), Vector3l(0, v1, v1), 1e6 * d0));
        } else if (domStr == "hex") {
            double d0, d1; argss >> d0 >> d1;
            dom.reset(new HexDomain(VectorXd{d0, d1, d1, d1}, 1e6 * d0));
        } else if (domStr == "pyr") {
            double d0, d1; argss >> d0 >> d1;
            dom.reset(new PyrDomain(Vector3d(d0, d1, d1), 1e6 * d0));
        } else if (domStr == "jct") {
            double d0, d3; argss >> d0 >> d3;
            dom.reset(new JctDomain(VectorXd{d0, d0, d0, d3}, VectorXl{0, 0, 0, 0}, 2e6 * d0));
        } else if (domStr == "tee") {
            double d0, d4; long v0; argss >> d0 >> d4 >> v0;
            dom.reset(new TeeDomain(VectorXd{d0, d0, d0, d0, d4}, VectorXl{v0, v0, v0, v0, 0}, 3e6 * d0));
        } else if (domStr == "tube") {
            double d0, d1, d3; long v1, v3; argss >> d0 >> d1 >> d3 >> v1 >> v3;
            dom.reset(new TubeDomain(VectorXd{d0, d1, d1, d3}, VectorXl{0, v1, v1, v3}, 1e6 * d0));
        } else if (domStr == "octet") {                                  // main.cpp:376-387
            double d0, d1, d2, d3, dT; long v0, v1, v2, v3; argss >> d0 >> d1 >> d2 >> d3 >> v0 >> v1 >> v2 >> v3 >> dT;
            MC_ASSERT_MSG(argss.fail() || (d0 > 0. && d1 > 0. && d2 > 0. && d3 > 0.), "Dimensions must be positive");
            if (!argss.fail()) dom.reset(new OctetDomain(VectorXd{d0, d1, d2, d3}, VectorXl{v0, v1, v2, v3}, dT));
        } else MC_ASSERT_MSG(false, "Invalid domain");
        MC_ASSERT_MSG(!argss.fail(), "Invalid domain arguments");
        std::cout << *dom << std::endl << std::endl;

        std::string probStr; argss >> probStr;
        long nemit = 0, size = 0, maxscat = 0, maxloop = 0, nsim = 0;
        std::unique_ptr<FieldProblem> prob;
        if (probStr == "check") {
            Matrix3d rot;
            argss >> rot(0, 0) >> rot(0, 1) >> rot(0, 2) >> rot(1, 0) >> rot(1, 1) >> rot(1, 2) >> rot(2, 0) >> rot(2, 1) >> rot(2, 2);
            MC_ASSERT_MSG(!argss.fail(), "Invalid problem arguments");
            checkDomain(mat.get(), dom.get(), rot, clk);
        } else if (probStr == "traj") {
            Vector3d pos, dir;
            argss >> pos(0) >> pos(1) >> pos(2) >> dir(0) >> dir(1) >> dir(2) >> maxscat >> maxloop;
            MC_ASSERT_MSG(!argss.fail(), "Invalid problem arguments");
            TrajProblem tp(mat.get(), dom.get(), pos, dir, maxscat, maxloop);
            std::cout << tp << std::endl << std::endl;
            solveTraj(&tp, clk);
        } else if (probStr == "temp") { argss >> nemit >> maxscat >> maxloop >> nsim; prob.reset(new TempProblem(mat.get(), dom.get(), nemit, maxscat, maxloop)); }
        else if (probStr == "flux") { argss >> nemit >> maxscat >> maxloop >> nsim; prob.reset(new FluxProblem(mat.get(), dom.get(), nemit, maxscat, maxloop)); }
        else if (probStr == "multi") { argss >> nemit >> maxscat >> maxloop >> nsim; prob.reset(new MultiProblem(mat.get(), dom.get(), nemit, maxscat, maxloop)); }
        else if (probStr == "cumtemp") { argss >> nemit >> size >> maxscat >> maxloop >> nsim; prob.reset(new CumTempProblem(mat.get(), dom.get(), nemit, size, maxscat, maxloop)); }
        else if (probStr == "cumflux") { argss >> nemit >> size >> maxscat >> maxloop >> nsim; prob.reset(new CumFluxProblem(mat.get(), dom.get(), nemit, size, maxscat, maxloop)); }
        else MC_ASSERT_MSG(false, "Invalid problem");
        MC_ASSERT_MSG(!argss.fail(), "Invalid problem arguments");

        if (prob) std::cout << *prob << std::endl << std::endl;
        if (!prob) {}
        else if (nsim == 1) solveField(prob.get(), clk);
        else {
            std::vector<ArrayXXd> sol;
            for (long i = 0; i < nsim; ++i) sol.push_back(solveField(prob.get(), clk));
            if (!sol.empty()) calcStats(sol);
        }
    } catch (const std::exception& e) {
        std::cerr << "montecarlo: " << e.what() << std::endl;
        return 1;
    }
    std::cout << "Total time: " << clk.stopwatch() << std::endl << std::endl;
    return 0;
}
