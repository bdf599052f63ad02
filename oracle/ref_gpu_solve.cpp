// oracle/ref_gpu_solve.cpp — TEST INFRASTRUCTURE: the drop-in of INTEGRATION.md §1, compiled and linked for real.
//
// This translation unit defines FieldProblem::solve (problem.h:141) for the REFERENCE'S OWN classes: it flattens the
// reference's Material / Domain / FieldProblem objects into the C-ABI descriptors of include/mcb.h and runs the CUDA path
// (libmcb.so).  oracle/Makefile links it with the reference's untouched objects -- main.o included -- into
// oracle/_ref/montecarlo_gpu; the reference's CPU body of the same function (problem.cpp:370-445) is demoted to a weak symbol
// with objcopy, so every caller (solveField, main.cpp:146-179, from inside `#pragma omp parallel`) and every vtable slot binds
// to this definition.  Argument grammar, seeds, progress bar and stdout blocks are the reference's own.
//
// Calling pattern: solve() is entered by EVERY thread of the parallel region and the reference splits the
// particle loop with an orphaned `omp for`.  Here the thread that arrives as OpenMP thread 0 runs the whole range on the GPU
// and the others return a zero partial, so the caller's `sol += partial` still sums to ONE solve (no T-fold over-count), for
// any thread count.  The Philox seed is taken from thread 0's generator (two words, like host/problem.cpp), or from the
// environment variable MCB_SEED when set (tests: the reference assigns random_device seeds to threads in arrival order).
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <boost/fusion/algorithm/iteration.hpp>
#include <boost/optional.hpp>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <typeinfo>
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
#include "constants.h"
#undef private
#undef protected
#include "../include/mcb.h"
#include "ref_flatten.h"

namespace {

struct FlatMaterial {
    std::vector<double> vel, tau, flux, scat;
    mcb_material_desc desc(const Material* m) const {
        mcb_material_desc d; std::memset(&d, 0, sizeof d);
        d.nw = m->nw_; d.np = m->np_; d.temp = m->T_; d.vel = vel.data(); d.tau = tau.data(); d.flux_pdf = flux.data(); d.scat_pdf = scat.data();
        d.energy_sum = m->energySum_; d.flux_sum = m->fluxSum_; d.scat_sum = m->scatSum_;
        return d;
    }
};
// Material keeps vel_ / tau_ / omega_ but not its pdfs (locals of the constructor, material.cpp:147-158): a maintainer would keep
// them as members; here dos and domega are re-read from the dispersion file the object names (material.cpp:86-107) and the
// pdfs recomputed with the constructor's own expressions
void flattenMaterial(const Material* m, FlatMaterial& out) {
    const long nw = m->nw_, np = m->np_;
    std::ifstream f(m->disp_.c_str());
    long fnw = 0, fnp = 0;
    if (!(f >> fnw >> fnp) || fnw != nw || fnp != np) refflat::die("cannot re-read the dispersion file");
    ArrayXXd dos(nw, np); ArrayXd domega(nw);
    for (long w = 0; w < nw; ++w) {
        double omega, dw; f >> omega >> dw; domega(w) = dw;
        for (long p = 0; p < np; ++p) { double v, d; f >> v >> d; dos(w, p) = d; }
    }
    if (!f) refflat::die("short dispersion file");
    out.vel.resize(nw * np); out.tau.resize(nw * np); out.flux.resize(nw * np); out.scat.resize(nw * np);
    for (long w = 0; w < nw; ++w) {
        const double x = HBAR / (KB * m->T_) * m->omega_(w);
        const double dedT = KB * (std::fabs(x) < Dbl::epsilon() ? 1. - x * x / 12. : std::pow(x / (2. * std::sinh(x / 2.)), 2));
        for (long p = 0; p < np; ++p) {
            const long k = w + nw * p;                             // column-major like ArrayXXd
            const double energy = dedT * dos(w, p) * domega(w);
            out.vel[k] = m->vel_(w, p); out.tau[k] = m->tau_(w, p);
            out.flux[k] = m->vel_(w, p) * energy; out.scat[k] = energy / m->tau_(w, p);
        }
    }
}

int problemKind(const FieldProblem* p, long* size, long* step) {
    *size = 0; *step = 0;
    if (dynamic_cast<const TempProblem*>(p)) return MCB_PROB_TEMP;
    if (dynamic_cast<const FluxProblem*>(p)) return MCB_PROB_FLUX;
    if (dynamic_cast<const MultiProblem*>(p)) return MCB_PROB_MULTI;
    if (const CumTempProblem* c = dynamic_cast<const CumTempProblem*>(p)) { *size = c->size_; *step = c->step_; return MCB_PROB_CUMTEMP; }
    if (const CumFluxProblem* c = dynamic_cast<const CumFluxProblem*>(p)) { *size = c->size_; *step = c->step_; return MCB_PROB_CUMFLUX; }
    refflat::die("unknown FieldProblem class");
    return -1;
}

// the progress bar of Progress::incrCount (problem.cpp:91-109) advanced by n particles at once: same ticks, same text
void advanceProgress(Progress* prog, long n, long esc) {
    if (!prog) return;
    prog->esc_ += esc;
    const long target = prog->count_ + n;
    while (prog->next_ < prog->div_ && prog->vec_.at(prog->next_) <= target) {
        prog->count_ = prog->vec_.at(prog->next_) - 1;
        prog->incrCount();
    }
    if (prog->count_ < target) { prog->count_ = target - 1; prog->incrCount(); }
}

mcb_ctx* g_ctx = 0;
const void* g_mat = 0; const void* g_dom = 0;
unsigned long long g_seed = 0;

} // namespace

ArrayXXd FieldProblem::solve(Rng& gen, Progress* prog) const
{
    // 64-bit Philox seed from the caller's engine (every thread draws its two words, so the engines stay in step)
    unsigned long long seed = ((unsigned long long)(gen() & 0xFFFFFFFFul) << 32) | (unsigned long long)(gen() & 0xFFFFFFFFul);
    if (const char* s = std::getenv("MCB_SEED")) seed = std::strtoull(s, 0, 10);
    ArrayXXd sol = initSolution();                                // rows() x cols, zero (field.cpp:25-45)
    int tid = 0;
#ifdef _OPENMP
    if (omp_in_parallel()) tid = omp_get_thread_num();
#endif
    if (tid != 0) return sol;                                     // the caller sums the partials: thread 0 brings the whole solve

    int dev = 0;
    if (const char* s = std::getenv("MCB_DEVICE")) dev = std::atoi(s);
    if (!g_ctx && mcb_create(dev, &g_ctx) != MCB_OK) { std::cerr << "mcb_create: " << mcb_last_error(0) << std::endl; std::abort(); }
    if (g_mat != (const void*)mat()) {
        FlatMaterial fm; flattenMaterial(mat(), fm);
        mcb_material_desc md = fm.desc(mat());
        if (mcb_upload_material(g_ctx, &md) != MCB_OK) { std::cerr << "mcb_upload_material: " << mcb_last_error(g_ctx) << std::endl; std::abort(); }
        g_mat = mat();
    }
    if (g_dom != (const void*)dom()) {
        refflat::FlatDomain fd; refflat::flattenDomain(dom(), refflat::cellsOf(dom()), fd);
        mcb_domain_desc dd = fd.desc();
        if (mcb_upload_domain(g_ctx, &dd) != MCB_OK) { std::cerr << "mcb_upload_domain: " << mcb_last_error(g_ctx) << std::endl; std::abort(); }
        g_dom = dom();
    }
    long size = 0, step = 0;
    mcb_problem_desc pd; std::memset(&pd, 0, sizeof pd);
    pd.kind = problemKind(this, &size, &step); pd.rows = (int32_t)rows(); pd.size = size; pd.step = step;
    pd.nemit = nemit_; pd.maxscat = maxscat_; pd.maxloop = maxloop_; pd.power = power_;
    std::vector<int64_t> counts(emitPdf_.data(), emitPdf_.data() + emitPdf_.size());
    pd.emit_count = counts.data();
    mcb_stats st;
    if (mcb_solve(g_ctx, &pd, seed, 0, nemit_, sol.data(), &st) != MCB_OK) { std::cerr << "mcb_solve: " << mcb_last_error(g_ctx) << std::endl; std::abort(); }
    advanceProgress(prog, (long)st.emitted, (long)st.esc);
    if (std::getenv("MCB_VERBOSE"))
        std::fprintf(stderr, "mcb: seed %llu emitted %lld steps %lld esc %lld launches %lld device_ms %.3f\n", seed,
                     (long long)st.emitted, (long long)st.steps, (long long)st.esc, (long long)st.launches, st.device_ms);
    g_seed = seed;
    return sol;
}
