// problem.cpp — host mirror of problem.cpp:31-160 (Clock, Progress, Problem) and :315-648 (FieldProblem
// family).  The particle/step loops of FieldProblem::solve (problem.cpp:383-437) and the normalisation
// (:439-444) run on the GPU behind mcb_solve (include/mcb.h).
#include "problem.h"
#include <algorithm>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "domain.h"
#include "material.h"

//---------------------------------------- Clock / Progress
Clock::Clock() { std::time(&start_); }
std::string Clock::stopwatch() {
    std::time_t now; std::time(&now);
    long diff = static_cast<long>(std::difftime(now, start_));
    std::ostringstream ss;
    ss << std::setfill('0') << std::setw(2) << diff / 3600 << ':' << std::setw(2) << diff / 60 % 60 << ':' << std::setw(2) << diff % 60;
    return ss.str();
}
std::string Clock::timestamp() {
    std::time_t now; std::time(&now);
    char buffer[80];
    std::strftime(buffer, sizeof buffer, "%Y-%m-%d %H:%M:%S", std::localtime(&now));
    return std::string(buffer);
}

Progress::Progress() : tot_(0), count_(0), div_(0), next_(0), esc_(0) {}
Progress::Progress(long tot, long div) : tot_(tot), count_(0), div_(div), next_(0), esc_(0) {
    MC_ASSERT_MSG(tot_ >= div_, "Too many divisions");
    for (long i = 1; i <= div_; ++i) vec_.push_back(tot_ * i / div_);
}
long Progress::incrCount() { advance(1, 0); return count_; }
long Progress::incrEsc() {
#pragma omp critical(mcb_progress)
    { esc_++; }
    return esc_;
}
void Progress::advance(long n, long e) {
#pragma omp critical(mcb_progress)
    {
        esc_ += e;
        count_ += n;
        while (next_ < div_ && count_ >= vec_.at((size_t)next_)) {
            next_++;
            std::cout << clk_.stopwatch() << ' ' << '[' << std::string((size_t)next_, '|') << std::string((size_t)(div_ - next_), '-') << ']';
            std::cout << " esc: " << esc_ << std::endl;
        }
        if (n > 0 && count_ == tot_) std::cout << std::endl;
    }
}

//---------------------------------------- Problem
Problem::Problem() : mat_(0), dom_(0) {}
Problem::Problem(const Material* mat, const Domain* dom) : mat_(mat), dom_(dom) {
    MC_ASSERT_MSG(mat_ && dom_, "Null material or domain");
    MC_ASSERT_MSG(dom_->isInit(), "Domain setup not complete");
}
Problem::~Problem() {}
std::string Problem::info() const {
    std::ostringstream ss;
    ss << "  mat:     " << mat_ << std::endl;
    ss << "  dom:     " << dom_;
    return ss.str();
}
std::ostream& operator<<(std::ostream& os, const Problem& prob) { return os << prob.info(); }

//---------------------------------------- flattening (the drop-in boundary)
mcb_domain_desc FlatDomain::desc() const {
    mcb_domain_desc d;
    d.nsdom = (int32_t)sdoms.size(); d.sdoms = sdoms.data();
    d.nplane = (int32_t)planes.size(); d.planes = planes.data();
    d.npair = (int32_t)pairs.size(); d.pairs = pairs.data();
    d.nemitter = (int32_t)emitters.size(); d.emitters = emitters.data();
    d.ncols = (int64_t)cell_vol.size(); d.cell_vol = cell_vol.empty() ? nullptr : cell_vol.data();
    return d;
}

FlatDomain flattenDomain(const Domain* dom) {
    MC_ASSERT_MSG(dom && dom->isInit(), "Domain setup not complete");
    FlatDomain f; f.cols = 0;
    std::map<const Boundary*, int32_t> planeId;
    std::map<const Subdomain*, int32_t> sdomId;
    const Subdomain::Pointers& sp = dom->sdomPtrs();
    for (size_t s = 0; s < sp.size(); ++s) {
        sdomId[sp[s]] = (int32_t)s;
        mcb_sdom_desc d; sp[s]->describe(d);
        d.plane_begin = (int32_t)f.planes.size(); d.plane_count = (int32_t)sp[s]->bdryPtrs().size();
        for (const Boundary* b : sp[s]->bdryPtrs()) {
            planeId[b] = (int32_t)f.planes.size();
            mcb_plane_desc p; b->describe(p);
            p.sdom = (int32_t)s;
            f.planes.push_back(p);
        }
        f.sdoms.push_back(d);
        f.cols += sp[s]->shape().prod();
        const Vector3l shp = sp[s]->shape();                               // Field(rows, dom, fun) nesting: k, j, i (field.cpp:62-78)
        for (long k = 0; k < shp(2); ++k) for (long j = 0; j < shp(1); ++j) for (long i = 0; i < shp(0); ++i)
            f.cell_vol.push_back(sp[s]->cellVol(Vector3l(i, j, k)));
    }
    for (size_t s = 0; s < sp.size(); ++s) {
        const Boundary::Pointers& bp = sp[s]->bdryPtrs();
        for (size_t k = 0; k < bp.size(); ++k) {
            mcb_plane_desc& p = f.planes[(size_t)f.sdoms[s].plane_begin + k];
            const Boundary::Pointers partners = bp[k]->partners();
            p.pair_begin = (int32_t)f.pairs.size(); p.pair_count = (int32_t)partners.size();
            for (const Boundary* q : partners) {
                MC_ASSERT_MSG(planeId.count(q), "Boundary paired with a boundary outside the domain");
                f.pairs.push_back(planeId[q]);
            }
        }
    }
    for (const Emitter* e : dom->emitPtrs()) {
        mcb_emitter_desc d;
        if (e->emitBdry()) { d.kind = MCB_EMIT_BDRY; d.index = planeId.at(e->emitBdry()); }
        else               { d.kind = MCB_EMIT_SDOM; d.index = sdomId.at(e->emitSdom()); }
        d.weight = e->emitWeight();
        f.emitters.push_back(d);
    }
    return f;
}

//---------------------------------------- device contexts (one per device, tables cached by object identity)
namespace {
struct DeviceContext {
    mcb_ctx* ctx = 0; const Material* mat = 0; const Domain* dom = 0; long cols = 0;
    ~DeviceContext() { if (ctx) mcb_destroy(ctx); }
};
std::mutex g_mu;
std::map<int, std::unique_ptr<DeviceContext>> g_ctx;
thread_local int t_device = 0;
thread_local mcb_stats t_stats = mcb_stats();

void check(mcb_ctx* c, int rc, const char* what) {
    if (rc != MCB_OK) throw std::runtime_error(std::string(what) + ": " + mcb_last_error(c));
}
}
void FieldProblem::device(int ordinal) { t_device = ordinal; }
mcb_stats FieldProblem::lastStats() { return t_stats; }

namespace {
// Runs f(ctx) with the device context of the calling thread's device, after making sure the tables of (mat, dom) are
// resident.  One GPU stream per device: calls from several host threads are serialised.
template <typename F>
void withDevice(const Material* mat, const Domain* dom, F f) {
    std::lock_guard<std::mutex> lock(g_mu);
    std::unique_ptr<DeviceContext>& dc = g_ctx[t_device];
    if (!dc) {
        dc.reset(new DeviceContext);
        int rc = mcb_create(t_device, &dc->ctx);
        if (rc != MCB_OK) { std::string msg = mcb_last_error(0); dc.reset(); g_ctx.erase(t_device); throw std::runtime_error("mcb_create: " + msg); }
    }
    if (dc->mat != mat) { mcb_material_desc md = mat->desc(); check(dc->ctx, mcb_upload_material(dc->ctx, &md), "mcb_upload_material"); dc->mat = mat; }
    if (dc->dom != dom) {
        FlatDomain fd = flattenDomain(dom); mcb_domain_desc dd = fd.desc();
        check(dc->ctx, mcb_upload_domain(dc->ctx, &dd), "mcb_upload_domain"); dc->dom = dom; dc->cols = fd.cols;
    }
    f(dc->ctx);
}
}


//---------------------------------------- TrajProblem
TrajProblem::TrajProblem() : maxscat_(0), maxloop_(0) {}
TrajProblem::TrajProblem(const Material* m, const Domain* d, const Prop& prop, const Vector3d& pos, const Vector3d& dir, long maxscat, long maxloop)
    : Problem(m, d), maxscat_(maxscat), maxloop_(maxloop), prop_(prop), pos_(pos), dir_(dir) {}
TrajProblem::TrajProblem(const Material* m, const Domain* d, const Prop& prop, const Vector3d& pos, long maxscat, long maxloop)
    : Problem(m, d), maxscat_(maxscat), maxloop_(maxloop), prop_(prop), pos_(pos) {}
TrajProblem::TrajProblem(const Material* m, const Domain* d, const Vector3d& pos, const Vector3d& dir, long maxscat, long maxloop)
    : Problem(m, d), maxscat_(maxscat), maxloop_(maxloop), pos_(pos), dir_(dir) {}
TrajProblem::TrajProblem(const Material* m, const Domain* d, const Vector3d& pos, long maxscat, long maxloop)
    : Problem(m, d), maxscat_(maxscat), maxloop_(maxloop), pos_(pos) {}
TrajProblem::TrajProblem(const Material* m, const Domain* d, long maxscat, long maxloop)
    : Problem(m, d), maxscat_(maxscat), maxloop_(maxloop) {}

std::string TrajProblem::info() const {
    std::ostringstream ss;
    ss << "TrajProblem " << static_cast<const Problem*>(this) << std::endl;
    ss << Problem::info() << std::endl;
    if (prop_) ss << "  prop:    " << prop_->w << " " << prop_->p << std::endl;
    if (pos_) ss << "  pos:     [" << (*pos_)(0) << " " << (*pos_)(1) << " " << (*pos_)(2) << "]" << std::endl;
    if (dir_) ss << "  dir:     [" << (*dir_)(0) << " " << (*dir_)(1) << " " << (*dir_)(2) << "]" << std::endl;
    ss << "  maxscat: " << maxscat_ << std::endl;
    ss << "  maxloop: " << maxloop_;
    return ss.str();
}
Progress TrajProblem::initProgress() const { return Progress(); }

ArrayXXd TrajProblem::solve(Rng& gen, Progress* prog) const {
    unsigned long long hi = gen(), lo = gen();
    const unsigned long long seed = (hi << 32) | lo;
    mcb_traj_desc t = mcb_traj_desc();
    t.maxscat = maxscat_;
    t.maxloop = (maxloop_ != 0 ? maxloop_ : loopFactor_ * maxscat_);              // problem.cpp:256
    t.sdom = -1;
    if (prop_) { t.has_prop = 1; t.w = prop_->w; t.p = prop_->p; }
    const Subdomain::Pointers& sp = dom()->sdomPtrs();
    if (pos_) {
        t.has_pos = 1;
        for (int k = 0; k < 3; ++k) t.pos[k] = (*pos_)(k);
        const Subdomain* s = dom()->locate(*pos_);                               // problem.cpp:241-242
        MC_ASSERT_MSG(s, "Position not inside domain");
        for (size_t i = 0; i < sp.size(); ++i) if (sp[i] == s) t.sdom = (int32_t)i;
        if (dir_) { t.has_dir = 1; for (int k = 0; k < 3; ++k) t.dir[k] = (*dir_)(k); }
    }
    const long nmax = std::max(t.maxloop, 1l);
    std::vector<double> pts((size_t)(3 * (2 * nmax + 1)));
    std::vector<int32_t> ssd((size_t)nmax), sin_((size_t)nmax), sink((size_t)nmax), sout((size_t)nmax), soutk((size_t)nmax);
    mcb_traj_out o = mcb_traj_out();
    o.max_points = 2 * nmax + 1; o.points = pts.data(); o.max_steps = nmax;
    o.step_sdom = ssd.data(); o.step_in = sin_.data(); o.step_in_kind = sink.data(); o.step_out = sout.data(); o.step_out_kind = soutk.data();
    withDevice(mat(), dom(), [&](mcb_ctx* ctx) { check(ctx, mcb_traj(ctx, &t, seed, &o), "mcb_traj"); });

    // the per-trip lines of problem.cpp:260-275
    auto typeOf = [&](int s, int k) { return k >= 0 ? sp.at((size_t)s)->bdryPtrs().at((size_t)k)->type() : std::string("Null"); };
    for (long i = 0; i < o.nsteps; ++i) {
        std::cout << std::setw(2) << ssd[(size_t)i] << ": " << std::setw(2) << sin_[(size_t)i] << " " << std::setw(5) << typeOf(ssd[(size_t)i], sin_[(size_t)i]) << " -> ";
        if (o.escaped == 1 && i == o.nsteps - 1) break;        // killed inside advect: the reference breaks before the right-hand side
        std::cout << std::setw(2) << sout[(size_t)i] << " " << std::setw(5) << typeOf(ssd[(size_t)i], sout[(size_t)i]) << std::endl;
    }
    if (o.escaped) { if (prog) prog->incrEsc(); std::cout << "Escaped" << std::endl; }
    ArrayXXd traj(3, o.npoints);
    for (long j = 0; j < o.npoints; ++j) for (int k = 0; k < 3; ++k) traj(k, j) = pts[(size_t)(3 * j + k)];
    return traj;
}

//---------------------------------------- FieldProblem
FieldProblem::FieldProblem() : nemit_(0), maxscat_(0), maxloop_(0), power_(0.) {}

FieldProblem::FieldProblem(const Material* mat, const Domain* dom, long nemit, long maxscat, long maxloop) : Problem(mat, dom) {
    const Emitter::Pointers emitPtrs = dom->emitPtrs();
    const long nemitter = (long)emitPtrs.size();
    MC_ASSERT_MSG(nemitter > 0, "No phonons to emit");
    VectorXd weight((size_t)nemitter);
    double weightSum = 0.;
    for (long i = 0; i < nemitter; ++i) { weight[(size_t)i] = emitPtrs.at((size_t)i)->emitWeight(); weightSum += weight[(size_t)i]; }
    emitPdf_.resize((size_t)nemitter);
    nemit_ = 0;
    for (long i = 0; i < nemitter; ++i) {                                   // problem.cpp:329-335
        double frac = weight[(size_t)i] / weightSum;
        double rounded = std::ceil(frac * nemit - 0.5);
        emitPdf_[(size_t)i] = std::max(1l, static_cast<long>(rounded));
        nemit_ += emitPdf_[(size_t)i];
    }
    maxscat_ = maxscat;
    maxloop_ = (maxloop != 0 ? maxloop : loopFactor_ * maxscat_);
    power_ = weightSum / nemit_ * mat->fluxSum() / 4.;                      // problem.cpp:341
}
FieldProblem::~FieldProblem() {}

std::string FieldProblem::info() const {
    std::ostringstream ss;
    ss << Problem::info() << std::endl;
    ss << "  emitpdf: [";
    for (size_t i = 0; i < emitPdf_.size(); ++i) ss << (i ? " " : "") << emitPdf_[i];
    ss << "]" << std::endl;
    ss << "  nemit:   " << nemit_ << std::endl;
    ss << "  maxscat: " << maxscat_ << std::endl;
    ss << "  maxloop: " << maxloop_ << std::endl;
    ss << "  power:   " << power_;
    return ss.str();
}

Progress FieldProblem::initProgress() const { return Progress(nemit_, std::min(20l, nemit_)); }

ArrayXXd FieldProblem::initSolution() const {
    long cols = 0;
    for (const Subdomain* s : dom()->sdomPtrs()) cols += s->shape().prod();   // Field::init field.cpp:25-45
    return ArrayXXd(rows(), cols);
}

mcb_problem_desc FieldProblem::desc() const {
    emit64_.assign(emitPdf_.begin(), emitPdf_.end());
    mcb_problem_desc d;
    d.kind = kind(); d.rows = (int32_t)rows(); d.size = size(); d.step = step();
    d.nemit = nemit_; d.maxscat = maxscat_; d.maxloop = maxloop_; d.power = power_; d.emit_count = emit64_.data();
    return d;
}

ArrayXXd FieldProblem::solveSeeded(unsigned long long seed, long n_begin, long n_end, Progress* prog) const {
    ArrayXXd out = initSolution();
    mcb_stats st = mcb_stats();
    withDevice(mat(), dom(), [&](mcb_ctx* ctx) {
        mcb_problem_desc pd = desc();
        check(ctx, mcb_solve(ctx, &pd, seed, n_begin, n_end, out.data(), &st), "mcb_solve");
    });
    t_stats = st;
    if (prog) prog->advance((long)st.emitted, (long)st.esc);
    return out;
}

// Reference signature (problem.cpp:370).  `gen` only supplies the 64-bit Philox seed.  Called from inside
// `#pragma omp parallel` (main.cpp:155) every thread takes the static chunk the reference's orphaned
// `omp for schedule(static)` (problem.cpp:383) would give it, so the partials still add up to one solve.
ArrayXXd FieldProblem::solve(Rng& gen, Progress* prog) const {
    unsigned long long hi = gen(), lo = gen();
    unsigned long long seed = (hi << 32) | lo;
    long n_begin = 0, n_end = nemit_;
#ifdef _OPENMP
    if (omp_in_parallel()) {
        // all threads must agree on the stream: take thread 0's seed
        static unsigned long long shared_seed;
#pragma omp barrier
#pragma omp master
        { shared_seed = seed; }
#pragma omp barrier
        seed = shared_seed;
        const long T = omp_get_num_threads(), t = omp_get_thread_num();
        const long q = nemit_ / T, r = nemit_ % T;                          // static schedule: first r chunks get q+1
        n_begin = t * q + std::min(t, r);
        n_end = n_begin + q + (t < r ? 1 : 0);
    }
#endif
    return solveSeeded(seed, n_begin, n_end, prog);
}

//---------------------------------------- the five tallies
TempProblem::TempProblem() {}
TempProblem::TempProblem(const Material* m, const Domain* d, long nemit, long maxscat, long maxloop) : FieldProblem(m, d, nemit, maxscat, maxloop) {}
std::string TempProblem::info() const { std::ostringstream ss; ss << "TempProblem " << static_cast<const Problem*>(this) << std::endl << FieldProblem::info(); return ss.str(); }
long TempProblem::rows() const { return 1; }
int TempProblem::kind() const { return MCB_PROB_TEMP; }

FluxProblem::FluxProblem() {}
FluxProblem::FluxProblem(const Material* m, const Domain* d, long nemit, long maxscat, long maxloop) : FieldProblem(m, d, nemit, maxscat, maxloop) {}
std::string FluxProblem::info() const { std::ostringstream ss; ss << "FluxProblem " << static_cast<const Problem*>(this) << std::endl << FieldProblem::info(); return ss.str(); }
long FluxProblem::rows() const { return 3; }
int FluxProblem::kind() const { return MCB_PROB_FLUX; }

MultiProblem::MultiProblem() {}
MultiProblem::MultiProblem(const Material* m, const Domain* d, long nemit, long maxscat, long maxloop) : FieldProblem(m, d, nemit, maxscat, maxloop) {}
std::string MultiProblem::info() const { std::ostringstream ss; ss << "MultiProblem " << static_cast<const Problem*>(this) << std::endl << FieldProblem::info(); return ss.str(); }
long MultiProblem::rows() const { return 4; }
int MultiProblem::kind() const { return MCB_PROB_MULTI; }

namespace { long cumStep(long maxscat, long size) { MC_ASSERT_MSG(size > 0, "size must be positive"); long s = (maxscat - 1) / size; if ((maxscat - 1) % size != 0) s++; return s; } }

CumTempProblem::CumTempProblem() : size_(0), step_(0) {}
CumTempProblem::CumTempProblem(const Material* m, const Domain* d, long nemit, long size, long maxscat, long maxloop)
    : FieldProblem(m, d, nemit, maxscat, maxloop), size_(size), step_(cumStep(maxscat, size)) {}
std::string CumTempProblem::info() const {
    std::ostringstream ss; ss << "CumTempProblem " << static_cast<const Problem*>(this) << std::endl << FieldProblem::info() << std::endl << "  size:    " << size_; return ss.str();
}
long CumTempProblem::rows() const { return size_ + 1; }
int CumTempProblem::kind() const { return MCB_PROB_CUMTEMP; }

CumFluxProblem::CumFluxProblem() : size_(0), step_(0) {}
CumFluxProblem::CumFluxProblem(const Material* m, const Domain* d, long nemit, long size, long maxscat, long maxloop)
    : FieldProblem(m, d, nemit, maxscat, maxloop), size_(size), step_(cumStep(maxscat, size)) {}
std::string CumFluxProblem::info() const {
    std::ostringstream ss; ss << "CumFluxProblem " << static_cast<const Problem*>(this) << std::endl << FieldProblem::info() << std::endl << "  size:    " << size_; return ss.str();
}
long CumFluxProblem::rows() const { return 3 * (size_ + 1); }
int CumFluxProblem::kind() const { return MCB_PROB_CUMFLUX; }
