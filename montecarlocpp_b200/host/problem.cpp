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
#include <thread>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "domain.h"
#include "field.h"
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

//---------------------------------------- device contexts (one per device, tables cached by object uid)
namespace {
struct DeviceContext {
    mcb_ctx* ctx = 0; unsigned long mat = 0, dom = 0; long cols = 0;
    std::mutex mu;                                  // one GPU stream per device: calls from several host threads are serialised
    ~DeviceContext() { if (ctx) mcb_destroy(ctx); }
};
std::mutex g_mu;                                    // guards the map and the device selection
std::map<int, std::unique_ptr<DeviceContext>> g_ctx;
std::vector<int> g_devices(1, 0);                   // process-wide (an OpenMP worker sees what the main thread selected)
thread_local mcb_stats t_stats = mcb_stats();

void check(mcb_ctx* c, int rc, const char* what) {
    if (rc != MCB_OK) throw std::runtime_error(std::string(what) + ": " + mcb_last_error(c));
}
}
int FieldProblem::deviceCount() { int n = 0; mcb_device_count(&n); return n; }
void FieldProblem::device(int ordinal) { devices(std::vector<int>(1, ordinal)); }
void FieldProblem::devices(const std::vector<int>& ordinals) {
    std::vector<int> d = ordinals;
    if (d.empty()) { const int n = deviceCount(); for (int i = 0; i < n; ++i) d.push_back(i); if (d.empty()) d.push_back(0); }
    std::lock_guard<std::mutex> lock(g_mu);
    g_devices = d;
}
std::vector<int> FieldProblem::devices() { std::lock_guard<std::mutex> lock(g_mu); return g_devices; }
mcb_stats FieldProblem::lastStats() { return t_stats; }

namespace {
// Runs f(ctx) with the context of device `dev`, after making sure the tables of (mat, dom) are resident there.  The cache is
// keyed on the objects' uids (a freed Domain's address is routinely reused by the next one; its uid is not).
template <typename F>
void withDevice(int dev, const Material* mat, const Domain* dom, F f) {
    DeviceContext* dc = 0;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        std::unique_ptr<DeviceContext>& slot = g_ctx[dev];
        if (!slot) {
            slot.reset(new DeviceContext);
            int rc = mcb_create(dev, &slot->ctx);
            if (rc != MCB_OK) { std::string msg = mcb_last_error(0); slot.reset(); g_ctx.erase(dev); throw std::runtime_error("mcb_create: " + msg); }
        }
        dc = slot.get();
    }
    std::lock_guard<std::mutex> lock(dc->mu);
    if (dc->mat != mat->uid()) { mcb_material_desc md = mat->desc(); check(dc->ctx, mcb_upload_material(dc->ctx, &md), "mcb_upload_material"); dc->mat = mat->uid(); }
    if (dc->dom != dom->uid()) {
        FlatDomain fd = flattenDomain(dom); mcb_domain_desc dd = fd.desc();
        check(dc->ctx, mcb_upload_domain(dc->ctx, &dd), "mcb_upload_domain"); dc->dom = dom->uid(); dc->cols = fd.cols;
    }
    f(dc->ctx);
}
int firstDevice() { std::lock_guard<std::mutex> lock(g_mu); return g_devices.front(); }
}


//---------------------------------------- Field::accumulate / CellVolF
VectorXd CellVolF::operator()(const Subdomain* sdom, const Vector3l& index) const { return VectorXd(1, sdom->cellVol(index)); }

Field& Field::accumulate(const Subdomain* sdom, const Vector3d& ipos, const Vector3d& fpos, const VectorXd& amount) {
    MC_ASSERT_MSG(dom_, "Field has no domain");
    MC_ASSERT_MSG((long)amount.size() == data_.rows(), "Amount has the wrong number of rows");
    int32_t index = -1;
    const Subdomain::Pointers& sp = dom_->sdomPtrs();
    for (size_t i = 0; i < sp.size(); ++i) if (sp[i] == sdom) index = (int32_t)i;
    MC_ASSERT_MSG(index >= 0, "Subdomain not in domain");
    const double b[3] = {ipos(0), ipos(1), ipos(2)}, e[3] = {fpos(0), fpos(1), fpos(2)};
    // the tally walk lives on the device (k_accumulate behind mcb_accumulate); only the geometry tables are needed
    DeviceContext* dc = 0;
    const int dev = firstDevice();
    {
        std::lock_guard<std::mutex> lock(g_mu);
        std::unique_ptr<DeviceContext>& slot = g_ctx[dev];
        if (!slot) {
            slot.reset(new DeviceContext);
            if (mcb_create(dev, &slot->ctx) != MCB_OK) { std::string msg = mcb_last_error(0); slot.reset(); g_ctx.erase(dev); throw std::runtime_error("mcb_create: " + msg); }
        }
        dc = slot.get();
    }
    std::lock_guard<std::mutex> lock(dc->mu);
    if (dc->dom != dom_->uid()) {
        FlatDomain fd = flattenDomain(dom_); mcb_domain_desc dd = fd.desc();
        check(dc->ctx, mcb_upload_domain(dc->ctx, &dd), "mcb_upload_domain"); dc->dom = dom_->uid(); dc->cols = fd.cols;
    }
    check(dc->ctx, mcb_accumulate(dc->ctx, (int32_t)data_.rows(), 1, &index, b, e, amount.data(), data_.data()), "mcb_accumulate");
    return *this;
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
    withDevice(firstDevice(), mat(), dom(), [&](mcb_ctx* ctx) { check(ctx, mcb_traj(ctx, &t, seed, &o), "mcb_traj"); });

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
    const std::vector<int> devs = devices();
    const long G = (long)devs.size();
    mcb_stats st = mcb_stats();
    if (G <= 1) {
        withDevice(devs.front(), mat(), dom(), [&](mcb_ctx* ctx) {
            mcb_problem_desc pd = desc();
            check(ctx, mcb_solve(ctx, &pd, seed, n_begin, n_end, out.data(), &st), "mcb_solve");
        });
    } else {
        // Phonons are independent histories: device g takes a contiguous share of [n_begin, n_end) (the static partition the
        // reference gives its OpenMP threads, problem.cpp:383-384) on its own host thread; the raw tallies are summed over
        // NVLink (ncclAllReduce inside mcb_allreduce: the `sol += partial` of main.cpp:162-165) and normalised once.
        const long n = n_end - n_begin, q = n / G, r = n % G;
        std::vector<mcb_stats> sts((size_t)G, mcb_stats());
        std::vector<mcb_ctx*> ctxs((size_t)G, (mcb_ctx*)0);
        std::vector<std::string> errs((size_t)G);
        mcb_problem_desc pd = desc();
        std::vector<std::thread> workers;
        for (long g = 0; g < G; ++g)
            workers.emplace_back([&, g]() {
                const long b = n_begin + g * q + std::min(g, r), e = b + q + (g < r ? 1 : 0);
                try {
                    withDevice(devs[(size_t)g], mat(), dom(), [&](mcb_ctx* ctx) {
                        ctxs[(size_t)g] = ctx;
                        check(ctx, mcb_solve_raw(ctx, &pd, seed, b, e, &sts[(size_t)g]), "mcb_solve_raw");
                    });
                } catch (const std::exception& ex) { errs[(size_t)g] = ex.what(); }
            });
        for (std::thread& w : workers) w.join();
        for (const std::string& e : errs) if (!e.empty()) throw std::runtime_error(e);
        check(ctxs[0], mcb_allreduce(ctxs.data(), (int)G, &pd), "mcb_allreduce");
        check(ctxs[0], mcb_finalize(ctxs[0], &pd, out.data()), "mcb_finalize");
        for (const mcb_stats& s : sts) {
            st.emitted += s.emitted; st.steps += s.steps; st.esc += s.esc; st.launches += s.launches; st.step_launches += s.step_launches;
            st.slot_steps += s.slot_steps; st.state_stores += s.state_stores; st.steady_launches += s.steady_launches;
            st.steady_steps += s.steady_steps; st.steady_stores += s.steady_stores;
            st.device_ms = std::max(st.device_ms, s.device_ms); st.step_ms = std::max(st.step_ms, s.step_ms); st.steady_ms = std::max(st.steady_ms, s.steady_ms);
        }
        st.cols = sts[0].cols; st.launches += 1;
    }
    t_stats = st;
    if (prog) prog->advance((long)st.emitted, (long)st.esc);
    return out;
}

// Reference signature (problem.cpp:370).  `gen` only supplies the 64-bit Philox seed.  The reference calls solve() from EVERY
// thread of `#pragma omp parallel` (main.cpp:155) and splits the particle loop with an orphaned `omp for` (problem.cpp:383);
// the caller then sums the threads' partial fields.  Here OpenMP thread 0 runs the whole range on the device(s) and the other
// threads return a zero partial, so the caller's sum is still exactly ONE solve, whatever the thread count (a GPU solve
// entered T times without this would count every phonon T times).  Every thread draws its two words, so the callers'
// engines stay in step with a serial run.
ArrayXXd FieldProblem::solve(Rng& gen, Progress* prog) const {
    unsigned long long hi = gen(), lo = gen();
    const unsigned long long seed = (hi << 32) | lo;
#ifdef _OPENMP
    if (omp_in_parallel() && omp_get_thread_num() != 0) return initSolution();
#endif
    return solveSeeded(seed, 0, nemit_, prog);
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
