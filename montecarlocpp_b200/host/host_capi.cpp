// host_capi.cpp — a small C surface over the C++ host mirror so that tests and bench.py (Python, ctypes)
// can drive the SAME objects a C++ caller would: Material, the Domain classes, the FieldProblem family.
// Not part of the reference API; every call maps 1:1 to a constructor or method of the mirror.
#include <cstring>
#include <omp.h>
#include <stdexcept>
#include <vector>
#include <memory>
#include <string>
#include "domain.h"
#include "material.h"
#include "problem.h"

namespace { thread_local std::string g_err; }
#define MCBH_TRY(body, fail) try { body } catch (const std::exception& e) { g_err = e.what(); return fail; }

struct mcbh_domain { std::unique_ptr<Domain> dom; FlatDomain flat; };
struct mcbh_problem { std::unique_ptr<FieldProblem> prob; };

static int solve_like_the_reference(const mcbh_problem* p, uint32_t mt_seed, int nthreads, double* out, mcb_stats* stats, int64_t* progress_count) {
    try {
        ArrayXXd sol = p->prob->initSolution();
        Progress prog = p->prob->initProgress();
        std::string err; mcb_stats st = mcb_stats();
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
        {
            ArrayXXd partial;
            bool ok = true;
            try {
                Rng gen(mt_seed + (uint32_t)omp_get_thread_num());
                partial = p->prob->solve(gen, &prog);
            } catch (const std::exception& e) {
                ok = false;
#pragma omp critical(mcbh_err)
                { err = e.what(); }
            }
#pragma omp critical(mcbh_sum)
            {
                if (ok) { sol += partial; if (omp_get_thread_num() == 0) st = FieldProblem::lastStats(); }
            }
        }
        if (!err.empty()) throw std::runtime_error(err);
        std::memcpy(out, sol.data(), sizeof(double) * (size_t)sol.size());
        if (stats) *stats = st;
        if (progress_count) *progress_count = prog.count();
        return MCB_OK;
    } catch (const std::exception& e) { g_err = e.what(); return MCB_EINVAL; }
}

extern "C" {

const char* mcbh_last_error(void) { return g_err.c_str(); }

Material* mcbh_material_create(const char* disp, const char* relax, double temp) {
    MCBH_TRY(return new Material(disp, relax, temp);, nullptr)
}
void mcbh_material_free(Material* m) { delete m; }
int mcbh_material_desc(const Material* m, mcb_material_desc* out) { if (!m || !out) return MCB_EINVAL; *out = m->desc(); return MCB_OK; }
double mcbh_material_cond(const Material* m) { return m->cond(); }

// kind: bulk film slab wire (3 dims, 3 divs) | jct tube (4, 4) | tee (5, 5): the constructor vectors
mcbh_domain* mcbh_domain_create(const char* kind, const double* dim, int ndim, const int64_t* div, int ndiv, double dT) {
    MCBH_TRY(
        std::string k(kind);
        VectorXd d(dim, dim + ndim); VectorXl v(div, div + ndiv);
        std::unique_ptr<mcbh_domain> h(new mcbh_domain);
        auto need = [&](size_t nd, size_t nv) { MC_ASSERT_MSG(d.size() == nd && v.size() == nv, "wrong number of dims/divs for domain " + k); };
        if (k == "bulk" || k == "film" || k == "slab" || k == "wire") {
            need(3, 3);
            Vector3d D(d[0], d[1], d[2]); Vector3l V(v[0], v[1], v[2]);
            if (k == "bulk") h->dom.reset(new BulkDomain(D, V, dT));
            else if (k == "film") h->dom.reset(new FilmDomain(D, V, dT));
            else if (k == "slab") h->dom.reset(new SlabDomain(D, V, dT));
            else h->dom.reset(new WireDomain(D, V, dT));
        } else if (k == "jct") { need(4, 4); h->dom.reset(new JctDomain(d, v, dT)); }
        else if (k == "tee") { need(5, 5); h->dom.reset(new TeeDomain(d, v, dT)); }
        else if (k == "tube") { need(4, 4); h->dom.reset(new TubeDomain(d, v, dT)); }
        else if (k == "octet") { need(4, 4); h->dom.reset(new OctetDomain(d, v, dT)); }
        else if (k == "hex") { need(4, 0); h->dom.reset(new HexDomain(d, dT)); }
        else if (k == "pyr") { need(3, 0); h->dom.reset(new PyrDomain(Vector3d(d[0], d[1], d[2]), dT)); }
        else MC_ASSERT_MSG(false, "Invalid domain");
        h->flat = flattenDomain(h->dom.get());
        return h.release();
    , nullptr)
}
void mcbh_domain_free(mcbh_domain* d) { delete d; }
int mcbh_domain_desc(const mcbh_domain* d, mcb_domain_desc* out) { if (!d || !out) return MCB_EINVAL; *out = d->flat.desc(); return MCB_OK; }
int64_t mcbh_domain_cols(const mcbh_domain* d) { return d->flat.cols; }
// Domain::average applied to a rows x cols solution (column-major): returns the number of columns of the result
// (cols for every shipped domain but OctetDomain, whose average is rows x 1, domain.cpp:1252-1257)
int64_t mcbh_domain_average(const mcbh_domain* d, const double* sol, int64_t rows, double* out) {
    MCBH_TRY(
        ArrayXXd in(rows, d->flat.cols);
        std::memcpy(in.data(), sol, sizeof(double) * (size_t)in.size());
        ArrayXXd avg = d->dom->average(in);
        std::memcpy(out, avg.data(), sizeof(double) * (size_t)avg.size());
        return (int64_t)avg.cols();
    , -1)
}

mcbh_problem* mcbh_problem_create(const Material* mat, const mcbh_domain* dom, int kind, int64_t nemit, int64_t size,
                                  int64_t maxscat, int64_t maxloop) {
    MCBH_TRY(
        MC_ASSERT_MSG(mat && dom, "Null material or domain");
        std::unique_ptr<mcbh_problem> h(new mcbh_problem);
        const Domain* D = dom->dom.get();
        switch (kind) {
        case MCB_PROB_TEMP: h->prob.reset(new TempProblem(mat, D, nemit, maxscat, maxloop)); break;
        case MCB_PROB_FLUX: h->prob.reset(new FluxProblem(mat, D, nemit, maxscat, maxloop)); break;
        case MCB_PROB_MULTI: h->prob.reset(new MultiProblem(mat, D, nemit, maxscat, maxloop)); break;
        case MCB_PROB_CUMTEMP: h->prob.reset(new CumTempProblem(mat, D, nemit, size, maxscat, maxloop)); break;
        case MCB_PROB_CUMFLUX: h->prob.reset(new CumFluxProblem(mat, D, nemit, size, maxscat, maxloop)); break;
        default: MC_ASSERT_MSG(false, "Invalid problem");
        }
        return h.release();
    , nullptr)
}
void mcbh_problem_free(mcbh_problem* p) { delete p; }
int mcbh_problem_desc(const mcbh_problem* p, mcb_problem_desc* out) { if (!p || !out) return MCB_EINVAL; *out = p->prob->desc(); return MCB_OK; }

// FieldProblem::solve(gen, prog) with gen = mt19937(mt_seed), exactly as a C++ caller would invoke it
int mcbh_problem_solve(const mcbh_problem* p, int device, uint32_t mt_seed, double* out, mcb_stats* stats) {
    MCBH_TRY(
        FieldProblem::device(device);
        Rng gen(mt_seed);
        ArrayXXd sol = p->prob->solve(gen, nullptr);
        std::memcpy(out, sol.data(), sizeof(double) * (size_t)sol.size());
        if (stats) *stats = FieldProblem::lastStats();
        return MCB_OK;
    , MCB_EINVAL)
}
int mcbh_problem_solve_seeded(const mcbh_problem* p, int device, uint64_t seed, int64_t n_begin, int64_t n_end, double* out, mcb_stats* stats) {
    MCBH_TRY(
        FieldProblem::device(device);
        ArrayXXd sol = p->prob->solveSeeded(seed, n_begin, n_end, nullptr);
        std::memcpy(out, sol.data(), sizeof(double) * (size_t)sol.size());
        if (stats) *stats = FieldProblem::lastStats();
        return MCB_OK;
    , MCB_EINVAL)
}

// process-wide device selection (n = 0: every visible sm_100 device); returns the number selected
int mcbh_set_devices(const int* ordinals, int n) {
    MCBH_TRY(
        FieldProblem::devices(std::vector<int>(ordinals, ordinals + (n > 0 ? n : 0)));
        return (int)FieldProblem::devices().size();
    , -1)
}
// The reference's calling pattern (main.cpp:155-166): solve() entered by every thread of an OpenMP region with its own
// generator (mt19937(mt_seed + thread)), the partial fields summed under `omp critical`.  Must equal ONE solve.
int mcbh_problem_solve_omp(const mcbh_problem* p, uint32_t mt_seed, int nthreads, double* out, mcb_stats* stats, int64_t* progress_count) {
    return solve_like_the_reference(p, mt_seed, nthreads, out, stats, progress_count);
}

// TriangularPrismImpl::cellVol (cell = MCB_CELL_TRIPRISM) / TetrahedronImpl::cellVol (MCB_CELL_TETRAHEDRON)
double mcbh_simplex_cell_vol(int cell, const int64_t index[3], const int64_t shape[3], double vol) {
    const Vector3l i(index[0], index[1], index[2]), s(shape[0], shape[1], shape[2]);
    return cell == MCB_CELL_TRIPRISM ? TriangularPrismImpl::cellVol(i, s, vol) : TetrahedronImpl::cellVol(i, s, vol);
}

// TrajProblem(mat, dom, [prop], [pos], [dir], maxscat, maxloop).solve(mt19937(mt_seed)): polyline out (3 x npoints,
// column-major); prints the per-trip lines to stdout like the reference.  Returns npoints or < 0.
int64_t mcbh_traj(const Material* mat, const mcbh_domain* dom, int has_prop, int64_t w, int64_t p, int has_pos, const double* pos,
                  int has_dir, const double* dir, int64_t maxscat, int64_t maxloop, int device, uint32_t mt_seed,
                  double* points, int64_t max_points) {
    MCBH_TRY(
        FieldProblem::device(device);
        const Domain* D = dom->dom.get();
        std::unique_ptr<TrajProblem> tp;
        TrajProblem::Prop pr(w, p);
        Vector3d P = has_pos ? Vector3d(pos[0], pos[1], pos[2]) : Vector3d(); Vector3d Dv = has_dir ? Vector3d(dir[0], dir[1], dir[2]) : Vector3d();
        if (has_prop && has_pos && has_dir) tp.reset(new TrajProblem(mat, D, pr, P, Dv, maxscat, maxloop));
        else if (has_prop && has_pos) tp.reset(new TrajProblem(mat, D, pr, P, maxscat, maxloop));
        else if (has_pos && has_dir) tp.reset(new TrajProblem(mat, D, P, Dv, maxscat, maxloop));
        else if (has_pos) tp.reset(new TrajProblem(mat, D, P, maxscat, maxloop));
        else tp.reset(new TrajProblem(mat, D, maxscat, maxloop));
        Rng gen(mt_seed);
        Progress prog = tp->initProgress();
        ArrayXXd sol = tp->solve(gen, &prog);
        const int64_t n = std::min<int64_t>(sol.cols(), max_points);
        std::memcpy(points, sol.data(), sizeof(double) * 3 * (size_t)n);
        return (int64_t)sol.cols();
    , -1)
}

} // extern "C"
