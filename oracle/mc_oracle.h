/*
 * mc_oracle.h — C entry points of the CPU ORACLE (test infrastructure, NOT product).
 *
 * The oracle is a dependency-free C++17 restatement of the reference algorithm for
 * the hot path (FieldProblem::solve and everything it calls).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (libmcb.so) never links or calls it.
 *
 * PARITY PINNED AGAINST THE REFERENCE ITSELF: the reference ships no tests, golden
 * vectors or fixtures, but its own sources are compiled here (oracle/Makefile
 * target `ref`, against the Eigen/Boost stand-ins of oracle/shim/) into oracle/_ref/, and
 * this restatement reproduces that binary's output word for word on the same mt19937
 * stream (tests/test_reference_pin.py, tests/golden/ref_*.json), besides the known-answer
 * tests of tests/test_oracle_cpu.py.
 *
 * Descriptors are the structs of include/mcb.h so the same tables can be handed to
 * the CUDA library.
 */
#ifndef MC_ORACLE_H_
#define MC_ORACLE_H_

#include <stdint.h>
#include "../include/mcb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_material orc_material;
typedef struct orc_domain   orc_domain;
typedef struct orc_problem  orc_problem;

#define ORC_RNG_MT19937 0   /* reference-faithful stream consumption, one engine per thread */
#define ORC_RNG_PHILOX  1   /* counter-based, keyed by (seed, particle id): the GPU's stream */

const char* orc_last_error(void);

/* Material::Material(disp, relax, temp)  material.cpp:82-162 */
orc_material* orc_material_load(const char* disp_path, const char* relax_path, double temp);
void   orc_material_free(orc_material*);
int    orc_material_desc(const orc_material*, mcb_material_desc* out); /* pointers into the oracle */
double orc_material_cond(const orc_material*);                         /* Material::cond() */
int    orc_material_alias(const orc_material*, int which, double* wprob, int32_t* walias,
                          double* pprob, int32_t* palias);

/* Shipped domains (domain.cpp): kind in {"bulk","film","jct","tee","tube","hex","pyr"}; dim/div are the
 * constructor vectors (NOT the CLI shorthand).  hex/pyr take no divisions (ndiv = 0). */
orc_domain* orc_domain_create(const char* kind, const double* dim, int ndim,
                              const int64_t* div, int ndiv, double dT);
/* One Parallelepiped<...> with arbitrary boundary kinds (MCB_BDRY_*), wall temperatures
 * T[6] and volumetric gradT (subdomain.h:145-158); Peri faces are paired 0<->3, 1<->4,
 * 2<->5 by pure translation (as BulkDomain::init does, domain.cpp:137-141). */
orc_domain* orc_domain_box(const double origin[3], const double mat[9], const int64_t div[3],
                           const double grad_t[3], const int32_t kinds[6], const double T[6]);
/* One non-box cell (tri-prism, tetrahedron, prism, pyramid: subdomain.h:200-630) with Spec/Diff/Isot faces. */
orc_domain* orc_domain_cell(int cell, const double origin[3], const double* cols, int ncols, const int64_t div[3],
                            const double grad_t[3], const int32_t* kinds, const double* T);
void   orc_domain_free(orc_domain*);
int    orc_domain_desc(const orc_domain*, mcb_domain_desc* out);
int64_t orc_domain_cols(const orc_domain*);
/* cell volumes per column: Field(1, dom, CellVolF())  problem.cpp:302-306,441-442 */
int    orc_domain_cell_vol(const orc_domain*, double* vol);

/* FieldProblem subclasses  problem.cpp:315-342, 451-648 */
orc_problem* orc_problem_create(const orc_material*, const orc_domain*, int kind,
                                int64_t nemit, int64_t size, int64_t maxscat, int64_t maxloop);
void   orc_problem_free(orc_problem*);
int    orc_problem_desc(const orc_problem*, mcb_problem_desc* out);

/* FieldProblem::solve for n in [n_begin,n_end) inside an OpenMP parallel region of
 * `nthreads` threads + the critical-section sum of main.cpp:151-166.  out_field rows x cols. */
int    orc_solve(const orc_problem*, int rng_mode, uint64_t seed,
                 int64_t n_begin, int64_t n_end, int nthreads,
                 double* out_field, mcb_stats* stats);
/* The raw tally (before postProc / volume / power). */
int    orc_solve_raw(const orc_problem*, int rng_mode, uint64_t seed,
                     int64_t n_begin, int64_t n_end, int nthreads,
                     double* raw_field, mcb_stats* stats);
/* problem.cpp:439-444 on a raw tally, in place */
int    orc_finalize(const orc_problem*, double* field);

/* Philox mode: state of particles [n_begin,n_end) after emission + nsteps loop trips. */
int    orc_trace(const orc_problem*, uint64_t seed, int64_t n_begin, int64_t n_end,
                 int64_t nsteps, mcb_trace_out* out);

/* TrajProblem::solve (problem.cpp:226-299), Philox word source (particle id 0): same records as mcb_traj */
int    orc_traj(const orc_material*, const orc_domain*, const mcb_traj_desc*, uint64_t seed, mcb_traj_out* out);
int    orc_traj_rng(const orc_material*, const orc_domain*, const mcb_traj_desc*, int rng_mode, uint64_t seed, mcb_traj_out* out);
/* Domain::locate (domain.cpp:59-67): index of the first subdomain containing pos, -1 if none */
int    orc_domain_locate(const orc_domain*, const double pos[3]);

int    orc_cell_index(const orc_domain*, int64_t n, const double* pos, const int32_t* sdom,
                      int64_t* index);
int    orc_accumulate(const orc_domain*, int32_t rows, int64_t n, const int32_t* sdom,
                      const double* bpos, const double* epos, const double* amount,
                      double* field);
void   orc_philox_words(uint64_t seed, uint64_t particle, uint32_t event, uint32_t block,
                        uint32_t out[4]);
/* first `n` doubles of uniform_01 / uniform(-1,1) / uniform_int(0,m-1) from mt19937(seed),
 * to pin the Boost-equivalent distributions in tests. which: 0,1,2 */
void   orc_mt_draws(uint32_t seed, int which, int64_t m, int64_t n, double* out);

int    orc_max_threads(void);
/* evaluation order of the three draws in `Vector3d coord(dist(gen), dist(gen), dist(gen))` (subdomain.cpp:279, :312, :354):
 * 0 (default) = left to right, 1 = right to left (g++'s choice when it compiles the reference into oracle/_ref) */
void   orc_set_arg_order(int right_to_left);

#ifdef __cplusplus
}
#endif
#endif
