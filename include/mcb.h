/*
 * mcb.h — C ABI of the B200-native phonon Monte Carlo hot path.
 *
 * This is the drop-in boundary for ONE path of nickdou/montecarlocpp:
 * the body of `FieldProblem::solve` (reference montecarlo/problem.cpp:370-445),
 * i.e. emit -> drift -> boundary interaction -> intrinsic scattering -> cell tally.
 *
 * The reference has no FFI layer; its seam is the pure virtual
 *     virtual ArrayXXd Problem::solve(Rng& gen, Progress* prog) const = 0;   (problem.h:83)
 * plus the getters that loop consumes.  A binding therefore flattens the const
 * objects the loop reads into the POD descriptors below and calls mcb_solve().
 * Every entry point names the reference interface it replaces (file:line relative
 * to /root/reference/montecarlo/).
 *
 * Conventions
 *   - plain C, no C++/torch types; all pointers are HOST pointers unless a name
 *     ends in `_dev`; the library copies what it needs during the call.
 *   - matrices are 3x3 column-major (Eigen default): m[r + 3*c].
 *   - 2-D tables are column-major like Eigen ArrayXXd(rows, cols): a[r + rows*c].
 *   - every function returns 0 on success, a negative MCB_E* code on failure;
 *     mcb_last_error() gives a message.  There is NO CPU fallback: with no usable
 *     CUDA device mcb_create() fails with MCB_ENODEVICE.
 */
#ifndef MCB_H_
#define MCB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB_ABI_VERSION 2

/* error codes */
#define MCB_OK          0
#define MCB_EINVAL     -1   /* bad argument / inconsistent descriptor          */
#define MCB_ENODEVICE  -2   /* no CUDA device / wrong architecture             */
#define MCB_ECUDA      -3   /* CUDA runtime error (message has the detail)     */
#define MCB_ESTATE     -4   /* call order violated (e.g. solve before upload)  */
#define MCB_ELIMIT     -5   /* table too large for the shared-memory staging   */

/* boundary kinds — Boundary::type() (boundary.cpp:278,303,344,448,509) */
#define MCB_BDRY_SPEC   0   /* SpecBoundary::scatter   boundary.cpp:283-287 */
#define MCB_BDRY_DIFF   1   /* DiffBoundary::scatter   boundary.cpp:308-312 */
#define MCB_BDRY_INTER  2   /* InterBoundary::scatter  boundary.cpp:349-359 */
#define MCB_BDRY_ISOT   3   /* IsotBoundary::scatter   boundary.cpp:455-460 */
#define MCB_BDRY_PERI   4   /* PeriBoundary::scatter   boundary.cpp:516-522 */

/* emitting-surface shapes — Boundary::Shape (boundary.cpp:119-258) */
#define MCB_SHAPE_NONE          0
#define MCB_SHAPE_PARALLELOGRAM 1   /* boundary.cpp:147-152 */
#define MCB_SHAPE_TRIANGLE      2   /* boundary.cpp:182-187 */
#define MCB_SHAPE_POLYGON       3   /* boundary.cpp:243-251; nvert = N-1 fan vertices */

/* cell (subdomain) kinds — *Impl::drawPos / cellVol (subdomain.cpp:269-441) */
#define MCB_CELL_PARALLELEPIPED 0   /* subdomain.cpp:269-281 */
#define MCB_CELL_TRIPRISM       1   /* subdomain.cpp:283-320 */
#define MCB_CELL_TETRAHEDRON    2   /* subdomain.cpp:322-377 */
#define MCB_CELL_PRISM          3   /* subdomain.cpp:379-409 */
#define MCB_CELL_PYRAMID        4   /* subdomain.cpp:411-441 */

/* emitter kinds — Emitter (boundary.h:186-201) */
#define MCB_EMIT_SDOM   0   /* EmitSubdomain  subdomain.cpp:228-263 */
#define MCB_EMIT_BDRY   1   /* EmitBoundary   boundary.cpp:387-431  */

/* problem kinds — FieldProblem subclasses (problem.cpp:451-648) */
#define MCB_PROB_TEMP     0
#define MCB_PROB_FLUX     1
#define MCB_PROB_MULTI    2
#define MCB_PROB_CUMTEMP  3
#define MCB_PROB_CUMFLUX  4

#define MCB_MAX_VERTS   8   /* Polygon<9> has 8 fan vertices (boundary.h:116) */
#define MCB_MAX_BASE    9   /* Prism/Pyramid base edge vectors (subdomain.h)  */

/*
 * Material tables — what Material (material.h:23-68) holds after its constructor
 * (material.cpp:82-162).  vel/tau replace Material::vel()/tau() (material.cpp:174-184);
 * flux_pdf/scat_pdf are the weights behind fluxDist_/scatDist_ (material.cpp:151-158),
 * from which the library rebuilds the two-level Walker-alias sampler of
 * Material::Dist (material.cpp:51-75).
 */
typedef struct mcb_material_desc {
    int64_t nw, np;            /* frequency bins, polarisations                      */
    double  temp;              /* Material::temp()                                   */
    const double* vel;         /* [nw*np] group velocity  vel_(w,p)                  */
    const double* tau;         /* [nw*np] relaxation time tau_(w,p)                  */
    const double* flux_pdf;    /* [nw*np] vel*dedT*dos*domega  (emission sampler)    */
    const double* scat_pdf;    /* [nw*np] dedT*dos*domega/tau  (scattering sampler)  */
    double  energy_sum;        /* Material::energySum()  material.cpp:186-189        */
    double  flux_sum;          /* Material::fluxSum()    material.cpp:191-194        */
    double  scat_sum;          /* Material::scatSum()    material.cpp:195-198        */
} mcb_material_desc;

/*
 * One boundary plane — Boundary (boundary.h:34-63) + the per-kind members of its
 * subclasses.  Normals point INTO the owning cell (subdomain.h:149-154).
 */
typedef struct mcb_plane_desc {
    double  normal[3];         /* Boundary::normal()  boundary.cpp:87-90             */
    double  offset;            /* Boundary::offset()  boundary.cpp:92-95             */
    int32_t kind;              /* MCB_BDRY_*                                          */
    int32_t sdom;              /* owner: index of Boundary::sdom() in sdomPtrs()      */
    int32_t pair_begin;        /* Inter: pairs_ (boundary.h:166); Peri: pair_ (:248)  */
    int32_t pair_count;        /*   -> range in mcb_domain_desc.pairs (plane ids)     */
    double  rot[9];            /* rotMatrix(normal): Diff::rot_ / EmitBoundary::rot_  */
    double  peri_rot[9];       /* PeriBoundary::rot_    (boundary.cpp:543-544)        */
    double  peri_transl[3];    /* PeriBoundary::transl_ (boundary.cpp:545-546)        */
    double  T;                 /* EmitBoundary::T_ AFTER makePair (boundary.cpp:538)  */
    double  origin[3];         /* EmitBoundary::o_                                    */
    int32_t shape;             /* MCB_SHAPE_*                                         */
    int32_t nvert;             /* 2 for parallelogram/triangle (i_, j_), N-1 polygon  */
    double  verts[3 * MCB_MAX_VERTS];  /* column v at verts[3*v .. 3*v+2]             */
} mcb_plane_desc;

/*
 * One subdomain — Subdomain / EmitSubdomain (subdomain.h:35-126).
 */
typedef struct mcb_sdom_desc {
    double  origin[3];         /* o_                                                  */
    double  mat[9];            /* mat_  (edge vectors as columns)                     */
    double  inv[9];            /* inv_ = mat_.inverse()  subdomain.cpp:43             */
    int64_t div[3];            /* div_                                                */
    int64_t shape[3];          /* shape_  subdomain.cpp:44,51                         */
    int64_t max[3];            /* max_    subdomain.cpp:44,52                         */
    int32_t accum;             /* accumFlag(): -2,-1,0,1,2,3,4  subdomain.cpp:47-70   */
    int32_t cell;              /* MCB_CELL_*                                          */
    double  eps;               /* eps_  subdomain.cpp:45                              */
    double  vol;               /* vol()                                               */
    double  grad_t[3];         /* EmitSubdomain::gradT_                               */
    double  emit_rot[9];       /* EmitSubdomain::rot_ = rotMatrix(gradT.normalized()) */
    int32_t plane_begin;       /* bdryPtrs() -> range in mcb_domain_desc.planes       */
    int32_t plane_count;
    int32_t nbase;             /* Prism/Pyramid: N = columns of mat_ (else 0)         */
    int32_t pad_;
    double  base[3 * MCB_MAX_BASE]; /* Prism/Pyramid mat_ columns: col 0 = axis/apex  */
                               /* vector, cols 1..N-1 = base fan (subdomain.h:378,501)*/
} mcb_sdom_desc;

/* One entry of Domain::emitPtrs() (domain.cpp:88-102), in that order. */
typedef struct mcb_emitter_desc {
    int32_t kind;              /* MCB_EMIT_SDOM | MCB_EMIT_BDRY                       */
    int32_t index;             /* sdom id, or plane id                                */
    double  weight;            /* Emitter::emitWeight()                               */
} mcb_emitter_desc;

typedef struct mcb_domain_desc {
    int32_t nsdom;   const mcb_sdom_desc*    sdoms;     /* Domain::sdomPtrs() order   */
    int32_t nplane;  const mcb_plane_desc*   planes;
    int32_t npair;   const int32_t*          pairs;     /* plane ids                  */
    int32_t nemitter;const mcb_emitter_desc* emitters;  /* Domain::emitPtrs() order   */
    /* Field(1, dom, CellVolF()).data().row(0) (problem.cpp:302-306, 441-442): Subdomain::cellVol of every field column,
     * in Field::init order.  May be NULL when every gridded subdomain is a parallelepiped (vol / shape.prod()). */
    int64_t ncols;   const double*           cell_vol;
} mcb_domain_desc;

/*
 * FieldProblem members (problem.h:121-147) after the constructor (problem.cpp:315-342).
 */
typedef struct mcb_problem_desc {
    int32_t kind;              /* MCB_PROB_*                                          */
    int32_t rows;              /* rows()  problem.cpp:468,501,534,576,624             */
    int64_t size;              /* Cum*: size_ (else 0)                                */
    int64_t step;              /* Cum*: step_  problem.cpp:562-563 (else 0)           */
    int64_t nemit;             /* nemit_ = emitPdf_.sum()  problem.cpp:337            */
    int64_t maxscat;           /* maxscat_                                            */
    int64_t maxloop;           /* maxloop_ (already defaulted, problem.cpp:339); limits of this library: maxloop < 2^32 - 1 (the
                                * loop trip is the 32-bit Philox event counter), maxscat < 2^31, nemit < 2^36 (< 2^32 once
                                * maxloop >= 2^31): the slot state packs pid | loop trip into one 64-bit word            */
    double  power;             /* power_   problem.cpp:341                            */
    const int64_t* emit_count; /* emitPdf_ [nemitter]  problem.cpp:329-335            */
} mcb_problem_desc;

/* Counters replacing Progress (problem.cpp:62-118) for one solve call. */
typedef struct mcb_stats {
    int64_t emitted;           /* particles started (Progress::count())              */
    int64_t steps;             /* trips of the loop body problem.cpp:401-435          */
    int64_t esc;               /* Progress::esc(): escaped or failed Inter hand-off   */
    int64_t launches;          /* CUDA kernel launches issued by this call            */
    int64_t cols;              /* field columns (cells)                               */
    double  device_ms;         /* CUDA-event time of the device work                  */
    double  step_ms;           /* CUDA-event time spent in the move-collide kernel    */
    int64_t step_launches;     /* launches of the move-collide kernel                 */
    int64_t slot_steps;        /* slots visited by the move-collide kernel x S (>= steps) */
    int64_t state_stores;      /* slot state write-backs (each: <= 72 B load + 72 B store) */
    /* the same four counters restricted to the STEADY phase: launches issued while particles were still
     * being emitted, i.e. with a full resident population (the launches a streaming roofline refers to) */
    int64_t steady_launches;
    int64_t steady_steps;
    int64_t steady_stores;
    double  steady_ms;
    int64_t compactions;       /* K3 passes over the survivors of the decay phase ... (ABI version 2) */
    int64_t sorts;             /* ... of which were counting sorts by cell (mcb_options::sort_mode)     */
    /* the TAIL: launches that run the last survivors (at most one tile per CTA) to termination; their duration is set by
     * the longest remaining history (a phonon's loop trips are sequential), not by throughput */
    double  tail_ms;
    int64_t tail_steps;
} mcb_stats;

/* Tunables of the device schedule (not part of the physics). 0 = library default.
 * With `slots` and `steps_per_launch` both 0 (the default) every phonon of the solve is kept resident when the two state
 * buffers fit in half of the available device memory, and each launch runs several loop trips per state round trip on a
 * population that only decays; setting either one selects the streaming schedule (resident population refilled while it
 * streams), the one an HBM roofline can be quoted on. */
typedef struct mcb_options {
    int64_t slots;             /* resident particle slots; streaming default: 32 ... 96  */
                               /* tiles of 32 slots per warp (~ a third of the phonons)   */
    int32_t steps_per_launch;  /* S: loop-body trips per state load/store while particles */
                               /* are left to emit (1 = streaming)                        */
    int32_t block;             /* threads per CTA (capped by the kernel's launch bounds)  */
    int32_t ctas_per_sm;       /* persistent grid = ctas_per_sm * SM count (default 1:   */
                               /* the tables + histograms fill one CTA's shared memory)  */
    int32_t tally_mode;        /* 0 auto, 1 shared-memory histograms (1-D difference     */
                               /* arrays / one per CTA), 2 global field (fp64 RED in L2), */
                               /* 3 one histogram per CTA                                 */
    int32_t decay_mode;        /* 0: once nothing is left to emit choose S per launch    */
                               /*    from the measured termination rate; every launch     */
                               /*    stores its survivors densely (compaction fused into  */
                               /*    the step kernel); the last survivors run to the end; */
                               /* 1: keep steps_per_launch throughout, separate           */
                               /*    compaction passes;                                   */
                               /* 2: like 0 with separate compaction passes (round-2      */
                               /*    schedule, kept for A/B)                              */
    int32_t emit_mode;         /* must be 0 (emission is fused into the step kernel)      */
    int32_t compact_pct;       /* decay phase: compact the survivors when fewer than this */
                               /* percentage of the visited slots is live (0 = default)  */
    int32_t sort_mode;         /* K3 with a key: 0 = unordered compaction (default);      */
                               /* 1 = every compaction is a counting sort of the          */
                               /*     survivors by (subdomain, tally cell), i.e. by the   */
                               /*     field column their position lies in;                */
                               /* k >= 2 = sort whenever the population has shrunk by a   */
                               /*     factor k since the last sort, plain compaction      */
                               /*     in between.  Results do not depend on the order     */
                               /*     of the slots (the RNG stream is keyed by particle). */
                               /* A sort is its own pass (decay_mode 0 then behaves as 2) */
    int32_t decay_pct;         /* decay phase: share (%) of the live phonons that may     */
                               /* terminate per launch when S is chosen (0 = default: 12  */
                               /* with the fused compaction, 30 with separate passes)     */
} mcb_options;                 /* sort_mode, decay_pct: ABI version 2 */

typedef struct mcb_ctx mcb_ctx;

/* Lifetime.  `device` is a CUDA ordinal.  Fails (no fallback) if it is not sm_100. */
int  mcb_create(int device, mcb_ctx** out);
void mcb_destroy(mcb_ctx* ctx);
const char* mcb_last_error(const mcb_ctx* ctx);   /* ctx may be NULL: last create() error */
int  mcb_abi_version(void);
/* "<sha1 of the CUDA sources>|<extra compile flags>": lets tests refuse a library that was not built from the tree it sits
 * in, or that was built with experiment flags (no reference analogue) */
const char* mcb_build_info(void);

int  mcb_set_options(mcb_ctx* ctx, const mcb_options* opt);
int  mcb_get_options(const mcb_ctx* ctx, mcb_options* opt);

/* Replaces the reads of Material (material.h:54-64) inside the loop. */
int  mcb_upload_material(mcb_ctx* ctx, const mcb_material_desc* mat);
/* Replaces the reads of Domain/Subdomain/Boundary (domain.h:65-66, subdomain.h:61-75,
 * boundary.h:51-62) inside the loop. */
int  mcb_upload_domain(mcb_ctx* ctx, const mcb_domain_desc* dom);

/* Number of field columns: Field::init (field.cpp:25-45). */
int  mcb_field_cols(const mcb_ctx* ctx, int64_t* cols);

/*
 * The hot path: FieldProblem::solve (problem.cpp:370-445) for particles
 * n in [n_begin, n_end) of [0, nemit).  `out_field` is rows x cols column-major,
 * already post-processed, divided by cell volume and multiplied by power_
 * (problem.cpp:439-444), exactly what solve() returns; partial ranges add up
 * linearly like the per-thread partials at main.cpp:162-165.
 * The RNG is Philox4x32-10 keyed by (seed, particle id) — see DESIGN.md.
 */
int  mcb_solve(mcb_ctx* ctx, const mcb_problem_desc* prob, uint64_t seed,
               int64_t n_begin, int64_t n_end,
               double* out_field, mcb_stats* stats);

/*
 * Same, but leaves the RAW tally (sum of sign*amount per cell, before
 * postProc / volume / power) in DEVICE memory at `raw_field_dev` (rows x cols doubles,
 * caller-allocated, zeroed by the caller or accumulated into).  Used by multi-GPU
 * callers that all-reduce the raw tally over NCCL before finalising.
 */
int  mcb_solve_raw_dev(mcb_ctx* ctx, const mcb_problem_desc* prob, uint64_t seed,
                       int64_t n_begin, int64_t n_end,
                       double* raw_field_dev, mcb_stats* stats);

/* problem.cpp:439-444 on a raw device tally (in place): postProc, /cellVol, *power_. */
int  mcb_finalize_dev(mcb_ctx* ctx, const mcb_problem_desc* prob, double* field_dev);

/* The stream the library launches on (a cudaStream_t, created cudaStreamNonBlocking).  STREAM CONTRACT: mcb_solve,
 * mcb_solve_raw_dev and mcb_solve_raw return after synchronising this stream, so the raw tally is complete when they return;
 * mcb_finalize_dev launches k_finalize on this stream and synchronises it.  Work a caller puts on ANOTHER stream between the
 * two (e.g. an NCCL all-reduce of the raw tally) is NOT ordered before k_finalize: issue it on this stream, or synchronise /
 * cudaStreamWaitEvent it before calling mcb_finalize_dev (bench.py and mcb_allreduce use the library's own stream). */
int  mcb_stream(const mcb_ctx* ctx, void** stream);

/* ---- several GPUs in one process (C++ callers; replaces the thread fan-out + `omp critical` sum of main.cpp:155-166).
 * Phonons are independent histories: context g (one per device) solves its particle range with mcb_solve_raw, which
 * leaves the RAW tally in the context's own device buffer; mcb_allreduce sums those buffers in place over NVLink
 * (ncclAllReduce, ncclDouble, ncclSum on the contexts' streams; NCCL is bound with dlopen at the first call); mcb_finalize
 * normalises (problem.cpp:439-444) one context's buffer and copies it to the host.  The Philox key is the global particle
 * id, so the result does not depend on the number of devices up to fp summation order. */
int  mcb_device_count(int* count);                 /* leading sm_100 devices visible to this process */
int  mcb_solve_raw(mcb_ctx* ctx, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end, mcb_stats* stats);
int  mcb_allreduce(mcb_ctx* const* ctxs, int n, const mcb_problem_desc* prob);
int  mcb_finalize(mcb_ctx* ctx, const mcb_problem_desc* prob, double* out_field);

/* ---------------------------------------------------------------- diagnostics ---
 * Per-particle state after emission and `nsteps` loop trips, for particles
 * [n_begin, n_end): the integer-parity probe of SURVEY §8c KA6.  Arrays are length
 * (n_end-n_begin); pos/dir are 3 per particle (xyz interleaved).  No tally.
 */
typedef struct mcb_trace_out {
    double*  pos;        /* [3n] */
    double*  dir;        /* [3n] */
    double*  scat_next;  /* [n]  */
    int64_t* w;          /* [n]  */
    int64_t* p;          /* [n]  */
    int32_t* sign;       /* [n]  +1 / -1 */
    int32_t* alive;      /* [n]  */
    int32_t* sdom;       /* [n]  */
    int64_t* nscat;      /* [n]  */
    int64_t* steps;      /* [n]  loop trips executed */
    int32_t* cell;       /* [3n] coord2index(coord(pos)) in the final sdom */
} mcb_trace_out;

int  mcb_trace(mcb_ctx* ctx, const mcb_problem_desc* prob, uint64_t seed,
               int64_t n_begin, int64_t n_end, int64_t nsteps, mcb_trace_out* out);

/* ------------------------------------------------------------- trajectories ---
 * TrajProblem::solve (problem.cpp:226-299): ONE particle traced for up to maxloop loop trips, recording the
 * polyline TrkPhonon keeps (phonon.cpp:129-170: start position, the position after every move, and the
 * image position after a periodic wrap) and, per loop trip, the boundary it sat on and the boundary it reached
 * (what the reference prints as `sdom: bdry type -> bdry type`).  Members of TrajProblem (problem.h:89-119):
 * optional prop_, pos_, dir_; maxscat_, maxloop_.  `sdom` is Domain::locate(pos) (domain.cpp:59-67) when has_pos.
 */
typedef struct mcb_traj_desc {
    int32_t has_prop, has_pos, has_dir, sdom;
    int64_t w, p;
    double  pos[3], dir[3];
    int64_t maxscat, maxloop;          /* maxloop already defaulted (100 * maxscat when 0, problem.cpp:262) */
} mcb_traj_desc;

typedef struct mcb_traj_out {
    int64_t  max_points;  double*  points;     /* [3 * max_points] xyz interleaved; needs >= 2*maxloop + 1 */
    int64_t  max_steps;                        /* >= maxloop */
    int32_t* step_sdom;                        /* [max_steps] subdomain index at the start of the trip        */
    int32_t* step_in;                          /* [max_steps] index of the boundary it sits on within that    */
    int32_t* step_in_kind;                     /*             subdomain's bdryPtrs(), -1 = none; MCB_BDRY_*   */
    int32_t* step_out;                         /* [max_steps] boundary reached by advect, -1 = none (scatter) */
    int32_t* step_out_kind;
    int64_t  npoints, nsteps;                  /* filled by the call */
    int32_t  escaped;                          /* 0; 1 = killed inside advect (subdomain.cpp:182-189);   */
                                               /* 2 = failed Inter hand-off (boundary.cpp:357-358)       */
    int32_t  pad_;
} mcb_traj_out;

int  mcb_traj(mcb_ctx* ctx, const mcb_traj_desc* traj, uint64_t seed, mcb_traj_out* out);

/* K3 probe (no reference analogue: the reference's loop `break`s, problem.cpp:411,425,434, and never reorders phonons).
 * Particles [n_begin, n_end) after emission and `nsteps` loop trips (no tally) are compacted -- unordered (sorted = 0) or
 * by the counting sort behind mcb_options::sort_mode (sorted = 1) -- and the resulting slots are described in slot order:
 * bin key, field column of the phonon's cell (Field::init column range of its subdomain + coord2index, field.cpp:25-45,
 * subdomain.cpp:148-159) and particle id, -1 for an inactive slot.  Arrays of length n_end - n_begin; bin = column /
 * *cols_per_bin.  Integer parity probe: the active slots are a permutation of the survivors, contiguous from slot 0, and
 * with sorted = 1 their bins are non-decreasing. */
int  mcb_sort_probe(mcb_ctx* ctx, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end,
                    int64_t nsteps, int32_t sorted, int64_t* bin, int64_t* col, int64_t* pid, int64_t* cols_per_bin);

/* Subdomain::coord + coord2index (subdomain.cpp:148-159) on caller-supplied
 * positions: bit-exact integer parity probe. pos [3n], sdom [n] -> index [3n]. */
int  mcb_cell_index(mcb_ctx* ctx, int64_t n, const double* pos, const int32_t* sdom,
                    int64_t* index);

/* Field::accumulate (field.cpp:92-220) on caller-supplied segments:
 * sdom [n], bpos/epos [3n], amount [rows*n] -> adds into field [rows*cols] (host). */
int  mcb_accumulate(mcb_ctx* ctx, int32_t rows, int64_t n, const int32_t* sdom,
                    const double* bpos, const double* epos, const double* amount,
                    double* field);

/* The Walker-alias tables the library built from a pdf (Material::Dist,
 * material.cpp:51-75): which = 0 flux, 1 scat.  wprob/walias [nw], pprob/palias [nw*np]
 * with entry (w,p) at [w*np + p]. */
int  mcb_get_alias(const mcb_ctx* ctx, int which, double* wprob, int32_t* walias,
                   double* pprob, int32_t* palias);

/* Philox4x32-10 words as the device generates them (key = seed, counter =
 * (particle id, event, block)): out [4] words. */
int  mcb_philox_words(uint64_t seed, uint64_t particle, uint32_t event, uint32_t block,
                      uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MCB_H_ */
