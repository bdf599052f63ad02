// problem.h — host mirror of Problem / FieldProblem and its five tallies (problem.h:29-239).
// FieldProblem::solve keeps the reference signature; its body flattens the const Material/Domain it was
// built on, uploads them once per device context, and runs the loop on the GPU through include/mcb.h.
#ifndef MCB_HOST_PROBLEM_H
#define MCB_HOST_PROBLEM_H
#include <ctime>
#include <iosfwd>
#include <optional>
#include <string>
#include <vector>
#include "../../include/mcb.h"
#include "mc_types.h"

class Clock {
    std::time_t start_;
public:
    Clock();
    std::string stopwatch();
    static std::string timestamp();
};

class Progress {
    long tot_, count_, div_, next_, esc_;
    std::vector<long> vec_;
    Clock clk_;
public:
    Progress();
    Progress(long tot, long div);
    void clock(const Clock& clk) { clk_ = clk; }
    long count() const { return count_; }
    long esc() const { return esc_; }
    long incrCount();
    long incrEsc();
    // the device finishes particles in bulk: advance by n finished particles / e escapes at once,
    // printing the same 20-tick bar lines the reference prints one particle at a time
    void advance(long n, long e);
};

class Material;
class Domain;

class Problem {
    const Material* mat_;
    const Domain* dom_;
protected:
    virtual std::string info() const;
public:
    Problem();
    Problem(const Material* mat, const Domain* dom);
    virtual ~Problem();
    const Material* mat() const { return mat_; }
    const Domain* dom() const { return dom_; }
    virtual Progress initProgress() const = 0;
    virtual ArrayXXd solve(Rng& gen, Progress* prog) const = 0;
    friend std::ostream& operator<<(std::ostream& os, const Problem& prob);
};

// Single-particle trajectory (problem.h:86-119, problem.cpp:163-299): the geometry self-check behind the
// `traj` and `check` CLI modes.  solve() prints one `sdom: bdry type -> bdry type` line per loop trip and returns
// the 3 x N polyline TrkPhonon records; the trace itself runs on the device (mcb_traj).
class TrajProblem : public Problem {
public:
    struct Prop { long w, p; Prop(long w_ = 0, long p_ = 0) : w(w_), p(p_) {} };     // Phonon::Prop (phonon.h:21-33)
private:
    static const long loopFactor_ = 100;
    long maxscat_, maxloop_;
    std::optional<Prop> prop_;
    std::optional<Vector3d> pos_, dir_;
    std::string info() const;
public:
    TrajProblem();
    TrajProblem(const Material* mat, const Domain* dom, const Prop& prop, const Vector3d& pos, const Vector3d& dir,
                long maxscat = 100, long maxloop = 0);
    TrajProblem(const Material* mat, const Domain* dom, const Prop& prop, const Vector3d& pos, long maxscat = 100, long maxloop = 0);
    TrajProblem(const Material* mat, const Domain* dom, const Vector3d& pos, const Vector3d& dir, long maxscat = 100, long maxloop = 0);
    TrajProblem(const Material* mat, const Domain* dom, const Vector3d& pos, long maxscat = 100, long maxloop = 0);
    TrajProblem(const Material* mat, const Domain* dom, long maxscat = 100, long maxloop = 0);
    Progress initProgress() const;
    ArrayXXd solve(Rng& gen, Progress* prog) const;
};

// Flattened Domain (what mcb_upload_domain consumes); storage owned here.
struct FlatDomain {
    std::vector<mcb_sdom_desc> sdoms;
    std::vector<mcb_plane_desc> planes;
    std::vector<int32_t> pairs;
    std::vector<mcb_emitter_desc> emitters;
    std::vector<double> cell_vol;       // Field(1, dom, CellVolF()): Subdomain::cellVol per column
    long cols;
    mcb_domain_desc desc() const;
};
FlatDomain flattenDomain(const Domain* dom);

class FieldProblem : public Problem {
    static const long loopFactor_ = 100;
    long nemit_, maxscat_, maxloop_;
    double power_;
    VectorXl emitPdf_;
protected:
    std::string info() const;
public:
    FieldProblem();
    FieldProblem(const Material* mat, const Domain* dom, long nemit, long maxscat, long maxloop);
    virtual ~FieldProblem();

    Progress initProgress() const;
    ArrayXXd initSolution() const;
    ArrayXXd solve(Rng& gen, Progress* prog) const;

    long nemit() const { return nemit_; }
    long maxscat() const { return maxscat_; }
    long maxloop() const { return maxloop_; }
    double power() const { return power_; }
    const VectorXl& emitPdf() const { return emitPdf_; }
    mcb_problem_desc desc() const;                      // pointers valid for this object's lifetime
    // device selection / counters of the latest solve on this thread (not in the reference).  The selection is
    // process-wide: solve() shards the particle range over every selected device (one host thread + one mcb context per
    // device), sums the raw tallies with NCCL (mcb_allreduce) and normalises once.
    static void device(int ordinal);                              // one device
    static void devices(const std::vector<int>& ordinals);        // several; empty = every visible sm_100 device
    static std::vector<int> devices();
    static int deviceCount();
    static mcb_stats lastStats();
    ArrayXXd solveSeeded(unsigned long long seed, long n_begin, long n_end, Progress* prog) const;

private:
    virtual long rows() const = 0;
    virtual int kind() const = 0;                       // MCB_PROB_*
    virtual long size() const { return 0; }
    virtual long step() const { return 0; }
    mutable std::vector<int64_t> emit64_;
};

// Field(1, dom, CellVolF()): the cell volumes the normalisation divides by (problem.h:151-155, problem.cpp:302-306, 441-442)
class Subdomain;
class CellVolF {
public:
    VectorXd operator()(const Subdomain* sdom, const Vector3l& index) const;
};

#define MCB_DECLARE_FIELD_PROBLEM(Name)                                                              \
    class Name : public FieldProblem {                                                              \
        std::string info() const;                                                                   \
    public:                                                                                         \
        Name();                                                                                     \
        Name(const Material* mat, const Domain* dom, long nemit, long maxscat, long maxloop = 0);   \
    private:                                                                                        \
        long rows() const;                                                                          \
        int kind() const;                                                                           \
    };
MCB_DECLARE_FIELD_PROBLEM(TempProblem)
MCB_DECLARE_FIELD_PROBLEM(FluxProblem)
MCB_DECLARE_FIELD_PROBLEM(MultiProblem)
#undef MCB_DECLARE_FIELD_PROBLEM

#define MCB_DECLARE_CUM_PROBLEM(Name)                                                                       \
    class Name : public FieldProblem {                                                                     \
        long size_, step_;                                                                                 \
        std::string info() const;                                                                          \
    public:                                                                                                \
        Name();                                                                                            \
        Name(const Material* mat, const Domain* dom, long nemit, long size, long maxscat, long maxloop = 0);\
    private:                                                                                               \
        long rows() const;                                                                                 \
        int kind() const;                                                                                  \
        long size() const { return size_; }                                                                \
        long step() const { return step_; }                                                                \
    };
MCB_DECLARE_CUM_PROBLEM(CumTempProblem)
MCB_DECLARE_CUM_PROBLEM(CumFluxProblem)
#undef MCB_DECLARE_CUM_PROBLEM
#endif
