// mcb_kernels.cuh — the CUDA kernels (sm_100a) of the phonon Monte Carlo hot path.
//
//   k_step      K2 with K1 fused in: S trips of the loop body (problem.cpp:401-435) per resident slot: advect -> tally ->
//               boundary or intrinsic scattering; after its last tile every warp refills the slots that ended inactive with the
//               next particles (problem.cpp:386-399).  State streams HBM -> registers -> HBM once per launch (TMA bulk
//               prefetch of the warp's next 2304-B group).  Material + geometry tables are staged into shared memory with a
//               TMA bulk copy (cp.async.bulk + mbarrier).  Tallies: 1-D grids -> fixed-point difference-array histograms in
//               shared memory; N-D grids -> warp-balanced item walk into one fixed-point histogram per CTA, or straight to L2
//               with fp64 RED.  Template modes: payload rows, tally destination, N-D walk (none / serial / serial +
//               cooperative pieces / warp-balanced items), axis-aligned boxes only, padded histogram columns.
//   k_emit      K1 for the first fill (every slot free): one thread per particle; k_emit_commit advances the particle counter.
//               K3 fused in as well: in the decay phase (nothing left to emit) a launch stores its survivors DENSELY into the
//               other state buffer (StepParams::compact; one cursor atomic per tile) and publishes the slot count for the launch
//               behind it -- terminated phonons simply vanish (no reference analogue; replaces the `break`s at
//               problem.cpp:411,425,434).
//   k_compact   K3 as its own pass: unordered stream compaction (mcb_options::decay_mode 1 / 2 only).
//   k_sort_*    K3 with a key: counting sort of the survivors by (subdomain, tally cell) (mcb_options::sort_mode), k_sort_probe
//               behind mcb_sort_probe.
//   k_finalize  K4: postProc, / cellVol, * power_ (problem.cpp:439-444).
//   k_traj      TrajProblem::solve (problem.cpp:226-299) for one particle.
//   k_cell_index / k_accumulate / k_gather_trace / k_philox: diagnostics behind mcb_cell_index / mcb_accumulate /
//               mcb_trace / mcb_philox_words.
#pragma once
#include "mcb_device.cuh"
#include "../../include/mcb.h"

namespace mcb {

// --------------------------------------------------------------- TMA bulk copy + mbarrier (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}

// the same three operations on plain 32-bit shared-memory addresses (k_step's per-warp TMA state prefetch)
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t phase) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void lds2(uint32_t a, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void lds2(uint32_t a, double& x, unsigned long long& y) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=d"(x), "=l"(y) : "r"(a));
}

// streaming state access (warp-tiled layout, mcb_device.cuh: StateView): 128-bit accesses that do not allocate in L1
__device__ __forceinline__ void ld_stream2(const void* p, double& a, double& b) {
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
__device__ __forceinline__ void ld_stream2(const void* p, double& a, unsigned long long& b) {
    asm volatile("ld.global.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=d"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ unsigned long long ld_stream1(const void* p) {
    unsigned long long v; asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ void st_stream2(void* p, double a, double b) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_stream2(void* p, double a, unsigned long long b) {
    asm volatile("st.global.L1::no_allocate.v2.b64 [%0], {%1, %2};" ::"l"(p), "d"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void red_shared_u32(unsigned* p, unsigned v) {          // fire-and-forget shared-memory add
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream1(void* p, unsigned long long v) {
    asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ------------------------------------------------------------------------------ particle
// The integer state stays PACKED in registers exactly as it lies in HBM (meta = mlo | nscat << 32, pidstep = ps): three
// registers instead of nine, no pack / unpack at the load and the store; fields are extracted where they are used.
struct Particle {
    double px, py, pz, dx, dy, dz, sn;
    uint32_t mlo;                 // wp:20 | sign:1 | active:1 | killed:1 | sdom:9   (low word of meta)
    uint32_t nscat;               // high word of meta
    unsigned long long ps;        // pid:36 | step:28
    __device__ __forceinline__ uint32_t wp() const { return mlo & 0xFFFFFu; }
    __device__ __forceinline__ bool sign() const { return (mlo >> 20) & 1u; }
    __device__ __forceinline__ bool active() const { return (mlo >> 21) & 1u; }
    __device__ __forceinline__ bool killed() const { return (mlo >> 22) & 1u; }
    __device__ __forceinline__ uint32_t sdom() const { return mlo >> 23; }
    // pid|step: the split is a per-solve constant (StepParams::step_bits / step_mask, constant-bank operands)
    __device__ __forceinline__ uint32_t step(const StepParams& P) const { return (uint32_t)ps & P.step_mask; }
    __device__ __forceinline__ unsigned long long pid(const StepParams& P) const { return ps >> P.step_bits; }
    __device__ __forceinline__ uint32_t pid_lo(const StepParams& P) const { return (uint32_t)(ps >> P.step_bits); }
    __device__ __forceinline__ uint32_t pid_hi(const StepParams& P) const { return (uint32_t)((ps >> P.step_bits) >> 32); }
    __device__ __forceinline__ void set_wp(uint32_t v) { mlo = (mlo & ~0xFFFFFu) | v; }
    __device__ __forceinline__ void set_sdom(uint32_t v) { mlo = (mlo & 0x7FFFFFu) | (v << 23); }
    __device__ __forceinline__ void stop() { mlo &= ~(1u << 21); }                       // alive but finished (or never started)
    __device__ __forceinline__ void kill() { mlo = (mlo | (1u << 22)) & ~(1u << 21); }  // Phonon::kill phonon.cpp:47-50
    __device__ __forceinline__ void init(uint32_t wp_, bool sign_, bool active_, uint32_t sdom_, unsigned long long pid_, uint32_t step_bits) {
        mlo = wp_ | ((uint32_t)sign_ << 20) | ((uint32_t)active_ << 21) | (sdom_ << 23); nscat = 0; ps = pid_ << step_bits;
    }
    __device__ __forceinline__ unsigned long long meta() const { return (unsigned long long)mlo | ((unsigned long long)nscat << 32); }
    // the slot's bytes: four 16-B vectors at v, v+512, v+1024, v+1536 and the 8-B pid|step word at w (StateView layout)
    static __device__ __forceinline__ unsigned char* vec_ptr(const StateView& st, long long i) {
        return st.base + (i >> 5) * MCB_GROUP_BYTES + (i & 31) * 16;
    }
    static __device__ __forceinline__ unsigned char* word_ptr(const StateView& st, long long i) {
        return st.base + (i >> 5) * MCB_GROUP_BYTES + 2048 + (i & 31) * 8;
    }
    __device__ __forceinline__ unsigned long long load(const StateView& st, long long i) {       // returns the raw meta word
        const unsigned char* v = vec_ptr(st, i);
        unsigned long long m;
        ld_stream2(v, px, py); ld_stream2(v + 512, pz, dx); ld_stream2(v + 1024, dy, dz); ld_stream2(v + 1536, sn, m);
        ps = ld_stream1(word_ptr(st, i));
        mlo = (uint32_t)m; nscat = (uint32_t)(m >> 32);
        return m;
    }
    // the same slot out of a staged copy of its 2304-B group in shared memory (k_step's TMA prefetch): av = address of the
    // lane's first vector (buffer + 16 lane), aw = address of its pid|step word (buffer + 2048 + 8 lane)
    __device__ __forceinline__ void load_shared(uint32_t av, uint32_t aw) {
        unsigned long long m;
        lds2(av, px, py); lds2(av + 512u, pz, dx); lds2(av + 1024u, dy, dz); lds2(av + 1536u, sn, m);
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(ps) : "r"(aw));
        mlo = (uint32_t)m; nscat = (uint32_t)(m >> 32);
    }
    // store to slot `lane` of group g
    __device__ __forceinline__ void store_group(const StateView& st, int g, unsigned lane) const {
        unsigned char* v = st.base + (size_t)g * MCB_GROUP_BYTES + lane * 16u;
        st_stream2(v, px, py); st_stream2(v + 512, pz, dx); st_stream2(v + 1024, dy, dz); st_stream2(v + 1536, sn, meta());
        st_stream1(v + 2048 - lane * 8u, ps);
    }
    __device__ __forceinline__ void store(const StateView& st, long long i) const {
        unsigned char* v = vec_ptr(st, i);
        st_stream2(v, px, py); st_stream2(v + 512, pz, dx); st_stream2(v + 1024, dy, dz); st_stream2(v + 1536, sn, meta());
        st_stream1(word_ptr(st, i), ps);
    }
    static __device__ __forceinline__ unsigned long long load_meta(const StateView& st, long long i) {
        return *reinterpret_cast<const unsigned long long*>(vec_ptr(st, i) + 1536 + 8);
    }
    static __device__ __forceinline__ void clear_meta(const StateView& st, long long i) {          // marks the slot inactive
        *reinterpret_cast<unsigned long long*>(vec_ptr(st, i) + 1536 + 8) = 0ull;
    }
};

// Material::Dist::drawProp (material.cpp:71-75) on the two-level Walker tables; entry (w,p) at w*np+p
template <typename WA, typename PA>
__device__ __forceinline__ uint32_t draw_prop(Rng& g, const Tables& T, const double* wprob, const WA* walias,
                                              const double* pprob, const PA* palias) {
    const uint32_t r = g.uint_below((uint32_t)T.nw, T.inv_bucket_w);
    const double t = g.u01();
    const uint32_t w = t < wprob[r] ? r : (uint32_t)walias[r];
    const uint32_t q = g.uint_below((uint32_t)T.np, T.inv_bucket_p);
    const double u = g.u01();
    const uint32_t p = u < pprob[w * T.np + q] ? q : (uint32_t)palias[w * T.np + q];
    return w * (uint32_t)T.np + p;
}
// Material::drawScatNext (material.cpp:215-224); lambda = vel*tau
__device__ __forceinline__ double draw_scat_next(Rng& g, double lambda) {
    double d = 0.0;
    while (d < 2.2250738585072014e-308) d = lambda * neg_log1m_u32(g.next());     // -log(1 - uniform_01)
    return d;
}

// TriangularPrismImpl::drawPos (subdomain.cpp:309-320) / TetrahedronImpl::drawPos (:351-377): fold the unit cube
// onto the prism / tetrahedron spanned by the three columns at a[0], a[3], a[6]
__device__ __forceinline__ void fold_simplex(bool tet, double& c0, double& c1, double& c2) {
    if (c0 + c1 > 1.0) { c0 = 1.0 - c0; c1 = 1.0 - c1; }
    if (!tet) return;
    if (c1 + c2 > 1.0) { const double t = c2; c2 = 1.0 - c0 - c1; c1 = 1.0 - t; }
    else if (c0 + c1 + c2 > 1.0) { const double t = c2; c2 = c0 + c1 + c2 - 1.0; c0 = 1.0 - c1 - t; }
}
// boost discrete_distribution draw on a small alias table (volDist_ / areaDist_)
__device__ __forceinline__ int draw_small(Rng& g, int n, const double* prob, const int32_t* alias) {
    const uint32_t r = g.uint_below((uint32_t)n);
    const double t = g.u01();
    return t < prob[r] ? (int)r : alias[r];
}

// Emitter::emit (boundary.cpp:378-385): position, direction and sign from one emitter
__device__ __forceinline__ void emit_from(const DEmitter& E, Rng& g, Particle& ph) {
    double px, py, pz, dx, dy, dz; uint32_t sign;
    if (E.kind == MCB_EMIT_SDOM) {
        // *Impl::drawPos subdomain.cpp:275-281, 309-320, 351-377, 401-409, 433-441 ; drawDir :255-258 ; emitSign :260-263
        const double* m = E.a;
        double loc[9];
        if (E.shape == MCB_CELL_PRISM || E.shape == MCB_CELL_PYRAMID) {
            const int ind = draw_small(g, E.nsub, E.sprob, E.salias);      // sub-wedge, then local = (col ind+1, col ind+2, col 0)
            for (int k = 0; k < 3; ++k) { loc[k] = E.a[3 * (ind + 1) + k]; loc[3 + k] = E.a[3 * (ind + 2) + k]; loc[6 + k] = E.a[k]; }
            m = loc;
        }
        double c0 = g.u01(), c1 = g.u01(), c2 = g.u01();
        if (E.shape == MCB_CELL_TRIPRISM || E.shape == MCB_CELL_PRISM) fold_simplex(false, c0, c1, c2);
        else if (E.shape == MCB_CELL_TETRAHEDRON || E.shape == MCB_CELL_PYRAMID) fold_simplex(true, c0, c1, c2);
        double mx, my, mz; matvec(m, c0, c1, c2, mx, my, mz);
        px = E.o[0] + mx; py = E.o[1] + my; pz = E.o[2] + mz;
        double ax, ay, az; draw_aniso(g, true, ax, ay, az);
        matvec(E.rot, ax, ay, az, dx, dy, dz);
        sign = dot3(dx, dy, dz, E.g[0], E.g[1], E.g[2]) < 0.0 ? 1u : 0u;
    } else {
        // EmitBoundary::drawPos / drawDir / emitSign boundary.cpp:418-431 ; shapes :147-152, :182-187, :243-251
        int n = 0;
        if (E.shape == MCB_SHAPE_POLYGON) n = draw_small(g, E.nsub, E.sprob, E.salias);     // fan triangle (verts n, n+1)
        double r1 = g.u01(), r2 = g.u01();
        if (E.shape != MCB_SHAPE_PARALLELOGRAM && !(r1 + r2 < 1.0)) { r1 = 1.0 - r1; r2 = 1.0 - r2; }
        const double* vi = E.a + 3 * n; const double* vj = E.a + 3 * (n + 1);
        px = E.o[0] + (r1 * vi[0] + r2 * vj[0]);
        py = E.o[1] + (r1 * vi[1] + r2 * vj[1]);
        pz = E.o[2] + (r1 * vi[2] + r2 * vj[2]);
        double ax, ay, az; draw_aniso(g, false, ax, ay, az);
        matvec(E.rot, ax, ay, az, dx, dy, dz);
        sign = E.g[0] >= 0.0 ? 1u : 0u;
    }
    normalize3(dx, dy, dz);                                  // Phonon ctor phonon.cpp:33-37
    ph.px = px; ph.py = py; ph.pz = pz; ph.dx = dx; ph.dy = dy; ph.dz = dz;
    ph.init(0u, sign != 0u, false, (uint32_t)E.sdom, 0ull, 0u);       // the caller sets wp, active and the particle id
}

// problem.cpp:386-399 for particle `pid`: pick emitter, drawFluxProp, Emitter::emit, drawScatNext.  Out of line (it is long and
// runs once per particle); its inputs travel BY VALUE: a reference to the kernel parameter struct would force every thread to
// copy the whole struct into local memory.
struct EmitArgs {
    const DEmitter* emitters; const long long* emit_cdf; const double* f_wprob; const double* f_pprob;
    const int32_t* f_walias; const int32_t* f_palias; const double* lambda;
    unsigned long long seed; double inv_bucket_w, inv_bucket_p; int32_t nemitter, nw, np, active; uint32_t step_bits;
};
static __device__ __noinline__ void emit_particle_impl(const EmitArgs a, unsigned long long pid, Particle& ph) {
    // emitter = upper_bound(emitCdf, n)  (problem.cpp:386-387)
    int lo = 0, hi = a.nemitter;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((long long)pid < a.emit_cdf[mid]) hi = mid; else lo = mid + 1; }
    const DEmitter& E = a.emitters[lo];
    Rng g; g.begin(a.seed, pid, 0u);
    Tables T; T.nw = a.nw; T.np = a.np; T.inv_bucket_w = a.inv_bucket_w; T.inv_bucket_p = a.inv_bucket_p;
    const uint32_t wp = draw_prop(g, T, a.f_wprob, a.f_walias, a.f_pprob, a.f_palias);
    emit_from(E, g, ph);
    ph.init(wp, ph.sign(), a.active != 0, ph.sdom(), pid, a.step_bits);
    ph.sn = draw_scat_next(g, a.lambda[wp]);
}
__device__ __forceinline__ void emit_particle(const StepParams& P, const Tables& T, unsigned long long pid, Particle& ph) {
    EmitArgs a;
    a.emitters = P.emitters; a.emit_cdf = P.emit_cdf; a.f_wprob = P.f_wprob; a.f_pprob = P.f_pprob; a.f_walias = P.f_walias; a.f_palias = P.f_palias;
    a.lambda = T.lambda; a.seed = P.seed; a.inv_bucket_w = T.inv_bucket_w; a.inv_bucket_p = T.inv_bucket_p;
    a.nemitter = P.nemitter; a.nw = T.nw; a.np = T.np; a.active = P.maxloop > 0 ? 1 : 0; a.step_bits = P.step_bits;
    emit_particle_impl(a, pid, ph);
}

// Subdomain::isInside subdomain.cpp:108-116.  BOX: every subdomain of the domain is an axis-aligned box (the host checks).
template <bool BOX>
__device__ __forceinline__ bool is_inside(const Tables& T, const DSdom& sd, double x, double y, double z) {
    bool in = true;
    if (BOX || sd.aabb) {            // axis-aligned box: n_b.x is x[b]
        // six signed distances against -eps, OR-ed in one predicate (setp.lt.or): no branches, no min/max detours
        const double neps = -sd.eps;
        const double s0 = x + sd.offl[0], s1 = sd.offh[0] - x, s2 = y + sd.offl[1], s3 = sd.offh[1] - y, s4 = z + sd.offl[2], s5 = sd.offh[2] - z;
        uint32_t out;
        asm("{\n .reg .pred p;\n setp.lt.f64 p, %1, %7;\n setp.lt.or.f64 p, %2, %7, p;\n setp.lt.or.f64 p, %3, %7, p;\n"
            " setp.lt.or.f64 p, %4, %7, p;\n setp.lt.or.f64 p, %5, %7, p;\n setp.lt.or.f64 p, %6, %7, p;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(out) : "d"(s0), "d"(s1), "d"(s2), "d"(s3), "d"(s4), "d"(s5), "d"(neps));
        return out == 0u;
    }
    if (sd.is_box) {                 // planes b and b+3 share n.pos up to sign: three dot products
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const DPlaneHot h = T.hot[sd.plane_begin + b];
            const double s = dot3(h.nx, h.ny, h.nz, x, y, z);
            const double off3 = T.hot[sd.plane_begin + b + 3].off;
            in = in && !(s + h.off < -sd.eps) && !(off3 - s < -sd.eps);
        }
        return in;
    }
    for (int b = 0; b < sd.plane_count; ++b) {
        const DPlaneHot h = T.hot[sd.plane_begin + b];
        in = in && !(dot3(h.nx, h.ny, h.nz, x, y, z) + h.off < -sd.eps);
    }
    return in;
}

struct Segment {                  // what one advect produced
    double bx, by, bz, ex, ey, ez, d;
    int hit; uint32_t nscat_before; bool ok;
    int next_plane;               // set by collide(): the boundary the particle sits on afterwards (-1: none)
};

// First half of a loop trip (problem.cpp:403-412): Subdomain::advect + Phonon::move.  Returns escapes (0/1).
template <bool BOX>
__device__ __forceinline__ uint32_t advect_move(const Tables& T, Particle& ph, Segment& sg) {
    const DSdom& sd = T.sdom[ph.sdom()];
    // Subdomain::advect subdomain.cpp:161-192
    double d = ph.sn; int hit = -1;
    if (BOX || sd.aabb) {
        // Axis-aligned box (every shipped domain): n_b = +e_b, n_{b+3} = -e_b, so n.dir and n.pos are single components
        // -- the very values the dot products produce.  Per axis at most one face is approached: face b when dir_b < 0,
        //   t = -(off_b + pos_b) / dir_b = (off_b + pos_b) / |dir_b|, else face b+3, t = (off_{b+3} - pos_b) / |dir_b|
        // (boundary.cpp:107-110; same operations, the two sign flips cancel exactly).  dir_b = 0 gives NaN or +inf,
        // which loses every comparison.  Candidates are taken in the reference's declaration order (faces 0,1,2,3,4,5)
        // with its strict `<`: among equal distances the lowest face index wins, and a tie with scatNext scatters.
        const double dr[3] = {ph.dx, ph.dy, ph.dz}, ps[3] = {ph.px, ph.py, ph.pz};
        int key = 6;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const bool front = dr[b] < 0.0;
            const double num = front ? sd.offl[b] + ps[b] : sd.offh[b] - ps[b];
            const double t = div_fast(num, fabs(dr[b]));
            const int k = front ? b : b + 3;
            if (t < d || (t == d && k < key)) { d = t; key = k; }
        }
        if (key < 6) hit = sd.plane_begin + key;
    } else if (sd.is_box) {
        // A parallelepiped's faces b and b+3 have exactly opposite normals, so n.dir < 0 holds for at most one
        // of each pair: three divisions, no divergence.  Candidates are then compared in the reference's
        // declaration order (faces 0,1,2 then 3,4,5) with its strict `<`.
        double t3[3]; int id3[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const DPlaneHot h = T.hot[sd.plane_begin + b];
            const double c = dot3(h.nx, h.ny, h.nz, ph.dx, ph.dy, ph.dz);
            const double s = dot3(h.nx, h.ny, h.nz, ph.px, ph.py, ph.pz);
            const double off3 = T.hot[sd.plane_begin + b + 3].off;
            const bool front = c < 0.0;                     // face b is approached; else face b+3 (if c > 0)
            const double num = front ? -(h.off + s) : -(off3 - s);
            const double den = front ? c : -c;
            t3[b] = num / den;                              // boundary.cpp:107-110
            id3[b] = c == 0.0 ? -1 : (front ? b : b + 3);
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) if (id3[b] == b && t3[b] < d) { d = t3[b]; hit = sd.plane_begin + b; }
#pragma unroll
        for (int b = 0; b < 3; ++b) if (id3[b] == b + 3 && t3[b] < d) { d = t3[b]; hit = sd.plane_begin + b + 3; }
    } else {
        for (int b = 0; b < sd.plane_count; ++b) {
            const DPlaneHot h = T.hot[sd.plane_begin + b];
            const double c = dot3(h.nx, h.ny, h.nz, ph.dx, ph.dy, ph.dz);
            if (c >= 0.0) continue;
            const double t = -(h.off + dot3(h.nx, h.ny, h.nz, ph.px, ph.py, ph.pz)) / c;   // boundary.cpp:107-110
            if (t < d) { d = t; hit = sd.plane_begin + b; }
        }
    }
    // Phonon::move phonon.cpp:95-105
    double sn = ph.sn - d;
    if (sn < c_k[27]) sn = 0.0;
    ph.sn = sn;
    sg.bx = ph.px; sg.by = ph.py; sg.bz = ph.pz;
    sg.ex = sg.bx + ph.dx * d; sg.ey = sg.by + ph.dy * d; sg.ez = sg.bz + ph.dz * d;
    ph.px = sg.ex; ph.py = sg.ey; ph.pz = sg.ez;
    sg.d = d; sg.hit = hit; sg.nscat_before = ph.nscat;
    ph.ps++;                                                                   // step is the low field of ps
    if (d < -sd.eps || !is_inside<BOX>(T, sd, sg.ex, sg.ey, sg.ez)) {         // subdomain.cpp:182-189
        ph.kill(); sg.ok = false; return 1u;                                  // problem.cpp:408-412
    }
    sg.ok = true;
    return 0u;
}

// The random events of a loop trip -- Material::scatter (material.cpp:226-231) and DiffBoundary::scatter (boundary.cpp:308-312)
// -- are drawn TOGETHER: both start the event's word sequence with Philox block 0 and both end in "two uniforms -> sqrt ->
// sincos(pi r) -> unit vector" (drawIso / drawAniso, random.cpp:16-44), so the lanes of a warp that scatter intrinsically and
// the lanes that hit a diffuse wall share one pass through the Philox rounds and the sincos kernel instead of taking two
// divergent ones (wire / film: the two events split a warp about evenly).
//  * intrinsic: seven words (w int, w real, p int, p real, iso mu, iso phi, free path) at fixed positions of blocks 0 and 1,
//    valid when no draw is rejected (probability ~ nw / 2^32); a rejection (or a 1-entry table, whose uniform_int draws no
//    word) replays the event through the sequential generator, so the consumed stream is always the CPU oracle's;
//  * diffuse wall: words 0 and 1 of block 0 (aniso r, phi), then the wall's rotation.
// (Drawing the event ahead of the flight, to interleave its dependency chain with advect / tally, was measured:
// -4 ... -14 %, the extra live registers cost more than the overlap gains.)
template <bool BOX>
__device__ __forceinline__ void scatter_draw(const StepParams& P, const Tables& T, Particle& ph, bool intr, const DPlaneCold* cb) {
    uint32_t a[4];
    philox4x32_10_rk(ph.pid_lo(P), ph.pid_hi(P), ph.step(P), 0u, P.rk, a);
    uint32_t w1 = a[0], w2 = a[1];                 // the two words of the direction draw
    bool ok = true;
    uint32_t wp = 0; double dist = 0.0;
    if (intr) {
        ok = T.nw > 1 && T.np > 1;
        if (ok) {
            uint32_t b[4];
            philox4x32_10_rk(ph.pid_lo(P), ph.pid_hi(P), ph.step(P), 1u, P.rk, b);
            uint32_t r = (uint32_t)(((double)a[0] + 0.5) * T.inv_bucket_w);
            uint32_t q = (uint32_t)(((double)a[2] + 0.5) * T.inv_bucket_p);
            ok = r < (uint32_t)T.nw && q < (uint32_t)T.np;
            r = min(r, (uint32_t)T.nw - 1u); q = min(q, (uint32_t)T.np - 1u);
            const uint32_t w = (double)a[1] * c_k[25] < T.wprob[r] ? r : (uint32_t)T.walias[r];
            const uint32_t k = w * (uint32_t)T.np + q;
            wp = (double)a[3] * c_k[25] < T.pprob[k] ? k : w * (uint32_t)T.np + (uint32_t)T.palias[k];
            dist = T.lambda[wp] * neg_log1m_u32(b[2]);                 // drawScatNext material.cpp:215-224
            ok = ok && !(dist < c_k[27]);
            w1 = b[0]; w2 = b[1];
        }
    }
    if (ok) {
        const double u = fma((double)w1, c_k[26], -1.0);              // uniform_real(-1, 1), exact
        double sp, cp; sincospi_unit(fma((double)w2, c_k[26], -1.0), &sp, &cp);
        // drawIso: z = u, sin = sqrt(1 - u^2) ; drawAniso(false): sin^2 = |u|, z = sqrt(1 - |u|)
        const double au = fabs(u);
        const double sth = sqrt(intr ? fma(-u, u, 1.0) : au);
        double z = u;
        if (!intr) z = sqrt(1.0 - au);
        const double lx = sth * cp, ly = sth * sp;
        if (intr) {
            // (sth cp, sth sp, u) is unit to ~2e-16 by construction; Phonon::dir's normalisation (phonon.cpp:88-93) would move it
            // by <= 1 ulp per component, which is the size of the other documented deviations: it is skipped
            ph.set_wp(wp); ph.dx = lx; ph.dy = ly; ph.dz = z; ph.sn = dist;
        } else {
            matvec(cb->m, lx, ly, z, ph.dx, ph.dy, ph.dz);
            renorm_unit(ph.dx, ph.dy, ph.dz);
        }
    } else {                                                           // intrinsic, replayed word by word
        Rng g; g.begin(P.seed, ph.pid(P), ph.step(P));
        ph.set_wp(draw_prop(g, T, T.wprob, T.walias, T.pprob, T.palias));
        draw_iso(g, ph.dx, ph.dy, ph.dz);
        ph.sn = draw_scat_next(g, T.lambda[ph.wp()]);
        renorm_unit(ph.dx, ph.dy, ph.dz);
    }
    ph.nscat++;
}

// The random part of Material::scatter (material.cpp:226-231): its seven words (w int, w real, p int, p real, iso mu,
// iso phi, free path) sit at fixed positions of two Philox blocks, valid when no draw is rejected (probability ~ nw / 2^32);
// a rejection (ok = false) replays the event through the sequential generator in collide(), so the consumed stream is
// always the CPU oracle's.  (Drawing the event ahead of the flight, to interleave its dependency chain with advect / tally,
// was measured: -4 ... -14 %, the extra live registers cost more than the overlap gains.)
struct Draw {
    uint32_t wp; double dx, dy, dz, dist; bool ok;
};
__device__ __forceinline__ void draw_event(const StepParams& P, const Tables& T, uint32_t pid_lo, uint32_t pid_hi, uint32_t event, Draw& o) {
    uint32_t a[4], b[4];
    philox4x32_10_rk(pid_lo, pid_hi, event, 0u, P.rk, a);
    philox4x32_10_rk(pid_lo, pid_hi, event, 1u, P.rk, b);
    uint32_t r = (uint32_t)(((double)a[0] + 0.5) * T.inv_bucket_w);
    uint32_t q = (uint32_t)(((double)a[2] + 0.5) * T.inv_bucket_p);
    bool ok = r < (uint32_t)T.nw && q < (uint32_t)T.np;
    r = min(r, (uint32_t)T.nw - 1u); q = min(q, (uint32_t)T.np - 1u);
    const uint32_t w = (double)a[1] * c_k[25] < T.wprob[r] ? r : (uint32_t)T.walias[r];
    const uint32_t k = w * (uint32_t)T.np + q;
    const uint32_t wp = (double)a[3] * c_k[25] < T.pprob[k] ? k : w * (uint32_t)T.np + (uint32_t)T.palias[k];
    const double c = fma((double)b[0], c_k[26], -1.0);                // drawIso random.cpp:16-27
    const double sth = sqrt(fma(-c, c, 1.0));
    double sp, cp; sincospi_unit(fma((double)b[1], c_k[26], -1.0), &sp, &cp);
    const double dist = T.lambda[wp] * neg_log1m_u32(b[2]);           // drawScatNext material.cpp:215-224
    o.ok = ok && !(dist < c_k[27]);
    // (sth cp, sth sp, c) is unit to ~2e-16 by construction; Phonon::dir's normalisation (phonon.cpp:88-93) would move it by
    // <= 1 ulp per component, which is the size of the other documented deviations: it is skipped
    o.wp = wp; o.dx = sth * cp; o.dy = sth * sp; o.dz = c; o.dist = dist;
}

// Second half of a loop trip (problem.cpp:418-434): Boundary::scatter or Material::scatter, then the stop test.
// UNI: intrinsic and diffuse-wall events share one pass through scatter_draw (kernels of domains with N-D tally grids: +2 ... +4 %;
// the 1-D kernels keep the separate paths: their 80-register budget spills with it, C2 slab -1.7 %).
template <bool BOX, bool UNI>
__device__ __forceinline__ uint32_t collide(const StepParams& P, const Tables& T, Particle& ph, Segment& sg) {
    uint32_t esc = 0;
    const int hit = sg.hit;
    sg.next_plane = hit;
    if (UNI && (hit < 0 || T.cold[hit].kind == MCB_BDRY_DIFF)) {                // material.cpp:226-231 | boundary.cpp:308-312
        scatter_draw<BOX>(P, T, ph, hit < 0, hit >= 0 ? &T.cold[hit] : nullptr);
    } else if (hit >= 0) {                                                     // problem.cpp:418-429
        const DPlaneCold& cb = T.cold[hit];
        const int kind = cb.kind;
        if (kind == MCB_BDRY_SPEC) {                                           // boundary.cpp:283-287
            const DPlaneHot h = T.hot[hit];
            const double c2 = 2.0 * dot3(h.nx, h.ny, h.nz, ph.dx, ph.dy, ph.dz);
            ph.dx -= c2 * h.nx; ph.dy -= c2 * h.ny; ph.dz -= c2 * h.nz;
            renorm_unit(ph.dx, ph.dy, ph.dz);
        } else if (kind == MCB_BDRY_DIFF) {                                    // boundary.cpp:308-312
            Rng g; g.begin(P.seed, ph.pid(P), ph.step(P));
            double ax, ay, az; draw_aniso(g, false, ax, ay, az);
            matvec(cb.m, ax, ay, az, ph.dx, ph.dy, ph.dz);
            renorm_unit(ph.dx, ph.dy, ph.dz);
            ph.nscat++;
        } else if (kind == MCB_BDRY_PERI) {                                    // boundary.cpp:516-522
            double nx, ny, nz; matvec(cb.m, sg.ex, sg.ey, sg.ez, nx, ny, nz);
            ph.px = nx + cb.t[0]; ph.py = ny + cb.t[1]; ph.pz = nz + cb.t[2];
            matvec(cb.m, ph.dx, ph.dy, ph.dz, nx, ny, nz);
            ph.dx = nx; ph.dy = ny; ph.dz = nz;
            renorm_unit(ph.dx, ph.dy, ph.dz);
            sg.next_plane = T.pairs[cb.pair_begin];
            ph.set_sdom((uint32_t)T.cold[sg.next_plane].sdom);
        } else if (kind == MCB_BDRY_INTER) {                                   // boundary.cpp:349-359
            int target = -1;
            if (cb.pair_count == 1) target = T.pairs[cb.pair_begin];
            else for (int q = 0; q < cb.pair_count && target < 0; ++q) {
                const int cand = T.pairs[cb.pair_begin + q];
                if (is_inside<BOX>(T, T.sdom[T.cold[cand].sdom], sg.ex, sg.ey, sg.ez)) target = cand;
            }
            if (target < 0) { ph.kill(); esc = 1; }                            // problem.cpp:422-426
            else ph.set_sdom((uint32_t)T.cold[target].sdom);
            sg.next_plane = target;
        } else {                                                               // Isot boundary.cpp:455-460
            ph.kill();
        }
    } else {                                                                   // Material::scatter material.cpp:226-231
        bool fast = false;
        if (T.nw > 1 && T.np > 1) {
            Draw d2; draw_event(P, T, ph.pid_lo(P), ph.pid_hi(P), ph.step(P), d2);
            fast = d2.ok;
            if (fast) { ph.set_wp(d2.wp); ph.dx = d2.dx; ph.dy = d2.dy; ph.dz = d2.dz; ph.sn = d2.dist; }
        }
        if (!fast) {
            Rng g; g.begin(P.seed, ph.pid(P), ph.step(P));
            ph.set_wp(draw_prop(g, T, T.wprob, T.walias, T.pprob, T.palias));
            draw_iso(g, ph.dx, ph.dy, ph.dz);
            ph.sn = draw_scat_next(g, T.lambda[ph.wp()]);
            renorm_unit(ph.dx, ph.dy, ph.dz);
        }
        ph.nscat++;
    }
    if (ph.nscat >= P.maxscat32 || ph.step(P) >= P.maxloop32) ph.stop();                       // :434, :401
    return esc;
}

// sum(q) of one histogram entry from its three limbs (mcb_device.cuh: deposit / deposit_fx)
__device__ __forceinline__ long long limbs_value(uint32_t p0, uint32_t p1, uint32_t p2) {
    const uint32_t sb = p1 - (p2 << 16), sa = p0 - (sb << 16);
    return (long long)(int32_t)p2 * 4294967296ll + (long long)sb * 65536ll + (long long)sa;
}
// Flush the CTA's `ncopy` three-limb histograms (MCB_TM_BLOCK) into the global fp64 field: per entry the copies are recombined
// and summed as 64-bit integers (exact, order-free), converted once and added with one fp64 RED; the limbs are re-armed
// with zeros.  Entries are taken in the field's own column-major order.
template <int NCOMP>
__device__ __forceinline__ void flush_block_fx(unsigned char* hist, unsigned ncopy, const StepParams& P, unsigned tid, unsigned nthr) {
    for (long long e = tid; e < P.field_len; e += nthr) {
        const long long c = e / P.rows; const int r = (int)(e - c * P.rows);
        long long acc = 0;
        for (unsigned w = 0; w < ncopy; ++w) {
            uint32_t* q = reinterpret_cast<uint32_t*>(hist + (size_t)w * P.hist_bytes + (size_t)c * P.fx_cstride) + 3 * r;
            acc += limbs_value(q[0], q[1], q[2]);
            q[0] = 0u; q[1] = 0u; q[2] = 0u;
        }
        if (acc != 0) atomicAdd(P.field + e, (double)acc * P.fx_inv[r % NCOMP]);
    }
}
// Flush the 1-D difference-array histograms (mcb_device.cuh: deposit_fx) with the whole CTA, in three phases:
//  1. per field entry (r, c): the `ninst` instances' direct and difference entries are recombined from their three limbs,
//     summed as 64-bit integers (exact, order-free) into the scratch arrays and re-armed with zeros;
//  2. per (row, 32-column chunk), one warp: inclusive scan of the difference sums, chunk totals kept;
//  3. per entry: direct + running difference sum (scan + the totals of the chunks before it) is converted once and added
//     to the global field (column-major, fp64 RED).
// Callers place a __syncthreads() before (all deposits done); the next deposits only touch the instances, which are zero
// after phase 1.
template <int NCOMP>
__device__ __forceinline__ void flush_tally1d(unsigned char* hist, unsigned ninst, long long* scratch, const StepParams& P,
                                              unsigned warp, unsigned nwarps, unsigned lane) {
    const int E = P.rows * P.cols, nthr = (int)nwarps * 32, tid = (int)(warp * 32u + lane);
    long long* sdir = scratch; long long* sdiff = scratch + E; long long* ctot = scratch + 2 * E;
    for (int e = tid; e < E; e += nthr) {
        const int r = e / P.cols, c = e - r * P.cols;
        long long direct = 0, diff = 0;
        for (unsigned w = 0; w < ninst; ++w) {
            unsigned char* q0 = hist + (size_t)w * P.hist_bytes + (size_t)(3 * r) * P.fx_ps + 4u * (unsigned)c;
            uint32_t v[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                uint32_t* q = reinterpret_cast<uint32_t*>(q0 + (k >= 3 ? P.fx_diff_off : 0u) + (unsigned)(k % 3) * P.fx_ps);
                v[k] = *q; *q = 0u;
            }
            direct += limbs_value(v[0], v[1], v[2]); diff += limbs_value(v[3], v[4], v[5]);
        }
        sdir[e] = direct; sdiff[e] = diff;
    }
    __syncthreads();
    const int nchunk = (P.cols + 31) / 32, ntask = P.rows * nchunk;
    for (int task = (int)warp; task < ntask; task += (int)nwarps) {
        const int r = task / nchunk, c = (task - r * nchunk) * 32 + (int)lane;
        long long run = c < P.cols ? sdiff[r * P.cols + c] : 0ll;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xFFFFFFFFu, run, o); if ((int)lane >= o) run += v; }
        if (c < P.cols) sdiff[r * P.cols + c] = run;
        if (lane == 31u) ctot[task] = run;
    }
    __syncthreads();
    for (int e = tid; e < E; e += nthr) {
        const int r = e / P.cols, c = e - r * P.cols;
        long long total = sdir[e] + sdiff[e];
        for (int k = 0; k < (c >> 5); ++k) total += ctot[r * nchunk + k];
        if (total != 0) atomicAdd(P.field + (long long)c * P.rows + r, (double)total * P.fx_inv[r % NCOMP]);
    }
}

// ------------------------------------------------------------------------------- k_step
// NCOMP: payload rows per deposit (1: Temp/CumTemp dt ; 3: Flux/CumFlux dpos ; 4: Multi dt,dpos)
// TM   : MCB_TM_WARP / MCB_TM_BLOCK / MCB_TM_GLOBAL
// NDM  : 0 only 1-D / single-cell tally grids; 1 some subdomain has a 2-D/3-D grid; 2 same, and the grid is fine enough
//        (>= 64 cells along an axis) for the warp-cooperative N-D walk to pay for its registers
// BOX  : every subdomain is an axis-aligned box (all shipped box domains): only the single-component plane code is built
// PAD  : NDM == 0 with TM == MCB_TM_WARP uses the 1-D difference-array histograms (mcb_device.cuh: tally_1d); PAD > 0 is
//        the compile-time padded column count of a histogram plane (cols <= PAD), 0 = run-time plane stride
// Dynamic shared memory: [mbarrier 16 B][material blob][geometry blob][histogram(s)]
#ifndef MCB_L2_PREFETCH
#define MCB_L2_PREFETCH 1
#endif
#ifndef MCB_BLOCK_MAX
#define MCB_BLOCK_MAX 768         // 1-D / single-cell tallies: 80 registers x 24 warps (round 2, with the TMA state prefetch: 896 x 72 -11 %, 640 x 96 -1 %)
#endif
#ifndef MCB_BLOCK_MAX_ND1
#define MCB_BLOCK_MAX_ND1 512     // serial N-D walk: 128 registers, no spills (round 2, three-limb CTA histogram: 768 x 80 regs spills 600 B, -27 %; 640 -32 %)
#endif
#ifndef MCB_BLOCK_MAX_ND
#define MCB_BLOCK_MAX_ND 512      // the cooperative N-D walk keeps two crossing iterators live: give it 128 registers
#endif
#ifndef MCB_BLOCK_MAX_ND3
#define MCB_BLOCK_MAX_ND3 640     // warp-balanced N-D tally (mcb_device.cuh: tally_nd_balanced): 96 registers, 4 KB of shared memory per warp
                                  // (512 x 128 registers: C4 -8 %, bulk -2 %; 768 x 80 spills: -20 %)
#endif
template <int NCOMP, int TM, int NDM, bool BOX, int PAD>
__global__ void __launch_bounds__(NDM == 3 ? MCB_BLOCK_MAX_ND3 : (NDM == 2 ? MCB_BLOCK_MAX_ND : (NDM == 1 ? MCB_BLOCK_MAX_ND1 : MCB_BLOCK_MAX)), 1) k_step(const StepParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    unsigned char* s_mat = smem + P.so_mat;
    unsigned char* s_geo = smem + P.so_geo;
    double* s_hist = reinterpret_cast<double*>(smem + P.so_hist);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    constexpr bool T1D = NDM == 0 && TM == MCB_TM_WARP && MCB_TALLY_FX;          // 1-D difference-array histograms
    const unsigned ninst = T1D ? (unsigned)P.hist_copies : 0u;
    const long long hist_words = TM == MCB_TM_GLOBAL ? 0ll : (long long)P.hist_copies * (P.hist_bytes / 4u);

    // --- stage tables: one elected thread arms the mbarrier and issues the TMA bulk copies
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, P.mv.bytes + P.gv.bytes);
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < P.mv.bytes; o += CH) tma_bulk_g2s(s_mat + o, P.mat_blob + o, min(CH, P.mv.bytes - o), bar);
        for (uint32_t o = 0; o < P.gv.bytes; o += CH) tma_bulk_g2s(s_geo + o, P.geo_blob + o, min(CH, P.gv.bytes - o), bar);
    }
    for (long long i = threadIdx.x; i < hist_words; i += blockDim.x) reinterpret_cast<uint32_t*>(s_hist)[i] = 0u;
    while (!mbar_try_wait(bar, 0)) {}
    __syncthreads();

    Tables T;
    T.lambda = reinterpret_cast<const double*>(smem + P.so_lambda);
    T.inv_vel = reinterpret_cast<const double*>(smem + P.so_inv_vel);
    T.wprob = reinterpret_cast<const double*>(smem + P.so_wprob);
    T.pprob = reinterpret_cast<const double*>(smem + P.so_pprob);
    T.walias = reinterpret_cast<const uint16_t*>(smem + P.so_walias);
    T.palias = reinterpret_cast<const uint8_t*>(smem + P.so_palias);
    T.hot = reinterpret_cast<const DPlaneHot*>(smem + P.so_hot);
    T.cold = reinterpret_cast<const DPlaneCold*>(smem + P.so_cold);
    T.sdom = reinterpret_cast<const DSdom*>(smem + P.so_sdom);
    T.pairs = reinterpret_cast<const int32_t*>(smem + P.so_pairs);
    T.nw = P.mv.nw; T.np = P.mv.np; T.inv_bucket_w = P.mv.inv_bucket_w; T.inv_bucket_p = P.mv.inv_bucket_p;
    // CTA histogram: `hist_copies` interleaved copies (a lane picks its copy by lane id) thin out same-word hits
    T.hist = TM == MCB_TM_BLOCK ? reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(s_hist) + (lane & (unsigned)(P.hist_copies - 1)) * P.hist_bytes)
                                : P.field;
    // 1-D difference-array histograms: shared by the CTA, a lane picks its copy by lane id
    const uint32_t my_hist = T1D ? smem_u32(s_hist) + (lane & (unsigned)(P.hist_copies - 1)) * P.hist_bytes : 0u;
    const bool cum = P.kind == MCB_PROB_CUMTEMP || P.kind == MCB_PROB_CUMFLUX;
    constexpr bool FX = TM != MCB_TM_GLOBAL;                      // fixed-point shared-memory histograms (mcb_device.cuh: deposit)

    // counters of this launch: warp-uniform values from ballots, kept in a 16-byte slot per warp in shared memory that lane 0
    // reads, bumps and writes back once per tile -- no per-thread counter registers live across the loop trip, no atomics
    __shared__ uint4 s_wcnt[32];                                // per warp: steps, live, stores, freed slots  (< 2^32 per launch)
    __shared__ unsigned s_esc;
    if (lane == 0) s_wcnt[warp] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) s_esc = 0u;
    // fixed-point histograms are flushed by the CTA between tiles at least every fx_flush_trips loop trips (the host keeps
    // steps_per_launch below that), which bounds the number of deposits an entry can receive (set_fixed_point, mcb_api.cu)
    int since_flush = 0;
    // Free slots are only LISTED during the tile loop (refilling a slot the moment its phonon terminates would run the long
    // emission path for a few dead lanes per warp at ~5 % lane efficiency).  Every WARP appends to its own segment of the list
    // (cursor = the warp's counter slot; a global cursor made every warp wait for a returning L2 atomic once per tile).
    // K1 fused into the launch: a warp lists the slots that end inactive (its own segment of free_list) and, after its last
    // tile, emits the next particles into them with dense lanes (problem.cpp:386-399).  Particle ids come from one atomic
    // cursor (any slot may carry any particle: the Philox stream is keyed by the id), so no emission kernels run between the
    // k_step launches (two launches and 25-45 us of every loop trip in round 1).  The emission call sits behind the tile loop,
    // where nothing of the step is live: a call inside the loop made the 80-register kernels spill loop-carried values.
    uint32_t* const my_free = P.free_list + (size_t)(blockIdx.x * nwarps + warp) * P.free_seg;
    __shared__ unsigned s_emitted;
    if (threadIdx.x == 0) s_emitted = 0u;
    const bool list_free = P.emit_enable && P.ctr->next < P.n_end;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.ctr->live[P.parity ^ 1] = 0ull;

    // TMA state prefetch: a warp's 32 slots are one contiguous 2304-B group (StateView), bulk-copied into the warp's staging
    // buffer while the warp works on the group before it.  The buffer is free again as soon as the lanes have moved their
    // slot into registers, so ONE buffer per warp keeps a full tile in flight.
    const bool staged = P.so_stage != 0u;
    const uint32_t a_buf = smem_u32(smem) + P.so_stage + warp * MCB_GROUP_BYTES;        // shared-memory addresses
    const uint32_t a_bar = smem_u32(smem) + P.so_wbar + warp * 8u;
    // slots this launch visits: the host's count, or -- behind a compacting launch -- the count that launch published
    long long nslots_ll = P.nslots;
    if (P.use_dev_n) nslots_ll = min(nslots_ll, (long long)P.ctr->n_slots);
    const int ngroups = (int)((nslots_ll + 31) >> 5);           // slots < 2^31 (plan_run)
    const int gstride = (int)(gridDim.x * nwarps);
    uint32_t wphase = 0;
    if (staged) {
        const int g0 = (int)(blockIdx.x * nwarps + warp);
        if (lane == 0) { mbar_init(reinterpret_cast<uint64_t*>(smem + P.so_wbar) + warp, 1); fence_mbar_init(); }
        __syncwarp();
        if (lane == 0 && g0 < ngroups) {
            mbar_expect_tx_a(a_bar, MCB_GROUP_BYTES);
            tma_bulk_g2s_a(a_buf, P.st.base + (size_t)g0 * MCB_GROUP_BYTES, MCB_GROUP_BYTES, a_bar);
        }
    }
    // warp-balanced N-D tally: this warp's record area; its mark words start at zero and are re-armed after every trip
    const uint32_t nd_base = NDM == 3 ? smem_u32(smem) + P.so_nd + warp * P.nd_warp_bytes : 0u;
    if (NDM == 3) for (uint32_t i = lane; i < (P.nd_warp_bytes - MCB_NDB_FIXED) / 4u; i += 32u) sts_s32(nd_base + MCB_NDB_FIXED + 4u * i, 0);
    __syncthreads();                                            // the counter slots are armed

    // tile = one group of 32 slots per warp: group g = (blockIdx + k gridDim) nwarps + warp; the trip count is CTA-uniform
    const int nslots = (int)nslots_ll;
    for (int g = (int)(blockIdx.x * nwarps + warp), gt = (int)(blockIdx.x * nwarps); gt < ngroups; g += gstride, gt += gstride) {
        if (FX && P.do_tally) {
            // between tiles (CTA-uniform): flush before an entry could have received more than fx_flush_trips rounds of deposits
            if (since_flush + P.steps_per_launch > P.fx_flush_trips) {
                __syncthreads();
                if (T1D) flush_tally1d<NCOMP>(reinterpret_cast<unsigned char*>(s_hist), ninst, reinterpret_cast<long long*>(smem + P.so_scratch), P, warp, nwarps, lane);
                else flush_block_fx<NCOMP>(reinterpret_cast<unsigned char*>(s_hist), (unsigned)P.hist_copies, P, threadIdx.x, blockDim.x);
                __syncthreads();
                since_flush = 0;
            }
            since_flush += P.steps_per_launch;
        }
        if (g >= ngroups) continue;                             // warp-uniform: this warp has no group in the CTA's last tile
        const int i = g * 32 + (int)lane;
        Particle ph;
        // the whole slot is loaded at once (one memory round trip, not meta first and the rest behind its branch); while
        // the population is full nearly every slot is active, so nothing extra is read.  Slots past nslots in the last group
        // are zero-filled memory (inactive).
        if (staged) {
            while (!mbar_try_wait_a(a_bar, wphase)) {}
            wphase ^= 1u;
            ph.load_shared(a_buf + lane * 16u, a_buf + 2048u + lane * 8u);
        } else ph.load(P.st, i);
        if (i >= nslots || !ph.active()) ph.mlo = 0u; else ph.mlo &= ~(1u << 22);
        const bool was_active = ph.active();
        unsigned tile_steps = 0;
        unsigned run_mask = __ballot_sync(0xFFFFFFFFu, was_active);
#if MCB_L2_PREFETCH
        // no room for the staging buffers (N-D kernels: the record areas take it): at least pull the warp's next group into L2
        if (!staged && lane == 0 && g + gstride < ngroups)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.st.base + (size_t)(g + gstride) * MCB_GROUP_BYTES), "r"(MCB_GROUP_BYTES) : "memory");
#endif
        if (staged && lane == 0 && g + gstride < ngroups) {
            // Refill the staging buffer with the warp's next group.  The ballot above consumed every lane's last-loaded
            // word, so all the lanes' shared-memory reads of the buffer have completed (loads of a warp complete in order);
            // the bulk copy is ordered behind them by that dependency, like a consumer-release / producer-acquire pair --
            // no proxy fence (its MEMBAR also waited for the previous tile's global stores: -5 %).
            mbar_expect_tx_a(a_bar, MCB_GROUP_BYTES);
            tma_bulk_g2s_a(a_buf, P.st.base + (size_t)(g + gstride) * MCB_GROUP_BYTES, MCB_GROUP_BYTES, a_bar);
        }
        for (int s = 0; s < P.steps_per_launch && run_mask != 0u; ++s) {
            // one loop trip (problem.cpp:401-435) in three phases; the tally phase is warp-synchronous
            const bool run = ph.active();
            tile_steps += __popc(run_mask);
            Segment sg; sg.ok = false; sg.hit = -1; sg.d = 0.0; sg.nscat_before = 0;
            sg.bx = sg.by = sg.bz = sg.ex = sg.ey = sg.ez = 0.0;
            bool escaped = false;
            if (run) escaped = advect_move<BOX>(T, ph, sg) != 0u;
            if (P.do_tally) {
                // accumAmt (problem.cpp:473-476,506-509,539-544,581-589,629-637) times sign (:414)
                const DSdom& sd = T.sdom[ph.sdom()];
                const int rbase = cum ? NCOMP * (int)(((long long)sg.nscat_before + P.cum_step - 1) / P.cum_step) : 0;
                if (T1D) {
                    // fixed point: sign and the launch's power-of-two scale in one factor per component (exact); only the
                    // dt row can exceed the fixed-point range (a flight of one of the few very slow modes): those
                    // flights are deposited exactly, fp64 RED straight to the global field, by the generic walk
                    const bool neg = !ph.sign();
                    double amt[NCOMP];
                    bool slow = false;
                    if (NCOMP != 3) {
                        const double dt = sg.d * T.inv_vel[ph.wp()];
                        slow = sg.ok && !(dt <= P.fx_max[0]);
                        amt[0] = dt * (neg ? -P.fx_scale[0] : P.fx_scale[0]);
                    }
                    if (NCOMP != 1) {
                        const double f = neg ? -P.fx_scale[NCOMP - 1] : P.fx_scale[NCOMP - 1];
                        amt[NCOMP - 3] = (sg.ex - sg.bx) * f; amt[NCOMP - 2] = (sg.ey - sg.by) * f; amt[NCOMP - 1] = (sg.ez - sg.bz) * f;
                    }
                    if (__any_sync(0xFFFFFFFFu, slow)) {
                        if (slow) {
                            double raw[NCOMP];
#pragma unroll
                            for (int k = 0; k < NCOMP; ++k) raw[k] = amt[k] * P.fx_inv[k];
                            tally_segments<NCOMP, MCB_TM_GLOBAL, false, false>(sd, P.field, P.rows, P.cols, rbase, true, sg.bx, sg.by, sg.bz, sg.ex, sg.ey, sg.ez, raw, lane);
                        }
                        __syncwarp();
                    }
                    Tally1D t1;
                    tally1d_setup<BOX>(sd, sg.ok && !slow, sg.bx, sg.by, sg.bz, sg.ex, sg.ey, sg.ez, t1);
                    const uint32_t ps3 = 3u * (PAD > 0 ? (uint32_t)PAD * 4u : P.fx_ps);
                    tally1d_deposit<NCOMP, PAD>(t1, my_hist, (uint32_t)rbase * ps3, P.fx_ps, P.fx_diff_off, amt);
                } else {
                    __syncwarp();
                    const double sg_ = ph.sign() ? 1.0 : -1.0;
                    double amt[NCOMP];
                    if (NCOMP == 1) amt[0] = sg_ * (sg.d * T.inv_vel[ph.wp()]);
                    else if (NCOMP == 4) { amt[0] = sg_ * (sg.d * T.inv_vel[ph.wp()]); amt[1 % NCOMP] = sg_ * (sg.ex - sg.bx); amt[2 % NCOMP] = sg_ * (sg.ey - sg.by); amt[3 % NCOMP] = sg_ * (sg.ez - sg.bz); }
                    else { amt[0] = sg_ * (sg.ex - sg.bx); amt[1 % NCOMP] = sg_ * (sg.ey - sg.by); amt[2 % NCOMP] = sg_ * (sg.ez - sg.bz); }
                    FxArgs fx{P.fx_cstride, false, &P};
                    double raw[NDM == 3 ? NCOMP : 1];
                    if (FX) {
                        // fixed-point histograms: scale by the launch's power of two (exact); a payload beyond the chosen range
                        // (rare: a long flight of a very slow mode) is deposited exactly through the global fp64 path instead
                        bool fits = true;
#pragma unroll
                        for (int k = 0; k < NCOMP; ++k) { fits = fits && fabs(amt[k]) <= P.fx_max[k]; if (NDM == 3) raw[k] = amt[k]; amt[k] *= P.fx_scale[k]; }
                        fx.slow = !fits;
                    }
                    if (NDM == 3) {
                        // N-D grids: one work item per crossed cell, dealt out over the warp; the 1-D and single-cell grids of
                        // a mixed domain keep the serial iterator
                        const bool nd_seg = sg.ok && sd.accum >= 3, flat_seg = sg.ok && sd.accum < 3;
                        if (__any_sync(0xFFFFFFFFu, flat_seg))
                            tally_segments<NCOMP, TM, false, true, false>(sd, T.hist, P.rows, P.cols, rbase, flat_seg, sg.bx, sg.by, sg.bz, sg.ex, sg.ey, sg.ez, amt, lane, fx);
                        if (FX && fx.slow) {
#pragma unroll
                            for (int k = 0; k < NCOMP; ++k) amt[k] = raw[k];
                        }
                        tally_nd_balanced<NCOMP, TM, BOX>(nd_base, sd, T.hist, P.rows, P.cols, rbase, nd_seg, FX && fx.slow, sg.bx, sg.by, sg.bz, sg.ex, sg.ey, sg.ez, amt, lane, fx);
                    } else
                    tally_segments<NCOMP, TM, (NDM > 0), true, (NDM == 2)>(sd, T.hist, P.rows, P.cols, rbase, sg.ok, sg.bx, sg.by, sg.bz, sg.ex, sg.ey, sg.ez, amt, lane, fx);
                }
            }
            if (sg.ok) escaped = collide<BOX, (NDM > 0)>(P, T, ph, sg) != 0u;
            if (escaped) red_shared_u32(&s_esc, 1u);                           // rare: Progress::incrEsc problem.cpp:111-118
            run_mask = __ballot_sync(0xFFFFFFFFu, ph.active());
        }
        unsigned st_mask;
        if (P.compact) {
            // K3 fused into the launch: the tile's survivors take the next slots of the output buffer (one atomic per tile; with
            // several loop trips per tile its latency hides behind the other warps), terminated phonons are dropped
            st_mask = run_mask;
            if (run_mask) {
                unsigned long long ob = 0ull;
                if (lane == 0) ob = atomicAdd(&P.ctr->out_cursor, (unsigned long long)__popc(run_mask));
                ob = __shfl_sync(0xFFFFFFFFu, ob, 0);
                if ((run_mask >> lane) & 1u) ph.store(P.st_out, (long long)(ob + __popc(run_mask & ((1u << lane) - 1u))));
            }
        } else {
            if (was_active) ph.store_group(P.st, g, lane);
            st_mask = __ballot_sync(0xFFFFFFFFu, was_active);
        }
        // end of the tile: list the slots that ended inactive (only ids below nslots), bump the warp's counters
        const unsigned fm = list_free ? __ballot_sync(0xFFFFFFFFu, i < nslots) & ~run_mask : 0u;
        uint4 wc = make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) wc = s_wcnt[warp];                       // the warp's slot is read and written by lane 0 only
        const unsigned nlisted = __shfl_sync(0xFFFFFFFFu, wc.w, 0);
        if ((fm >> lane) & 1u) my_free[nlisted + __popc(fm & ((1u << lane) - 1u))] = (uint32_t)i;
        if (lane == 0) {
            wc.x += tile_steps; wc.y += __popc(run_mask); wc.z += __popc(st_mask); wc.w += __popc(fm);
            s_wcnt[warp] = wc;
        }
        __syncwarp();
    }
    // K1: refill the slots this warp listed (see above)
    if (list_free) {
        const unsigned n_w = s_wcnt[warp].w;
        if (n_w) {
            unsigned long long base = 0ull;
            if (lane == 0) base = atomicAdd(&P.ctr->next, (unsigned long long)n_w);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            const unsigned long long room = base < P.n_end ? P.n_end - base : 0ull;
            const unsigned ne = room < (unsigned long long)n_w ? (unsigned)room : n_w;
#pragma unroll 1
            for (unsigned j = lane; j < ne; j += 32u) {
                Particle np_;
                emit_particle(P, T, base + j, np_);
                np_.store(P.st, (long long)my_free[j]);
            }
            if (lane == 0 && ne) { red_shared_u32(&s_emitted, ne); if (P.maxloop > 0) s_wcnt[warp].y += ne; }     // refilled slots are live
        }
    }

    // --- the CTA's counters: one global atomic each; the warps' free-list counts
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long st = 0, lv = 0, so = 0;
        for (unsigned w = 0; w < nwarps; ++w) { st += s_wcnt[w].x; lv += s_wcnt[w].y; so += s_wcnt[w].z; }
        if (st) atomicAdd(&P.ctr->steps, st);
        if (s_esc) atomicAdd(&P.ctr->esc, (unsigned long long)s_esc);
        if (lv) atomicAdd(&P.ctr->live[P.parity], lv);
        if (so) atomicAdd(&P.ctr->stores, so);
        if (s_emitted) atomicAdd(&P.ctr->emitted, (unsigned long long)s_emitted);
        if (P.host_ctr || P.compact) {
            // the last CTA to get here mirrors the counters into pinned host memory: the host only waits for the event behind
            // the launch -- no device-to-host copy sits between two k_step launches
            __threadfence();
            if (atomicAdd(&P.ctr->done, 1ull) == (unsigned long long)gridDim.x - 1ull) {
                __threadfence();
                if (P.compact) {
                    // publish the number of slots written for the launch behind this one, re-arm the cursor and make the
                    // rest of the last 32-slot group inactive (the buffer holds stale state from two launches ago)
                    const unsigned long long n_out = *reinterpret_cast<volatile unsigned long long*>(&P.ctr->out_cursor);
                    P.ctr->n_slots = n_out; P.ctr->out_cursor = 0ull;
                    for (unsigned long long k = n_out; k < ((n_out + 31ull) & ~31ull); ++k) Particle::clear_meta(P.st_out, (long long)k);
                    __threadfence();
                }
                if (P.host_ctr) {
                    const volatile unsigned long long* src = reinterpret_cast<const volatile unsigned long long*>(P.ctr);
                    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(P.host_ctr);
                    for (int k = 0; k < (int)(sizeof(Counters) / 8); ++k) dst[k] = src[k];
                }
                P.ctr->done = 0ull;
                __threadfence_system();
            }
        }
    }
    // --- flush the shared-memory histogram(s): sum the copies, transpose row-major -> the field's column-major layout,
    //     one fp64 RED per non-zero entry
    if (TM != MCB_TM_GLOBAL && P.do_tally) {
        if (T1D) flush_tally1d<NCOMP>(reinterpret_cast<unsigned char*>(s_hist), ninst, reinterpret_cast<long long*>(smem + P.so_scratch), P, warp, nwarps, lane);
        else flush_block_fx<NCOMP>(reinterpret_cast<unsigned char*>(s_hist), (unsigned)P.hist_copies, P, threadIdx.x, blockDim.x);
    }
}

#ifdef MCB_AUX_KERNELS      // the non-template kernels live in ONE translation unit (mcb_api.cu)
// ------------------------------------------------------------------------------- k_emit
// K1 for the FIRST fill (every slot is free; afterwards k_step refills its own slots): particle next + j into slot j
// (problem.cpp:386-399), one thread per particle; k_emit_commit then advances the particle counter.
__device__ __forceinline__ unsigned long long emit_quota(const StepParams& P) {
    const unsigned long long next = P.ctr->next, room = P.n_end > next ? P.n_end - next : 0ull, nfree = (unsigned long long)P.nslots;
    return nfree < room ? nfree : room;
}
__global__ void __launch_bounds__(256) k_emit(const StepParams P) {
    const unsigned long long next = P.ctr->next, n = emit_quota(P);
    Tables T;
    T.lambda = reinterpret_cast<const double*>(P.mat_blob + P.mv.off_lambda);
    T.nw = P.mv.nw; T.np = P.mv.np; T.inv_bucket_w = P.mv.inv_bucket_w; T.inv_bucket_p = P.mv.inv_bucket_p;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (unsigned long long)gridDim.x * blockDim.x) {
        Particle ph;
        emit_particle(P, T, next + j, ph);
        ph.store(P.st, (long long)j);
    }
}
__global__ void k_emit_commit(const StepParams P) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long n = emit_quota(P);
        P.ctr->next += n; P.ctr->emitted += n; P.ctr->live[0] = 0ull; P.ctr->live[1] = 0ull;
    }
}

// ------------------------------------------------------------------------------- k_traj
// TrajProblem::solve (problem.cpp:226-299) for ONE particle (id 0), recording TrkPhonon's polyline
// (phonon.cpp:129-170) and the per-trip boundary trace.  Diagnostic: a single thread, tables read from global memory.
struct TrajDev {
    double* points; long long max_points; long long max_steps;
    int32_t *step_sdom, *step_in, *step_in_kind, *step_out, *step_out_kind;
    long long* counts;            // [0] npoints, [1] nsteps, [2] escaped
};
__global__ void k_traj(const StepParams P, const mcb_traj_desc t, const TrajDev o) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Tables T;
    T.lambda = reinterpret_cast<const double*>(P.mat_blob + P.mv.off_lambda);
    T.inv_vel = reinterpret_cast<const double*>(P.mat_blob + P.mv.off_inv_vel);
    T.wprob = reinterpret_cast<const double*>(P.mat_blob + P.mv.off_wprob);
    T.pprob = reinterpret_cast<const double*>(P.mat_blob + P.mv.off_pprob);
    T.walias = reinterpret_cast<const uint16_t*>(P.mat_blob + P.mv.off_walias);
    T.palias = reinterpret_cast<const uint8_t*>(P.mat_blob + P.mv.off_palias);
    T.hot = reinterpret_cast<const DPlaneHot*>(P.geo_blob + P.gv.off_hot);
    T.cold = reinterpret_cast<const DPlaneCold*>(P.geo_blob + P.gv.off_cold);
    T.sdom = reinterpret_cast<const DSdom*>(P.geo_blob + P.gv.off_sdom);
    T.pairs = reinterpret_cast<const int32_t*>(P.geo_blob + P.gv.off_pairs);
    T.nw = P.mv.nw; T.np = P.mv.np; T.inv_bucket_w = P.mv.inv_bucket_w; T.inv_bucket_p = P.mv.inv_bucket_p; T.hist = nullptr;

    long long npts = 0, nsteps = 0, escaped = 0;
    auto push = [&](double x, double y, double z) {
        if (npts < o.max_points) { o.points[3 * npts] = x; o.points[3 * npts + 1] = y; o.points[3 * npts + 2] = z; }
        npts++;
    };
    Particle ph; ph.init(0u, true, true, 0u, 0ull, P.step_bits);
    Rng g; g.begin(P.seed, 0ull, 0u);
    ph.set_wp(t.has_prop ? (uint32_t)(t.w * T.np + t.p) : draw_prop(g, T, T.wprob, T.walias, T.pprob, T.palias));   // :232
    int cur = -1;
    if (t.has_pos) {                                                                                          // :233-243
        ph.px = t.pos[0]; ph.py = t.pos[1]; ph.pz = t.pos[2];
        if (t.has_dir) { ph.dx = t.dir[0]; ph.dy = t.dir[1]; ph.dz = t.dir[2]; }
        else draw_iso(g, ph.dx, ph.dy, ph.dz);
        normalize3(ph.dx, ph.dy, ph.dz);
        ph.set_sdom((uint32_t)t.sdom);
    } else {                                                                                                  // :244-253
        const DEmitter& E = P.emitters[g.uint_below((uint32_t)P.nemitter)];
        const uint32_t wp = ph.wp();
        emit_from(E, g, ph);
        ph.init(wp, true, true, ph.sdom(), 0ull, P.step_bits);         // TrajProblem traces with sign +1 (problem.cpp:244-253)
        cur = E.kind == MCB_EMIT_BDRY ? E.index : -1;
    }
    push(ph.px, ph.py, ph.pz);
    ph.sn = draw_scat_next(g, T.lambda[ph.wp()]);                                                             // :254
    for (long long i = 0; i < P.maxloop; ++i) {                                                               // :258
        const DSdom& sd = T.sdom[ph.sdom()];
        const long long k = nsteps++;
        const int in_local = (cur >= sd.plane_begin && cur < sd.plane_begin + sd.plane_count) ? cur - sd.plane_begin : -1;
        if (k < o.max_steps) {
            o.step_sdom[k] = (int32_t)ph.sdom(); o.step_in[k] = in_local; o.step_in_kind[k] = cur >= 0 ? T.cold[cur].kind : -1;
            o.step_out[k] = -1; o.step_out_kind[k] = -1;
        }
        Segment sg; sg.ok = false; sg.hit = -1; sg.next_plane = -1;
        const uint32_t esc = advect_move<false>(T, ph, sg);                                                          // :265
        push(ph.px, ph.py, ph.pz);
        if (esc) { escaped = 1; break; }                                                                      // :267-271
        if (k < o.max_steps) { o.step_out[k] = sg.hit >= 0 ? sg.hit - sd.plane_begin : -1; o.step_out_kind[k] = sg.hit >= 0 ? T.cold[sg.hit].kind : -1; }
        const bool peri = sg.hit >= 0 && T.cold[sg.hit].kind == MCB_BDRY_PERI;
        if (collide<false, false>(P, T, ph, sg)) { escaped = 2; break; }                                                    // :277-294
        if (peri) push(ph.px, ph.py, ph.pz);
        cur = sg.next_plane;
        if (!ph.active()) break;                                                                              // :295
    }
    o.counts[0] = npts; o.counts[1] = nsteps; o.counts[2] = escaped;
}

// ---------------------------------------------------------------------------- k_compact
// Unordered stream compaction of active slots from `src` into `dst` (tail of the solve).
__global__ void k_compact(StateView src, StateView dst, long long n, Counters* ctr) {
    const unsigned lane = threadIdx.x & 31u;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        Particle ph; ph.mlo = 0;
        if (i < n) ph.load(src, i);
        const bool act = (i < n) && ph.active();
        const unsigned m = __ballot_sync(0xFFFFFFFFu, act);
        if (!m) continue;
        unsigned long long pos = 0;
        const int leader = __ffs(m) - 1;
        if ((int)lane == leader) pos = atomicAdd(&ctr->compact_cursor, (unsigned long long)__popc(m));
        pos = __shfl_sync(0xFFFFFFFFu, pos, leader);
        if (act) ph.store(dst, (long long)(pos + __popc(m & ((1u << lane) - 1u))));
    }
}

// ------------------------------------------------------------------------------ k_sort_*
// K3 with a key (mcb_options::sort_mode): the survivors are compacted AND ordered by (subdomain, tally cell), i.e. by the
// field column their position lies in -- Field::init gives every subdomain a contiguous column range (field.cpp:25-45) and
// a cell the column  offset + sum_d coord2index(coord(pos))_d * stride_d  (subdomain.cpp:148-159, field.cpp:38).  A counting
// sort in three launches, every one of them over the slots the step kernel would visit:
//   k_sort_count    key (bin = column / cols_per_bin, at most MCB_SORT_MAX_BINS bins) of every active slot; per-CTA histogram
//                   in shared memory, then one global atomic per non-empty bin and CTA; the key is kept per slot (4 B)
//   k_sort_scan     exclusive prefix sum of the bin counts by one CTA -> the first destination slot of every bin
//   k_sort_scatter  every active slot takes the next destination slot of its bin (atomic cursor per bin) and is moved there
// Slots of one bin end up contiguous; their order inside the bin is the order of the atomics (like k_compact's cursor, it
// only permutes the summation order of the tallies: the RNG stream of a phonon is keyed by its particle id).
#define MCB_SORT_MAX_BINS 16384
#define MCB_SORT_NOKEY 0xFFFFFFFFu
__device__ __forceinline__ long long sort_column(const DSdom& sd, double px, double py, double pz) {
    if (sd.col_offset < 0) return 0;                                  // subdomain without cells: any bin
    double c[3]; sdom_coord(sd, px, py, pz, c);
    return (long long)sd.col_offset + coord2index1(c[0], sd.max[0]) + (long long)coord2index1(c[1], sd.max[1]) * sd.stride1 +
           (long long)coord2index1(c[2], sd.max[2]) * sd.stride2;
}
__global__ void __launch_bounds__(256) k_sort_count(StateView src, long long n, const unsigned char* geo_blob, GeometryView gv,
                                                    uint32_t cols_per_bin, uint32_t nbins, uint32_t* keys, uint32_t* bins) {
    extern __shared__ uint32_t s_bins[];
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) s_bins[b] = 0u;
    __syncthreads();
    const DSdom* sds = reinterpret_cast<const DSdom*>(geo_blob + gv.off_sdom);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // position and meta only: three of the slot's five pieces
        const unsigned char* v = Particle::vec_ptr(src, i);
        double px, py, pz, dx, sn; unsigned long long m;
        ld_stream2(v + 1536, sn, m);
        uint32_t key = MCB_SORT_NOKEY;
        if (MCB_META_ACTIVE(m)) {
            ld_stream2(v, px, py); ld_stream2(v + 512, pz, dx);
            key = (uint32_t)(sort_column(sds[MCB_META_SDOM(m)], px, py, pz) / cols_per_bin);
            key = min(key, nbins - 1u);
            atomicAdd(&s_bins[key], 1u);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) { const uint32_t v = s_bins[b]; if (v) atomicAdd(&bins[b], v); }
}
// one CTA of 1024 threads: bins[b] <- number of active slots in bins < b; total[0] <- number of active slots
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* bins, uint32_t nbins, unsigned long long* total) {
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (nbins + 1023u) / 1024u, b0 = threadIdx.x * per, b1 = min(nbins, b0 + per);
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; ++b) sum += bins[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) {                       // inclusive scan of the per-thread sums
        const uint32_t v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;
    for (uint32_t b = b0; b < b1; ++b) { const uint32_t v = bins[b]; bins[b] = run; run += v; }
    if (threadIdx.x == 1023u) *total = (unsigned long long)s_part[1023];
}
__global__ void __launch_bounds__(256) k_sort_scatter(StateView src, StateView dst, long long n, const uint32_t* keys, uint32_t* bins) {
    const unsigned lane = threadIdx.x & 31u;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;                      // the trip count is uniform across the CTA
        const uint32_t key = i < n ? keys[i] : MCB_SORT_NOKEY;
        // lanes of a warp that share the bin take consecutive destination slots with ONE atomic (neighbouring source slots of
        // an already sorted population mostly do)
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
        if (key == MCB_SORT_NOKEY) continue;
        Particle ph; ph.load(src, i);
        const int leader = __ffs(peers) - 1;
        uint32_t pos = 0;
        if ((int)lane == leader) pos = atomicAdd(&bins[key], (uint32_t)__popc(peers));
        pos = __shfl_sync(peers, pos, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        ph.store(dst, (long long)pos);
    }
}
// diagnostic behind mcb_sort_probe: bin key, field column and particle id of every slot (slot order); -1 for inactive slots
__global__ void k_sort_probe(StateView st, long long n, const unsigned char* geo_blob, GeometryView gv, uint32_t cols_per_bin,
                             uint32_t nbins, uint32_t step_bits, long long* bin, long long* col, long long* pid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Particle ph; const unsigned long long m = ph.load(st, i);
    bin[i] = col[i] = pid[i] = -1;
    if (!MCB_META_ACTIVE(m)) return;
    const DSdom* sds = reinterpret_cast<const DSdom*>(geo_blob + gv.off_sdom);
    const long long cc = sort_column(sds[MCB_META_SDOM(m)], ph.px, ph.py, ph.pz);
    col[i] = cc; bin[i] = (long long)min((uint32_t)(cc / cols_per_bin), nbins - 1u); pid[i] = (long long)(ph.ps >> step_bits);
}

// --------------------------------------------------------------------------- k_finalize
// problem.cpp:439-444: postProc (problem.cpp:478-481,511-514,546-551,591-599,639-648), / cellVol, * power_
__global__ void k_finalize(double* field, int rows, long long cols, int kind, long long size,
                           double energy_sum, double power, const double* cell_vol) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double* f = field + c * rows;
    const double vol = cell_vol[c];
    if (kind == MCB_PROB_CUMTEMP) {
        for (long long i = 0; i < size; ++i) f[i + 1] += f[i];
    } else if (kind == MCB_PROB_CUMFLUX) {
        for (long long i = 0; i < size; ++i) for (int q = 0; q < 3; ++q) f[3 * (i + 1) + q] += f[3 * i + q];
    }
    for (int r = 0; r < rows; ++r) {
        double v = f[r];
        if (kind == MCB_PROB_TEMP || kind == MCB_PROB_CUMTEMP || (kind == MCB_PROB_MULTI && r == 0)) v = v / energy_sum;
        f[r] = power * (v / vol);
    }
}

// ------------------------------------------------------------------------- diagnostics
__global__ void k_cell_index(const unsigned char* geo_blob, GeometryView gv, long long n, const double* pos,
                             const int32_t* sdom, long long* index) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DSdom* sds = reinterpret_cast<const DSdom*>(geo_blob + gv.off_sdom);
    const DSdom& sd = sds[sdom[i]];
    double c[3]; sdom_coord(sd, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], c);
    for (int d = 0; d < 3; ++d) index[3 * i + d] = coord2index1(c[d], sd.max[d]);
}

__global__ void k_accumulate(const unsigned char* geo_blob, GeometryView gv, int rows, long long n, const int32_t* sdom,
                             const double* bpos, const double* epos, const double* amount, double* field) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DSdom* sds = reinterpret_cast<const DSdom*>(geo_blob + gv.off_sdom);
    const DSdom& sd = sds[sdom[i]];
    const double* a = amount + (long long)rows * i;
    // generic row count: deposit one component at a time (same weights, same order per component)
    for (int r = 0; r < rows; ++r)
        tally_segments<1, MCB_TM_GLOBAL, true, false>(sd, field, rows, 0, r, true, bpos[3 * i], bpos[3 * i + 1], bpos[3 * i + 2],
                                               epos[3 * i], epos[3 * i + 1], epos[3 * i + 2], a + r, threadIdx.x & 31u);
}

// the device generator's words for one (seed, particle, event, block) — behind mcb_philox_words
__global__ void k_philox(unsigned long long seed, unsigned long long pid, uint32_t event, uint32_t block, uint32_t* out) {
    uint32_t w[4];
    philox4x32_10((uint32_t)pid, (uint32_t)(pid >> 32), event, block, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    for (int i = 0; i < 4; ++i) out[i] = w[i];
}

// scatter the slot state to per-particle trace arrays (slot order is not particle order)
__global__ void k_gather_trace(StateView st, long long nslots, unsigned long long n_begin, long long n, int np, uint32_t step_bits,
                               const unsigned char* geo_blob, GeometryView gv,
                               double* pos, double* dir, double* sn, long long* w, long long* p, int32_t* sign,
                               int32_t* alive, int32_t* sdom, long long* nscat, long long* steps, int32_t* cell) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    Particle ph;
    const unsigned long long meta = ph.load(st, i), ps = ph.ps;
    const long long o = (long long)((ps >> step_bits) - n_begin);
    if (o < 0 || o >= n) return;
    pos[3 * o] = ph.px; pos[3 * o + 1] = ph.py; pos[3 * o + 2] = ph.pz;
    dir[3 * o] = ph.dx; dir[3 * o + 1] = ph.dy; dir[3 * o + 2] = ph.dz;
    sn[o] = ph.sn;
    const uint32_t wp = MCB_META_WP(meta);
    w[o] = wp / np; p[o] = wp % np;
    sign[o] = MCB_META_SIGN(meta) ? 1 : -1;
    alive[o] = MCB_META_KILLED(meta) ? 0 : 1;
    sdom[o] = (int32_t)MCB_META_SDOM(meta);
    nscat[o] = MCB_META_NSCAT(meta);
    steps[o] = (long long)(ps & ((1ull << step_bits) - 1ull));
    const DSdom* sds = reinterpret_cast<const DSdom*>(geo_blob + gv.off_sdom);
    const DSdom& sd = sds[MCB_META_SDOM(meta)];
    double c[3]; sdom_coord(sd, ph.px, ph.py, ph.pz, c);
    for (int d = 0; d < 3; ++d) cell[3 * o + d] = (int32_t)coord2index1(c[d], sd.max[d]);
}

#endif // MCB_AUX_KERNELS

} // namespace mcb
