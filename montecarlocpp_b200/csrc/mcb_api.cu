// mcb_api.cu — host side of the C ABI declared in include/mcb.h.
//
// Flattens the POD descriptors into device tables, drives the slot schedule (first fill, step launches that
// refill their own slots, compacting launches of the decay phase, optional sort), finalises the field.  No CPU fallback: every entry point needs a
// working sm_100 device.  Reference citations: file:line relative to /root/reference/montecarlo/.
#define MCB_AUX_KERNELS
#include "mcb_kernels.cuh"

#include <algorithm>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

using namespace mcb;

namespace mcb {
cudaError_t launch_step_n1(const StepParams&, int, int, int, int, int, int, size_t, cudaStream_t);
cudaError_t launch_step_n3(const StepParams&, int, int, int, int, int, int, size_t, cudaStream_t);
cudaError_t launch_step_n4(const StepParams&, int, int, int, int, int, int, size_t, cudaStream_t);
}

namespace {

thread_local std::string g_create_err;

#define CUDA_TRY(ctx, expr)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                  \
            return MCB_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }                       // scratch buffers of the diagnostic entry points are freed on every return path
};

// boost::random::discrete_distribution<long,double>: Walker alias table (random.h:26), as used by
// Material::Dist (material.cpp:51-75).  Published algorithm (Boost.Random 1.5x, discrete_distribution.hpp).
void build_alias(const double* w, size_t n, double* prob, int32_t* alias) {
    std::vector<std::pair<double, int32_t>> below, above;
    double sum = 0.0;
    for (size_t i = 0; i < n; ++i) sum += w[i];
    const double avg = sum / (double)n;
    for (size_t i = 0; i < n; ++i) {
        const double v = w[i] / avg;
        if (v < 1.0) below.emplace_back(v, (int32_t)i); else above.emplace_back(v, (int32_t)i);
    }
    for (size_t i = 0; i < n; ++i) { prob[i] = 0.0; alias[i] = 0; }
    size_t b = 0, a = 0;
    while (b < below.size() && a < above.size()) {
        prob[below[b].second] = below[b].first; alias[below[b].second] = above[a].second;
        above[a].first -= (1.0 - below[b].first);
        if (above[a].first < 1.0) { below[b] = above[a]; ++a; } else { ++b; }
    }
    for (; b < below.size(); ++b) prob[below[b].second] = 1.0;
    for (; a < above.size(); ++a) prob[above[a].second] = 1.0;
}

struct AliasTables {            // entry (w,p) at [w*np + p]
    std::vector<double> wprob, pprob; std::vector<int32_t> walias, palias;
    void build(const double* pdf, long nw, long np) {     // pdf(w,p) at [w + nw*p] (column-major)
        wprob.resize(nw); walias.resize(nw); pprob.resize(nw * np); palias.resize(nw * np);
        std::vector<double> rowsum(nw), row(np);
        for (long w = 0; w < nw; ++w) { double s = 0.0; for (long p = 0; p < np; ++p) s += pdf[w + nw * p]; rowsum[w] = s; }
        build_alias(rowsum.data(), nw, wprob.data(), walias.data());
        for (long w = 0; w < nw; ++w) {
            for (long p = 0; p < np; ++p) row[p] = pdf[w + nw * p];
            build_alias(row.data(), np, &pprob[w * np], &palias[w * np]);
        }
    }
};

inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

} // namespace

struct mcb_ctx {
    int device = 0; int sm_count = 148; size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evA[2] = {nullptr, nullptr}, evB[2] = {nullptr, nullptr}, evC[2] = {nullptr, nullptr};
    std::string err;
    mcb_options opt{};
    // material
    bool has_mat = false; long nw = 0, np = 0; double energy_sum = 0, flux_sum = 0;
    double inv_vel_fx = 0;       // slowness 1/vel that all but 0.1 % of the drawn modes stay below: sets the fixed-point
                                 // range of the dt payload (slower flights take the exact fp64 path)
    double flight_max = 0;       // largest possible flight inside one subdomain (sum of its edge-vector lengths)
    MaterialView mv{}; DevBuf<unsigned char> mat_blob;
    AliasTables flux_alias, scat_alias;
    DevBuf<double> f_wprob, f_pprob; DevBuf<int32_t> f_walias, f_palias;
    // geometry
    bool has_dom = false; GeometryView gv{}; DevBuf<unsigned char> geo_blob;
    std::vector<DSdom> h_sdom; int nemitter = 0; int any_nd = 0;      // 0 / 1 / 2: see k_step's NDM
    int nd_mark_words = 0;       // warp-balanced N-D tally: 32-bit mark words a warp needs = max items of one flight (1 + sum of max_)
    long long nd_max_cells = 0;  // ... and the largest max_ of an N-D grid axis (16-bit record fields)
    int all_box = 0;             // every subdomain is an axis-aligned box (k_step's BOX)
    int t1_ok = 0;               // every 1-D tally grid has unit column stride (the difference-array histograms apply)
    DevBuf<DEmitter> emitters; DevBuf<double> cell_vol; long long cols = 0;
    // problem / run state
    DevBuf<long long> emit_cdf;
    DevBuf<unsigned char> state[2]; long long slots_alloc = 0;          // warp-tiled slot state (mcb_device.cuh: StateView)
    DevBuf<Counters> ctr; Counters* h_ctr = nullptr;     // pinned mirror (two slots), written by k_step itself through ...
    Counters* d_hctr = nullptr;                          // ... its device address
    DevBuf<double> field;
    DevBuf<uint32_t> free_list;  // fused emission: per-warp segments of free slot ids
    DevBuf<uint32_t> sort_keys, sort_bins;   // K3 with a key (mcb_options::sort_mode): bin of every slot, bin counts / cursors
};

namespace {

StateView view_of(mcb_ctx* c, int which) { return StateView{c->state[which].p}; }

// bits of the pid|step word given to the loop trip count: 28 by default (nemit < 2^36), up to 32 for very long histories
uint32_t step_bits_for(long long maxloop) {
    uint32_t b = MCB_STEP_BITS_MIN;
    while (b < MCB_STEP_BITS_MAX && (maxloop >> b) != 0) ++b;
    return b;
}

int check_problem(mcb_ctx* c, const mcb_problem_desc* p) {
    if (!p) { c->err = "null problem"; return MCB_EINVAL; }
    if (!c->has_mat || !c->has_dom) { c->err = "upload material and domain before solving"; return MCB_ESTATE; }
    int rows;
    switch (p->kind) {
    case MCB_PROB_TEMP: rows = 1; break;
    case MCB_PROB_FLUX: rows = 3; break;
    case MCB_PROB_MULTI: rows = 4; break;
    case MCB_PROB_CUMTEMP: rows = (int)p->size + 1; break;
    case MCB_PROB_CUMFLUX: rows = 3 * ((int)p->size + 1); break;
    default: c->err = "Invalid problem"; return MCB_EINVAL;
    }
    if (p->rows != rows) { c->err = "problem rows inconsistent with kind/size"; return MCB_EINVAL; }
    if ((p->kind == MCB_PROB_CUMTEMP || p->kind == MCB_PROB_CUMFLUX) && (p->size <= 0 || p->step <= 0)) {
        c->err = "Cum* problems need size > 0 and step > 0"; return MCB_EINVAL;
    }
    if (!p->emit_count) { c->err = "null emit_count"; return MCB_EINVAL; }
    long long tot = 0;
    for (int i = 0; i < c->nemitter; ++i) { if (p->emit_count[i] < 0) { c->err = "negative emit_count"; return MCB_EINVAL; } tot += p->emit_count[i]; }
    if (tot != p->nemit) { c->err = "nemit != sum(emit_count)"; return MCB_EINVAL; }
    if (p->maxloop < 0 || p->maxloop > MCB_MAX_LOOP) { c->err = "maxloop out of range (< 2^32 - 1: the loop trip is the 32-bit Philox event counter)"; return MCB_EINVAL; }
    if (p->maxscat < 0 || p->maxscat > 0x7FFFFFFFll) { c->err = "maxscat out of range"; return MCB_EINVAL; }
    if ((unsigned long long)p->nemit >> (64 - step_bits_for(p->maxloop))) { c->err = "nemit out of range (< 2^36; < 2^32 when maxloop >= 2^31)"; return MCB_EINVAL; }
    if (p->kind == MCB_PROB_CUMTEMP || p->kind == MCB_PROB_CUMFLUX) {
        // bin of a deposit = ceil(nscat_before / step) with nscat_before <= maxscat - 1 must be a row of the field (problem.cpp:562-563)
        const long long top = p->maxscat > 0 ? (p->maxscat - 1 + p->step - 1) / p->step : 0;
        if (top > p->size) { c->err = "Cum* problem: step too small for maxscat and size (bin index beyond the field rows)"; return MCB_EINVAL; }
    }
    return MCB_OK;
}

// payload rows per deposit: Temp/CumTemp 1 (dt), Flux/CumFlux 3 (dpos), Multi 4 (dt, dpos); one translation unit each
cudaError_t launch_step(const StepParams& P, int tm, int nd, int box, int pad, int grid, int block, size_t smem, cudaStream_t s) {
    switch (P.kind) {
    case MCB_PROB_TEMP: case MCB_PROB_CUMTEMP: return mcb::launch_step_n1(P, tm, nd, box, pad, grid, block, smem, s);
    case MCB_PROB_FLUX: case MCB_PROB_CUMFLUX: return mcb::launch_step_n3(P, tm, nd, box, pad, grid, block, smem, s);
    default: return mcb::launch_step_n4(P, tm, nd, box, pad, grid, block, smem, s);
    }
}

int ensure_slots(mcb_ctx* c, long long slots) {
    slots = (slots + 31) / 32 * 32;          // whole warps
    if (slots <= c->slots_alloc) return MCB_OK;
    for (int w = 0; w < 2; ++w) CUDA_TRY(c, c->state[w].alloc(state_bytes(slots)));
    c->slots_alloc = slots;
    return MCB_OK;
}

struct RunPlan {
    long long slots; int S, block, grid; int tm, copies; size_t smem;
    int ndm;                            // k_step's NDM: 0 no N-D grid, 1 serial walk, 2 + cooperative pieces, 3 warp-balanced items
    uint32_t nd_off, nd_warp_bytes;     // NDM 3: the per-warp record areas
    // 1-D difference-array histograms (tm == MCB_TM_WARP, no N-D grid): padded columns (0 = run-time stride), plane stride,
    // warps per histogram group, flush interval in loop trips, offset of the flush scratch
    int t1d, pad, trips; uint32_t ps, inst_bytes, scratch_off, stage_off, wbar_off;
    uint32_t cstride;                   // MCB_TM_BLOCK: bytes per histogram column
};

#define MCB_FX_FLUSH_TRIPS 16
#ifndef MCB_COOP_ND_CELLS
#define MCB_COOP_ND_CELLS 64      // grids at least this fine along an axis take the warp-cooperative N-D walk
#endif
#ifndef MCB_ND_BALANCED
#define MCB_ND_BALANCED 1      // N-D tally grids: the warp-balanced item walk (k_step NDM 3) instead of the per-lane serial walk
#endif
#define MCB_TILES_MIN 32
#define MCB_TILES_MAX 96
#ifndef MCB_DECAY_S
#define MCB_DECAY_S 16         // loop trips per launch once nothing is left to emit (before the hazard is known)
#endif
#ifndef MCB_DECAY_ADAPT_PCT
#define MCB_DECAY_ADAPT_PCT 30 // decay phase: loop trips per launch chosen so that about this share of the live phonons terminates (0: fixed)
#endif
#define MCB_DECAY_S_MIN 4
#define MCB_DECAY_S_MAX 64
// decay phase with the compaction fused into k_step (the default): every launch stores its survivors densely, so a launch may
// be short -- about this share of the live phonons terminates per launch (the lanes of a tile are 1 - share / 2 occupied on
// average) -- as long as it still carries enough loop trips to dilute its fixed costs (table staging, flush, launch gap)
#ifndef MCB_DECAY_FUSED_PCT
#define MCB_DECAY_FUSED_PCT 12
#endif
#define MCB_DECAY_FUSED_S_MIN 2
#ifndef MCB_DECAY_MIN_WORK
#define MCB_DECAY_MIN_WORK 4000000ll   // phonon-steps per launch (~150 us)
#endif
#ifndef MCB_COMPACT_PCT
#define MCB_COMPACT_PCT 90
#endif

int plan_run(mcb_ctx* c, const mcb_problem_desc* prob, long long nparticles, RunPlan* r) {
    const mcb_options& o = c->opt;
    const int per_sm = o.ctas_per_sm > 0 ? o.ctas_per_sm : 1;
    size_t base = 16 + (size_t)c->mv.bytes + (size_t)c->gv.bytes;
    const size_t budget = c->smem_optin / (size_t)per_sm > 1024 ? c->smem_optin / (size_t)per_sm - 1024 : 0;
    // CTA shapes (__launch_bounds__ of the k_step instances): 1-D tallies 768 threads x 80 registers; the N-D walks are built for
    // fewer, fatter threads
    r->ndm = c->any_nd; r->nd_off = 0; r->nd_warp_bytes = 0;
#if MCB_ND_BALANCED
    int nd3_block = 0;
    if (c->any_nd && c->nd_max_cells <= MCB_NDB_MAX_CELLS) {
        // per-warp record areas (very fine grids need many mark words): take the largest CTA that still leaves room for one copy
        // of the CTA histogram, else the largest that fits at all (the tally then goes to the global field)
        const size_t wb = (size_t)(MCB_NDB_FIXED + 4 * (c->nd_mark_words + 2) + 15) / 16 * 16;
        const size_t hist1 = o.tally_mode == 2 ? 0 : (size_t)c->cols * 4u * (size_t)((3 * prob->rows) | 1);
        const int top = o.block > 0 ? std::min(o.block, MCB_BLOCK_MAX_ND3) : MCB_BLOCK_MAX_ND3;
        for (int pass = 0; pass < 2 && !nd3_block; ++pass)
            for (int blk = top; blk >= 256 && !nd3_block; blk -= 128)
                if (base + 16 + (size_t)(blk / 32) * wb + (pass == 0 ? hist1 : 0) <= budget) nd3_block = blk;
        if (nd3_block) { r->ndm = 3; r->nd_warp_bytes = (uint32_t)wb; }
    }
#endif
    int block_max = c->any_nd == 2 ? MCB_BLOCK_MAX_ND : (c->any_nd == 1 ? MCB_BLOCK_MAX_ND1 : MCB_BLOCK_MAX);
#if MCB_ND_BALANCED
    if (r->ndm == 3) block_max = nd3_block;
#endif
    r->block = o.block > 0 ? std::min(o.block, block_max) : block_max;
    if (r->block % 32 != 0 || o.block > 1024) { c->err = "block must be a multiple of 32, <= 1024"; return MCB_EINVAL; }
    r->S = o.steps_per_launch > 0 ? o.steps_per_launch : 16;
    // streaming schedule: 32 ... 96 tiles per warp, about a third of the solve's phonons (3.6 M ... 10.9 M resident phonons for the
    // 1-D kernels).  Long launches dilute a launch's fixed costs (table staging, flush, emission tail: C3 at 1e8 phonons +9 % with
    // 96 tiles instead of 32), while a population close to the whole problem would leave hardly any steady launches.
    long long slots = o.slots;
    if (slots <= 0) {
        const long long per_tile = (long long)c->sm_count * r->block;
        const long long t = (std::max<long long>(nparticles, 1) / 3 + per_tile - 1) / per_tile;
        slots = per_tile * std::min<long long>(MCB_TILES_MAX, std::max<long long>(MCB_TILES_MIN, t));
    }
    if (o.slots <= 0 && o.steps_per_launch <= 0) {
        // The library's default schedule (no options set) keeps EVERY phonon of the solve resident when the two state buffers fit
        // in half of the device memory that is available: the first fill emits them all with dense lanes and every launch then
        // runs S = 16 loop trips per state round trip on a population that only decays (C1 +30 %, C3 / C5 +35 ... +50 % over a
        // 3.6-M-slot population that is refilled while it streams).  An explicit slots / steps_per_launch keeps the streaming
        // schedule the HBM roofline is quoted on.
        // (the memory query is skipped when the buffers this context already holds are large enough: cudaMemGetInfo was measured
        // to take 20-30 ms every few calls, as much as a whole C2 solve)
        size_t free_b = 0, total_b = 0;
        if (nparticles <= c->slots_alloc) slots = std::max(slots, (long long)nparticles);
        else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t have = 2 * state_bytes(c->slots_alloc);
            const long long cap = (long long)((free_b + have) / 2 / (2 * (MCB_GROUP_BYTES / 32)));
            slots = std::max(slots, std::min<long long>(nparticles, cap));
        }
        cudaGetLastError();
    }
    slots = std::min(slots, std::max<long long>(nparticles, 1));
    if (slots > 0x7FFFFF00ll) { c->err = "too many resident slots (< 2^31)"; return MCB_ELIMIT; }
    r->slots = slots;
    const long long tiles = (slots + r->block - 1) / r->block;
    r->grid = (int)std::min<long long>((long long)c->sm_count * per_sm, std::max<long long>(tiles, 1));
    const size_t nwarps = (size_t)r->block / 32;
    if (r->ndm == 3) {                  // the record areas sit between the tables and the histogram
        r->nd_off = (uint32_t)((base + 15) / 16 * 16);
        base = r->nd_off + nwarps * r->nd_warp_bytes;
    }
    r->t1d = 0; r->pad = 0; r->trips = MCB_FX_FLUSH_TRIPS; r->ps = 0; r->inst_bytes = 0; r->scratch_off = 0; r->stage_off = 0; r->wbar_off = 0;
    r->copies = 1; r->cstride = 0; r->tm = MCB_TM_GLOBAL;
    const size_t stage = nwarps * (MCB_GROUP_BYTES + 8) + 128;       // per-warp staging buffers of the TMA state prefetch + mbarriers
    auto place_stage = [&](size_t end) {                               // staging area behind everything else, if it fits
        if (end + stage > budget) return end;
        r->stage_off = (uint32_t)((end + 127) / 128 * 128); r->wbar_off = r->stage_off + (uint32_t)(nwarps * MCB_GROUP_BYTES);
        return (size_t)r->wbar_off + nwarps * 8;
    };
    const long long fdep = r->ndm == 2 ? 3 : 1;                     // deposits of one flight into one entry (cooperative N-D walk: <= 3)
    // (1) 1-D / single-cell tallies: difference-array histograms (mcb_device.cuh: deposit_fx).  An instance holds
    // [direct | difference][row][limb 0 | 1 | 2] planes of `ps` bytes and is shared by the whole CTA; up to 4 copies (picked by
    // lane id) thin out same-word hits inside a warp instruction.  The state staging buffers come first, copies take the rest.
    if (c->any_nd == 0 && c->t1_ok && o.tally_mode != 2 && o.tally_mode != 3) {
        const int pad = c->cols <= 32 ? 32 : (c->cols <= 128 ? 128 : (c->cols <= 512 ? 512 : 0));
        const uint32_t ps = 4u * (uint32_t)(pad > 0 ? pad : (int)((c->cols + 31) / 32 * 32));
        const size_t inst = (size_t)prob->rows * 2u * MCB_T1D_LIMBS * ps;
        const size_t scratch = ((size_t)prob->rows * (size_t)c->cols * 2 + (size_t)prob->rows * (size_t)((c->cols + 31) / 32)) * 8 + 16;
        for (int st = 1; st >= 0 && !r->t1d; --st)
            for (int cp = 4; cp >= 1; cp >>= 1)
                if (base + inst * (size_t)cp + scratch + (st ? stage : 0) <= budget) {
                    r->t1d = 1; r->pad = c->all_box ? pad : 0; r->copies = cp; r->ps = ps; r->inst_bytes = (uint32_t)inst;
                    r->tm = MCB_TM_WARP;
                    r->scratch_off = (uint32_t)((base + inst * (size_t)cp + 15) / 16 * 16);
                    r->smem = place_stage(r->scratch_off + scratch);
                    break;
                }
    }
    // (2) any grid: one three-limb histogram per CTA (mcb_device.cuh: deposit), walked cell by cell; up to 4 copies by lane id
    if (!r->t1d && o.tally_mode != 2) {
        const uint32_t cstride = 4u * (uint32_t)((3 * prob->rows) | 1);
        const size_t inst = (size_t)c->cols * cstride;
        for (int cp = 4; cp >= 1; cp >>= 1)
            if (base + inst * (size_t)cp <= budget && inst < 0xFFFFFFFFull) {
                r->tm = MCB_TM_BLOCK; r->copies = cp; r->cstride = cstride; r->inst_bytes = (uint32_t)inst;
                r->smem = place_stage(base + inst * (size_t)cp);
                break;
            }
        if (r->tm != MCB_TM_BLOCK && (o.tally_mode == 1 || o.tally_mode == 3)) { c->err = "tally_mode: the shared-memory histogram does not fit"; return MCB_ELIMIT; }
    }
    // (3) the global field in L2 (fp64 RED)
    if (r->tm == MCB_TM_GLOBAL) r->smem = place_stage(base);
    if (r->tm != MCB_TM_GLOBAL)          // an entry receives at most (threads sharing the copy) x fdep deposits per loop trip; 2^16 between two flushes
        r->trips = (int)std::max<long long>(1, std::min<long long>(256, 65536ll / (std::max(1, r->block / r->copies) * fdep)));
    if (r->smem > c->smem_optin) { c->err = "material + geometry tables exceed the shared-memory staging area"; return MCB_ELIMIT; }
    return MCB_OK;
}

void fill_params(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, StepParams* P) {
    std::memset(P, 0, sizeof *P);
    P->mat_blob = c->mat_blob.p; P->mv = c->mv; P->geo_blob = c->geo_blob.p; P->gv = c->gv;
    P->emitters = c->emitters.p; P->emit_cdf = c->emit_cdf.p; P->nemitter = c->nemitter;
    P->f_wprob = c->f_wprob.p; P->f_pprob = c->f_pprob.p; P->f_walias = c->f_walias.p; P->f_palias = c->f_palias.p;
    P->kind = prob->kind; P->rows = prob->rows; P->cols = (int32_t)c->cols; P->cum_step = prob->step > 0 ? prob->step : 1;
    P->maxscat = prob->maxscat; P->maxloop = prob->maxloop; P->seed = seed;
    P->maxscat32 = (uint32_t)prob->maxscat; P->maxloop32 = (uint32_t)prob->maxloop;
    P->step_bits = step_bits_for(prob->maxloop); P->step_mask = P->step_bits >= 32 ? 0xFFFFFFFFu : (1u << P->step_bits) - 1u;
    for (int r = 0; r < 10; ++r) { P->rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u; P->rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u; }
    P->ctr = c->ctr.p; P->field_len = (long long)prob->rows * c->cols;
    P->so_mat = 16; P->so_geo = 16 + c->mv.bytes; P->so_hist = 16 + c->mv.bytes + c->gv.bytes;
    P->so_lambda = P->so_mat + c->mv.off_lambda; P->so_inv_vel = P->so_mat + c->mv.off_inv_vel; P->so_wprob = P->so_mat + c->mv.off_wprob;
    P->so_pprob = P->so_mat + c->mv.off_pprob; P->so_walias = P->so_mat + c->mv.off_walias; P->so_palias = P->so_mat + c->mv.off_palias;
    P->so_hot = P->so_geo + c->gv.off_hot; P->so_cold = P->so_geo + c->gv.off_cold; P->so_sdom = P->so_geo + c->gv.off_sdom; P->so_pairs = P->so_geo + c->gv.off_pairs;
    P->hist_copies = 1;
}
void apply_plan(const RunPlan& plan, const mcb_problem_desc* prob, StepParams* P) {
    P->tally_smem = plan.tm; P->hist_copies = plan.copies;
    P->fx_ps = plan.ps; P->fx_diff_off = (uint32_t)prob->rows * MCB_T1D_LIMBS * plan.ps; P->hist_bytes = plan.inst_bytes; P->so_scratch = plan.scratch_off;
    P->so_stage = plan.stage_off; P->so_wbar = plan.wbar_off; P->fx_cstride = plan.cstride;
    P->so_nd = plan.nd_off; P->nd_warp_bytes = plan.nd_warp_bytes;
    if (plan.ndm == 3) P->so_hist = (uint32_t)(plan.nd_off + (size_t)(plan.block / 32) * plan.nd_warp_bytes);
}

// Fixed-point scale of the shared-memory tallies (mcb_device.cuh: deposit, deposit_fx).  Payload component k of a flight is
// accepted up to fx_max = 2^E > flight_max (* the 99.9 % slowness for the dt row; anything larger takes the exact fp64 path)
// and deposited as q = rint(v 2^(Q-1-E)), |q| <= 2^(Q-1), in three carry-free limbs.  A histogram is flushed at least every
// `trips` loop trips (k_step flushes between tiles and the schedule keeps steps_per_launch <= trips), so an entry receives at
// most N = (threads sharing the copy) x trips x f <= 2^n <= 2^16 deposits between flushes (f = 1 deposit per flight and cell;
// 3 for the cooperative N-D walk, whose pieces can share a cell): the 16-bit limb fields sum below 2^32 and the top limb
// stays inside int32 for Q = 64 - n, capped at 50 (the rounding trick holds |q| < 2^51): quantum 2^-47 .. 2^-49 of fx_max.
void set_fixed_point(mcb_ctx* c, const mcb_problem_desc* prob, const RunPlan& plan, StepParams* P) {
    const long long bound = (long long)std::max(1, plan.block / plan.copies) * plan.trips * (plan.ndm == 2 ? 3 : 1);
    int n = 0; while ((1ll << n) < bound) ++n;                                             // bound <= 2^n <= 2^16
    const int QB = std::min(50, 64 - n);
    P->fx_limb_bits = 16;
    for (int k = 0; k < 4; ++k) {
        const bool is_dt = (prob->kind == MCB_PROB_TEMP || prob->kind == MCB_PROB_CUMTEMP || prob->kind == MCB_PROB_MULTI) && k == 0;
        const double amax = c->flight_max * (is_dt ? c->inv_vel_fx : 1.0);
        int e = 0; std::frexp(amax > 0.0 ? amax : 1.0, &e);                                // amax < 2^e
        const int E = e;
        P->fx_max[k] = std::ldexp(1.0, E); P->fx_scale[k] = std::ldexp(1.0, QB - 1 - E); P->fx_inv[k] = std::ldexp(1.0, E + 1 - QB);
    }
    P->fx_flush_trips = plan.trips;
}

int upload_cdf(mcb_ctx* c, const mcb_problem_desc* prob) {
    std::vector<long long> cdf(c->nemitter);
    long long acc = 0;
    for (int i = 0; i < c->nemitter; ++i) { acc += prob->emit_count[i]; cdf[i] = acc; }     // problem.cpp:374-378
    CUDA_TRY(c, c->emit_cdf.alloc(cdf.size()));
    CUDA_TRY(c, cudaMemcpyAsync(c->emit_cdf.p, cdf.data(), cdf.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MCB_OK;
}

// K3: move the active slots of state[cur] (the first `nslots`) to the front of state[other], whose first `bound` slots are
// cleared first (bound >= number of active slots).  sorted = false: unordered stream compaction (k_compact); true: counting
// sort by (subdomain, tally cell) bin (k_sort_count / k_sort_scan / k_sort_scatter, mcb_kernels.cuh).
void sort_geometry(const mcb_ctx* c, uint32_t* cols_per_bin, uint32_t* nbins) {
    const long long cols = std::max<long long>(c->cols, 1);
    *cols_per_bin = (uint32_t)((cols + MCB_SORT_MAX_BINS - 1) / MCB_SORT_MAX_BINS);
    *nbins = (uint32_t)((cols + *cols_per_bin - 1) / *cols_per_bin);
}
int compact_slots(mcb_ctx* c, int cur, long long nslots, long long bound, bool sorted, long long* launches) {
    const int other = cur ^ 1;
    const unsigned grid = (unsigned)std::min<long long>((nslots + 255) / 256, (long long)c->sm_count * 8);
    CUDA_TRY(c, cudaMemsetAsync(&c->ctr.p->compact_cursor, 0, sizeof(unsigned long long), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->state[other].p, 0, state_bytes(bound), c->stream));
    if (!sorted) {
        k_compact<<<grid, 256, 0, c->stream>>>(view_of(c, cur), view_of(c, other), nslots, c->ctr.p);
        CUDA_TRY(c, cudaGetLastError());
        *launches += 1;
        return MCB_OK;
    }
    uint32_t cpb, nbins; sort_geometry(c, &cpb, &nbins);
    CUDA_TRY(c, c->sort_keys.alloc((size_t)nslots)); CUDA_TRY(c, c->sort_bins.alloc(nbins));
    if (nbins * 4u > 48u * 1024u) CUDA_TRY(c, cudaFuncSetAttribute(k_sort_count, cudaFuncAttributeMaxDynamicSharedMemorySize, MCB_SORT_MAX_BINS * 4));
    CUDA_TRY(c, cudaMemsetAsync(c->sort_bins.p, 0, (size_t)nbins * 4, c->stream));
    k_sort_count<<<grid, 256, (size_t)nbins * 4, c->stream>>>(view_of(c, cur), nslots, c->geo_blob.p, c->gv, cpb, nbins, c->sort_keys.p, c->sort_bins.p);
    k_sort_scan<<<1, 1024, 0, c->stream>>>(c->sort_bins.p, nbins, &c->ctr.p->compact_cursor);
    k_sort_scatter<<<grid, 256, 0, c->stream>>>(view_of(c, cur), view_of(c, other), nslots, c->sort_keys.p, c->sort_bins.p);
    CUDA_TRY(c, cudaGetLastError());
    *launches += 3;
    return MCB_OK;
}

// The schedule: launches of k_step over the resident slots until every particle of
// [n_begin, n_end) has been emitted and has terminated; the tail is compacted.
int run_solve(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end,
              double* raw_field_dev, mcb_stats* stats) {
    int rc = check_problem(c, prob);
    if (rc) return rc;
    if (n_begin < 0 || n_end > prob->nemit || n_begin > n_end) { c->err = "bad particle range"; return MCB_EINVAL; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    RunPlan plan;
    rc = plan_run(c, prob, n_end - n_begin, &plan);
    if (rc) return rc;
    rc = ensure_slots(c, plan.slots);
    if (rc) return rc;
    rc = upload_cdf(c, prob);
    if (rc) return rc;

    StepParams P; fill_params(c, prob, seed, &P);
    P.field = raw_field_dev; P.do_tally = 1;
    P.steps_per_launch = plan.S; P.n_end = (unsigned long long)n_end;
    apply_plan(plan, prob, &P);
    set_fixed_point(c, prob, plan, &P);

    Counters init{}; init.next = (unsigned long long)n_begin;
    CUDA_TRY(c, cudaMemcpyAsync(c->ctr.p, &init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    int cur = 0; long long nslots = plan.slots;
    CUDA_TRY(c, cudaMemsetAsync(c->state[cur].p, 0, state_bytes(nslots), c->stream));        // every slot inactive

    long long launches = 0, step_launches = 0, slot_steps = 0;
    CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
    float step_ms_total = 0.f;
    const long long total = n_end - n_begin;
    // The host runs one launch AHEAD of the counters it reads (depth-2 ring of events + pinned counters), so the
    // GPU always has the next launch queued.  Decisions (stop, compaction, tail mode) use the counters of the
    // previous launch; after the last particle dies one extra, empty launch has already been queued.
    const long long tail_slots = (long long)plan.grid * plan.block;      // one tile per CTA: finish in one launch
    int S_cur = plan.S;
    if (MCB_DECAY_ADAPT_PCT > 0 && c->opt.steps_per_launch <= 0 && nslots >= total) S_cur = MCB_DECAY_S_MIN;   // everything resident: start short, then adapt
    bool host_all_emitted = false;
    unsigned long long last_live = (unsigned long long)nslots;
    {   // fused emission: one free-slot segment per k_step warp, long enough for every slot the warp visits
        const long long tiles0 = (nslots + plan.block - 1) / plan.block;
        const long long grid0 = std::min<long long>(plan.grid, std::max<long long>(tiles0, 1));
        P.free_seg = (uint32_t)(((tiles0 + grid0 - 1) / grid0) * 32);
        if (nslots < total) {               // (every phonon resident after the first fill: nothing is ever refilled)
            CUDA_TRY(c, c->free_list.alloc((size_t)(grid0 * (plan.block / 32)) * P.free_seg));
            P.free_list = c->free_list.p;
        }
    }
    const long long compact_pct = c->opt.compact_pct > 0 ? c->opt.compact_pct : MCB_COMPACT_PCT;
    // decay_mode 0 (default): K3 is fused into k_step -- once nothing is left to emit every launch stores its survivors densely
    // into the other state buffer (StepParams::compact) and the launch behind it reads the slot count from device memory, so the
    // host's one-launch-old counters never leave dead lanes in a tile.  Modes 1 / 2 (and the sorting compaction) keep the separate
    // K3 pass of round 2: a launch then visits the survivors of the launch BEFORE the previous one.
    const bool fused = c->opt.decay_mode == 0 && c->opt.sort_mode == 0;
    const int decay_pct = c->opt.decay_pct > 0 ? c->opt.decay_pct : (fused ? MCB_DECAY_FUSED_PCT : MCB_DECAY_ADAPT_PCT);
    bool prev_compacting = false, any_compacted = false, quiet = true;
    // tail accounting: launches that run the last survivors (at most one tile per CTA) to termination are bound by the
    // length of the longest history, not by throughput (mcb_stats::tail_ms / tail_steps)
    bool tail_launch[2] = {false, false}, tail_next = false; float tail_ms = 0.f; unsigned long long tail_steps = 0;
    static const bool launch_log = std::getenv("MCB_LAUNCH_LOG") != nullptr;      // per-launch trace on stderr (diagnostic)
    long long log_nslots[2] = {0, 0}; int log_S[2] = {0, 0};
    long long steady_launches = 0, compactions = 0, sorts = 0, sorted_pop = nslots; float steady_ms = 0.f;
    unsigned long long steady_steps = 0, steady_stores = 0, prev_steps = 0, prev_stores = 0;
    if (total > 0) for (long long it = 0;; ++it) {
        const int slot = (int)(it & 1);
        // fixed-point histograms are flushed between tiles: keep a tile's loop trips within the flush interval
        if (plan.tm != MCB_TM_GLOBAL) S_cur = std::min(S_cur, plan.trips);
        P.st = view_of(c, cur); P.nslots = nslots; P.steps_per_launch = S_cur; P.parity = slot; P.host_ctr = c->d_hctr + slot;
        const long long tiles = (nslots + plan.block - 1) / plan.block;
        const int grid = (int)std::min<long long>(plan.grid, std::max<long long>(tiles, 1));
        // K1: the first fill (every slot free) is its own dense kernel; afterwards k_step refills the slots that end inactive
        // itself (32 at a time, ids from the atomic cursor Counters::next) while particles are left to emit
        P.emit_enable = (host_all_emitted || !P.free_list) ? 0 : 1;
        // K3 fused: once nothing is left to emit the launch stores its survivors densely into the other buffer -- unless (almost)
        // nobody terminated in the last completed launch, or none has completed yet (a population that ends by maxscat loses
        // nobody for its first maxscat loop trips): such a launch stores in place and spares its tiles the wait for the cursor
        const bool compacting = fused && !P.emit_enable && !quiet;
        P.compact = compacting ? 1 : 0; P.st_out = view_of(c, cur ^ 1); P.use_dev_n = any_compacted ? 1 : 0;
        if (it == 0) {
            k_emit<<<(unsigned)(c->sm_count * 8), 256, 0, c->stream>>>(P);
            k_emit_commit<<<1, 32, 0, c->stream>>>(P);
            CUDA_TRY(c, cudaGetLastError());
            launches += 2;
        }
        tail_launch[slot] = tail_next; log_nslots[slot] = nslots; log_S[slot] = S_cur;
        CUDA_TRY(c, cudaEventRecord(c->evA[slot], c->stream));
        CUDA_TRY(c, launch_step(P, plan.tm, plan.ndm, c->all_box, plan.pad, grid, plan.block, plan.smem, c->stream));
        CUDA_TRY(c, cudaEventRecord(c->evB[slot], c->stream));
        // the launch's last CTA has mirrored the counters into h_ctr[slot] (StepParams::host_ctr): evB is all the host waits for
        launches++; step_launches++; slot_steps += nslots * (long long)std::min<long long>(S_cur, prob->maxloop);
        const bool before_compacted = prev_compacting;            // launch it-1
        prev_compacting = compacting;
        if (compacting) { cur ^= 1; any_compacted = true; }        // the launches behind this one read what this one writes
        if (it == 0) continue;
        const int prev = slot ^ 1;
        CUDA_TRY(c, cudaEventSynchronize(c->evB[prev]));
        float ms = 0.f; cudaEventElapsedTime(&ms, c->evA[prev], c->evB[prev]); step_ms_total += ms;
        const unsigned long long live = c->h_ctr[prev].live[prev], next = c->h_ctr[prev].next;     // launch it-1 had parity `prev`
        const bool all_emitted = next >= (unsigned long long)n_end;
        host_all_emitted = all_emitted;
        if (!all_emitted) {          // launch it-1 ran with a full population: steady-phase accounting
            steady_launches++; steady_ms += ms;
            steady_steps += c->h_ctr[prev].steps - prev_steps; steady_stores += c->h_ctr[prev].stores - prev_stores;
        }
        // hazard of the population in launch it-1 (fraction of the live phonons that terminated per loop trip): sets the decay S below
        const unsigned long long dsteps = c->h_ctr[prev].steps - prev_steps;
        if (tail_launch[prev]) { tail_ms += ms; tail_steps += dsteps; }
        if (launch_log) std::fprintf(stderr, "mcb launch %lld: slots<=%lld S=%d %.3f ms, %llu steps, live after %llu%s\n", it - 1, log_nslots[prev], log_S[prev],
                                     ms, dsteps, live, tail_launch[prev] ? " (tail)" : "");
        double hazard = -1.0;
        if (all_emitted && last_live >= live && dsteps > 0) hazard = (double)(last_live - live) / (double)dsteps;
        last_live = live;
        prev_steps = c->h_ctr[prev].steps; prev_stores = c->h_ctr[prev].stores;
        if (all_emitted && live == 0) {
            CUDA_TRY(c, cudaEventSynchronize(c->evB[slot]));
            cudaEventElapsedTime(&ms, c->evA[slot], c->evB[slot]); step_ms_total += ms;
            break;
        }
        // fused K3: launch it-1 stored exactly its `live` survivors, launch `it` (queued above) visits them, and the launch queued
        // next visits the survivors of launch `it`: at most `live` (the kernel takes the exact count from device memory)
        if (before_compacted && (long long)live < nslots) nslots = std::max<long long>((long long)live, 1);
        if (!fused && all_emitted && (long long)live * 100 < nslots * compact_pct && nslots > tail_slots) {
            // tail: compact the survivors so later launches stream only live state.  `live` is one launch old,
            // i.e. an upper bound (nothing is emitted any more); unused destination slots stay inactive.
            // sort_mode: the compaction is a counting sort by (subdomain, tally cell) -- always (1), or whenever the
            // population has shrunk by the factor sort_mode since the last sort
            const long long bound = std::max<long long>((long long)live, 1);
            const int sm = c->opt.sort_mode;
            const bool sorted = sm == 1 || (sm >= 2 && bound * sm <= sorted_pop);
            if ((rc = compact_slots(c, cur, nslots, bound, sorted, &launches))) return rc;
            compactions++;
            if (sorted) { sorts++; sorted_pop = bound; }
            cur ^= 1; nslots = bound;
        }
        // decay phase (nothing left to emit): the population only shrinks, so run several loop trips per state round trip;
        // once the survivors fit one tile per CTA let every thread run its phonon to termination
        if (all_emitted && c->opt.decay_mode != 1) {
            S_cur = std::max(plan.S, MCB_DECAY_S);
#if MCB_DECAY_ADAPT_PCT > 0
            // a young population terminates fast (C2: 6 % per loop trip -- 16 trips would leave 37 % of the lanes alive), an old
            // one slowly: run as many loop trips per state round trip as let about decay_pct % of the phonons terminate (12 with the
            // fused compaction, 30 with separate passes), but enough of them to dilute the launch's fixed costs
            const int s_min = fused ? MCB_DECAY_FUSED_S_MIN : MCB_DECAY_S_MIN;
            if (hazard > 0.0) S_cur = (int)std::min<double>(MCB_DECAY_S_MAX, std::max<double>(s_min, std::floor(0.01 * decay_pct / hazard + 0.5)));
            else if (hazard == 0.0) S_cur = MCB_DECAY_S_MAX;
            if (fused && live > 0) S_cur = (int)std::min<long long>(MCB_DECAY_S_MAX, std::max<long long>(S_cur, (MCB_DECAY_MIN_WORK + (long long)live - 1) / (long long)live));
#endif
            // a launch compacts only when the last completed launch says it will lose at least ~1 % of its phonons: every tile of
            // a compacting launch waits for one atomic on a single cursor (1e7 phonons at S = 4: 3e5 same-address atomics, ~0.6 ms)
            quiet = !(hazard > 0.0 && hazard * S_cur >= 0.01) && (long long)live * 100 >= nslots * 95;     // (slow losses add up: compact below 95 %)
            if ((long long)live <= tail_slots) { S_cur = (int)std::min<long long>(std::max<long long>(prob->maxloop, 1), 1 << 22); tail_next = true; }
        }
    }
    CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(&c->h_ctr[0], c->ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        float ms = 0.f; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        stats->emitted = (int64_t)c->h_ctr[0].emitted; stats->steps = (int64_t)c->h_ctr[0].steps; stats->esc = (int64_t)c->h_ctr[0].esc;
        stats->launches = launches; stats->cols = c->cols; stats->device_ms = ms; stats->step_ms = step_ms_total;
        stats->step_launches = step_launches; stats->slot_steps = slot_steps; stats->state_stores = (int64_t)c->h_ctr[0].stores;
        stats->steady_launches = steady_launches; stats->steady_steps = (int64_t)steady_steps;
        stats->steady_stores = (int64_t)steady_stores; stats->steady_ms = steady_ms;
        stats->compactions = compactions; stats->sorts = sorts; stats->tail_ms = tail_ms; stats->tail_steps = (int64_t)tail_steps;
    }
    return MCB_OK;
}

} // namespace

// =============================================================================== C ABI
extern "C" {

int mcb_abi_version(void) { return MCB_ABI_VERSION; }

#ifndef MCB_SRC_HASH
#define MCB_SRC_HASH "unknown"
#endif
#ifndef MCB_EXTRA_FLAGS
#define MCB_EXTRA_FLAGS ""
#endif
const char* mcb_build_info(void) { return MCB_SRC_HASH "|" MCB_EXTRA_FLAGS; }

const char* mcb_last_error(const mcb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int mcb_create(int device, mcb_ctx** out) {
    if (!out) { g_create_err = "null out pointer"; return MCB_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                       " (there is no CPU fallback)";
        return MCB_ENODEVICE;
    }
    if (device < 0 || device >= ndev) { g_create_err = "device ordinal out of range"; return MCB_EINVAL; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return MCB_ECUDA; }
    if (prop.major != 10) {
        g_create_err = std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                       "; this library is built for sm_100a only";
        return MCB_ENODEVICE;
    }
    mcb_ctx* c = new mcb_ctx;
    c->device = device; c->sm_count = prop.multiProcessorCount; c->smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&c->evA[0])) != cudaSuccess || (e = cudaEventCreate(&c->evA[1])) != cudaSuccess ||
        (e = cudaEventCreate(&c->evB[0])) != cudaSuccess || (e = cudaEventCreate(&c->evB[1])) != cudaSuccess ||
        (e = cudaEventCreate(&c->evC[0])) != cudaSuccess || (e = cudaEventCreate(&c->evC[1])) != cudaSuccess ||
        (e = c->ctr.alloc(1)) != cudaSuccess || (e = cudaHostAlloc(&c->h_ctr, 2 * sizeof(Counters), cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer(&c->d_hctr, c->h_ctr, 0)) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e); delete c; return MCB_ECUDA;
    }
    *out = c;
    return MCB_OK;
}

void mcb_destroy(mcb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->mat_blob.release(); c->f_wprob.release(); c->f_pprob.release(); c->f_walias.release(); c->f_palias.release();
    c->geo_blob.release(); c->emitters.release(); c->cell_vol.release(); c->emit_cdf.release();
    for (int w = 0; w < 2; ++w) c->state[w].release();
    c->ctr.release(); c->field.release(); c->free_list.release(); c->sort_keys.release(); c->sort_bins.release();
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    if (c->ev0) cudaEventDestroy(c->ev0); if (c->ev1) cudaEventDestroy(c->ev1);
    for (int k = 0; k < 2; ++k) { if (c->evA[k]) cudaEventDestroy(c->evA[k]); if (c->evB[k]) cudaEventDestroy(c->evB[k]); if (c->evC[k]) cudaEventDestroy(c->evC[k]); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int mcb_set_options(mcb_ctx* c, const mcb_options* o) {
    if (!c || !o) return MCB_EINVAL;
    if (o->slots < 0 || o->steps_per_launch < 0 || o->block < 0 || o->ctas_per_sm < 0 || o->tally_mode < 0 || o->tally_mode > 3 || o->decay_mode < 0 || o->decay_mode > 2 || o->decay_pct < 0 || o->decay_pct > 100 || o->emit_mode != 0 || o->compact_pct < 0 || o->compact_pct > 100 || o->sort_mode < 0) {
        c->err = "negative / unknown option"; return MCB_EINVAL;
    }
    c->opt = *o;
    return MCB_OK;
}
int mcb_get_options(const mcb_ctx* c, mcb_options* o) { if (!c || !o) return MCB_EINVAL; *o = c->opt; return MCB_OK; }

int mcb_stream(const mcb_ctx* c, void** s) { if (!c || !s) return MCB_EINVAL; *s = (void*)c->stream; return MCB_OK; }

int mcb_upload_material(mcb_ctx* c, const mcb_material_desc* m) {
    if (!c) return MCB_EINVAL;
    if (!m || !m->vel || !m->tau || !m->flux_pdf || !m->scat_pdf) { c->err = "null material table"; return MCB_EINVAL; }
    if (m->nw <= 0 || m->np <= 0) { c->err = "Invalid dispersion table (nw, np must be > 0)"; return MCB_EINVAL; }
    if (m->nw > 65535 || m->np > 255 || m->nw * m->np >= MCB_MAX_WP) { c->err = "material table too large (nw<=65535, np<=255)"; return MCB_ELIMIT; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const long nw = m->nw, np = m->np, n = nw * np;
    for (long i = 0; i < n; ++i)
        if (!(m->vel[i] > 0.0) || !std::isfinite(m->tau[i]) || !(m->tau[i] >= 0.0)) { c->err = "vel must be > 0 and tau finite, >= 0"; return MCB_EINVAL; }
    c->flux_alias.build(m->flux_pdf, nw, np);
    c->scat_alias.build(m->scat_pdf, nw, np);
    MaterialView v{}; v.nw = (int32_t)nw; v.np = (int32_t)np;
    uint32_t off = 0;
    v.off_lambda = off;  off = align16(off + (uint32_t)(n * 8));
    v.off_inv_vel = off; off = align16(off + (uint32_t)(n * 8));
    v.off_wprob = off;   off = align16(off + (uint32_t)(nw * 8));
    v.off_pprob = off;   off = align16(off + (uint32_t)(n * 8));
    v.off_walias = off;  off = align16(off + (uint32_t)(nw * 2));
    v.off_palias = off;  off = align16(off + (uint32_t)(n * 1));
    v.bytes = off;
    auto bucket_of = [](uint32_t n) { uint32_t b = 0xFFFFFFFFu / n; if (0xFFFFFFFFu % n == n - 1u) ++b; return b; };   // uniform_int_distribution, random.h:25
    v.inv_bucket_w = 1.0 / (double)bucket_of((uint32_t)nw); v.inv_bucket_p = 1.0 / (double)bucket_of((uint32_t)np);
    if ((size_t)v.bytes + 16 > c->smem_optin) { c->err = "material tables exceed the shared-memory staging area"; return MCB_ELIMIT; }
    std::vector<unsigned char> blob(v.bytes, 0);
    double* lambda = reinterpret_cast<double*>(&blob[v.off_lambda]);
    double* inv_vel = reinterpret_cast<double*>(&blob[v.off_inv_vel]);
    double* wprob = reinterpret_cast<double*>(&blob[v.off_wprob]);
    double* pprob = reinterpret_cast<double*>(&blob[v.off_pprob]);
    uint16_t* walias = reinterpret_cast<uint16_t*>(&blob[v.off_walias]);
    uint8_t* palias = reinterpret_cast<uint8_t*>(&blob[v.off_palias]);
    for (long w = 0; w < nw; ++w) {
        wprob[w] = c->scat_alias.wprob[w]; walias[w] = (uint16_t)c->scat_alias.walias[w];
        for (long p = 0; p < np; ++p) {
            const long src = w + nw * p, dst = w * np + p;
            lambda[dst] = m->vel[src] * m->tau[src];          // vel(phn) * tau(phn)   material.cpp:221
            inv_vel[dst] = 1.0 / m->vel[src];
            pprob[dst] = c->scat_alias.pprob[dst]; palias[dst] = (uint8_t)c->scat_alias.palias[dst];
        }
    }
    {   // 99.9 % quantile of 1/vel over the modes as they are drawn (emission ~ fluxPdf, scattering ~ scatPdf)
        std::vector<std::pair<double, double>> sw((size_t)n);
        double fs = 0.0, ss = 0.0;
        for (long i = 0; i < n; ++i) { fs += m->flux_pdf[i]; ss += m->scat_pdf[i]; }
        for (long i = 0; i < n; ++i)
            sw[(size_t)i] = {1.0 / m->vel[i], (fs > 0.0 ? m->flux_pdf[i] / fs : 0.0) + (ss > 0.0 ? m->scat_pdf[i] / ss : 0.0)};
        std::sort(sw.begin(), sw.end(), [](const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first > b.first; });
        double tail = 0.0; c->inv_vel_fx = sw.back().first;
        for (const auto& e : sw) { tail += e.second; if (tail > 2e-3) { c->inv_vel_fx = e.first; break; } }
    }
    CUDA_TRY(c, c->mat_blob.alloc(v.bytes));
    CUDA_TRY(c, cudaMemcpy(c->mat_blob.p, blob.data(), v.bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(c, c->f_wprob.alloc(nw)); CUDA_TRY(c, c->f_walias.alloc(nw));
    CUDA_TRY(c, c->f_pprob.alloc(n));  CUDA_TRY(c, c->f_palias.alloc(n));
    CUDA_TRY(c, cudaMemcpy(c->f_wprob.p, c->flux_alias.wprob.data(), nw * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->f_walias.p, c->flux_alias.walias.data(), nw * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->f_pprob.p, c->flux_alias.pprob.data(), n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->f_palias.p, c->flux_alias.palias.data(), n * 4, cudaMemcpyHostToDevice));
    c->mv = v; c->nw = nw; c->np = np; c->energy_sum = m->energy_sum; c->flux_sum = m->flux_sum; c->has_mat = true;
    return MCB_OK;
}

int mcb_get_alias(const mcb_ctx* c, int which, double* wprob, int32_t* walias, double* pprob, int32_t* palias) {
    if (!c || !c->has_mat || which < 0 || which > 1) return MCB_EINVAL;
    const AliasTables& a = which == 0 ? c->flux_alias : c->scat_alias;
    std::memcpy(wprob, a.wprob.data(), a.wprob.size() * 8); std::memcpy(walias, a.walias.data(), a.walias.size() * 4);
    std::memcpy(pprob, a.pprob.data(), a.pprob.size() * 8); std::memcpy(palias, a.palias.data(), a.palias.size() * 4);
    return MCB_OK;
}

int mcb_upload_domain(mcb_ctx* c, const mcb_domain_desc* d) {
    if (!c) return MCB_EINVAL;
    if (!d || !d->sdoms || !d->planes || d->nsdom <= 0 || d->nplane <= 0) { c->err = "empty domain"; return MCB_EINVAL; }
    if (d->nsdom > MCB_MAX_SDOM) { c->err = "too many subdomains"; return MCB_ELIMIT; }
    if (d->nemitter <= 0 || !d->emitters) { c->err = "Domain has no emitters"; return MCB_EINVAL; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    std::vector<DPlaneHot> hot(d->nplane); std::vector<DPlaneCold> cold(d->nplane);
    for (int i = 0; i < d->nplane; ++i) {
        const mcb_plane_desc& p = d->planes[i];
        if (p.kind < MCB_BDRY_SPEC || p.kind > MCB_BDRY_PERI) { c->err = "unknown boundary kind"; return MCB_EINVAL; }
        if (p.sdom < 0 || p.sdom >= d->nsdom) { c->err = "plane owner out of range"; return MCB_EINVAL; }
        if (p.pair_count < 0 || p.pair_begin < 0 || p.pair_begin + p.pair_count > d->npair) { c->err = "pair range out of bounds"; return MCB_EINVAL; }
        if ((p.kind == MCB_BDRY_PERI && p.pair_count != 1) || (p.kind == MCB_BDRY_INTER && p.pair_count < 1)) {
            c->err = "Boundary not paired (Peri needs 1 partner, Inter >= 1)"; return MCB_EINVAL;   // isInit() boundary.cpp:339,503
        }
        hot[i] = {p.normal[0], p.normal[1], p.normal[2], p.offset};
        DPlaneCold& q = cold[i]; std::memset(&q, 0, sizeof q);
        q.kind = p.kind; q.sdom = p.sdom; q.pair_begin = p.pair_begin; q.pair_count = p.pair_count;
        const double* m = p.kind == MCB_BDRY_PERI ? p.peri_rot : p.rot;
        for (int k = 0; k < 9; ++k) q.m[k] = m[k];
        for (int k = 0; k < 3; ++k) q.t[k] = p.peri_transl[k];
    }
    for (int i = 0; i < d->npair; ++i) if (d->pairs[i] < 0 || d->pairs[i] >= d->nplane) { c->err = "pair id out of range"; return MCB_EINVAL; }
    std::vector<DSdom> sd(d->nsdom); std::vector<double> cell_vol; int any_nd = 0; double flight_max = 0.0;
    int all_box = 1, t1_ok = 1;
    long long cols = 0, nd_mark_words = 0, nd_max_cells = 0;
    for (int s = 0; s < d->nsdom; ++s) {
        const mcb_sdom_desc& S = d->sdoms[s]; DSdom& D = sd[s]; std::memset(&D, 0, sizeof D);
        if (S.plane_count <= 0 || S.plane_begin < 0 || S.plane_begin + S.plane_count > d->nplane) { c->err = "sdom plane range out of bounds"; return MCB_EINVAL; }
        if (!(S.vol > 0.0)) { c->err = "Volume too small, check vector order"; return MCB_EINVAL; }       // subdomain.h:189
        for (int k = 0; k < 3; ++k) { D.o[k] = S.origin[k]; D.div[k] = (double)S.div[k]; D.max[k] = (int32_t)S.max[k]; }
        for (int k = 0; k < 9; ++k) D.inv[k] = S.inv[k];
        D.eps = S.eps; D.accum = S.accum; D.plane_begin = S.plane_begin; D.plane_count = S.plane_count;
        D.stride1 = (int32_t)S.shape[0]; D.stride2 = (int32_t)(S.shape[0] * S.shape[1]);
        D.is_box = (S.cell == MCB_CELL_PARALLELEPIPED && S.plane_count == 6) ? 1 : 0;
        for (int b = 0; b < 3 && D.is_box; ++b)
            for (int k = 0; k < 3; ++k)
                if (d->planes[S.plane_begin + b + 3].normal[k] != -d->planes[S.plane_begin + b].normal[k]) D.is_box = 0;
        D.aabb = D.is_box;
        for (int b = 0; b < 3 && D.aabb; ++b) {
            const mcb_plane_desc& pl = d->planes[S.plane_begin + b];
            for (int k = 0; k < 3; ++k) if (pl.normal[k] != (k == b ? 1.0 : 0.0)) D.aabb = 0;
        }
        for (int b = 0; b < 3; ++b) { D.offl[b] = d->planes[S.plane_begin + b].offset; D.offh[b] = D.is_box ? d->planes[S.plane_begin + b + 3].offset : 0.0; }
        if (!D.aabb) all_box = 0;
        if (S.accum >= 0 && S.accum <= 2) {          // 1-D tally grid along axis `accum` (subdomain.cpp:47-70)
            const int ax = S.accum;
            D.t1_axis = ax; D.t1_o = S.origin[ax]; D.t1_inv = S.inv[4 * ax]; D.t1_div = (double)S.div[ax]; D.t1_max = (int32_t)S.max[ax];
            const long long stride = ax == 0 ? 1 : (ax == 1 ? S.shape[0] : S.shape[0] * S.shape[1]);
            if (stride != 1) t1_ok = 0;
            if (D.aabb) for (int k = 0; k < 3; ++k) if (k != ax && S.inv[ax + 3 * k] != 0.0) all_box = 0;   // inv_ row must be diagonal for the 1-term coord
        }
        {   // a flight stays inside its (convex) subdomain: its length is bounded by the sum of the edge-vector lengths
            double diam = 0.0;
            const int ncol = S.nbase > 0 ? std::min<int>(S.nbase, MCB_MAX_BASE) : 3;
            const double* cols3 = S.nbase > 0 ? S.base : S.mat;
            for (int k = 0; k < ncol; ++k) diam += std::sqrt(cols3[3 * k] * cols3[3 * k] + cols3[3 * k + 1] * cols3[3 * k + 1] + cols3[3 * k + 2] * cols3[3 * k + 2]);
            if (S.cell == MCB_CELL_PARALLELEPIPED && S.nbase <= 0) {      // parallelepiped: the longest of its four body diagonals
                double best = 0.0;
                for (int sg = 0; sg < 4; ++sg) {
                    double v[3];
                    for (int k = 0; k < 3; ++k) v[k] = S.mat[k] + ((sg & 1) ? -1.0 : 1.0) * S.mat[3 + k] + ((sg & 2) ? -1.0 : 1.0) * S.mat[6 + k];
                    best = std::max(best, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
                }
                diam = std::min(diam, best * (1.0 + 1e-12));
            }
            flight_max = std::max(flight_max, diam + 4.0 * std::fabs(S.eps));
        }
        if (S.accum >= 3) {
            nd_mark_words = std::max<long long>(nd_mark_words, 1 + S.max[0] + S.max[1] + S.max[2]);
            nd_max_cells = std::max<long long>(nd_max_cells, std::max(S.max[0], std::max(S.max[1], S.max[2])));
            // BOX kernels take coord() as div * (inv_dd * (p_d - o_d)): every off-diagonal entry of inv_ must be an exact zero
            if (D.aabb) for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) if (r != k && S.inv[r + 3 * k] != 0.0) all_box = 0;
        }
        if (S.accum >= 3) any_nd = std::max(any_nd, (S.shape[0] >= MCB_COOP_ND_CELLS || S.shape[1] >= MCB_COOP_ND_CELLS || S.shape[2] >= MCB_COOP_ND_CELLS) ? 2 : 1);
        const long long sp = S.shape[0] * S.shape[1] * S.shape[2];
        if (S.accum < -2 || S.accum > 4 || sp < 0) { c->err = "bad accum flag / shape"; return MCB_EINVAL; }
        if (sp == 0 && S.accum != -2) { c->err = "subdomain without cells (shape 0) must have accum flag -2"; return MCB_EINVAL; }
        if (sp == 0) D.col_offset = -1;                                                                    // field.cpp:34
        else {
            D.col_offset = (int32_t)cols; cols += sp;
            if (d->cell_vol) {                                       // Subdomain::cellVol evaluated by the caller (CellVolF)
                if (cols > d->ncols) { c->err = "cell_vol has fewer entries than the field has columns"; return MCB_EINVAL; }
                for (long long k = 0; k < sp; ++k) cell_vol.push_back(d->cell_vol[cols - sp + k]);
            } else {
                if (S.cell != MCB_CELL_PARALLELEPIPED && sp != 1) { c->err = "gridded non-box cells need mcb_domain_desc.cell_vol (Subdomain::cellVol per column)"; return MCB_EINVAL; }
                for (long long k = 0; k < sp; ++k) cell_vol.push_back(S.vol / (double)sp);                 // subdomain.cpp:269-273, 396-399, 428-431
            }
        }
        if (cols > 0x7FFFFFFFll) { c->err = "too many cells"; return MCB_ELIMIT; }
    }
    std::vector<DEmitter> em(d->nemitter);
    for (int i = 0; i < d->nemitter; ++i) {
        const mcb_emitter_desc& e = d->emitters[i]; DEmitter& E = em[i]; std::memset(&E, 0, sizeof E);
        E.kind = e.kind; E.index = e.index;
        if (e.kind == MCB_EMIT_SDOM) {
            if (e.index < 0 || e.index >= d->nsdom) { c->err = "emitter sdom out of range"; return MCB_EINVAL; }
            const mcb_sdom_desc& S = d->sdoms[e.index];
            if (S.cell < MCB_CELL_PARALLELEPIPED || S.cell > MCB_CELL_PYRAMID) { c->err = "unknown cell kind"; return MCB_EINVAL; }
            E.sdom = e.index; E.shape = S.cell;
            for (int k = 0; k < 3; ++k) { E.o[k] = S.origin[k]; E.g[k] = S.grad_t[k]; }
            for (int k = 0; k < 9; ++k) { E.a[k] = S.mat[k]; E.rot[k] = S.emit_rot[k]; }
            if (S.cell == MCB_CELL_PRISM || S.cell == MCB_CELL_PYRAMID) {
                // volDist_ over PrismImpl::volume / PyramidImpl::volume (subdomain.cpp:385-394, 417-426): vol(i) uses columns
                // i and i+1 (sic: i = 0 pairs the axis with itself and is zero), while drawPos uses columns ind+1, ind+2
                const int N = S.nbase;
                if (N < 4 || N > MCB_MAX_BASE) { c->err = "prism/pyramid needs 4..9 mat columns"; return MCB_EINVAL; }
                for (int k = 0; k < 3 * N; ++k) E.a[k] = S.base[k];
                double vol[MCB_MAX_BASE];
                const double div = S.cell == MCB_CELL_PRISM ? 2.0 : 6.0;
                for (int i = 0; i < N - 2; ++i) {
                    const double* a = &S.base[3 * i]; const double* b = &S.base[3 * (i + 1)]; const double* z = &S.base[0];
                    const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
                    vol[i] = (cx * z[0] + cy * z[1] + cz * z[2]) / div;
                }
                E.nsub = N - 2;
                build_alias(vol, (size_t)E.nsub, E.sprob, E.salias);
            }
        } else if (e.kind == MCB_EMIT_BDRY) {
            if (e.index < 0 || e.index >= d->nplane) { c->err = "emitter plane out of range"; return MCB_EINVAL; }
            const mcb_plane_desc& p = d->planes[e.index];
            if (p.shape < MCB_SHAPE_PARALLELOGRAM || p.shape > MCB_SHAPE_POLYGON) { c->err = "emitting boundary without a shape"; return MCB_EINVAL; }
            if (p.nvert < 2 || p.nvert > MCB_MAX_VERTS || (p.shape != MCB_SHAPE_POLYGON && p.nvert != 2)) { c->err = "bad shape vertex count"; return MCB_EINVAL; }
            E.sdom = p.sdom; E.shape = p.shape;
            for (int k = 0; k < 3; ++k) E.o[k] = p.origin[k];
            for (int k = 0; k < 3 * p.nvert; ++k) E.a[k] = p.verts[k];
            for (int k = 0; k < 9; ++k) E.rot[k] = p.rot[k];
            E.g[0] = p.T;
            if (p.shape == MCB_SHAPE_POLYGON) {                      // areaDist_ over the fan triangles (boundary.cpp:195-214)
                double area[MCB_MAX_VERTS];
                for (int n = 0; n < p.nvert - 1; ++n) {
                    const double* a = &p.verts[3 * n]; const double* b = &p.verts[3 * (n + 1)];
                    const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
                    area[n] = std::sqrt(cx * cx + cy * cy + cz * cz) / 2.0;
                }
                E.nsub = p.nvert - 1;
                build_alias(area, (size_t)E.nsub, E.sprob, E.salias);
            }
        } else { c->err = "unknown emitter kind"; return MCB_EINVAL; }
    }
    GeometryView v{}; v.nsdom = d->nsdom; v.nplane = d->nplane; v.npair = d->npair;
    uint32_t off = 0;
    v.off_hot = off;   off = align16(off + (uint32_t)(hot.size() * sizeof(DPlaneHot)));
    v.off_cold = off;  off = align16(off + (uint32_t)(cold.size() * sizeof(DPlaneCold)));
    v.off_sdom = off;  off = align16(off + (uint32_t)(sd.size() * sizeof(DSdom)));
    v.off_pairs = off; off = align16(off + (uint32_t)(std::max(d->npair, 1) * sizeof(int32_t)));
    v.bytes = off;
    std::vector<unsigned char> blob(v.bytes, 0);
    std::memcpy(&blob[v.off_hot], hot.data(), hot.size() * sizeof(DPlaneHot));
    std::memcpy(&blob[v.off_cold], cold.data(), cold.size() * sizeof(DPlaneCold));
    std::memcpy(&blob[v.off_sdom], sd.data(), sd.size() * sizeof(DSdom));
    if (d->npair > 0) std::memcpy(&blob[v.off_pairs], d->pairs, d->npair * sizeof(int32_t));
    CUDA_TRY(c, c->geo_blob.alloc(v.bytes));
    CUDA_TRY(c, cudaMemcpy(c->geo_blob.p, blob.data(), v.bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(c, c->emitters.alloc(em.size()));
    CUDA_TRY(c, cudaMemcpy(c->emitters.p, em.data(), em.size() * sizeof(DEmitter), cudaMemcpyHostToDevice));
    CUDA_TRY(c, c->cell_vol.alloc(cell_vol.size()));
    if (!cell_vol.empty()) CUDA_TRY(c, cudaMemcpy(c->cell_vol.p, cell_vol.data(), cell_vol.size() * 8, cudaMemcpyHostToDevice));
    c->gv = v; c->h_sdom = sd; c->nemitter = d->nemitter; c->cols = cols; c->any_nd = any_nd; c->nd_mark_words = (int)std::min<long long>(nd_mark_words, 1 << 20); c->nd_max_cells = nd_max_cells; c->flight_max = flight_max; c->has_dom = true;
    c->all_box = all_box; c->t1_ok = t1_ok;
    return MCB_OK;
}

int mcb_field_cols(const mcb_ctx* c, int64_t* cols) {
    if (!c || !cols || !c->has_dom) return MCB_EINVAL;
    *cols = c->cols; return MCB_OK;
}

int mcb_solve_raw_dev(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end,
                      double* raw_field_dev, mcb_stats* stats) {
    if (!c) return MCB_EINVAL;
    if (!raw_field_dev) { c->err = "null device field"; return MCB_EINVAL; }
    return run_solve(c, prob, seed, n_begin, n_end, raw_field_dev, stats);
}

int mcb_finalize_dev(mcb_ctx* c, const mcb_problem_desc* prob, double* field_dev) {
    if (!c) return MCB_EINVAL;
    int rc = check_problem(c, prob);
    if (rc) return rc;
    if (!field_dev) { c->err = "null device field"; return MCB_EINVAL; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->cols > 0) {
        k_finalize<<<(unsigned)((c->cols + 127) / 128), 128, 0, c->stream>>>(field_dev, prob->rows, c->cols, prob->kind, prob->size,
                                                                            c->energy_sum, prob->power, c->cell_vol.p);
        CUDA_TRY(c, cudaGetLastError());
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MCB_OK;
}

// ---------------------------------------------------------------- several devices in one process (C++ callers, no torch)
int mcb_device_count(int* count) {
    if (!count) return MCB_EINVAL;
    int n = 0, ok = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    for (int d = 0; d < n; ++d) { cudaDeviceProp p; if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok; else break; }
    *count = ok;
    return MCB_OK;
}

int mcb_solve_raw(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end, mcb_stats* stats) {
    if (!c) return MCB_EINVAL;
    int rc = check_problem(c, prob);
    if (rc) return rc;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t len = (size_t)prob->rows * (size_t)c->cols;
    CUDA_TRY(c, c->field.alloc(len));
    CUDA_TRY(c, cudaMemsetAsync(c->field.p, 0, std::max<size_t>(len, 1) * sizeof(double), c->stream));
    return run_solve(c, prob, seed, n_begin, n_end, c->field.p, stats);
}

int mcb_finalize(mcb_ctx* c, const mcb_problem_desc* prob, double* out_field) {
    if (!c) return MCB_EINVAL;
    if (!out_field) { c->err = "null output field"; return MCB_EINVAL; }
    int rc = check_problem(c, prob);
    if (rc) return rc;
    const size_t len = (size_t)prob->rows * (size_t)c->cols;
    if (!c->field.p || c->field.n < len) { c->err = "mcb_finalize before mcb_solve_raw"; return MCB_ESTATE; }
    rc = mcb_finalize_dev(c, prob, c->field.p);
    if (rc) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out_field, c->field.p, len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MCB_OK;
}

namespace {
// NCCL is bound at run time (dlopen), not at link time: a process that already carries another libnccl.so.2 (torch ships
// its own) keeps using that one, and single-GPU users need no NCCL at all.
struct Nccl {
    typedef struct ncclComm* comm_t;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false; std::string err;
    Nccl() {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) { err = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
        CommInitAll = (int (*)(comm_t*, int, const int*))dlsym(h, "ncclCommInitAll");
        CommDestroy = (int (*)(comm_t))dlsym(h, "ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
        GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        ok = CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd && GetErrorString;
        if (!ok) err = "libnccl.so.2 lacks the expected symbols";
    }
};
std::mutex g_nccl_mu;
std::map<std::vector<int>, std::vector<Nccl::comm_t>> g_comms;      // one communicator clique per device list
}

// replaces the omp-critical `sol += partial` of main.cpp:162-165 when the particle range is sharded over GPUs
int mcb_allreduce(mcb_ctx* const* ctxs, int n, const mcb_problem_desc* prob) {
    if (!ctxs || n <= 0 || !ctxs[0]) return MCB_EINVAL;
    mcb_ctx* c0 = ctxs[0];
    int rc = check_problem(c0, prob);
    if (rc) return rc;
    if (n == 1) return MCB_OK;
    const size_t len = (size_t)prob->rows * (size_t)c0->cols;
    std::vector<int> devs;
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || ctxs[i]->cols != c0->cols || !ctxs[i]->field.p || ctxs[i]->field.n < len) { c0->err = "mcb_allreduce: contexts do not hold the same raw field (mcb_solve_raw first)"; return MCB_ESTATE; }
        devs.push_back(ctxs[i]->device);
    }
    for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) if (devs[i] == devs[j]) { c0->err = "mcb_allreduce: one context per device"; return MCB_EINVAL; }
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    static Nccl nccl;
    if (!nccl.ok) { c0->err = nccl.err; return MCB_ECUDA; }
    std::vector<Nccl::comm_t>& comms = g_comms[devs];
    if (comms.empty()) {
        comms.resize(n);
        const int r = nccl.CommInitAll(comms.data(), n, devs.data());
        if (r != 0) { c0->err = std::string("ncclCommInitAll: ") + nccl.GetErrorString(r); g_comms.erase(devs); return MCB_ECUDA; }
    }
    int r = nccl.GroupStart();
    for (int i = 0; i < n && r == 0; ++i) {
        CUDA_TRY(c0, cudaSetDevice(ctxs[i]->device));
        r = nccl.AllReduce(ctxs[i]->field.p, ctxs[i]->field.p, len, /*ncclDouble*/ 8, /*ncclSum*/ 0, comms[i], ctxs[i]->stream);
    }
    const int r2 = nccl.GroupEnd();
    if (r != 0 || r2 != 0) { c0->err = std::string("ncclAllReduce: ") + nccl.GetErrorString(r != 0 ? r : r2); return MCB_ECUDA; }
    for (int i = 0; i < n; ++i) { CUDA_TRY(c0, cudaSetDevice(ctxs[i]->device)); CUDA_TRY(c0, cudaStreamSynchronize(ctxs[i]->stream)); }
    return MCB_OK;
}

int mcb_solve(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end,
              double* out_field, mcb_stats* stats) {
    if (!c) return MCB_EINVAL;
    if (!out_field) { c->err = "null output field"; return MCB_EINVAL; }
    int rc = check_problem(c, prob);
    if (rc) return rc;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t len = (size_t)prob->rows * (size_t)c->cols;
    CUDA_TRY(c, c->field.alloc(len));
    CUDA_TRY(c, cudaMemsetAsync(c->field.p, 0, std::max<size_t>(len, 1) * sizeof(double), c->stream));
    rc = run_solve(c, prob, seed, n_begin, n_end, c->field.p, stats);
    if (rc) return rc;
    rc = mcb_finalize_dev(c, prob, c->field.p);
    if (rc) return rc;
    if (stats) stats->launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(out_field, c->field.p, len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MCB_OK;
}

// the production kernels without the tally: every particle of the range is emitted into its own slot of state[0] (k_emit),
// then one k_step launch runs `nsteps` loop trips (mcb_trace, mcb_sort_probe)
// stop_after: the loop bound is cut to `nsteps` (every phonon ends inactive, mcb_trace); else the survivors stay active
static int trace_run(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end, int64_t nsteps, bool stop_after, StepParams* Pout) {
    const long long n = n_end - n_begin;
    CUDA_TRY(c, cudaSetDevice(c->device));
    RunPlan plan;
    mcb_options saved = c->opt; c->opt.slots = n; c->opt.tally_mode = 2; c->opt.block = 0;
    int rc = plan_run(c, prob, n, &plan);
    c->opt = saved;
    if (rc) return rc;
    if ((rc = ensure_slots(c, n))) return rc;
    if ((rc = upload_cdf(c, prob))) return rc;
    StepParams& P = *Pout; fill_params(c, prob, seed, &P);
    apply_plan(plan, prob, &P);
    if (stop_after) { P.maxloop = std::min<long long>(prob->maxloop, nsteps); P.maxloop32 = (uint32_t)P.maxloop; }
    const long long trips = std::min<long long>(prob->maxloop, nsteps);
    P.field = nullptr; P.do_tally = 0;
    P.steps_per_launch = (int)std::max<long long>(1, std::min<long long>(trips, 0x7FFFFFFF)); P.n_end = (unsigned long long)n_end;
    P.st = view_of(c, 0); P.nslots = n; P.emit_enable = 0;
    Counters init{}; init.next = (unsigned long long)n_begin;
    CUDA_TRY(c, cudaMemcpyAsync(c->ctr.p, &init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->state[0].p, 0, state_bytes(c->slots_alloc), c->stream));
    k_emit<<<(unsigned)(c->sm_count * 8), 256, 0, c->stream>>>(P);     // first fill: particle n_begin + j into slot j
    k_emit_commit<<<1, 32, 0, c->stream>>>(P);
    CUDA_TRY(c, cudaGetLastError());
    if (trips > 0) CUDA_TRY(c, launch_step(P, plan.tm, plan.ndm, c->all_box, plan.pad, plan.grid, plan.block, plan.smem, c->stream));
    return MCB_OK;
}

int mcb_trace(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end, int64_t nsteps,
              mcb_trace_out* o) {
    if (!c) return MCB_EINVAL;
    int rc = check_problem(c, prob);
    if (rc) return rc;
    if (!o || n_begin < 0 || n_end > prob->nemit || n_begin > n_end || nsteps < 0) { c->err = "bad trace arguments"; return MCB_EINVAL; }
    const long long n = n_end - n_begin;
    if (n == 0) return MCB_OK;
    StepParams P;
    if ((rc = trace_run(c, prob, seed, n_begin, n_end, nsteps, true, &P))) return rc;
    DevBuf<double> dpos, ddir, dsn; DevBuf<long long> dw, dp, dnscat, dsteps; DevBuf<int32_t> dsign, dalive, dsdom, dcell;
    CUDA_TRY(c, dpos.alloc(3 * n)); CUDA_TRY(c, ddir.alloc(3 * n)); CUDA_TRY(c, dsn.alloc(n));
    CUDA_TRY(c, dw.alloc(n)); CUDA_TRY(c, dp.alloc(n)); CUDA_TRY(c, dnscat.alloc(n)); CUDA_TRY(c, dsteps.alloc(n));
    CUDA_TRY(c, dsign.alloc(n)); CUDA_TRY(c, dalive.alloc(n)); CUDA_TRY(c, dsdom.alloc(n)); CUDA_TRY(c, dcell.alloc(3 * n));
    k_gather_trace<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(P.st, n, (unsigned long long)n_begin, n, (int)c->np, P.step_bits,
        c->geo_blob.p, c->gv, dpos.p, ddir.p, dsn.p, dw.p, dp.p, dsign.p, dalive.p, dsdom.p, dnscat.p, dsteps.p, dcell.p);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
#define MCB_COPY_OUT(dst, src, count, type) if (dst) CUDA_TRY(c, cudaMemcpy(dst, src.p, (size_t)(count) * sizeof(type), cudaMemcpyDeviceToHost))
    MCB_COPY_OUT(o->pos, dpos, 3 * n, double); MCB_COPY_OUT(o->dir, ddir, 3 * n, double); MCB_COPY_OUT(o->scat_next, dsn, n, double);
    MCB_COPY_OUT(o->w, dw, n, long long); MCB_COPY_OUT(o->p, dp, n, long long); MCB_COPY_OUT(o->sign, dsign, n, int32_t);
    MCB_COPY_OUT(o->alive, dalive, n, int32_t); MCB_COPY_OUT(o->sdom, dsdom, n, int32_t); MCB_COPY_OUT(o->nscat, dnscat, n, long long);
    MCB_COPY_OUT(o->steps, dsteps, n, long long); MCB_COPY_OUT(o->cell, dcell, 3 * n, int32_t);
#undef MCB_COPY_OUT
    dpos.release(); ddir.release(); dsn.release(); dw.release(); dp.release(); dnscat.release(); dsteps.release();
    dsign.release(); dalive.release(); dsdom.release(); dcell.release();
    return MCB_OK;
}

// K3 probe: the particles [n_begin, n_end) after emission and `nsteps` loop trips (no tally) are compacted -- unordered
// (sorted = 0) or by the counting sort of mcb_options::sort_mode (sorted = 1) -- and the resulting slots are described in slot
// order: bin key, field column and particle id (-1: inactive slot).  Parity probe for the integer work of the sort.
int mcb_sort_probe(mcb_ctx* c, const mcb_problem_desc* prob, uint64_t seed, int64_t n_begin, int64_t n_end, int64_t nsteps,
                   int32_t sorted, int64_t* bin, int64_t* col, int64_t* pid, int64_t* cols_per_bin) {
    if (!c) return MCB_EINVAL;
    int rc = check_problem(c, prob);
    if (rc) return rc;
    if (!bin || !col || !pid || n_begin < 0 || n_end > prob->nemit || n_begin > n_end || nsteps < 0) { c->err = "bad sort-probe arguments"; return MCB_EINVAL; }
    const long long n = n_end - n_begin;
    uint32_t cpb, nbins; sort_geometry(c, &cpb, &nbins);
    if (cols_per_bin) *cols_per_bin = cpb;
    if (n == 0) return MCB_OK;
    StepParams P;
    if ((rc = trace_run(c, prob, seed, n_begin, n_end, nsteps, false, &P))) return rc;
    long long launches = 0;
    if ((rc = compact_slots(c, 0, n, n, sorted != 0, &launches))) return rc;
    DevBuf<long long> dbin, dcol, dpid;
    CUDA_TRY(c, dbin.alloc(n)); CUDA_TRY(c, dcol.alloc(n)); CUDA_TRY(c, dpid.alloc(n));
    k_sort_probe<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(view_of(c, 1), n, c->geo_blob.p, c->gv, cpb, nbins, P.step_bits, dbin.p, dcol.p, dpid.p);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(bin, dbin.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(c, cudaMemcpy(col, dcol.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(c, cudaMemcpy(pid, dpid.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return MCB_OK;
}

int mcb_traj(mcb_ctx* c, const mcb_traj_desc* t, uint64_t seed, mcb_traj_out* o) {
    if (!c) return MCB_EINVAL;
    if (!c->has_mat || !c->has_dom) { c->err = "upload material and domain before solving"; return MCB_ESTATE; }
    if (!t || !o || !o->points || !o->step_sdom || !o->step_in || !o->step_in_kind || !o->step_out || !o->step_out_kind ||
        o->max_points < 1 || o->max_steps < 1) { c->err = "bad trajectory buffers"; return MCB_EINVAL; }
    if (t->maxloop < 0 || t->maxloop > MCB_MAX_LOOP || t->maxscat < 0) { c->err = "maxloop / maxscat out of range"; return MCB_EINVAL; }
    if (t->has_prop && (t->w < 0 || t->w >= c->nw || t->p < 0 || t->p >= c->np)) { c->err = "prop out of range"; return MCB_EINVAL; }
    if (t->has_pos && (t->sdom < 0 || t->sdom >= c->gv.nsdom)) { c->err = "Position not inside domain"; return MCB_EINVAL; }   // problem.cpp:242
    CUDA_TRY(c, cudaSetDevice(c->device));
    StepParams P; std::memset(&P, 0, sizeof P);
    P.mat_blob = c->mat_blob.p; P.mv = c->mv; P.geo_blob = c->geo_blob.p; P.gv = c->gv;
    P.emitters = c->emitters.p; P.nemitter = c->nemitter; P.seed = seed; P.maxscat = t->maxscat; P.maxloop = t->maxloop;
    P.maxscat32 = (uint32_t)std::min<long long>(t->maxscat, 0x7FFFFFFFll); P.maxloop32 = (uint32_t)t->maxloop;
    P.step_bits = step_bits_for(t->maxloop); P.step_mask = P.step_bits >= 32 ? 0xFFFFFFFFu : (1u << P.step_bits) - 1u;
    for (int r = 0; r < 10; ++r) { P.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u; P.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u; }
    DevBuf<double> dpts; DevBuf<int32_t> ds[5]; DevBuf<long long> dcnt;
    CUDA_TRY(c, dpts.alloc((size_t)o->max_points * 3)); CUDA_TRY(c, dcnt.alloc(3));
    for (int k = 0; k < 5; ++k) CUDA_TRY(c, ds[k].alloc((size_t)o->max_steps));
    TrajDev dev{dpts.p, o->max_points, o->max_steps, ds[0].p, ds[1].p, ds[2].p, ds[3].p, ds[4].p, dcnt.p};
    k_traj<<<1, 32, 0, c->stream>>>(P, *t, dev);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    long long cnt[3];
    CUDA_TRY(c, cudaMemcpy(cnt, dcnt.p, sizeof cnt, cudaMemcpyDeviceToHost));
    o->npoints = cnt[0]; o->nsteps = cnt[1]; o->escaped = (int32_t)cnt[2];
    const long long np = std::min<long long>(cnt[0], o->max_points), ns = std::min<long long>(cnt[1], o->max_steps);
    CUDA_TRY(c, cudaMemcpy(o->points, dpts.p, (size_t)np * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    int32_t* host[5] = {o->step_sdom, o->step_in, o->step_in_kind, o->step_out, o->step_out_kind};
    for (int k = 0; k < 5; ++k) CUDA_TRY(c, cudaMemcpy(host[k], ds[k].p, (size_t)ns * sizeof(int32_t), cudaMemcpyDeviceToHost));
    dpts.release(); dcnt.release(); for (int k = 0; k < 5; ++k) ds[k].release();
    return MCB_OK;
}

int mcb_cell_index(mcb_ctx* c, int64_t n, const double* pos, const int32_t* sdom, int64_t* index) {
    if (!c) return MCB_EINVAL;
    if (!c->has_dom) { c->err = "upload a domain first"; return MCB_ESTATE; }
    if (n < 0 || (n > 0 && (!pos || !sdom || !index))) { c->err = "bad arguments"; return MCB_EINVAL; }
    if (n == 0) return MCB_OK;
    for (int64_t i = 0; i < n; ++i) if (sdom[i] < 0 || sdom[i] >= c->gv.nsdom) { c->err = "sdom out of range"; return MCB_EINVAL; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    DevBuf<double> dpos; DevBuf<int32_t> ds; DevBuf<long long> di;
    CUDA_TRY(c, dpos.alloc(3 * n)); CUDA_TRY(c, ds.alloc(n)); CUDA_TRY(c, di.alloc(3 * n));
    CUDA_TRY(c, cudaMemcpy(dpos.p, pos, 3 * n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(ds.p, sdom, n * 4, cudaMemcpyHostToDevice));
    k_cell_index<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->geo_blob.p, c->gv, n, dpos.p, ds.p, di.p);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(index, di.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    dpos.release(); ds.release(); di.release();
    return MCB_OK;
}

int mcb_accumulate(mcb_ctx* c, int32_t rows, int64_t n, const int32_t* sdom, const double* bpos, const double* epos,
                   const double* amount, double* field) {
    if (!c) return MCB_EINVAL;
    if (!c->has_dom) { c->err = "upload a domain first"; return MCB_ESTATE; }
    if (rows <= 0 || n < 0 || !field || (n > 0 && (!sdom || !bpos || !epos || !amount))) { c->err = "bad arguments"; return MCB_EINVAL; }
    for (int64_t i = 0; i < n; ++i) if (sdom[i] < 0 || sdom[i] >= c->gv.nsdom) { c->err = "sdom out of range"; return MCB_EINVAL; }
    if (n == 0 || c->cols == 0) return MCB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t len = (size_t)rows * (size_t)c->cols;
    DevBuf<double> db, de, da, df; DevBuf<int32_t> ds;
    CUDA_TRY(c, db.alloc(3 * n)); CUDA_TRY(c, de.alloc(3 * n)); CUDA_TRY(c, da.alloc(rows * n)); CUDA_TRY(c, df.alloc(len)); CUDA_TRY(c, ds.alloc(n));
    CUDA_TRY(c, cudaMemcpy(db.p, bpos, 3 * n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(de.p, epos, 3 * n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(da.p, amount, (size_t)rows * n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(ds.p, sdom, n * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(df.p, field, len * 8, cudaMemcpyHostToDevice));
    k_accumulate<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->geo_blob.p, c->gv, rows, n, ds.p, db.p, de.p, da.p, df.p);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(field, df.p, len * 8, cudaMemcpyDeviceToHost));
    db.release(); de.release(); da.release(); df.release(); ds.release();
    return MCB_OK;
}

int mcb_philox_words(uint64_t seed, uint64_t particle, uint32_t event, uint32_t block, uint32_t out[4]) {
    // runs the DEVICE generator on the current device (no host mirror, no fallback)
    if (!out) return MCB_EINVAL;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 16) != cudaSuccess) { g_create_err = "no CUDA device (there is no CPU fallback)"; return MCB_ENODEVICE; }
    k_philox<<<1, 1>>>(seed, particle, event, block, d);
    cudaError_t e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); return MCB_ECUDA; }
    return MCB_OK;
}

} // extern "C"
