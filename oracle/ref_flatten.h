// oracle/ref_flatten.h — TEST INFRASTRUCTURE, not product code.
//
// The reference-side half of the binding INTEGRATION.md describes, compiled against the reference's REAL classes: it walks
// Domain::sdomPtrs() / emitPtrs(), Subdomain::bdryPtrs() and the private members listed there and fills the C-ABI descriptors
// of include/mcb.h.  Shared by oracle/ref_driver.cpp (`flatten <file>`: descriptors written to a file for the tests) and
// oracle/ref_gpu_solve.cpp (FieldProblem::solve of the reference replaced by the CUDA path, linked into the reference's
// own main(): oracle/_ref/montecarlo_gpu).
// The including file wraps the reference's headers in `#define private public` (a maintainer would add the getters instead).
#pragma once
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace refflat {

inline void die(const char* msg) { std::fprintf(stderr, "ref binding: %s\n", msg); std::exit(2); }

// ---------------------------------------------------------------------------------------------- flatten
inline void put3(double* d, const Vector3d& v) { for (int k = 0; k < 3; ++k) d[k] = v(k); }
inline void put9(double* d, const Matrix3d& m) { for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) d[3 * c + r] = m(r, c); }   // column-major
inline Matrix3d rotTo(const Vector3d& n) { return Eigen::Quaternion<double>::FromTwoVectors(Vector3d::UnitZ(), n).matrix(); }     // boundary.cpp:33-37

// cell kind and (prism / pyramid) base columns need the concrete template type: visit the domain's own containers
struct CellInfo { int cell; int nbase; double base[3 * MCB_MAX_BASE]; };
struct CellVisitor {
    std::vector<CellInfo>* out;
    void push(int cell, const Matrix3Xd* mat) const {
        CellInfo c; std::memset(&c, 0, sizeof c); c.cell = cell;
        if (mat) {
            if (mat->cols() > MCB_MAX_BASE) die("too many base columns");
            c.nbase = (int)mat->cols();
            for (long j = 0; j < mat->cols(); ++j) for (int k = 0; k < 3; ++k) c.base[3 * j + k] = (*mat)(k, j);
        }
        out->push_back(c);
    }
    template <class A, class B, class C, class D, class E, class F> void operator()(const Parallelepiped<A, B, C, D, E, F>&) const { push(MCB_CELL_PARALLELEPIPED, 0); }
    template <class A, class B, class C, class D, class E> void operator()(const TriangularPrism<A, B, C, D, E>&) const { push(MCB_CELL_TRIPRISM, 0); }
    template <class A, class B, class C, class D> void operator()(const Tetrahedron<A, B, C, D>&) const { push(MCB_CELL_TETRAHEDRON, 0); }
    template <class B, class T, class S> void operator()(const Prism<B, T, S>& p) const { push(MCB_CELL_PRISM, &p.mat_); }
    template <class B, class S> void operator()(const Pyramid<B, S>& p) const { push(MCB_CELL_PYRAMID, &p.mat_); }
};

template <class S> bool periOf(const Boundary* b, mcb_plane_desc& d, const Boundary** pair) {
    const PeriBoundary<S>* p = dynamic_cast<const PeriBoundary<S>*>(b);
    if (!p) return false;
    put9(d.peri_rot, p->rot_); put3(d.peri_transl, p->transl_); *pair = p->pair_;
    return true;
}
inline void shapeOf(const Boundary::Shape& sh, mcb_plane_desc& d) {
    if (const Parallelogram* q = dynamic_cast<const Parallelogram*>(&sh)) { d.shape = MCB_SHAPE_PARALLELOGRAM; d.nvert = 2; put3(d.verts, q->i_); put3(d.verts + 3, q->j_); return; }
    if (const Triangle* q = dynamic_cast<const Triangle*>(&sh)) { d.shape = MCB_SHAPE_TRIANGLE; d.nvert = 2; put3(d.verts, q->i_); put3(d.verts + 3, q->j_); return; }
    const Matrix3Xd* v = 0;
    if (const Polygon<4>* q = dynamic_cast<const Polygon<4>*>(&sh)) v = &q->verts_;
    else if (const Polygon<5>* q = dynamic_cast<const Polygon<5>*>(&sh)) v = &q->verts_;
    else if (const Polygon<6>* q = dynamic_cast<const Polygon<6>*>(&sh)) v = &q->verts_;
    else if (const Polygon<7>* q = dynamic_cast<const Polygon<7>*>(&sh)) v = &q->verts_;
    else if (const Polygon<8>* q = dynamic_cast<const Polygon<8>*>(&sh)) v = &q->verts_;
    else if (const Polygon<9>* q = dynamic_cast<const Polygon<9>*>(&sh)) v = &q->verts_;
    if (!v) die("unknown boundary shape");
    d.shape = MCB_SHAPE_POLYGON; d.nvert = (int32_t)v->cols();
    for (long j = 0; j < v->cols(); ++j) for (int k = 0; k < 3; ++k) d.verts[3 * j + k] = (*v)(k, j);
}

struct FlatDomain {
    std::vector<mcb_sdom_desc> sdoms; std::vector<mcb_plane_desc> planes; std::vector<int32_t> pairs;
    std::vector<mcb_emitter_desc> emitters; std::vector<double> cellVol;
    mcb_domain_desc desc() const {
        mcb_domain_desc d; std::memset(&d, 0, sizeof d);
        d.nsdom = (int32_t)sdoms.size(); d.sdoms = sdoms.data(); d.nplane = (int32_t)planes.size(); d.planes = planes.data();
        d.npair = (int32_t)pairs.size(); d.pairs = pairs.data(); d.nemitter = (int32_t)emitters.size(); d.emitters = emitters.data();
        d.ncols = (int64_t)cellVol.size(); d.cell_vol = cellVol.data();
        return d;
    }
};
// Domain::sdomPtrs() / emitPtrs() / Subdomain::bdryPtrs() and the members listed in INTEGRATION.md -> the descriptors of include/mcb.h
inline void flattenDomain(const Domain* dom, const std::vector<CellInfo>& cells, FlatDomain& out) {
    std::vector<mcb_sdom_desc>& sdoms = out.sdoms; std::vector<mcb_plane_desc>& planes = out.planes; std::vector<int32_t>& pairs = out.pairs;
    std::vector<mcb_emitter_desc>& emitters = out.emitters; std::vector<double>& cellVol = out.cellVol;
    const Subdomain::Pointers& sp = dom->sdomPtrs();
    if (cells.size() != sp.size()) die("cell visitor / sdomPtrs mismatch");
    std::map<const Boundary*, int32_t> planeId; std::map<const Subdomain*, int32_t> sdomId;
    std::vector<std::vector<const Boundary*> > partners;
    for (size_t s = 0; s < sp.size(); ++s) {
        const Subdomain* sd = sp[s]; sdomId[sd] = (int32_t)s;
        mcb_sdom_desc d; std::memset(&d, 0, sizeof d);
        put3(d.origin, sd->o_); put9(d.mat, sd->mat_); put9(d.inv, sd->inv_);
        for (int k = 0; k < 3; ++k) { d.div[k] = sd->div_(k); d.shape[k] = sd->shape_(k); d.max[k] = sd->max_(k); }
        d.accum = sd->accum_; d.cell = cells[s].cell; d.eps = sd->eps_; d.vol = sd->vol_;
        d.nbase = cells[s].nbase; std::memcpy(d.base, cells[s].base, sizeof d.base);
        const EmitSubdomain* es = dynamic_cast<const EmitSubdomain*>(sd);
        put9(d.emit_rot, Matrix3d::Identity());
        if (es) {
            put3(d.grad_t, es->gradT_);
            if (es->gradT_.norm() > 0.) put9(d.emit_rot, es->rot_);            // rotMatrix(0/0) is NaN in the reference and never used
        }
        d.plane_begin = (int32_t)planes.size(); d.plane_count = (int32_t)sd->bdryPtrs().size();
        for (size_t b = 0; b < sd->bdryPtrs().size(); ++b) {
            const Boundary* bd = sd->bdryPtrs()[b];
            planeId[bd] = (int32_t)planes.size();
            mcb_plane_desc p; std::memset(&p, 0, sizeof p);
            put3(p.normal, bd->normal()); p.offset = bd->offset(); p.sdom = (int32_t)s; p.shape = MCB_SHAPE_NONE;
            put9(p.rot, rotTo(bd->normal()));
            std::vector<const Boundary*> prt;
            const std::string ty = bd->type();
            if (ty == "Spec") p.kind = MCB_BDRY_SPEC;
            else if (ty == "Diff") { p.kind = MCB_BDRY_DIFF; put9(p.rot, dynamic_cast<const DiffBoundary*>(bd)->rot_); }
            else if (ty == "Inter") { p.kind = MCB_BDRY_INTER; const InterBoundary* ib = dynamic_cast<const InterBoundary*>(bd); prt.assign(ib->pairs_.begin(), ib->pairs_.end()); }
            else {
                const EmitBoundary* eb = dynamic_cast<const EmitBoundary*>(bd);
                if (!eb) die("unknown boundary type");
                p.kind = ty.compare(0, 4, "Isot") == 0 ? MCB_BDRY_ISOT : MCB_BDRY_PERI;
                put9(p.rot, eb->rot_); p.T = eb->T_; put3(p.origin, eb->o_);
                shapeOf(eb->shape(), p);
                if (p.kind == MCB_BDRY_PERI) {
                    const Boundary* pr = 0;
                    if (!(periOf<Parallelogram>(bd, p, &pr) || periOf<Triangle>(bd, p, &pr) || periOf<Polygon<4> >(bd, p, &pr) ||
                          periOf<Polygon<5> >(bd, p, &pr) || periOf<Polygon<6> >(bd, p, &pr) || periOf<Polygon<7> >(bd, p, &pr) ||
                          periOf<Polygon<8> >(bd, p, &pr) || periOf<Polygon<9> >(bd, p, &pr))) die("unknown periodic boundary");
                    if (pr) prt.push_back(pr);
                }
            }
            partners.push_back(prt);
            planes.push_back(p);
        }
        sdoms.push_back(d);
        const Vector3l shp = sd->shape();                                   // Field(rows, dom, fun) nesting: k, j, i (field.cpp:62-78)
        for (long k = 0; k < shp(2); ++k) for (long j = 0; j < shp(1); ++j) for (long i = 0; i < shp(0); ++i)
            cellVol.push_back(sd->cellVol(Vector3l(i, j, k)));
    }
    for (size_t q = 0; q < planes.size(); ++q) {
        planes[q].pair_begin = (int32_t)pairs.size(); planes[q].pair_count = (int32_t)partners[q].size();
        for (size_t k = 0; k < partners[q].size(); ++k) {
            if (!planeId.count(partners[q][k])) die("boundary paired with a boundary outside the domain");
            pairs.push_back(planeId[partners[q][k]]);
        }
    }
    for (size_t e = 0; e < dom->emitPtrs().size(); ++e) {
        const Emitter* em = dom->emitPtrs()[e];
        mcb_emitter_desc d; std::memset(&d, 0, sizeof d);
        if (em->emitBdry()) { d.kind = MCB_EMIT_BDRY; d.index = planeId.at(em->emitBdry()); }
        else { d.kind = MCB_EMIT_SDOM; d.index = sdomId.at(em->emitSdom()); }
        d.weight = em->emitWeight();
        emitters.push_back(d);
    }
}

// cell kinds of a shipped Domain in sdomPtrs() order (the concrete container type is needed for Prism / Pyramid columns)
inline std::vector<CellInfo> cellsOf(const Domain* dom) {
    std::vector<CellInfo> cells; CellVisitor vis = {&cells};
    if (const BulkDomain* d = dynamic_cast<const BulkDomain*>(dom)) vis(d->sdom_);
    else if (const FilmDomain* d = dynamic_cast<const FilmDomain*>(dom)) vis(d->sdom_);
    else if (const HexDomain* d = dynamic_cast<const HexDomain*>(dom)) vis(d->sdom_);
    else if (const PyrDomain* d = dynamic_cast<const PyrDomain*>(dom)) vis(d->sdom_);
    else if (const JctDomain* d = dynamic_cast<const JctDomain*>(dom)) boost::fusion::for_each(d->sdomCont_, vis);
    else if (const TeeDomain* d = dynamic_cast<const TeeDomain*>(dom)) boost::fusion::for_each(d->sdomCont_, vis);
    else if (const TubeDomain* d = dynamic_cast<const TubeDomain*>(dom)) boost::fusion::for_each(d->sdomCont_, vis);
    else if (const OctetDomain* d = dynamic_cast<const OctetDomain*>(dom)) boost::fusion::for_each(d->sdomCont_, vis);
    else die("unknown Domain class");
    return cells;
}

} // namespace refflat
