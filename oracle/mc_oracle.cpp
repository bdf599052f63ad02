/*
 * mc_oracle.cpp — CPU ORACLE for the phonon Monte Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  A dependency-free C++17 restatement of the algorithm of
 * nickdou/montecarlocpp's FieldProblem::solve and everything it calls, in fp64 and in the
 * reference's operation order at expression level.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/montecarlo/).
 *
 * Parity is pinned against the reference itself (oracle/_ref) — see mc_oracle.h.  Third-party arithmetic that is NOT under
 * /root/reference and is restated here from its published algorithm:
 *   - Eigen 3.2.x (unpinned; no build files shipped): Hyperplane(n,e) offset = -n.e,
 *     signedDistance = n.p + offset, ParametrizedLine::intersection = -(offset+n.o)/(n.d),
 *     pointAt = o + d*t, Quaternion::FromTwoVectors(...).matrix(), Matrix3d::inverse (cofactor),
 *     normalized() = v / sqrt(x^2+y^2+z^2).  Eigen's internal summation tree for 3-vectors is
 *     not reproduced (left-to-right here; <= 1 ulp, benign).  For the antiparallel case of
 *     FromTwoVectors Eigen picks an SVD-dependent axis; any rotation with R z = n is
 *     statistically equivalent (drawAniso is azimuthally uniform) — we use a turn about x.
 *   - Boost.Random 1.5x (unpinned): mt19937 (bit-identical to std::mt19937), uniform_01<double>
 *     and uniform_real_distribution<double> consume ONE 32-bit word (u = x * 2^-32),
 *     uniform_int_distribution<long> = bucketed rejection on one word (no word when the range
 *     is empty), discrete_distribution<long,double> = Walker alias table (1 int + 1 real draw).
 *
 * Deliberate differences from the reference (all documented in DESIGN.md):
 *   - `Vector3d coord(dist(gen), dist(gen), dist(gen))` (subdomain.cpp:279) has unspecified
 *     argument evaluation order in C++; we draw x, y, z left to right.
 *   - In ORC_RNG_PHILOX mode the word source is Philox4x32-10 keyed by (seed, particle id)
 *     with counter (particle id, event, block): event 0 = emission, event i+1 = loop trip i.
 *     The distributions on top of the words are unchanged.
 */
#include "mc_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;
void set_err(const std::string& s) { g_err = s; }

/* constants.h:15-19 — the reference's literal values (HBAR is not CODATA; keep it) */
const double PI   = 3.141592653589793;
const double HBAR = 1.054560652927e-034;
const double KB   = 1.380648e-023;
const double DMIN = std::numeric_limits<double>::min();
const double DEPS = std::numeric_limits<double>::epsilon();

/* ------------------------------------------------------------------ small linear algebra */
struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double  operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(const V3& a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(const V3& a, double s) { return V3(a.x / s, a.y / s, a.z / s); }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(const V3& a) { return a / norm(a); }

struct M3 {               /* column-major like Eigen: m[r + 3*c] */
    double m[9];
    M3() { for (double& v : m) v = 0.; }
    double  operator()(int r, int c) const { return m[r + 3 * c]; }
    double& operator()(int r, int c) { return m[r + 3 * c]; }
    V3 col(int c) const { return V3(m[3 * c], m[3 * c + 1], m[3 * c + 2]); }
    static M3 identity() { M3 r; r(0, 0) = r(1, 1) = r(2, 2) = 1.; return r; }
    static M3 diag(double a, double b, double c) { M3 r; r(0, 0) = a; r(1, 1) = b; r(2, 2) = c; return r; }
};
inline V3 operator*(const M3& A, const V3& v) {   /* col0*v0 + col1*v1 + col2*v2 */
    return V3(A(0, 0) * v.x + A(0, 1) * v.y + A(0, 2) * v.z,
              A(1, 0) * v.x + A(1, 1) * v.y + A(1, 2) * v.z,
              A(2, 0) * v.x + A(2, 1) * v.y + A(2, 2) * v.z);
}
inline M3 transpose(const M3& A) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(j, i); return r; }
inline double determinant(const M3& A) {
    return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1))
         - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0))
         + A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
}
/* Eigen Matrix3d::inverse(): cofactor matrix / determinant (Eigen/src/LU/Inverse.h, size 3) */
inline M3 inverse(const M3& A) {
    M3 c;
    c(0, 0) =  (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1));
    c(1, 0) = -(A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0));
    c(2, 0) =  (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    c(0, 1) = -(A(0, 1) * A(2, 2) - A(0, 2) * A(2, 1));
    c(1, 1) =  (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0));
    c(2, 1) = -(A(0, 0) * A(2, 1) - A(0, 1) * A(2, 0));
    c(0, 2) =  (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1));
    c(1, 2) = -(A(0, 0) * A(1, 2) - A(0, 2) * A(1, 0));
    c(2, 2) =  (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0));
    double det = A(0, 0) * c(0, 0) + A(0, 1) * c(1, 0) + A(0, 2) * c(2, 0);
    double inv = 1. / det;
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = c.m[i] * inv;
    return r;
}

/* boundary.cpp:33-37 / subdomain.cpp:26-30 rotMatrix(n) = Quatd::FromTwoVectors(UnitZ, n).matrix()
 * (Eigen/src/Geometry/Quaternion.h setFromTwoVectors + toRotationMatrix) */
M3 rotMatrix(const V3& n) {
    V3 v0(0., 0., 1.);
    double nn = norm(n);
    if (!(nn > 0.)) return M3::identity();        /* gradT == 0: Eigen yields NaN; never used */
    V3 v1 = n / nn;
    double c = dot(v1, v0);
    double qx, qy, qz, qw;
    if (c < -1. + 1e-12) {                        /* antiparallel: Eigen's SVD branch */
        qx = 1.; qy = 0.; qz = 0.; qw = 0.;       /* 180 degrees about x (see header) */
    } else {
        V3 axis = cross(v0, v1);
        double s = std::sqrt((1. + c) * 2.);
        double invs = 1. / s;
        qx = axis.x * invs; qy = axis.y * invs; qz = axis.z * invs; qw = s * 0.5;
    }
    double tx = 2. * qx, ty = 2. * qy, tz = 2. * qz;
    double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    M3 r;
    r(0, 0) = 1. - (tyy + tzz); r(0, 1) = txy - twz;        r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;        r(1, 1) = 1. - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;        r(2, 1) = tyz + twx;        r(2, 2) = 1. - (txx + tyy);
    return r;
}

/* boundary.cpp:28-31 reflMatrix(n) = I - 2 n n^T */
M3 reflMatrix(const V3& n) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r(i, j) = (i == j ? 1. : 0.) - 2. * (n[i] * n[j]);
    return r;
}

/* ------------------------------------------------------------------------------ RNG */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 philox.h): published algorithm. */
inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* A source of 32-bit words: mt19937 (random.h:22) or the per-particle Philox stream. */
class Words {
public:
    int mode;
    std::mt19937 mt;
    uint32_t key[2]; uint32_t pid[2]; uint32_t event; uint32_t idx; uint32_t buf[4];
    explicit Words(int m, uint64_t seed) : mode(m), mt((uint32_t)seed), event(0), idx(0) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); pid[0] = pid[1] = 0;
    }
    /* Philox only: start the word sequence of (particle, event) */
    void begin(uint64_t particle, uint32_t ev) {
        pid[0] = (uint32_t)particle; pid[1] = (uint32_t)(particle >> 32); event = ev; idx = 0;
    }
    uint32_t next() {
        if (mode == ORC_RNG_MT19937) return (uint32_t)mt();
        if ((idx & 3u) == 0) {
            uint32_t ctr[4] = {pid[0], pid[1], event, idx >> 2};
            philox4x32_10(ctr, key, buf);
        }
        return buf[(idx++) & 3u];
    }
};

static int g_arg_rtl = 0;   /* orc_set_arg_order(): see Domain::emit */

/* boost::random::uniform_01<double> on a 32-bit engine: x / 2^32, one word (random.h:23) */
inline double uniform01(Words& g) {
    for (;;) {
        double r = (double)g.next() * (1. / 4294967296.);
        if (r < 1.) return r;
    }
}
/* boost::random::uniform_real_distribution<double>(-1, 1): x/2^32*(max-min)+min (random.h:24) */
inline double uniformOne(Words& g) {
    for (;;) {
        double r = (double)g.next() / 4294967296. * (1. - (-1.)) + (-1.);
        if (r < 1.) return r;
    }
}
/* boost::random::uniform_int_distribution<long>(0, n-1) on a 32-bit engine (random.h:25):
 * empty range consumes nothing; otherwise bucketed rejection on one word. */
inline long uniformInt(Words& g, long n) {
    uint32_t range = (uint32_t)(n - 1);
    if (range == 0) return 0;
    const uint32_t brange = 0xFFFFFFFFu;
    uint32_t bucket = brange / (range + 1u);
    if (brange % (range + 1u) == range) ++bucket;
    for (;;) {
        uint32_t r = g.next() / bucket;
        if (r <= range) return (long)r;
    }
}

/* boost::random::discrete_distribution<long,double> (random.h:26): Walker alias table */
struct Discrete {
    std::vector<double> prob;
    std::vector<long>   alias;
    Discrete() {}
    Discrete(const double* first, const double* last) {
        size_t size = (size_t)(last - first);
        std::vector<std::pair<double, long>> below, above;
        double sum = 0.;
        for (const double* it = first; it != last; ++it) sum += *it;
        double avg = sum / (double)size;
        long i = 0;
        for (const double* it = first; it != last; ++it, ++i) {
            double val = *it / avg;
            if (val < 1.) below.emplace_back(val, i); else above.emplace_back(val, i);
        }
        prob.assign(size, 0.); alias.assign(size, 0);
        auto b = below.begin(), be = below.end();
        auto a = above.begin(), ae = above.end();
        while (b != be && a != ae) {
            prob[b->second] = b->first; alias[b->second] = a->second;
            a->first -= (1. - b->first);
            if (a->first < 1.) { *b = *a++; } else { ++b; }
        }
        for (; b != be; ++b) prob[b->second] = 1.;
        for (; a != ae; ++a) prob[a->second] = 1.;
    }
    long operator()(Words& g) const {
        long r = uniformInt(g, (long)prob.size());
        double test = uniform01(g);
        return test < prob[r] ? r : alias[r];
    }
};

/* random.cpp:16-27 */
V3 drawIso(Words& g) {
    double cosTheta = uniformOne(g);
    double sinTheta = std::sqrt(1. - cosTheta * cosTheta);
    double phi = PI * uniformOne(g);
    return V3(sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta);
}
/* random.cpp:29-44 */
V3 drawAniso(Words& g, bool bidir) {
    double r = uniformOne(g);
    int sign = (bidir ? (r < 0. ? -1 : 1) : 1);
    double sinSqTheta = std::abs(r);
    double sinTheta = std::sqrt(sinSqTheta);
    double cosTheta = sign * std::sqrt(1. - sinSqTheta);
    double phi = PI * uniformOne(g);
    return V3(sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta);
}

/* ---------------------------------------------------------------------------- Phonon */
/* phonon.h:36-42, phonon.cpp:33-127.  line_ duplicates (pos_, dir_) and is dropped. */
struct Phonon {
    bool alive = true, sign = true;
    long w = 0, p = 0;
    V3 pos, dir;
    double time = 0., scatNext = 0.;
    long nscat = 0;
    Phonon() {}
    Phonon(bool s, long w_, long p_, const V3& pos_, const V3& dir_)
        : alive(true), sign(s), w(w_), p(p_), pos(pos_), dir(normalized(dir_)), time(0.), scatNext(0.), nscat(0) {}
    int sgn() const { return sign ? 1 : -1; }
    void setDir(const V3& d, bool scatter) { dir = normalized(d); if (scatter) nscat++; }   /* :88-93 */
    void move(double distance, double vel) {                                                 /* :95-105 */
        scatNext -= distance;
        if (scatNext < DMIN) scatNext = 0.;
        time += distance / vel;
        pos = pos + dir * distance;
    }
};

/* -------------------------------------------------------------------------- Material */
struct Dist {                                   /* material.cpp:51-75 */
    Discrete wDist; std::vector<Discrete> pDist;
    Dist() {}
    Dist(const std::vector<double>& pdf, long nw, long np) {   /* pdf(w,p) at [w + nw*p] */
        std::vector<double> rowSum(nw, 0.);
        for (long w = 0; w < nw; ++w) { double s = 0.; for (long p = 0; p < np; ++p) s += pdf[w + nw * p]; rowSum[w] = s; }
        wDist = Discrete(rowSum.data(), rowSum.data() + nw);
        pDist.reserve(nw);
        for (long w = 0; w < nw; ++w) {
            std::vector<double> row(np);
            for (long p = 0; p < np; ++p) row[p] = pdf[w + nw * p];
            pDist.emplace_back(row.data(), row.data() + np);
        }
    }
    void draw(Words& g, long& w, long& p) const { w = wDist(g); p = pDist.at(w)(g); }
};

double sum_colmajor(const std::vector<double>& a) { double s = 0.; for (double v : a) s += v; return s; }

} // namespace

struct orc_material {
    long nw = 0, np = 0; double T = 0., k = 0.;
    std::vector<double> omega, tau, vel, energyPdf, fluxPdf, scatPdf;
    Dist energyDist, fluxDist, scatDist;
    double energySum = 0., fluxSum = 0., scatSum = 0.;
    double velAt(const Phonon& ph) const { return vel[ph.w + nw * ph.p]; }   /* material.cpp:180-184 */
    double tauAt(const Phonon& ph) const { return tau[ph.w + nw * ph.p]; }   /* material.cpp:174-178 */
    /* material.cpp:215-224 */
    void drawScatNext(Phonon& ph, Words& g) const {
        double distance = 0.;
        while (distance < DMIN) distance = velAt(ph) * tauAt(ph) * -std::log(1. - uniform01(g));
        ph.scatNext = distance;
    }
    /* material.cpp:226-231 */
    void scatter(Phonon& ph, Words& g) const {
        scatDist.draw(g, ph.w, ph.p);
        ph.setDir(drawIso(g), true);
        drawScatNext(ph, g);
    }
};

namespace {

/* material.cpp:23-45 extractArray: one text line per row, `cols` numbers per line */
bool extractArray(std::istream& is, std::vector<double>& data, long rows, long cols) {
    is >> std::ws;
    std::string line; long i = 0;
    data.assign((size_t)(rows * cols), 0.);
    while (i < rows && std::getline(is, line)) {
        std::stringstream ss(line);
        for (long j = 0; j < cols; ++j) { double e; ss >> e; if (!ss) return false; data[i * cols + j] = e; }
        i++;
    }
    return i == rows;
}

} // namespace

/* ------------------------------------------------------------------ geometry objects */
struct OBoundary {                              /* boundary.h:34-63 + subclasses */
    int kind = MCB_BDRY_SPEC;
    V3 normal; double offset = 0.;
    int sdom = -1;
    std::vector<int> pairs;                     /* Inter pairs_ / Peri pair_ (plane ids) */
    M3 rot, refl, periRot; V3 periTransl;
    double T = 0.; V3 o;
    int shape = MCB_SHAPE_PARALLELOGRAM; std::vector<V3> verts;
    bool emitRegistered = false;
    double distancePos(const V3& p) const { return dot(normal, p) + offset; }            /* boundary.cpp:102-105 */
    double distancePhn(const Phonon& ph) const { return -(offset + dot(normal, ph.pos)) / dot(normal, ph.dir); } /* :107-110 */
    double area() const {                                                                /* boundary.cpp:142-145,177-180,237-241 */
        if (shape == MCB_SHAPE_PARALLELOGRAM) return norm(cross(verts[0], verts[1]));
        if (shape == MCB_SHAPE_TRIANGLE) return norm(cross(verts[0], verts[1])) / 2.;
        double s = 0.; for (size_t n = 0; n + 1 < verts.size(); ++n) s += norm(cross(verts[n], verts[n + 1])) / 2.;
        return s;
    }
    double emitWeight() const { return area() * std::abs(T); }                          /* boundary.cpp:413-416 */
};

struct OSubdomain {                             /* subdomain.h:35-126 */
    int cell = MCB_CELL_PARALLELEPIPED;
    double vol = 0.; V3 o; M3 mat, inv;
    long div[3] = {0, 0, 0}, shape[3] = {1, 1, 1}, max[3] = {0, 0, 0};
    int accum = -1; double eps = 0.;
    V3 gradT; M3 emitRot;
    std::vector<V3> base;                       /* Prism / Pyramid: mat_ columns (subdomain.h:378,501) */
    Discrete volDist;                           /* Prism / Pyramid: volDist_ */
    std::vector<int> planes;                    /* bdryPtrs_ (plane ids, declaration order) */
    std::vector<int> emitPlanes;                /* emitPtrs_ of this sdom */
    double emitWeight() const { return 2. * vol * norm(gradT); }                        /* subdomain.cpp:250-253 */
    long shapeProd() const { return shape[0] * shape[1] * shape[2]; }
};

struct OEmitter { int kind; int index; };

struct orc_domain {
    std::vector<OSubdomain> sdoms;
    std::vector<OBoundary>  planes;
    std::vector<OEmitter>   emitters;           /* Domain::emitPtrs() order */
    std::vector<long>       colOffset;          /* Field::init stride(0) per sdom, -1 if none */
    long cols = 0;
    /* flattened copies handed out through orc_domain_desc */
    std::vector<mcb_sdom_desc> fs; std::vector<mcb_plane_desc> fp; std::vector<int32_t> fpairs;
    std::vector<mcb_emitter_desc> fe; std::vector<double> fcellvol;

    /* subdomain.cpp:108-116 */
    bool isInside(int s, const V3& pos) const {
        const OSubdomain& sd = sdoms[s];
        for (int b : sd.planes) if (planes[b].distancePos(pos) < -sd.eps) return false;
        return true;
    }
    /* subdomain.cpp:148-151 */
    V3 coord(int s, const V3& pos) const {
        const OSubdomain& sd = sdoms[s];
        V3 t = sd.inv * (pos - sd.o);
        return V3((double)sd.div[0] * t.x, (double)sd.div[1] * t.y, (double)sd.div[2] * t.z);
    }
    /* subdomain.cpp:153-159 */
    void coord2index(int s, const V3& c, long idx[3]) const {
        const OSubdomain& sd = sdoms[s];
        for (int d = 0; d < 3; ++d) {
            long v = static_cast<long>(std::floor(c[d]));
            idx[d] = std::min(std::max(v, 0l), sd.max[d]);
        }
    }
    /* subdomain.cpp:161-192; returns hit plane id or -1 */
    int advect(int s, Phonon& ph, double vel) const {
        const OSubdomain& sd = sdoms[s];
        double minDistance = ph.scatNext;
        int newBdry = -1;
        for (int b : sd.planes) {
            const OBoundary& B = planes[b];
            if (dot(B.normal, ph.dir) >= 0.) continue;
            double distance = B.distancePhn(ph);
            if (distance < minDistance) { minDistance = distance; newBdry = b; }
        }
        ph.move(minDistance, vel);
        bool negDist = (minDistance < -sd.eps);
        bool outside = !isInside(s, ph.pos);
        if (negDist || outside) { ph.alive = false; return -1; }
        return newBdry;
    }
    /* Boundary::scatter family (boundary.cpp:283-287,308-312,349-359,455-460,516-522);
     * returns the plane the particle now sits on, or -1 (failed Inter hand-off). */
    int scatter(int b, Phonon& ph, Words& g) const {
        const OBoundary& B = planes[b];
        switch (B.kind) {
        case MCB_BDRY_SPEC: {
            /* refl_.selfadjointView<Upper>() * dir: symmetric product from the upper triangle */
            const M3& R = B.refl; const V3& d = ph.dir;
            V3 r(R(0, 0) * d.x + R(0, 1) * d.y + R(0, 2) * d.z,
                 R(0, 1) * d.x + R(1, 1) * d.y + R(1, 2) * d.z,
                 R(0, 2) * d.x + R(1, 2) * d.y + R(2, 2) * d.z);
            ph.setDir(r, false);
            return b;
        }
        case MCB_BDRY_DIFF:
            ph.setDir(B.rot * drawAniso(g, false), true);
            return b;
        case MCB_BDRY_INTER:
            if (B.pairs.size() == 1) return B.pairs.front();
            for (int q : B.pairs) if (isInside(planes[q].sdom, ph.pos)) return q;
            ph.alive = false;
            return -1;
        case MCB_BDRY_ISOT:
            ph.alive = false;
            return b;
        case MCB_BDRY_PERI:
            ph.pos = B.periRot * ph.pos + B.periTransl;
            ph.setDir(B.periRot * ph.dir, false);
            return B.pairs.front();
        }
        return -1;
    }
    /* ParallelepipedImpl::cellVol subdomain.cpp:269-273 (other cells: N3, not built yet) */
    /* Subdomain::cellVol(index): ParallelepipedImpl :269-273, TriangularPrismImpl :283-307, TetrahedronImpl :322-349,
     * PrismImpl :396-399, PyramidImpl :428-431.  The partial-cell formulas (and their normalisation, which counts a full
     * cell as vol/prod(shape) although the simplex fills only 1/2 resp. 1/6 of the spanning box) are kept literally. */
    double cellVol(int s, const long idx[3]) const {
        const OSubdomain& sd = sdoms[s];
        const double prod = (double)sd.shapeProd();
        if (sd.cell == MCB_CELL_PARALLELEPIPED) return sd.vol / prod;
        if (sd.cell == MCB_CELL_PRISM || sd.cell == MCB_CELL_PYRAMID) return sd.vol;
        const int nd = sd.cell == MCB_CELL_TRIPRISM ? 2 : 3;
        double shp[3] = {(double)sd.shape[0], (double)sd.shape[1], (double)sd.shape[2]};
        double q = 0.; for (int d = 0; d < nd; ++d) q += (double)idx[d] / shp[d];
        const double f0 = 1. - q;
        if (f0 <= 0.) return 0.;
        double one = 0.; for (int d = 0; d < nd; ++d) one += 1. / shp[d];
        const double f1 = f0 - one;
        if (f1 >= 0.) return sd.vol / prod;
        if (nd == 2) {
            double frac = std::pow(f0, 2);
            for (int i = 0; i < 2; i++) {                                   /* corners (1,0), (0,1): one step along one axis */
                double f = f0 - 1. / shp[i];
                if (f > 0.) frac += -1 * std::pow(f, 2);
            }
            return sd.vol * frac / (2. * shp[2]);
        }
        static const int pts[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
        double frac = std::pow(f0, 3);
        for (int i = 0; i < 6; i++) {
            int sum = pts[i][0] + pts[i][1] + pts[i][2];
            int sign = (sum % 2) ? -1 : 1;
            double f = f0 - ((double)pts[i][0] / shp[0] + (double)pts[i][1] / shp[1] + (double)pts[i][2] / shp[2]);
            if (f > 0.) frac += sign * std::pow(f, 3);
        }
        return sd.vol * frac / 6.;
    }

    /* Emitter::emit boundary.cpp:378-385 */
    Phonon emit(const OEmitter& e, long w, long p, Words& g) const {
        V3 pos, dir; bool sign;
        if (e.kind == MCB_EMIT_SDOM) {
            const OSubdomain& sd = sdoms[e.index];
            M3 m = sd.mat;
            if (sd.cell == MCB_CELL_PRISM || sd.cell == MCB_CELL_PYRAMID) {     /* PrismImpl/PyramidImpl::drawPos :401-409, :433-441 */
                long ind = sd.volDist(g);
                const V3 &a = sd.base.at((size_t)ind + 1), &b = sd.base.at((size_t)ind + 2), &z = sd.base[0];
                for (int k = 0; k < 3; ++k) { m(k, 0) = a[k]; m(k, 1) = b[k]; m(k, 2) = z[k]; }
            }
            /* ParallelepipedImpl::drawPos :275-281 ; TriangularPrismImpl :309-320 ; TetrahedronImpl :351-377 */
            /* `Vector3d coord(dist(gen), dist(gen), dist(gen))`: argument evaluation order is unspecified in C++.
             * Default: left to right (what the GPU path implements).  g_arg_rtl: right to left, which is what g++
             * does when it compiles the reference (oracle/_ref) -- used to compare against that binary word for word. */
            double c0, c1, c2;
            if (g_arg_rtl) { c2 = uniform01(g); c1 = uniform01(g); c0 = uniform01(g); }
            else           { c0 = uniform01(g); c1 = uniform01(g); c2 = uniform01(g); }
            const bool tet = sd.cell == MCB_CELL_TETRAHEDRON || sd.cell == MCB_CELL_PYRAMID;
            if (sd.cell != MCB_CELL_PARALLELEPIPED) {
                if (c0 + c1 > 1.) { c0 = 1. - c0; c1 = 1. - c1; }
                if (tet) {
                    if (c1 + c2 > 1.) { double tmp = c2; c2 = 1. - c0 - c1; c1 = 1. - tmp; }
                    else if (c0 + c1 + c2 > 1.) { double tmp = c2; c2 = c0 + c1 + c2 - 1.; c0 = 1. - c1 - tmp; }
                }
            }
            pos = sd.o + m * V3(c0, c1, c2);
            dir = sd.emitRot * drawAniso(g, true);                 /* subdomain.cpp:255-258 */
            sign = dot(dir, sd.gradT) < 0.;                        /* subdomain.cpp:260-263 */
        } else {
            const OBoundary& B = planes[e.index];
            V3 sp;
            if (B.shape == MCB_SHAPE_PARALLELOGRAM) {              /* boundary.cpp:147-152 */
                double r1 = uniform01(g), r2 = uniform01(g);
                sp = r1 * B.verts[0] + r2 * B.verts[1];
            } else if (B.shape == MCB_SHAPE_TRIANGLE) {            /* boundary.cpp:182-187 */
                double r1 = uniform01(g), r2 = uniform01(g);
                sp = (r1 + r2 < 1.) ? r1 * B.verts[0] + r2 * B.verts[1]
                                    : (1. - r1) * B.verts[0] + (1. - r2) * B.verts[1];
            } else {                                               /* boundary.cpp:243-251 */
                std::vector<double> areas;
                for (size_t n = 0; n + 1 < B.verts.size(); ++n) areas.push_back(norm(cross(B.verts[n], B.verts[n + 1])) / 2.);
                Discrete ad(areas.data(), areas.data() + areas.size());
                long n = ad(g);
                double r1 = uniform01(g), r2 = uniform01(g);
                sp = (r1 + r2 < 1.) ? r1 * B.verts[n] + r2 * B.verts[n + 1]
                                    : (1. - r1) * B.verts[n] + (1. - r2) * B.verts[n + 1];
            }
            pos = B.o + sp;                                        /* boundary.cpp:418-421 */
            dir = B.rot * drawAniso(g, false);                     /* boundary.cpp:423-426 */
            sign = B.T >= 0.;                                      /* boundary.cpp:428-431 */
        }
        return Phonon(sign, w, p, pos, dir);
    }
    int emitSdom(const OEmitter& e) const { return e.kind == MCB_EMIT_SDOM ? e.index : planes[e.index].sdom; }
    double emitWeight(const OEmitter& e) const {
        return e.kind == MCB_EMIT_SDOM ? sdoms[e.index].emitWeight() : planes[e.index].emitWeight();
    }
};

namespace {

/* Subdomain::Subdomain subdomain.cpp:41-71 (+ EmitSubdomain :232-236) */
void initSubdomain(OSubdomain& sd, double vol, const V3& o, const M3& mat, const long div[3], const V3& gradT) {
    sd.vol = vol; sd.o = o; sd.mat = mat; sd.inv = inverse(mat);
    double maxn = 0.;
    for (int c = 0; c < 3; ++c) maxn = std::max(maxn, norm(mat.col(c)));
    sd.eps = 100. * DEPS * maxn;
    bool anyNeg = false; int pos = 0;
    for (int d = 0; d < 3; ++d) {
        sd.div[d] = div[d]; sd.shape[d] = std::max(div[d], 1l); sd.max[d] = std::max(div[d], 1l) - 1;
        if (div[d] < 0) anyNeg = true;
        if (div[d] > 0) pos++;
    }
    int dim = anyNeg ? -1 : pos;
    switch (dim) {
    case -1: for (int d = 0; d < 3; ++d) { sd.shape[d] = 0; sd.max[d] = 0; } sd.accum = -2; break;
    case 0: sd.accum = -1; break;
    case 1: { int dir = 0; for (int d = 1; d < 3; ++d) if (sd.div[d] > sd.div[dir]) dir = d; sd.accum = dir; break; }
    default: sd.accum = dim + 1; break;
    }
    sd.gradT = gradT;
    sd.emitRot = rotMatrix(gradT);               /* rotMatrix(gradT.normalized()) */
}

OBoundary makeBoundary(int kind, const V3& o, const V3& i, const V3& j, double T) {
    OBoundary B;
    B.kind = kind;
    B.shape = MCB_SHAPE_PARALLELOGRAM; B.verts = {i, j};
    B.normal = normalized(cross(i, j));          /* Parallelogram::normal boundary.cpp:137-140 */
    B.offset = -dot(B.normal, o);                /* Eigen Hyperplane(n, e) */
    B.o = o; B.T = T;
    B.rot = rotMatrix(B.normal);                 /* Diff: boundary.cpp:294-301; Emit: :390-392 */
    B.refl = reflMatrix(B.normal);               /* Spec: boundary.cpp:268-271 */
    return B;
}

/* Parallelepiped<Bac,Lef,Bot,Fro,Rig,Top> ctor subdomain.h:145-158 + init :187-192 +
 * Subdomain::addBdry subdomain.cpp:199-212 */
int addParallelepiped(orc_domain& D, const V3& o, const M3& mat, const long div[3], const V3& gradT,
                      const int kinds[6], const double T[6]) {
    OSubdomain sd;
    sd.cell = MCB_CELL_PARALLELEPIPED;
    initSubdomain(sd, determinant(mat), o, mat, div, gradT);
    int s = (int)D.sdoms.size();
    V3 c0 = mat.col(0), c1 = mat.col(1), c2 = mat.col(2);
    OBoundary b[6] = {
        makeBoundary(kinds[0], o,      c1, c2, T[0]),
        makeBoundary(kinds[1], o,      c2, c0, T[1]),
        makeBoundary(kinds[2], o,      c0, c1, T[2]),
        makeBoundary(kinds[3], o + c0, c2, c1, T[3]),
        makeBoundary(kinds[4], o + c1, c0, c2, T[4]),
        makeBoundary(kinds[5], o + c2, c1, c0, T[5])};
    for (int k = 0; k < 6; ++k) {
        b[k].sdom = s;
        int id = (int)D.planes.size();
        bool emitting = (b[k].kind == MCB_BDRY_ISOT || b[k].kind == MCB_BDRY_PERI);
        if (emitting && b[k].emitWeight() != 0.) { b[k].emitRegistered = true; sd.emitPlanes.push_back(id); }
        D.planes.push_back(b[k]);
        sd.planes.push_back(id);
    }
    D.sdoms.push_back(sd);
    return s;
}

OBoundary makeBoundaryShape(int kind, const V3& o, int shape, const std::vector<V3>& verts, double T) {
    OBoundary B;
    B.kind = kind; B.shape = shape; B.verts = verts;
    B.normal = normalized(cross(verts[0], verts[1]));   /* Parallelogram/Triangle/Polygon::normal boundary.cpp:137,172,231 */
    B.offset = -dot(B.normal, o);
    B.o = o; B.T = T;
    B.rot = rotMatrix(B.normal); B.refl = reflMatrix(B.normal);
    return B;
}
int finishCell(orc_domain& D, OSubdomain& sd, std::vector<OBoundary>& b) {
    int s = (int)D.sdoms.size();
    for (OBoundary& B : b) {
        B.sdom = s;
        int id = (int)D.planes.size();
        bool emitting = (B.kind == MCB_BDRY_ISOT || B.kind == MCB_BDRY_PERI);
        if (emitting && B.emitWeight() != 0.) { B.emitRegistered = true; sd.emitPlanes.push_back(id); }
        D.planes.push_back(B);
        sd.planes.push_back(id);
    }
    D.sdoms.push_back(sd);
    return s;
}
const int PAR = MCB_SHAPE_PARALLELOGRAM, TRI = MCB_SHAPE_TRIANGLE, POLY = MCB_SHAPE_POLYGON;

/* TriangularPrism<Bac,Lef,Bot,Dia,Top> subdomain.h:206-233 */
int addTriangularPrism(orc_domain& D, const V3& o, const M3& mat, const long div[3], const V3& gradT, const int kinds[5], const double T[5]) {
    OSubdomain sd; sd.cell = MCB_CELL_TRIPRISM;
    initSubdomain(sd, determinant(mat) / 2., o, mat, div, gradT);
    V3 c0 = mat.col(0), c1 = mat.col(1), c2 = mat.col(2);
    std::vector<OBoundary> b = {
        makeBoundaryShape(kinds[0], o,      PAR, {c1, c2}, T[0]),
        makeBoundaryShape(kinds[1], o,      PAR, {c2, c0}, T[1]),
        makeBoundaryShape(kinds[2], o,      TRI, {c0, c1}, T[2]),
        makeBoundaryShape(kinds[3], o + c0, PAR, {c2, c1 - c0}, T[3]),
        makeBoundaryShape(kinds[4], o + c2, TRI, {c1, c0}, T[4])};
    return finishCell(D, sd, b);
}
/* Tetrahedron<Bac,Lef,Bot,Dia> subdomain.h:289-314 */
int addTetrahedron(orc_domain& D, const V3& o, const M3& mat, const long div[3], const V3& gradT, const int kinds[4], const double T[4]) {
    OSubdomain sd; sd.cell = MCB_CELL_TETRAHEDRON;
    initSubdomain(sd, determinant(mat) / 6., o, mat, div, gradT);
    V3 c0 = mat.col(0), c1 = mat.col(1), c2 = mat.col(2);
    std::vector<OBoundary> b = {
        makeBoundaryShape(kinds[0], o,      TRI, {c1, c2}, T[0]),
        makeBoundaryShape(kinds[1], o,      TRI, {c2, c0}, T[1]),
        makeBoundaryShape(kinds[2], o,      TRI, {c0, c1}, T[2]),
        makeBoundaryShape(kinds[3], o + c0, TRI, {c2 - c0, c1 - c0}, T[3])};
    return finishCell(D, sd, b);
}
/* Prism<Bot,Top,Sid> subdomain.h:363-470 and Pyramid<Bot,Sid> :486-596.  cols = mat_ (col 0 = axis / apex vector, cols
 * 1..N-1 = base fan).  PrismImpl/PyramidImpl::volume (subdomain.cpp:385-394, 417-426) pairs columns i, i+1 for i < N-2. */
int addPrismLike(orc_domain& D, bool pyramid, const V3& o, const std::vector<V3>& cols, long div, const V3& gradT,
                 const std::vector<int>& kinds, const std::vector<double>& T) {
    const long N = (long)cols.size();
    OSubdomain sd; sd.cell = pyramid ? MCB_CELL_PYRAMID : MCB_CELL_PRISM;
    std::vector<double> vol((size_t)(N - 2));
    double total = 0.;
    for (long i = 0; i < N - 2; ++i) { vol[(size_t)i] = dot(cross(cols[(size_t)i], cols[(size_t)i + 1]), cols[0]) / (pyramid ? 6. : 2.); total += vol[(size_t)i]; }
    M3 base;                                                         /* matBase = (col 1, col N-1, col 0) */
    for (int k = 0; k < 3; ++k) { base(k, 0) = cols[1][k]; base(k, 1) = cols[(size_t)N - 1][k]; base(k, 2) = cols[0][k]; }
    long dv[3] = {div < 0 ? -1l : 0l, div < 0 ? -1l : 0l, div < 0 ? -1l : 0l};
    initSubdomain(sd, total, o, base, dv, gradT);
    sd.base = cols;
    sd.volDist = Discrete(vol.data(), vol.data() + (N - 2));
    std::vector<V3> bot(cols.begin() + 1, cols.end()), top(bot.rbegin(), bot.rend());
    std::vector<OBoundary> b;
    size_t t = 0;
    b.push_back(makeBoundaryShape(kinds[t], o, POLY, bot, T[t])); ++t;
    if (!pyramid) { b.push_back(makeBoundaryShape(kinds[t], o + cols[0], POLY, top, T[t])); ++t; }
    for (long ind = 0; ind < N; ++ind, ++t) {                        /* InitSidesF :428-452 / :553-578 */
        V3 p; if (ind != 0) p = p + cols[(size_t)ind];
        V3 i = cols[0]; if (pyramid && ind != 0) i = i - cols[(size_t)ind];
        V3 j = -p; if (ind != N - 1) j = j + cols[(size_t)ind + 1];
        b.push_back(makeBoundaryShape(kinds[t], o + p, pyramid ? TRI : PAR, {i, j}, T[t]));
    }
    return finishCell(D, sd, b);
}

/* makePair(InterBoundary&, InterBoundary&) boundary.cpp:361-369 */
void pairInter(orc_domain& D, int a, int b) { D.planes[a].pairs.push_back(b); D.planes[b].pairs.push_back(a); }
/* makePair(PeriBoundary&, PeriBoundary&, transl, rot) boundary.cpp:524-550 */
void pairPeri(orc_domain& D, int a, int b, const V3& transl, const M3& rot = M3::identity()) {
    OBoundary& A = D.planes[a]; OBoundary& B = D.planes[b];
    double T1 = A.T, T2 = B.T;
    A.T = T1 - T2; B.T = T2 - T1;
    A.periRot = rot; B.periRot = transpose(rot);
    A.periTransl = transl; B.periTransl = -(transpose(rot) * transl);
    A.pairs = {b}; B.pairs = {a};
}
inline int pl(const orc_domain& D, int s, int k) { return D.sdoms[s].planes[k]; }

/* Domain::addSdom domain.cpp:88-102 for every sdom in order, then Field::init field.cpp:25-45 */
void finishDomain(orc_domain& D) {
    D.emitters.clear();
    for (int s = 0; s < (int)D.sdoms.size(); ++s) {
        if (D.sdoms[s].emitWeight() != 0.) D.emitters.push_back({MCB_EMIT_SDOM, s});
        for (int b : D.sdoms[s].emitPlanes) D.emitters.push_back({MCB_EMIT_BDRY, b});
    }
    D.cols = 0; D.colOffset.assign(D.sdoms.size(), -1);
    for (size_t s = 0; s < D.sdoms.size(); ++s) {
        long sp = D.sdoms[s].shapeProd();
        if (sp == 0) continue;
        D.colOffset[s] = D.cols; D.cols += sp;
    }
    /* flatten */
    D.fs.clear(); D.fp.clear(); D.fpairs.clear(); D.fe.clear();
    for (const OBoundary& B : D.planes) {
        mcb_plane_desc p; std::memset(&p, 0, sizeof p);
        p.normal[0] = B.normal.x; p.normal[1] = B.normal.y; p.normal[2] = B.normal.z; p.offset = B.offset;
        p.kind = B.kind; p.sdom = B.sdom;
        p.pair_begin = (int32_t)D.fpairs.size(); p.pair_count = (int32_t)B.pairs.size();
        for (int q : B.pairs) D.fpairs.push_back(q);
        std::memcpy(p.rot, B.rot.m, sizeof p.rot);
        std::memcpy(p.peri_rot, B.periRot.m, sizeof p.peri_rot);
        p.peri_transl[0] = B.periTransl.x; p.peri_transl[1] = B.periTransl.y; p.peri_transl[2] = B.periTransl.z;
        /* only EmitBoundary (Isot, Peri) keeps T_, o_ and its shape (boundary.h:203-223) */
        if (B.kind == MCB_BDRY_ISOT || B.kind == MCB_BDRY_PERI) {
            p.T = B.T; p.origin[0] = B.o.x; p.origin[1] = B.o.y; p.origin[2] = B.o.z;
            p.shape = B.shape; p.nvert = (int32_t)B.verts.size();
            for (size_t v = 0; v < B.verts.size() && v < MCB_MAX_VERTS; ++v) {
                p.verts[3 * v] = B.verts[v].x; p.verts[3 * v + 1] = B.verts[v].y; p.verts[3 * v + 2] = B.verts[v].z;
            }
        }
        D.fp.push_back(p);
    }
    for (const OSubdomain& S : D.sdoms) {
        mcb_sdom_desc s; std::memset(&s, 0, sizeof s);
        s.origin[0] = S.o.x; s.origin[1] = S.o.y; s.origin[2] = S.o.z;
        std::memcpy(s.mat, S.mat.m, sizeof s.mat); std::memcpy(s.inv, S.inv.m, sizeof s.inv);
        for (int d = 0; d < 3; ++d) { s.div[d] = S.div[d]; s.shape[d] = S.shape[d]; s.max[d] = S.max[d]; }
        s.accum = S.accum; s.cell = S.cell; s.eps = S.eps; s.vol = S.vol;
        s.grad_t[0] = S.gradT.x; s.grad_t[1] = S.gradT.y; s.grad_t[2] = S.gradT.z;
        std::memcpy(s.emit_rot, S.emitRot.m, sizeof s.emit_rot);
        s.plane_begin = S.planes.empty() ? 0 : S.planes.front(); s.plane_count = (int32_t)S.planes.size();
        s.nbase = (int32_t)S.base.size();
        for (size_t v = 0; v < S.base.size() && v < MCB_MAX_BASE; ++v) { s.base[3 * v] = S.base[v].x; s.base[3 * v + 1] = S.base[v].y; s.base[3 * v + 2] = S.base[v].z; }
        D.fs.push_back(s);
    }
    for (const OEmitter& e : D.emitters) D.fe.push_back({e.kind, e.index, D.emitWeight(e)});
    D.fcellvol.assign((size_t)D.cols, 0.);
    orc_domain_cell_vol(&D, D.fcellvol.data());
}

const int SPEC = MCB_BDRY_SPEC, DIFF = MCB_BDRY_DIFF, INTER = MCB_BDRY_INTER, ISOT = MCB_BDRY_ISOT, PERI = MCB_BDRY_PERI;
const double T0[6] = {0., 0., 0., 0., 0., 0.};

} // namespace

/* ----------------------------------------------------------------------------- Field */
namespace {

/* Field::accumulate field.cpp:92-220; data is rows x cols column-major */
void accumulate(const orc_domain& D, std::vector<double>& data, long rows, int s,
                const V3& bpos, const V3& epos, const double* amount) {
    const OSubdomain& sd = D.sdoms[s];
    int flag = sd.accum;
    if (flag < -1) return;
    long off = D.colOffset[s];
    long s1 = sd.shape[0], s2 = sd.shape[0] * sd.shape[1];
    auto col = [&](const long idx[3]) -> double* { return &data[(size_t)rows * (size_t)(off + idx[0] + s1 * idx[1] + s2 * idx[2])]; };
    if (flag < 0) {
        long z[3] = {0, 0, 0}; double* c = col(z);
        for (long r = 0; r < rows; ++r) c[r] += amount[r];
        return;
    }
    V3 bcoord = D.coord(s, bpos), ecoord = D.coord(s, epos);
    V3 dcoord = ecoord - bcoord;
    long bindex[3], eindex[3];
    D.coord2index(s, bcoord, bindex); D.coord2index(s, ecoord, eindex);

    if (flag < 3) {
        int d = flag;
        long b = bindex[d], e = eindex[d];
        long bvec[3] = {0, 0, 0}, evec[3] = {0, 0, 0};
        bvec[d] = b; evec[d] = e;
        if (b == e) { double* c = col(bvec); for (long r = 0; r < rows; ++r) c[r] += amount[r]; return; }
        std::vector<double> cellAmount(rows);
        double ad = std::abs(dcoord[d]);
        for (long r = 0; r < rows; ++r) cellAmount[r] = amount[r] / ad;
        int pm;
        double* cb = col(bvec); double* ce = col(evec);
        if (b < e) {
            double fb = (1 + b - bcoord[d]), fe = (ecoord[d] - e);
            for (long r = 0; r < rows; ++r) cb[r] += cellAmount[r] * fb;
            for (long r = 0; r < rows; ++r) ce[r] += cellAmount[r] * fe;
            pm = 1;
        } else {
            double fb = (bcoord[d] - b), fe = (1 + e - ecoord[d]);
            for (long r = 0; r < rows; ++r) cb[r] += cellAmount[r] * fb;
            for (long r = 0; r < rows; ++r) ce[r] += cellAmount[r] * fe;
            pm = -1;
        }
        long nvec[3] = {0, 0, 0};
        for (long n = b + pm; n != e; n += pm) {
            nvec[d] = n; double* c = col(nvec);
            for (long r = 0; r < rows; ++r) c[r] += cellAmount[r];
        }
    } else {
        struct Step { long v[3]; };
        std::map<double, Step> borders;
        auto ins = borders.begin();
        for (int d = 0; d < 3; d++) {
            if (std::abs(dcoord[d]) < DMIN) continue;
            long b = bindex[d], e = eindex[d];
            int pm;
            if (b == e) continue;
            else if (b < e) { b++; e++; pm = 1; }
            else pm = -1;
            ins = borders.begin();
            Step step{{0, 0, 0}}; step.v[d] = pm;
            bool search = !borders.empty();
            for (long n = b; n != e; n += pm) {
                double param = (n - bcoord[d]) / dcoord[d];
                if (search) {
                    auto found = borders.find(param);
                    if (found != borders.end()) {
                        for (int q = 0; q < 3; ++q) found->second.v[q] += step.v[q];
                        ins = found;
                        continue;
                    }
                }
                ins = borders.insert(ins, std::make_pair(param, step));
            }
        }
        borders.insert(ins, std::make_pair(1., Step{{0, 0, 0}}));
        long index[3] = {bindex[0], bindex[1], bindex[2]};
        double param = 0.;
        for (auto it = borders.begin(); it != borders.end(); ++it) {
            double* c = col(index);
            double wgt = (it->first - param);
            for (long r = 0; r < rows; ++r) c[r] += amount[r] * wgt;
            param = it->first;
            for (int q = 0; q < 3; ++q) index[q] += it->second.v[q];
        }
    }
}

} // namespace

/* --------------------------------------------------------------------------- Problem */
struct orc_problem {
    const orc_material* mat; const orc_domain* dom;
    int kind; long rows; long size = 0, step = 0;
    long nemit = 0, maxscat = 0, maxloop = 0; double power = 0.;
    std::vector<int64_t> emitPdf;

    /* accumAmt family problem.cpp:473-476,506-509,539-544,581-589,629-637 (sign applied at :414) */
    void accumAmt(const Phonon& pre, const Phonon& post, double* vec) const {
        for (long r = 0; r < rows; ++r) vec[r] = 0.;
        double dtime = post.time - pre.time;
        V3 dpos = post.pos - pre.pos;
        switch (kind) {
        case MCB_PROB_TEMP: vec[0] = dtime; break;
        case MCB_PROB_FLUX: vec[0] = dpos.x; vec[1] = dpos.y; vec[2] = dpos.z; break;
        case MCB_PROB_MULTI: vec[0] = dtime; vec[1] = dpos.x; vec[2] = dpos.y; vec[3] = dpos.z; break;
        case MCB_PROB_CUMTEMP: { long index = (pre.nscat + step - 1) / step; vec[index] = dtime; break; }
        case MCB_PROB_CUMFLUX: { long index = 3 * ((pre.nscat + step - 1) / step);
            vec[index] = dpos.x; vec[index + 1] = dpos.y; vec[index + 2] = dpos.z; break; }
        }
    }
    /* postProc family problem.cpp:478-481,511-514,546-551,591-599,639-648, then :439-444 */
    void finalize(double* f) const {
        long cols = dom->cols;
        double es = mat->energySum;
        switch (kind) {
        case MCB_PROB_TEMP: for (long i = 0; i < rows * cols; ++i) f[i] = f[i] / es; break;
        case MCB_PROB_FLUX: break;
        case MCB_PROB_MULTI: for (long c = 0; c < cols; ++c) f[rows * c] /= es; break;
        case MCB_PROB_CUMTEMP:
            for (long c = 0; c < cols; ++c) for (long i = 0; i < size; ++i) f[rows * c + i + 1] += f[rows * c + i];
            for (long i = 0; i < rows * cols; ++i) f[i] = f[i] / es;
            break;
        case MCB_PROB_CUMFLUX:
            for (long c = 0; c < cols; ++c) for (long i = 0; i < size; ++i)
                for (int q = 0; q < 3; ++q) f[rows * c + 3 * (i + 1) + q] += f[rows * c + 3 * i + q];
            break;
        }
        std::vector<double> vol(cols);
        orc_domain_cell_vol(dom, vol.data());
        for (long c = 0; c < cols; ++c) for (long r = 0; r < rows; ++r) f[rows * c + r] = power * (f[rows * c + r] / vol[c]);
    }
};

namespace {

/* One particle history: problem.cpp:386-436.  `trace_steps` >= 0 stops after that many trips. */
struct HistoryOut { long steps = 0; bool esc = false; int sdom = -1; };

HistoryOut runHistory(const orc_problem& P, long n, const long* cdfBegin, const long* cdfEnd, Words& g,
                      std::vector<double>* field, long traceSteps, Phonon* finalState) {
    const orc_domain& D = *P.dom; const orc_material& M = *P.mat;
    HistoryOut out;
    long emitIndex = std::upper_bound(cdfBegin, cdfEnd, n) - cdfBegin;            /* :386-387 */
    const OEmitter& e = D.emitters.at((size_t)emitIndex);
    int sdom = D.emitSdom(e);
    if (g.mode == ORC_RNG_PHILOX) g.begin((uint64_t)n, 0);
    long w, p; M.fluxDist.draw(g, w, p);                                           /* :393 */
    Phonon phn = D.emit(e, w, p, g);                                               /* :397 */
    M.drawScatNext(phn, g);                                                        /* :399 */
    std::vector<double> amount(P.rows);
    long limit = traceSteps >= 0 ? std::min(traceSteps, P.maxloop) : P.maxloop;
    for (long i = 0; i < limit; i++) {                                             /* :401 */
        if (g.mode == ORC_RNG_PHILOX) g.begin((uint64_t)n, (uint32_t)(i + 1));
        Phonon pre(phn);                                                           /* :403 */
        double vel = M.velAt(phn);                                                 /* :405 */
        int bdry = D.advect(sdom, phn, vel);                                       /* :406 */
        out.steps++;
        if (!phn.alive) { out.esc = true; break; }                                 /* :408-412 */
        if (field) {
            P.accumAmt(pre, phn, amount.data());                                   /* :414 */
            int sg = phn.sgn();
            for (long r = 0; r < P.rows; ++r) amount[r] = sg * amount[r];
            accumulate(D, *field, P.rows, sdom, pre.pos, phn.pos, amount.data());  /* :416 */
        }
        if (bdry >= 0) {                                                           /* :418-429 */
            bdry = D.scatter(bdry, phn, g);
            if (bdry < 0) { out.esc = true; break; }
            sdom = D.planes[bdry].sdom;
        } else {
            M.scatter(phn, g);                                                     /* :430-433 */
        }
        if (!phn.alive || phn.nscat >= P.maxscat) break;                           /* :434 */
    }
    out.sdom = sdom;
    if (finalState) *finalState = phn;
    return out;
}

int solveImpl(const orc_problem* P, int rng_mode, uint64_t seed, int64_t n_begin, int64_t n_end,
              int nthreads, double* out, mcb_stats* stats, bool finalize) {
    if (!P || !out) { set_err("null argument"); return MCB_EINVAL; }
    if (n_begin < 0 || n_end > P->nemit || n_begin > n_end) { set_err("bad particle range"); return MCB_EINVAL; }
    const long rows = P->rows, cols = P->dom->cols;
    std::vector<long> cdf(P->emitPdf.begin(), P->emitPdf.end());                   /* :374-381 */
    for (size_t i = 1; i < cdf.size(); ++i) cdf[i] += cdf[i - 1];
    if (nthreads <= 0) nthreads = orc_max_threads();
    std::vector<double> sol((size_t)(rows * cols), 0.);
    long totalSteps = 0, totalEsc = 0;
    std::vector<std::vector<double>> partials((size_t)nthreads);
#pragma omp parallel num_threads(nthreads) reduction(+ : totalSteps, totalEsc)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        Words g(rng_mode, seed + (uint64_t)tid);       /* DEBUG seeds 0,1,2,... main.cpp:29-39 */
        if (rng_mode == ORC_RNG_PHILOX) g = Words(rng_mode, seed);
        std::vector<double> fld((size_t)(rows * cols), 0.);                        /* Field fld(rows(), dom()) :372 */
#pragma omp for schedule(static)
        for (long n = n_begin; n < n_end; ++n) {                                   /* :383-384 */
            HistoryOut h = runHistory(*P, n, cdf.data(), cdf.data() + cdf.size(), g, &fld, -1, nullptr);
            totalSteps += h.steps; if (h.esc) totalEsc++;
        }
        partials[(size_t)tid].swap(fld);
    }
    /* main.cpp:162-165 `sol += partial` (summed here in thread order so runs are reproducible;
     * finalisation is linear so applying it after the sum equals summing finalised partials) */
    for (int t = 0; t < nthreads; ++t)
        if (!partials[(size_t)t].empty()) for (size_t i = 0; i < sol.size(); ++i) sol[i] += partials[(size_t)t][i];
    if (finalize) P->finalize(sol.data());
    std::memcpy(out, sol.data(), sol.size() * sizeof(double));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->emitted = n_end - n_begin; stats->steps = totalSteps; stats->esc = totalEsc; stats->cols = cols;
    }
    return MCB_OK;
}

} // namespace

/* ============================================================================ C API */
extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

void orc_set_arg_order(int right_to_left) { g_arg_rtl = right_to_left ? 1 : 0; }

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Material::Material material.cpp:82-162 */
orc_material* orc_material_load(const char* disp, const char* relax, double temp) {
    std::unique_ptr<orc_material> M(new orc_material);
    M->T = temp;
    std::ifstream dispFile(disp);
    if (!dispFile) { set_err(std::string("Error opening dispersion file ") + disp); return nullptr; }
    dispFile >> M->nw >> M->np >> std::ws;
    if (!(M->nw > 0 && M->np > 0)) { set_err("Invalid dispersion file"); return nullptr; }
    const long nw = M->nw, np = M->np;
    std::vector<double> dispData;
    if (!extractArray(dispFile, dispData, nw, 2 + 2 * np)) { set_err("Array extraction failed (disp)"); return nullptr; }
    M->omega.resize(nw); M->vel.resize(nw * np);
    std::vector<double> domega(nw * np), dos(nw * np);
    const long dc = 2 + 2 * np;
    for (long w = 0; w < nw; ++w) {
        M->omega[w] = dispData[w * dc + 0];
        for (long p = 0; p < np; ++p) {
            domega[w + nw * p] = dispData[w * dc + 1];
            M->vel[w + nw * p] = dispData[w * dc + 2 + 2 * p];
            dos[w + nw * p]    = dispData[w * dc + 3 + 2 * p];
        }
    }
    std::ifstream relaxFile(relax);
    if (!relaxFile) { set_err(std::string("Error opening relaxation time file ") + relax); return nullptr; }
    std::vector<double> coeffs;                       /* np rows x 8: coeffs(4*j+k, p) at [p*8 + 4*j + k] */
    if (!extractArray(relaxFile, coeffs, np, 8)) { set_err("Array extraction failed (relax)"); return nullptr; }
    /* tau  material.cpp:116-134 */
    M->tau.assign(nw * np, 0.);
    for (long p = 0; p < np; ++p) {
        std::vector<double> tauinv(nw, 0.);
        for (int j = 0; j < 2; ++j) {
            const double* c = &coeffs[p * 8 + 4 * j];
            if (c[0] < 0.) { set_err("Scattering times cannot be negative"); return nullptr; }
            if (c[0] <= DMIN) continue;
            for (long w = 0; w < nw; ++w)
                tauinv[w] += (c[0] * std::pow(M->omega[w], c[1]) * std::pow(M->T, c[2]) * std::exp(-c[3] / M->T));
        }
        for (long w = 0; w < nw; ++w) M->tau[w + nw * p] = 1. / tauinv[w];
    }
    for (double t : M->tau) if (!std::isfinite(t)) { set_err("Scattering time model produced infinite values"); return nullptr; }
    /* dedT  material.cpp:136-145 */
    std::vector<double> dedT(nw * np);
    for (long w = 0; w < nw; ++w) {
        double x = HBAR / (KB * M->T) * M->omega[w];
        double val = KB * (std::fabs(x) < DEPS ? 1. - x * x / 12. : std::pow(x / (2. * std::sinh(x / 2.)), 2));
        for (long p = 0; p < np; ++p) dedT[w + nw * p] = val;
    }
    /* distributions  material.cpp:147-158 */
    M->energyPdf.resize(nw * np); M->fluxPdf.resize(nw * np); M->scatPdf.resize(nw * np);
    for (long i = 0; i < nw * np; ++i) M->energyPdf[i] = dedT[i] * dos[i] * domega[i];
    M->energyDist = Dist(M->energyPdf, nw, np); M->energySum = sum_colmajor(M->energyPdf);
    for (long i = 0; i < nw * np; ++i) M->fluxPdf[i] = M->vel[i] * M->energyPdf[i];
    M->fluxDist = Dist(M->fluxPdf, nw, np); M->fluxSum = sum_colmajor(M->fluxPdf);
    for (long i = 0; i < nw * np; ++i) M->scatPdf[i] = M->energyPdf[i] / M->tau[i];
    M->scatDist = Dist(M->scatPdf, nw, np); M->scatSum = sum_colmajor(M->scatPdf);
    /* k  material.cpp:160-161 */
    double ks = 0.;
    for (long i = 0; i < nw * np; ++i) ks += M->tau[i] * std::pow(M->vel[i], 2) * M->energyPdf[i];
    M->k = ks / 3.;
    return M.release();
}
void orc_material_free(orc_material* m) { delete m; }
double orc_material_cond(const orc_material* m) { return m->k; }
int orc_material_desc(const orc_material* m, mcb_material_desc* o) {
    if (!m || !o) return MCB_EINVAL;
    o->nw = m->nw; o->np = m->np; o->temp = m->T;
    o->vel = m->vel.data(); o->tau = m->tau.data(); o->flux_pdf = m->fluxPdf.data(); o->scat_pdf = m->scatPdf.data();
    o->energy_sum = m->energySum; o->flux_sum = m->fluxSum; o->scat_sum = m->scatSum;
    return MCB_OK;
}
int orc_material_alias(const orc_material* m, int which, double* wprob, int32_t* walias, double* pprob, int32_t* palias) {
    if (!m) return MCB_EINVAL;
    const Dist& d = which == 0 ? m->fluxDist : m->scatDist;
    for (long w = 0; w < m->nw; ++w) {
        wprob[w] = d.wDist.prob[w]; walias[w] = (int32_t)d.wDist.alias[w];
        for (long p = 0; p < m->np; ++p) {
            pprob[w * m->np + p] = d.pDist[w].prob[p]; palias[w * m->np + p] = (int32_t)d.pDist[w].alias[p];
        }
    }
    return MCB_OK;
}

orc_domain* orc_domain_box(const double origin[3], const double mat[9], const int64_t div[3],
                           const double grad_t[3], const int32_t kinds[6], const double T[6]) {
    std::unique_ptr<orc_domain> D(new orc_domain);
    M3 A; std::memcpy(A.m, mat, sizeof A.m);
    long dv[3] = {(long)div[0], (long)div[1], (long)div[2]};
    int k[6]; for (int i = 0; i < 6; ++i) k[i] = kinds[i];
    if (!(determinant(A) >= DMIN)) { set_err("Volume too small, check vector order"); return nullptr; }
    for (int i = 0; i < 6; ++i) if (k[i] == INTER) { set_err("Inter boundary needs a partner subdomain"); return nullptr; }
    for (int i = 0; i < 3; ++i) if ((k[i] == PERI) != (k[i + 3] == PERI)) { set_err("Peri faces must come in opposite pairs"); return nullptr; }
    addParallelepiped(*D, V3(origin[0], origin[1], origin[2]), A, dv, V3(grad_t[0], grad_t[1], grad_t[2]), k, T);
    for (int i = 0; i < 3; ++i) if (k[i] == PERI) pairPeri(*D, pl(*D, 0, i), pl(*D, 0, i + 3), A.col(i));
    finishDomain(*D);
    return D.release();
}

orc_domain* orc_domain_create(const char* kind, const double* dim, int ndim, const int64_t* div, int ndiv, double dT) {
    std::unique_ptr<orc_domain> D(new orc_domain);
    std::string k(kind);
    auto need = [&](int nd, int nv) { if (ndim != nd || ndiv != nv) { set_err("wrong number of dims/divs for domain " + k); return false; } return true; };
    auto box = [&](double ox, double oy, double oz, double a, double b, double c, long d0, long d1, long d2,
                   const V3& gradT, std::initializer_list<int> kinds) {
        long dv[3] = {d0, d1, d2}; int kk[6]; int i = 0; for (int q : kinds) kk[i++] = q;
        return addParallelepiped(*D, V3(ox, oy, oz), M3::diag(a, b, c), dv, gradT, kk, T0);
    };
    if (k == "bulk" || k == "film") {                 /* domain.cpp:137-148, 188-199 */
        if (!need(3, 3)) return nullptr;
        V3 gradT(-dT / dim[0], 0., 0.);
        if (k == "bulk") box(0, 0, 0, dim[0], dim[1], dim[2], div[0], div[1], div[2], gradT, {PERI, SPEC, SPEC, PERI, SPEC, SPEC});
        else             box(0, 0, 0, dim[0], dim[1], dim[2], div[0], div[1], div[2], gradT, {PERI, DIFF, SPEC, PERI, DIFF, SPEC});
        pairPeri(*D, pl(*D, 0, 0), pl(*D, 0, 3), V3(dim[0], 0., 0.));
    } else if (k == "jct") {                          /* domain.h:187-190, domain.cpp:352-385 */
        if (!need(4, 4)) return nullptr;
        V3 gradT(-dT / (2. * dim[0]), 0., 0.);
        box(0., 0., 0.,         2. * dim[0], dim[1], dim[3], 2 * div[0], div[1], div[3], gradT, {PERI, SPEC, DIFF, PERI, INTER, DIFF});
        box(0., dim[1], 0.,     dim[0], dim[2], dim[3],      div[0], div[2], div[3],     gradT, {PERI, INTER, DIFF, INTER, SPEC, DIFF});
        box(dim[0], dim[1], 0., dim[0], dim[2], dim[3],      div[0], div[2], div[3],     gradT, {INTER, INTER, DIFF, PERI, SPEC, DIFF});
        pairInter(*D, pl(*D, 0, 4), pl(*D, 1, 1));
        pairInter(*D, pl(*D, 0, 4), pl(*D, 2, 1));
        pairInter(*D, pl(*D, 1, 3), pl(*D, 2, 0));
        V3 transl(2. * dim[0], 0., 0.);
        pairPeri(*D, pl(*D, 0, 0), pl(*D, 0, 3), transl);
        pairPeri(*D, pl(*D, 1, 0), pl(*D, 2, 3), transl);
    } else if (k == "tee") {                          /* domain.h:225-229, domain.cpp:429-468 */
        if (!need(5, 5)) return nullptr;
        V3 gradT(-dT / (2. * dim[0] + dim[1]), 0., 0.);
        box(0., 0., 0.,              dim[0], dim[2], dim[4], div[0], div[2], div[4], gradT, {PERI, DIFF, SPEC, INTER, DIFF, SPEC});
        box(dim[0], 0., 0.,          dim[1], dim[2], dim[4], div[1], div[2], div[4], gradT, {INTER, SPEC, SPEC, INTER, INTER, SPEC});
        box(dim[0], dim[2], 0.,      dim[1], dim[3], dim[4], div[1], div[3], div[4], gradT, {DIFF, INTER, SPEC, DIFF, SPEC, SPEC});
        box(dim[0] + dim[1], 0., 0., dim[0], dim[2], dim[4], div[0], div[2], div[4], gradT, {INTER, DIFF, SPEC, PERI, DIFF, SPEC});
        pairInter(*D, pl(*D, 0, 3), pl(*D, 1, 0));
        pairInter(*D, pl(*D, 1, 4), pl(*D, 2, 1));
        pairInter(*D, pl(*D, 1, 3), pl(*D, 3, 0));
        pairPeri(*D, pl(*D, 0, 0), pl(*D, 3, 3), V3(2. * dim[0] + dim[1], 0., 0.));
    } else if (k == "tube") {                         /* domain.h:264-267, domain.cpp:509-540 */
        if (!need(4, 4)) return nullptr;
        V3 gradT(-dT / dim[0], 0., 0.);
        box(0., dim[1], 0.,     dim[0], dim[3], dim[2], div[0], div[3], div[2], gradT, {PERI, DIFF, SPEC, PERI, DIFF, INTER});
        box(0., dim[1], dim[2], dim[0], dim[3], dim[3], div[0], div[3], div[3], gradT, {PERI, INTER, INTER, PERI, DIFF, DIFF});
        box(0., 0., dim[2],     dim[0], dim[1], dim[3], div[0], div[1], div[3], gradT, {PERI, SPEC, DIFF, PERI, INTER, DIFF});
        pairInter(*D, pl(*D, 1, 1), pl(*D, 2, 4));
        pairInter(*D, pl(*D, 1, 2), pl(*D, 0, 5));
        V3 transl(dim[0], 0., 0.);
        pairPeri(*D, pl(*D, 0, 0), pl(*D, 0, 3), transl);
        pairPeri(*D, pl(*D, 1, 0), pl(*D, 1, 3), transl);
        pairPeri(*D, pl(*D, 2, 0), pl(*D, 2, 3), transl);
    } else if (k == "hex") {                          /* domain.h:142-143, domain.cpp:238-256 (+ init(), which the reference forgets) */
        if (!need(4, 0)) return nullptr;
        std::vector<V3> cols = {V3(dim[0], 0., 0.), V3(0., dim[1], -dim[3]), V3(0., 2. * dim[1], 0.), V3(0., 2. * dim[1], dim[2]),
                                V3(0., dim[1], dim[2] + dim[3]), V3(0., 0., dim[2])};
        addPrismLike(*D, false, V3(), cols, 0, V3(-dT / dim[0], 0., 0.), {PERI, PERI, SPEC, SPEC, SPEC, SPEC, SPEC, SPEC}, std::vector<double>(8, 0.));
        pairPeri(*D, pl(*D, 0, 0), pl(*D, 0, 1), V3(dim[0], 0., 0.));
    } else if (k == "pyr") {                          /* domain.h:165, domain.cpp:299-314 */
        if (!need(3, 0)) return nullptr;
        std::vector<V3> cols = {V3(dim[0], 0.5 * dim[1], 0.5 * dim[2]), V3(0., dim[1], 0.), V3(0., dim[1], dim[2]), V3(0., 0., dim[2])};
        addPrismLike(*D, true, V3(), cols, 0, V3(-dT / dim[0], 0., 0.), {SPEC, SPEC, SPEC, SPEC, SPEC}, std::vector<double>(5, 0.));
    } else {
        set_err("Invalid domain " + k + " (oracle builds bulk, film, jct, tee, tube, hex, pyr, orc_domain_box and orc_domain_cell)");
        return nullptr;
    }
    for (const OSubdomain& s : D->sdoms) if (!(s.vol >= DMIN)) { set_err("Volume too small, check vector order"); return nullptr; }
    finishDomain(*D);
    return D.release();
}
/* One non-box cell (MCB_CELL_TRIPRISM / TETRAHEDRON: cols = 3 mat columns, div[3]; PRISM / PYRAMID: cols = N mat_ columns,
 * div[0] only) with the given boundary kinds and wall temperatures in declaration order; no pairing (Spec/Diff/Isot). */
orc_domain* orc_domain_cell(int cell, const double origin[3], const double* cols, int ncols, const int64_t div[3],
                            const double grad_t[3], const int32_t* kinds, const double* T) {
    std::unique_ptr<orc_domain> D(new orc_domain);
    V3 o(origin[0], origin[1], origin[2]), g(grad_t[0], grad_t[1], grad_t[2]);
    std::vector<V3> c; for (int i = 0; i < ncols; ++i) c.push_back(V3(cols[3 * i], cols[3 * i + 1], cols[3 * i + 2]));
    long dv[3] = {(long)div[0], (long)div[1], (long)div[2]};
    if (cell == MCB_CELL_TRIPRISM || cell == MCB_CELL_TETRAHEDRON) {
        if (ncols != 3) { set_err("tri-prism / tetrahedron need 3 columns"); return nullptr; }
        M3 m; for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) m(k, i) = c[(size_t)i][k];
        int kk[5]; for (int i = 0; i < (cell == MCB_CELL_TRIPRISM ? 5 : 4); ++i) kk[i] = kinds[i];
        if (cell == MCB_CELL_TRIPRISM) addTriangularPrism(*D, o, m, dv, g, kk, T); else addTetrahedron(*D, o, m, dv, g, kk, T);
    } else if (cell == MCB_CELL_PRISM || cell == MCB_CELL_PYRAMID) {
        if (ncols < 4 || ncols > MCB_MAX_BASE) { set_err("prism / pyramid need 4..9 columns"); return nullptr; }
        const int nb = ncols + (cell == MCB_CELL_PRISM ? 2 : 1);
        addPrismLike(*D, cell == MCB_CELL_PYRAMID, o, c, dv[0], g, std::vector<int>(kinds, kinds + nb), std::vector<double>(T, T + nb));
    } else { set_err("orc_domain_cell: unknown cell kind"); return nullptr; }
    if (!(D->sdoms[0].vol >= DMIN)) { set_err("Volume too small, check vector order"); return nullptr; }
    finishDomain(*D);
    return D.release();
}
void orc_domain_free(orc_domain* d) { delete d; }
int64_t orc_domain_cols(const orc_domain* d) { return d->cols; }
int orc_domain_desc(const orc_domain* d, mcb_domain_desc* o) {
    if (!d || !o) return MCB_EINVAL;
    o->nsdom = (int32_t)d->fs.size(); o->sdoms = d->fs.data();
    o->nplane = (int32_t)d->fp.size(); o->planes = d->fp.data();
    o->npair = (int32_t)d->fpairs.size(); o->pairs = d->fpairs.data();
    o->nemitter = (int32_t)d->fe.size(); o->emitters = d->fe.data();
    o->ncols = d->cols; o->cell_vol = d->fcellvol.data();
    return MCB_OK;
}
/* Field(1, dom, CellVolF()) field.cpp:53-80: columns in (k, j, i) nesting, i fastest */
int orc_domain_cell_vol(const orc_domain* d, double* vol) {
    long n = 0;
    for (size_t s = 0; s < d->sdoms.size(); ++s) {
        const long* sh = d->sdoms[s].shape;
        for (long k = 0; k < sh[2]; ++k) for (long j = 0; j < sh[1]; ++j) for (long i = 0; i < sh[0]; ++i) {
            long idx[3] = {i, j, k};
            vol[n++] = d->cellVol((int)s, idx);
        }
    }
    return MCB_OK;
}

/* FieldProblem::FieldProblem problem.cpp:315-342 (+ Cum* ctors :558-565, :606-612) */
orc_problem* orc_problem_create(const orc_material* mat, const orc_domain* dom, int kind,
                                int64_t nemit, int64_t size, int64_t maxscat, int64_t maxloop) {
    if (!mat || !dom) { set_err("null material/domain"); return nullptr; }
    if (dom->emitters.empty()) { set_err("Domain has no emitters"); return nullptr; }
    std::unique_ptr<orc_problem> P(new orc_problem);
    P->mat = mat; P->dom = dom; P->kind = kind;
    long nemitter = (long)dom->emitters.size();
    std::vector<double> weight(nemitter);
    double weightSum = 0.;
    for (long i = 0; i < nemitter; ++i) { weight[i] = dom->emitWeight(dom->emitters[i]); weightSum += weight[i]; }
    P->emitPdf.resize(nemitter);
    long total = 0;
    for (long i = 0; i < nemitter; ++i) {
        double frac = weight[i] / weightSum;
        double rounded = std::ceil(frac * nemit - 0.5);
        P->emitPdf[i] = std::max(1l, static_cast<long>(rounded));
        total += P->emitPdf[i];
    }
    P->nemit = total; P->maxscat = maxscat;
    P->maxloop = (maxloop != 0 ? maxloop : 100 * maxscat);
    P->power = weightSum / P->nemit * mat->fluxSum / 4.;
    switch (kind) {
    case MCB_PROB_TEMP: P->rows = 1; break;
    case MCB_PROB_FLUX: P->rows = 3; break;
    case MCB_PROB_MULTI: P->rows = 4; break;
    case MCB_PROB_CUMTEMP: case MCB_PROB_CUMFLUX:
        if (size <= 0) { set_err("Cum* problems need size > 0"); return nullptr; }
        P->size = size; P->step = (maxscat - 1) / size; if ((maxscat - 1) % size != 0) P->step++;
        P->rows = kind == MCB_PROB_CUMTEMP ? size + 1 : 3 * (size + 1);
        break;
    default: set_err("Invalid problem"); return nullptr;
    }
    return P.release();
}
void orc_problem_free(orc_problem* p) { delete p; }
int orc_problem_desc(const orc_problem* p, mcb_problem_desc* o) {
    if (!p || !o) return MCB_EINVAL;
    o->kind = p->kind; o->rows = (int32_t)p->rows; o->size = p->size; o->step = p->step; o->nemit = p->nemit;
    o->maxscat = p->maxscat; o->maxloop = p->maxloop; o->power = p->power; o->emit_count = p->emitPdf.data();
    return MCB_OK;
}

int orc_solve(const orc_problem* P, int rng_mode, uint64_t seed, int64_t n_begin, int64_t n_end, int nthreads,
              double* out, mcb_stats* stats) { return solveImpl(P, rng_mode, seed, n_begin, n_end, nthreads, out, stats, true); }
int orc_solve_raw(const orc_problem* P, int rng_mode, uint64_t seed, int64_t n_begin, int64_t n_end, int nthreads,
                  double* out, mcb_stats* stats) { return solveImpl(P, rng_mode, seed, n_begin, n_end, nthreads, out, stats, false); }
int orc_finalize(const orc_problem* P, double* f) { if (!P || !f) return MCB_EINVAL; P->finalize(f); return MCB_OK; }

int orc_trace(const orc_problem* P, uint64_t seed, int64_t n_begin, int64_t n_end, int64_t nsteps, mcb_trace_out* o) {
    if (!P || !o) return MCB_EINVAL;
    std::vector<long> cdf(P->emitPdf.begin(), P->emitPdf.end());
    for (size_t i = 1; i < cdf.size(); ++i) cdf[i] += cdf[i - 1];
    Words g(ORC_RNG_PHILOX, seed);
    for (long n = n_begin; n < n_end; ++n) {
        Phonon ph;
        HistoryOut h = runHistory(*P, n, cdf.data(), cdf.data() + cdf.size(), g, nullptr, nsteps, &ph);
        long i = n - n_begin;
        if (o->pos) { o->pos[3 * i] = ph.pos.x; o->pos[3 * i + 1] = ph.pos.y; o->pos[3 * i + 2] = ph.pos.z; }
        if (o->dir) { o->dir[3 * i] = ph.dir.x; o->dir[3 * i + 1] = ph.dir.y; o->dir[3 * i + 2] = ph.dir.z; }
        if (o->scat_next) o->scat_next[i] = ph.scatNext;
        if (o->w) o->w[i] = ph.w;
        if (o->p) o->p[i] = ph.p;
        if (o->sign) o->sign[i] = ph.sgn();
        if (o->alive) o->alive[i] = ph.alive ? 1 : 0;
        if (o->sdom) o->sdom[i] = h.sdom;
        if (o->nscat) o->nscat[i] = ph.nscat;
        if (o->steps) o->steps[i] = h.steps;
        if (o->cell) {
            long idx[3]; P->dom->coord2index(h.sdom, P->dom->coord(h.sdom, ph.pos), idx);
            o->cell[3 * i] = (int32_t)idx[0]; o->cell[3 * i + 1] = (int32_t)idx[1]; o->cell[3 * i + 2] = (int32_t)idx[2];
        }
    }
    return MCB_OK;
}

int orc_domain_locate(const orc_domain* d, const double pos[3]) {
    for (size_t s = 0; s < d->sdoms.size(); ++s) if (d->isInside((int)s, V3(pos[0], pos[1], pos[2]))) return (int)s;
    return -1;
}

/* TrajProblem::solve problem.cpp:226-299 with TrkPhonon's recording (phonon.cpp:129-170) */
int orc_traj(const orc_material* M, const orc_domain* D, const mcb_traj_desc* t, uint64_t seed, mcb_traj_out* o) {
    return orc_traj_rng(M, D, t, ORC_RNG_PHILOX, seed, o);
}
/* rng_mode = ORC_RNG_MT19937: one sequential mt19937(seed) like `Rng gen(s)` in solveTraj (main.cpp:86-103), to compare with
 * the reference binary word for word; ORC_RNG_PHILOX: the device's per-event streams of particle 0 */
int orc_traj_rng(const orc_material* M, const orc_domain* D, const mcb_traj_desc* t, int rng_mode, uint64_t seed, mcb_traj_out* o) {
    if (!M || !D || !t || !o) { set_err("null argument"); return MCB_EINVAL; }
    Words g(rng_mode, seed);
    g.begin(0, 0);
    auto push = [&](const V3& p) {
        if (o->npoints < o->max_points) { o->points[3 * o->npoints] = p.x; o->points[3 * o->npoints + 1] = p.y; o->points[3 * o->npoints + 2] = p.z; }
        o->npoints++;
    };
    o->npoints = 0; o->nsteps = 0; o->escaped = 0;
    int sdom = -1, bdry = -1;
    Phonon phn;
    long w = t->w, p = t->p;
    if (!t->has_prop) M->scatDist.draw(g, w, p);                                   /* :232 */
    if (t->has_pos) {                                                              /* :233-243 */
        V3 pos(t->pos[0], t->pos[1], t->pos[2]);
        V3 dir = t->has_dir ? V3(t->dir[0], t->dir[1], t->dir[2]) : drawIso(g);
        phn = Phonon(true, w, p, pos, dir);
        sdom = t->sdom; bdry = -1;
        if (sdom < 0 || sdom >= (int)D->sdoms.size()) { set_err("Position not inside domain"); return MCB_EINVAL; }
    } else {                                                                       /* :244-253 */
        const OEmitter& e = D->emitters.at((size_t)uniformInt(g, (long)D->emitters.size()));
        phn = D->emit(e, w, p, g);
        sdom = D->emitSdom(e);
        bdry = e.kind == MCB_EMIT_BDRY ? e.index : -1;
    }
    push(phn.pos);                                                                 /* TrkPhonon ctor */
    M->drawScatNext(phn, g);                                                       /* :254 */
    auto local = [&](int s, int b) {                                               /* find(sdom->bdryPtrs(), bdry) */
        if (b < 0) return -1;
        const std::vector<int>& pl = D->sdoms[s].planes;
        for (size_t k = 0; k < pl.size(); ++k) if (pl[k] == b) return (int)k;
        return -1;
    };
    for (long i = 0; i < t->maxloop; i++) {                                        /* :258 */
        g.begin(0, (uint32_t)(i + 1));
        const long k = o->nsteps++;
        if (k < o->max_steps) {
            o->step_sdom[k] = sdom; o->step_in[k] = local(sdom, bdry); o->step_in_kind[k] = bdry >= 0 ? D->planes[bdry].kind : -1;
            o->step_out[k] = -1; o->step_out_kind[k] = -1;
        }
        double vel = M->velAt(phn);
        bdry = D->advect(sdom, phn, vel);                                          /* :265 */
        push(phn.pos);                                                             /* Phonon::move -> TrkPhonon::pos */
        if (!phn.alive) { o->escaped = 1; break; }                                 /* :267-271 */
        if (k < o->max_steps) { o->step_out[k] = local(sdom, bdry); o->step_out_kind[k] = bdry >= 0 ? D->planes[bdry].kind : -1; }
        if (bdry >= 0) {                                                           /* :277-290 */
            const bool peri = D->planes[bdry].kind == MCB_BDRY_PERI;
            bdry = D->scatter(bdry, phn, g);
            if (peri) push(phn.pos);                                               /* PeriBoundary::scatter sets pos */
            if (bdry < 0) { o->escaped = 2; break; }
            sdom = D->planes[bdry].sdom;
        } else {
            M->scatter(phn, g);                                                    /* :291-294 */
        }
        if (!phn.alive || phn.nscat >= t->maxscat) break;                          /* :295 */
    }
    return MCB_OK;
}

int orc_cell_index(const orc_domain* d, int64_t n, const double* pos, const int32_t* sdom, int64_t* index) {
    for (int64_t i = 0; i < n; ++i) {
        long idx[3]; V3 p(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        d->coord2index(sdom[i], d->coord(sdom[i], p), idx);
        index[3 * i] = idx[0]; index[3 * i + 1] = idx[1]; index[3 * i + 2] = idx[2];
    }
    return MCB_OK;
}
int orc_accumulate(const orc_domain* d, int32_t rows, int64_t n, const int32_t* sdom, const double* bpos,
                   const double* epos, const double* amount, double* field) {
    std::vector<double> data(field, field + (size_t)rows * (size_t)d->cols);
    for (int64_t i = 0; i < n; ++i)
        accumulate(*d, data, rows, sdom[i], V3(bpos[3 * i], bpos[3 * i + 1], bpos[3 * i + 2]),
                   V3(epos[3 * i], epos[3 * i + 1], epos[3 * i + 2]), amount + (size_t)rows * (size_t)i);
    std::memcpy(field, data.data(), data.size() * sizeof(double));
    return MCB_OK;
}
void orc_philox_words(uint64_t seed, uint64_t particle, uint32_t event, uint32_t block, uint32_t out[4]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)particle, (uint32_t)(particle >> 32), event, block};
    philox4x32_10(ctr, key, out);
}
void orc_mt_draws(uint32_t seed, int which, int64_t m, int64_t n, double* out) {
    Words g(ORC_RNG_MT19937, seed);
    for (int64_t i = 0; i < n; ++i)
        out[i] = which == 0 ? uniform01(g) : (which == 1 ? uniformOne(g) : (double)uniformInt(g, (long)m));
}

} // extern "C"
