// field.h — host mirror of Field (field.h:26-46, field.cpp:25-90) and Statistics<S> (field.h:48-66, field.cpp:226-264, the
// Welford accumulator the driver uses across nsim repetitions).  Field keeps the reference's column layout (one column per
// cell, subdomains in Domain::sdomPtrs() order, cells i-fastest) and its functor constructor (CellVolF, OctetDomain::WeightF).
// Field::accumulate (field.cpp:92-220) is on the hot path: the mirror's method runs the segment through the device's tally
// walk (mcb_accumulate) -- there is no host restatement of it in the product.
#ifndef MCB_HOST_FIELD_H
#define MCB_HOST_FIELD_H
#include <map>
#include "domain.h"
#include "mc_types.h"

class Field {
    const Domain* dom_;
    std::map<const Subdomain*, long> offset_;      // first column of a subdomain (-1: no columns), field.cpp:25-45
    ArrayXXd data_;
    void init(long rows) {
        long n = 0;
        for (const Subdomain* s : dom_->sdomPtrs()) { const long cells = s->shape().prod(); offset_[s] = cells > 0 ? n : -1; n += cells; }
        data_ = ArrayXXd(rows, n);
    }
public:
    Field() : dom_(0) {}
    Field(long rows, const Domain* dom) : dom_(dom) { init(rows); }
    // every column from fun(subdomain, cell index): field.cpp:53-80
    template <typename F>
    Field(long rows, const Domain* dom, const F& fun) : dom_(dom) {
        init(rows);
        long n = 0;
        for (const Subdomain* s : dom_->sdomPtrs()) {
            const Vector3l shape = s->shape();
            for (long k = 0; k < shape(2); ++k) for (long j = 0; j < shape(1); ++j) for (long i = 0; i < shape(0); ++i) {
                const VectorXd v = fun(s, Vector3l(i, j, k));
                MC_ASSERT_MSG((long)v.size() == rows, "Field functor: wrong number of rows");
                for (long r = 0; r < rows; ++r) data_(r, n) = v[(size_t)r];
                n++;
            }
        }
    }
    const ArrayXXd& data() const { return data_; }
    // path-length weighted deposit of `amount` over the cells crossed by ipos -> fpos inside sdom (device walk)
    Field& accumulate(const Subdomain* sdom, const Vector3d& ipos, const Vector3d& fpos, const VectorXd& amount);
};

template <typename S>
class Statistics {
    long n_;
    S z_, m_, s_;
public:
    Statistics() : n_(0) {}
    Statistics(const S& zero) : n_(0), z_(zero), m_(zero), s_(zero) {}
    S mean() const { return m_; }
    S variance() const { return n_ < 2 ? z_ : s_ / (double)(n_ - 1); }
    void add(const S& x) {
        n_++;
        S d = x - m_;
        m_ += d / (double)n_;
        s_ += d * (x - m_);
    }
};
#endif
