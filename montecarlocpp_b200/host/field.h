// field.h — host mirror of Statistics<S> (field.h:48-66, field.cpp:226-264), the Welford accumulator the
// driver uses across nsim repetitions.  Field::accumulate itself (field.cpp:92-220) runs on the device.
#ifndef MCB_HOST_FIELD_H
#define MCB_HOST_FIELD_H
#include "mc_types.h"

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
