// constants.h — the reference's literal constants (constants.h:15-29).  HBAR is NOT the CODATA value;
// it is kept digit for digit because material tables derived from it must match.
#ifndef MCB_HOST_CONSTANTS_H
#define MCB_HOST_CONSTANTS_H
#include <algorithm>
#include <cmath>
#include <limits>

typedef std::numeric_limits<double> Dbl;

const double PI   = 3.141592653589793;
const double HBAR = 1.054560652927e-034;
const double KB   = 1.380648e-023;

template <typename Scalar>
bool isApprox(const Scalar& a, const Scalar& b) {            // constants.h:21-29
    typedef std::numeric_limits<Scalar> Lim;
    Scalar tol = 100 * Lim::epsilon();
    Scalar mx = std::max(std::abs(a), std::abs(b));
    return std::abs(a - b) <= std::max(Lim::min(), tol * mx);
}
#endif
