// oracle/shim/boost/random/uniform_int_distribution.hpp — TEST INFRASTRUCTURE.  Boost's generate_uniform_int for a
// 32-bit engine and a range below 2^32: nothing is drawn for an empty range; the full range maps directly; otherwise
// bucketed rejection (bucket = brange / (range + 1), one more if the division is exact at the top).
#ifndef MCB_SHIM_BOOST_UNIFORM_INT
#define MCB_SHIM_BOOST_UNIFORM_INT
#include <cstdlib>
namespace boost { namespace random {
template <class Int = int> class uniform_int_distribution {
    Int min_, max_;
public:
    typedef Int result_type;
    uniform_int_distribution(Int mn = 0, Int mx = 9) : min_(mn), max_(mx) {}
    template <class Engine> result_type operator()(Engine& eng) const {
        typedef unsigned long long U;
        const U range = (U)max_ - (U)min_;
        const U brange = (U)(eng.max)() - (U)(eng.min)();
        if (range == 0) return min_;
        if (range > brange) std::abort();                     // multi-word composition: not needed by the reference
        if (range == brange) return (Int)((U)(eng() - (eng.min)()) + (U)min_);
        U bucket = brange / (range + 1);
        if (brange % (range + 1) == range) ++bucket;
        for (;;) {
            const U r = (U)(eng() - (eng.min)()) / bucket;
            if (r <= range) return (Int)(r + (U)min_);
        }
    }
};
} using random::uniform_int_distribution; }
#endif
