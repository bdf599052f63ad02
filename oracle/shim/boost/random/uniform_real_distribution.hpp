// oracle/shim/boost/random/uniform_real_distribution.hpp — TEST INFRASTRUCTURE.  Boost's generate_uniform_real for an
// integer engine: numerator / divisor * (max - min) + min with divisor = engine range + 1, redrawn unless < max.
#ifndef MCB_SHIM_BOOST_UNIFORM_REAL
#define MCB_SHIM_BOOST_UNIFORM_REAL
namespace boost { namespace random {
template <class Real = double> class uniform_real_distribution {
    Real min_, max_;
public:
    typedef Real result_type;
    uniform_real_distribution(Real mn = Real(0), Real mx = Real(1)) : min_(mn), max_(mx) {}
    template <class Engine> result_type operator()(Engine& eng) const {
        const Real divisor = Real((eng.max)() - (eng.min)()) + Real(1);
        for (;;) {
            const Real numerator = Real(eng() - (eng.min)());
            const Real r = numerator / divisor * (max_ - min_) + min_;
            if (r < max_) return r;
        }
    }
};
} using random::uniform_real_distribution; }
#endif
