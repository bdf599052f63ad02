// oracle/shim/boost/random/uniform_01.hpp — TEST INFRASTRUCTURE.  Boost.Random's uniform_01<double> on a 32-bit integer
// engine: (x - min) * 1 / (max - min + 1), i.e. one word times 2^-32, redrawn if the product rounds to 1 (never for 2^-32).
#ifndef MCB_SHIM_BOOST_UNIFORM_01
#define MCB_SHIM_BOOST_UNIFORM_01
namespace boost { namespace random {
template <class Real = double> class uniform_01 {
public:
    typedef Real result_type;
    template <class Engine> result_type operator()(Engine& eng) const {
        const Real factor = Real(1) / (Real((eng.max)() - (eng.min)()) + Real(1));
        for (;;) {
            const Real r = Real(eng() - (eng.min)()) * factor;
            if (r < Real(1)) return r;
        }
    }
};
} using random::uniform_01; }
#endif
