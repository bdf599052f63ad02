// oracle/shim/boost/random/discrete_distribution.hpp — TEST INFRASTRUCTURE.  Boost's discrete_distribution: Walker alias
// table built by pairing below-average with above-average weights in input order; a draw is one uniform_int plus one
// uniform_01.
#ifndef MCB_SHIM_BOOST_DISCRETE
#define MCB_SHIM_BOOST_DISCRETE
#include <utility>
#include <vector>
#include "uniform_01.hpp"
#include "uniform_int_distribution.hpp"
namespace boost { namespace random {
template <class Int = int, class W = double> class discrete_distribution {
    std::vector<std::pair<W, Int> > table_;
public:
    typedef Int result_type;
    discrete_distribution() { table_.push_back(std::make_pair(W(1), Int(0))); }
    template <class It> discrete_distribution(It first, It last) {
        std::vector<std::pair<W, Int> > below, above;
        std::size_t size = 0; W sum = 0;
        for (It it = first; it != last; ++it) { sum += *it; ++size; }
        const W average = sum / W(size);
        Int i = 0;
        for (It it = first; it != last; ++it, ++i) {
            const W val = *it / average;
            if (val < W(1)) below.push_back(std::make_pair(val, i)); else above.push_back(std::make_pair(val, i));
        }
        table_.resize(size);
        typename std::vector<std::pair<W, Int> >::iterator b = below.begin(), be = below.end(), a = above.begin(), ae = above.end();
        while (b != be && a != ae) {
            table_[(std::size_t)b->second] = std::make_pair(b->first, a->second);
            a->first -= (W(1) - b->first);
            if (a->first < W(1)) *b = *a++; else ++b;
        }
        for (; b != be; ++b) table_[(std::size_t)b->second].first = W(1);
        for (; a != ae; ++a) table_[(std::size_t)a->second].first = W(1);
    }
    template <class Engine> result_type operator()(Engine& eng) const {
        const Int r = uniform_int_distribution<Int>(0, (Int)table_.size() - 1)(eng);
        const W test = uniform_01<W>()(eng);
        return test < table_[(std::size_t)r].first ? r : table_[(std::size_t)r].second;
    }
};
} using random::discrete_distribution; }
#endif
