// oracle/shim/boost/optional.hpp — TEST INFRASTRUCTURE.  boost::optional as used by problem.h (empty / engaged, *, ->, bool).
#ifndef MCB_SHIM_BOOST_OPTIONAL
#define MCB_SHIM_BOOST_OPTIONAL
#include <optional>
namespace boost {
template <class T> class optional : public std::optional<T> {
public:
    optional() {}
    optional(const T& v) : std::optional<T>(v) {}
};
}
#endif
