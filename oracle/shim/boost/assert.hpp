// oracle/shim/boost/assert.hpp — TEST INFRASTRUCTURE.  Like Boost's: active unless NDEBUG / BOOST_DISABLE_ASSERTS; a violated
// contract prints the reference's own message and aborts.
#ifndef MCB_SHIM_BOOST_ASSERT
#define MCB_SHIM_BOOST_ASSERT
#include <cstdio>
#include <cstdlib>
#if defined(NDEBUG) || defined(BOOST_DISABLE_ASSERTS)
#define BOOST_ASSERT_MSG(expr, msg) ((void)0)
#define BOOST_ASSERT(expr) ((void)0)
#else
#define BOOST_ASSERT_MSG(expr, msg) \
    ((expr) ? (void)0 : (std::fprintf(stderr, "%s:%d: assertion failed: %s (%s)\n", __FILE__, __LINE__, msg, #expr), std::abort()))
#define BOOST_ASSERT(expr) BOOST_ASSERT_MSG(expr, "")
#endif
#endif
