// oracle/shim/boost/static_assert.hpp — TEST INFRASTRUCTURE.
#ifndef MCB_SHIM_BOOST_STATIC_ASSERT
#define MCB_SHIM_BOOST_STATIC_ASSERT
#define BOOST_STATIC_ASSERT_MSG(expr, msg) static_assert(expr, msg)
#define BOOST_STATIC_ASSERT(expr) static_assert(expr, #expr)
#endif
