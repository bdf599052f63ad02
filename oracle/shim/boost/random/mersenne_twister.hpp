// oracle/shim/boost/random/mersenne_twister.hpp — TEST INFRASTRUCTURE.  boost::random::mt19937 and std::mt19937 are the
// same generator (same parameters, same 32-bit seeding recurrence), so the standard one is used.
#ifndef MCB_SHIM_BOOST_MT
#define MCB_SHIM_BOOST_MT
#include <random>
namespace boost { namespace random { typedef std::mt19937 mt19937; } using random::mt19937; }
#endif
