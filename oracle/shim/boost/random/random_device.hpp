// oracle/shim/boost/random/random_device.hpp — TEST INFRASTRUCTURE.  Reads /dev/urandom like Boost's; when the
// environment variable MCREF_SEED is set the device instead returns MCREF_SEED, MCREF_SEED+1, ... (one per call), which
// makes a run of the reference binary reproducible without its -DDEBUG build (that build also switches to TrkPhonon).
#ifndef MCB_SHIM_BOOST_RANDOM_DEVICE
#define MCB_SHIM_BOOST_RANDOM_DEVICE
#include <atomic>
#include <cstdio>
#include <cstdlib>
namespace boost { namespace random {
class random_device {
public:
    typedef unsigned int result_type;
    random_device() {}
    result_type operator()() {
        static std::atomic<unsigned int> next(0);
        if (const char* s = std::getenv("MCREF_SEED")) return (result_type)std::strtoul(s, 0, 10) + next.fetch_add(1u);
        result_type r = 0;
        std::FILE* f = std::fopen("/dev/urandom", "rb");
        if (!f || std::fread(&r, sizeof r, 1, f) != 1) std::abort();
        std::fclose(f);
        return r;
    }
};
} using random::random_device; }
#endif
