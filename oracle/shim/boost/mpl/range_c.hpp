// oracle/shim/boost/mpl/range_c.hpp — TEST INFRASTRUCTURE: forwards to the Fusion/MPL stand-in.
#include "../fusion/fusion_shim.hpp"
