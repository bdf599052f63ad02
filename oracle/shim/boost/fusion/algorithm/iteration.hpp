// oracle/shim/boost/fusion/algorithm/iteration.hpp — TEST INFRASTRUCTURE: forwards to the Fusion/MPL stand-in.
#include "../../fusion/fusion_shim.hpp"
