// oracle/shim/boost/fusion/algorithm/transformation.hpp — TEST INFRASTRUCTURE: forwards to the Fusion/MPL stand-in.
#include "../../fusion/fusion_shim.hpp"
