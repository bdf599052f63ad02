// oracle/shim/boost/fusion/sequence/intrinsic.hpp — TEST INFRASTRUCTURE: forwards to the Fusion/MPL stand-in.
#include "../../fusion/fusion_shim.hpp"
