// oracle/shim/boost/fusion/container/vector/vector50.hpp — TEST INFRASTRUCTURE: forwards to the Fusion/MPL stand-in.
#include "../../../fusion/fusion_shim.hpp"
