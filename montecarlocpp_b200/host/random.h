// random.h — host mirror of random.h:21-26.  On the accelerated path the host engine is only used to
// derive the 64-bit Philox seed of a solve (FieldProblem::solve); all per-phonon draws happen on the
// device from a counter-based Philox4x32-10 stream keyed by (seed, particle id).
#ifndef MCB_HOST_RANDOM_H
#define MCB_HOST_RANDOM_H
#include "mc_types.h"
typedef std::random_device Dev;
#endif
