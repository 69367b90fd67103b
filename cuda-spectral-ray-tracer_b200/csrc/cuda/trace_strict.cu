// Strict build of the path tracer: compiled with -fmad=false so every float operation rounds
// separately, exactly like the host oracle (g++ -ffp-contract=off).  See trace_impl.cuh.
#define SRT_FP_NS strictfp_
#include "trace_impl.cuh"
