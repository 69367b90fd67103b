// FMA-contracted build of the path tracer (nvcc default -fmad=true, what the reference's own
// CUDA build uses).  See trace_impl.cuh.
#define SRT_FP_NS fastfp
#include "trace_impl.cuh"
