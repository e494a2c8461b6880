// fabe13_hostcore.h -- host instantiation of the shared per-element core (fabe13_hostcore.cpp); see there.
#ifndef FABE13_HOSTCORE_H
#define FABE13_HOSTCORE_H

#include <stddef.h>

namespace fabe13 {

// one element (the scalar API of fabe13.h; k may be null)
void host_sincos_one(double x, int core, double* s, double* c, long long* k);
void host_sincosf_one(float x, int core, float* s, float* c);
double host_derived_one(double x, int op, int core);  // op = FABE13_OP_TAN / _COT / _SINC

// arrays: either output may be null; op = 0 (sincos) or a derived function written to sin_out; in == sin_out allowed.
// `threads` > 1 splits large arrays (>= 65536 elements per thread) into contiguous pieces.  stream_stores: 1 = write the
// results with non-temporal stores, 0 = ordinary stores, -1 = by size (non-temporal from 2^20 elements on).
void host_sincos_array(const double* in, double* sin_out, double* cos_out, size_t n, int op, int core, int threads,
                       int stream_stores = -1);
void host_sincosf_array(const float* in, float* sin_out, float* cos_out, size_t n, int core, int threads);

// test hook: the large-argument reduction (2^45 <= |x| < Inf) as the product runs it (r, q), its exact form alone, and
// whether the fast form vouched for its own result
void host_reduce_huge(const double* in, double* r, int* q, double* r_exact, int* q_exact, int* fast_ok, size_t n);

const char* host_isa_name();  // "AVX-512", "AVX2" or "scalar": the clone of the array loop this CPU runs

}  // namespace fabe13

#endif  // FABE13_HOSTCORE_H
