/*
 * fabe13_cuda.h -- additive CUDA entry points of libfabe13.so (B200 / sm_100a).
 *
 * The reference has no device API; these are the functions a binding for the array path would call when the data
 * already lives in GPU memory.  Each one replaces a call to the reference's fabe13_sincos (fabe13/fabe13.h:51-56,
 * implementation fabe13/fabe13.c:222-259) for device-resident arrays.  Plain C ABI: pointers, sizes, an opaque
 * stream handle (a cudaStream_t passed as void*).  All functions returning int return 0 on success or a
 * cudaError_t value; the last non-zero value is also kept for fabe13_cuda_last_error().
 */
#ifndef FABE13_CUDA_H
#define FABE13_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-element core selection (option "core"). */
#define FABE13_CUDA_CORE_POLY 0 /* direct minimax kernels: 24 FP64 instructions per element (default) */
#define FABE13_CUDA_CORE_PSI 1  /* Psi-Hyperbasis rational core + correction polynomials */

/* Asynchronous sincos over n device-resident doubles on `cuda_stream` (NULL = legacy default stream).
 * d_in, d_sin, d_cos are device pointers on the current device, 8-byte aligned; d_in == d_sin is allowed.
 * The call is ONE kernel launch (per 2^30 elements) and needs no scratch memory: arguments >= 2^45 are reduced
 * inside the streaming kernel (by their own lanes when a warp holds one or two, from per-warp lists in shared memory
 * otherwise, whole warps at once where they are dense) with a 31 KB read-only table of the powers of two modulo 2 pi
 * followed by Cody-Waite, correctly rounded to half an ulp; the rare remainder that comes out within 2^-20 of 0 or within
 * 2^-18 of pi/4 is recomputed by an exact 192-bit Payne-Hanek product.  Option "hybrid_min_n" (default: never) re-enables, from that many elements on, the launch
 * sequence with a global worklist for sparse huge arguments (n/32 32-bit entries from a library-owned stream-ordered
 * memory pool) and a second-pass kernel. */
int fabe13_sincos_device(const double* d_in, double* d_sin, double* d_cos, size_t n, void* cuda_stream);

/* Same, and also writes the quadrant index per element (test hook for the k-parity contract):
 * d_k[i] = rint(x*2/pi) by the reference's sequence for |x| < 2^45; k mod 4 (0..3) for 2^45 <= |x| < Inf;
 * 0 for NaN/Inf. */
int fabe13_sincos_device_k(const double* d_in, double* d_sin, double* d_cos, long long* d_k, size_t n,
                           void* cuda_stream);

/* Array forms of fabe13_sin / fabe13_cos (reference scalar API fabe13/fabe13.h:71-72): one output array, 16 bytes of
 * traffic per element instead of 24.  Same numerics as fabe13_sincos (bit-identical values).  Device pointers on a
 * stream, or host pointers (synchronous), exactly as the sincos entries. */
int fabe13_sin_device(const double* d_in, double* d_out, size_t n, void* cuda_stream);
int fabe13_cos_device(const double* d_in, double* d_out, size_t n, void* cuda_stream);
int fabe13_sin_host(const double* in, double* out, size_t n);
int fabe13_cos_host(const double* in, double* out, size_t n);

/* Array forms of fabe13_tan / fabe13_cot / fabe13_sinc (reference scalar API fabe13/fabe13.h:73-75, implementation
 * fabe13/fabe13.c:264-266): tan = sin/cos (NaN if cos == 0), cot = cos/sin (NaN if sin == 0), sinc = 1 for
 * |x| < 1e-9 and sin(x)/x otherwise -- one correctly rounded division after the sincos core, NaN results canonical.
 * Element i equals the scalar entry point's value bit for bit.  One output array, 16 bytes of traffic per element. */
int fabe13_tan_device(const double* d_in, double* d_out, size_t n, void* cuda_stream);
int fabe13_cot_device(const double* d_in, double* d_out, size_t n, void* cuda_stream);
int fabe13_sinc_device(const double* d_in, double* d_out, size_t n, void* cuda_stream);
int fabe13_tan_host(const double* in, double* out, size_t n);
int fabe13_cot_host(const double* in, double* out, size_t n);
int fabe13_sinc_host(const double* in, double* out, size_t n);

/* FP32 arrays (the reference lists an FP32 path on its roadmap only, README.md:225; there is no reference
 * implementation to be in parity with).  Each element is widened to FP64, reduced and evaluated in FP64 (every
 * range, Payne-Hanek included; plain arguments by a route with shorter polynomials than the FP64 entries need) and
 * both results are rounded once to float: within 0.5003 ulp of the true values (0.500185 measured over all 2^32
 * inputs), the correctly rounded float in all but 6e-6 of the cases.  NaN/Inf -> 0x7FC00000 in both outputs; |x| <= 2^-27 -> (x, 1.0f).  12 bytes of traffic per
 * element.  Arrays must be 4-byte aligned; 256-bit accesses are used when the three arrays share one phase modulo
 * 32 bytes.  The host entry takes device, pinned/registered (processed in place over PCIe) or pageable memory
 * (staged in 2M-element pieces by the same worker threads as the FP64 path). */
int fabe13_sincosf_device(const float* d_in, float* d_sin, float* d_cos, size_t n, void* cuda_stream);
int fabe13_sincosf_host(const float* in, float* sin_out, float* cos_out, size_t n);

/* Same as fabe13_sincos_device with caller-provided scratch of at least fabe13_sincos_workspace_bytes(n) bytes
 * (no allocation on the call path; usable inside CUDA graph capture on any toolkit).  The size may be 0 (small n,
 * or option "worklist" = 2); a NULL workspace is accepted exactly then. */
size_t fabe13_sincos_workspace_bytes(size_t n);
int fabe13_sincos_device_ws(const double* d_in, double* d_sin, double* d_cos, size_t n, void* workspace,
                            size_t workspace_bytes, void* cuda_stream);

/* One slice per GPU, no exchange between them: launches on every listed device, then waits for all.
 * d_in[g], d_sin[g], d_cos[g] live on devices[g] and hold n_per_dev[g] elements. */
int fabe13_sincos_multi(int ngpu, const int* devices, const double* const* d_in, double* const* d_sin,
                        double* const* d_cos, const size_t* n_per_dev);

/* Host-pointer entry with a size_t length and an error return; fabe13_sincos() forwards here.
 * Pinned host memory (cudaHostAlloc / cudaHostRegister / fabe13_cuda_host_alloc) is copied directly by the copy
 * engines, H2D and D2H overlapped with the kernel chunk by chunk; pageable memory goes through pinned staging
 * buffers filled by worker threads.  Pinned arrays of at most "zero_copy_max" elements (default 2^24) skip the copy
 * engines: the kernels read the host buffer and write the results back directly over PCIe (as do pageable calls of
 * at most 32768 elements, through a bounce buffer).  With option "host_devices" > 1 the array is split into contiguous slices,
 * one per GPU. */
int fabe13_sincos_host(const double* in, double* sin_out, double* cos_out, size_t n);

/* Which host-pointer calls run where.  Below option "min_n" elements (default 16384; FABE13_CUDA_MIN_N) -- below
 * "min_n_pageable" (default 262144) if any of the arrays is pageable memory -- a host-pointer call (fabe13_sincos and
 * every *_host entry) is evaluated on the calling thread by the HOST instantiation of the same per-element core the
 * kernels compile for the device (bit-identical results): a kernel launch plus a synchronisation costs ~13 us, the
 * host core ~1.3 ns per element, and pageable arrays pay two more host copies and a PCIe round trip on top.  Both
 * defaults are measured crossovers (profiles/r02_abi_latency.txt).  This mirrors the reference's own small-n route
 * (fabe13/fabe13.c:226-229: n < 32 -> scalar loop).  Setting both to 0 sends every call to the GPU.
 * On a CUDA failure every int entry point returns the cudaError_t and leaves the outputs unspecified; the void
 * fabe13_sincos(), which has no error channel (fabe13/fabe13.h:51-56), then fills sin_out / cos_out on the host core
 * (option "fallback" = 1, the default; "host_threads" threads) and the error stays readable through
 * fabe13_cuda_last_error().  (An in-place call -- in == sin_out -- whose pipeline failed half way has already
 * overwritten part of its input; the completion is exact only for calls with separate arrays.)
 * Read-only counters: "host_core_calls", "fallbacks".
 * Option "host_assist" = k (default 0: off; FABE13_CUDA_HOST_ASSIST) lets k host threads work on a LARGE page-locked
 * array (at least "host_assist_min_n" elements, default 2^24) together with the GPU: the copy-engine pipeline takes
 * chunks from the front of the array, the host threads run the host core on pieces from its back, and both claim their
 * next piece from one shared cursor, so the split follows their speeds during the call.  A host-resident array is
 * bound by PCIe (78 GB/s for this 8 B in : 16 B out mix = 3.25 Gelem/s per GPU) and by the host's memory, not by the
 * kernel; the host cores add what the host's memory still carries on top of the link.  The bits do not depend on
 * where the two sides meet.  Read-only counters "assist_gpu_elems" / "assist_host_elems" tell how such calls were split. */

/* Pinned host allocations for the fast host path. */
void* fabe13_cuda_host_alloc(size_t bytes);
void fabe13_cuda_host_free(void* p);
/* Opt-in for buffers the caller already owns (malloc, mmap, a language runtime's arrays): page-lock and map them once,
 * and every later fabe13_sincos / _host call on them takes the pinned path (copy engines or zero-copy kernels reading
 * and writing the caller's memory directly, no staging threads).  Registering costs about as much as one staged call
 * over the same bytes, so it pays for buffers that are reused.  Thin wrappers over cudaHostRegister/Unregister. */
int fabe13_cuda_host_register(void* p, size_t bytes);
int fabe13_cuda_host_unregister(void* p);

/* Give back what the library holds between calls: the host pipeline's device buffers (up to 8 slots x 3 x chunk), its
 * pinned staging buffers, the small-call bounce buffer and the cached scratch of the stream-ordered pool (which keeps
 * at most 256 MB on its own).  _trim keeps streams and events, so the next call only re-allocates buffers; _shutdown
 * also destroys streams, events and the pool (the library re-initialises on the next call).  Both wait for host-pointer
 * calls in flight; do not call them concurrently with device-pointer calls on other threads. */
int fabe13_cuda_trim(void);
int fabe13_cuda_shutdown(void);

/* Contiguous slice [begin, end) of an n-element array owned by `rank` of `world` ranks; interior boundaries are
 * multiples of 4 elements so every slice keeps 32-byte alignment.  The remainder goes to the last rank. */
void fabe13_shard_bounds(size_t n, int world, int rank, size_t* begin, size_t* end);

/* Options: "core" (0/1), "unroll" (1/2), "blocks_per_sm" (0 = one tile per block, k = persistent grid of k blocks
 * per SM), "tiles_per_block", "small_n" (arrays below this size take the one-element-per-thread kernel; default 0 = never),
 * "worklist" (where arguments >= 2^45 are reduced: 0 = by size [default]: inside the streaming kernel, one launch,
 * below "hybrid_min_n" elements [default 2^31: always], plus a global list for sparse stragglers from there on; 1 = global list
 * and second-pass kernel only [the round-1 design]; 2 = shared memory only at every size; 3 = hybrid at every size),
 * "hybrid_min_n", "local_min" / "dense_hint" / "direct_max" (routing thresholds of the streaming kernel),
 * "worklist_shift" (worklist=1: capacity n >> shift), "second_pass_blocks_per_sm",
 * "pdl" (programmatic dependent launch of the kernels of one call), "chunk_elems", "host_devices", "stage_threads",
 * "zero_copy_max", "stream_copy" (non-temporal stores for the staging copies of pageable arrays and for the host
 * core's results on arrays of 2^20 elements or more: 1 [default] / 0),
 * "min_n", "min_n_pageable", "fallback", "host_threads", "host_assist", "host_assist_min_n" (see above).
 * Environment variables FABE13_CUDA_<OPTION IN CAPS> set the initial values.
 * Read-only keys for get: "launches" (kernels launched so far), "host_core_calls", "fallbacks", "assist_gpu_elems",
 * "assist_host_elems", "device_count", "sm_count". */
int fabe13_cuda_set_option(const char* key, long long value);
int fabe13_cuda_get_option(const char* key, long long* value);

int fabe13_cuda_device_count(void);
/* last non-zero cudaError_t this library saw ON THE CALLING THREAD (thread-local: clear / call / check needs no lock;
 * errors of the library's worker threads are handed to the thread that made the call), 0 if none */
int fabe13_cuda_last_error(void);
const char* fabe13_cuda_last_error_string(void);
void fabe13_cuda_clear_error(void);

/* Test hook, never on a product path: evaluates the shared per-element core on the HOST for n elements
 * (same code the kernels compile for the device).  Lets CPU-only CI check the algorithm without a GPU. */
void fabe13_debug_host_core(const double* in, double* sin_out, double* cos_out, long long* k_out, size_t n, int core);
void fabe13_debug_host_sincosf(const float* in, float* sin_out, float* cos_out, size_t n, int core);
/* Same for the derived functions: op = 3 (tan), 4 (cot), 5 (sinc). */
void fabe13_debug_host_derived(const double* in, double* out, size_t n, int op, int core);
/* The host ARRAY loop exactly as the small-n route and the fallback run it (vectorised plain path + patched lanes;
 * op = 0 for sincos; either output may be NULL), so CPU-only CI can compare it with the per-element routine. */
void fabe13_debug_host_array(const double* in, double* sin_out, double* cos_out, size_t n, int op, int core, int threads);
/* ... with the way the results are stored spelled out: stream_stores = 1 non-temporal stores (what arrays of 2^20
 * elements or more get), 0 ordinary stores, -1 by size. */
void fabe13_debug_host_array_stores(const double* in, double* sin_out, double* cos_out, size_t n, int op, int core, int threads,
                                    int stream_stores);

/* The reduction of large arguments (2^45 <= |x| < Inf; any other in[i] gets r = NaN, q = -1, fast_ok = 0) on the host: r[i] and
 * q[i] = rint(x 2/pi) mod 4 as the product computes them (the fast form: tabulated remainders of the powers of two modulo
 * 2 pi, then Cody-Waite; the exact 192-bit form where the fast one hands over), r_exact / q_exact from the exact form
 * alone, fast_ok[i] = 1 where the fast form vouched for its own result. */
void fabe13_debug_reduce_huge(const double* in, double* r, int* q, double* r_exact, int* q_exact, int* fast_ok, size_t n);

/* Bench/test hooks: fill a device array with the deterministic synthetic inputs used by bench.py and the parity
 * tests (counter-based splitmix64; the same generator exists on the host in oracle/fabe13_oracle.c).
 * uniform: x_i = fma(hi-lo, u_i, lo), u_i in [0,1);  loguniform: |x| log-uniform over [1, 2^1024), random sign. */
int fabe13_debug_fill_uniform_device(double* d_out, size_t n, unsigned long long seed, unsigned long long first_index,
                                     double lo, double hi, void* cuda_stream);
int fabe13_debug_fill_loguniform_device(double* d_out, size_t n, unsigned long long seed,
                                        unsigned long long first_index, void* cuda_stream);

#ifdef __cplusplus
}
#endif

#endif /* FABE13_CUDA_H */
