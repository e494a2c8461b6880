/*
 * fabe13.h -- C API of the FP64 trigonometric library, B200 (sm_100a) build.
 *
 * This header declares the same entry points, with the same signatures and the same meaning, as the reference
 * library's public header (reference fabe13/fabe13.h:22 for FABE13_ALIGNMENT, :51-56 fabe13_sincos, :62 and :68 the
 * two getters, :71-78 the scalar functions), so a program written against the reference compiles and links against
 * libfabe13.so from this repository without source changes.  The array entry point runs on the GPU (see
 * fabe13_cuda.h for the device-pointer variants); the scalar entry points evaluate the same per-element algorithm
 * on the host.
 *
 * Behavioural contract kept from the reference (file:line in fabe13/fabe13.c):
 *   - n <= 0 is a silent no-op (:223); no NULL checks; buffers are caller-owned; nothing is retained.
 *   - NaN and +-Inf inputs give the canonical quiet NaN 0x7FF8000000000000 in BOTH outputs (:563-566, :599-600).
 *   - |x| <= 2^-27 gives sin = x bit-for-bit (so +-0 and subnormals pass through) and cos = 1.0 (:71, :567, :597-598).
 *   - the quadrant index k = rint(x * 2/pi) is formed by the reference's exact operation sequence (:568-571).
 * Deliberate differences (documented in DESIGN.md): results are accurate to ~2e-16 for every finite x up to
 * DBL_MAX (the reference degrades as 3.9e-17*|x| and returns garbage beyond ~1e15), and elements in the n%8 tail
 * or in calls with n < 32 get the same accuracy as the bulk (the reference's scalar core is off by up to 1e-1).
 * in == sin_out (in-place) is allowed.
 */
#ifndef FABE13_H
#define FABE13_H

#ifdef __cplusplus
extern "C" {
#endif

/* Alignment (bytes) the reference recommends for the three arrays.  Here 32-byte alignment is enough for the
 * 256-bit path; any 8-byte-aligned pointers are accepted and the widest common vector width is used. */
#define FABE13_ALIGNMENT 64

/* sin_out[i] = sin(in[i]), cos_out[i] = cos(in[i]) for i in [0, n).
 * The pointers may be host pointers (pageable or pinned; the call stages them through the GPU and returns when the
 * results are in sin_out/cos_out) or device pointers of the current CUDA device (the call enqueues on the legacy
 * default stream and synchronises it).  Errors are reported through fabe13_cuda_last_error(); there is no CPU
 * fallback. */
void fabe13_sincos(const double* in, double* sin_out, double* cos_out, int n);

/* Name of the active back-end, e.g. "CUDA sm_100a (NVIDIA B200, 148 SMs)". */
const char* fabe13_get_active_implementation_name(void);

/* Lanes that execute one instruction together: 32 (a warp). */
int fabe13_get_active_simd_width(void);

/* Scalar functions (host).  sin/cos use the per-element core of the array path; sinc/tan/cot are derived from it
 * with the reference's guards (|x| < 1e-9 -> 1; cos == 0 -> NaN; sin == 0 -> NaN) and one IEEE division -- the same
 * code the array forms in fabe13_cuda.h compile for the device; every NaN they return is the canonical quiet NaN
 * (the reference propagates whatever NaN the division makes).  atan/asin/acos forward to libm. */
double fabe13_sin(double x);
double fabe13_cos(double x);
double fabe13_sinc(double x);
double fabe13_tan(double x);
double fabe13_cot(double x);
double fabe13_atan(double x);
double fabe13_asin(double x);
double fabe13_acos(double x);

#ifdef __cplusplus
}
#endif

#endif /* FABE13_H */
