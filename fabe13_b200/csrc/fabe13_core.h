// fabe13_core.h -- the per-element FP64 sincos algorithm, written once for device and host.
//
// The CUDA kernels (fabe13_kernels.cu) instantiate these functions per lane; the scalar entry points of the
// C API (fabe13_sin/fabe13_cos/..., reference fabe13/fabe13.c:262-266) instantiate the very same code on the
// host, so a scalar call and an array element give bit-identical results.
//
// Stages (reference line numbers are for fabe13/fabe13.c, AVX-512 lane algorithm :561-600):
//   1. quadrant index  k = rint(x*2/pi) by the reference's exact op sequence (:568-571)   -> bit-exact k
//   2. Cody-Waite      r = x - k*pi/2 with a correct 3-word pi/2 and FMAs (replaces :572-576, whose low word
//                      is 2x too large -- SURVEY finding 3)                                -> |r - exact| <= 2^-54
//      Payne-Hanek     for |x| >= 2^45: 256-bit window of 2/pi times the 53-bit mantissa (the reference has no
//                      such path; its "Payne-Hanek" is the 2-term product above)
//   3. core            direct minimax kernels (replaces the Estrin degree-7 pair :383-389, :580-583) or the
//                      Psi-Hyperbasis rational core (:311-330) plus the correction polynomials it lacks
//   4. fix-up          quadrant rotation (:586-596), |x| <= 2^-27 -> (x, 1) (:567, :597-598), NaN/Inf -> canonical
//                      quiet NaN in both outputs (:563-566, :599-600); all on integer bit patterns
//
// All FP64 operations are spelled as explicit round-to-nearest mul/add/fma so neither nvcc (-fmad) nor gcc
// (-ffp-contract) can re-associate or fuse them differently on the two sides.
#ifndef FABE13_CORE_H
#define FABE13_CORE_H

#include <stdint.h>
#include <string.h>

#include "fabe13_constants.h"

#if defined(__CUDACC__)
#define FABE13_HD __host__ __device__ __forceinline__
#else
#define FABE13_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define FABE13_MUL(a, b) __dmul_rn((a), (b))
#define FABE13_ADD(a, b) __dadd_rn((a), (b))
#define FABE13_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
// host: the translation unit is compiled with -ffp-contract=off, so * and + stay separate roundings
#define FABE13_MUL(a, b) ((a) * (b))
#define FABE13_ADD(a, b) ((a) + (b))
#define FABE13_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

enum { FABE13_CORE_POLY = 0, FABE13_CORE_PSI = 1 };

// The FP64 constants of the hot path.  On the device they live in the constant bank, so every DFMA takes its
// coefficient as a c[bank][offset] operand: no per-thread materialisation of 64-bit immediates (which costs ~50
// UMOV/MOV per thread -- a lot when a thread handles only four elements).  On the host they are a static table.
struct Fabe13Coef {
    double k_hi, k_lo, magic, neg_magic;
    double pio2_1, pio2_2, pio2_3;
    double s1, s2, s3, s4, s5, s6;
    double c1, c2, c3, c4, c5, c6;
    double neg_half, one;
};
#define FABE13_COEF_INIT                                                                                             \
    {                                                                                                                \
        FABE13_K_TWO_OVER_PI_HI, FABE13_K_TWO_OVER_PI_LO, FABE13_ROUND_MAGIC, -FABE13_ROUND_MAGIC, -FABE13_PIO2_1,     \
            -FABE13_PIO2_2, -FABE13_PIO2_3, FABE13_S1, FABE13_S2, FABE13_S3, FABE13_S4, FABE13_S5, FABE13_S6, FABE13_C1, \
            FABE13_C2, FABE13_C3, FABE13_C4, FABE13_C5, FABE13_C6, -0.5, 1.0                                           \
    }
// the Psi-Hyperbasis core's constants (reference cubics with their signs folded in, fitted correction polynomials)
struct Fabe13PsiCoef {
    double three_eighths, one;
    double a2, neg_a3, neg_a1, b2, neg_b3, neg_b1;
    double ds[FABE13_PSI_CORR_TERMS], dc[FABE13_PSI_CORR_TERMS];
};
#define FABE13_PSI_COEF_INIT                                                                                     \
    {                                                                                                            \
        0.375, 1.0, FABE13_HX_A2, -FABE13_HX_A3, -FABE13_HX_A1, FABE13_HX_B2, -FABE13_HX_B3, -FABE13_HX_B1, FABE13_PSI_DS, \
            FABE13_PSI_DC                                                                                        \
    }
#if defined(__CUDACC__)
static __constant__ Fabe13Coef fabe13_coef_device = FABE13_COEF_INIT;
static __constant__ Fabe13PsiCoef fabe13_psi_coef_device = FABE13_PSI_COEF_INIT;
#endif
static const Fabe13Coef fabe13_coef_host = FABE13_COEF_INIT;
static const Fabe13PsiCoef fabe13_psi_coef_host = FABE13_PSI_COEF_INIT;
#if defined(__CUDA_ARCH__)
#define FABE13_CF(field) (fabe13_coef_device.field)
#define FABE13_PCF(field) (fabe13_psi_coef_device.field)
#else
#define FABE13_CF(field) (fabe13_coef_host.field)
#define FABE13_PCF(field) (fabe13_psi_coef_host.field)
#endif

#define FABE13_QNAN_BITS 0x7FF8000000000000ull
#define FABE13_ONE_BITS 0x3FF0000000000000ull
#define FABE13_ABS_MASK 0x7FFFFFFFFFFFFFFFull
#define FABE13_INF_BITS 0x7FF0000000000000ull
#define FABE13_TINY_BITS 0x3E40000000000000ull /* 2^-27, inclusive (reference :71, _CMP_LE_OS :567) */
#define FABE13_HUGE_BITS 0x42C0000000000000ull /* 2^45: first |x| on the Payne-Hanek path */

FABE13_HD uint64_t fabe13_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, sizeof u);
    return u;
#endif
}

FABE13_HD double fabe13_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, sizeof x);
    return x;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// 1. quadrant index, reference op sequence (fabe13/fabe13.c:568-571; same in :494-497, :426-429, :299-301)
//    returns k as a double; *k_int receives k as an integer (valid for |k| < 2^51, i.e. far beyond the 2^45 cut)
// ---------------------------------------------------------------------------------------------------------------
FABE13_HD double fabe13_quadrant_index(double x, double* biased_out) {
    const double p_hi = FABE13_MUL(x, FABE13_CF(k_hi));
    const double e1 = FABE13_FMA(x, FABE13_CF(k_hi), -p_hi);
    const double p_lo = FABE13_FMA(x, FABE13_CF(k_lo), e1);
    const double sum = FABE13_ADD(p_hi, p_lo);
    // rint() by the 1.5*2^52 bias: the biased sum carries k in its low mantissa bits (two's complement)
    const double biased = FABE13_ADD(sum, FABE13_CF(magic));
    *biased_out = biased;
    return FABE13_ADD(biased, FABE13_CF(neg_magic));
}

// k as an integer, and q = k mod 4 from the low word only (32-bit work on the device)
FABE13_HD int64_t fabe13_k_from_biased(double biased) { return (int64_t)(fabe13_bits(biased) - 0x4338000000000000ull); }
FABE13_HD uint32_t fabe13_q_from_biased(double biased) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(biased) & 3u;
#else
    return (uint32_t)fabe13_bits(biased) & 3u;
#endif
}

// 2a. Cody-Waite with a correct three-word pi/2.  x - k*P1 is exact (a multiple of 2^-52 below 2), the two
//     following FMAs each round once.
FABE13_HD double fabe13_cody_waite(double x, double k) {
    double r = FABE13_FMA(k, FABE13_CF(pio2_1), x);  // the table holds the negated words
    r = FABE13_FMA(k, FABE13_CF(pio2_2), r);
    r = FABE13_FMA(k, FABE13_CF(pio2_3), r);
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// 2b. Payne-Hanek for |x| >= 2^45 (works for any normal |x| >= 2^-10).
//     |x| = m * 2^e with the 53-bit integer mantissa m.  2/pi = sum_j W[j] 2^(-64 j).  Words whose product with
//     m*2^e is a multiple of 4 cannot change (k mod 4, fraction) and are skipped; the next four words give the
//     integer bits mod 4 and >= 126 fraction bits (the closest a double gets to a multiple of pi/2 costs ~62).
//     Returns q = rint(|x|*2/pi) mod 4 and r = |x| - rint(|x|*2/pi)*pi/2, |r| <= pi/4, to < 1 ulp.
// ---------------------------------------------------------------------------------------------------------------
FABE13_HD void fabe13_mul_64x64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    *lo = (uint64_t)p;
    *hi = (uint64_t)(p >> 64);
#endif
}

FABE13_HD int fabe13_clz64(uint64_t v) {  // v != 0
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

FABE13_HD double fabe13_u64_to_double(uint64_t v) {  // exact for v < 2^53
#if defined(__CUDA_ARCH__)
    return __ull2double_rn(v);
#else
    return (double)v;
#endif
}

FABE13_HD double fabe13_payne_hanek(uint64_t abs_bits, const uint64_t* W, int* q_out) {
    const int e = (int)(abs_bits >> 52) - 1075;  // |x| = m * 2^e
    const uint64_t m = (abs_bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
    const int j0 = ((e - 2) >> 6) + 1;  // first word whose product is not a multiple of 4
    const int s0 = e - 64 * j0;         // scale of that word: in [-62, 1]
    const uint64_t w0 = W[j0], w1 = W[j0 + 1], w2 = W[j0 + 2], w3 = W[j0 + 3];

    // low 256 bits of m * (w0:w1:w2:w3), top three words only (the lowest contributes just its carry)
    uint64_t hi3, lo3, hi2, lo2, hi1, lo1, hi0, lo0;
    fabe13_mul_64x64(m, w3, &hi3, &lo3);
    fabe13_mul_64x64(m, w2, &hi2, &lo2);
    fabe13_mul_64x64(m, w1, &hi1, &lo1);
    fabe13_mul_64x64(m, w0, &hi0, &lo0);
    (void)lo3;
    (void)hi0;
    const uint64_t R1 = lo2 + hi3;
    const uint64_t c1 = (R1 < lo2) ? 1u : 0u;
    const uint64_t R2 = lo1 + (hi2 + c1);  // hi2 < 2^53: no overflow in the inner sum
    const uint64_t c2 = (R2 < lo1) ? 1u : 0u;
    const uint64_t R3 = lo0 + (hi1 + c2);

    // binary point sits at bit 192 - s0 of R3:R2:R1:R0; shift left so it lands between bits 62 and 61 of A2
    const int L = 62 + s0;  // 0..63
    const uint64_t A2 = (R3 << L) | ((R2 >> 1) >> (63 - L));
    const uint64_t A1 = (R2 << L) | ((R1 >> 1) >> (63 - L));

    unsigned q = (unsigned)(A2 >> 62);
    uint64_t f_hi = (A2 << 2) | (A1 >> 62);  // fraction, scale 2^-128
    uint64_t f_lo = A1 << 2;

    // round to nearest: fraction >= 1/2 -> next integer, negative remainder
    const uint64_t neg = f_hi >> 63;
    const uint64_t flip = (uint64_t)0 - neg;
    q += (unsigned)neg;
    f_lo = (f_lo ^ flip) + neg;
    f_hi = (f_hi ^ flip) + ((neg != 0 && f_lo == 0) ? 1u : 0u);

    int lead = 0;
    if (f_hi == 0) {  // > 64 bits of cancellation: never for doubles (worst case ~62), kept for safety
        f_hi = f_lo;
        f_lo = 0;
        lead = 64;
    }
    double r = 0.0;
    if (f_hi != 0) {
        const int lz = fabe13_clz64(f_hi);
        const uint64_t mant = (f_hi << lz) | ((f_lo >> 1) >> (63 - lz));  // top bit set
        lead += lz;
        // |fraction| = mant * 2^(-64 - lead); split into an exact 53-bit head and 11-bit tail
        const double scale = fabe13_from_bits((uint64_t)(1023 - 64 - lead) << 52);
        const double f_head = FABE13_MUL(fabe13_u64_to_double(mant & ~0x7FFull), scale);
        const double f_tail = FABE13_MUL(fabe13_u64_to_double(mant & 0x7FFull), scale);
        // times pi/2 in double-double
        const double p = FABE13_MUL(f_head, FABE13_PH_PIO2_HI);
        double t = FABE13_FMA(f_head, FABE13_PH_PIO2_HI, -p);
        t = FABE13_FMA(f_head, FABE13_PH_PIO2_LO, t);
        t = FABE13_FMA(f_tail, FABE13_PH_PIO2_HI, t);
        r = FABE13_ADD(p, t);
    }
    *q_out = (int)(q & 3u);
    return neg ? -r : r;
}

// ---------------------------------------------------------------------------------------------------------------
// 3a. direct minimax core on |r| <= 0.8:  sin r = r + (r z) S(z),  cos r = 1 + z (-1/2 + z C(z)),  z = r^2
//     15 FP64 instructions (3 MUL + 12 FMA)
// ---------------------------------------------------------------------------------------------------------------
FABE13_HD void fabe13_core_poly(double r, double* s, double* c) {
    const double z = FABE13_MUL(r, r);
    double ps = FABE13_FMA(z, FABE13_CF(s6), FABE13_CF(s5));
    double pc = FABE13_FMA(z, FABE13_CF(c6), FABE13_CF(c5));
    ps = FABE13_FMA(z, ps, FABE13_CF(s4));
    pc = FABE13_FMA(z, pc, FABE13_CF(c4));
    ps = FABE13_FMA(z, ps, FABE13_CF(s3));
    pc = FABE13_FMA(z, pc, FABE13_CF(c3));
    ps = FABE13_FMA(z, ps, FABE13_CF(s2));
    pc = FABE13_FMA(z, pc, FABE13_CF(c2));
    ps = FABE13_FMA(z, ps, FABE13_CF(s1));
    pc = FABE13_FMA(z, pc, FABE13_CF(c1));
    const double rz = FABE13_MUL(r, z);
    pc = FABE13_FMA(z, pc, FABE13_CF(neg_half));
    *s = FABE13_FMA(rz, ps, r);
    *c = FABE13_FMA(z, pc, FABE13_CF(one));
}

// ---------------------------------------------------------------------------------------------------------------
// 3b. Psi-Hyperbasis rational core (reference fabe13/fabe13.c:311-330) plus the correction term that makes it
//     a sin/cos:  psi = r/(1 + 3/8 z);  sin r = psi*Pa(psi^2) + (r z) Ds(z);  cos r = Pb(psi^2) + z^2 Dc(z).
//     Pa/Pb are the reference's Taylor-matched cubics; Ds/Dc are fitted in tools/gen_constants.py.
// ---------------------------------------------------------------------------------------------------------------
FABE13_HD void fabe13_core_psi(double r, double* s, double* c) {
    const double z = FABE13_MUL(r, r);
    const double den = FABE13_FMA(FABE13_PCF(three_eighths), z, FABE13_PCF(one));
    const double psi = r / den;  // IEEE-754 correctly rounded on both sides
    const double p = FABE13_MUL(psi, psi);
    double pa = FABE13_FMA(FABE13_PCF(neg_a3), p, FABE13_PCF(a2));
    double pb = FABE13_FMA(FABE13_PCF(neg_b3), p, FABE13_PCF(b2));
    pa = FABE13_FMA(pa, p, FABE13_PCF(neg_a1));
    pb = FABE13_FMA(pb, p, FABE13_PCF(neg_b1));
    pa = FABE13_FMA(pa, p, FABE13_PCF(one));
    const double cos_psi = FABE13_FMA(pb, p, FABE13_PCF(one));
    const double sin_psi = FABE13_MUL(psi, pa);
    double qs = FABE13_PCF(ds)[FABE13_PSI_CORR_TERMS - 1];
    double qc = FABE13_PCF(dc)[FABE13_PSI_CORR_TERMS - 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = FABE13_PSI_CORR_TERMS - 2; i >= 0; --i) {
        qs = FABE13_FMA(z, qs, FABE13_PCF(ds)[i]);
        qc = FABE13_FMA(z, qc, FABE13_PCF(dc)[i]);
    }
    const double rz = FABE13_MUL(r, z);
    const double zz = FABE13_MUL(z, z);
    *s = FABE13_FMA(rz, qs, sin_psi);
    *c = FABE13_FMA(zz, qc, cos_psi);
}

template <int CORE>
FABE13_HD void fabe13_core(double r, double* s, double* c) {
    if (CORE == FABE13_CORE_PSI)
        fabe13_core_psi(r, s, c);
    else
        fabe13_core_poly(r, s, c);
}

// ---------------------------------------------------------------------------------------------------------------
// 4. fix-up on bit patterns: rotate by q = k mod 4 (reference :586-596), then the tiny and special overrides
//    (:597-600).  No FP64-pipe instructions: selects and XORs on the integer pipe.
// ---------------------------------------------------------------------------------------------------------------
// hi-word classification of |x|.  The tiny test uses the high word only: it also catches 2^-27 < |x| < 2^-27(1+2^-20),
// where the polynomial path would return exactly (x, 1.0) as well (x^3/6 < ulp(x)/2 and x^2/2 < 2^-54), so the
// result is the same bit for bit as with the reference's inclusive <= 2^-27 compare (fabe13.c:71, :567).
FABE13_HD uint32_t fabe13_abs_hi(uint64_t x_bits) { return (uint32_t)(x_bits >> 32) & 0x7FFFFFFFu; }
FABE13_HD bool fabe13_is_special(uint32_t ahi) { return ahi >= 0x7FF00000u; }                    // NaN, +-Inf
FABE13_HD bool fabe13_is_tiny(uint32_t ahi) { return ahi <= FABE13_TINY_HI_WORD; }               // |x| <= 2^-27 (+sliver)
FABE13_HD bool fabe13_is_huge(uint32_t ahi) { return (ahi - FABE13_PH_HI_WORD) < (0x7FF00000u - FABE13_PH_HI_WORD); }
// neither tiny, nor Payne-Hanek range, nor special: one subtract and one compare
FABE13_HD bool fabe13_is_plain(uint32_t ahi) {
    return (ahi - (FABE13_TINY_HI_WORD + 1u)) < (FABE13_PH_HI_WORD - (FABE13_TINY_HI_WORD + 1u));
}

// STASH selects what a lane in the Payne-Hanek range leaves behind in the first pass of the array path: x itself in
// its sin slot (FABE13_STASH_SIN; free, it shares the override select) or in its cos slot (FABE13_STASH_COS, used
// when only cosines are produced).  The second pass picks the argument up from there.  FABE13_STASH_NONE is the
// complete fix-up used wherever the reduction already handled every range.
enum { FABE13_STASH_NONE = 0, FABE13_STASH_SIN = 1, FABE13_STASH_COS = 2 };

template <int STASH>
FABE13_HD void fabe13_fixup(uint64_t x_bits, double sin_r, double cos_r, uint32_t q, uint64_t* s_bits, uint64_t* c_bits) {
    const uint64_t sb = fabe13_bits(sin_r);
    const uint64_t cb = fabe13_bits(cos_r);
    const bool swap = (q & 1u) != 0;
    uint64_t so = swap ? cb : sb;  // q: 0 -> ( s,  c)  1 -> ( c, -s)  2 -> (-s, -c)  3 -> (-c,  s)
    uint64_t co = swap ? sb : cb;
    so ^= (uint64_t)(q & 2u) << 62;
    co ^= (uint64_t)((q + 1u) & 2u) << 62;
    // overrides, merged into one select per output: tiny -> (x, 1); NaN/Inf -> (qNaN, qNaN); [stash -> x]
    const uint32_t ahi = fabe13_abs_hi(x_bits);
    const bool special = fabe13_is_special(ahi);
    const bool override_ = (STASH != FABE13_STASH_NONE) ? !fabe13_is_plain(ahi) : (special || fabe13_is_tiny(ahi));
    const uint64_t alt_s = special ? FABE13_QNAN_BITS : x_bits;
    uint64_t alt_c = special ? FABE13_QNAN_BITS : FABE13_ONE_BITS;
    if (STASH == FABE13_STASH_COS) alt_c = fabe13_is_huge(ahi) ? x_bits : alt_c;
    so = override_ ? alt_s : so;
    co = override_ ? alt_c : co;
    *s_bits = so;
    *c_bits = co;
}

// One element, every path inline (used by the host scalar API, by small-N and edge elements on the device and
// by the Payne-Hanek second pass).  k_out: k for |x| < 2^45; q in 0..3 on the Payne-Hanek path; 0 for NaN/Inf.
template <int CORE>
FABE13_HD void fabe13_sincos_one(uint64_t x_bits, const uint64_t* ph_table, uint64_t* s_bits, uint64_t* c_bits,
                                 int64_t* k_out) {
    const uint32_t ahi = fabe13_abs_hi(x_bits);
    double r;
    uint32_t q;
    int64_t k;
    if (fabe13_is_huge(ahi)) {
        int qa;
        r = fabe13_payne_hanek(x_bits & FABE13_ABS_MASK, ph_table, &qa);
        if (x_bits >> 63) {  // odd symmetry: k(-x) = -k(x), r(-x) = -r(x)
            r = -r;
            qa = (4 - qa) & 3;
        }
        q = (uint32_t)qa;
        k = qa;
    } else {
        const double x = fabe13_from_bits(x_bits);
        double biased;
        const double kd = fabe13_quadrant_index(x, &biased);
        r = fabe13_cody_waite(x, kd);
        q = fabe13_q_from_biased(biased);
        k = fabe13_is_special(ahi) ? 0 : fabe13_k_from_biased(biased);
    }
    double sr, cr;
    fabe13_core<CORE>(r, &sr, &cr);
    fabe13_fixup<FABE13_STASH_NONE>(x_bits, sr, cr, q, s_bits, c_bits);
    *k_out = k;
}

#endif  // FABE13_CORE_H
