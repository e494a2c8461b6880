// fabe13_hostcore.cpp -- the HOST instantiation of the shared per-element core (fabe13_core.h) over arrays.
//
// This is product code, not the oracle: the very functions the kernels compile for the device, compiled by the host
// compiler, so every element gets the same bits as on the GPU (tests/test_gpu_parity.py compares them).  It serves
//   * host-pointer calls below option "min_n" (reference: n < 32 -> scalar loop, fabe13/fabe13.c:226-229 -- there to
//     dodge SIMD set-up, here to dodge a kernel launch: ~13 us against ~2 ns per element),
//   * the void fabe13_sincos() entry when CUDA cannot be initialised or a launch fails (the reference's ABI has no
//     error channel, fabe13/fabe13.h:51-56, so the arrays must still be filled),
//   * the scalar API (fabe13_sin, ...: reference fabe13/fabe13.c:262-266).
// Layout of the loop mirrors the fused kernel: every lane runs the plain path (reference-exact k, Cody-Waite, core,
// rotation) branch-free, so the host compiler vectorises it (AVX-512 / AVX2 clones picked at run time); blocks that
// hold a tiny, special or Payne-Hanek-range element patch those lanes with the every-range routine afterwards.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "fabe13_core.h"
#include "fabe13_hostcore.h"

namespace fabe13 {
namespace {

const uint32_t h_ph_stream[FABE13_PH_STREAM_WORDS] = FABE13_PH_STREAM;

constexpr size_t kStreamFrom = (size_t)1 << 20;  // elements (8 MB per array)
constexpr int kBlock = 128;  // elements per block: inputs are copied to a local array first, so in == sin_out is safe

// one block (m8 <= kBlock elements, a multiple of 8) of plain-path arithmetic; returns the number of lanes that are NOT plain (to be patched by the caller)
#define FABE13_HOST_BLOCK_BODY(CORE)                                                                              \
    int rare = 0;                                                                                                 \
    for (int i = 0; i < m8; ++i) {                                                                                \
        const double v = x[i];                                                                                    \
        double biased;                                                                                            \
        const double kd = fabe13_quadrant_index(v, &biased);                                                      \
        const double r = fabe13_cody_waite(v, kd);                                                                \
        double sr, cr;                                                                                            \
        fabe13_core<CORE>(r, &sr, &cr);                                                                           \
        uint64_t sb, cb;                                                                                          \
        fabe13_rotate(sr, cr, fabe13_q_from_biased(biased), &sb, &cb);                                            \
        s[i] = sb;                                                                                                \
        c[i] = cb;                                                                                                \
        rare += fabe13_is_plain(fabe13_abs_hi(fabe13_bits(v))) ? 0 : 1;                                           \
    }                                                                                                             \
    return rare;

template <int CORE>
int block_generic(const double* __restrict x, uint64_t* __restrict s, uint64_t* __restrict c, int m8) {
    FABE13_HOST_BLOCK_BODY(CORE)
}
#if defined(__x86_64__)
template <int CORE>
__attribute__((target("avx2,fma"))) int block_avx2(const double* __restrict x, uint64_t* __restrict s, uint64_t* __restrict c, int m8) {
    FABE13_HOST_BLOCK_BODY(CORE)
}
template <int CORE>
__attribute__((target("avx512f,avx512dq,avx512vl,fma"))) int block_avx512(const double* __restrict x, uint64_t* __restrict s,
                                                                                uint64_t* __restrict c, int m8) {
    FABE13_HOST_BLOCK_BODY(CORE)
}
#endif

typedef int (*BlockFn)(const double*, uint64_t*, uint64_t*, int);

template <int CORE>
BlockFn pick_block() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl"))
        return block_avx512<CORE>;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return block_avx2<CORE>;
#endif
    return block_generic<CORE>;
}

template <int CORE>
void one(uint64_t xb, uint64_t* sb, uint64_t* cb) {
    int64_t k;
    fabe13_sincos_one<CORE>(xb, h_ph_stream, sb, cb, &k);
}

// Results of a LARGE array leave the block buffers with non-temporal stores: nobody reads them again soon, and an
// ordinary store would first fetch every destination line for ownership -- 40 instead of 24 bytes of memory traffic per
// element, on a loop that is bound by the host's memory bandwidth once a few threads run it.  8-byte MOVNTI up to the
// first 32-byte boundary and after the last one, 32-byte VMOVNTDQ between them; the write-combining buffers merge
// them into whole lines.  The caller fences before it reports completion.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void put_stream_avx2(double* dst, const uint64_t* src, size_t m) {
    size_t i = 0;
    for (; i < m && ((uintptr_t)(dst + i) & 31u) != 0; ++i) _mm_stream_si64((long long*)(dst + i), (long long)src[i]);
    for (; i + 4 <= m; i += 4)
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_loadu_si256((const __m256i*)(src + i)));
    for (; i < m; ++i) _mm_stream_si64((long long*)(dst + i), (long long)src[i]);
}
void put_stream_sse2(double* dst, const uint64_t* src, size_t m) {
    for (size_t i = 0; i < m; ++i) _mm_stream_si64((long long*)(dst + i), (long long)src[i]);
}
#endif

typedef void (*PutFn)(double*, const uint64_t*, size_t);

void put_plain(double* dst, const uint64_t* src, size_t m) { memcpy(dst, src, m * sizeof(double)); }

PutFn pick_put(bool stream) {
#if defined(__x86_64__)
    if (stream) return __builtin_cpu_supports("avx2") ? put_stream_avx2 : put_stream_sse2;
#endif
    (void)stream;
    return put_plain;
}

void stores_complete(bool stream) {
#if defined(__x86_64__)
    if (stream) _mm_sfence();
#endif
    (void)stream;
}

template <int CORE>
void range(const double* in, double* sin_out, double* cos_out, size_t n, int op, bool stream) {
    static const BlockFn block = pick_block<CORE>();
    const PutFn put = pick_put(stream);
    alignas(64) double x[kBlock];
    alignas(64) uint64_t s[kBlock], c[kBlock];
    for (size_t off = 0; off < n; off += kBlock) {
        const size_t m = n - off < (size_t)kBlock ? n - off : (size_t)kBlock;
        memcpy(x, in + off, m * sizeof(double));
        const size_t m8 = (m + 7) & ~(size_t)7;  // whole vectors; short calls do not pay for a full block
        for (size_t i = m; i < m8; ++i) x[i] = 1.0;  // a plain lane
        if (block(x, s, c, (int)m8) != 0) {
            for (size_t i = 0; i < m; ++i) {
                const uint64_t xb = fabe13_bits(x[i]);
                if (!fabe13_is_plain(fabe13_abs_hi(xb))) one<CORE>(xb, &s[i], &c[i]);
            }
        }
        if (op != FABE13_OP_SINCOS) {
            for (size_t i = 0; i < m; ++i) s[i] = fabe13_derived_rt(op, fabe13_bits(x[i]), s[i], c[i]);
        }
        if (sin_out) put(sin_out + off, s, m);
        if (cos_out) put(cos_out + off, c, m);
    }
    stores_complete(stream);
}

template <int CORE>
void rangef(const float* in, float* sin_out, float* cos_out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float sf, cf;
        fabe13_sincosf_one<CORE>(in[i], h_ph_stream, &sf, &cf);
        sin_out[i] = sf;
        cos_out[i] = cf;
    }
}

// split [0, n) into `threads` contiguous pieces (multiples of kBlock) and run fn(begin, end) on each
template <typename Fn>
void parallel(size_t n, int threads, size_t min_per_thread, Fn fn) {
    size_t want = n / min_per_thread;
    if (want < 1) want = 1;
    if ((size_t)threads > want) threads = (int)want;
    if (threads <= 1) {
        fn((size_t)0, n);
        return;
    }
    const size_t per = ((n / (size_t)threads) + kBlock - 1) / kBlock * kBlock;
    std::vector<std::thread> th;
    th.reserve((size_t)threads);
    size_t done_to = per < n ? per : n;  // [0, done_to) is the calling thread's piece; pieces behind it go to threads
    try {
        for (int t = 1; t < threads; ++t) {
            const size_t b = per * (size_t)t;
            if (b >= n) break;
            const size_t e = (t == threads - 1 || b + per > n) ? n : b + per;
            th.emplace_back([=] { fn(b, e); });
            done_to = e;
        }
    } catch (...) {  // a thread could not be started: what it and its successors would have done runs here (below)
    }
    fn((size_t)0, per < n ? per : n);
    const size_t threaded_end = th.empty() ? (per < n ? per : n) : done_to;
    if (threaded_end < n) fn(threaded_end, n);
    for (std::thread& x : th) x.join();
}

}  // namespace

void host_sincos_one(double x, int core, double* s, double* c, long long* k) {
    uint64_t sb, cb;
    int64_t kk;
    if (core == FABE13_CORE_PSI)
        fabe13_sincos_one<FABE13_CORE_PSI>(fabe13_bits(x), h_ph_stream, &sb, &cb, &kk);
    else
        fabe13_sincos_one<FABE13_CORE_POLY>(fabe13_bits(x), h_ph_stream, &sb, &cb, &kk);
    *s = fabe13_from_bits(sb);
    *c = fabe13_from_bits(cb);
    if (k) *k = (long long)kk;
}

void host_sincosf_one(float x, int core, float* s, float* c) {
    if (core == FABE13_CORE_PSI)
        fabe13_sincosf_one<FABE13_CORE_PSI>(x, h_ph_stream, s, c);
    else
        fabe13_sincosf_one<FABE13_CORE_POLY>(x, h_ph_stream, s, c);
}

double host_derived_one(double x, int op, int core) {
    double s, c;
    host_sincos_one(x, core, &s, &c, nullptr);
    return fabe13_from_bits(fabe13_derived_rt(op, fabe13_bits(x), fabe13_bits(s), fabe13_bits(c)));
}

void host_sincos_array(const double* in, double* sin_out, double* cos_out, size_t n, int op, int core, int threads,
                       int stream_stores) {
    if (n == 0 || (!sin_out && !cos_out)) return;
    // non-temporal result stores from kStreamFrom elements on (an array that size does not stay in the caches anyway)
    const bool stream = stream_stores > 0 || (stream_stores < 0 && n >= kStreamFrom);
    parallel(n, threads, (size_t)1 << 16, [=](size_t b, size_t e) {
        if (core == FABE13_CORE_PSI)
            range<FABE13_CORE_PSI>(in + b, sin_out ? sin_out + b : nullptr, cos_out ? cos_out + b : nullptr, e - b, op, stream);
        else
            range<FABE13_CORE_POLY>(in + b, sin_out ? sin_out + b : nullptr, cos_out ? cos_out + b : nullptr, e - b, op, stream);
    });
}

void host_sincosf_array(const float* in, float* sin_out, float* cos_out, size_t n, int core, int threads) {
    if (n == 0) return;
    parallel(n, threads, (size_t)1 << 16, [=](size_t b, size_t e) {
        if (core == FABE13_CORE_PSI)
            rangef<FABE13_CORE_PSI>(in + b, sin_out + b, cos_out + b, e - b);
        else
            rangef<FABE13_CORE_POLY>(in + b, sin_out + b, cos_out + b, e - b);
    });
}

void host_reduce_huge(const double* in, double* r, int* q, double* r_exact, int* q_exact, int* fast_ok, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const uint64_t xb = fabe13_bits(in[i]);
        if (!fabe13_is_huge(fabe13_abs_hi(xb))) {  // not an argument of this reduction (the table starts at 2^45): say so
            r[i] = r_exact[i] = fabe13_from_bits(FABE13_QNAN_BITS);
            q[i] = q_exact[i] = -1;
            fast_ok[i] = 0;
            continue;
        }
        uint32_t qq;
        r[i] = fabe13_payne_hanek(xb, h_ph_stream, &qq);
        q[i] = (int)qq;
        r_exact[i] = fabe13_payne_hanek_exact(xb, h_ph_stream, &qq);
        q_exact[i] = (int)qq;
        uint64_t rb;
        fast_ok[i] = fabe13_payne_hanek_fast(xb, &rb, &qq) ? 1 : 0;
    }
}

const char* host_isa_name() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl"))
        return "AVX-512";
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return "AVX2";
#endif
    return "scalar";
}

}  // namespace fabe13
