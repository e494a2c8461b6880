// host_sanitize.cu -- the host side of libfabe13 (scalar API, the host instantiation of the per-element core, option
// table, shard arithmetic, the oracle-independent debug hooks) under AddressSanitizer + UndefinedBehaviorSanitizer.
// No GPU needed: nothing here launches a kernel.  Stands in for SURVEY section 5's "host ASan/UBSan build of shim".
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -cudart static \
//        -Xcompiler -fsanitize=address,undefined,-fno-sanitize-recover=all,-ffp-contract=off,-mfma \
//        -o tools/host_sanitize tools/host_sanitize.cu fabe_b200/csrc/fabe13_kernels.cu fabe_b200/csrc/fabe13_shim.cu \
//        fabe_b200/csrc/fabe13_hostcore.cpp -lpthread
//   tools/host_sanitize        (exit code 0 and "host_sanitize: ok" = clean)
// and the same with -Xcompiler -fsanitize=thread instead of address,undefined for the threaded section at the end.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../include/fabe13.h"
#include "../include/fabe13_cuda.h"

static uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double from_bits(uint64_t u) {
    double x;
    memcpy(&x, &u, sizeof x);
    return x;
}

int main() {
    std::vector<double> in;
    // every exponent, with all-zero, all-one and alternating mantissas, both signs: walks the whole 2/pi table
    for (uint64_t e = 0; e < 2048; ++e)
        for (uint64_t m : {0x0ull, 0xFFFFFFFFFFFFFull, 0xAAAAAAAAAAAAAull, 0x5555555555555ull, 0x8000000000000ull, 0x1ull})
            for (uint64_t s = 0; s < 2; ++s) in.push_back(from_bits((s << 63) | (e << 52) | m));
    // multiples of pi/2 and their neighbours; the double closest to a multiple of pi/2
    for (int k = -2000; k <= 2000; ++k) {
        const double x = k * 1.5707963267948966;
        in.push_back(x);
        in.push_back(nextafter(x, INFINITY));
        in.push_back(nextafter(x, -INFINITY));
    }
    in.push_back(ldexp(6381956970095103.0, 797));
    in.push_back(1e22);
    in.push_back(0x1p45);
    in.push_back(nextafter(0x1p45, 0.0));
    uint64_t seed = 0xFABE13;
    for (int i = 0; i < 2000000; ++i) in.push_back(from_bits(splitmix64(seed)));  // random bit patterns
    const size_t n = in.size();
    // exactly-sized heap arrays: ASan sees any access past either end
    std::vector<double> s(n), c(n), d(n);
    std::vector<long long> k(n);
    double acc = 0;
    for (int core = 0; core < 2; ++core) {
        fabe13_debug_host_core(in.data(), s.data(), c.data(), k.data(), n, core);
        for (size_t i = 0; i < n; ++i)
            if (s[i] == s[i]) acc += s[i] + c[i];
        for (int op = 3; op <= 5; ++op) {
            fabe13_debug_host_derived(in.data(), d.data(), n, op, core);
            for (size_t i = 0; i < n; i += 97)
                if (d[i] == d[i] && fabs(d[i]) < 1e300) acc += d[i];
        }
    }
    // the host ARRAY loop (small-n route and fallback of the product): odd sizes on exactly-sized arrays, in place,
    // single outputs, derived functions, several threads -- bit for bit against the per-element routine
    for (size_t m : {(size_t)1, (size_t)7, (size_t)8, (size_t)127, (size_t)128, (size_t)129, (size_t)1000, (size_t)65537, (size_t)300001}) {
        const size_t off = (m * 7919) % (n - m);
        for (int core = 0; core < 2; ++core) {
            std::vector<double> x(in.begin() + off, in.begin() + off + m), ws(m), wc(m), as(m), ac(m), inplace(x);
            fabe13_debug_host_core(x.data(), ws.data(), wc.data(), nullptr, m, core);
            fabe13_debug_host_array(x.data(), as.data(), ac.data(), m, 0, core, 3);
            fabe13_debug_host_array(inplace.data(), inplace.data(), nullptr, m, 0, core, 1);
            if (memcmp(ws.data(), as.data(), m * 8) || memcmp(wc.data(), ac.data(), m * 8) || memcmp(ws.data(), inplace.data(), m * 8)) {
                fprintf(stderr, "host array loop differs from the per-element routine (m=%zu core=%d)\n", m, core);
                return 1;
            }
            fabe13_debug_host_array(x.data(), nullptr, ac.data(), m, 0, core, 2);
            if (memcmp(wc.data(), ac.data(), m * 8)) return 1;
            // the same with non-temporal result stores (what large arrays get), outputs at odd 8-byte phases
            if (m > 3) {
                fabe13_debug_host_array_stores(x.data(), as.data() + 1, ac.data() + 3, m - 3, 0, core, 2, 1);
                if (memcmp(ws.data(), as.data() + 1, (m - 3) * 8) || memcmp(wc.data(), ac.data() + 3, (m - 3) * 8)) {
                    fprintf(stderr, "host array loop with streaming stores differs (m=%zu core=%d)\n", m, core);
                    return 1;
                }
            }
            for (int op = 3; op <= 5; ++op) {
                fabe13_debug_host_derived(x.data(), ws.data(), m, op, core);
                fabe13_debug_host_array(x.data(), as.data(), nullptr, m, op, core, 2);
                if (memcmp(ws.data(), as.data(), m * 8)) {
                    fprintf(stderr, "host array loop, derived op %d differs (m=%zu core=%d)\n", op, m, core);
                    return 1;
                }
            }
        }
    }
    // the C entry points themselves.  Small n stays on the host core whether or not a GPU is visible; on a box without
    // one, the void entry falls back to the host core for large n too and the int entries report the error.
    {
        const size_t m = 5000;
        std::vector<double> x(in.begin(), in.begin() + m), ws(m), wc(m), as(m), ac(m);
        fabe13_debug_host_core(x.data(), ws.data(), wc.data(), nullptr, m, 0);
        fabe13_sincos(x.data(), as.data(), ac.data(), (int)m);
        if (memcmp(ws.data(), as.data(), m * 8) || memcmp(wc.data(), ac.data(), m * 8)) {
            fprintf(stderr, "fabe13_sincos (small n) differs from the per-element routine\n");
            return 1;
        }
        if (fabe13_sin_host(x.data(), as.data(), m) != 0 || memcmp(ws.data(), as.data(), m * 8)) return 1;
        if (fabe13_cos_host(x.data(), ac.data(), m) != 0 || memcmp(wc.data(), ac.data(), m * 8)) return 1;
        if (fabe13_cuda_device_count() == 0) {
            const size_t big = 400000;  // above the pageable cut (262144): the GPU route, which cannot run here
            std::vector<double> bx(in.begin(), in.begin() + big), bs(big), bc(big), hs(big), hc(big);
            fabe13_debug_host_core(bx.data(), hs.data(), hc.data(), nullptr, big, 0);
            fabe13_cuda_clear_error();
            fabe13_sincos(bx.data(), bs.data(), bc.data(), (int)big);  // no GPU: host fallback, error recorded
            if (memcmp(hs.data(), bs.data(), big * 8) || memcmp(hc.data(), bc.data(), big * 8) || fabe13_cuda_last_error() == 0) {
                fprintf(stderr, "fabe13_sincos fallback without a GPU is broken\n");
                return 1;
            }
            if (fabe13_sincos_host(bx.data(), bs.data(), bc.data(), big) == 0) {
                fprintf(stderr, "fabe13_sincos_host must report the missing GPU\n");
                return 1;
            }
            fabe13_cuda_clear_error();
        }
        (void)fabe13_cuda_trim();
        (void)fabe13_cuda_shutdown();
    }
    std::vector<float> fin(n), fs(n), fc(n);
    for (size_t i = 0; i < n; ++i) {
        const uint64_t b = splitmix64(seed);
        uint32_t w = (uint32_t)b;
        memcpy(&fin[i], &w, 4);
    }
    for (int core = 0; core < 2; ++core) fabe13_debug_host_sincosf(fin.data(), fs.data(), fc.data(), n, core);
    for (size_t i = 0; i < n; i += 1000) {  // scalar API (the array core, one element)
        const double x = in[i];
        acc += fabe13_sin(x) == fabe13_sin(x) ? fabe13_cos(x) : 0;
        const double t = fabe13_tan(x), ct = fabe13_cot(x), sc = fabe13_sinc(x);
        if (t == t && fabs(t) < 1e300) acc += t;
        if (ct == ct && fabs(ct) < 1e300) acc += ct;
        if (sc == sc) acc += sc;
        (void)fabe13_atan(x);
        (void)fabe13_asin(x);
        (void)fabe13_acos(x);
    }
    // no-ops and argument checks that must not touch memory
    fabe13_sincos(nullptr, nullptr, nullptr, 0);
    fabe13_sincos(nullptr, nullptr, nullptr, -5);
    (void)fabe13_sincos_host(nullptr, nullptr, nullptr, 0);
    (void)fabe13_sincosf_host(nullptr, nullptr, nullptr, 0);
    (void)fabe13_get_active_simd_width();
    // shard arithmetic
    for (size_t total : {(size_t)0, (size_t)1, (size_t)3, (size_t)1000003, (size_t)1000000000, (size_t)1 << 40})
        for (int world : {1, 2, 3, 8, 64}) {
            size_t prev = 0;
            for (int r = 0; r < world; ++r) {
                size_t b, e;
                fabe13_shard_bounds(total, world, r, &b, &e);
                if (b != prev || e < b || e > total) {
                    fprintf(stderr, "shard bounds broken: n=%zu world=%d rank=%d [%zu,%zu)\n", total, world, r, b, e);
                    return 1;
                }
                prev = e;
            }
            if (prev != total) return 1;
        }
    // option table: every documented key, out-of-range values, unknown and empty keys
    const char* keys[] = {"core", "unroll", "blocks_per_sm", "tiles_per_block", "small_n", "worklist", "hybrid_min_n", "local_min",
                          "dense_hint", "direct_max", "worklist_shift", "second_pass_blocks_per_sm", "pdl", "chunk_elems", "host_devices",
                          "stage_threads", "zero_copy_max", "stream_copy", "min_n", "min_n_pageable", "fallback", "host_threads", "launches",
                          "host_core_calls", "fallbacks", "no_such_option", ""};
    for (const char* key : keys) {
        long long v = 0;
        const int rc = fabe13_cuda_get_option(key, &v);
        if (rc == 0) (void)fabe13_cuda_set_option(key, v);
        (void)fabe13_cuda_set_option(key, -1);
        (void)fabe13_cuda_set_option(key, (long long)1 << 62);
    }
    {
        long long v = 0;
        if (fabe13_cuda_set_option(nullptr, 1) != -1 || fabe13_cuda_get_option(nullptr, &v) != -1 ||
            fabe13_cuda_get_option("core", nullptr) != -1) {
            fprintf(stderr, "null option arguments must be refused with -1\n");
            return 1;
        }
    }
    // the entry points are re-entrant (reference: globals written once at load, fabe13.c:209-214): scalar calls, option
    // reads and writes and the host instantiation from eight threads at once (the interesting build for this part is
    // -fsanitize=thread)
    {
        std::vector<std::thread> th;
        std::vector<double> sums(8, 0.0);
        for (int w = 0; w < 8; ++w)
            th.emplace_back([&, w] {
                std::vector<double> ls(4096), lc(4096);
                for (int rep = 0; rep < 20; ++rep) {
                    fabe13_debug_host_core(in.data() + (size_t)w * 4096, ls.data(), lc.data(), nullptr, 4096, rep & 1);
                    (void)fabe13_cuda_set_option("core", rep & 1);
                    long long v = 0;
                    (void)fabe13_cuda_get_option("core", &v);
                    fabe13_sincos(in.data() + (size_t)w * 4096, ls.data(), lc.data(), 4096);  // small-n route, per thread
                    fabe13_cuda_clear_error();
                    (void)fabe13_cuda_last_error();
                    for (int i = 0; i < 256; ++i) {
                        const double y = fabe13_sin(0.001 * i + w);
                        if (y == y) sums[(size_t)w] += y + fabe13_cos(1e300 * (i + 1)) + fabe13_tan(0.5 + i);
                    }
                }
            });
        for (std::thread& t : th) t.join();
        for (double v : sums) acc += v;
        (void)fabe13_cuda_set_option("core", 0);
    }
    (void)fabe13_sincos_workspace_bytes(0);
    (void)fabe13_sincos_workspace_bytes((size_t)1 << 30);
    printf("host_sanitize: ok (%zu doubles x 2 cores, checksum %.17g)\n", n, acc);
    return 0;
}
