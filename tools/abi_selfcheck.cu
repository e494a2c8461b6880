// abi_selfcheck.cu -- C-level self check of the array entry points (compute-sanitizer is closed on the GPU pool, so
// memory safety is probed the manual way: guard words around every output, every alignment offset, every kernel path,
// and a bit-for-bit comparison with the host instantiation of the core).
//   nvcc -O1 -std=c++17 -o tools/abi_selfcheck tools/abi_selfcheck.cu -Iinclude -Lfabe_b200/lib -lfabe13 -Xlinker -rpath -Xlinker '$ORIGIN/../fabe_b200/lib'
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "fabe13.h"
#include "fabe13_cuda.h"

int main() {
    const size_t n = 600011, pad = 8;
    std::vector<double> x(n), s(n + 2 * pad), c(n + 2 * pad), hs(n), hc(n), hd[3];
    for (size_t i = 0; i < n; ++i)
        x[i] = (i % 97 == 3) ? ldexp(1.0 + (i % 1000) / 1000.0, (int)(45 + i % 970)) : (i % 89 == 5 ? NAN : (double)i * 0.37 - 1e5);
    for (size_t i = 200000; i < 260000; ++i) x[i] = ldexp(1.0 + (i % 1000) / 1000.0, (int)(45 + i % 970));  // dense-warp route
    fabe13_debug_host_core(x.data(), hs.data(), hc.data(), nullptr, n, 0);
    for (int op = 3; op <= 5; ++op) {  // tan, cot, sinc
        hd[op - 3].resize(n);
        fabe13_debug_host_derived(x.data(), hd[op - 3].data(), n, op, 0);
    }
    double *dx, *ds, *dc;
    cudaMalloc(&dx, (n + 2 * pad) * 8);
    cudaMalloc(&ds, (n + 2 * pad) * 8);
    cudaMalloc(&dc, (n + 2 * pad) * 8);
    const double guard = 12345.678;
    std::vector<double> g(n + 2 * pad, guard);
    int bad = 0;
    for (int off = 0; off < 4; ++off) {
        for (long long small : {0LL, 1LL << 30}) {
            fabe13_cuda_set_option("small_n", small);
            for (long long shift : {0LL, 1LL, 2LL, 3LL, 20LL}) {  // worklist designs 0..3; 20 = design 1 with a tiny list
                fabe13_cuda_set_option("worklist", shift == 20 ? 1 : shift);
                fabe13_cuda_set_option("worklist_shift", shift == 20 ? 20 : 3);
                for (int mode = 0; mode < 6; ++mode) {  // sincos, sin only, cos only, tan, cot, sinc
                    cudaMemcpy(ds, g.data(), (n + 2 * pad) * 8, cudaMemcpyHostToDevice);
                    cudaMemcpy(dc, g.data(), (n + 2 * pad) * 8, cudaMemcpyHostToDevice);
                    cudaMemcpy(dx + off, x.data(), n * 8, cudaMemcpyHostToDevice);
                    int rc = mode == 0   ? fabe13_sincos_device(dx + off, ds + pad + off, dc + pad + off, n, nullptr)
                             : mode == 1 ? fabe13_sin_device(dx + off, ds + pad + off, n, nullptr)
                             : mode == 2 ? fabe13_cos_device(dx + off, dc + pad + off, n, nullptr)
                             : mode == 3 ? fabe13_tan_device(dx + off, ds + pad + off, n, nullptr)
                             : mode == 4 ? fabe13_cot_device(dx + off, ds + pad + off, n, nullptr)
                                         : fabe13_sinc_device(dx + off, ds + pad + off, n, nullptr);
                    cudaDeviceSynchronize();
                    cudaMemcpy(s.data(), ds, (n + 2 * pad) * 8, cudaMemcpyDeviceToHost);
                    cudaMemcpy(c.data(), dc, (n + 2 * pad) * 8, cudaMemcpyDeviceToHost);
                    int mism = 0, guards = 0;
                    for (size_t i = 0; i < n; ++i) {
                        if (mode < 2 && memcmp(&s[pad + off + i], &hs[i], 8)) ++mism;
                        if ((mode == 0 || mode == 2) && memcmp(&c[pad + off + i], &hc[i], 8)) ++mism;
                        if (mode >= 3 && memcmp(&s[pad + off + i], &hd[mode - 3][i], 8)) ++mism;
                    }
                    for (size_t i = 0; i < n + 2 * pad; ++i) {
                        const bool inside = i >= pad + off && i < pad + off + n;
                        if ((!inside || mode == 2) && s[i] != guard) ++guards;
                        if ((!inside || mode == 1 || mode >= 3) && c[i] != guard) ++guards;
                    }
                    printf("off %d small_n %lld worklist %lld mode %d rc %d mismatches %d guard-hits %d\n", off, small, shift, mode, rc, mism, guards);
                    bad += mism + guards + (rc != 0);
                }
            }
        }
    }
    std::vector<double> ps(n), pc(n);
    fabe13_cuda_set_option("min_n_pageable", 0);
    fabe13_cuda_set_option("min_n", 0);  // the GPU routes of the host entry, also for the 1000-element call below
    fabe13_sincos_host(x.data(), ps.data(), pc.data(), n);
    for (size_t i = 0; i < n; ++i) bad += memcmp(&ps[i], &hs[i], 8) != 0 || memcmp(&pc[i], &hc[i], 8) != 0;
    fabe13_sincos_host(x.data(), ps.data(), pc.data(), 1000);
    for (size_t i = 0; i < 1000; ++i) bad += memcmp(&ps[i], &hs[i], 8) != 0;
    printf("total problems %d, last CUDA error %d\n", bad, fabe13_cuda_last_error());
    return bad != 0;
}
