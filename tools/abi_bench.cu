// abi_bench.cu -- latency / throughput of the C ABI entry points measured from C (no Python in the loop).
//   nvcc -O2 -o tools/abi_bench tools/abi_bench.cu -Iinclude -Lfabe_b200/lib -lfabe13 -Xlinker -rpath -Xlinker '$ORIGIN/../fabe_b200/lib'
// Prints, per n: wall time of fabe13_sincos_device + stream sync, GPU time between events, and the host-pointer
// entry (pinned and pageable) wall time.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>

#include "fabe13.h"
#include "fabe13_cuda.h"

static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    const size_t nmax = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)100000000;
    double *d_in, *d_s, *d_c;
    cudaMalloc(&d_in, nmax * 8);
    cudaMalloc(&d_s, nmax * 8);
    cudaMalloc(&d_c, nmax * 8);
    fabe13_debug_fill_uniform_device(d_in, nmax, 0xFABE1305ull, 0, -1e4, 1e4, nullptr);
    cudaDeviceSynchronize();
    double* h_in = (double*)fabe13_cuda_host_alloc(nmax * 8);
    double* h_s = (double*)fabe13_cuda_host_alloc(nmax * 8);
    double* h_c = (double*)fabe13_cuda_host_alloc(nmax * 8);
    cudaMemcpy(h_in, d_in, nmax * 8, cudaMemcpyDeviceToHost);
    std::vector<double> p_in(h_in, h_in + nmax), p_s(nmax), p_c(nmax);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%s\n", fabe13_get_active_implementation_name());
    long long min_n_saved = 0;
    fabe13_cuda_get_option("min_n", &min_n_saved);
    long long min_np_saved = 0;
    fabe13_cuda_get_option("min_n_pageable", &min_np_saved);
    fabe13_cuda_set_option("min_n", 0);  // first table: every host-pointer call on the GPU
    fabe13_cuda_set_option("min_n_pageable", 0);
    printf("%12s %14s %14s %14s %14s %16s %16s\n", "n", "dev wall us", "dev gpu us", "dev Gelem/s", "back2back us", "host pinned us", "host pageable us");
    for (size_t n = 1; n <= nmax; n *= 10) {
        const int reps = n <= 1000000 ? 200 : (n <= 10000000 ? 50 : 10);
        for (int i = 0; i < 5; ++i) fabe13_sincos_device(d_in, d_s, d_c, n, st);
        cudaStreamSynchronize(st);
        double wall = 1e30, gpu = 1e30;
        for (int i = 0; i < reps; ++i) {
            const double t0 = now_us();
            cudaEventRecord(e0, st);
            fabe13_sincos_device(d_in, d_s, d_c, n, st);
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            const double t1 = now_us();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (t1 - t0 < wall) wall = t1 - t0;
            if (ms * 1e3 < gpu) gpu = ms * 1e3;
        }
        // back-to-back: queue `reps` calls, one sync: steady-state cost per call when the caller does not wait
        const double t0 = now_us();
        for (int i = 0; i < reps; ++i) fabe13_sincos_device(d_in, d_s, d_c, n, st);
        cudaStreamSynchronize(st);
        const double b2b = (now_us() - t0) / reps;
        double hp = 1e30, hg = 1e30;
        const int hreps = n <= 1000000 ? 20 : 3;
        fabe13_sincos_host(h_in, h_s, h_c, n);
        for (int i = 0; i < hreps; ++i) {
            const double a = now_us();
            fabe13_sincos_host(h_in, h_s, h_c, n);
            const double b = now_us();
            if (b - a < hp) hp = b - a;
        }
        fabe13_sincos_host(p_in.data(), p_s.data(), p_c.data(), n);
        for (int i = 0; i < hreps; ++i) {
            const double a = now_us();
            fabe13_sincos_host(p_in.data(), p_s.data(), p_c.data(), n);
            const double b = now_us();
            if (b - a < hg) hg = b - a;
        }
        printf("%12zu %14.1f %14.1f %14.3f %14.1f %16.1f %16.1f\n", n, wall, gpu, n / gpu / 1e3, b2b, hp, hg);
        fflush(stdout);
        if (memcmp(h_s, p_s.data(), n * 8) != 0) printf("  !! pinned and pageable host paths disagree\n");
    }
    fabe13_cuda_set_option("min_n", min_n_saved);
    // the small-n route: the legacy void entry on pageable arrays, host core (default min_n) against the GPU forced
    // (min_n = 0: bounce buffer + zero-copy kernel, or the staged pipeline beyond 32768 elements) -- picks "min_n"
    long long min_n_default = 0;
    fabe13_cuda_get_option("min_n", &min_n_default);
    printf("\nsmall-n route (pageable arrays through void fabe13_sincos; options min_n %lld, min_n_pageable %lld)\n", min_n_default, min_np_saved);
    printf("%12s %16s %16s %16s\n", "n", "host core us", "GPU pageable us", "GPU pinned us");
    const size_t small_sizes[] = {1, 10, 32, 100, 1000, 2000, 4000, 8000, 12000, 16000, 24000, 32000, 65536, 131072, 262144, 524288, 1048576};
    for (size_t n : small_sizes) {
        if (n > nmax) break;
        const int reps = n <= 1000 ? 20000 : 2000;
        double t_host = 0, t_gpu = 0, t_pin = 0;
        fabe13_cuda_set_option("min_n", (long long)1 << 31);
        fabe13_cuda_set_option("min_n_pageable", (long long)1 << 31);
        fabe13_sincos(p_in.data(), p_s.data(), p_c.data(), (int)n);
        double a = now_us();
        for (int i = 0; i < reps; ++i) fabe13_sincos(p_in.data(), p_s.data(), p_c.data(), (int)n);
        t_host = (now_us() - a) / reps;
        fabe13_cuda_set_option("min_n", 0);
        fabe13_cuda_set_option("min_n_pageable", 0);
        const int greps = 200;
        fabe13_sincos(p_in.data(), p_s.data(), p_c.data(), (int)n);
        a = now_us();
        for (int i = 0; i < greps; ++i) fabe13_sincos(p_in.data(), p_s.data(), p_c.data(), (int)n);
        t_gpu = (now_us() - a) / greps;
        fabe13_sincos(h_in, h_s, h_c, (int)n);
        a = now_us();
        for (int i = 0; i < greps; ++i) fabe13_sincos(h_in, h_s, h_c, (int)n);
        t_pin = (now_us() - a) / greps;
        fabe13_cuda_set_option("min_n", min_n_default);
        fabe13_cuda_set_option("min_n_pageable", min_np_saved);
        printf("%12zu %16.3f %16.2f %16.2f\n", n, t_host, t_gpu, t_pin);
        fflush(stdout);
    }
    long long launches = 0;
    fabe13_cuda_get_option("launches", &launches);
    printf("kernel launches so far: %lld, last error %d\n", launches, fabe13_cuda_last_error());
    return 0;
}
