// bw_probe.cu -- standalone HBM bandwidth probe for the access mix of the sincos path: 8 B read + 16 B written per
// element (1R:2W), against a plain 1R:1W copy.  Not part of the product; used to find the ceiling the streaming
// kernel can be held to and the best load/store flavour, vector width and grid shape.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe tools/bw_probe.cu && ./bw_probe [elements]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

enum { LD_DEFAULT, LD_NA, LD_NC_NA, LD_CS };
enum { ST_DEFAULT, ST_CS, ST_NA, ST_WT };

template <int LD>
__device__ __forceinline__ void ld256(const double* p, uint64_t (&b)[4]) {
    if (LD == LD_DEFAULT) asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(b[0]), "=l"(b[1]), "=l"(b[2]), "=l"(b[3]) : "l"(p));
    if (LD == LD_NA) asm volatile("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(b[0]), "=l"(b[1]), "=l"(b[2]), "=l"(b[3]) : "l"(p));
    if (LD == LD_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(b[0]), "=l"(b[1]), "=l"(b[2]), "=l"(b[3]) : "l"(p));
    if (LD == LD_CS) asm volatile("ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(b[0]), "=l"(b[1]), "=l"(b[2]), "=l"(b[3]) : "l"(p));
}
template <int ST>
__device__ __forceinline__ void st256(double* p, const uint64_t (&b)[4]) {
    if (ST == ST_DEFAULT) asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]) : "memory");
    if (ST == ST_CS) asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]) : "memory");
    if (ST == ST_NA) asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]) : "memory");
    if (ST == ST_WT) asm volatile("st.global.wt.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]) : "memory");
}

// WRITES = 1: copy;  WRITES = 2: the sincos mix.  FLOPS > 0 adds that many dependent DFMA per element (to see when the
// FP64 pipe starts to matter).
template <int LD, int ST, int UNROLL, int WRITES, int FLOPS>
__global__ void __launch_bounds__(256) stream_kernel(const double* __restrict__ in, double* __restrict__ o1, double* __restrict__ o2, uint32_t nvec) {
    const uint32_t stride = gridDim.x * 256;
    for (uint32_t v = blockIdx.x * 256 + threadIdx.x; v < nvec; v += UNROLL * stride) {
        uint64_t b[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (v + u * stride < nvec) ld256<LD>(in + (size_t)(v + u * stride) * 4, b[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (v + u * stride < nvec) {
                uint64_t c[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double d = __longlong_as_double((long long)b[u][j]);
                    double a = d, e = d;
#pragma unroll
                    for (int f = 0; f < FLOPS / 2; ++f) { a = __fma_rn(a, d, 0.5); e = __fma_rn(e, d, 0.25); }
                    b[u][j] = (uint64_t)__double_as_longlong(a);
                    c[j] = (uint64_t)__double_as_longlong(e);
                }
                st256<ST>(o1 + (size_t)(v + u * stride) * 4, b[u]);
                if (WRITES == 2) st256<ST>(o2 + (size_t)(v + u * stride) * 4, c);
            }
        }
    }
}

template <int LD, int ST, int UNROLL, int WRITES, int FLOPS>
void run(const char* name, const double* in, double* o1, double* o2, size_t n, int grid) {
    const uint32_t nvec = (uint32_t)(n / 4);
    if (grid == 0) grid = (int)((nvec + 256u * UNROLL - 1) / (256u * UNROLL));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) stream_kernel<LD, ST, UNROLL, WRITES, FLOPS><<<grid, 256>>>(in, o1, o2, nvec);
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0;
    const int reps = 10;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        stream_kernel<LD, ST, UNROLL, WRITES, FLOPS><<<grid, 256>>>(in, o1, o2, nvec);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
        sum += ms;
    }
    const double bytes = (double)n * 8.0 * (1 + WRITES);
    printf("%-44s grid %7d  avg %8.1f GB/s  best %8.1f GB/s  (%.3f ms)\n", name, grid, bytes / (sum / reps * 1e-3) / 1e9,
           bytes / (best * 1e-3) / 1e9, sum / reps);
    fflush(stdout);
}

int main(int argc, char** argv) {
    size_t n = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)500000000;
    n &= ~(size_t)3;
    double *in, *o1, *o2;
    CK(cudaMalloc(&in, n * 8));
    CK(cudaMalloc(&o1, n * 8));
    CK(cudaMalloc(&o2, n * 8));
    CK(cudaMemset(in, 0, n * 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("elements %zu, SMs %d\n", n, sms);
    {   // cudaMemcpy D2D as the reference point of MEASURED_PEAKS.json (1R:1W)
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        for (int i = 0; i < 3; ++i) CK(cudaMemcpyAsync(o1, in, n * 8, cudaMemcpyDeviceToDevice));
        float best = 1e30f;
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(a); CK(cudaMemcpyAsync(o1, in, n * 8, cudaMemcpyDeviceToDevice)); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
        }
        printf("%-44s               best %8.1f GB/s\n", "cudaMemcpy D2D (1R:1W)", (double)n * 16.0 / (best * 1e-3) / 1e9);
        for (int i = 0; i < 3; ++i) CK(cudaMemsetAsync(o1, 0, n * 8));
        best = 1e30f;
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(a); CK(cudaMemsetAsync(o1, 0, n * 8)); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
        }
        printf("%-44s               best %8.1f GB/s\n", "cudaMemset (write only)", (double)n * 8.0 / (best * 1e-3) / 1e9);
    }
    const int G[] = {0, sms * 2, sms * 3, sms * 4, sms * 6, sms * 8, sms * 16};
    puts("--- 1R:1W copy, 256-bit, NA load / CS store");
    for (int g : G) run<LD_NA, ST_CS, 2, 1, 0>("copy NA/CS unroll2", in, o1, o2, n, g);
    puts("--- 1R:2W (sincos mix), NA load / CS store, unroll x grid");
    for (int g : G) run<LD_NA, ST_CS, 1, 2, 0>("1R2W NA/CS unroll1", in, o1, o2, n, g);
    for (int g : G) run<LD_NA, ST_CS, 2, 2, 0>("1R2W NA/CS unroll2", in, o1, o2, n, g);
    for (int g : G) run<LD_NA, ST_CS, 4, 2, 0>("1R2W NA/CS unroll4", in, o1, o2, n, g);
    puts("--- 1R:2W load/store flavours (unroll 2, grid = 4/SM and one-shot)");
    for (int g : {sms * 4, 0}) {
        run<LD_DEFAULT, ST_DEFAULT, 2, 2, 0>("1R2W default/default", in, o1, o2, n, g);
        run<LD_NA, ST_DEFAULT, 2, 2, 0>("1R2W NA/default", in, o1, o2, n, g);
        run<LD_NC_NA, ST_CS, 2, 2, 0>("1R2W NC.NA/CS", in, o1, o2, n, g);
        run<LD_CS, ST_CS, 2, 2, 0>("1R2W CS/CS", in, o1, o2, n, g);
        run<LD_NA, ST_NA, 2, 2, 0>("1R2W NA/NA", in, o1, o2, n, g);
        run<LD_NA, ST_WT, 2, 2, 0>("1R2W NA/WT", in, o1, o2, n, g);
    }
    puts("--- 1R:2W with dependent DFMA per element (NA/CS, unroll 2, grid = 4/SM): where the FP64 pipe starts to bind");
    run<LD_NA, ST_CS, 2, 2, 8>("1R2W +8 DFMA", in, o1, o2, n, sms * 4);
    run<LD_NA, ST_CS, 2, 2, 16>("1R2W +16 DFMA", in, o1, o2, n, sms * 4);
    run<LD_NA, ST_CS, 2, 2, 24>("1R2W +24 DFMA", in, o1, o2, n, sms * 4);
    run<LD_NA, ST_CS, 2, 2, 32>("1R2W +32 DFMA", in, o1, o2, n, sms * 4);
    run<LD_NA, ST_CS, 2, 2, 48>("1R2W +48 DFMA", in, o1, o2, n, sms * 4);
    run<LD_NA, ST_CS, 2, 2, 64>("1R2W +64 DFMA", in, o1, o2, n, sms * 4);
    run<LD_NA, ST_CS, 1, 2, 24>("1R2W +24 DFMA unroll1 8/SM", in, o1, o2, n, sms * 8);
    run<LD_NA, ST_CS, 1, 2, 24>("1R2W +24 DFMA unroll1 one-shot", in, o1, o2, n, 0);
    run<LD_NA, ST_CS, 2, 2, 24>("1R2W +24 DFMA unroll2 one-shot", in, o1, o2, n, 0);
    puts("--- sustained: 300 back-to-back launches (1R:2W +24 DFMA, unroll 1, one tile per block), avg per 50");
    {
        const uint32_t nvec = (uint32_t)(n / 4);
        const int grid = (int)((nvec + 255u) / 256u);
        cudaEvent_t ev[7];
        for (auto& evt : ev) CK(cudaEventCreate(&evt));
        CK(cudaEventRecord(ev[0]));
        for (int b = 0; b < 6; ++b) {
            for (int i = 0; i < 50; ++i) stream_kernel<LD_NA, ST_CS, 1, 2, 24><<<grid, 256>>>(in, o1, o2, nvec);
            CK(cudaEventRecord(ev[b + 1]));
        }
        CK(cudaDeviceSynchronize());
        for (int b = 0; b < 6; ++b) {
            float ms;
            CK(cudaEventElapsedTime(&ms, ev[b], ev[b + 1]));
            printf("launches %3d..%3d: %8.1f GB/s\n", b * 50, b * 50 + 49, (double)n * 24.0 * 50 / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
