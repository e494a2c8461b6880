// tma_probe.cu -- does staging the sincos stream through shared memory with bulk async copies (the TMA engine:
// cp.async.bulk + mbarrier) move a 1R:2W FP64 stream faster than plain 256-bit LDG/STG?  Three kernels with the same
// per-element work (FLOPS dependent DFMA, two outputs), 8 B read + 16 B written per element:
//   ldg_tile        one tile of 1024 doubles per block, LDG.E.NA.256 / STG.E.EF.256 (the product kernel's shape)
//   tma_tile        one tile per block: bulk load into shared memory, compute from/to shared memory, two bulk stores
//   tma_persistent  grid of SMs*k blocks, STAGES-deep ring per block: loads of later tiles in flight while one is computed
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu && tools/tma_probe [elements]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int kThreads = 256;
constexpr int kTile = 1024;               // doubles per tile: 8 KB in, 2 x 8 KB out
constexpr int kFlops = 24;                // dependent DFMA per element, like the sincos lane

__device__ __forceinline__ void work(double d, double& a, double& e) {
    a = d;
    e = d;
#pragma unroll
    for (int f = 0; f < kFlops / 2; ++f) {
        a = __fma_rn(a, d, 0.5);
        e = __fma_rn(e, d, 0.25);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads) ldg_tile(const double* __restrict__ in, double* __restrict__ o1, double* __restrict__ o2, uint32_t ntiles) {
    const uint32_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    const size_t e = (size_t)tile * kTile + threadIdx.x * 4;
    uint64_t b[4], c[4];
    asm volatile("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(b[0]), "=l"(b[1]), "=l"(b[2]), "=l"(b[3]) : "l"(in + e));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double a, x;
        work(__longlong_as_double((long long)b[j]), a, x);
        b[j] = (uint64_t)__double_as_longlong(a);
        c[j] = (uint64_t)__double_as_longlong(x);
    }
    asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(o1 + e), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]) : "memory");
    asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(o2 + e), "l"(c[0]), "l"(c[1]), "l"(c[2]), "l"(c[3]) : "memory");
}

struct __align__(128) Stage {
    double in[kTile];
    double o1[kTile];
    double o2[kTile];
};

__device__ __forceinline__ void compute_stage(Stage& st) {
    double4 v = reinterpret_cast<const double4*>(st.in)[threadIdx.x];  // hmm: 32-byte vector from shared memory
    double4 r1, r2;
    work(v.x, r1.x, r2.x);
    work(v.y, r1.y, r2.y);
    work(v.z, r1.z, r2.z);
    work(v.w, r1.w, r2.w);
    reinterpret_cast<double4*>(st.o1)[threadIdx.x] = r1;
    reinterpret_cast<double4*>(st.o2)[threadIdx.x] = r2;
}

__global__ void __launch_bounds__(kThreads) tma_tile(const double* __restrict__ in, double* __restrict__ o1, double* __restrict__ o2, uint32_t ntiles) {
    extern __shared__ __align__(128) unsigned char raw[];
    Stage& st = *reinterpret_cast<Stage*>(raw);
    __shared__ uint64_t bar;
    const uint32_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, kTile * 8);
        bulk_load(st.in, in + (size_t)tile * kTile, kTile * 8, &bar);
    }
    mbar_wait(&bar, 0);
    compute_stage(st);
    fence_async_proxy();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_store(o1 + (size_t)tile * kTile, st.o1, kTile * 8);
        bulk_store(o2 + (size_t)tile * kTile, st.o2, kTile * 8);
        bulk_commit();
        bulk_wait_read<0>();
    }
}

template <int STAGES>
__global__ void __launch_bounds__(kThreads) tma_persistent(const double* __restrict__ in, double* __restrict__ o1, double* __restrict__ o2, uint32_t ntiles) {
    extern __shared__ __align__(128) unsigned char raw[];
    Stage* st = reinterpret_cast<Stage*>(raw);
    __shared__ uint64_t full[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // tiles of this block: blockIdx.x, + gridDim.x, ...  (all blocks advance together: one moving window in DRAM)
    const uint32_t first = blockIdx.x, step = gridDim.x;
    uint32_t issued = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const uint32_t t = first + issued * step;
            if (t < ntiles) {
                mbar_expect_tx(&full[s], kTile * 8);
                bulk_load(st[s].in, in + (size_t)t * kTile, kTile * 8, &full[s]);
            }
            ++issued;
        }
    }
    uint32_t it = 0;
    for (uint32_t t = first; t < ntiles; t += step, ++it) {
        const int s = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        mbar_wait(&full[s], parity);
        // the stores issued from this stage STAGES iterations ago must have finished reading its output buffers
        if (threadIdx.x == 0) bulk_wait_read<STAGES - 1>();
        __syncthreads();
        compute_stage(st[s]);
        fence_async_proxy();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store(o1 + (size_t)t * kTile, st[s].o1, kTile * 8);
            bulk_store(o2 + (size_t)t * kTile, st[s].o2, kTile * 8);
            bulk_commit();
            const uint32_t tn = first + (it + STAGES) * step;  // refill this stage's input buffer (already consumed)
            if (tn < ntiles) {
                mbar_expect_tx(&full[s], kTile * 8);
                bulk_load(st[s].in, in + (size_t)tn * kTile, kTile * 8, &full[s]);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait_read<0>();
}

template <typename F>
void timeit(const char* name, size_t n, int grid, F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0;
    const int reps = 10;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
        sum += ms;
    }
    CK(cudaGetLastError());
    const double bytes = (double)n * 24.0;
    printf("%-44s grid %7d  avg %8.1f GB/s  best %8.1f GB/s  (%.3f ms)\n", name, grid, bytes / (sum / reps * 1e-3) / 1e9,
           bytes / (best * 1e-3) / 1e9, sum / reps);
    fflush(stdout);
}

int main(int argc, char** argv) {
    size_t n = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)500000000;
    n -= n % kTile;
    const uint32_t ntiles = (uint32_t)(n / kTile);
    double *in, *o1, *o2;
    CK(cudaMalloc(&in, n * 8));
    CK(cudaMalloc(&o1, n * 8));
    CK(cudaMalloc(&o2, n * 8));
    CK(cudaMemset(in, 0, n * 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("elements %zu, tiles %u, SMs %d, %d dependent DFMA per element\n", n, ntiles, sms, kFlops);
    timeit("ldg_tile (LDG.256 / STG.256, 1 tile per block)", n, (int)ntiles, [&] { ldg_tile<<<ntiles, kThreads>>>(in, o1, o2, ntiles); });
    CK(cudaFuncSetAttribute(tma_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Stage)));
    timeit("tma_tile (bulk copies, 1 tile per block)", n, (int)ntiles, [&] { tma_tile<<<ntiles, kThreads, sizeof(Stage)>>>(in, o1, o2, ntiles); });
    CK(cudaFuncSetAttribute(tma_persistent<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(Stage))));
    CK(cudaFuncSetAttribute(tma_persistent<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * sizeof(Stage))));
    CK(cudaFuncSetAttribute(tma_persistent<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(Stage))));
    for (int k : {1, 2, 3, 4}) {
        const int g = sms * k;
        char name[96];
        if (k * 2 * (int)sizeof(Stage) <= 220 * 1024) {
            snprintf(name, sizeof name, "tma_persistent 2 stages, %d blocks/SM", k);
            timeit(name, n, g, [&] { tma_persistent<2><<<g, kThreads, 2 * sizeof(Stage)>>>(in, o1, o2, ntiles); });
        }
        if (k * 3 * (int)sizeof(Stage) <= 220 * 1024) {
            snprintf(name, sizeof name, "tma_persistent 3 stages, %d blocks/SM", k);
            timeit(name, n, g, [&] { tma_persistent<3><<<g, kThreads, 3 * sizeof(Stage)>>>(in, o1, o2, ntiles); });
        }
        if (k * 4 * (int)sizeof(Stage) <= 220 * 1024) {
            snprintf(name, sizeof name, "tma_persistent 4 stages, %d blocks/SM", k);
            timeit(name, n, g, [&] { tma_persistent<4><<<g, kThreads, 4 * sizeof(Stage)>>>(in, o1, o2, ntiles); });
        }
    }
    // verify one variant moved the right data
    CK(cudaMemset(o1, 0xFF, 4096));
    tma_persistent<3><<<sms, kThreads, 3 * sizeof(Stage)>>>(in, o1, o2, ntiles);
    CK(cudaDeviceSynchronize());
    double h[4];
    CK(cudaMemcpy(h, o1, sizeof h, cudaMemcpyDeviceToHost));
    printf("check: o1[0] = %.6f (expected the DFMA chain of 0.0)\n", h[0]);
    return 0;
}
