// fabe13_shim.cu -- the C ABI of libfabe13.so: the reference's fabe13.h entry points plus the CUDA extension
// (include/fabe13_cuda.h), over the CUDA runtime.
//
// Mirrors the reference's dispatcher layer (fabe13/fabe13.c:132-135 global state, :186-219 initialisation and
// getters, :222-259 fabe13_sincos, :262-269 scalar API) with the ISA switch replaced by: device pointers -> one
// stream-ordered launch sequence; host pointers -> chunked H2D | kernel | D2H pipeline, except that host-pointer calls
// below option "min_n" run the HOST INSTANTIATION of the same per-element core (fabe13_hostcore.cpp; same bits as the
// device) -- the counterpart of the reference's n < 32 scalar route (fabe13/fabe13.c:226-229): a kernel launch costs
// ~13 us, the host core ~2 ns per element.  A CUDA failure is recorded (sticky, per thread), reported on stderr and
// returned by every int entry point; the void fabe13_sincos, whose ABI has no error channel (fabe13/fabe13.h:51-56),
// then fills its outputs with the host core so that callers never read unwritten memory.
#include <cuda_runtime.h>
#include <math.h>
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <thread>
#include <vector>

#pragma GCC visibility push(default)  // the library is built with -fvisibility=hidden; only the C ABI is exported
#include "../../include/fabe13.h"
#include "../../include/fabe13_cuda.h"
#pragma GCC visibility pop
#include "fabe13_core.h"
#include "fabe13_hostcore.h"
#include "fabe13_internal.h"

namespace {

using fabe13::DeviceInfo;
using fabe13::LaunchTuning;

constexpr int kMaxDevices = 64;
constexpr int kSlots = 16;                                 // max staging workers (pageable callers), one slot each
constexpr int kPipeSlots = 8;                              // pipeline depth for pinned callers (copy engines only)
constexpr size_t kLaunchMaxElems = (size_t)1 << 30;        // per launch sequence; keeps 32-bit indices and alignment
constexpr size_t kZeroCopyPageableMax = 32768;             // pageable callers: bounce through one small pinned block
constexpr unsigned long long kPoolKeepBytes = 256ull << 20;  // scratch the stream-ordered pool keeps between calls

// ---------------------------------------------------------------------------------------------------------------
// configuration (reference: compile-time constants fabe13/fabe13.c:71-73; here run-time options)
// ---------------------------------------------------------------------------------------------------------------
struct Config {
    std::atomic<long long> core{FABE13_CORE_POLY};
    std::atomic<long long> worklist{0};
    // worklist = 0: arrays of at least this many elements would take the hybrid sequence (zero counters | stream |
    // second pass).  Default: never -- with sparse Payne-Hanek lanes reduced by their own lanes (direct_max) the single
    // launch wins on every mix measured, and costs 3 % less per call at the 8-GPU slice size (profiles/r02_tune_stragglers.md)
    std::atomic<long long> hybrid_min_n{(long long)1 << 31};
    std::atomic<long long> local_min{2};
    // first-element votes (of 32) that send a warp's tile down the dense route.  8 since the fast reduction made that
    // route cheap (round 1: 16): the every-range routine over all 128 elements now costs less than fast path + list from
    // about a fifth of huge arguments on (random 30 / 40 / 50 / 60 %: +5 / +11 / +10 / +4 %, nothing lost below;
    // profiles/r02_tune_dense_hint.md).  (The kernel file's own kDenseHint only stands in for a non-positive option.)
    std::atomic<long long> dense_hint{8};
    std::atomic<long long> direct_max{2};
    std::atomic<long long> blocks_per_sm{0};
    std::atomic<long long> tiles_per_block{1};
    std::atomic<long long> pdl{1};
    std::atomic<long long> second_pass_blocks_per_sm{8};
    std::atomic<long long> unroll{1};
    // arrays below this many elements take the one-element-per-thread kernel.  Default 0: the streaming kernel serves every
    // size -- replayed from a CUDA graph it is the faster one at 1000 elements (1.6 vs 1.8 us per call: its launches are
    // programmatic-dependent, so back-to-back calls overlap their launch latency) as at 1e6 (3.9 vs 7.0 us: 256-bit
    // accesses, a quarter of the blocks); only around 1e5 elements the other kernel's larger grid wins, by 7 %
    // (profiles/r02_graph_latency.md)
    std::atomic<long long> small_n{0};
    std::atomic<long long> worklist_shift{3};
    std::atomic<long long> chunk_elems{(long long)1 << 22};
    std::atomic<long long> host_devices{1};
    std::atomic<long long> stage_threads{0};  // 0 = auto: min(kSlots, 3/4 of the hardware threads)
    std::atomic<long long> zero_copy_max{(long long)1 << 24};
    std::atomic<long long> stream_copy{1};  // staging copies of the pageable path with non-temporal stores
    // host-pointer calls below these many elements stay on the host core: measured crossovers (profiles/r02_abi_latency.txt)
    // of one host thread against the zero-copy kernel (page-locked arrays: 20 us either way at 16k elements) and
    // against the staged path (pageable arrays: bounce copies + PCIe round trip, 360 vs 400 us at 256k elements, equal at 512k)
    std::atomic<long long> min_n{16384};
    std::atomic<long long> min_n_pageable{262144};
    std::atomic<long long> fallback{1};     // void fabe13_sincos(): on a CUDA failure fill the outputs on the host core
    std::atomic<long long> host_threads{0}; // threads of that fallback (0 = auto: the hardware threads, at most 16)
    // Large page-locked host arrays: this many host threads take pieces from the array's TAIL with the host core while
    // the copy-engine pipeline works through it from the front (0 = off, the default: a host-pointer call is then a
    // GPU call).  A host-resident array is bound by what the host's memory and the PCIe link carry, not by the kernel
    // (DESIGN.md 3.5); the host core produces the same bits, so where the two ranges meet makes no difference.
    std::atomic<long long> host_assist{0};
    std::atomic<long long> host_assist_min_n{(long long)1 << 24};
};

Config g_cfg;
std::once_flag g_cfg_once;
std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_host_core_calls{0};  // array calls served by the host core (small n)
std::atomic<unsigned long long> g_fallbacks{0};        // fabe13_sincos calls completed on the host core after a CUDA failure
std::atomic<unsigned long long> g_assist_host_elems{0};  // option host_assist: elements the host threads took (all calls)
std::atomic<unsigned long long> g_assist_gpu_elems{0};   // ... and elements the GPU pipeline took in those calls
std::atomic<int> g_ndev{-1};                           // visible CUDA devices, asked once (-1 = not asked yet)
thread_local int tl_last_error = 0;                    // per thread: clear / call / check needs no lock

struct OptionDesc {
    const char* key;
    const char* env;
    std::atomic<long long> Config::*field;
    long long lo, hi;
};

const OptionDesc kOptions[] = {
    {"core", "FABE13_CUDA_CORE", &Config::core, 0, 1},
    {"worklist", "FABE13_CUDA_WORKLIST", &Config::worklist, 0, 3},
    {"hybrid_min_n", "FABE13_CUDA_HYBRID_MIN_N", &Config::hybrid_min_n, 0, (long long)1 << 31},
    {"local_min", "FABE13_CUDA_LOCAL_MIN", &Config::local_min, 1, 129},
    {"dense_hint", "FABE13_CUDA_DENSE_HINT", &Config::dense_hint, 1, 33},
    {"direct_max", "FABE13_CUDA_DIRECT_MAX", &Config::direct_max, 0, 128},
    {"blocks_per_sm", "FABE13_CUDA_BLOCKS_PER_SM", &Config::blocks_per_sm, 0, 64},
    {"tiles_per_block", "FABE13_CUDA_TILES_PER_BLOCK", &Config::tiles_per_block, 1, 1024},
    {"pdl", "FABE13_CUDA_PDL", &Config::pdl, 0, 1},
    {"second_pass_blocks_per_sm", "FABE13_CUDA_SECOND_PASS_BLOCKS_PER_SM", &Config::second_pass_blocks_per_sm, 1, 32},
    {"unroll", "FABE13_CUDA_UNROLL", &Config::unroll, 1, 2},
    {"small_n", "FABE13_CUDA_SMALL_N", &Config::small_n, 0, (long long)1 << 30},
    {"worklist_shift", "FABE13_CUDA_WORKLIST_SHIFT", &Config::worklist_shift, 0, 20},
    {"chunk_elems", "FABE13_CUDA_CHUNK_ELEMS", &Config::chunk_elems, 1024, (long long)1 << 28},
    {"host_devices", "FABE13_CUDA_HOST_DEVICES", &Config::host_devices, 1, kMaxDevices},
    {"stage_threads", "FABE13_CUDA_STAGE_THREADS", &Config::stage_threads, 0, kSlots},
    {"zero_copy_max", "FABE13_CUDA_ZERO_COPY_MAX", &Config::zero_copy_max, 0, (long long)1 << 30},
    {"stream_copy", "FABE13_CUDA_STREAM_COPY", &Config::stream_copy, 0, 1},
    {"min_n", "FABE13_CUDA_MIN_N", &Config::min_n, 0, (long long)1 << 31},
    {"min_n_pageable", "FABE13_CUDA_MIN_N_PAGEABLE", &Config::min_n_pageable, 0, (long long)1 << 31},
    {"fallback", "FABE13_CUDA_FALLBACK", &Config::fallback, 0, 1},
    {"host_threads", "FABE13_CUDA_HOST_THREADS", &Config::host_threads, 0, 256},
    {"host_assist", "FABE13_CUDA_HOST_ASSIST", &Config::host_assist, 0, 256},
    {"host_assist_min_n", "FABE13_CUDA_HOST_ASSIST_MIN_N", &Config::host_assist_min_n, (long long)1 << 20, (long long)1 << 40},
};

void load_env() {
    std::call_once(g_cfg_once, [] {
        for (const OptionDesc& o : kOptions) {
            const char* v = getenv(o.env);
            if (!v || !*v) continue;
            long long x;
            if (strcmp(o.key, "core") == 0 && (strcmp(v, "psi") == 0 || strcmp(v, "poly") == 0 || strcmp(v, "estrin") == 0))
                x = strcmp(v, "psi") == 0 ? 1 : 0;
            else
                x = atoll(v);
            if (x >= o.lo && x <= o.hi) (g_cfg.*(o.field)).store(x);
        }
    });
}

LaunchTuning current_tuning() {
    load_env();
    LaunchTuning t;
    t.worklist = (int)g_cfg.worklist.load();
    t.hybrid_min_n = (int)(g_cfg.hybrid_min_n.load() > 0x7FFFFFFFll ? 0x7FFFFFFFll : g_cfg.hybrid_min_n.load());
    t.local_min = (int)g_cfg.local_min.load();
    t.dense_hint = (int)g_cfg.dense_hint.load();
    t.direct_max = (int)g_cfg.direct_max.load();
    t.blocks_per_sm = (int)g_cfg.blocks_per_sm.load();
    t.tiles_per_block = (int)g_cfg.tiles_per_block.load();
    t.pdl = (int)g_cfg.pdl.load();
    t.second_pass_blocks_per_sm = (int)g_cfg.second_pass_blocks_per_sm.load();
    t.unroll = (int)g_cfg.unroll.load();
    t.small_n = (int)g_cfg.small_n.load();
    t.worklist_shift = (int)g_cfg.worklist_shift.load();
    return t;
}

int record(cudaError_t e, const char* where) {
    if (e == cudaSuccess) return 0;
    tl_last_error = (int)e;
    fprintf(stderr, "fabe13: CUDA error %d (%s) in %s\n", (int)e, cudaGetErrorString(e), where);
    const char* ab = getenv("FABE13_CUDA_ABORT");
    if (ab && *ab == '1') abort();
    (void)cudaGetLastError();  // clear the non-sticky error state so later calls are not poisoned
    return (int)e;
}

// an error code reported by a worker thread (recorded in ITS thread-local slot) becomes this thread's last error too
int adopt(int rc) {
    if (rc != 0) tl_last_error = rc;
    return rc;
}

#define FABE13_TRY(expr)                               \
    do {                                               \
        cudaError_t e__ = (expr);                      \
        if (e__ != cudaSuccess) return record(e__, #expr); \
    } while (0)

// The C ABI must not let a C++ exception (std::system_error from std::thread, std::bad_alloc) cross into C callers.
template <typename Fn>
int guarded(const char* who, Fn fn) {
    try {
        return fn();
    } catch (const std::bad_alloc&) {
        return record(cudaErrorMemoryAllocation, who);
    } catch (...) {
        return record(cudaErrorUnknown, who);
    }
}

// visible CUDA devices; asked once (a process does not gain or lose GPUs), so a box without a GPU never re-enters
// the driver's failing initialisation on later calls
int device_count() {
    int n = g_ndev.load(std::memory_order_acquire);
    if (n >= 0) return n;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
    }
    g_ndev.store(n, std::memory_order_release);
    return n;
}

// NVTX ranges (tracing): no-ops unless a tool is attached
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range&) = delete;
    Range& operator=(const Range&) = delete;
};

// results of large arrays leave the host core with non-temporal stores (-1: by size) unless option stream_copy is off
int host_stream_stores() { return g_cfg.stream_copy.load(std::memory_order_relaxed) != 0 ? -1 : 0; }

int host_core_threads() {
    long long t = g_cfg.host_threads.load();
    if (t <= 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        t = hw ? (hw > 16 ? 16 : hw) : 1;
    }
    return (int)t;
}

// ---------------------------------------------------------------------------------------------------------------
// per-device state, created on first use (reference: load-time constructor fabe13/fabe13.c:209-214)
// ---------------------------------------------------------------------------------------------------------------
struct Slot {
    cudaStream_t stream = nullptr;
    double* d_in = nullptr;   // cap elements each
    double* d_sin = nullptr;
    double* d_cos = nullptr;
    void* d_ws = nullptr;
    double* h_in = nullptr;   // pinned staging, allocated only when a pageable buffer shows up
    double* h_sin = nullptr;
    double* h_cos = nullptr;
    size_t cap = 0;        // elements per device buffer
    size_t ws_bytes = 0;   // bytes of worklist scratch
    size_t stage_cap = 0;  // elements per pinned staging buffer
};

struct DeviceState {
    std::mutex init_mu;
    bool ready = false;
    DeviceInfo info{};
    cudaMemPool_t pool = nullptr;  // stream-ordered scratch for the device-pointer entry points
    char name[320] = {0};
    // host path.  Locking: a host call holds pipe_rw SHARED while its chunks are in flight, so several callers share
    // the slots (their chunks interleave on the slots' streams; stream order keeps a slot's device buffers consistent
    // as long as one chunk's copies and kernels are enqueued atomically: slot_mu).  Resizing or freeing slots takes
    // pipe_rw EXCLUSIVE, i.e. waits until no call is in flight.  The pinned staging buffers of the pageable path and
    // the small-call bounce buffer have one user at a time (stage_mu, small_mu).
    std::shared_mutex pipe_rw;
    std::mutex slot_mu[kSlots];
    std::mutex stage_mu;
    std::mutex small_mu;
    std::mutex ev_mu;
    std::vector<cudaEvent_t> ev_free;     // blocking-sync events: waiting host threads sleep instead of spinning
    cudaStream_t small_stream = nullptr;  // small host calls: one kernel reading/writing pinned host memory directly
    double* small_stage = nullptr;        // pinned, 3 * small_cap doubles (in | sin | cos), for pageable callers
    size_t small_cap = 0;
    Slot slots[kSlots];
};

DeviceState g_dev[kMaxDevices];

int device_state(int dev, DeviceState** out) {
    if (dev < 0 || dev >= kMaxDevices) return record(cudaErrorInvalidDevice, "device_state");
    DeviceState& s = g_dev[dev];
    std::lock_guard<std::mutex> lk(s.init_mu);
    if (!s.ready) {
        cudaDeviceProp prop;
        FABE13_TRY(cudaGetDeviceProperties(&prop, dev));
        s.info.device = dev;
        s.info.sm_count = prop.multiProcessorCount;
        snprintf(s.name, sizeof s.name, "CUDA sm_%d%d (%s, %d SMs)", prop.major, prop.minor, prop.name,
                 prop.multiProcessorCount);
        cudaMemPoolProps pp;
        memset(&pp, 0, sizeof pp);
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = dev;
        FABE13_TRY(cudaMemPoolCreate(&s.pool, &pp));
        // the worklist scratch is reused call after call (125 MB for a 1e9-element launch): keep up to two of those
        // between calls and give the rest back, so a GPU shared with another allocator is not drained
        unsigned long long keep = kPoolKeepBytes;
        FABE13_TRY(cudaMemPoolSetAttribute(s.pool, cudaMemPoolAttrReleaseThreshold, &keep));
        s.ready = true;
    }
    *out = &s;
    return 0;
}

int current_device_state(DeviceState** out) {
    int dev = 0;
    FABE13_TRY(cudaGetDevice(&dev));
    return device_state(dev, out);
}

// ---------------------------------------------------------------------------------------------------------------
// device-pointer path
// ---------------------------------------------------------------------------------------------------------------
int run_device(const double* d_in, double* d_sin, double* d_cos, long long* d_k, size_t n, void* user_ws,
               size_t user_ws_bytes, cudaStream_t stream, int op = FABE13_OP_SINCOS) {
    if (n == 0) return 0;
    DeviceState* ds;
    if (int rc = current_device_state(&ds)) return rc;
    const LaunchTuning t = current_tuning();
    const int core = (int)g_cfg.core.load();
    for (size_t off = 0; off < n; off += kLaunchMaxElems) {
        const size_t m = (n - off < kLaunchMaxElems) ? n - off : kLaunchMaxElems;
        void* ws = nullptr;
        bool pooled = false;
        const size_t need = m >= (size_t)t.small_n ? fabe13::workspace_bytes(m, t) : 0;
        if (need > 0) {
            if (user_ws) {
                if (user_ws_bytes < need) return record(cudaErrorInvalidValue, "workspace too small");
                ws = user_ws;
            } else {
                // no scratch to be had (device memory exhausted by the caller's own arrays): not an error -- without
                // a workspace the launch keeps every Payne-Hanek lane in the block (worklist design 2)
                if (cudaMallocFromPoolAsync(&ws, need, ds->pool, stream) == cudaSuccess) {
                    pooled = true;
                } else {
                    (void)cudaGetLastError();
                    ws = nullptr;
                }
            }
        }
        int launches = 0;
        cudaError_t e = fabe13::launch_sincos(d_in + off, d_sin ? d_sin + off : nullptr, d_cos ? d_cos + off : nullptr,
                                              d_k ? d_k + off : nullptr, m, op, core, t, ds->info, ws, stream, &launches);
        g_launches.fetch_add((unsigned long long)launches);
        if (pooled) {
            cudaError_t e2 = cudaFreeAsync(ws, stream);
            if (e == cudaSuccess) e = e2;
        }
        if (e != cudaSuccess) return record(e, "launch_sincos");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host-pointer path: chunked pipeline through per-device slots
// ---------------------------------------------------------------------------------------------------------------
// The staging copies of the pageable path are bound by the host's memory bandwidth, and neither destination is read
// by the CPU again soon (the staging buffer is read by the DMA engine, the caller's output array by the caller much
// later): non-temporal stores skip the read-for-ownership of every destination line, a quarter of the traffic.
// Option stream_copy = 0 restores memcpy.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void stream_copy_avx2(char* d, const char* s, size_t bytes) {  // d is 32-byte aligned
    size_t i = 0;
    for (; i + 128 <= bytes; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
    }
    _mm_sfence();  // the DMA engine (or another thread) must see the data once this returns
    if (i < bytes) memcpy(d + i, s + i, bytes - i);
}
#endif

void staging_copy(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2 && g_cfg.stream_copy.load(std::memory_order_relaxed) != 0 && bytes >= 4096) {
        char* d = static_cast<char*>(dst);
        const char* s = static_cast<const char*>(src);
        const size_t lead = (32 - ((uintptr_t)d & 31)) & 31;  // stores must be 32-byte aligned; loads need not be
        if (lead) {
            memcpy(d, s, lead);
            d += lead;
            s += lead;
            bytes -= lead;
        }
        stream_copy_avx2(d, s, bytes);
        return;
    }
#endif
    memcpy(dst, src, bytes);
}

enum PtrKind { PTR_PAGEABLE, PTR_PINNED, PTR_DEVICE };

PtrKind classify_byte(const void* p, ptrdiff_t* map_delta) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return PTR_PAGEABLE;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return PTR_DEVICE;
    if (a.type == cudaMemoryTypeHost) {
        if (map_delta) *map_delta = (const char*)a.devicePointer - (const char*)a.hostPointer;
        return PTR_PINNED;
    }
    return PTR_PAGEABLE;
}

// What kind of memory [p, p + bytes) is.  Pinned only if the LAST byte is page-locked too and maps to the device with
// the same offset as the first one: a pointer into a smaller cudaHostRegister region (or a sub-array reaching past the
// end of a pinned block) is treated as pageable and staged, instead of faulting inside a copy or a zero-copy kernel.
PtrKind classify(const void* p, size_t bytes) {
    ptrdiff_t d0 = 0, d1 = 0;
    const PtrKind k = classify_byte(p, &d0);
    if (k != PTR_PINNED || bytes <= 1) return k;
    const PtrKind k1 = classify_byte((const char*)p + bytes - 1, &d1);
    return (k1 == PTR_PINNED && d0 == d1) ? PTR_PINNED : PTR_PAGEABLE;
}

bool slots_fit(const DeviceState& s, size_t chunk, bool device_bufs, bool staging, size_t need_ws, int nslots) {
    for (int i = 0; i < nslots; ++i) {
        const Slot& sl = s.slots[i];
        if (!sl.stream) return false;
        if (device_bufs && sl.cap < chunk) return false;
        if (need_ws > 0 && sl.ws_bytes < need_ws) return false;
        if (staging && sl.stage_cap < chunk) return false;
    }
    return true;
}

// (re)size the first `nslots` slots of the current device for `chunk` elements; pinned staging only if asked for,
// device buffers only if asked for (the FP32 path works on its staging buffers in place).  Capacities are kept per
// slot: the pinned pipeline uses kPipeSlots large slots, the staging workers up to kSlots small ones.
// Caller holds pipe_rw exclusively: no call is in flight on any slot.
int grow_slots(DeviceState& s, size_t chunk, bool device_bufs, bool staging, size_t need_ws, int nslots) {
    for (int i = 0; i < nslots; ++i) {
        Slot& sl = s.slots[i];
        if (!sl.stream) FABE13_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if (device_bufs && sl.cap < chunk) {
            if (sl.d_in) FABE13_TRY(cudaFree(sl.d_in));
            if (sl.d_sin) FABE13_TRY(cudaFree(sl.d_sin));
            if (sl.d_cos) FABE13_TRY(cudaFree(sl.d_cos));
            sl.d_in = sl.d_sin = sl.d_cos = nullptr;
            sl.cap = 0;
            FABE13_TRY(cudaMalloc(&sl.d_in, chunk * sizeof(double)));
            FABE13_TRY(cudaMalloc(&sl.d_sin, chunk * sizeof(double)));
            FABE13_TRY(cudaMalloc(&sl.d_cos, chunk * sizeof(double)));
            sl.cap = chunk;
        }
        if (need_ws > 0 && sl.ws_bytes < need_ws) {
            if (sl.d_ws) FABE13_TRY(cudaFree(sl.d_ws));
            sl.d_ws = nullptr;
            sl.ws_bytes = 0;
            FABE13_TRY(cudaMalloc(&sl.d_ws, need_ws));
            sl.ws_bytes = need_ws;
        }
        if (staging && sl.stage_cap < chunk) {
            if (sl.h_in) cudaFreeHost(sl.h_in);
            if (sl.h_sin) cudaFreeHost(sl.h_sin);
            if (sl.h_cos) cudaFreeHost(sl.h_cos);
            sl.h_in = sl.h_sin = sl.h_cos = nullptr;
            sl.stage_cap = 0;
            FABE13_TRY(cudaHostAlloc(&sl.h_in, chunk * sizeof(double), cudaHostAllocDefault));
            FABE13_TRY(cudaHostAlloc(&sl.h_sin, chunk * sizeof(double), cudaHostAllocDefault));
            FABE13_TRY(cudaHostAlloc(&sl.h_cos, chunk * sizeof(double), cudaHostAllocDefault));
            sl.stage_cap = chunk;
        }
    }
    return 0;
}

// Returns with `lk` holding pipe_rw shared and the first nslots slots large enough.
int acquire_slots(DeviceState& s, std::shared_lock<std::shared_mutex>& lk, size_t chunk, bool device_bufs, bool staging,
                  size_t need_ws, int nslots) {
    for (;;) {
        lk.lock();
        if (slots_fit(s, chunk, device_bufs, staging, need_ws, nslots)) return 0;
        lk.unlock();
        std::unique_lock<std::shared_mutex> x(s.pipe_rw);  // waits for the calls in flight
        if (int rc = grow_slots(s, chunk, device_bufs, staging, need_ws, nslots)) return rc;
    }
}

int take_event(DeviceState& s, cudaEvent_t* ev) {
    {
        std::lock_guard<std::mutex> lk(s.ev_mu);
        if (!s.ev_free.empty()) {
            *ev = s.ev_free.back();
            s.ev_free.pop_back();
            return 0;
        }
    }
    FABE13_TRY(cudaEventCreateWithFlags(ev, cudaEventBlockingSync | cudaEventDisableTiming));
    return 0;
}

void give_event(DeviceState& s, cudaEvent_t ev) {
    std::lock_guard<std::mutex> lk(s.ev_mu);
    s.ev_free.push_back(ev);
}

size_t pick_chunk(size_t n) {
    const size_t cap = (size_t)g_cfg.chunk_elems.load();
    size_t c = (n + 5) / 6;  // at least ~6 chunks so the three stages overlap
    if (c < 131072) c = 131072;
    if (c > cap) c = cap;
    if (c > n) c = n;
    return (c + 3) & ~(size_t)3;
}

// enqueue one chunk on a slot: H2D, kernels, 2x D2H (then `done`, if given).  src/dst are pinned (user memory or
// staging).  The slot's mutex makes the sequence atomic with respect to other callers sharing the slot.
int enqueue_chunk(DeviceState& s, int slot, const double* src, double* dst_sin, double* dst_cos, size_t m,
                  const LaunchTuning& t, int core, int op, cudaEvent_t done) {
    Slot& sl = s.slots[slot];
    std::lock_guard<std::mutex> lk(s.slot_mu[slot]);
    {
        Range r("fabe13:h2d");
        FABE13_TRY(cudaMemcpyAsync(sl.d_in, src, m * sizeof(double), cudaMemcpyHostToDevice, sl.stream));
    }
    {
        Range r("fabe13:kernel");
        int launches = 0;
        cudaError_t e = fabe13::launch_sincos(sl.d_in, dst_sin ? sl.d_sin : nullptr, dst_cos ? sl.d_cos : nullptr, nullptr, m, op, core,
                                              t, s.info, (m >= (size_t)t.small_n && fabe13::workspace_bytes(m, t) > 0) ? sl.d_ws : nullptr, sl.stream, &launches);
        g_launches.fetch_add((unsigned long long)launches);
        if (e != cudaSuccess) return record(e, "launch_sincos(host path)");
    }
    Range r("fabe13:d2h");
    if (dst_sin) FABE13_TRY(cudaMemcpyAsync(dst_sin, sl.d_sin, m * sizeof(double), cudaMemcpyDeviceToHost, sl.stream));
    if (dst_cos) FABE13_TRY(cudaMemcpyAsync(dst_cos, sl.d_cos, m * sizeof(double), cudaMemcpyDeviceToHost, sl.stream));
    if (done) FABE13_TRY(cudaEventRecord(done, sl.stream));
    return 0;
}

// Small and medium host-resident arrays: skip the copy engines.  The kernels read the pinned host buffer over PCIe
// and write both results straight back (zero-copy through the unified address space): no pipeline fill/drain, one
// launch for small n.  Small pageable calls are bounced through a pinned staging block.  Returns -1 if this route is not available
// (memory registered without a device mapping), in which case the caller takes the copy pipeline.
int host_small(DeviceState& s, const double* in, double* sin_out, double* cos_out, size_t n, bool pinned,
               const LaunchTuning& t, int core, int op) {
    Range range("fabe13:zero_copy");
    std::lock_guard<std::mutex> lk(s.small_mu);
    if (!s.small_stream) FABE13_TRY(cudaStreamCreateWithFlags(&s.small_stream, cudaStreamNonBlocking));
    const double* src = in;
    double *dst_s = sin_out, *dst_c = cos_out;  // either may be null (single-output calls)
    if (!pinned) {
        if (s.small_cap < n) {
            if (s.small_stage) cudaFreeHost(s.small_stage);
            s.small_stage = nullptr;
            s.small_cap = 0;
            const size_t cap = kZeroCopyPageableMax > n ? kZeroCopyPageableMax : n;
            FABE13_TRY(cudaHostAlloc(&s.small_stage, 3 * cap * sizeof(double), cudaHostAllocDefault));
            s.small_cap = cap;
        }
        memcpy(s.small_stage, in, n * sizeof(double));
        src = s.small_stage;
        if (sin_out) dst_s = s.small_stage + s.small_cap;
        if (cos_out) dst_c = s.small_stage + 2 * s.small_cap;
    }
    void *d_in = nullptr, *d_s = nullptr, *d_c = nullptr;
    if (cudaHostGetDevicePointer(&d_in, (void*)src, 0) != cudaSuccess ||
        (dst_s && cudaHostGetDevicePointer(&d_s, dst_s, 0) != cudaSuccess) ||
        (dst_c && cudaHostGetDevicePointer(&d_c, dst_c, 0) != cudaSuccess)) {
        (void)cudaGetLastError();
        return -1;
    }
    int launches = 0;
    void* ws = nullptr;
    const size_t need_ws = n >= (size_t)t.small_n ? fabe13::workspace_bytes(n, t) : 0;
    if (need_ws > 0 && cudaMallocFromPoolAsync(&ws, need_ws, s.pool, s.small_stream) != cudaSuccess) {
        (void)cudaGetLastError();  // no scratch: the launch keeps every Payne-Hanek lane in the block
        ws = nullptr;
    }
    cudaError_t e = fabe13::launch_sincos((const double*)d_in, (double*)d_s, (double*)d_c, nullptr, n, op, core, t, s.info, ws,
                                          s.small_stream, &launches);
    g_launches.fetch_add((unsigned long long)launches);
    if (ws) {
        cudaError_t e2 = cudaFreeAsync(ws, s.small_stream);
        if (e == cudaSuccess) e = e2;
    }
    if (e != cudaSuccess) return record(e, "launch_sincos(zero-copy)");
    FABE13_TRY(cudaStreamSynchronize(s.small_stream));
    if (!pinned) {
        if (sin_out) memcpy(sin_out, dst_s, n * sizeof(double));
        if (cos_out) memcpy(cos_out, dst_c, n * sizeof(double));
    }
    return 0;
}

int default_stage_workers() {
    int w = (int)g_cfg.stage_threads.load();
    if (w <= 0) {
        // the staging copies are what bounds this path (measured: throughput still grows from 8 to 12 threads on
        // a 16-thread host), so take most of the machine; the caller's own thread is worker 0
        const unsigned hw = std::thread::hardware_concurrency();
        const unsigned want = hw - hw / 4;
        w = (int)(want > (unsigned)kSlots ? (unsigned)kSlots : (want ? want : 1));
    }
    return w;
}

// run work(w) for w = 0 .. workers-1, worker 0 on the calling thread
template <typename Fn>
void run_workers(int workers, Fn work) {
    if (workers <= 1) {
        work(0);
        return;
    }
    // (a thread that cannot be started must not take the process down -- a joinable std::thread destroyed during
    // unwinding calls std::terminate -- nor leave its share of the chunks undone: the calling thread runs it)
    std::vector<std::thread> th;
    th.reserve((size_t)workers);
    int started = 1;
    try {
        for (; started < workers; ++started) th.emplace_back(work, started);
    } catch (...) {
    }
    work(0);
    for (int w = started; w < workers; ++w) work(w);
    for (std::thread& x : th) x.join();
}


// Option host_assist: a large page-locked array worked on from both ends.  The calling thread feeds the copy-engine
// pipeline with chunks from the FRONT of the array, `helpers` host threads run the host core on pieces from the BACK;
// both sides claim their next piece from one atomic word (front and back cursor, in units of kAssistUnit elements),
// so the split follows the speeds of the two sides as they are during this call -- no tuning, no history.  The feeder
// keeps at most kAssistInFlight chunks queued (it polls the oldest chunk's event: no interrupt per chunk, a sleeping
// poll every 50 us on 1.3 ms chunks) -- claims must be lazy, or the GPU side would have taken the whole array before
// the first host thread starts.  Results are the same bits whichever side computes an element.
constexpr size_t kAssistUnit = (size_t)1 << 18;  // elements per host claim (2 MB per array)
constexpr int kAssistInFlight = 6;               // chunks queued on the GPU side (three stages, two deep)

int host_slice_assisted(DeviceState& s, const double* in, double* sin_out, double* cos_out, size_t n, int op, int helpers,
                        const LaunchTuning& t, int core) {
    Range range("fabe13:assisted_pipeline");
    size_t chunk = pick_chunk(n);
    chunk = (chunk + kAssistUnit - 1) / kAssistUnit * kAssistUnit;  // whole units
    const uint32_t units_per_chunk = (uint32_t)(chunk / kAssistUnit);
    const uint32_t units = (uint32_t)((n + kAssistUnit - 1) / kAssistUnit);
    const size_t need_ws = fabe13::workspace_bytes(chunk, t);
    std::shared_lock<std::shared_mutex> lk(s.pipe_rw, std::defer_lock);
    if (int rc = acquire_slots(s, lk, chunk, true, false, need_ws, kPipeSlots)) return rc;

    std::atomic<uint64_t> cursors{(uint64_t)units << 32};  // low word: first unit not yet taken from the front; high word: end of the untaken range
    std::atomic<int> stop{0};
    std::atomic<unsigned long long> host_elems{0};
    auto claim_front = [&](uint32_t want, uint32_t* begin) -> uint32_t {
        uint64_t c = cursors.load(std::memory_order_relaxed);
        for (;;) {
            const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
            if (lo >= hi) return 0;
            const uint32_t take = hi - lo < want ? hi - lo : want;
            if (cursors.compare_exchange_weak(c, ((uint64_t)hi << 32) | (lo + take), std::memory_order_relaxed)) {
                *begin = lo;
                return take;
            }
        }
    };
    auto claim_back = [&](uint32_t* begin) -> bool {
        uint64_t c = cursors.load(std::memory_order_relaxed);
        for (;;) {
            const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
            if (lo >= hi) return false;
            if (cursors.compare_exchange_weak(c, ((uint64_t)(hi - 1) << 32) | lo, std::memory_order_relaxed)) {
                *begin = hi - 1;
                return true;
            }
        }
    };
    auto helper = [&] {
        Range r("fabe13:host_assist");
        uint32_t u;
        unsigned long long mine = 0;
        while (stop.load(std::memory_order_relaxed) == 0 && claim_back(&u)) {
            const size_t off = (size_t)u * kAssistUnit;
            const size_t m = n - off < kAssistUnit ? n - off : kAssistUnit;
            fabe13::host_sincos_array(in + off, sin_out ? sin_out + off : nullptr, cos_out ? cos_out + off : nullptr, m, op, core, 1,
                                      host_stream_stores() ? 1 : 0);
            mine += m;
        }
        host_elems.fetch_add(mine, std::memory_order_relaxed);
    };
    std::vector<std::thread> th;
    th.reserve((size_t)helpers);
    try {
        for (int i = 0; i < helpers; ++i) th.emplace_back(helper);
    } catch (...) {  // fewer helpers than asked for: the array is finished by those that did start and by the pipeline
    }

    // the feeder: this thread
    cudaEvent_t ev[kAssistInFlight] = {};
    int rc = 0;
    unsigned long long gpu_elems = 0;
    for (int i = 0; i < kAssistInFlight && rc == 0; ++i) {
        cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) rc = record(e, "cudaEventCreateWithFlags(assist)");
    }
    auto wait_event = [&](cudaEvent_t e) -> int {
        for (;;) {
            const cudaError_t q = cudaEventQuery(e);
            if (q == cudaSuccess) return 0;
            if (q != cudaErrorNotReady) return record(q, "cudaEventQuery(assist)");
            (void)cudaGetLastError();  // "not ready" is no error: do not leave it for the next launch's error check to find
            struct timespec ts = {0, 50000};
            nanosleep(&ts, nullptr);
        }
    };
    size_t queued = 0;  // chunks enqueued so far; chunk i waits on ev[i % kAssistInFlight]
    while (rc == 0) {
        if (queued >= (size_t)kAssistInFlight) rc = wait_event(ev[queued % kAssistInFlight]);  // the oldest chunk in flight
        if (rc != 0) break;
        uint32_t u0;
        const uint32_t got = claim_front(units_per_chunk, &u0);
        if (got == 0) break;
        const size_t off = (size_t)u0 * kAssistUnit;
        size_t m = (size_t)got * kAssistUnit;
        if (off + m > n) m = n - off;
        rc = enqueue_chunk(s, (int)(queued % kPipeSlots), in + off, sin_out ? sin_out + off : nullptr, cos_out ? cos_out + off : nullptr, m, t,
                           core, op, ev[queued % kAssistInFlight]);
        gpu_elems += m;
        ++queued;
    }
    if (rc != 0) stop.store(1);
    // (also after a failure: chunks already queued are still copying into the caller's arrays)
    for (int i = 0; i < kPipeSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(s.slots[i].stream);
        if (e != cudaSuccess && rc == 0) rc = record(e, "wait for the pipeline (assist)");
    }
    for (std::thread& x : th) x.join();
    for (cudaEvent_t e : ev)
        if (e) (void)cudaEventDestroy(e);
    if (rc != 0) (void)cudaGetLastError();
    g_assist_host_elems.fetch_add(host_elems.load());
    g_assist_gpu_elems.fetch_add(gpu_elems);
    return rc;
}

// One device, one contiguous slice.  Caller has made `dev` current.  `helpers` > 0: option host_assist for this slice.
int host_slice(int dev, const double* in, double* sin_out, double* cos_out, size_t n, bool pinned, int op, int helpers = 0) {
    DeviceState* ds;
    if (int rc = device_state(dev, &ds)) return rc;
    const LaunchTuning t = current_tuning();
    const int core = (int)g_cfg.core.load();
    // one kernel working on host memory directly: up to zero_copy_max elements for pinned arrays (faster than the
    // copy pipeline up to ~1e7 elements, measured), small calls only for pageable ones (they need a bounce buffer)
    const size_t zc_max = (size_t)g_cfg.zero_copy_max.load();
    if (n <= (pinned ? zc_max : (zc_max < kZeroCopyPageableMax ? zc_max : kZeroCopyPageableMax))) {
        const int rc = host_small(*ds, in, sin_out, cos_out, n, pinned, t, core, op);
        if (rc >= 0) return rc;
    }
    if (pinned && helpers > 0 && n >= (size_t)g_cfg.host_assist_min_n.load(std::memory_order_relaxed))
        return host_slice_assisted(*ds, in, sin_out, cos_out, n, op, helpers, t, core);
    size_t chunk = pick_chunk(n);
    int stage_workers = 0;
    if (!pinned) {
        // staged: pieces of at most 8 MB, and small enough that every worker gets about three of them (a worker handles
        // its pieces one after the other -- copy in, run, copy out -- so it is the other workers that overlap its stages)
        stage_workers = default_stage_workers();
        chunk = n / ((size_t)stage_workers * 3);
        if (chunk > ((size_t)1 << 20)) chunk = (size_t)1 << 20;
        if (chunk < ((size_t)1 << 16)) chunk = (size_t)1 << 16;
        chunk = (chunk + 3) & ~(size_t)3;
    }
    const size_t nchunks = (n + chunk - 1) / chunk;
    const size_t need_ws = fabe13::workspace_bytes(chunk, t);  // depends on the tuning in force
    const bool sleep_wait = n >= ((size_t)1 << 22);  // long waits: sleep on a blocking event; short ones: spin (lower latency)

    if (pinned) {
        Range range("fabe13:pinned_pipeline");
        std::shared_lock<std::shared_mutex> lk(ds->pipe_rw, std::defer_lock);
        if (int rc = acquire_slots(*ds, lk, chunk, true, false, need_ws, kPipeSlots)) return rc;
        // copy engines read and write the caller's pinned memory directly; one host thread feeds all slots
        cudaEvent_t ev[kPipeSlots] = {};
        const int used = nchunks < (size_t)kPipeSlots ? (int)nchunks : kPipeSlots;
        int rc = 0;
        if (sleep_wait)
            for (int i = 0; i < used && rc == 0; ++i) rc = take_event(*ds, &ev[i]);
        for (size_t i = 0; i < nchunks && rc == 0; ++i) {
            const size_t off = i * chunk;
            const size_t m = n - off < chunk ? n - off : chunk;
            const int slot = (int)(i % kPipeSlots);
            // the completion event goes behind the slot's LAST chunk only (one record per slot and call: a blocking-sync
            // event raises an interrupt when it completes, which is not something to put behind every copy)
            const bool last_on_slot = i + kPipeSlots >= nchunks;
            rc = enqueue_chunk(*ds, slot, in + off, sin_out ? sin_out + off : nullptr, cos_out ? cos_out + off : nullptr, m,
                               t, core, op, last_on_slot ? ev[slot] : nullptr);
        }
        // (also after a failure: earlier chunks are still copying into the caller's arrays -- do not hand the arrays
        // back mid-flight)
        for (int i = 0; i < used; ++i) {
            cudaError_t e = (sleep_wait && ev[i] && rc == 0) ? cudaEventSynchronize(ev[i]) : cudaStreamSynchronize(ds->slots[i].stream);
            if (e != cudaSuccess && rc == 0) rc = record(e, "wait for the pipeline");
        }
        for (int i = 0; i < used; ++i)
            if (ev[i]) give_event(*ds, ev[i]);
        if (rc != 0) (void)cudaGetLastError();
        return rc;
    }

    // pageable memory: worker w owns slot w and the chunks i = w (mod workers); it copies the chunk into pinned
    // staging, runs it, waits, and copies the results out.  Workers overlap each other's stages.
    Range range("fabe13:staged_pipeline");
    int workers = stage_workers;
    if ((size_t)workers > nchunks) workers = (int)nchunks;
    std::lock_guard<std::mutex> stage_lk(ds->stage_mu);
    std::shared_lock<std::shared_mutex> lk(ds->pipe_rw, std::defer_lock);
    if (int rc = acquire_slots(*ds, lk, chunk, true, true, need_ws, workers)) return rc;
    std::atomic<int> rc_all{0};
    auto work = [&](int w) {
        if (cudaSetDevice(dev) != cudaSuccess) {
            rc_all.store(record(cudaErrorInvalidDevice, "cudaSetDevice(worker)"));
            return;
        }
        Slot& sl = ds->slots[w];
        cudaEvent_t ev = nullptr;
        if (sleep_wait && take_event(*ds, &ev) != 0) {
            rc_all.store(tl_last_error);
            return;
        }
        for (size_t i = (size_t)w; i < nchunks && rc_all.load() == 0; i += (size_t)workers) {
            const size_t off = i * chunk;
            const size_t m = n - off < chunk ? n - off : chunk;
            {
                Range r("fabe13:stage_in");
                staging_copy(sl.h_in, in + off, m * sizeof(double));
            }
            int rc = enqueue_chunk(*ds, w, sl.h_in, sin_out ? sl.h_sin : nullptr, cos_out ? sl.h_cos : nullptr, m, t, core, op, ev);
            if (rc == 0) {
                const cudaError_t e = ev ? cudaEventSynchronize(ev) : cudaStreamSynchronize(sl.stream);
                if (e != cudaSuccess) rc = record(e, "wait for chunk (worker)");
            }
            if (rc != 0) {
                rc_all.store(rc);
                break;
            }
            Range r("fabe13:stage_out");
            if (sin_out) staging_copy(sin_out + off, sl.h_sin, m * sizeof(double));
            if (cos_out) staging_copy(cos_out + off, sl.h_cos, m * sizeof(double));
        }
        if (ev) give_event(*ds, ev);
    };
    run_workers(workers, work);
    return adopt(rc_all.load());
}

// host-pointer calls below "min_n" elements: the host instantiation of the per-element core (one thread, no CUDA call
// beyond finding out that the arrays are not device memory: a device pointer, or a mix of host and device pointers,
// takes the ordinary route and is served or refused there).  One cudaPointerGetAttributes costs ~0.1 us -- as much as
// the arithmetic of 80 elements -- so calls of fewer than 256 elements ask about `in` only (0.15 us at n = 10 instead
// of 0.35); from 256 elements on every array is asked about.  Reference: the n < 32 scalar route, fabe13/fabe13.c:226-229.
constexpr size_t kClassifyAllFrom = 256;
bool stays_on_host(const void* in, const void* out0, const void* out1, size_t n) {
    if (n >= (size_t)g_cfg.min_n.load(std::memory_order_relaxed)) return false;
    if (device_count() == 0) return true;
    if (classify_byte(in, nullptr) == PTR_DEVICE) return false;
    if (n < kClassifyAllFrom) return true;
    return (!out0 || classify_byte(out0, nullptr) != PTR_DEVICE) && (!out1 || classify_byte(out1, nullptr) != PTR_DEVICE);
}

int run_host(const double* in, double* sin_out, double* cos_out, size_t n, int op = FABE13_OP_SINCOS) {
    if (n == 0) return 0;
    load_env();
    if (!in || (!sin_out && !cos_out)) return record(cudaErrorInvalidValue, "null array");
    if (stays_on_host(in, sin_out, cos_out, n)) {
        fabe13::host_sincos_array(in, sin_out, cos_out, n, op, (int)g_cfg.core.load(), 1);
        g_host_core_calls.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    if (device_count() == 0) {  // (without a device every array is pageable host memory)
        if (n >= (size_t)g_cfg.min_n_pageable.load(std::memory_order_relaxed)) return record(cudaErrorNoDevice, "no CUDA device is visible");
        fabe13::host_sincos_array(in, sin_out, cos_out, n, op, (int)g_cfg.core.load(), host_core_threads(), host_stream_stores());
        g_host_core_calls.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    Range range("fabe13_sincos_host");
    const size_t bytes = n * sizeof(double);
    const PtrKind ki = classify(in, bytes);
    const PtrKind ks = sin_out ? classify(sin_out, bytes) : classify(cos_out, bytes);
    const PtrKind kc = cos_out ? classify(cos_out, bytes) : ks;
    if (ki == PTR_DEVICE || ks == PTR_DEVICE || kc == PTR_DEVICE) {
        if (!(ki == PTR_DEVICE && ks == PTR_DEVICE && kc == PTR_DEVICE))
            return record(cudaErrorInvalidValue, "fabe13_sincos: mixed host and device pointers");
        if (int rc = run_device(in, sin_out, cos_out, nullptr, n, nullptr, 0, nullptr, op)) return rc;
        FABE13_TRY(cudaStreamSynchronize(nullptr));
        return 0;
    }
    const bool pinned = (ki == PTR_PINNED && ks == PTR_PINNED && kc == PTR_PINNED);
    if (!pinned && n < (size_t)g_cfg.min_n_pageable.load(std::memory_order_relaxed)) {
        // pageable arrays pay two extra host copies and a PCIe round trip: the host core is faster up to here (the
        // default cut is the crossover of ONE host thread; from 131072 elements on the array is split over up to
        // "host_threads" threads, 65536 elements each at least -- which is also what a caller gets who raises the cut
        // to keep pageable arrays of every size on the CPU: see DESIGN.md 3.5 for when that is the better choice)
        fabe13::host_sincos_array(in, sin_out, cos_out, n, op, (int)g_cfg.core.load(), host_core_threads(), host_stream_stores());
        g_host_core_calls.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    int cur = 0;
    FABE13_TRY(cudaGetDevice(&cur));
    const int ndev_avail = device_count();
    int ndev = (int)g_cfg.host_devices.load();
    if (ndev > ndev_avail) ndev = ndev_avail;
    const int helpers = (int)g_cfg.host_assist.load(std::memory_order_relaxed);
    if (ndev <= 1 || n < ((size_t)1 << 20)) return host_slice(cur, in, sin_out, cos_out, n, pinned, op, helpers);

    // contiguous slice per GPU, one host thread each; no exchange between slices
    std::vector<std::thread> th;
    th.reserve((size_t)ndev);
    std::vector<int> rcs((size_t)ndev, 0);
    auto slice = [&](int g) {
        size_t b, e;
        fabe13_shard_bounds(n, ndev, g, &b, &e);
        const int dev = (cur + g) % ndev_avail;
        if (cudaSetDevice(dev) != cudaSuccess) {
            rcs[(size_t)g] = record(cudaErrorInvalidDevice, "cudaSetDevice(slice)");
            return;
        }
        if (e > b)
            rcs[(size_t)g] = host_slice(dev, in + b, sin_out ? sin_out + b : nullptr, cos_out ? cos_out + b : nullptr, e - b, pinned, op,
                                        helpers / ndev);
    };
    int started = 0;
    try {
        for (; started < ndev; ++started) th.emplace_back(slice, started);
    } catch (...) {
    }
    for (std::thread& x : th) x.join();
    for (int g = started; g < ndev; ++g) slice(g);  // threads that could not be started: their slices run here, one by one
    (void)cudaSetDevice(cur);
    for (int rc : rcs)
        if (rc) return adopt(rc);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// FP32 arrays.  Device pointers: one launch.  Pinned (or registered) host arrays: the kernel reads and writes them in
// place over PCIe (12 B/element either way).  Pageable host arrays: 4M-element pieces through a pinned bounce buffer.
// ---------------------------------------------------------------------------------------------------------------
int run_device_f32(const float* d_in, float* d_sin, float* d_cos, size_t n, cudaStream_t stream) {
    if (n == 0) return 0;
    DeviceState* ds;
    if (int rc = current_device_state(&ds)) return rc;
    const int core = (int)g_cfg.core.load();
    for (size_t off = 0; off < n; off += kLaunchMaxElems) {
        const size_t m = (n - off < kLaunchMaxElems) ? n - off : kLaunchMaxElems;
        int launches = 0;
        cudaError_t e = fabe13::launch_sincosf(d_in + off, d_sin + off, d_cos + off, m, core, ds->info, stream, &launches);
        g_launches.fetch_add((unsigned long long)launches);
        if (e != cudaSuccess) return record(e, "launch_sincosf");
    }
    return 0;
}

int run_host_f32(const float* in, float* sin_out, float* cos_out, size_t n) {
    if (n == 0) return 0;
    load_env();
    if (!in || !sin_out || !cos_out) return record(cudaErrorInvalidValue, "fabe13_sincosf_host: null array");
    if (stays_on_host(in, sin_out, cos_out, n)) {
        fabe13::host_sincosf_array(in, sin_out, cos_out, n, (int)g_cfg.core.load(), 1);
        g_host_core_calls.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    if (device_count() == 0) {
        if (n >= (size_t)g_cfg.min_n_pageable.load(std::memory_order_relaxed)) return record(cudaErrorNoDevice, "no CUDA device is visible");
        fabe13::host_sincosf_array(in, sin_out, cos_out, n, (int)g_cfg.core.load(), host_core_threads());
        g_host_core_calls.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    Range range("fabe13_sincosf_host");
    const size_t bytes = n * sizeof(float);
    const PtrKind ki = classify(in, bytes), ks = classify(sin_out, bytes), kc = classify(cos_out, bytes);
    if (ki == PTR_DEVICE || ks == PTR_DEVICE || kc == PTR_DEVICE) {
        if (!(ki == PTR_DEVICE && ks == PTR_DEVICE && kc == PTR_DEVICE))
            return record(cudaErrorInvalidValue, "fabe13_sincosf_host: mixed host and device pointers");
        if (int rc = run_device_f32(in, sin_out, cos_out, n, nullptr)) return rc;
        FABE13_TRY(cudaStreamSynchronize(nullptr));
        return 0;
    }
    const bool pinned = (ki == PTR_PINNED && ks == PTR_PINNED && kc == PTR_PINNED);
    if (!pinned && n < (size_t)g_cfg.min_n_pageable.load(std::memory_order_relaxed)) {
        fabe13::host_sincosf_array(in, sin_out, cos_out, n, (int)g_cfg.core.load(), host_core_threads());
        g_host_core_calls.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    DeviceState* ds;
    if (int rc = current_device_state(&ds)) return rc;
    if (pinned) {
        void *d_in = nullptr, *d_s = nullptr, *d_c = nullptr;
        if (cudaHostGetDevicePointer(&d_in, (void*)in, 0) == cudaSuccess && cudaHostGetDevicePointer(&d_s, sin_out, 0) == cudaSuccess &&
            cudaHostGetDevicePointer(&d_c, cos_out, 0) == cudaSuccess) {
            std::lock_guard<std::mutex> lk(ds->small_mu);
            if (!ds->small_stream) FABE13_TRY(cudaStreamCreateWithFlags(&ds->small_stream, cudaStreamNonBlocking));
            if (int rc = run_device_f32((const float*)d_in, (float*)d_s, (float*)d_c, n, ds->small_stream)) return rc;
            FABE13_TRY(cudaStreamSynchronize(ds->small_stream));
            return 0;
        }
        (void)cudaGetLastError();  // registered without a device mapping: stage it like pageable memory
    }
    // pageable memory: like the FP64 path, worker w owns slot w (its pinned staging buffers and its stream) and the
    // pieces i = w (mod workers): copy in, one kernel over the staging buffers in place (zero-copy over PCIe), copy out.
    // The staging copies bound this path (host memory bandwidth), hence the threads.
    int workers = default_stage_workers();
    // pieces of at most 8 MB per staging buffer, small enough that every worker gets about three of them
    size_t piece = n / ((size_t)workers * 3);                     // floats per piece
    if (piece > ((size_t)1 << 21)) piece = (size_t)1 << 21;
    if (piece < ((size_t)1 << 17)) piece = (size_t)1 << 17;
    piece = (piece + 7) & ~(size_t)7;                             // keeps every piece's 32-byte phase
    const size_t piece_doubles = (piece * sizeof(float) + sizeof(double) - 1) / sizeof(double);  // buffers are sized in doubles
    const size_t npieces = (n + piece - 1) / piece;
    if ((size_t)workers > npieces) workers = (int)npieces;
    std::lock_guard<std::mutex> stage_lk(ds->stage_mu);
    std::shared_lock<std::shared_mutex> lk(ds->pipe_rw, std::defer_lock);
    if (int rc = acquire_slots(*ds, lk, piece_doubles, false, true, 0, workers)) return rc;  // staging only: no device buffers
    int dev = 0;
    FABE13_TRY(cudaGetDevice(&dev));
    std::atomic<int> rc_all{0};
    auto work = [&](int w) {
        if (cudaSetDevice(dev) != cudaSuccess) {
            rc_all.store(record(cudaErrorInvalidDevice, "cudaSetDevice(worker)"));
            return;
        }
        Slot& sl = ds->slots[w];
        float* h_in = reinterpret_cast<float*>(sl.h_in);
        float* h_s = reinterpret_cast<float*>(sl.h_sin);
        float* h_c = reinterpret_cast<float*>(sl.h_cos);
        void *d_in = nullptr, *d_s = nullptr, *d_c = nullptr;
        if (cudaHostGetDevicePointer(&d_in, h_in, 0) != cudaSuccess || cudaHostGetDevicePointer(&d_s, h_s, 0) != cudaSuccess ||
            cudaHostGetDevicePointer(&d_c, h_c, 0) != cudaSuccess) {
            rc_all.store(record(cudaErrorInvalidValue, "cudaHostGetDevicePointer(staging)"));
            return;
        }
        for (size_t i = (size_t)w; i < npieces && rc_all.load() == 0; i += (size_t)workers) {
            const size_t off = i * piece;
            const size_t m = n - off < piece ? n - off : piece;
            staging_copy(h_in, in + off, m * sizeof(float));
            int rc;
            {
                std::lock_guard<std::mutex> slot_lk(ds->slot_mu[w]);
                rc = run_device_f32((const float*)d_in, (float*)d_s, (float*)d_c, m, sl.stream);
            }
            if (rc == 0) {
                cudaError_t e = cudaStreamSynchronize(sl.stream);
                if (e != cudaSuccess) rc = record(e, "wait for piece (worker)");
            }
            if (rc != 0) {
                rc_all.store(rc);
                return;
            }
            staging_copy(sin_out + off, h_s, m * sizeof(float));
            staging_copy(cos_out + off, h_c, m * sizeof(float));
        }
    };
    run_workers(workers, work);
    return adopt(rc_all.load());
}

int cfg_core() {
    load_env();
    return (int)g_cfg.core.load();
}

int derived_device(const double* d_in, double* d_out, size_t n, void* cuda_stream, int op, const char* who) {
    if (!d_out && n) return record(cudaErrorInvalidValue, who);
    return run_device(d_in, d_out, nullptr, nullptr, n, nullptr, 0, (cudaStream_t)cuda_stream, op);
}

int derived_host(const double* in, double* out, size_t n, int op, const char* who) {
    if (n == 0) return 0;
    if (!out) return record(cudaErrorInvalidValue, who);
    return run_host(in, out, nullptr, n, op);
}

// give back what the host path holds on one device (caller made it current): slot buffers, pinned staging, the bounce
// buffer, the pool's cached scratch.  With `destroy`, also streams, events and the pool itself.
int release_device(DeviceState& s, bool destroy) {
    std::lock_guard<std::mutex> init_lk(s.init_mu);
    if (!s.ready) return 0;
    std::lock_guard<std::mutex> stage_lk(s.stage_mu);        // same order as the staged paths: stage_mu, then pipe_rw
    std::unique_lock<std::shared_mutex> pipe_lk(s.pipe_rw);  // waits for host calls in flight
    std::lock_guard<std::mutex> small_lk(s.small_mu);
    int rc = 0;
    auto keep = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == 0) rc = record(e, what);
    };
    for (Slot& sl : s.slots) {
        if (sl.stream) keep(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize(slot)");
        if (sl.d_in) keep(cudaFree(sl.d_in), "cudaFree(slot)");
        if (sl.d_sin) keep(cudaFree(sl.d_sin), "cudaFree(slot)");
        if (sl.d_cos) keep(cudaFree(sl.d_cos), "cudaFree(slot)");
        if (sl.d_ws) keep(cudaFree(sl.d_ws), "cudaFree(slot)");
        if (sl.h_in) keep(cudaFreeHost(sl.h_in), "cudaFreeHost(slot)");
        if (sl.h_sin) keep(cudaFreeHost(sl.h_sin), "cudaFreeHost(slot)");
        if (sl.h_cos) keep(cudaFreeHost(sl.h_cos), "cudaFreeHost(slot)");
        sl.d_in = sl.d_sin = sl.d_cos = nullptr;
        sl.d_ws = nullptr;
        sl.h_in = sl.h_sin = sl.h_cos = nullptr;
        sl.cap = sl.ws_bytes = sl.stage_cap = 0;
        if (destroy && sl.stream) {
            keep(cudaStreamDestroy(sl.stream), "cudaStreamDestroy(slot)");
            sl.stream = nullptr;
        }
    }
    if (s.small_stream) keep(cudaStreamSynchronize(s.small_stream), "cudaStreamSynchronize(small)");
    if (s.small_stage) keep(cudaFreeHost(s.small_stage), "cudaFreeHost(bounce)");
    s.small_stage = nullptr;
    s.small_cap = 0;
    if (s.pool) keep(cudaMemPoolTrimTo(s.pool, 0), "cudaMemPoolTrimTo");
    if (destroy) {
        if (s.small_stream) keep(cudaStreamDestroy(s.small_stream), "cudaStreamDestroy(small)");
        s.small_stream = nullptr;
        {
            std::lock_guard<std::mutex> ev_lk(s.ev_mu);
            for (cudaEvent_t ev : s.ev_free) keep(cudaEventDestroy(ev), "cudaEventDestroy");
            s.ev_free.clear();
        }
        if (s.pool) keep(cudaMemPoolDestroy(s.pool), "cudaMemPoolDestroy");
        s.pool = nullptr;
        s.ready = false;
    }
    return rc;
}

int release_all(bool destroy) {
    if (device_count() == 0) return 0;
    int cur = 0;
    FABE13_TRY(cudaGetDevice(&cur));
    int rc = 0;
    for (int d = 0; d < kMaxDevices && d < device_count(); ++d) {
        {
            std::lock_guard<std::mutex> lk(g_dev[d].init_mu);
            if (!g_dev[d].ready) continue;
        }
        if (cudaSetDevice(d) != cudaSuccess) {
            (void)cudaGetLastError();
            continue;
        }
        const int r = release_device(g_dev[d], destroy);
        if (r && !rc) rc = r;
    }
    (void)cudaSetDevice(cur);
    return rc;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

void fabe13_sincos(const double* in, double* sin_out, double* cos_out, int n) {
    if (n <= 0) return;  // reference fabe13/fabe13.c:223
    const int rc = guarded("fabe13_sincos", [&] { return run_host(in, sin_out, cos_out, (size_t)n); });
    if (rc == 0) return;
    // The reference's ABI cannot report a failure (void, fabe13/fabe13.h:51-56): rather than leave the outputs
    // unwritten, fill them with the host instantiation of the same core (same bits).  The error stays recorded.
    if (g_cfg.fallback.load() == 0 || !in || !sin_out || !cos_out) return;
    if (device_count() != 0 && (classify_byte(in, nullptr) == PTR_DEVICE || classify_byte(sin_out, nullptr) == PTR_DEVICE ||
                                classify_byte(cos_out, nullptr) == PTR_DEVICE))
        return;  // device memory: nothing the host can do
    const int rc2 = guarded("fabe13_sincos (host fallback)", [&] {
        fabe13::host_sincos_array(in, sin_out, cos_out, (size_t)n, FABE13_OP_SINCOS, cfg_core(), host_core_threads(), host_stream_stores());
        return 0;
    });
    if (rc2 == 0) {
        g_fallbacks.fetch_add(1);
        tl_last_error = rc;  // what went wrong on the CUDA side stays visible
    }
}

int fabe13_sincos_host(const double* in, double* sin_out, double* cos_out, size_t n) {
    return guarded("fabe13_sincos_host", [&] { return run_host(in, sin_out, cos_out, n); });
}

const char* fabe13_get_active_implementation_name(void) {
    static char host_name[160];
    static std::once_flag host_name_once;
    DeviceState* ds;
    if (device_count() == 0) {
        std::call_once(host_name_once, [] {
            snprintf(host_name, sizeof host_name, "CUDA sm_100a (no device visible; host core %s)", fabe13::host_isa_name());
        });
        return host_name;
    }
    if (current_device_state(&ds) != 0) return "CUDA sm_100a (device initialisation failed)";
    return ds->name;
}

int fabe13_get_active_simd_width(void) { return 32; }

// scalar API: host instantiation of the array core (reference fabe13/fabe13.c:262-269)
double fabe13_sin(double x) {
    double s, c;
    fabe13::host_sincos_one(x, cfg_core(), &s, &c, nullptr);
    return s;
}
double fabe13_cos(double x) {
    double s, c;
    fabe13::host_sincos_one(x, cfg_core(), &s, &c, nullptr);
    return c;
}
// sinc / tan / cot: the reference's guards (fabe13/fabe13.c:264-266: |x| < 1e-9 -> 1, cos == 0 -> NAN, sin == 0 -> NAN)
// around one division; the same fabe13_derived the array kernels compile for the device
double fabe13_sinc(double x) { return fabe13::host_derived_one(x, FABE13_OP_SINC, cfg_core()); }
double fabe13_tan(double x) { return fabe13::host_derived_one(x, FABE13_OP_TAN, cfg_core()); }
double fabe13_cot(double x) { return fabe13::host_derived_one(x, FABE13_OP_COT, cfg_core()); }
double fabe13_atan(double x) { return atan(x); }
double fabe13_asin(double x) { return asin(x); }
double fabe13_acos(double x) { return acos(x); }

int fabe13_sincos_device(const double* d_in, double* d_sin, double* d_cos, size_t n, void* cuda_stream) {
    return guarded("fabe13_sincos_device", [&] { return run_device(d_in, d_sin, d_cos, nullptr, n, nullptr, 0, (cudaStream_t)cuda_stream); });
}

// single-output array variants (the array forms of fabe13_sin / fabe13_cos): 16 B/element instead of 24
int fabe13_sin_device(const double* d_in, double* d_out, size_t n, void* cuda_stream) {
    if (!d_out && n) return record(cudaErrorInvalidValue, "fabe13_sin_device: null output");
    return guarded("fabe13_sin_device", [&] { return run_device(d_in, d_out, nullptr, nullptr, n, nullptr, 0, (cudaStream_t)cuda_stream); });
}
int fabe13_cos_device(const double* d_in, double* d_out, size_t n, void* cuda_stream) {
    if (!d_out && n) return record(cudaErrorInvalidValue, "fabe13_cos_device: null output");
    return guarded("fabe13_cos_device", [&] { return run_device(d_in, nullptr, d_out, nullptr, n, nullptr, 0, (cudaStream_t)cuda_stream); });
}
int fabe13_sin_host(const double* in, double* out, size_t n) {
    if (n == 0) return 0;
    if (!out) return record(cudaErrorInvalidValue, "fabe13_sin_host: null output");
    return guarded("fabe13_sin_host", [&] { return run_host(in, out, nullptr, n); });
}
int fabe13_cos_host(const double* in, double* out, size_t n) {
    if (n == 0) return 0;
    if (!out) return record(cudaErrorInvalidValue, "fabe13_cos_host: null output");
    return guarded("fabe13_cos_host", [&] { return run_host(in, nullptr, out, n); });
}

// array forms of fabe13_tan / fabe13_cot / fabe13_sinc: the same per-element code as the scalar entry points
int fabe13_tan_device(const double* d_in, double* d_out, size_t n, void* cuda_stream) {
    return guarded("fabe13_tan_device", [&] { return derived_device(d_in, d_out, n, cuda_stream, FABE13_OP_TAN, "fabe13_tan_device: null output"); });
}
int fabe13_cot_device(const double* d_in, double* d_out, size_t n, void* cuda_stream) {
    return guarded("fabe13_cot_device", [&] { return derived_device(d_in, d_out, n, cuda_stream, FABE13_OP_COT, "fabe13_cot_device: null output"); });
}
int fabe13_sinc_device(const double* d_in, double* d_out, size_t n, void* cuda_stream) {
    return guarded("fabe13_sinc_device", [&] { return derived_device(d_in, d_out, n, cuda_stream, FABE13_OP_SINC, "fabe13_sinc_device: null output"); });
}
int fabe13_tan_host(const double* in, double* out, size_t n) {
    return guarded("fabe13_tan_host", [&] { return derived_host(in, out, n, FABE13_OP_TAN, "fabe13_tan_host: null output"); });
}
int fabe13_cot_host(const double* in, double* out, size_t n) {
    return guarded("fabe13_cot_host", [&] { return derived_host(in, out, n, FABE13_OP_COT, "fabe13_cot_host: null output"); });
}
int fabe13_sinc_host(const double* in, double* out, size_t n) {
    return guarded("fabe13_sinc_host", [&] { return derived_host(in, out, n, FABE13_OP_SINC, "fabe13_sinc_host: null output"); });
}

// FP32 arrays (the reference's roadmap item README.md:225; no reference implementation exists)
int fabe13_sincosf_device(const float* d_in, float* d_sin, float* d_cos, size_t n, void* cuda_stream) {
    if (n && (!d_in || !d_sin || !d_cos)) return record(cudaErrorInvalidValue, "fabe13_sincosf_device: null array");
    return guarded("fabe13_sincosf_device", [&] { return run_device_f32(d_in, d_sin, d_cos, n, (cudaStream_t)cuda_stream); });
}
int fabe13_sincosf_host(const float* in, float* sin_out, float* cos_out, size_t n) {
    return guarded("fabe13_sincosf_host", [&] { return run_host_f32(in, sin_out, cos_out, n); });
}

int fabe13_sincos_device_k(const double* d_in, double* d_sin, double* d_cos, long long* d_k, size_t n,
                           void* cuda_stream) {
    return guarded("fabe13_sincos_device_k", [&] { return run_device(d_in, d_sin, d_cos, d_k, n, nullptr, 0, (cudaStream_t)cuda_stream); });
}

size_t fabe13_sincos_workspace_bytes(size_t n) {
    const LaunchTuning t = current_tuning();
    return fabe13::workspace_bytes(n < kLaunchMaxElems ? n : kLaunchMaxElems, t);
}

int fabe13_sincos_device_ws(const double* d_in, double* d_sin, double* d_cos, size_t n, void* workspace,
                            size_t workspace_bytes, void* cuda_stream) {
    if (!workspace && fabe13_sincos_workspace_bytes(n) > 0)
        return record(cudaErrorInvalidValue, "fabe13_sincos_device_ws: null workspace");
    return guarded("fabe13_sincos_device_ws",
                   [&] { return run_device(d_in, d_sin, d_cos, nullptr, n, workspace, workspace_bytes, (cudaStream_t)cuda_stream); });
}

int fabe13_sincos_multi(int ngpu, const int* devices, const double* const* d_in, double* const* d_sin,
                        double* const* d_cos, const size_t* n_per_dev) {
    return guarded("fabe13_sincos_multi", [&]() -> int {
        int cur = 0;
        FABE13_TRY(cudaGetDevice(&cur));
        int rc = 0;
        for (int g = 0; g < ngpu && rc == 0; ++g) {
            cudaError_t e = cudaSetDevice(devices[g]);
            if (e != cudaSuccess) {
                rc = record(e, "cudaSetDevice(multi)");
                break;
            }
            rc = run_device(d_in[g], d_sin[g], d_cos[g], nullptr, n_per_dev[g], nullptr, 0, nullptr);
        }
        for (int g = 0; g < ngpu; ++g) {
            if (cudaSetDevice(devices[g]) != cudaSuccess) continue;
            cudaError_t e = cudaStreamSynchronize(nullptr);
            if (e != cudaSuccess && rc == 0) rc = record(e, "cudaStreamSynchronize(multi)");
        }
        (void)cudaSetDevice(cur);
        return rc;
    });
}

void* fabe13_cuda_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        record(cudaErrorMemoryAllocation, "fabe13_cuda_host_alloc");
        return nullptr;
    }
    return p;
}

void fabe13_cuda_host_free(void* p) {
    if (p) (void)cudaFreeHost(p);
}

int fabe13_cuda_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return record(cudaErrorInvalidValue, "fabe13_cuda_host_register");
    FABE13_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}

int fabe13_cuda_host_unregister(void* p) {
    FABE13_TRY(cudaHostUnregister(p));
    return 0;
}

int fabe13_cuda_trim(void) {
    return guarded("fabe13_cuda_trim", [] { return release_all(false); });
}

int fabe13_cuda_shutdown(void) {
    return guarded("fabe13_cuda_shutdown", [] { return release_all(true); });
}

void fabe13_shard_bounds(size_t n, int world, int rank, size_t* begin, size_t* end) {
    if (world < 1) world = 1;
    if (rank < 0) rank = 0;
    if (rank >= world) rank = world - 1;
    const size_t per = (n / (size_t)world) & ~(size_t)3;
    *begin = per * (size_t)rank;
    *end = (rank == world - 1) ? n : per * (size_t)(rank + 1);
}

int fabe13_cuda_set_option(const char* key, long long value) {
    if (!key) return -1;
    load_env();
    for (const OptionDesc& o : kOptions) {
        if (strcmp(o.key, key) == 0) {
            if (value < o.lo || value > o.hi) return -2;
            (g_cfg.*(o.field)).store(value);
            return 0;
        }
    }
    return -1;
}

int fabe13_cuda_get_option(const char* key, long long* value) {
    if (!key || !value) return -1;
    load_env();
    for (const OptionDesc& o : kOptions) {
        if (strcmp(o.key, key) == 0) {
            *value = (g_cfg.*(o.field)).load();
            return 0;
        }
    }
    if (strcmp(key, "launches") == 0) {
        *value = (long long)g_launches.load();
        return 0;
    }
    if (strcmp(key, "host_core_calls") == 0) {
        *value = (long long)g_host_core_calls.load();
        return 0;
    }
    if (strcmp(key, "fallbacks") == 0) {
        *value = (long long)g_fallbacks.load();
        return 0;
    }
    if (strcmp(key, "assist_host_elems") == 0) {
        *value = (long long)g_assist_host_elems.load();
        return 0;
    }
    if (strcmp(key, "assist_gpu_elems") == 0) {
        *value = (long long)g_assist_gpu_elems.load();
        return 0;
    }
    if (strcmp(key, "device_count") == 0) {
        *value = fabe13_cuda_device_count();
        return 0;
    }
    if (strcmp(key, "sm_count") == 0) {
        DeviceState* ds;
        if (int rc = current_device_state(&ds)) return rc;
        *value = ds->info.sm_count;
        return 0;
    }
    return -1;
}

int fabe13_cuda_device_count(void) { return device_count(); }

int fabe13_cuda_last_error(void) { return tl_last_error; }

const char* fabe13_cuda_last_error_string(void) { return cudaGetErrorString((cudaError_t)tl_last_error); }

void fabe13_cuda_clear_error(void) { tl_last_error = 0; }

void fabe13_debug_host_core(const double* in, double* sin_out, double* cos_out, long long* k_out, size_t n, int core) {
    for (size_t i = 0; i < n; ++i) fabe13::host_sincos_one(in[i], core, &sin_out[i], &cos_out[i], k_out ? &k_out[i] : nullptr);
}

void fabe13_debug_host_sincosf(const float* in, float* sin_out, float* cos_out, size_t n, int core) {
    for (size_t i = 0; i < n; ++i) fabe13::host_sincosf_one(in[i], core, &sin_out[i], &cos_out[i]);
}

void fabe13_debug_host_derived(const double* in, double* out, size_t n, int op, int core) {
    for (size_t i = 0; i < n; ++i) out[i] = fabe13::host_derived_one(in[i], op, core);
}

// the host array loop (vectorised clones + patching) as the product calls it, for CPU-only tests of that code
void fabe13_debug_host_array(const double* in, double* sin_out, double* cos_out, size_t n, int op, int core, int threads) {
    fabe13::host_sincos_array(in, sin_out, cos_out, n, op, core, threads);
}

void fabe13_debug_reduce_huge(const double* in, double* r, int* q, double* r_exact, int* q_exact, int* fast_ok, size_t n) {
    fabe13::host_reduce_huge(in, r, q, r_exact, q_exact, fast_ok, n);
}

void fabe13_debug_host_array_stores(const double* in, double* sin_out, double* cos_out, size_t n, int op, int core, int threads,
                                    int stream_stores) {
    fabe13::host_sincos_array(in, sin_out, cos_out, n, op, core, threads, stream_stores);
}

int fabe13_debug_fill_uniform_device(double* d_out, size_t n, unsigned long long seed, unsigned long long first_index,
                                     double lo, double hi, void* cuda_stream) {
    DeviceState* ds;
    if (int rc = current_device_state(&ds)) return rc;
    FABE13_TRY(fabe13::fill_uniform(d_out, n, seed, first_index, lo, hi, ds->info, (cudaStream_t)cuda_stream));
    return 0;
}

int fabe13_debug_fill_loguniform_device(double* d_out, size_t n, unsigned long long seed,
                                        unsigned long long first_index, void* cuda_stream) {
    DeviceState* ds;
    if (int rc = current_device_state(&ds)) return rc;
    FABE13_TRY(fabe13::fill_loguniform(d_out, n, seed, first_index, ds->info, (cudaStream_t)cuda_stream));
    return 0;
}

}  // extern "C"
