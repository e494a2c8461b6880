// pcie_probe.cu -- the ceiling of the host-resident path: bare cudaMemcpyAsync H2D || D2H in the path's own mix
// (8 bytes in : 16 bytes out per element) between pinned host memory and 1..N GPUs at once, no kernel in between.
// What fabe13_sincos_host can reach end to end is bounded by what this prints (the reference
// call site is fabe13/benchmark_fabe13.c:212-218 -- arrays in host memory).
//
//   make -C tools pcie_probe && tools/pcie_probe [--gpus 1,2,4,8] [--secs 1.0] [--chunk-mb 32] [--quick]
//
// For every GPU count it runs, with the copies of all GPUs inside one common time window:
//   style   threads = one process, one host thread per GPU;  procs = one forked process per GPU (bench.py's layout)
//   alloc   default   cudaHostAlloc(Default)
//           portable  cudaHostAlloc(Portable | Mapped)            (what fabe13_cuda_host_alloc returns)
//           wc-in     as portable, input buffer write-combined    (H2D source never read by the CPU)
//           huge-reg  mmap + MADV_HUGEPAGE (or MAP_HUGETLB if the pool has pages) + cudaHostRegister
//   bind    each GPU's thread / process pinned to its own core (the GPU's NUMA node's cores if sysfs names one)
//   dirs    both (the path's mix), h2d only, d2h only
// and a host-memory line: the same 8:16 byte mix moved by T CPU threads with non-temporal stores (what the host DRAM
// gives the CPU arm of bench.py -- and what 8 GPUs' DMA has to share).
// One cudaMemcpyAsync per copy; nothing batched.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum { ALLOC_DEFAULT = 0, ALLOC_PORTABLE, ALLOC_WC_IN, ALLOC_HUGE_REG, ALLOC_COUNT };
static const char* kAllocName[] = {"default", "portable", "wc-in", "huge-reg"};
enum { DIR_BOTH = 0, DIR_H2D, DIR_D2H };
static const char* kDirName[] = {"both", "h2d", "d2h"};

struct Shared {  // one page shared by threads or forked processes
    std::atomic<int> arrived;
    std::atomic<int> go;
    std::atomic<int> failed;
    double h2d_bytes[16], d2h_bytes[16], t_begin[16], t_end[16];
    int huge_kind[16];  // 0 none, 1 THP madvise, 2 MAP_HUGETLB
};

struct Cfg {
    int ngpu;
    int alloc;
    int dir;
    bool bind;
    double secs;
    size_t chunk_in;  // bytes per H2D copy; D2H copies are twice that
    int depth;        // chunks in flight per direction
};

static std::vector<int> cpus_of_node(int node) {
    std::vector<int> out;
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return out;
    char buf[1024] = {0};
    if (fgets(buf, sizeof buf, f)) {
        for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int a, b;
            if (sscanf(tok, "%d-%d", &a, &b) == 2)
                for (int c = a; c <= b; ++c) out.push_back(c);
            else if (sscanf(tok, "%d", &a) == 1)
                out.push_back(a);
        }
    }
    fclose(f);
    return out;
}

static int numa_node_of_gpu(int dev) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return -1;
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static void bind_to_core(int g, int node) {
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    sched_getaffinity(0, sizeof allowed, &allowed);
    std::vector<int> cand;
    if (node >= 0)
        for (int c : cpus_of_node(node))
            if (CPU_ISSET(c, &allowed)) cand.push_back(c);
    if (cand.empty())
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) cand.push_back(c);
    if (cand.empty()) return;
    cpu_set_t one;
    CPU_ZERO(&one);
    CPU_SET(cand[(size_t)g % cand.size()], &one);
    sched_setaffinity(0, sizeof one, &one);
}

static void* huge_alloc(size_t bytes, int* kind) {
    const size_t two_mb = (size_t)2 << 20;
    bytes = (bytes + two_mb - 1) & ~(two_mb - 1);
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
    if (p != MAP_FAILED) {
        *kind = 2;
    } else {
        p = mmap(nullptr, bytes + two_mb, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) return nullptr;
        p = (void*)(((uintptr_t)p + two_mb - 1) & ~(uintptr_t)(two_mb - 1));
        madvise(p, bytes, MADV_HUGEPAGE);
        *kind = 1;
    }
    memset(p, 1, bytes);  // touch: the pages exist (and are huge where the kernel obliges) before they are locked
    return p;
}

#define CK(x)                                                                                 \
    do {                                                                                      \
        cudaError_t e__ = (x);                                                                \
        if (e__ != cudaSuccess) {                                                             \
            fprintf(stderr, "pcie_probe: %s -> %s\n", #x, cudaGetErrorString(e__));           \
            sh->failed.store(1);                                                              \
            goto done;                                                                        \
        }                                                                                     \
    } while (0)

// one GPU's share of a configuration; g indexes Shared, dev is the CUDA ordinal
static void worker(const Cfg cfg, int g, int dev, Shared* sh) {
    const size_t in_bytes = cfg.chunk_in * (size_t)cfg.depth, out_bytes = 2 * in_bytes;
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in((size_t)cfg.depth, nullptr), ev_out((size_t)cfg.depth, nullptr);
    bool registered = false, arrived = false;
    double h2d = 0, d2h = 0;
    int kind = 0;
    CK(cudaSetDevice(dev));
    if (cfg.bind) bind_to_core(g, numa_node_of_gpu(dev));
    if (cfg.alloc == ALLOC_HUGE_REG) {
        h_in = huge_alloc(in_bytes, &kind);
        h_out = huge_alloc(out_bytes, &kind);
        if (!h_in || !h_out) {
            sh->failed.store(1);
            goto done;
        }
        CK(cudaHostRegister(h_in, in_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
        CK(cudaHostRegister(h_out, out_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
        registered = true;
    } else {
        const unsigned base = cfg.alloc == ALLOC_DEFAULT ? cudaHostAllocDefault : (cudaHostAllocPortable | cudaHostAllocMapped);
        CK(cudaHostAlloc(&h_in, in_bytes, base | (cfg.alloc == ALLOC_WC_IN ? cudaHostAllocWriteCombined : 0)));
        CK(cudaHostAlloc(&h_out, out_bytes, base));
        memset(h_in, 1, in_bytes);
        memset(h_out, 1, out_bytes);
    }
    sh->huge_kind[g] = kind;
    CK(cudaMalloc(&d_in, in_bytes));
    CK(cudaMalloc(&d_out, out_bytes));
    CK(cudaMemset(d_out, 0, out_bytes));
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int i = 0; i < cfg.depth; ++i) {
        CK(cudaEventCreateWithFlags(&ev_in[(size_t)i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_out[(size_t)i], cudaEventDisableTiming));
    }
    // warm-up: one copy each way
    CK(cudaMemcpyAsync(d_in, h_in, cfg.chunk_in, cudaMemcpyHostToDevice, s_in));
    CK(cudaMemcpyAsync(h_out, d_out, 2 * cfg.chunk_in, cudaMemcpyDeviceToHost, s_out));
    CK(cudaStreamSynchronize(s_in));
    CK(cudaStreamSynchronize(s_out));
    // common start
    sh->arrived.fetch_add(1);
    arrived = true;
    while (sh->go.load() == 0 && sh->failed.load() == 0) sched_yield();
    if (sh->failed.load()) goto done;
    {
        const double t0 = now_s();
        sh->t_begin[g] = t0;
        long round = 0;
        // rings of `depth` chunks per direction: chunk i is re-issued as soon as its previous copy has finished
        while (now_s() - t0 < cfg.secs) {
            const size_t i = (size_t)(round % cfg.depth);
            if (round >= cfg.depth) {
                if (cfg.dir != DIR_D2H) CK(cudaEventSynchronize(ev_in[i]));
                if (cfg.dir != DIR_H2D) CK(cudaEventSynchronize(ev_out[i]));
            }
            if (cfg.dir != DIR_D2H) {
                CK(cudaMemcpyAsync((char*)d_in + i * cfg.chunk_in, (char*)h_in + i * cfg.chunk_in, cfg.chunk_in, cudaMemcpyHostToDevice, s_in));
                CK(cudaEventRecord(ev_in[i], s_in));
                h2d += (double)cfg.chunk_in;
            }
            if (cfg.dir != DIR_H2D) {
                // the path writes two output arrays: two copies of chunk size, as the pipeline issues them
                CK(cudaMemcpyAsync((char*)h_out + 2 * i * cfg.chunk_in, (char*)d_out + 2 * i * cfg.chunk_in, cfg.chunk_in, cudaMemcpyDeviceToHost, s_out));
                CK(cudaMemcpyAsync((char*)h_out + (2 * i + 1) * cfg.chunk_in, (char*)d_out + (2 * i + 1) * cfg.chunk_in, cfg.chunk_in,
                                   cudaMemcpyDeviceToHost, s_out));
                CK(cudaEventRecord(ev_out[i], s_out));
                d2h += 2.0 * (double)cfg.chunk_in;
            }
            ++round;
        }
        CK(cudaStreamSynchronize(s_in));
        CK(cudaStreamSynchronize(s_out));
        sh->t_end[g] = now_s();
        sh->h2d_bytes[g] = h2d;
        sh->d2h_bytes[g] = d2h;
    }
done:
    if (!arrived) sh->arrived.fetch_add(1);  // never leave the others waiting
    for (cudaEvent_t e : ev_in)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_out)
        if (e) cudaEventDestroy(e);
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    if (registered) {
        cudaHostUnregister(h_in);
        cudaHostUnregister(h_out);
    } else if (cfg.alloc != ALLOC_HUGE_REG) {
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
    }
    // (huge-reg buffers are left to process exit / the next mmap: munmap needs the unaligned base)
}

struct Row {
    double h2d_gbs, d2h_gbs;
    bool ok;
    int huge_kind;
};

static Row collect(const Cfg& cfg, Shared* sh) {
    Row r{0, 0, sh->failed.load() == 0, 0};
    if (!r.ok) return r;
    double t0 = 1e300, t1 = 0, h = 0, d = 0;
    for (int g = 0; g < cfg.ngpu; ++g) {
        if (sh->t_begin[g] < t0) t0 = sh->t_begin[g];
        if (sh->t_end[g] > t1) t1 = sh->t_end[g];
        h += sh->h2d_bytes[g];
        d += sh->d2h_bytes[g];
        if (sh->huge_kind[g] > r.huge_kind) r.huge_kind = sh->huge_kind[g];
    }
    r.h2d_gbs = h / (t1 - t0) / 1e9;
    r.d2h_gbs = d / (t1 - t0) / 1e9;
    return r;
}

// One (GPU count, style) group of configurations.  CUDA contexts are expensive to create (about a second each on an
// 8-GPU box), so a group creates them once: style "threads" = ONE forked child that runs every configuration with one
// thread per GPU; style "procs" = one forked child per GPU (bench.py's layout), each walking the same list of
// configurations, meeting the others at a barrier in shared memory before every one.  The parent never touches CUDA
// (fork after CUDA initialisation is not supported).
struct Group {
    Shared* sh;   // one Shared per configuration
    Row* rows;
    int ncfg;
};

static Group new_group(int ncfg) {
    Group g;
    g.ncfg = ncfg;
    g.sh = (Shared*)mmap(nullptr, sizeof(Shared) * (size_t)ncfg, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    g.rows = (Row*)mmap(nullptr, sizeof(Row) * (size_t)ncfg, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset((void*)g.sh, 0, sizeof(Shared) * (size_t)ncfg);
    memset((void*)g.rows, 0, sizeof(Row) * (size_t)ncfg);
    return g;
}

static void free_group(Group& g) {
    munmap(g.sh, sizeof(Shared) * (size_t)g.ncfg);
    munmap(g.rows, sizeof(Row) * (size_t)g.ncfg);
}

static void run_group(const std::vector<Cfg>& cfgs, bool procs, Group& g) {
    const int ngpu = cfgs[0].ngpu;
    std::vector<pid_t> kids;
    if (!procs) {
        const pid_t pid = fork();
        if (pid == 0) {
            for (size_t c = 0; c < cfgs.size(); ++c) {
                Shared* sh = &g.sh[c];
                std::vector<std::thread> th;
                for (int d = 0; d < ngpu; ++d) th.emplace_back(worker, cfgs[c], d, d, sh);
                while (sh->arrived.load() < ngpu) sched_yield();
                sh->go.store(1);
                for (std::thread& t : th) t.join();
                g.rows[c] = collect(cfgs[c], sh);
            }
            _exit(0);
        }
        kids.push_back(pid);
    } else {
        for (int d = 0; d < ngpu; ++d) {
            const pid_t pid = fork();
            if (pid == 0) {
                for (size_t c = 0; c < cfgs.size(); ++c) {
                    Shared* sh = &g.sh[c];
                    if (d == 0) {  // rank 0 opens the window once everybody has arrived
                        std::thread opener([sh, ngpu] {
                            while (sh->arrived.load() < ngpu) sched_yield();
                            sh->go.store(1);
                        });
                        worker(cfgs[c], d, d, sh);
                        opener.join();
                    } else {
                        worker(cfgs[c], d, d, sh);
                    }
                }
                _exit(0);
            }
            kids.push_back(pid);
        }
    }
    bool ok = true;
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ok = false;
    }
    if (procs)
        for (size_t c = 0; c < cfgs.size(); ++c) g.rows[c] = collect(cfgs[c], &g.sh[c]);
    if (!ok)
        for (size_t c = 0; c < cfgs.size(); ++c) g.rows[c].ok = false;
}

// host memory: T threads each read 8 bytes and write 16 per element (non-temporal stores), the CPU arm's traffic mix
static double host_mix_gbs(int threads, double secs) {
    const size_t n = (size_t)16 << 20;  // doubles per thread: 128 MB in, 256 MB out
    std::vector<double*> in((size_t)threads), out((size_t)threads);
    for (int t = 0; t < threads; ++t) {
        in[(size_t)t] = (double*)aligned_alloc(64, n * 8);
        out[(size_t)t] = (double*)aligned_alloc(64, n * 16);
        memset(in[(size_t)t], 1, n * 8);
        memset(out[(size_t)t], 1, n * 16);
    }
    std::atomic<int> go{0};
    std::vector<double> bytes((size_t)threads, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([&, t] {
            while (go.load() == 0) sched_yield();
            const double t0 = now_s();
            double moved = 0;
            while (now_s() - t0 < secs) {
                const double* a = in[(size_t)t];
                double* b = out[(size_t)t];
#if defined(__x86_64__)
                for (size_t i = 0; i < n; i += 2) {
                    const __m128d v = _mm_load_pd(a + i);
                    _mm_stream_pd(b + i, v);
                    _mm_stream_pd(b + n + i, v);
                }
                _mm_sfence();
#else
                for (size_t i = 0; i < n; ++i) {
                    b[i] = a[i];
                    b[n + i] = a[i];
                }
#endif
                moved += 24.0 * (double)n;
            }
            bytes[(size_t)t] = moved / (now_s() - t0);
        });
    go.store(1);
    for (std::thread& t : th) t.join();
    double total = 0;
    for (double b : bytes) total += b;
    for (int t = 0; t < threads; ++t) {
        free(in[(size_t)t]);
        free(out[(size_t)t]);
    }
    return total / 1e9;
}

static void print_topology(int ngpu_visible) {
    // runs in a child: touches CUDA
    printf("host: %ld online CPUs", sysconf(_SC_NPROCESSORS_ONLN));
    for (int node = 0; node < 8; ++node) {
        std::vector<int> c = cpus_of_node(node);
        if (!c.empty()) printf(", node%d: %zu cpus (%d..%d)", node, c.size(), c.front(), c.back());
    }
    printf("\n");
    for (int d = 0; d < ngpu_visible; ++d) {
        cudaDeviceProp p;
        char bus[32] = {0};
        if (cudaGetDeviceProperties(&p, d) != cudaSuccess) break;
        cudaDeviceGetPCIBusId(bus, sizeof bus, d);
        printf("gpu %d: %s, pci %s, numa_node %d, async engines %d\n", d, p.name, bus, numa_node_of_gpu(d), p.asyncEngineCount);
    }
    fflush(stdout);
}

int main(int argc, char** argv) {
    std::vector<int> counts;
    double secs = 1.0;
    size_t chunk_mb = 32;
    int depth = 4;
    bool quick = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) {
            for (char* tok = strtok(argv[++i], ","); tok; tok = strtok(nullptr, ",")) counts.push_back(atoi(tok));
        } else if (!strcmp(argv[i], "--secs") && i + 1 < argc) {
            secs = atof(argv[++i]);
        } else if (!strcmp(argv[i], "--chunk-mb") && i + 1 < argc) {
            chunk_mb = (size_t)atoll(argv[++i]);
        } else if (!strcmp(argv[i], "--depth") && i + 1 < argc) {
            depth = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--quick")) {
            quick = true;
        } else {
            fprintf(stderr, "usage: pcie_probe [--gpus 1,2,4,8] [--secs S] [--chunk-mb MB] [--depth K] [--quick]\n");
            return 2;
        }
    }
    // the device count, asked in a child so that this process stays free to fork CUDA users
    int* shared_count = (int*)mmap(nullptr, sizeof(int), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    *shared_count = 0;
    {
        const pid_t pid = fork();
        if (pid == 0) {
            int n = 0;
            if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
            *shared_count = n;
            if (n > 0) print_topology(n);
            _exit(0);
        }
        int st;
        waitpid(pid, &st, 0);
    }
    const int visible = *shared_count;
    if (visible <= 0) {
        fprintf(stderr, "pcie_probe: no CUDA device\n");
        return 1;
    }
    if (counts.empty())
        for (int c = 1; c <= visible; c *= 2) counts.push_back(c);
    printf("mix: H2D 8 B : D2H 16 B per element; copies of %zu MB (D2H: two per chunk), %d chunks in flight per direction, %.1f s per line\n",
           chunk_mb, depth, secs);
    printf("%5s %8s %9s %5s %5s %10s %10s %10s %12s\n", "gpus", "style", "alloc", "bind", "dirs", "H2D GB/s", "D2H GB/s", "sum GB/s",
           "Gelem/s eq");
    for (int ngpu : counts) {
        if (ngpu > visible || ngpu > 16) continue;
        for (int procs = 0; procs < 2; ++procs) {
            std::vector<Cfg> cfgs;
            for (int alloc = 0; alloc < ALLOC_COUNT; ++alloc) {
                for (int bind = 0; bind < 2; ++bind) {
                    for (int dir = 0; dir < 3; ++dir) {
                        // the full cross product only where it informs: directions alone and binding for the
                        // portable allocation; every allocation mode for the path's own mix
                        if (dir != DIR_BOTH && (alloc != ALLOC_PORTABLE || bind)) continue;
                        if (bind && alloc != ALLOC_PORTABLE && alloc != ALLOC_HUGE_REG) continue;
                        if (quick && (alloc == ALLOC_DEFAULT || alloc == ALLOC_WC_IN || (bind && alloc != ALLOC_PORTABLE))) continue;
                        cfgs.push_back(Cfg{ngpu, alloc, dir, bind != 0, secs, chunk_mb << 20, depth});
                    }
                }
            }
            Group g = new_group((int)cfgs.size());
            run_group(cfgs, procs != 0, g);
            for (size_t c = 0; c < cfgs.size(); ++c) {
                const Cfg& cfg = cfgs[c];
                const Row& r = g.rows[c];
                if (!r.ok) {
                    printf("%5d %8s %9s %5s %5s %10s\n", ngpu, procs ? "procs" : "threads", kAllocName[cfg.alloc], cfg.bind ? "yes" : "no",
                           kDirName[cfg.dir], "failed");
                    continue;
                }
                // elements/s this traffic corresponds to: bounded by the slower of in/8 and out/16
                double ge = 0;
                if (cfg.dir == DIR_BOTH) ge = (r.h2d_gbs / 8 < r.d2h_gbs / 16 ? r.h2d_gbs / 8 : r.d2h_gbs / 16);
                printf("%5d %8s %9s%s %5s %5s %10.1f %10.1f %10.1f %12.2f\n", ngpu, procs ? "procs" : "threads", kAllocName[cfg.alloc],
                       r.huge_kind == 2 ? "*" : (r.huge_kind == 1 ? "+" : " "), cfg.bind ? "yes" : "no", kDirName[cfg.dir], r.h2d_gbs, r.d2h_gbs,
                       r.h2d_gbs + r.d2h_gbs, ge);
            }
            fflush(stdout);
            free_group(g);
        }
    }
    printf("(huge-reg: * = MAP_HUGETLB pages, + = transparent huge pages requested with madvise)\n");
    const long cpus = sysconf(_SC_NPROCESSORS_ONLN);
    printf("\nhost memory, CPU only: 8 B read : 16 B written per element with non-temporal stores\n%8s %10s %12s\n", "threads", "GB/s", "Gelem/s eq");
    for (long t = 1; t <= cpus; t *= 2) {
        const double gbs = host_mix_gbs((int)t, secs < 0.5 ? secs : 0.5);
        printf("%8ld %10.1f %12.2f\n", t, gbs, gbs / 24);
        fflush(stdout);
    }
    if ((cpus & (cpus - 1)) != 0) {
        const double gbs = host_mix_gbs((int)cpus, secs < 0.5 ? secs : 0.5);
        printf("%8ld %10.1f %12.2f\n", cpus, gbs, gbs / 24);
    }
    return 0;
}
