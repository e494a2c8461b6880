// fabe13_kernels.cu -- sm_100a kernels for the FP64 array sincos path.
//
// Replaces the reference's SIMD array drivers fabe13_sincos_avx512 / _avx2 / _neon (fabe13/fabe13.c:545-614,
// :469-541, :399-464) and their scalar tail loops (:339-358).  Kernels:
//
//   sincos_fused_kernel       the hot path.  Tiles of 128*UNROLL 256-bit vectors in address order (by default one
//                             tile per block, so DRAM sees one advancing window); each thread issues LDG.E.NA.256 for
//                             its vectors, then two STG.E.EF.256 per vector.  Per lane: reference-exact k, Cody-Waite,
//                             polynomial/Psi core, integer-pipe rotation.  24 B/element of HBM traffic.  Lanes with
//                             |x| >= 2^45 are found with warp votes and reduced by Payne-Hanek in the same kernel:
//                             a dense warp runs the every-range routine for all its lanes; otherwise the warp compacts
//                             such lanes into its slice of shared memory and strides over that list with all lanes
//                             busy; a warp with a lone one (hybrid launches) appends it to a global worklist.
//   sincos_second_pass_kernel the global worklist: one Payne-Hanek reduction per entry, lanes densely packed; if a
//                             segment overflowed (option worklist=1 only) it also rescans the output for arguments
//                             still parked there (a finite value >= 2^45 can never be a sine).
//   sincos_inline_kernel      small n: one launch, one element per thread, every path inline.
//   sincosf_kernel            FP32 arrays: FP64 arithmetic inside, 12 B/element.
//   zero_counters_kernel      clears the worklist header; first of the three launches of a hybrid call, the other two
//                             follow it with programmatic dependent launch.
//
// Data layout in HBM: three contiguous FP64 arrays (in, sin, cos), optional int64 k array (test hook), and -- only
// for launches that use the global worklist -- a workspace = { kSegments counters, 128 B apart }
// { kSegments x seg_cap u32 element indices }.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "fabe13_core.h"
#include "fabe13_internal.h"

namespace fabe13 {
namespace {

// Blocks of 128 threads (a tile = 512 elements): measured against 256 and 64 in one process
// (profiles/r02_ab_block_threads_*.txt) -- 1-3 % faster than 256 on plain data at every size (287 vs 270-283 Gelem/s
// at 1e9, 284 vs 277-282 on the 8-GPU slice, 282 vs 274-277 at 1e8) and 2-4 % on sparse mixes, the same on the dense
// Payne-Hanek mix; 64-thread blocks lose a quarter (the block scheduler cannot launch them fast enough).
#ifndef FABE13_BLOCK_THREADS
#define FABE13_BLOCK_THREADS 128
#endif
constexpr int kThreads = FABE13_BLOCK_THREADS;
// resident blocks per SM the fused kernel is compiled for (register cap 65536 / 1536 threads -> 40); the Psi core is
// FP64-pipe bound, needs ~60 registers and is compiled for 1024 resident threads
#ifndef FABE13_FUSED_MIN_BLOCKS
#define FABE13_FUSED_MIN_BLOCKS (1536 / FABE13_BLOCK_THREADS)
#endif
// Worklist segments: blocks append to segment blockIdx % kSegments, so same-address atomics (which serialise in
// L2) are spread over kSegments counters placed one 128-byte line apart.
constexpr int kSegments = 32;
constexpr int kCounterStrideWords = 32;
constexpr int kWorklistHeaderBytes = kSegments * kCounterStrideWords * 4;

// 2/pi as a stream of 32-bit words, most significant first (fabe13_payne_hanek reads seven consecutive words)
__constant__ uint32_t c_ph_stream[FABE13_PH_STREAM_WORDS] = FABE13_PH_STREAM;

// Programmatic dependent launch: the three launches of one call (zero the counters, stream, second pass) are queued
// with the programmatic-serialisation attribute, so each one is staged while its predecessor still runs and the
// 2-4 us launch gaps between dependent kernels disappear.  A kernel launched that way must not touch what its
// predecessor produces before this wait (all prerequisite grids complete, their writes visible).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ... and the other half: the kernel queued behind this one may be staged (its blocks made resident, parked at their own
// griddepcontrol.wait) as soon as every block of this grid has started, i.e. during this grid's last wave, instead
// of after its last block has exited.  It still sees none of this grid's writes before its wait returns.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__global__ void __launch_bounds__(1024) zero_counters_kernel(unsigned int* counters) {
    counters[threadIdx.x] = 0;  // kSegments * kCounterStrideWords == 1024 words
}
static_assert(kSegments * kCounterStrideWords == 1024, "zero_counters_kernel clears exactly one 1024-word header");

struct Worklist {
    unsigned int* counters;  // kSegments counters, kCounterStrideWords apart; may exceed seg_cap (= overflow)
    uint32_t* entries;       // segment s owns entries[s*seg_cap, (s+1)*seg_cap)
    uint32_t seg_cap;
};

// which outputs a launch produces: both (the reference's fabe13_sincos), or one of them (16 B/element)
// (or a derived function of the pair -- tan, cot, sinc -- written to the first output array; fused kernel only)
enum { OUT_BOTH = 0, OUT_SIN = 1, OUT_COS = 2, OUT_TAN = FABE13_OP_TAN, OUT_COT = FABE13_OP_COT, OUT_SINC = FABE13_OP_SINC };
__host__ __device__ constexpr bool out_first(int out) { return out != OUT_COS; }                    // writes sin_out
__host__ __device__ constexpr bool out_second(int out) { return out == OUT_BOTH || out == OUT_COS; }  // writes cos_out
__host__ __device__ constexpr bool out_derived(int out) { return out >= OUT_TAN; }

struct StreamArgs {
    const double* in;
    double* sin_out;   // null when only cosines are produced
    double* cos_out;   // null when only sines are produced
    long long* k_out;  // null unless WITH_K
    uint32_t n;        // elements in this launch (<= 2^31)
    uint32_t head;     // scalar elements before the vector body
    uint32_t nvec;     // vectors in the body
    uint32_t ntiles;   // tiles of kThreads*UNROLL vectors
    uint32_t tiles_per_block;  // consecutive tiles one block takes per grid-stride step
    bool pdl;                  // launched with programmatic stream serialisation
    uint32_t local_min;        // warps with at least this many Payne-Hanek lanes reduce them locally (kLocalMin;
                               // 0xFFFFFFFF = never: everything goes to the global worklist, option worklist=1)
    uint32_t dense_hint;       // first-element votes (of 32) that send a warp down the dense route (kDenseHint; 33 = never)
    uint32_t direct_max;       // a warp with at most this many Payne-Hanek lanes lets each lane reduce its own (kDirectMax)
    Worklist wl;
};

// ---- 64/128/256-bit streaming accesses.  Loads do not allocate in L1 (each byte is read once); stores are
//      evict-first.  Plain (coherent) loads, not .nc: in == sin_out is allowed.
template <int VEC>
__device__ __forceinline__ void load_stream(const double* p, uint64_t (&b)[VEC]);
template <>
__device__ __forceinline__ void load_stream<4>(const double* p, uint64_t (&b)[4]) {
    asm volatile("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(b[0]), "=l"(b[1]), "=l"(b[2]), "=l"(b[3])
                 : "l"(p));
}
template <>
__device__ __forceinline__ void load_stream<2>(const double* p, uint64_t (&b)[2]) {
    asm volatile("ld.global.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(b[0]), "=l"(b[1]) : "l"(p));
}
template <>
__device__ __forceinline__ void load_stream<1>(const double* p, uint64_t (&b)[1]) {
    asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(b[0]) : "l"(p));
}

template <int VEC>
__device__ __forceinline__ void store_stream(void* p, const uint64_t (&b)[VEC]);
template <>
__device__ __forceinline__ void store_stream<4>(void* p, const uint64_t (&b)[4]) {
    asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3])
                 : "memory");
}
template <>
__device__ __forceinline__ void store_stream<2>(void* p, const uint64_t (&b)[2]) {
    asm volatile("st.global.cs.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(b[0]), "l"(b[1]) : "memory");
}
template <>
__device__ __forceinline__ void store_stream<1>(void* p, const uint64_t (&b)[1]) {
    asm volatile("st.global.cs.b64 [%0], %1;" ::"l"(p), "l"(b[0]) : "memory");
}

// 2^45 <= |v| < Inf: an argument parked in the sin array by the first pass (no sine has that magnitude)
__device__ __forceinline__ bool is_stashed(uint64_t bits) { return fabe13_is_huge(fabe13_abs_hi(bits)); }

// ---- the fused kernel's form of the fast path: the quadrant rotation only, NO overrides.  For ordinary data no
//      lane needs one, and the warp finds that out with the vote it takes anyway -- so the override selects (and the
//      values they select from) leave the hot path: ~7 of 62 instructions per element.  A warp that does hold a tiny,
//      special or Payne-Hanek-range lane patches those lanes with patch_lane() before it stores.
template <int CORE, bool WITH_K, int OUT>
__device__ __forceinline__ bool plain_lane(uint64_t xb, uint64_t& sb, uint64_t& cb, long long& k) {
    const double x = __longlong_as_double((long long)xb);
    double biased;
    const double kd = fabe13_quadrant_index(x, &biased);
    const double r = fabe13_cody_waite(x, kd);
    double sr, cr;
    fabe13_core<CORE>(r, &sr, &cr);
    const uint32_t ahi = fabe13_abs_hi(xb);
    fabe13_rotate(sr, cr, fabe13_q_from_biased(biased), &sb, &cb);
    if (out_derived(OUT)) sb = fabe13_derived_plain<OUT>(xb, sb, cb);  // (lanes that are not plain are patched later)
    if (WITH_K) k = fabe13_is_special(ahi) ? 0 : fabe13_k_from_biased(biased);
    return fabe13_is_plain(ahi);
}

// what a lane that is not plain must hold instead: NaN/Inf -> the canonical quiet NaN, |x| <= 2^-27 -> (x, 1) [or the
// derived function of that pair], Payne-Hanek range -> x itself, parked in the first output the launch writes
template <int OUT>
__device__ __forceinline__ void patch_lane(uint64_t xb, uint64_t& sb, uint64_t& cb) {
    const uint32_t ahi = fabe13_abs_hi(xb);
    if (fabe13_is_plain(ahi)) return;
    if (fabe13_is_special(ahi)) {
        sb = cb = FABE13_QNAN_BITS;
    } else if (fabe13_is_huge(ahi)) {
        if (OUT == OUT_COS) cb = xb; else sb = xb;
    } else {
        sb = out_derived(OUT) ? fabe13_derived<OUT>(xb, xb, FABE13_ONE_BITS) : xb;
        cb = FABE13_ONE_BITS;
    }
}

// ---- the default hot path: streaming tile + warp-local worklist (+ global worklist for stragglers) --------------
// Phase 1 is the streaming tile: every lane runs the fast path and the vectors are stored
// (a Payne-Hanek lane parks x itself in its output slot).  One warp ballot per tile says whether any lane was not
// plain; only then (rare for ordinary data) the warp counts its Payne-Hanek lanes and
//   * count >= kLocalMin: compacts (x, offset-in-tile) pairs into its own slice of shared memory and strides over
//     that list with all lanes busy -- Payne-Hanek reduction, core, rotation, two 8-byte stores that merge in L2
//     with the lines the warp has just written.  No block-wide barrier, no scratch in HBM, cost proportional to
//     the number of such lanes (rounded up to 32 per warp);
//   * count < kLocalMin (sparse stragglers; hybrid launches only): appends their indices to the global worklist
//     (one atomic per warp on one of 32 counters), and the second-pass kernel reduces them densely packed.  A lone lane would
//     otherwise cost the warp a whole ~250-instruction round at 1/32 utilisation and, worse, keep the block's
//     registers and shared memory occupied while it runs (throughput here follows occupancy closely: 5 instead of
//     6 resident blocks costs 25 %).  At most kLocalMin-1 of a warp's 128 elements are ever pushed, so the global
//     list cannot overflow its n/32 entries in the default geometry (and the rescan covers any other geometry).
// Measured (tools/tune_thresholds.py, profiles/r01_tune_thresholds.md): with the ~150-instruction round of the current
// reduction only a LONE lane is cheaper on the global list -> kLocalMin = 2 (option local_min).
constexpr uint32_t kLocalMin = 2;
// A warp with only a few Payne-Hanek lanes (at most kDirectMax of its 128 elements) skips the list altogether: the
// lanes that hold one reduce it themselves, reading the 2/pi words from the constant bank (with one or two lanes
// active a divergent constant read costs nothing) -- no prefix sum, no compaction, no table staged in shared memory.
// Measured (tools/tune_stragglers.py, profiles/r02_tune_stragglers.md): 1 % random mix 233 (list only) -> 250 Gelem/s with
// kDirectMax = 2, 237 with 4 or 8; the global list of the hybrid sequence reaches 213-223 on the same data.
constexpr uint32_t kDirectMax = 2;
// lanes (of 32) whose first element must be non-plain for the warp to take the dense route
constexpr uint32_t kDenseHint = 16;

template <int CORE, int OUT>
__device__ __forceinline__ void store_huge(const StreamArgs& a, size_t e, uint64_t xb, const uint32_t* table) {
    uint64_t sb, cb;
    const int q = fabe13_sincos_huge<CORE>(xb, table, &sb, &cb);
    if (out_derived(OUT)) sb = fabe13_derived<OUT>(xb, sb, cb);
    if (out_first(OUT)) a.sin_out[e] = __longlong_as_double((long long)sb);
    if (out_second(OUT)) a.cos_out[e] = __longlong_as_double((long long)cb);
    if (a.k_out) a.k_out[e] = q;
}

template <int UNROLL, int VEC>
struct FusedShared {
    static constexpr int kWarpElems = 32 * UNROLL * VEC;  // elements one warp owns per tile
    static constexpr int kWarps = kThreads / 32;
    // per-warp worklists: no block-wide barrier anywhere, a warp only ever touches its own slice
    uint64_t x[kWarps][kWarpElems];               // arguments of the warp's Payne-Hanek lanes, compacted
    uint16_t off[kWarps][kWarpElems];             // their element offsets inside the tile
};

// what a tiny (special = false) or NaN/Inf (special = true) argument returns: reference fabe13.c:597-600
__device__ __noinline__ ulonglong2 override_values(uint64_t x_bits, bool special) {
    ulonglong2 o;
    o.x = special ? FABE13_QNAN_BITS : x_bits;
    o.y = special ? FABE13_QNAN_BITS : FABE13_ONE_BITS;
    return o;
}

// fabe13_sincos_one for a warp running in lockstep (all 32 lanes call it together): same reduction, core and rotation;
// the tiny / NaN / Inf overrides -- selects that fabe13_sincos_one applies to every element -- are entered only if
// a warp vote finds a lane that needs one.  Same bits as fabe13_sincos_one for every input.
template <int CORE>
__device__ __forceinline__ void sincos_one_lockstep(uint64_t x_bits, const uint32_t* ph_stream, uint64_t* s_bits,
                                                    uint64_t* c_bits, int64_t* k_out) {
    const uint32_t ahi = fabe13_abs_hi(x_bits);
    double r;
    uint32_t q;
    if (fabe13_is_huge(ahi)) {
        r = fabe13_payne_hanek(x_bits, ph_stream, &q);
        *k_out = (int64_t)q;
    } else {
        const double x = fabe13_from_bits(x_bits);
        double biased;
        const double kd = fabe13_quadrant_index(x, &biased);
        r = fabe13_cody_waite(x, kd);
        q = fabe13_q_from_biased(biased);
        *k_out = fabe13_is_special(ahi) ? 0 : fabe13_k_from_biased(biased);
    }
    double sr, cr;
    fabe13_core<CORE>(r, &sr, &cr);
    fabe13_rotate(sr, cr, q, s_bits, c_bits);
    const bool special = fabe13_is_special(ahi);
    const bool override_ = special || fabe13_is_tiny(ahi);
    if (__any_sync(0xFFFFFFFFu, override_)) {
        // (a call, because ptxas turns a short block into predicated selects again, vote or no vote)
        const ulonglong2 o = override_values(x_bits, special);
        if (override_) {
            *s_bits = o.x;
            *c_bits = o.y;
        }
    }
}

// The dense-warp route (see fused_tile): one 256-bit vector of every lane through the every-range routine, element by
// element, all 32 lanes in lockstep.  Not inlined: its register allocation must not leak into the streaming path
// (inlined, ptxas spilled the fast path's live vectors around it).  Arguments and results travel in registers (the
// structs are passed and returned by value), and the element loop is unrolled: round 2's first version parked the
// vector in shared memory and ran a rolled loop over it, and once the fast reduction had taken the integer product
// out of the rounds, those twelve shared-memory wavefronts per element, next to the 32 of the table gather, were what
// bound the route -- the L1 data pipe was 96 % busy (profiles/r02_ncu_fused_kernel_c4_rolled_rounds.txt).
// (The 256-bit stores stay with the caller: inside a non-inlined function ptxas 12.9 once turned st.global.cs.v4.b64
// into a 64-bit store of the first element in one of the two kernels it was cloned into.)
// elements per call of the dense routine: the whole vector.  (While its fourth element runs, the routine holds twelve
// result registers, and ptxas parks three words in local memory once per vector to fit the kernel's 40 registers:
// six 32-bit local accesses per four elements.  Two calls of two elements each were tried and are worse -- the extra
// argument shuffling makes both the caller and the routine spill.)
#ifndef FABE13_DENSE_PIECE
#define FABE13_DENSE_PIECE 4
#endif
constexpr int kDensePiece = FABE13_DENSE_PIECE;
template <int VEC>
struct DenseIn {
    uint64_t x[VEC];
};
template <int VEC>
struct DenseOut {
    uint64_t s[VEC], c[VEC];
};

template <int CORE, int VEC, bool WITH_K, int OUT>
__device__ __forceinline__ DenseOut<VEC> dense_vector_body(const DenseIn<VEC>& in, long long* k_dst) {
    DenseOut<VEC> o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        uint64_t s1, c1;
        int64_t k;
        sincos_one_lockstep<CORE>(in.x[j], c_ph_stream, &s1, &c1, &k);
        if (out_derived(OUT)) s1 = fabe13_derived<OUT>(in.x[j], s1, c1);
        o.s[j] = s1;
        o.c[j] = c1;
        if (WITH_K && k_dst) k_dst[j] = k;
    }
    return o;
}
// (two entry points, so that the ordinary one does not carry a pointer it never uses through its 40 registers)
template <int CORE, int VEC, int OUT>
__device__ __noinline__ DenseOut<VEC> dense_vector(const DenseIn<VEC> in) {
    return dense_vector_body<CORE, VEC, false, OUT>(in, nullptr);
}
template <int CORE, int VEC, int OUT>
__device__ __noinline__ DenseOut<VEC> dense_vector_k(const DenseIn<VEC> in, long long* k_dst) {
    return dense_vector_body<CORE, VEC, true, OUT>(in, k_dst);
}

// one tile of kThreads*UNROLL vectors: stream it, then deal with its Payne-Hanek lanes
template <int CORE, int VEC, int UNROLL, bool WITH_K, int OUT>
__device__ __forceinline__ void fused_tile(const StreamArgs& a, uint32_t tile, FusedShared<UNROLL, VEC>& sh) {
    constexpr int kTileElems = kThreads * UNROLL * VEC;
    static_assert(kTileElems <= 65536, "offsets inside a tile are stored as 16-bit values");
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const double* in = a.in + a.head;
    double* so = out_first(OUT) ? a.sin_out + a.head : nullptr;
    double* co = out_second(OUT) ? a.cos_out + a.head : nullptr;
    auto& s_x = sh.x;
    auto& s_off = sh.off;
    const uint32_t v0 = tile * (kThreads * UNROLL) + threadIdx.x;
    uint64_t xb[UNROLL][VEC];
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint32_t v = v0 + u * kThreads;
        live[u] = v < a.nvec;
        if (live[u]) {
            load_stream<VEC>(in + (size_t)v * VEC, xb[u]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) xb[u][j] = 0x3FF0000000000000ull;  // 1.0: a plain lane
        }
    }
    bool all_plain = true, dense = false;
    uint64_t sb[UNROLL][VEC], cb[UNROLL][VEC];
    long long kk[UNROLL][VEC];
    // Dense-warp route.  One early vote on each lane's first element (its plain/non-plain predicate is needed by
    // the fast path anyway): if most of them are not plain, the tile is presumably full of Payne-Hanek arguments,
    // and running the fast path first only to redo nearly every lane would waste a quarter of the work.  Such a
    // warp hands all its elements to the every-range routine instead, one element slot of every lane per round
    // (dense_vector: out of line, vector in and out in registers), and the results leave as the same two 256-bit
    // stores per vector as on the fast path (scattered 8-byte stores of a dense tile are bound by L2 sector
    // transactions: measured 98 instead of 120 Gelem/s).  Lanes split for the reduction only (large-argument form vs
    // Cody-Waite) and share core and fix-up; a round in which no lane holds a large argument skips that branch
    // altogether, so interleaved data (every 4th element huge, say) costs one such reduction per tile, not four.
    // A wrong guess costs nothing but time: the routine is complete for every input.
    if (__popc(__ballot_sync(0xFFFFFFFFu, !fabe13_is_plain(fabe13_abs_hi(xb[0][0])))) >= (int)a.dense_hint) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            constexpr int kPiece = VEC < kDensePiece ? VEC : kDensePiece;
#pragma unroll
            for (int h = 0; h < VEC; h += kPiece) {
                DenseIn<kPiece> in;
#pragma unroll
                for (int j = 0; j < kPiece; ++j) in.x[j] = xb[u][h + j];
                DenseOut<kPiece> o;
                if (WITH_K)
                    o = dense_vector_k<CORE, kPiece, OUT>(in, live[u] ? a.k_out + a.head + (size_t)(v0 + u * kThreads) * VEC + h : nullptr);
                else
                    o = dense_vector<CORE, kPiece, OUT>(in);
#pragma unroll
                for (int j = 0; j < kPiece; ++j) {
                    sb[u][h + j] = o.s[j];
                    cb[u][h + j] = o.c[j];
                }
            }
        }
        dense = true;
    } else {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) all_plain &= plain_lane<CORE, WITH_K, OUT>(xb[u][j], sb[u][j], cb[u][j], kk[u][j]);
        }
    }
    // warp ballot: which lanes hold anything but plain arguments?  (zero for ordinary data: one VOTE per tile)
    const bool rare = __ballot_sync(0xFFFFFFFFu, !all_plain) != 0;
    if (rare) {  // tiny, special or Payne-Hanek-range lanes: give them their values / placeholders
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) patch_lane<OUT>(xb[u][j], sb[u][j], cb[u][j]);
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        if (live[u]) {
            const size_t e = (size_t)(v0 + u * kThreads) * VEC;
            if (out_first(OUT)) store_stream<VEC>(so + e, sb[u]);
            if (out_second(OUT)) store_stream<VEC>(co + e, cb[u]);
            if (WITH_K && !dense) {  // (the dense route has written its k values itself)
#pragma unroll
                for (int j = 0; j < VEC; ++j) a.k_out[a.head + e + j] = kk[u][j];
            }
        }
    }
    if (rare) {
        uint32_t n_huge = 0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) n_huge += is_stashed(xb[u][j]) ? 1u : 0u;
        }
        const uint32_t count_all = __reduce_add_sync(0xFFFFFFFFu, n_huge);
        const bool to_global = a.wl.counters != nullptr && count_all < a.local_min;  // hybrid launches: stragglers go to the second pass
        if (count_all != 0 && count_all <= a.direct_max && !to_global) {
            // a few stragglers: every lane that holds one reduces it itself.  The lane parks its vector in its own
            // shared-memory slots first (so that the loop below can index it and the registers are free for the
            // reduction); its result stores follow its own placeholder stores in program order.
            if (n_huge != 0) {
                uint32_t mask = 0;
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        s_x[warp][(u * VEC + j) * 32 + lane] = xb[u][j];
                        mask |= (is_stashed(xb[u][j]) ? 1u : 0u) << (u * VEC + j);
                    }
                }
#pragma unroll 1
                for (; mask != 0; mask &= mask - 1) {
                    const uint32_t r = (uint32_t)__ffs((int)mask) - 1u;
                    const size_t e = (size_t)a.head + ((size_t)v0 + (size_t)(r / VEC) * kThreads) * VEC + r % VEC;
                    store_huge<CORE, OUT>(a, e, s_x[warp][r * 32 + lane], c_ph_stream);
                }
            }
            __syncwarp();  // the slots are reused by the warp's next tile
        } else if (count_all != 0) {
            uint32_t incl = n_huge;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            const uint32_t count = count_all;
            uint32_t slot = incl - n_huge;
            if (!to_global) {
                // compact this warp's Payne-Hanek lanes into its shared-memory list ...
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        if (is_stashed(xb[u][j])) {
                            s_x[warp][slot] = xb[u][j];
                            s_off[warp][slot] = (uint16_t)((u * kThreads + threadIdx.x) * VEC + j);
                            ++slot;
                        }
                    }
                }
                // ... and reduce them with all lanes busy.  __syncwarp orders the list writes before the
                // reads, and the placeholder vector stores above before the 8-byte result stores below.
                __syncwarp();
                const size_t tile_base = (size_t)a.head + (size_t)tile * kTileElems;
                for (uint32_t i = lane; i < count; i += 32)
                    store_huge<CORE, OUT>(a, tile_base + s_off[warp][i], s_x[warp][i], c_ph_stream);
                __syncwarp();  // the list is reused by the warp's next tile
            } else {
                // a few stragglers: leave them, parked in their output slots, to the second-pass kernel
                // (the counters have been zeroed: this grid waited for its predecessors at its top)
                const uint32_t seg = blockIdx.x % kSegments;
                uint32_t base = 0;
                if (lane == 31) base = atomicAdd(a.wl.counters + seg * kCounterStrideWords, count);
                slot += __shfl_sync(0xFFFFFFFFu, base, 31);
                uint32_t* seg_entries = a.wl.entries + (size_t)seg * a.wl.seg_cap;
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        if (is_stashed(xb[u][j])) {
                            if (slot < a.wl.seg_cap) seg_entries[slot] = a.head + (v0 + u * kThreads) * VEC + j;
                            ++slot;
                        }
                    }
                }
            }
        }
    }
}

// ONE_TILE: block b owns tile b (the default geometry: no loop state to keep in registers); otherwise blocks stride
// over chunks of tiles_per_block consecutive tiles.
template <int CORE, int VEC, int UNROLL, bool WITH_K, int OUT, bool ONE_TILE>
__global__ void __launch_bounds__(kThreads, CORE == FABE13_CORE_PSI ? 1024 / kThreads : FABE13_FUSED_MIN_BLOCKS)
    sincos_fused_kernel(const StreamArgs a) {
    __shared__ FusedShared<UNROLL, VEC> sh;
    const uint32_t lane = threadIdx.x & 31u;
    // Every launch of this kernel carries the programmatic-serialisation attribute (option pdl): the kernel queued behind
    // it -- the second pass of a hybrid call, or the next call's streaming kernel -- is staged during this grid's last
    // wave; and this grid itself touches nothing before everything queued ahead of it has completed.
    grid_launch_dependents();
    grid_dependency_wait();
    if (ONE_TILE) {
        fused_tile<CORE, VEC, UNROLL, WITH_K, OUT>(a, blockIdx.x, sh);
    } else {
        for (uint32_t chunk = blockIdx.x; chunk * a.tiles_per_block < a.ntiles; chunk += gridDim.x) {
            const uint32_t tile_end = min(a.ntiles, (chunk + 1) * a.tiles_per_block);
            for (uint32_t tile = chunk * a.tiles_per_block; tile < tile_end; ++tile)
                fused_tile<CORE, VEC, UNROLL, WITH_K, OUT>(a, tile, sh);
        }
    }

    // edges: the < VEC elements in front of the aligned body and the < VEC behind it (block 0, first warp)
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        const uint32_t body_end = a.head + a.nvec * VEC;
        const uint32_t n_edge = a.head + (a.n - body_end);
        if (lane < n_edge) {
            const uint32_t idx = lane < a.head ? lane : body_end + (lane - a.head);
            const uint64_t xb = (uint64_t)__double_as_longlong(a.in[idx]);
            uint64_t sb, cb;
            int64_t k;
            fabe13_sincos_one<CORE>(xb, c_ph_stream, &sb, &cb, &k);
            if (out_derived(OUT)) sb = fabe13_derived<OUT>(xb, sb, cb);
            if (out_first(OUT)) a.sin_out[idx] = __longlong_as_double((long long)sb);
            if (out_second(OUT)) a.cos_out[idx] = __longlong_as_double((long long)cb);
            if (WITH_K) a.k_out[idx] = k;
        }
    }
}

struct SecondPassArgs {
    double* sin_out;   // either may be null (single-output launches); the stash is in sin_out if present, else cos_out
    double* cos_out;
    long long* k_out;
    uint32_t n;
    Worklist wl;
};

template <int CORE>
__device__ __forceinline__ void reduce_stashed(const SecondPassArgs& a, uint32_t idx, uint64_t xb, const uint32_t* table) {
    uint64_t sb, cb;
    const int q = fabe13_sincos_huge<CORE>(xb, table, &sb, &cb);  // the caller has checked is_stashed(xb)
    if (a.sin_out) a.sin_out[idx] = __longlong_as_double((long long)sb);
    if (a.cos_out) a.cos_out[idx] = __longlong_as_double((long long)cb);
    if (a.k_out) a.k_out[idx] = q;
}

__device__ __forceinline__ void stage_table(uint32_t* table) {
    if (threadIdx.x < FABE13_PH_STREAM_WORDS) table[threadIdx.x] = c_ph_stream[threadIdx.x];
    __syncthreads();
}

// Second pass.  (1) The compacted worklist: every active lane runs Payne-Hanek.  The 2/pi table is staged in
// shared memory because lanes index it by their own exponent (a divergent __constant__ read would serialise).
// (2) Only if some segment overflowed: rescan the sin array for arguments still stashed there.  (1) and (2) run
// concurrently across blocks, so both check that the slot still holds a stashed argument; if two threads race on
// one slot they read the same x and write the same bits.
template <int CORE>
__global__ void __launch_bounds__(kThreads) sincos_second_pass_kernel(const SecondPassArgs a) {
    __shared__ uint32_t table[FABE13_PH_STREAM_WORDS];
    __shared__ uint32_t seg_start[kSegments + 1];
    __shared__ int overflowed;
    grid_launch_dependents();  // (the next call's streaming kernel may be staged behind this one)
    grid_dependency_wait();    // everything below reads what the stream kernel wrote
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        const uint32_t raw = a.wl.counters[lane * kCounterStrideWords];
        const bool over = raw > a.wl.seg_cap;
        uint32_t incl = over ? a.wl.seg_cap : raw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += t;
        }
        seg_start[lane + 1] = incl;
        if (lane == 0) seg_start[0] = 0;
        const unsigned any_over = __ballot_sync(0xFFFFFFFFu, over);
        if (lane == 0) overflowed = any_over != 0;
    }
    stage_table(table);  // includes the __syncthreads that publishes seg_start / overflowed
    const uint32_t total = seg_start[kSegments];
    if (total == 0 && !overflowed) return;

    const uint32_t gtid = blockIdx.x * kThreads + threadIdx.x;
    const uint32_t stride = gridDim.x * kThreads;
    const double* stash = a.sin_out ? a.sin_out : a.cos_out;
    // Both loops are software-pipelined: the loads of the next iteration (index two steps ahead, argument one step
    // ahead) are in flight while the ~250-instruction reduction of the current element runs.
    auto entry = [&](uint32_t i) -> uint32_t {
        if (i >= total) return 0;
        uint32_t seg = 0;  // largest seg with seg_start[seg] <= i
#pragma unroll
        for (int step = kSegments / 2; step > 0; step >>= 1)
            if (seg_start[seg + step] <= i) seg += step;
        return a.wl.entries[(size_t)seg * a.wl.seg_cap + (i - seg_start[seg])];
    };
    auto load_bits = [&](uint32_t idx) -> uint64_t { return (uint64_t)__double_as_longlong(stash[idx]); };
    if (gtid < total) {
        uint32_t idx0 = entry(gtid), idx1 = entry(gtid + stride);
        uint64_t x0 = load_bits(idx0);
        for (uint32_t i = gtid; i < total; i += stride) {
            const uint32_t idx2 = entry(i + 2 * stride);
            const uint64_t x1 = (i + stride < total) ? load_bits(idx1) : 0;
            // still a stashed argument?  (when the rescan below is active another thread may already have reduced it)
            if (is_stashed(x0)) reduce_stashed<CORE>(a, idx0, x0, table);
            idx0 = idx1;
            idx1 = idx2;
            x0 = x1;
        }
    }
    if (overflowed && gtid < a.n) {
        uint64_t x0 = load_bits(gtid);
        for (uint32_t i = gtid; i < a.n; i += stride) {
            const uint64_t x1 = (i + stride < a.n) ? load_bits(i + stride) : 0;
            if (is_stashed(x0)) reduce_stashed<CORE>(a, i, x0, table);
            x0 = x1;
        }
    }
}

// Small n: launch latency is everything, so one kernel does every path inline.
template <int CORE>
__global__ void __launch_bounds__(kThreads) sincos_inline_kernel(const double* in, double* sin_out, double* cos_out,
                                                                  long long* k_out, uint32_t n, int op) {
    __shared__ uint32_t table[FABE13_PH_STREAM_WORDS];
    stage_table(table);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const uint64_t xb = (uint64_t)__double_as_longlong(in[i]);
        uint64_t sb, cb;
        int64_t k;
        fabe13_sincos_one<CORE>(xb, table, &sb, &cb, &k);
        if (op != FABE13_OP_SINCOS) sb = fabe13_derived_rt(op, xb, sb, cb);
        if (sin_out) sin_out[i] = __longlong_as_double((long long)sb);
        if (cos_out) cos_out[i] = __longlong_as_double((long long)cb);
        if (k_out) k_out[i] = k;
    }
}

// ---- FP32 arrays: sincosf -----------------------------------------------------------------------------------------
// The reference has no FP32 path (roadmap only, README.md:225); this is the array path's own algorithm applied to
// float data: each element is widened to FP64 (exact), reduced and evaluated in FP64 -- plain arguments by a shorter
// route than the FP64 entries need (fabe13_core.h section 6: 16 FP64 instructions, good to 3e-12), everything else by
// the every-range routine -- and both results are rounded once to FP32: within 0.5001 ulp of the true values, the
// correctly rounded float in all but ~1e-4 of the cases.  12 bytes of HBM traffic per element: one 256-bit load of eight floats
// per thread, two 256-bit stores; tiles in address order, one per block.
#ifndef FABE13_SINCOSF_MIN_BLOCKS
#define FABE13_SINCOSF_MIN_BLOCKS (1536 / FABE13_BLOCK_THREADS)
#endif
struct SincosfArgs {
    const float* in;
    float* sin_out;
    float* cos_out;
    uint32_t n;      // elements in this launch
    uint32_t head;   // scalar elements before the 32-byte aligned body
    uint32_t nvec;   // 8-float vectors in the body (0 when the three arrays are not congruent modulo 32 bytes)
};

struct BitsPair {
    uint64_t s, c;
};
// the every-range routine, out of line: eight inlined copies per thread would set the kernel's register count
template <int CORE>
__device__ __noinline__ BitsPair sincos_any_range(uint64_t xb) {
    BitsPair r;
    int64_t k;
    fabe13_sincos_one<CORE>(xb, c_ph_stream, &r.s, &r.c, &k);
    return r;
}

// one float: the short FP64 route for plain arguments (fabe13_sincosf_one in fabe13_core.h, restated here so that
// the rare branch is a call), results rounded once
template <int CORE>
__device__ __forceinline__ void sincosf_one(float xf, float& sf, float& cf) {
    const double x = (double)xf;
    const uint64_t xb = (uint64_t)__double_as_longlong(x);
    if (CORE == FABE13_CORE_POLY && fabe13_is_plain(fabe13_abs_hi(xb))) {
        fabe13_sincosf_one<CORE>(xf, c_ph_stream, &sf, &cf);  // the compiler keeps only the plain branch here
    } else {  // tiny, NaN/Inf, Payne-Hanek range (rare), or the Psi core
        const BitsPair r = sincos_any_range<CORE>(xb);
        sf = __double2float_rn(__longlong_as_double((long long)r.s));
        cf = __double2float_rn(__longlong_as_double((long long)r.c));
    }
}

template <int CORE>
__global__ void __launch_bounds__(kThreads, FABE13_SINCOSF_MIN_BLOCKS) sincosf_kernel(const SincosfArgs a) {
    const uint32_t v = blockIdx.x * kThreads + threadIdx.x;
    if (v < a.nvec) {
        uint64_t w[4];
        load_stream<4>(reinterpret_cast<const double*>(a.in + a.head) + (size_t)v * 4, w);
        uint64_t so[4], co[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s0, c0, s1, c1;
            sincosf_one<CORE>(__uint_as_float((uint32_t)w[j]), s0, c0);
            sincosf_one<CORE>(__uint_as_float((uint32_t)(w[j] >> 32)), s1, c1);
            so[j] = (uint64_t)__float_as_uint(s0) | ((uint64_t)__float_as_uint(s1) << 32);
            co[j] = (uint64_t)__float_as_uint(c0) | ((uint64_t)__float_as_uint(c1) << 32);
        }
        store_stream<4>(a.sin_out + a.head + (size_t)v * 8, so);
        store_stream<4>(a.cos_out + a.head + (size_t)v * 8, co);
    }
    // what the vectors do not cover: the head, the tail, or everything when the arrays are mutually misaligned
    const uint32_t body_end = a.head + a.nvec * 8;
    const uint32_t n_edge = a.head + (a.n - body_end);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_edge; i += gridDim.x * kThreads) {
        const uint32_t idx = i < a.head ? i : body_end + (i - a.head);
        float sf, cf;
        sincosf_one<CORE>(a.in[idx], sf, cf);
        a.sin_out[idx] = sf;
        a.cos_out[idx] = cf;
    }
}

// ---- bench/test hooks: the synthetic inputs of SURVEY.md section 8d, generated in place on the device.
//      Counter-based (splitmix64 of seed + global index), so any slice of any sharding reproduces the same array
//      and the host generator in oracle/fabe13_oracle.c yields identical bits.
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(kThreads) fill_uniform_kernel(double* out, size_t n, uint64_t seed, uint64_t first,
                                                                 double lo, double hi) {
    const double span = __dadd_rn(hi, -lo);
    for (size_t i = blockIdx.x * (size_t)kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const double u = __dmul_rn(__ull2double_rn(splitmix64(seed + first + i) >> 11), 0x1.0p-53);
        out[i] = __fma_rn(span, u, lo);
    }
}

__global__ void __launch_bounds__(kThreads) fill_loguniform_kernel(double* out, size_t n, uint64_t seed, uint64_t first) {
    for (size_t i = blockIdx.x * (size_t)kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const uint64_t h = splitmix64(seed + first + i);
        const uint64_t bits = ((h & 1u) << 63) | ((1023u + ((h >> 1) & 1023u)) << 52) | (splitmix64(h) >> 12);
        out[i] = __longlong_as_double((long long)bits);
    }
}

// Launch `kernel(arg)`; with `dependent` the launch carries the programmatic-stream-serialisation attribute.
template <typename Arg>
cudaError_t launch_1arg(void (*kernel)(Arg), const Arg& arg, int grid, int block, cudaStream_t stream, bool dependent) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = dependent ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, arg);
}

template <int CORE, int VEC, bool WITH_K, int OUT>
cudaError_t launch_stream_unroll(const StreamArgs& a, int unroll, int grid, cudaStream_t stream) {
    {
        const bool one_tile = (uint32_t)grid == a.ntiles && a.tiles_per_block == 1;
        if (unroll == 2) return launch_1arg(sincos_fused_kernel<CORE, VEC, 2, WITH_K, OUT, false>, a, grid, kThreads, stream, a.pdl);
        if (one_tile) return launch_1arg(sincos_fused_kernel<CORE, VEC, 1, WITH_K, OUT, true>, a, grid, kThreads, stream, a.pdl);
        return launch_1arg(sincos_fused_kernel<CORE, VEC, 1, WITH_K, OUT, false>, a, grid, kThreads, stream, a.pdl);
    }
}

template <int CORE, bool WITH_K, int OUT>
cudaError_t launch_stream_vec(const StreamArgs& a, int vec, int unroll, int grid, cudaStream_t stream) {
    switch (vec) {
        case 4: return launch_stream_unroll<CORE, 4, WITH_K, OUT>(a, unroll, grid, stream);
        case 2: return launch_stream_unroll<CORE, 2, WITH_K, OUT>(a, unroll, grid, stream);
        default: return launch_stream_unroll<CORE, 1, WITH_K, OUT>(a, unroll, grid, stream);
    }
}

template <int CORE, int OUT>
cudaError_t launch_stream_unroll_derived(const StreamArgs& a, int vec, int grid, cudaStream_t stream) {
    const bool one_tile = (uint32_t)grid == a.ntiles && a.tiles_per_block == 1;
#define FABE13_LAUNCH_DERIVED(V)                                                                                         \
    return one_tile ? launch_1arg(sincos_fused_kernel<CORE, V, 1, false, OUT, true>, a, grid, kThreads, stream, a.pdl)   \
                    : launch_1arg(sincos_fused_kernel<CORE, V, 1, false, OUT, false>, a, grid, kThreads, stream, a.pdl)
    if (vec == 4) FABE13_LAUNCH_DERIVED(4);
    if (vec == 2) FABE13_LAUNCH_DERIVED(2);
    FABE13_LAUNCH_DERIVED(1);
#undef FABE13_LAUNCH_DERIVED
}

template <int CORE>
cudaError_t launch_stream(const StreamArgs& a, int op, int vec, int unroll, int grid, cudaStream_t stream) {
    // derived functions: fused kernel only, one vector in flight per thread (unroll 1)
    if (op == FABE13_OP_TAN) return launch_stream_unroll_derived<CORE, OUT_TAN>(a, vec, grid, stream);
    if (op == FABE13_OP_COT) return launch_stream_unroll_derived<CORE, OUT_COT>(a, vec, grid, stream);
    if (op == FABE13_OP_SINC) return launch_stream_unroll_derived<CORE, OUT_SINC>(a, vec, grid, stream);
    if (a.sin_out && a.cos_out)
        return a.k_out ? launch_stream_vec<CORE, true, OUT_BOTH>(a, vec, unroll, grid, stream)
                       : launch_stream_vec<CORE, false, OUT_BOTH>(a, vec, unroll, grid, stream);
    if (a.sin_out) return launch_stream_vec<CORE, false, OUT_SIN>(a, vec, unroll, grid, stream);
    return launch_stream_vec<CORE, false, OUT_COS>(a, vec, unroll, grid, stream);
}

// Which Payne-Hanek worklist a launch over n elements uses (LaunchTuning::worklist; 0 = pick by size).
enum { WL_AUTO = 0, WL_GLOBAL = 1, WL_LOCAL = 2, WL_HYBRID = 3 };
int worklist_mode(size_t n, const LaunchTuning& t) {
    if (t.worklist != WL_AUTO) return t.worklist;
    // one launch at every size unless the caller asks for the hybrid sequence from some size on (option hybrid_min_n;
    // its two extra launches cost 3 % at 1.25e8 elements and it no longer wins on any sparse mix, see kDirectMax)
    return n >= (size_t)t.hybrid_min_n ? WL_HYBRID : WL_LOCAL;
}

uint32_t segment_capacity(size_t n, const LaunchTuning& t, int mode) {
    // hybrid: at most local_min-1 (default 1) of every 128 elements are pushed -> n/32 entries are plenty
    size_t cap = mode == WL_HYBRID ? n >> 5 : n >> t.worklist_shift;
    if (cap < 4096) cap = 4096;
    return (uint32_t)((cap + kSegments - 1) / kSegments);
}

template <int CORE>
cudaError_t launch_core(const double* d_in, double* d_sin, double* d_cos, long long* d_k, size_t n, int op,
                        const LaunchTuning& t, const DeviceInfo& dev, void* workspace, cudaStream_t stream,
                        int* launches) {
    int mode = worklist_mode(n, t);
    // tan, cot and sinc keep their Payne-Hanek lanes in the block at every size: the global list's overflow rescan
    // recognises a parked argument by its magnitude (no sine or cosine reaches 2^45), and a tangent can
    if (op != FABE13_OP_SINCOS) mode = WL_LOCAL;
    if (mode != WL_LOCAL && workspace == nullptr) mode = WL_LOCAL;  // no scratch: everything stays in the block
    if (n < (size_t)t.small_n) {
        const long long need = (long long)((n + kThreads - 1) / kThreads);
        const long long cap = (long long)dev.sm_count * 64;
        sincos_inline_kernel<CORE><<<(int)(need < cap ? need : cap), kThreads, 0, stream>>>(d_in, d_sin, d_cos, d_k, (uint32_t)n, op);
        *launches += 1;
        return cudaGetLastError();
    }
    if (d_k && !(d_sin && d_cos)) return cudaErrorInvalidValue;  // the k hook exists for the sincos entry only

    // widest vector all arrays can share: they must be congruent modulo the vector size
    const uintptr_t pi = (uintptr_t)d_in, ps = (uintptr_t)(d_sin ? d_sin : d_cos), pc = (uintptr_t)(d_cos ? d_cos : d_sin);
    int vec = 1;
    for (int v = 4; v > 1; v >>= 1) {
        const uintptr_t mask = (uintptr_t)v * 8 - 1;
        if ((pi & mask) == (ps & mask) && (pi & mask) == (pc & mask)) {
            vec = v;
            break;
        }
    }
    const uintptr_t vbytes = (uintptr_t)vec * 8;
    uint32_t head = (uint32_t)(((vbytes - (pi & (vbytes - 1))) & (vbytes - 1)) / 8);
    if (head > n) head = (uint32_t)n;
    const int unroll = (t.unroll == 2 && op == FABE13_OP_SINCOS) ? 2 : 1;  // derived functions: unroll 1 only

    StreamArgs a;
    a.in = d_in;
    a.sin_out = d_sin;
    a.cos_out = d_cos;
    a.k_out = d_k;
    a.n = (uint32_t)n;
    a.head = head;
    a.nvec = (uint32_t)((n - head) / vec);
    a.ntiles = (uint32_t)(((size_t)a.nvec + (size_t)kThreads * unroll - 1) / ((size_t)kThreads * unroll));
    // blocks_per_sm == 0: one chunk of tiles_per_block tiles per block (the block scheduler walks the array in
    // address order); otherwise a persistent grid of SMs*blocks_per_sm blocks strides over the chunks.
    a.tiles_per_block = (uint32_t)(t.tiles_per_block > 0 ? t.tiles_per_block : 1);
    long long grid = ((long long)a.ntiles + a.tiles_per_block - 1) / a.tiles_per_block;
    if (t.blocks_per_sm > 0 && grid > (long long)dev.sm_count * t.blocks_per_sm) grid = (long long)dev.sm_count * t.blocks_per_sm;
    if (grid < 1) grid = 1;

    // worklist=1 sends every Payne-Hanek lane to the global list: no local reduction, no dense route
    a.local_min = mode == WL_GLOBAL ? 0xFFFFFFFFu : (uint32_t)(t.local_min > 0 ? t.local_min : (int)kLocalMin);
    a.dense_hint = mode == WL_GLOBAL ? 33u : (uint32_t)(t.dense_hint > 0 ? t.dense_hint : (int)kDenseHint);
    a.direct_max = mode == WL_GLOBAL ? 0u : (uint32_t)(t.direct_max >= 0 ? t.direct_max : (int)kDirectMax);

    if (mode == WL_LOCAL) {  // one launch, no scratch
        a.wl.counters = nullptr;
        a.wl.entries = nullptr;
        a.wl.seg_cap = 0;
        a.pdl = t.pdl != 0;
        *launches += 1;
        return launch_stream<CORE>(a, op, vec, unroll, (int)grid, stream);
    }

    // global or hybrid: zero the counters -> streaming kernel -> second pass over the global worklist
    a.wl.counters = (unsigned int*)workspace;
    a.wl.entries = (uint32_t*)((char*)workspace + kWorklistHeaderBytes);
    a.wl.seg_cap = segment_capacity(n, t, mode);

    a.pdl = t.pdl != 0;
    cudaError_t err;
    if (a.pdl) {
        zero_counters_kernel<<<1, 1024, 0, stream>>>(a.wl.counters);  // an ordinary launch: waits for all earlier work
        err = cudaGetLastError();
    } else {
        err = cudaMemsetAsync(workspace, 0, kWorklistHeaderBytes, stream);
    }
    if (err != cudaSuccess) return err;
    err = launch_stream<CORE>(a, op, vec, unroll, (int)grid, stream);
    if (err != cudaSuccess) return err;

    SecondPassArgs b;
    b.sin_out = d_sin;
    b.cos_out = d_cos;
    b.k_out = d_k;
    b.n = (uint32_t)n;
    b.wl = a.wl;
    err = launch_1arg(sincos_second_pass_kernel<CORE>, b, dev.sm_count * t.second_pass_blocks_per_sm, kThreads, stream, a.pdl);
    *launches += a.pdl ? 3 : 2;
    return err;
}

}  // namespace

size_t workspace_bytes(size_t n, const LaunchTuning& t) {  // (an upper bound: the derived functions never need any)
    const int mode = worklist_mode(n, t);
    if (mode == WL_LOCAL) return 0;  // the warp-local worklists live in shared memory
    return (size_t)kWorklistHeaderBytes + (size_t)kSegments * segment_capacity(n, t, mode) * sizeof(uint32_t);
}

cudaError_t fill_uniform(double* d_out, size_t n, uint64_t seed, uint64_t first, double lo, double hi,
                         const DeviceInfo& dev, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    fill_uniform_kernel<<<dev.sm_count * 8, kThreads, 0, stream>>>(d_out, n, seed, first, lo, hi);
    return cudaGetLastError();
}

cudaError_t fill_loguniform(double* d_out, size_t n, uint64_t seed, uint64_t first, const DeviceInfo& dev,
                            cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    fill_loguniform_kernel<<<dev.sm_count * 8, kThreads, 0, stream>>>(d_out, n, seed, first);
    return cudaGetLastError();
}

cudaError_t launch_sincosf(const float* d_in, float* d_sin, float* d_cos, size_t n, int core, const DeviceInfo& dev,
                           cudaStream_t stream, int* launches) {
    if (n == 0) return cudaSuccess;
    if (n > ((size_t)1 << 31) || !d_in || !d_sin || !d_cos) return cudaErrorInvalidValue;
    const uintptr_t pi = (uintptr_t)d_in, ps = (uintptr_t)d_sin, pc = (uintptr_t)d_cos;
    SincosfArgs a;
    a.in = d_in;
    a.sin_out = d_sin;
    a.cos_out = d_cos;
    a.n = (uint32_t)n;
    a.head = 0;
    a.nvec = 0;
    if ((pi & 31) == (ps & 31) && (pi & 31) == (pc & 31) && (pi & 3) == 0) {  // one 32-byte phase for all three arrays
        size_t head = ((32 - (pi & 31)) & 31) / 4;
        if (head > n) head = n;
        a.head = (uint32_t)head;
        a.nvec = (uint32_t)((n - head) / 8);
    }
    const size_t edge = n - (size_t)a.nvec * 8;
    long long grid = ((long long)a.nvec + kThreads - 1) / kThreads;
    const long long edge_blocks = (long long)((edge + kThreads - 1) / kThreads);
    const long long cap = (long long)dev.sm_count * 64;
    if (grid < (edge_blocks < cap ? edge_blocks : cap)) grid = edge_blocks < cap ? edge_blocks : cap;
    if (grid < 1) grid = 1;
    if (core == FABE13_CORE_PSI)
        sincosf_kernel<FABE13_CORE_PSI><<<(unsigned)grid, kThreads, 0, stream>>>(a);
    else
        sincosf_kernel<FABE13_CORE_POLY><<<(unsigned)grid, kThreads, 0, stream>>>(a);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_sincos(const double* d_in, double* d_sin, double* d_cos, long long* d_k, size_t n, int op, int core,
                          const LaunchTuning& tuning, const DeviceInfo& dev, void* workspace, cudaStream_t stream,
                          int* launches) {
    if (n == 0) return cudaSuccess;
    if (n > ((size_t)1 << 31) || (!d_sin && !d_cos)) return cudaErrorInvalidValue;
    if (op != FABE13_OP_SINCOS && (op < FABE13_OP_TAN || op > FABE13_OP_SINC || !d_sin || d_cos || d_k)) return cudaErrorInvalidValue;
    if (core == FABE13_CORE_PSI)
        return launch_core<FABE13_CORE_PSI>(d_in, d_sin, d_cos, d_k, n, op, tuning, dev, workspace, stream, launches);
    return launch_core<FABE13_CORE_POLY>(d_in, d_sin, d_cos, d_k, n, op, tuning, dev, workspace, stream, launches);
}

}  // namespace fabe13
