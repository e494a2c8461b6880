// fabe13_internal.h -- interface between the C-ABI shim (fabe13_shim.cu) and the kernels (fabe13_kernels.cu).
#ifndef FABE13_INTERNAL_H
#define FABE13_INTERNAL_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fabe13 {

struct LaunchTuning {
    int worklist;       // Payne-Hanek worklist: 0 = by size (local below hybrid_min_n [default: at every size], hybrid from there on);
                        // 1 = global list + second-pass kernel; 2 = warp-local lists in shared memory, one launch,
                        // no scratch; 3 = hybrid (local for warps with >= 4 such lanes, global for stragglers)
    int hybrid_min_n;   // worklist = 0: arrays of at least this many elements take the hybrid sequence
    int local_min;      // hybrid: a warp with at least this many Payne-Hanek lanes (of 128) reduces them itself (default 2)
    int dense_hint;     // first-element votes (of 32) that send a warp down the dense route (default 16; 33 = never)
    int direct_max;     // a warp with at most this many Payne-Hanek lanes (of 128) lets each lane reduce its own (default 2)
    int blocks_per_sm;  // 0 = one tile per block (default); k > 0 = persistent grid of SMs * k blocks
    int tiles_per_block;  // consecutive tiles per block per grid-stride step
    int pdl;              // 1 = queue the launches of one call with programmatic dependent launch
    int second_pass_blocks_per_sm;  // grid of the Payne-Hanek second pass
    int unroll;         // 256-bit vectors in flight per thread: 1 or 2
    int small_n;        // n below this runs the single inline kernel (no worklist, one launch)
    int worklist_shift; // worklist capacity = n >> worklist_shift entries (0 = one entry per element)
};

struct DeviceInfo {
    int device;
    int sm_count;
};

// Bytes of scratch a launch over n elements needs for its Payne-Hanek worklist (header + entries); 0 when the
// launch keeps its worklists in shared memory.
size_t workspace_bytes(size_t n, const LaunchTuning& t);

// Enqueue the sincos path for n <= 2^31 elements on `stream`.  `workspace` must hold workspace_bytes(n) bytes
// (it is zero-initialised here, on the stream).  d_k may be null.  One of d_sin / d_cos may be null: then only the
// other output is produced (16 B/element instead of 24).  op = 0 (FABE13_OP_SINCOS) for these; FABE13_OP_TAN / _COT /
// _SINC write that function of the pair to d_sin (d_cos and d_k null).  Returns the number of kernels launched via
// *launches.  Every pointer is a device pointer on the current device.
cudaError_t launch_sincos(const double* d_in, double* d_sin, double* d_cos, long long* d_k, size_t n, int op, int core,
                          const LaunchTuning& tuning, const DeviceInfo& dev, void* workspace, cudaStream_t stream,
                          int* launches);

// FP32 arrays (12 B/element): the same per-element algorithm in FP64 arithmetic, results rounded once to float.
cudaError_t launch_sincosf(const float* d_in, float* d_sin, float* d_cos, size_t n, int core, const DeviceInfo& dev,
                           cudaStream_t stream, int* launches);

// Bench/test hooks: synthetic inputs generated on the device (bit-identical to oracle/fabe13_oracle.c's fills).
cudaError_t fill_uniform(double* d_out, size_t n, uint64_t seed, uint64_t first, double lo, double hi,
                         const DeviceInfo& dev, cudaStream_t stream);
cudaError_t fill_loguniform(double* d_out, size_t n, uint64_t seed, uint64_t first, const DeviceInfo& dev,
                            cudaStream_t stream);

}  // namespace fabe13

#endif  // FABE13_INTERNAL_H
