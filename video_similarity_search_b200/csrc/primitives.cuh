// Declarations of the device-wide integer primitives in primitives.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slic {

// out[i] = sum(in[0..i)).  in and out may alias only when total_out == nullptr.
// total_out (device int, optional) receives the grand total.
int exclusive_scan_i32(const int* in, int* out, int64_t n, int* total_out, cudaStream_t st);

// counts[b] = #{i : keys[i] == b}; keys must lie in [0, bins).
int histogram_i32(const int* keys, int64_t n, int* counts, int64_t bins, cudaStream_t st);

// Stable sort of (key, value) by non-negative int32 key: LSD radix sort, 8- or 9-bit digits, one launch per pass
// (one-sweep: decoupled look-back over the tiles' digit histograms) + one histogram launch.
// vals_in == nullptr sorts the identity permutation (vals_out = argsort, stable).  keys_out may be nullptr when the
// sorted keys are not wanted.  keys_in/vals_in must not alias keys_out/vals_out.
int stable_sort_pairs_i32(const int* keys_in, const int* vals_in, int64_t n, int key_bits, int* keys_out,
                          int* vals_out, cudaStream_t st);

int iota_i32(int* out, int64_t n, cudaStream_t st);

}  // namespace slic
