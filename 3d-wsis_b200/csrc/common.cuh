// Shared helpers for the sm_100a kernels: error plumbing, launch accounting, the 64-bit coordinate
// hash table, warp utilities.  No torch types; see include/wsis_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/wsis_b200.h"

namespace wsis {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define WSIS_CHECK(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      ::wsis::set_error(__VA_ARGS__);         \
      return 1;                               \
    }                                         \
  } while (0)

#define WSIS_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::wsis::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

// checks the launch itself (the reference checks cudaGetLastError after every launch,
// tensorview.h:90-101); stays asynchronous.
#define WSIS_LAUNCH_OK()                                                                       \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      ::wsis::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                                \
    }                                                                                          \
    ::wsis::count_launch();                                                                    \
  } while (0)

static inline cudaStream_t as_stream(wsis_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();

// Host-mapped (zero-copy) word that a kernel fills right before it traps: a trap kills the context, device memory
// becomes unreadable, but the host copy of this word still says WHICH wait gave up (wsis_debug_trap_word()).
unsigned int *trap_word_device();   // device-visible address (NULL if the mapping could not be made)

// ------------------------------------------------------------------------------------------------
// coordinate hash: key = b:16 | x:16 | y:16 | z:16, open addressing with linear probing, slots = 2^m
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ __forceinline__ unsigned long long pack_key(int b, int x, int y, int z) {
  return ((unsigned long long)(unsigned)(b & 0xFFFF) << 48) | ((unsigned long long)(unsigned)(x & 0xFFFF) << 32) |
         ((unsigned long long)(unsigned)(y & 0xFFFF) << 16) | (unsigned long long)(unsigned)(z & 0xFFFF);
}

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// returns the slot holding `key`, inserting it if absent
__device__ __forceinline__ int64_t hash_insert(unsigned long long *keys, int64_t mask, unsigned long long key) {
  int64_t s = (int64_t)(mix64(key) & (unsigned long long)mask);
  while (true) {
    unsigned long long prev = atomicCAS(keys + s, kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) return s;
    s = (s + 1) & mask;
  }
}

// returns the slot holding `key` or -1
__device__ __forceinline__ int64_t hash_find(const unsigned long long *__restrict__ keys, int64_t mask,
                                             unsigned long long key) {
  int64_t s = (int64_t)(mix64(key) & (unsigned long long)mask);
  while (true) {
    unsigned long long cur = __ldg(keys + s);
    if (cur == key) return s;
    if (cur == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
}

// tile record geometry (tilemap.cu builds the records, conv_umma.cu consumes them)
//   valid[K][4] u32 | {nU u32, amask u32, P u32, npack u32, members u16[32]} | loc[K][128] u16
__host__ __device__ __forceinline__ int rec_hdr_bytes(int K) { return 16 * K + 80; }
__host__ __device__ __forceinline__ int rec_stride_bytes(int K) { return rec_hdr_bytes(K) + 256 * K; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace wsis
