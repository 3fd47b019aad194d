// Shared helpers for libaadg_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/aadg_b200.h"

namespace aadg {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return AADG_ECUDA;
  }
  return AADG_OK;
}

#define AADG_CUDA_TRY(expr)                                                  \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      aadg::set_error("%s: %s", #expr, cudaGetErrorString(_e));              \
      return AADG_ECUDA;                                                     \
    }                                                                        \
  } while (0)

#define AADG_REQUIRE(cond, ...)                                              \
  do {                                                                       \
    if (!(cond)) {                                                           \
      aadg::set_error(__VA_ARGS__);                                          \
      return AADG_EINVAL;                                                    \
    }                                                                        \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t size, off;
  Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  void* take(size_t bytes) {
    size_t o = align_up(off, 256);
    if (o + bytes > size) return nullptr;
    off = o + bytes;
    return base + o;
  }
};

// ---- device-side PTX helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Aligned view of the dynamic shared memory that KEEPS ITS ADDRESS SPACE: the padding is computed on the 32-bit shared
// address and added to the original pointer.  Rounding the generic pointer through uintptr_t loses the provenance and
// every access through it compiles to a generic LD.E / ST.E with 64-bit address arithmetic instead of LDS / STS.
__device__ __forceinline__ uint8_t* align_smem(uint8_t* p, uint32_t a) {
  const uint32_t s = smem_u32(p);
  return p + (((s + a - 1u) & ~(a - 1u)) - s);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP).  dst/src 16-B aligned,
// bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

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

}  // namespace aadg
