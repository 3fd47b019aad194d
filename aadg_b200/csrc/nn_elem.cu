// HBM-bound layers of the segmentation net (bf16 NHWC, 8 channels = 16 bytes per thread access):
// batch-norm statistics / apply / backward (with fused ReLU, residual add and dropout), max-pool,
// bilinear up-sampling (align_corners=True), global average pool / broadcast, depthwise 3x3
// (dilated) convolution forward / data / weight gradient, im2col for the 3-channel stem, Adam with
// the bf16 weight re-layout.  They replace the cuDNN / ATen kernels behind smp.DeepLabV3Plus
// (models/__init__.py:17-23; search_dg.py:132,170-172; scheduler.py:10-11).
#include <cuda_bf16.h>
#include <stdlib.h>
#include <algorithm>

#include "common.cuh"

namespace aadg {
namespace nn {

typedef __nv_bfloat16 bf16;

struct V8 { float v[8]; };

__device__ __forceinline__ V8 ld8(const bf16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  V8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  return r;
}
__device__ __forceinline__ void st8(bf16* p, const V8& a) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ V8 ld8f(const float* p) {
  V8 r;
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// Philox4x32-10 -> 4 uniform words; key (seed lo, seed hi), counter (idx lo, idx hi, stream, 0)
__device__ __forceinline__ uint4 philox(uint2 key, uint4 ctr) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
// keep-mask of 8 consecutive elements starting at element index e (p = 0.5): bit i of the result
__device__ __forceinline__ unsigned int dropout_bits8(unsigned long long seed, unsigned long long e) {
  const uint4 r = philox(make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)),
                         make_uint4((unsigned int)(e >> 3), (unsigned int)(e >> 35), 0x5eedu, 0u));
  return r.x & 0xffu;
}

// grid-stride walk over (pixel p, 8-channel group g) without a division per iteration
#define AADG_FOR_PIXEL_GROUPS(P, G, p, g)                                                      \
  const long long _tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;                     \
  const long long _stride = (long long)gridDim.x * blockDim.x;                                 \
  const int _sg = (int)(_stride % (G));                                                        \
  const long long _sp = _stride / (G);                                                         \
  long long p = _tid / (G);                                                                    \
  int g = (int)(_tid % (G));                                                                   \
  for (; p < (P); p += _sp, g += _sg, (g >= (G) ? (g -= (G), ++p) : 0))

// ---- batch-norm ------------------------------------------------------------------------------------
// Thread-stationary layout for every pass over an activation tensor: block (TX, TY) with tx = 8-channel group
// and ty = pixel lane.  A thread keeps its group's per-channel coefficients in registers for the whole kernel
// and walks pixels p = blockIdx.x*TY + ty + k*gridDim.x*TY, BN_U pixels per trip with every 16-byte load
// issued before the first use (memory-level parallelism instead of per-element coefficient reloads).
constexpr int BN_U = 4;

__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ V8 unpack8(const uint4& u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  V8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  return r;
}

// per-thread partial sums a0/a1[8] of channel group threadIdx.x -> out0/out1[C].  No shared-memory atomics (a float
// atomicAdd on shared memory is a compare-and-swap loop): pixel lanes that share a warp are folded with shuffles,
// every (row, channel) partial then has exactly one writer in a [rows][C] shared array (rows * C == 2048 for the
// block shapes of reduce_block()), and each channel's column is summed and sent out with one global reduction.
template <int NACC>
__device__ __forceinline__ void block_channel_sum(int C, float* a0, float* a1, float* out0, float* out1) {
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x;
  __shared__ float s0[2048], s1[NACC > 1 ? 2048 : 1];
  int row, rows;
  if (TX < 32) {
    for (int o = TX; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], o);
        if (NACC > 1) a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], o);
      }
    }
    row = (ty * TX + tx) >> 5; rows = (TX * blockDim.y) >> 5;
    if ((((ty * TX + tx) & 31) >= TX)) row = -1;          // one writer per warp and channel group
  } else {
    row = ty; rows = blockDim.y;
  }
  if (row >= 0 && tx < (C >> 3)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0[row * C + tx * 8 + i] = a0[i];
      if (NACC > 1) s1[row * C + tx * 8 + i] = a1[i];
    }
  }
  __syncthreads();
  for (int c = ty * TX + tx; c < C; c += TX * blockDim.y) {
    float t0 = 0.f, t1 = 0.f;
    for (int r = 0; r < rows; ++r) { t0 += s0[r * C + c]; if (NACC > 1) t1 += s1[r * C + c]; }
    atomicAdd(&out0[c], t0);
    if (NACC > 1) atomicAdd(&out1[c], t1);
  }
}

__global__ void __launch_bounds__(256, 4) bn_stats_kernel(const bf16* __restrict__ x, long long P, int C, int ld,
                                                          float* sum, float* sumsq) {
  const int g = threadIdx.x;
  float a0[8] = {}, a1[8] = {};
  if (g < (C >> 3)) {
    const long long step = (long long)gridDim.x * blockDim.y;
    const bf16* xb = x + g * 8;
    for (long long p0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; p0 < P; p0 += BN_U * step) {
      uint4 xv[BN_U];
#pragma unroll
      for (int u = 0; u < BN_U; ++u) {
        const long long p = p0 + u * step;
        xv[u] = p < P ? ldg16(xb + p * ld) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < BN_U; ++u) {
        const V8 v = unpack8(xv[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[i] += v.v[i]; a1[i] = fmaf(v.v[i], v.v[i], a1[i]); }
      }
    }
  }
  block_channel_sum<2>(C, a0, a1, sum, sumsq);
}

// mean / invstd, fused scale-shift for the apply pass, running statistics (momentum 0.1, unbiased var)
__global__ void bn_finalize_kernel(float* sum, float* sumsq, const float* gamma, const float* beta,
                                   int C, float count, float eps, float momentum, float* mean, float* invstd,
                                   float* scale, float* shift, float* run_mean, float* run_var, int reset_sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = sum[c] / count;
  const float var = fmaxf(sumsq[c] / count - m * m, 0.f);
  if (reset_sums) { sum[c] = 0.f; sumsq[c] = 0.f; }     // ready for the next accumulation (no separate memset launch)
  const float is = rsqrtf(var + eps);
  mean[c] = m; invstd[c] = is;
  const float sc = gamma[c] * is;
  scale[c] = sc; shift[c] = beta[c] - m * sc;
  if (run_mean) {
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * m;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * var * (count / fmaxf(count - 1.f, 1.f));
  }
}

// y = act(x*scale + shift (+ res)) (* dropout) ; flags: 1 = relu, 2 = dropout(0.5)
__global__ void __launch_bounds__(256, 3) bn_apply_kernel(const bf16* __restrict__ x, int ldx, const float* scale,
                                                          const float* shift, const bf16* __restrict__ res, int ldr,
                                                          bf16* __restrict__ y, int ldy, long long P, int C, int flags,
                                                          unsigned long long seed, unsigned char* relu_bits) {
  const int G = C >> 3, g = threadIdx.x;
  if (g >= G) return;
  if (flags & 64) seed = *reinterpret_cast<const unsigned long long*>(seed);   // seed kept in device memory (CUDA graphs)
  const V8 sc = ld8f(scale + g * 8), sh = ld8f(shift + g * 8);
  const long long step = (long long)gridDim.x * blockDim.y;
  for (long long p0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; p0 < P; p0 += BN_U * step) {
    uint4 xv[BN_U], rv[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long p = p0 + u * step;
      if (p < P) {
        xv[u] = ldg16(x + p * ldx + g * 8);
        if (res) rv[u] = ldg16(res + p * ldr + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long p = p0 + u * step;
      if (p >= P) break;
      const long long e = p * G + g;
      V8 v = unpack8(xv[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = fmaf(v.v[i], sc.v[i], sh.v[i]);
      if (res) {
        const V8 r = unpack8(rv[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] += r.v[i];
      }
      if (flags & 1) {
        if (relu_bits) {      // one byte per 8 channels: backward reads it instead of the whole output tensor
          unsigned int bits = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) bits |= (v.v[i] > 0.f ? 1u : 0u) << i;
          relu_bits[e] = (unsigned char)bits;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = fmaxf(v.v[i], 0.f);
        if (flags & 32) {     // ReLU6
#pragma unroll
          for (int i = 0; i < 8; ++i) v.v[i] = fminf(v.v[i], 6.f);
        }
      }
      if (flags & 2) {
        const unsigned int keep = dropout_bits8(seed, (unsigned long long)e * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = (keep >> i) & 1 ? v.v[i] * 2.f : 0.f;
      }
      st8(y + p * ldy + g * 8, v);
    }
  }
}

// The gradient that reaches the batch-norm output: g = dy * relu'(.) * dropout.  Mask source by flag:
// 4 = recomputed from x exactly as forward evaluated it, 8 = bit mask (one byte per 8 channels), 1 = y tensor.
// MASK (template) = 0 none, 4, 8 or 1; the runtime flag 2 adds dropout.
template <int MASK> struct BnMaskSrc { };
template <> struct BnMaskSrc<8> { unsigned int b; };
template <> struct BnMaskSrc<1> { uint4 v; };
template <int MASK>
__device__ __forceinline__ void bn_mask_load(BnMaskSrc<MASK>& m, const bf16* y, int ldy, long long p, int G, int g) {
  if constexpr (MASK == 8) m.b = reinterpret_cast<const unsigned char*>(y)[p * G + g];
  if constexpr (MASK == 1) m.v = ldg16(y + p * ldy + g * 8);
}
template <int MASK>
__device__ __forceinline__ void bn_mask_apply(V8& d, const V8& xv, const V8& sc, const V8& sh, const BnMaskSrc<MASK>& m,
                                              int flags, unsigned long long seed, long long e) {
  const float hi = (flags & 32) ? 6.f : INFINITY;      // ReLU6: the gradient also vanishes where the output saturates
  if constexpr (MASK == 4) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = fmaf(xv.v[i], sc.v[i], sh.v[i]);
      d.v[i] = (t > 0.f && t < hi) ? d.v[i] : 0.f;
    }
  }
  if constexpr (MASK == 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d.v[i] = (m.b >> i) & 1 ? d.v[i] : 0.f;
  }
  if constexpr (MASK == 1) {
    const V8 yv = unpack8(m.v);
#pragma unroll
    for (int i = 0; i < 8; ++i) d.v[i] = (yv.v[i] > 0.f && yv.v[i] < hi) ? d.v[i] : 0.f;
  }
  if (flags & 2) {
    const unsigned int keep = dropout_bits8(seed, (unsigned long long)e * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) d.v[i] = (keep >> i) & 1 ? d.v[i] * 2.f : 0.f;
  }
}

// dbeta = sum g ; dgamma = sum g * xhat  (`beta` = the forward shift vector when flag 4 is set)
// (all backward kernels: DY2 adds a second incoming gradient tensor on load -- the two branches of a residual block
// are summed here instead of by a read-modify-write accumulation in the convolution that produced one of them)
template <int MASK, bool DY2>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const bf16* __restrict__ dy, int lddy,
                                                               const bf16* __restrict__ dy2, int lddy2,
                                                               const bf16* __restrict__ x, int ldx, const bf16* y, int ldy,
                                                               const float* mean, const float* invstd, const float* gamma,
                                                               const float* beta, long long P, int C, int flags,
                                                               unsigned long long seed, float* dgamma, float* dbeta) {
  const int G = C >> 3, g = threadIdx.x;
  float a0[8] = {}, a1[8] = {};
  if (flags & 64) seed = *reinterpret_cast<const unsigned long long*>(seed);
  if (g < G) {
    const V8 m = ld8f(mean + g * 8), is = ld8f(invstd + g * 8);
    V8 sc = {}, sh = {};
    if (MASK == 4) {
      const V8 ga = ld8f(gamma + g * 8);
      sh = ld8f(beta + g * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) sc.v[i] = ga.v[i] * is.v[i];
    }
    const long long step = (long long)gridDim.x * blockDim.y;
    for (long long p0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; p0 < P; p0 += BN_U * step) {
      uint4 dv[BN_U], xv[BN_U], ev[DY2 ? BN_U : 1];
      BnMaskSrc<MASK> mv[BN_U];
#pragma unroll
      for (int u = 0; u < BN_U; ++u) {
        const long long p = p0 + u * step;
        if (p < P) {
          dv[u] = ldg16(dy + p * lddy + g * 8);
          if (DY2) ev[u] = ldg16(dy2 + p * lddy2 + g * 8);
          xv[u] = ldg16(x + p * ldx + g * 8);
          bn_mask_load(mv[u], y, ldy, p, G, g);
        }
      }
#pragma unroll
      for (int u = 0; u < BN_U; ++u) {
        const long long p = p0 + u * step;
        if (p >= P) break;
        V8 d = unpack8(dv[u]);
        if (DY2) {
          const V8 d2 = unpack8(ev[u]);
#pragma unroll
          for (int i = 0; i < 8; ++i) d.v[i] += d2.v[i];
        }
        const V8 xx = unpack8(xv[u]);
        bn_mask_apply(d, xx, sc, sh, mv[u], flags, seed, p * G + g);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[i] = fmaf(d.v[i], xx.v[i] - m.v[i], a0[i]); a1[i] += d.v[i]; }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] *= is.v[i];
  }
  block_channel_sum<2>(C, a0, a1, dgamma, dbeta);
}

// dx = gamma*invstd * (g - dbeta/P - xhat*dgamma/P) ; optional dres = g (gradient of the residual input)
template <int MASK, bool DY2>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(const bf16* __restrict__ dy, int lddy,
                                                              const bf16* __restrict__ dy2, int lddy2,
                                                              const bf16* __restrict__ x, int ldx, const bf16* y, int ldy,
                                                              const float* mean, const float* invstd, const float* gamma,
                                                              const float* beta, const float* dgamma, const float* dbeta,
                                                              long long P, int C, int flags, unsigned long long seed,
                                                              bf16* __restrict__ dx, int lddx, bf16* dres, int lddr,
                                                              int dres_accumulate, float inv_count) {
  const int G = C >> 3, g = threadIdx.x;
  if (g >= G) return;
  if (flags & 64) seed = *reinterpret_cast<const unsigned long long*>(seed);
  const float invP = inv_count;      // 1 / (pixels the statistics were taken over): local P, or the global count (SyncBN)
  // dx = sc*g + kk + bb*(x - mean):  sc = gamma*invstd, bb = -sc*invstd*dgamma/P, kk = -sc*dbeta/P
  const V8 m = ld8f(mean + g * 8);
  V8 sc, sh = {}, bb, kk;
  {
    const V8 is = ld8f(invstd + g * 8), ga = ld8f(gamma + g * 8), dg = ld8f(dgamma + g * 8), db = ld8f(dbeta + g * 8);
    if (MASK == 4) sh = ld8f(beta + g * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc.v[i] = ga.v[i] * is.v[i];
      bb.v[i] = -sc.v[i] * is.v[i] * dg.v[i] * invP;
      kk.v[i] = -sc.v[i] * db.v[i] * invP;
    }
  }
  const long long step = (long long)gridDim.x * blockDim.y;
  for (long long p0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; p0 < P; p0 += BN_U * step) {
    uint4 dv[BN_U], xv[BN_U], ev[DY2 ? BN_U : 1];
    BnMaskSrc<MASK> mv[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long p = p0 + u * step;
      if (p < P) {
        dv[u] = ldg16(dy + p * lddy + g * 8);
        if (DY2) ev[u] = ldg16(dy2 + p * lddy2 + g * 8);
        xv[u] = ldg16(x + p * ldx + g * 8);
        bn_mask_load(mv[u], y, ldy, p, G, g);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long p = p0 + u * step;
      if (p >= P) break;
      V8 d = unpack8(dv[u]);
      if (DY2) {
        const V8 d2 = unpack8(ev[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) d.v[i] += d2.v[i];
      }
      const V8 xx = unpack8(xv[u]);
      bn_mask_apply(d, xx, sc, sh, mv[u], flags, seed, p * G + g);
      if (dres) {
        V8 r = d;
        if (dres_accumulate) {
          const V8 o = ld8(dres + p * lddr + g * 8);    // rare path: not prefetched
#pragma unroll
          for (int i = 0; i < 8; ++i) r.v[i] += o.v[i];
        }
        st8(dres + p * lddr + g * 8, r);
      }
      V8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = fmaf(bb.v[i], xx.v[i] - m.v[i], fmaf(sc.v[i], d.v[i], kk.v[i]));
      st8(dx + p * lddx + g * 8, o);
    }
  }
}

// a (+)= b, bf16 NHWC with strides (gradient fan-in of skip connections)
__global__ void add_kernel(bf16* a, int lda, const bf16* b, int ldb, long long P, int C) {
  const int G = C >> 3;
  AADG_FOR_PIXEL_GROUPS(P, G, p, g) {
    V8 x = ld8(a + p * lda + g * 8);
    const V8 yv = ld8(b + p * ldb + g * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) x.v[i] += yv.v[i];
    st8(a + p * lda + g * 8, x);
  }
}

// ---- max-pool 3x3 stride 2 pad 1 -------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const bf16* x, int N, int H, int W, int C, bf16* y, unsigned char* arg, int Ho, int Wo) {
  const int G = C >> 3;
  const long long total = (long long)N * Ho * Wo * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % G);
    long long p = e / G;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    V8 best; unsigned char bi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best.v[i] = -INFINITY; bi[i] = 0; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * 2 - 1 + r;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = ox * 2 - 1 + s;
        if (ix < 0 || ix >= W) continue;
        const V8 v = ld8(x + (((size_t)n * H + iy) * W + ix) * C + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (v.v[i] > best.v[i]) { best.v[i] = v.v[i]; bi[i] = (unsigned char)(r * 3 + s); }
      }
    }
    const size_t o = (((size_t)n * Ho + oy) * Wo + ox) * C + g * 8;
    st8(y + o, best);
    uint2 packed;
    packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(arg + o) = packed;
  }
}
// gather form, one thread per 2x2 block of input pixels (rows 2i, 2i+1; columns 2j, 2j+1) x 8 channels: the block
// is covered by exactly the four windows (i..i+1, j..j+1), each loaded once; input pixel (2i+a, 2j+b) is tap
// (r, s) = (1 + a - 2di, 1 + b - 2dj) of window (i + di, j + dj) whenever that tap index is in 0..2.
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const bf16* __restrict__ dy, const unsigned char* __restrict__ arg,
                                                          int N, int H, int W, int C, int Ho, int Wo, bf16* __restrict__ dx) {
  const int G = C >> 3;
  const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1;
  const long long total = (long long)N * Hb * Wb * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % G);
    long long p = e / G;
    const int j = (int)(p % Wb); p /= Wb;
    const int i = (int)(p % Hb);
    const int n = (int)(p / Hb);
    uint4 dv[2][2];
    uint2 av[2][2];
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        dv[di][dj] = make_uint4(0, 0, 0, 0);
        av[di][dj] = make_uint2(0xffffffffu, 0xffffffffu);          // no tap has code 255
        if (i + di < Ho && j + dj < Wo) {
          const size_t o = (((size_t)n * Ho + i + di) * Wo + j + dj) * C + g * 8;
          dv[di][dj] = __ldg(reinterpret_cast<const uint4*>(dy + o));
          av[di][dj] = __ldg(reinterpret_cast<const uint2*>(arg + o));
        }
      }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int iy = 2 * i + a, ix = 2 * j + b;
        if (iy >= H || ix >= W) continue;
        V8 acc;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
#pragma unroll
        for (int di = 0; di < 2; ++di)
#pragma unroll
          for (int dj = 0; dj < 2; ++dj) {
            const int r = 1 + a - 2 * di, s2 = 1 + b - 2 * dj;
            if (r < 0 || r > 2 || s2 < 0 || s2 > 2) continue;       // resolved at compile time
            const unsigned int code = (unsigned int)(r * 3 + s2);
            const V8 d = unpack8(dv[di][dj]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const unsigned int am = ((q < 4 ? av[di][dj].x : av[di][dj].y) >> ((q & 3) * 8)) & 0xffu;
              if (am == code) acc.v[q] += d.v[q];
            }
          }
        st8(dx + (((size_t)n * H + iy) * W + ix) * C + g * 8, acc);
      }
  }
}

// ---- bilinear resize, align_corners = True (nn.UpsamplingBilinear2d) -----------------------------------------
__device__ __forceinline__ void src_index(int o, float scale, int in, int& i0, int& i1, float& lam) {
  const float s = scale * (float)o;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = min(i0 + 1, in - 1);
  lam = s - (float)i0;
}
// forward: a thread produces 4 consecutive output pixels of one row x 8 channels; they share the two source rows
// and (for scale factors >= 2) at most 3 source columns, cached in registers between consecutive outputs
__global__ void __launch_bounds__(256) upsample_fwd_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int ldx,
                                                           bf16* __restrict__ y, int Ho, int Wo, int ldy) {
  const int G = C >> 3;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const int Wq = (Wo + 3) >> 2;
  const long long total = (long long)N * Ho * Wq * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % G);
    long long p = e / G;
    const int oq = (int)(p % Wq); p /= Wq;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    int y0, y1; float ly;
    src_index(oy, sy, H, y0, y1, ly);
    const bf16* r0 = x + ((size_t)n * H + y0) * W * ldx + g * 8;
    const bf16* r1 = x + ((size_t)n * H + y1) * W * ldx + g * 8;
    auto column = [&](int xx) {      // source column xx blended between the two source rows
      const V8 a = unpack8(ldg16(r0 + (size_t)xx * ldx)), b = unpack8(ldg16(r1 + (size_t)xx * ldx));
      V8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = (1.f - ly) * a.v[i] + ly * b.v[i];
      return o;
    };
    int hx0 = -1, hx1 = -1;
    V8 h0, h1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ox = oq * 4 + q;
      if (ox >= Wo) break;
      int x0, x1; float lx;
      src_index(ox, sx, W, x0, x1, lx);
      V8 a, b;
      if (x0 == hx0) a = h0; else if (x0 == hx1) a = h1; else a = column(x0);
      if (x1 == x0) b = a; else if (x1 == hx1) b = h1; else if (x1 == hx0) b = h0; else b = column(x1);
      h0 = a; hx0 = x0; h1 = b; hx1 = x1;
      V8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = (1.f - lx) * a.v[i] + lx * b.v[i];
      st8(y + (((size_t)n * Ho + oy) * Wo + ox) * ldy + g * 8, o);
    }
  }
}
// gather form of the transpose: input pixel collects from every output pixel that samples it.  The column weights
// of the (at most UPB_MAX) candidate output columns are computed once per thread, not once per candidate row.
constexpr int UPB_MAX = 12;
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const bf16* __restrict__ dy, int N, int Ho, int Wo, int C, int lddy,
                                                           bf16* __restrict__ dx, int H, int W, int lddx) {
  const int G = C >> 3;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)N * H * W * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % G);
    long long p = e / G;
    const int ix = (int)(p % W); p /= W;
    const int iy = (int)(p % H);
    const int n = (int)(p / H);
    // candidate outputs: scale*o in (i-1, i+1)
    const int oy_lo = sy > 0.f ? max(0, (int)floorf((float)(iy - 1) / sy)) : 0;
    const int oy_hi = sy > 0.f ? min(Ho - 1, (int)ceilf((float)(iy + 1) / sy)) : Ho - 1;
    const int ox_lo = sx > 0.f ? max(0, (int)floorf((float)(ix - 1) / sx)) : 0;
    const int ox_hi = sx > 0.f ? min(Wo - 1, (int)ceilf((float)(ix + 1) / sx)) : Wo - 1;
    V8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
    auto weight = [](int o, float scale, int in, int i) {
      int a0, a1; float l;
      src_index(o, scale, in, a0, a1, l);
      float wgt = 0.f;
      if (a0 == i) wgt += 1.f - l;
      if (a1 == i) wgt += l;
      return wgt;
    };
    if (ox_hi - ox_lo < UPB_MAX) {
      float wx[UPB_MAX];
#pragma unroll
      for (int j = 0; j < UPB_MAX; ++j) wx[j] = ox_lo + j <= ox_hi ? weight(ox_lo + j, sx, W, ix) : 0.f;
      for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        const float wy = weight(oy, sy, H, iy);
        if (wy == 0.f) continue;
        const bf16* row = dy + (((size_t)n * Ho + oy) * Wo + ox_lo) * lddy + g * 8;
#pragma unroll
        for (int j = 0; j < UPB_MAX; ++j) {
          if (wx[j] == 0.f) continue;
          const V8 d = unpack8(ldg16(row + (size_t)j * lddy));
          const float wgt = wy * wx[j];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wgt, d.v[i], acc.v[i]);
        }
      }
    } else {
      for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        const float wy = weight(oy, sy, H, iy);
        if (wy == 0.f) continue;
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
          const float wxx = weight(ox, sx, W, ix);
          if (wxx == 0.f) continue;
          const V8 d = ld8(dy + (((size_t)n * Ho + oy) * Wo + ox) * lddy + g * 8);
          const float wgt = wy * wxx;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wgt, d.v[i], acc.v[i]);
        }
      }
    }
    st8(dx + (((size_t)n * H + iy) * W + ix) * lddx + g * 8, acc);
  }
}

// ---- nearest x2 up-sampling (smp Unet DecoderBlock: F.interpolate(scale_factor=2, mode="nearest")) ---------
__global__ void nearest2x_fwd_kernel(const bf16* x, int N, int H, int W, int C, int ldx, bf16* y, int ldy) {
  const int G = C >> 3;
  const int Ho = 2 * H, Wo = 2 * W;
  const int oy = blockIdx.y, n = blockIdx.z;
  const bf16* xrow = x + ((size_t)n * H + (oy >> 1)) * W * ldx;
  bf16* yrow = y + ((size_t)n * Ho + oy) * Wo * ldy;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < Wo * G; e += gridDim.x * blockDim.x) {
    const int ox = e / G, g = e - ox * G;
    *reinterpret_cast<uint4*>(yrow + (size_t)ox * ldy + g * 8) =
        *reinterpret_cast<const uint4*>(xrow + (size_t)(ox >> 1) * ldx + g * 8);
  }
}
// dx[iy,ix] = sum of the 2x2 block of dy it was copied to
__global__ void nearest2x_bwd_kernel(const bf16* dy, int N, int H, int W, int C, int lddy, bf16* dx, int lddx) {
  const int G = C >> 3;
  const int Wo = 2 * W;
  const int iy = blockIdx.y, n = blockIdx.z;
  const bf16* d0 = dy + ((size_t)n * 2 * H + 2 * iy) * Wo * lddy;
  const bf16* d1 = d0 + (size_t)Wo * lddy;
  bf16* xrow = dx + ((size_t)n * H + iy) * W * lddx;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < W * G; e += gridDim.x * blockDim.x) {
    const int ix = e / G, g = e - ix * G;
    const V8 a = ld8(d0 + (size_t)(2 * ix) * lddy + g * 8), b = ld8(d0 + (size_t)(2 * ix + 1) * lddy + g * 8);
    const V8 c = ld8(d1 + (size_t)(2 * ix) * lddy + g * 8), d = ld8(d1 + (size_t)(2 * ix + 1) * lddy + g * 8);
    V8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = (a.v[i] + b.v[i]) + (c.v[i] + d.v[i]);
    st8(xrow + (size_t)ix * lddx + g * 8, o);
  }
}
// channel-slice copy (skip connection into a concat buffer)
__global__ void copy_kernel(const bf16* x, int ldx, bf16* y, int ldy, long long P, int C) {
  const int G = C >> 3;
  AADG_FOR_PIXEL_GROUPS(P, G, p, g) {
    *reinterpret_cast<uint4*>(y + p * ldy + g * 8) = *reinterpret_cast<const uint4*>(x + p * ldx + g * 8);
  }
}

// ---- global average pool / broadcast --------------------------------------------------------------------
// out[n][c] (fp32) = mean over pixels ; grid (N, ceil(G/32)), block (32 groups, 8 pixel lanes)
__global__ void gap_kernel(const bf16* x, int HW, int C, int ld, float* out, float scale) {
  const int n = blockIdx.x, g = blockIdx.y * 32 + threadIdx.x, G = C >> 3;
  __shared__ float sm[8][32][8];
  float a[8] = {};
  if (g < G)
    for (int p = threadIdx.y; p < HW; p += 8) {
      const V8 v = ld8(x + ((size_t)n * HW + p) * ld + g * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += v.v[i];
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[threadIdx.y][threadIdx.x][i] = a[i];
  __syncthreads();
  if (threadIdx.y == 0 && g < G) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x][i];
      out[(size_t)n * C + g * 8 + i] = t * scale;
    }
  }
}
// y[n, p, :] = v[n, :]  (bilinear resize of a 1x1 map); v bf16 [N, C]
__global__ void broadcast_kernel(const bf16* v, int C, bf16* y, int HW, int ldy, long long total_pix) {
  const int G = C >> 3;
  const long long total = total_pix * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / G;
    const int g = (int)(e - p * G);
    const long long n = p / HW;
    *reinterpret_cast<uint4*>(y + p * ldy + g * 8) = *reinterpret_cast<const uint4*>(v + n * C + g * 8);
  }
}
// y[n, p, :] += scale * v[n, :]  (gradient of a global average pool joining a feature map's gradient); v fp32 [N, C]
__global__ void broadcast_add_kernel(const float* v, int C, bf16* y, int HW, int ldy, long long total_pix, float scale) {
  const int G = C >> 3;
  const long long total = total_pix * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / G;
    const int g = (int)(e - p * G);
    const long long n = p / HW;
    V8 a = ld8(y + p * ldy + g * 8);
    const V8 b = ld8f(v + n * C + g * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) a.v[i] = fmaf(scale, b.v[i], a.v[i]);
    st8(y + p * ldy + g * 8, a);
  }
}
// fp32 [N, C] -> bf16 [N, C] (optionally scaled)
__global__ void f32_to_bf16_kernel(const float* x, bf16* y, long long n, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i] * scale);
}

// ---- depthwise 3x3 (dilated, stride 1, "same" padding) ---------------------------------------------------
// y[n,oy,ox,c] = sum_t w[t][c] * x[n, oy + (r-1)*dil*sign, ox + (s-1)*dil*sign, c]; sign = -1 gives the data gradient
// generic (dilated) form: one output pixel x 8 channels per thread (filter taps come from L1)
__global__ void dw3x3_kernel(const bf16* x, int N, int H, int W, int C, int ldx, const float* s_w, int dil, int sign,
                             bf16* y, int ldy, int accumulate) {
  const int G = C >> 3;
  const int oy = blockIdx.y, n = blockIdx.z;
  const bf16* xn = x + (size_t)n * H * W * ldx;
  bf16* yrow = y + ((size_t)n * H + oy) * W * ldy;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < W * G; e += gridDim.x * blockDim.x) {
    const int ox = e / G, g = e - ox * G;
    V8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy + (r - 1) * dil * sign;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ix = ox + (s2 - 1) * dil * sign;
        if (ix < 0 || ix >= W) continue;
        const V8 v = ld8(xn + ((size_t)iy * W + ix) * ldx + g * 8);
        const V8 wv = ld8f(s_w + (r * 3 + s2) * C + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wv.v[i], v.v[i], acc.v[i]);
      }
    }
    if (accumulate) {
      const V8 prev = ld8(yrow + (size_t)ox * ldy + g * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] += prev.v[i];
    }
    st8(yrow + (size_t)ox * ldy + g * 8, acc);
  }
}
// ---- strided depthwise 3x3 (MobileNetV2 down-sampling blocks): padding = dilation, output (ho, wo) ---------------
// forward: y[n,oy,ox,c] = sum_t w[t][c] * x[n, oy*stride + (r-1)*dil, ox*stride + (s-1)*dil, c]
__global__ void __launch_bounds__(256) dw3x3_strided_fwd_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int ldx,
                                                                const float* __restrict__ wgt, int dil, int stride,
                                                                bf16* __restrict__ y, int Ho, int Wo, int ldy) {
  const int G = C >> 3;
  const long long total = (long long)N * Ho * Wo * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % G);
    long long p = e / G;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    V8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * stride + (r - 1) * dil;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ix = ox * stride + (s2 - 1) * dil;
        if (ix < 0 || ix >= W) continue;
        const V8 v = unpack8(ldg16(x + (((size_t)n * H + iy) * W + ix) * ldx + g * 8));
        const V8 wv = ld8f(wgt + (r * 3 + s2) * C + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wv.v[i], v.v[i], acc.v[i]);
      }
    }
    st8(y + (((size_t)n * Ho + oy) * Wo + ox) * ldy + g * 8, acc);
  }
}
// data gradient (gather form): dx[n,iy,ix,c] = sum over taps with (iy - (r-1)*dil) divisible by stride of w[t][c] * dy[...]
__global__ void __launch_bounds__(256) dw3x3_strided_dgrad_kernel(const bf16* __restrict__ dy, int N, int Ho, int Wo, int C,
                                                                  int lddy, const float* __restrict__ wgt, int dil, int stride,
                                                                  bf16* __restrict__ dx, int H, int W, int lddx) {
  const int G = C >> 3;
  const long long total = (long long)N * H * W * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % G);
    long long p = e / G;
    const int ix = (int)(p % W); p /= W;
    const int iy = (int)(p % H);
    const int n = (int)(p / H);
    V8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ty = iy - (r - 1) * dil;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int tx = ix - (s2 - 1) * dil;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        const V8 v = unpack8(ldg16(dy + (((size_t)n * Ho + oy) * Wo + ox) * lddy + g * 8));
        const V8 wv = ld8f(wgt + (r * 3 + s2) * C + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wv.v[i], v.v[i], acc.v[i]);
      }
    }
    st8(dx + (((size_t)n * H + iy) * W + ix) * lddx + g * 8, acc);
  }
}
// weight gradient: dw[t][c] += sum over output pixels of dy[o][c] * x[o*stride + off_t][c]; block (TX groups, TY lanes)
__global__ void __launch_bounds__(256) dw3x3_strided_wgrad_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int ldx,
                                                                  const bf16* __restrict__ dy, int Ho, int Wo, int lddy,
                                                                  int dil, int stride, float* dw) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r = blockIdx.y;                      // filter row
  const long long P = (long long)N * Ho * Wo;
  float acc[3][8] = {};
  if (tx < (C >> 3))
    for (long long p = (long long)blockIdx.x * blockDim.y + ty; p < P; p += (long long)gridDim.x * blockDim.y) {
      const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
      const int iy = oy * stride + (r - 1) * dil;
      if (iy < 0 || iy >= H) continue;
      const V8 d = unpack8(ldg16(dy + (size_t)p * lddy + tx * 8));
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ix = ox * stride + (s2 - 1) * dil;
        if (ix < 0 || ix >= W) continue;
        const V8 v = unpack8(ldg16(x + (((size_t)n * H + iy) * W + ix) * ldx + tx * 8));
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[s2][i] = fmaf(d.v[i], v.v[i], acc[s2][i]);
      }
    }
  // fold the three taps of this filter row: reuse the two-array block reduction twice
  float z[8] = {};
  block_channel_sum<2>(C, acc[0], acc[1], dw + (size_t)(r * 3 + 0) * C, dw + (size_t)(r * 3 + 1) * C);
  __syncthreads();
  block_channel_sum<1>(C, acc[2], z, dw + (size_t)(r * 3 + 2) * C, nullptr);
}

// ---- stride 2, dilation 1 (every down-sampling depthwise layer of MobileNetV2): channel-stationary threads --------
// The generic strided kernels above decode (n, y, x, group) from a 64-bit flat index per element and re-read the taps
// from L1 per tap: measured instruction-bound (MobileNetV2 @256^2: forward 120 us, data gradient 300 us, weight
// gradient 250 us per launch against 30-90 us of HBM time).  Here a thread owns ONE 8-channel group for its whole life
// (its 9 x 8 taps live in registers), blockIdx.y walks the images and the pixel index inside an image is 32-bit: one
// 32-bit division per pixel.  block = 256 threads = (256 / G) pixel lanes x G groups (G = C / 8 <= 256).
//   forward        one output pixel per trip, 9 loads
//   data gradient  one 2x2 input quad (rows 2a, 2a+1; columns 2b, 2b+1) per trip: with stride 2 the quad is covered by
//                  exactly the four outputs (a..a+1, b..b+1), each loaded once, and every tap is used exactly once
//                  (same accumulation order as the generic gather: filter row, then filter column)
//   weight grad.   9 x 8 register accumulators over the thread's pixels, shared-memory then global reduction per CTA
__device__ __forceinline__ void s2_load_taps(const float* __restrict__ wgt, int C, int g, float (&wt)[9][8]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const V8 w8 = ld8f(wgt + (size_t)t * C + g * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) wt[t][i] = w8.v[i];
  }
}
__global__ void __launch_bounds__(256, 2) dw3x3_s2_fwd_kernel(const bf16* __restrict__ x, int H, int W, int C, int ldx,
                                                           const float* __restrict__ wgt, bf16* __restrict__ y, int Ho, int Wo,
                                                           int ldy) {
  const int G = C >> 3, lanes = 256 / G;
  const int lane = threadIdx.x / G, g = threadIdx.x - lane * G;
  if (lane >= lanes) return;
  const int n = blockIdx.y, P = Ho * Wo;
  float wt[9][8];
  s2_load_taps(wgt, C, g, wt);
  const bf16* xn = x + (size_t)n * H * W * ldx + g * 8;
  bf16* yn = y + (size_t)n * P * ldy + g * 8;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int q = blockIdx.x * lanes + lane; q < P; q += gridDim.x * lanes) {
    const int oy = q / Wo, ox = q - oy * Wo;
    const int iy0 = 2 * oy - 1, ix0 = 2 * ox - 1;
    uint4 raw[9];                                          // all nine loads in flight before the first use
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int iy = iy0 + r, ix = ix0 + s2;
        const bool ok = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
        raw[r * 3 + s2] = ok ? ldg16(xn + (size_t)(iy * W + ix) * ldx) : zero;
      }
    V8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const V8 v = unpack8(raw[t]);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wt[t][i], v.v[i], acc.v[i]);
    }
    st8(yn + (size_t)q * ldy, acc);
  }
}
__global__ void __launch_bounds__(256, 2) dw3x3_s2_dgrad_kernel(const bf16* __restrict__ dy, int Ho, int Wo, int C, int lddy,
                                                             const float* __restrict__ wgt, bf16* __restrict__ dx, int H, int W,
                                                             int lddx) {
  const int G = C >> 3, lanes = 256 / G;
  const int lane = threadIdx.x / G, g = threadIdx.x - lane * G;
  if (lane >= lanes) return;
  const int n = blockIdx.y;
  const int Hq = (H + 1) >> 1, Wq = (W + 1) >> 1;          // quads; Hq == Ho and Wq == Wo for a 3x3 / stride-2 / pad-1 layer
  float wt[9][8];
  s2_load_taps(wgt, C, g, wt);
  const bf16* dyn = dy + (size_t)n * Ho * Wo * lddy + g * 8;
  bf16* dxn = dx + (size_t)n * H * W * lddx + g * 8;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int q = blockIdx.x * lanes + lane; q < Hq * Wq; q += gridDim.x * lanes) {
    const int a = q / Wq, b = q - a * Wq;
    const bool a1 = a + 1 < Ho, b1 = b + 1 < Wo, in = a < Ho && b < Wo;
    const bf16* d = dyn + (size_t)(a * Wo + b) * lddy;
    const V8 d00 = unpack8(in ? ldg16(d) : zero);
    const V8 d01 = unpack8(in && b1 ? ldg16(d + lddy) : zero);
    const V8 d10 = unpack8(in && a1 ? ldg16(d + (size_t)Wo * lddy) : zero);
    const V8 d11 = unpack8(in && a1 && b1 ? ldg16(d + (size_t)(Wo + 1) * lddy) : zero);
    const int iy = 2 * a, ix = 2 * b;
    bf16* o = dxn + (size_t)(iy * W + ix) * lddx;
    V8 acc;
    // (even row, even column): tap (1,1) of output (a, b)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = wt[4][i] * d00.v[i];
    st8(o, acc);
    if (ix + 1 < W) {   // (even, odd): tap (1,0) of (a, b+1), tap (1,2) of (a, b)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wt[5][i], d00.v[i], wt[3][i] * d01.v[i]);
      st8(o + lddx, acc);
    }
    if (iy + 1 < H) {
      // (odd, even): tap (0,1) of (a+1, b), tap (2,1) of (a, b)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wt[7][i], d00.v[i], wt[1][i] * d10.v[i]);
      st8(o + (size_t)W * lddx, acc);
      if (ix + 1 < W) {   // (odd, odd): taps (0,0) of (a+1, b+1), (0,2) of (a+1, b), (2,0) of (a, b+1), (2,2) of (a, b)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          acc.v[i] = fmaf(wt[8][i], d00.v[i], fmaf(wt[6][i], d01.v[i], fmaf(wt[2][i], d10.v[i], wt[0][i] * d11.v[i])));
        st8(o + (size_t)(W + 1) * lddx, acc);
      }
    }
  }
}
// work item = (image, block of `lanes * S2_WG_PIX` output pixels); the grid is ONE wave of CTAs striding over the items
constexpr int S2_WG_PIX = 8;
__global__ void __launch_bounds__(256, 2) dw3x3_s2_wgrad_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int ldx,
                                                             const bf16* __restrict__ dy, int Ho, int Wo, int lddy, float* dw,
                                                             int blocks_per_image) {
  extern __shared__ float s_dw9[];                        // [9][C]
  const int G = C >> 3, lanes = 256 / G;
  const int lane = threadIdx.x / G, g = threadIdx.x - lane * G;
  const int P = Ho * Wo;
  for (int i = threadIdx.x; i < 9 * C; i += 256) s_dw9[i] = 0.f;
  __syncthreads();
  if (lane < lanes) {
    float acc[9][8] = {};
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const int items = N * blocks_per_image;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int n = item / blocks_per_image, blk = item - n * blocks_per_image;
      const bf16* xn = x + (size_t)n * H * W * ldx + g * 8;
      const bf16* dyn = dy + (size_t)n * P * lddy + g * 8;
      const int q0 = blk * lanes * S2_WG_PIX + lane;
#pragma unroll 1
      for (int j = 0; j < S2_WG_PIX; ++j) {
        const int q = q0 + j * lanes;
        if (q >= P) break;
        const int oy = q / Wo, ox = q - oy * Wo;
        const int iy0 = 2 * oy - 1, ix0 = 2 * ox - 1;
        const uint4 draw = ldg16(dyn + (size_t)q * lddy);
        uint4 raw[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const int iy = iy0 + r, ix = ix0 + s2;
            const bool ok = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
            raw[r * 3 + s2] = ok ? ldg16(xn + (size_t)(iy * W + ix) * ldx) : zero;
          }
        const V8 d = unpack8(draw);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const V8 v = unpack8(raw[t]);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[t][i] = fmaf(d.v[i], v.v[i], acc[t][i]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_dw9[t * C + g * 8 + i], acc[t][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * C; i += 256) atomicAdd(&dw[i], s_dw9[i]);
}

// dw[t][c] += sum_px dy[px][c] * x[px + off_t][c]; blockIdx.y = filter row r (3 taps, 24 register accumulators),
// the three launches' dy reads overlap in L2; block reduction in shared memory [3][C], then atomics
__global__ void __launch_bounds__(256, 3) dw3x3_wgrad_kernel(const bf16* x, int N, int H, int W, int C, int ldx,
                                                             const bf16* dy, int lddy, int dil, float* dw) {
  extern __shared__ float s_dw[];
  const int G = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r = blockIdx.y;
  const long long P = (long long)N * H * W;
  float acc[3][8] = {};
  if (tx < G)
    for (long long p = (long long)blockIdx.x * blockDim.y + ty; p < P; p += (long long)gridDim.x * blockDim.y) {
      const int ox = (int)(p % W), oy = (int)((p / W) % H);
      const int iy = oy + (r - 1) * dil;
      if (iy < 0 || iy >= H) continue;
      const V8 d = ld8(dy + (size_t)p * lddy + tx * 8);
      const bf16* xr = x + (size_t)(p + (long long)(iy - oy) * W) * ldx + tx * 8;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ix = ox + (s2 - 1) * dil;
        if (ix < 0 || ix >= W) continue;
        const V8 v = ld8(xr + (long long)(ix - ox) * ldx);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[s2][i] = fmaf(d.v[i], v.v[i], acc[s2][i]);
      }
    }
  for (int i = ty * blockDim.x + tx; i < 3 * C; i += blockDim.x * blockDim.y) s_dw[i] = 0.f;
  __syncthreads();
  if (tx < G) {
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_dw[t * C + tx * 8 + i], acc[t][i]);
  }
  __syncthreads();
  for (int i = ty * blockDim.x + tx; i < 3 * C; i += blockDim.x * blockDim.y) atomicAdd(&dw[(size_t)r * 3 * C + i], s_dw[i]);
}

// ---- stem im2col: fp32 NCHW [-1,1] image -> bf16 [N*Ho*Wo][KP] patches, k = (r*S + s)*3 + c --------------------
// one CTA per (image, output row): the R input rows it needs are staged as bf16 in shared memory
// [R][3][W + 2*pad] (zero padded), then every thread emits 16-byte groups of 8 patch values
__global__ void im2col_stem_kernel(const float* img, int N, int H, int W, int R, int S, int stride, int pad, int Ho,
                                   int Wo, int KP, bf16* col) {
  extern __shared__ bf16 s_rows[];
  const int oy = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int WP = W + 2 * pad;
  for (int e = threadIdx.x; e < R * 3 * WP; e += blockDim.x) {
    const int xp = e % WP, c = (e / WP) % 3, r = e / (3 * WP);
    const int iy = oy * stride - pad + r, ix = xp - pad;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = img[(((size_t)n * 3 + c) * H + iy) * W + ix];
    s_rows[e] = __float2bfloat16_rn(v);
  }
  // patch index k -> offset inside the staged rows (computed once per CTA: no divisions in the copy loop)
  __shared__ short koff[512];
  const int K = R * S * 3, KG = KP / 8;
  for (int k = threadIdx.x; k < KP; k += blockDim.x) {
    int off = -1;
    if (k < K) { const int c = k % 3, rs = k / 3, s2 = rs % S, r = rs / S; off = (r * 3 + c) * WP + s2; }
    koff[k] = (short)off;
  }
  __syncthreads();
  bf16* orow = col + ((size_t)n * Ho + oy) * Wo * KP;
  const bf16 zero = __float2bfloat16_rn(0.f);
  for (int e = threadIdx.x; e < Wo * KG; e += blockDim.x) {
    const int ox = e / KG, kg = e - ox * KG;
    const int xs = ox * stride;
    __align__(16) bf16 o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int off = koff[kg * 8 + i];
      o[i] = off >= 0 ? s_rows[off + xs] : zero;
    }
    *reinterpret_cast<uint4*>(orow + (size_t)ox * KP + kg * 8) = *reinterpret_cast<const uint4*>(o);
  }
}

// ---- stem input in space-to-depth form ------------------------------------------------------------------------------
// A 7x7 / stride-2 / pad-3 convolution over 3 channels is a 4x4 / stride-1 convolution over the 2x2 space-to-depth image
// (12 channels, padded to 16): s2d pixel (Y, X) holds input pixels (2Y+py, 2X+px), channel (py*2 + px)*3 + c.  Written
// here as bf16 [N, H/2 + 3, W/2 + 3, 16] with TWO zero rows / columns before and ONE after the image, so that output
// pixel (oy, ox) needs the s2d rows oy..oy+3 and, in each, the FOUR CONSECUTIVE s2d pixels ox..ox+3 = 64 contiguous
// bf16 = one 128-byte row of a tcgen05 A tile.  The convolution kernel reads those windows straight out of this buffer
// with an overlapping tensor map (pixel pitch 16 elements, 64 "channels": aadg_conv_fprop_windows_bf16) -- the
// 3.17 GB im2col patch buffer of round 1 (written once, read twice per step) is gone; this buffer is 0.3 GB.
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* __restrict__ img, int N, int H, int W,
                                                       bf16* __restrict__ out) {
  const int Hs = H / 2 + 3, Ws = W / 2 + 3;
  const long long total = (long long)N * Hs * Ws;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int xs = (int)(e % Ws);
    const long long q = e / Ws;
    const int ys = (int)(q % Hs), n = (int)(q / Hs);
    const int Y = ys - 2, X = xs - 2;
    uint32_t wv[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (Y >= 0 && Y < H / 2 && X >= 0 && X < W / 2) {
      float v[12];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int py = 0; py < 2; ++py) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(img + (((size_t)n * 3 + c) * H + 2 * Y + py) * W + 2 * X));
          v[(py * 2 + 0) * 3 + c] = t.x;
          v[(py * 2 + 1) * 3 + c] = t.y;
        }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        wv[k] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + e * 16);
    dst[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    dst[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
  }
}


// Row-pitched patch layout k = r*RP + s*3 + c (RP a multiple of 8, >= 3*S): the 3*S values of one filter row are
// contiguous in an [x][c]-interleaved staged image row, so a (pixel, filter row) unit is a copy of RP/2 32-bit
// words from shared memory (tail masked to zero) into RP/8 aligned 16-byte stores -- no per-element index table.
// One CTA per (image, output row); stride must be even (word-aligned source runs).
__global__ void __launch_bounds__(256) im2col_rows_kernel(const float* __restrict__ img, int N, int H, int W, int R, int S,
                                                          int stride, int pad, int Ho, int Wo, int RP, int KP,
                                                          bf16* __restrict__ col) {
  extern __shared__ bf16 s_rows[];                   // [R][WP*3 (+1 to keep rows word aligned)]
  const int oy = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int WP = W + 2 * pad;
  const int pitch = (WP * 3 + RP + 1) & ~1;          // elements; the slack keeps the last pixel's masked tail in bounds
  for (int e = threadIdx.x; e < R * pitch; e += blockDim.x) s_rows[e] = __float2bfloat16_rn(0.f);
  __syncthreads();
  for (int rc = 0; rc < R * 3; ++rc) {
    const int r = rc / 3, c = rc - r * 3;
    const int iy = oy * stride - pad + r;
    if (iy < 0 || iy >= H) continue;
    const float* src = img + (((size_t)n * 3 + c) * H + iy) * W;
    bf16* dst = s_rows + r * pitch + pad * 3 + c;
    for (int x = threadIdx.x; x < W; x += blockDim.x) dst[x * 3] = __float2bfloat16_rn(__ldg(src + x));
  }
  __syncthreads();
  const int groups = RP >> 3, valid = S * 3;
  bf16* orow = col + ((size_t)n * Ho + oy) * Wo * KP;
  for (int u = threadIdx.x; u < Wo * R * groups; u += blockDim.x) {
    const int q = u % groups, t = u / groups;
    const int r = t % R, ox = t / R;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(s_rows + r * pitch + ox * stride * 3) + q * 4;
    uint32_t wv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = q * 8 + 2 * j;
      uint32_t v = src[j];
      if (e + 1 >= valid) v = e < valid ? (v & 0xffffu) : 0u;
      wv[j] = v;
    }
    *reinterpret_cast<uint4*>(orow + (size_t)ox * KP + r * RP + q * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
  }
  const int tail = (KP - R * RP) >> 3;
  for (int u = threadIdx.x; u < Wo * tail; u += blockDim.x) {
    const int q = u % tail, ox = u / tail;
    *reinterpret_cast<uint4*>(orow + (size_t)ox * KP + R * RP + q * 8) = make_uint4(0, 0, 0, 0);
  }
}

// ---- Adam (torch.optim.Adam defaults, scheduler.py:10-11) over a flat parameter buffer ------------------------
__global__ void adam_kernel(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                            float eps, float bc1, float bc2_sqrt, float wd) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}
// the same update with (lr, beta1, beta2, eps, weight_decay, grad_scale) and the 1-based step count read from device
// memory, so that a captured CUDA graph replays it unchanged while the host advances the counter / changes the rate;
// grad_scale (1 / world size after a summing all-reduce) is folded into the gradient load
__global__ void adam_dev_kernel(float* p, const float* g, float* m, float* v, long long n, const float* hyper,
                                const long long* step_ptr) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gs = hyper[5];
  const float step = (float)*step_ptr;
  const float bc1 = 1.f - powf(b1, step), bc2_sqrt = sqrtf(1.f - powf(b2, step));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}
// fp32 master weights [taps][Cout][Cin] -> bf16 copy and bf16 transposed copy [taps][Cin][Cout]
struct WeightDesc { long long off_master, off_bf16, off_bf16_t; int taps, cout, cin, pad_; };
__global__ void weight_prep_kernel(const float* master, bf16* wb, bf16* wbt, const WeightDesc* descs) {
  const WeightDesc d = descs[blockIdx.y];
  const long long n = (long long)d.taps * d.cout * d.cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = master[d.off_master + i];
    const bf16 b = __float2bfloat16_rn(v);
    wb[d.off_bf16 + i] = b;
    if (d.off_bf16_t >= 0) {
      const int ci = (int)(i % d.cin);
      const long long q = i / d.cin;
      const int co = (int)(q % d.cout);
      const long long t = q / d.cout;
      wbt[d.off_bf16_t + (t * d.cin + ci) * d.cout + co] = b;
    }
  }
}

static int num_sms_nn() {
  static int v = 0;
  if (!v) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}
static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  return (int)std::max<long long>(1, std::min<long long>(g, 148 * 16));
}
static inline dim3 reduce_block(int C) {
  int tx = 1;
  while (tx < (C >> 3)) tx <<= 1;
  tx = std::min(tx, 256);
  return dim3(tx, 256 / tx);
}
// grid of a thread-stationary streaming pass: `per_sm` resident blocks on every SM, never more blocks than there
// are BN_U-pixel trips to hand out
static inline int stream_blocks(long long pixels, dim3 blk, int per_sm) {
  const long long trips = (pixels + (long long)blk.y * BN_U - 1) / ((long long)blk.y * BN_U);
  return (int)std::max<long long>(1, std::min<long long>(trips, (long long)num_sms_nn() * per_sm));
}

// generic depthwise fallbacks (any dilation / image size); csrc/dwconv.cu holds the TMA-tiled fast paths and the C ABI
int dwconv3x3_generic(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int direction,
                      void* y, int ldy, cudaStream_t st) {
  dim3 grid(std::max(1, std::min((w * (c / 8) + 255) / 256, 64)), h, n);
  dw3x3_kernel<<<grid, 256, 0, st>>>((const bf16*)x, n, h, w, c, ldx, wgt, dil, (direction & 1) ? -1 : 1, (bf16*)y, ldy,
                                     (direction & 2) ? 1 : 0);
  return check_launch("dwconv3x3");
}
int dwconv3x3_wgrad_generic(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int lddy, int dil,
                            float* dw, cudaStream_t st) {
  const long long pixels = (long long)n * h * w;
  const dim3 blk = reduce_block(c);
  const int blocks = (int)std::min<long long>((pixels + blk.y * 16 - 1) / (blk.y * 16), 148 * 6);
  const size_t smem = (size_t)3 * c * sizeof(float);
  dim3 grid(std::max(blocks, 1), 3);
  dw3x3_wgrad_kernel<<<grid, blk, smem, st>>>((const bf16*)x, n, h, w, c, ldx, (const bf16*)dy, lddy, dil, dw);
  return check_launch("dwconv3x3 wgrad");
}

// the channel-stationary stride-2 kernels: 3x3 / stride 2 / dilation 1 / pad 1, up to 2048 channels, 32-bit in-image
// indices; AADG_DW_S2=0 falls back to the generic strided kernels
static bool s2_fast(int c, int dil, int stride, int n, int h, int w, int ho, int wo) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("AADG_DW_S2"); on = e ? atoi(e) : 1; }
  return on && stride == 2 && dil == 1 && c >= 8 && c / 8 <= 256 && n <= 65535 && (long long)h * w < (1ll << 30) &&
         ho == (h - 1) / 2 + 1 && wo == (w - 1) / 2 + 1;
}
int dwconv3x3_strided(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int stride, int direction,
                      void* y, int ho, int wo, int ldy, cudaStream_t st) {
  if (s2_fast(c, dil, stride, n, h, w, ho, wo)) {
    const int lanes = 256 / (c / 8);
    if (direction == 0) {
      dim3 grid(std::max(1, std::min((ho * wo + lanes * 8 - 1) / (lanes * 8), 4096)), n);
      dw3x3_s2_fwd_kernel<<<grid, 256, 0, st>>>((const bf16*)x, h, w, c, ldx, wgt, (bf16*)y, ho, wo, ldy);
    } else {      // x = dy [n,ho,wo], y = dx [n,h,w]; one 2x2 input quad per thread and trip
      const int quads = ((h + 1) / 2) * ((w + 1) / 2);
      dim3 grid(std::max(1, std::min((quads + lanes * 4 - 1) / (lanes * 4), 4096)), n);
      dw3x3_s2_dgrad_kernel<<<grid, 256, 0, st>>>((const bf16*)x, ho, wo, c, ldx, wgt, (bf16*)y, h, w, ldy);
    }
    return check_launch("dwconv3x3 stride 2");
  }
  if (direction == 0)
    dw3x3_strided_fwd_kernel<<<grid_for((long long)n * ho * wo * (c / 8)), 256, 0, st>>>(
        (const bf16*)x, n, h, w, c, ldx, wgt, dil, stride, (bf16*)y, ho, wo, ldy);
  else      // x = dy [n,ho,wo], y = dx [n,h,w]
    dw3x3_strided_dgrad_kernel<<<grid_for((long long)n * h * w * (c / 8)), 256, 0, st>>>(
        (const bf16*)x, n, ho, wo, c, ldx, wgt, dil, stride, (bf16*)y, h, w, ldy);
  return check_launch("dwconv3x3 strided");
}
int dwconv3x3_strided_wgrad(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int ho, int wo, int lddy,
                            int dil, int stride, float* dw, cudaStream_t st) {
  if (s2_fast(c, dil, stride, n, h, w, ho, wo) && (size_t)9 * c * sizeof(float) <= 48 * 1024) {
    const int lanes = 256 / (c / 8);
    const int bpi = (ho * wo + lanes * S2_WG_PIX - 1) / (lanes * S2_WG_PIX);
    const int grid = (int)std::max<long long>(1, std::min<long long>((long long)n * bpi, (long long)num_sms_nn() * 2));   // 2 CTAs per SM: one wave
    dw3x3_s2_wgrad_kernel<<<grid, 256, (size_t)9 * c * sizeof(float), st>>>((const bf16*)x, n, h, w, c, ldx, (const bf16*)dy,
                                                                         ho, wo, lddy, dw, bpi);
    return check_launch("dwconv3x3 stride 2 wgrad");
  }
  const long long pixels = (long long)n * ho * wo;
  const dim3 blk = reduce_block(c);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((pixels + blk.y * 16 - 1) / (blk.y * 16), 148 * 4));
  dim3 grid(blocks, 3);
  dw3x3_strided_wgrad_kernel<<<grid, blk, 0, st>>>((const bf16*)x, n, h, w, c, ldx, (const bf16*)dy, ho, wo, lddy, dil, stride,
                                                  dw);
  return check_launch("dwconv3x3 strided wgrad");
}

}  // namespace nn
}  // namespace aadg

using namespace aadg;
using namespace aadg::nn;

#define NN_REQ_C(C) AADG_REQUIRE((C) > 0 && (C) % 8 == 0 && (C) <= 2048, "channels %d must be a multiple of 8 and <= 2048", (C))

extern "C" {

int aadg_bn_stats(const void* x, long long pixels, int c, int ld, float* sum, float* sumsq, void* stream) {
  NN_REQ_C(c);
  AADG_REQUIRE(pixels > 0 && pixels < (1ll << 31), "bad pixel count");
  const dim3 blk = reduce_block(c);
  bn_stats_kernel<<<stream_blocks(pixels, blk, 4), blk, 0, (cudaStream_t)stream>>>((const bf16*)x, pixels, c, ld, sum, sumsq);
  return check_launch("bn_stats");
}

int aadg_bn_finalize(float* sum, float* sumsq, const float* gamma, const float* beta, int c, float count,
                     float eps, float momentum, float* mean, float* invstd, float* scale, float* shift,
                     float* run_mean, float* run_var, int reset_sums, void* stream) {
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, gamma, beta, c, count, eps, momentum,
                                                                        mean, invstd, scale, shift, run_mean, run_var,
                                                                        reset_sums);
  return check_launch("bn_finalize");
}

int aadg_bn_apply(const void* x, int ldx, const float* scale, const float* shift, const void* res, int ldr, void* y,
                  int ldy, long long pixels, int c, int flags, unsigned long long seed, void* relu_bits,
                  void* stream) {
  NN_REQ_C(c);
  const dim3 blk = reduce_block(c);
  bn_apply_kernel<<<stream_blocks(pixels, blk, 3), blk, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, ldx, scale, shift, (const bf16*)res, ldr, (bf16*)y, ldy, pixels, c, flags, seed,
      (unsigned char*)relu_bits);
  return check_launch("bn_apply");
}

// phase bit 1 = reduce (dgamma / dbeta accumulate), bit 2 = apply; inv_count <= 0 -> 1 / pixels
static int bn_backward_impl(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const void* y,
                            int ldy, const float* mean, const float* invstd, const float* gamma, const float* shift,
                            long long pixels, int c, int flags, unsigned long long seed, float* dgamma, float* dbeta,
                            void* dx, int lddx, void* dres, int lddr, int dres_accumulate, void* stream, int phase = 3,
                            float inv_count = 0.f) {
  NN_REQ_C(c);
  AADG_REQUIRE(pixels > 0 && pixels < (1ll << 31), "bad pixel count");
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & 16) && (phase & 1)) {     // flag 16: dgamma / dbeta are accumulated into (the caller zeroed its gradient buffer)
    AADG_CUDA_TRY(cudaMemsetAsync(dgamma, 0, sizeof(float) * c, st));
    AADG_CUDA_TRY(cudaMemsetAsync(dbeta, 0, sizeof(float) * c, st));
  }
  if (!(inv_count > 0.f)) inv_count = 1.f / (float)pixels;
  const dim3 blk = reduce_block(c);
  AADG_REQUIRE(!(flags & 1) || (flags & 4) || y, "ReLU mask needs y / the bit mask (or flag 4 to recompute it from x)");
  AADG_REQUIRE(!(flags & 4) || shift, "flag 4 needs the forward shift vector");
  const int mask = (flags & 4) ? 4 : (flags & 8) ? 8 : (flags & 1) ? 1 : 0;
  const int grid = stream_blocks(pixels, blk, 2);
#define AADG_BN_BWD(MASK, DY2)                                                                                         \
  {                                                                                                                    \
    if (phase & 1)                                                                                                     \
      bn_bwd_reduce_kernel<MASK, DY2><<<grid, blk, 0, st>>>((const bf16*)dy, lddy, (const bf16*)dy2, lddy2,             \
                                                            (const bf16*)x, ldx, (const bf16*)y, ldy, mean, invstd,     \
                                                            gamma, shift, pixels, c, flags, seed, dgamma, dbeta);       \
    if (phase & 2)                                                                                                     \
      bn_bwd_apply_kernel<MASK, DY2><<<grid, blk, 0, st>>>((const bf16*)dy, lddy, (const bf16*)dy2, lddy2,              \
                                                           (const bf16*)x, ldx, (const bf16*)y, ldy, mean, invstd,      \
                                                           gamma, shift, dgamma, dbeta, pixels, c, flags, seed,         \
                                                           (bf16*)dx, lddx, (bf16*)dres, lddr, dres_accumulate,         \
                                                           inv_count);                                                  \
  }
  if (dy2) {
    if (mask == 4) AADG_BN_BWD(4, true) else if (mask == 8) AADG_BN_BWD(8, true) else if (mask == 1) AADG_BN_BWD(1, true) else AADG_BN_BWD(0, true)
  } else {
    if (mask == 4) AADG_BN_BWD(4, false) else if (mask == 8) AADG_BN_BWD(8, false) else if (mask == 1) AADG_BN_BWD(1, false) else AADG_BN_BWD(0, false)
  }
#undef AADG_BN_BWD
  return check_launch("bn_backward");
}

/* The two halves of the backward, for batch statistics shared across ranks (SyncBN): `reduce` ACCUMULATES this
 * rank's sum(g * xhat) and sum(g) into dgamma_sum / dbeta_sum (zero them first); the caller all-reduces them; `apply`
 * takes the global sums and inv_count = 1 / (global pixel count).  dy2 may be NULL. */
int aadg_bn_backward_reduce(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const void* y,
                            int ldy, const float* mean, const float* invstd, const float* gamma, const float* shift,
                            long long pixels, int c, int flags, unsigned long long seed, float* dgamma_sum,
                            float* dbeta_sum, void* stream) {
  return bn_backward_impl(dy, lddy, dy2, dy2 ? lddy2 : 0, x, ldx, y, ldy, mean, invstd, gamma, shift, pixels, c, flags | 16,
                          seed, dgamma_sum, dbeta_sum, nullptr, 0, nullptr, 0, 0, stream, 1);
}

int aadg_bn_backward_apply(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const void* y,
                           int ldy, const float* mean, const float* invstd, const float* gamma, const float* shift,
                           long long pixels, int c, int flags, unsigned long long seed, const float* dgamma_sum,
                           const float* dbeta_sum, float inv_count, void* dx, int lddx, void* dres, int lddr,
                           int dres_accumulate, void* stream) {
  AADG_REQUIRE(inv_count > 0.f, "inv_count = 1 / (pixels behind the statistics) must be positive");
  return bn_backward_impl(dy, lddy, dy2, dy2 ? lddy2 : 0, x, ldx, y, ldy, mean, invstd, gamma, shift, pixels, c, flags | 16,
                          seed, (float*)dgamma_sum, (float*)dbeta_sum, dx, lddx, dres, lddr, dres_accumulate, stream, 2,
                          inv_count);
}

int aadg_bn_backward(const void* dy, int lddy, const void* x, int ldx, const void* y, int ldy, const float* mean,
                     const float* invstd, const float* gamma, const float* shift, long long pixels, int c, int flags,
                     unsigned long long seed, float* dgamma, float* dbeta, void* dx, int lddx, void* dres, int lddr,
                     int dres_accumulate, void* stream) {
  return bn_backward_impl(dy, lddy, nullptr, 0, x, ldx, y, ldy, mean, invstd, gamma, shift, pixels, c, flags, seed, dgamma,
                          dbeta, dx, lddx, dres, lddr, dres_accumulate, stream);
}

/* the same backward with the incoming gradient given as the sum of two tensors (dy + dy2): the two gradient
 * branches that meet at a residual block's input are added on load */
int aadg_bn_backward2(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const void* y, int ldy,
                      const float* mean, const float* invstd, const float* gamma, const float* shift, long long pixels,
                      int c, int flags, unsigned long long seed, float* dgamma, float* dbeta, void* dx, int lddx,
                      void* dres, int lddr, int dres_accumulate, void* stream) {
  AADG_REQUIRE(dy2, "dy2 is required (use aadg_bn_backward for a single gradient tensor)");
  return bn_backward_impl(dy, lddy, dy2, lddy2, x, ldx, y, ldy, mean, invstd, gamma, shift, pixels, c, flags, seed, dgamma,
                          dbeta, dx, lddx, dres, lddr, dres_accumulate, stream);
}

int aadg_add_bf16(void* a, int lda, const void* b, int ldb, long long pixels, int c, void* stream) {
  NN_REQ_C(c);
  add_kernel<<<grid_for(pixels * (c / 8)), 256, 0, (cudaStream_t)stream>>>((bf16*)a, lda, (const bf16*)b, ldb, pixels, c);
  return check_launch("add");
}

int aadg_maxpool3x3s2_fwd(const void* x, int n, int h, int w, int c, void* y, void* argmax, void* stream) {
  NN_REQ_C(c);
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  maxpool_fwd_kernel<<<grid_for((long long)n * ho * wo * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, n, h, w, c, (bf16*)y, (unsigned char*)argmax, ho, wo);
  return check_launch("maxpool fwd");
}
int aadg_maxpool3x3s2_bwd(const void* dy, const void* argmax, int n, int h, int w, int c, void* dx, void* stream) {
  NN_REQ_C(c);
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  maxpool_bwd_kernel<<<grid_for((long long)n * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, (const unsigned char*)argmax, n, h, w, c, ho, wo, (bf16*)dx);
  return check_launch("maxpool bwd");
}

/* nearest x2 up-sampling: y bf16 [n,2h,2w,ldy] (channel slice) <- x bf16 [n,h,w,ldx]; and its transpose */
int aadg_upsample_nearest2x_fwd(const void* x, int n, int h, int w, int c, int ldx, void* y, int ldy, void* stream) {
  NN_REQ_C(c);
  AADG_REQUIRE(2 * h <= 65535 && n <= 65535, "tensor too large for the grid");
  dim3 grid(std::max(1, std::min((2 * w * (c / 8) + 255) / 256, 64)), 2 * h, n);
  nearest2x_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, n, h, w, c, ldx, (bf16*)y, ldy);
  return check_launch("nearest2x fwd");
}
int aadg_upsample_nearest2x_bwd(const void* dy, int n, int h, int w, int c, int lddy, void* dx, int lddx, void* stream) {
  NN_REQ_C(c);
  AADG_REQUIRE(h <= 65535 && n <= 65535, "tensor too large for the grid");
  dim3 grid(std::max(1, std::min((w * (c / 8) + 255) / 256, 64)), h, n);
  nearest2x_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, n, h, w, c, lddy, (bf16*)dx, lddx);
  return check_launch("nearest2x bwd");
}
int aadg_copy_bf16(const void* x, int ldx, void* y, int ldy, long long pixels, int c, void* stream) {
  NN_REQ_C(c);
  copy_kernel<<<grid_for(pixels * (c / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, pixels, c);
  return check_launch("copy");
}

int aadg_upsample_bilinear_fwd(const void* x, int n, int h, int w, int c, int ldx, void* y, int ho, int wo, int ldy,
                               void* stream) {
  NN_REQ_C(c);
  upsample_fwd_kernel<<<grid_for((long long)n * ho * ((wo + 3) / 4) * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, n, h, w, c, ldx, (bf16*)y, ho, wo, ldy);
  return check_launch("upsample fwd");
}
int aadg_upsample_bilinear_bwd(const void* dy, int n, int ho, int wo, int c, int lddy, void* dx, int h, int w, int lddx,
                               void* stream) {
  NN_REQ_C(c);
  upsample_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, n, ho, wo, c, lddy, (bf16*)dx, h, w, lddx);
  return check_launch("upsample bwd");
}

/* out fp32 [n,c] = scale * sum over the hw pixels of x bf16 [n,hw,ld] */
int aadg_global_sum(const void* x, int n, int hw, int c, int ld, float* out, float scale, void* stream) {
  NN_REQ_C(c);
  dim3 grid(n, (c / 8 + 31) / 32), blk(32, 8);
  gap_kernel<<<grid, blk, 0, (cudaStream_t)stream>>>((const bf16*)x, hw, c, ld, out, scale);
  return check_launch("global_sum");
}
int aadg_broadcast_pixels(const void* v, int n, int c, void* y, int hw, int ldy, void* stream) {
  NN_REQ_C(c);
  broadcast_kernel<<<grid_for((long long)n * hw * (c / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)v, c, (bf16*)y, hw, ldy,
                                                                                          (long long)n * hw);
  return check_launch("broadcast");
}
/* y bf16 [n, hw, ldy] (c channels) += scale * v fp32 [n, c] broadcast over the hw pixels */
int aadg_broadcast_add_pixels(const float* v, int n, int c, void* y, int hw, int ldy, float scale, void* stream) {
  NN_REQ_C(c);
  broadcast_add_kernel<<<grid_for((long long)n * hw * (c / 8)), 256, 0, (cudaStream_t)stream>>>(v, c, (bf16*)y, hw, ldy,
                                                                                              (long long)n * hw, scale);
  return check_launch("broadcast add");
}
int aadg_f32_to_bf16(const float* x, void* y, long long count, float scale, void* stream) {
  f32_to_bf16_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, count, scale);
  return check_launch("f32_to_bf16");
}

/* img fp32 [n,3,h,w] -> col bf16 [n*ho*wo][kp], k = (r*S+s)*3 + c, zero padded to kp (multiple of 8) */
int aadg_im2col_stem(const float* img, int n, int h, int w, int r, int s, int stride, int pad, int kp, void* col,
                     void* stream) {
  AADG_REQUIRE(kp % 8 == 0 && kp >= r * s * 3 && kp <= 512, "kp must be a multiple of 8, >= R*S*3 and <= 512");
  AADG_REQUIRE((size_t)r * 3 * (w + 2 * pad) < 32768, "image too wide for 16-bit patch offsets");
  const int ho = (h + 2 * pad - r) / stride + 1, wo = (w + 2 * pad - s) / stride + 1;
  const size_t smem = (size_t)r * 3 * (w + 2 * pad) * sizeof(bf16);
  AADG_REQUIRE(smem <= 200 * 1024, "image too wide for the staged im2col (%zu bytes of shared memory)", smem);
  static bool attr_set = false;
  if (!attr_set) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(im2col_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  im2col_stem_kernel<<<n * ho, 256, smem, (cudaStream_t)stream>>>(img, n, h, w, r, s, stride, pad, ho, wo, kp, (bf16*)col);
  return check_launch("im2col");
}

/* same patches in the row-pitched layout k = r*row_pitch + s*3 + c (row_pitch % 8 == 0, >= 3*s; kp % 8 == 0,
 * >= r*row_pitch; stride even): every filter row starts on a 16-byte boundary of the patch */
int aadg_im2col_stem_rows(const float* img, int n, int h, int w, int r, int s, int stride, int pad, int row_pitch, int kp,
                          void* col, void* stream) {
  AADG_REQUIRE(row_pitch % 8 == 0 && row_pitch >= 3 * s && kp % 8 == 0 && kp >= r * row_pitch,
               "row_pitch must be a multiple of 8 >= 3*S and kp a multiple of 8 >= R*row_pitch");
  AADG_REQUIRE(stride % 2 == 0 && stride > 0, "the row-pitched im2col needs an even stride");
  AADG_REQUIRE(((uintptr_t)col & 15) == 0, "col must be 16-byte aligned");
  const int ho = (h + 2 * pad - r) / stride + 1, wo = (w + 2 * pad - s) / stride + 1;
  const int pitch = ((w + 2 * pad) * 3 + row_pitch + 1) & ~1;
  const size_t smem = (size_t)r * pitch * sizeof(bf16);
  AADG_REQUIRE(smem <= 200 * 1024, "image too wide for the staged im2col (%zu bytes of shared memory)", smem);
  static bool attr_set = false;
  if (!attr_set) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(im2col_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  im2col_rows_kernel<<<n * ho, 256, smem, (cudaStream_t)stream>>>(img, n, h, w, r, s, stride, pad, ho, wo, row_pitch, kp,
                                                                 (bf16*)col);
  return check_launch("im2col rows");
}

/* fp32 NCHW image [n,3,h,w] (h, w even) -> bf16 space-to-depth stem input [n, h/2+3, w/2+3, 16] with its zero border
 * (see stem_s2d_kernel); `out` must have 64 spare elements after the tensor (the last overlapping windows end there) */
int aadg_stem_s2d(const float* img, int n, int h, int w, void* out, void* stream) {
  AADG_REQUIRE(img && out && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "stem input must be [n,3,h,w] with even h, w");
  AADG_REQUIRE(((uintptr_t)out & 15) == 0 && ((uintptr_t)img & 7) == 0, "stem tensors must be 16-byte (out) / 8-byte (img) aligned");
  const long long total = (long long)n * (h / 2 + 3) * (w / 2 + 3);
  stem_s2d_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(img, n, h, w, (bf16*)out);
  return check_launch("stem space-to-depth");
}

int aadg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long count, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  AADG_REQUIRE(step >= 1, "step counts from 1");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, count, lr, beta1,
                                                                beta2, eps, bc1, bc2, weight_decay);
  return check_launch("adam");
}

int aadg_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long count,
                       const float* hyper, const long long* step, void* stream) {
  AADG_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper && step, "null pointer");
  adam_dev_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, count, hyper, step);
  return check_launch("adam (device state)");
}

/* descs (device) int64x3 + int32x4 per weight: see WeightDesc */
int aadg_weight_prep(const float* master, void* w_bf16, void* w_bf16_t, const void* descs, int n_descs, void* stream) {
  if (n_descs <= 0) return AADG_OK;
  dim3 grid(64, n_descs);
  weight_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(master, (bf16*)w_bf16, (bf16*)w_bf16_t,
                                                            (const WeightDesc*)descs);
  return check_launch("weight_prep");
}

}  // extern "C"
