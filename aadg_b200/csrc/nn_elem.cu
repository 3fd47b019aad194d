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

// per-thread partial sums a0/a1[8] of channel group threadIdx.x -> out0/out1[C] (shared-memory then global atomics)
template <int NACC>
__device__ __forceinline__ void block_channel_sum(int C, const float* a0, const float* a1, float* out0, float* out1) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  __shared__ float s0[2048], s1[NACC > 1 ? 2048 : 1];
  for (int i = ty * blockDim.x + tx; i < C; i += blockDim.x * blockDim.y) { s0[i] = 0.f; if (NACC > 1) s1[i] = 0.f; }
  __syncthreads();
  if (tx < (C >> 3)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&s0[tx * 8 + i], a0[i]);
      if (NACC > 1) atomicAdd(&s1[tx * 8 + i], a1[i]);
    }
  }
  __syncthreads();
  for (int i = ty * blockDim.x + tx; i < C; i += blockDim.x * blockDim.y) {
    atomicAdd(&out0[i], s0[i]);
    if (NACC > 1) atomicAdd(&out1[i], s1[i]);
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
__global__ void bn_finalize_kernel(const float* sum, const float* sumsq, const float* gamma, const float* beta,
                                   int C, float count, float eps, float momentum, float* mean, float* invstd,
                                   float* scale, float* shift, float* run_mean, float* run_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = sum[c] / count;
  const float var = fmaxf(sumsq[c] / count - m * m, 0.f);
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
  if constexpr (MASK == 4) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d.v[i] = fmaf(xv.v[i], sc.v[i], sh.v[i]) > 0.f ? d.v[i] : 0.f;
  }
  if constexpr (MASK == 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d.v[i] = (m.b >> i) & 1 ? d.v[i] : 0.f;
  }
  if constexpr (MASK == 1) {
    const V8 yv = unpack8(m.v);
#pragma unroll
    for (int i = 0; i < 8; ++i) d.v[i] = yv.v[i] > 0.f ? d.v[i] : 0.f;
  }
  if (flags & 2) {
    const unsigned int keep = dropout_bits8(seed, (unsigned long long)e * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) d.v[i] = (keep >> i) & 1 ? d.v[i] * 2.f : 0.f;
  }
}

// dbeta = sum g ; dgamma = sum g * xhat  (`beta` = the forward shift vector when flag 4 is set)
template <int MASK>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const bf16* __restrict__ dy, int lddy,
                                                               const bf16* __restrict__ x, int ldx, const bf16* y, int ldy,
                                                               const float* mean, const float* invstd, const float* gamma,
                                                               const float* beta, long long P, int C, int flags,
                                                               unsigned long long seed, float* dgamma, float* dbeta) {
  const int G = C >> 3, g = threadIdx.x;
  float a0[8] = {}, a1[8] = {};
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
      uint4 dv[BN_U], xv[BN_U];
      BnMaskSrc<MASK> mv[BN_U];
#pragma unroll
      for (int u = 0; u < BN_U; ++u) {
        const long long p = p0 + u * step;
        if (p < P) {
          dv[u] = ldg16(dy + p * lddy + g * 8);
          xv[u] = ldg16(x + p * ldx + g * 8);
          bn_mask_load(mv[u], y, ldy, p, G, g);
        }
      }
#pragma unroll
      for (int u = 0; u < BN_U; ++u) {
        const long long p = p0 + u * step;
        if (p >= P) break;
        V8 d = unpack8(dv[u]);
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
template <int MASK>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(const bf16* __restrict__ dy, int lddy,
                                                              const bf16* __restrict__ x, int ldx, const bf16* y, int ldy,
                                                              const float* mean, const float* invstd, const float* gamma,
                                                              const float* beta, const float* dgamma, const float* dbeta,
                                                              long long P, int C, int flags, unsigned long long seed,
                                                              bf16* __restrict__ dx, int lddx, bf16* dres, int lddr,
                                                              int dres_accumulate) {
  const int G = C >> 3, g = threadIdx.x;
  if (g >= G) return;
  const float invP = 1.f / (float)P;
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
    uint4 dv[BN_U], xv[BN_U];
    BnMaskSrc<MASK> mv[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long p = p0 + u * step;
      if (p < P) {
        dv[u] = ldg16(dy + p * lddy + g * 8);
        xv[u] = ldg16(x + p * ldx + g * 8);
        bn_mask_load(mv[u], y, ldy, p, G, g);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long p = p0 + u * step;
      if (p >= P) break;
      V8 d = unpack8(dv[u]);
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
// gather form: input pixel (iy,ix) collects dy of every window whose arg-max it is
__global__ void maxpool_bwd_kernel(const bf16* dy, const unsigned char* arg, int N, int H, int W, int C, int Ho,
                                   int Wo, bf16* dx) {
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
    for (int oy = (iy) / 2; oy <= (iy + 1) / 2; ++oy) {
      if (oy >= Ho) continue;
      const int r = iy - (oy * 2 - 1);
      if (r < 0 || r > 2) continue;
      for (int ox = (ix) / 2; ox <= (ix + 1) / 2; ++ox) {
        if (ox >= Wo) continue;
        const int s = ix - (ox * 2 - 1);
        if (s < 0 || s > 2) continue;
        const size_t o = (((size_t)n * Ho + oy) * Wo + ox) * C + g * 8;
        const uint2 packed = *reinterpret_cast<const uint2*>(arg + o);
        const V8 d = ld8(dy + o);
        const int code = r * 3 + s;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int a = ((i < 4 ? packed.x : packed.y) >> ((i & 3) * 8)) & 0xff;
          if (a == code) acc.v[i] += d.v[i];
        }
      }
    }
    st8(dx + (((size_t)n * H + iy) * W + ix) * C + g * 8, acc);
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
__global__ void upsample_fwd_kernel(const bf16* x, int N, int H, int W, int C, int ldx, bf16* y, int Ho, int Wo, int ldy) {
  const int G = C >> 3;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)N * Ho * Wo * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % G);
    long long p = e / G;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    int y0, y1, x0, x1; float ly, lx;
    src_index(oy, sy, H, y0, y1, ly);
    src_index(ox, sx, W, x0, x1, lx);
    const bf16* b = x + (size_t)n * H * W * ldx + g * 8;
    const V8 v00 = ld8(b + ((size_t)y0 * W + x0) * ldx), v01 = ld8(b + ((size_t)y0 * W + x1) * ldx);
    const V8 v10 = ld8(b + ((size_t)y1 * W + x0) * ldx), v11 = ld8(b + ((size_t)y1 * W + x1) * ldx);
    V8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o.v[i] = (1.f - ly) * ((1.f - lx) * v00.v[i] + lx * v01.v[i]) + ly * ((1.f - lx) * v10.v[i] + lx * v11.v[i]);
    st8(y + (((size_t)n * Ho + oy) * Wo + ox) * ldy + g * 8, o);
  }
}
// gather form of the transpose: input pixel collects from every output pixel that samples it
__global__ void upsample_bwd_kernel(const bf16* dy, int N, int Ho, int Wo, int C, int lddy, bf16* dx, int H, int W, int lddx) {
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
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1; float ly;
      src_index(oy, sy, H, y0, y1, ly);
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1; float lx;
        src_index(ox, sx, W, x0, x1, lx);
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        const V8 d = ld8(dy + (((size_t)n * Ho + oy) * Wo + ox) * lddy + g * 8);
        const float wgt = wy * wx;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(wgt, d.v[i], acc.v[i]);
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
// fp32 [N, C] -> bf16 [N, C] (optionally scaled)
__global__ void f32_to_bf16_kernel(const float* x, bf16* y, long long n, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i] * scale);
}

// ---- depthwise 3x3 (dilated, stride 1, "same" padding) ---------------------------------------------------
// y[n,oy,ox,c] = sum_t w[t][c] * x[n, oy + (r-1)*dil*sign, ox + (s-1)*dil*sign, c]; sign = -1 gives the data gradient
// generic (dilated) form: one output pixel x 8 channels per thread (filter taps come from L1)
__global__ void dw3x3_kernel(const bf16* x, int N, int H, int W, int C, int ldx, const float* s_w, int dil, int sign,
                             bf16* y, int ldy) {
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
    st8(yrow + (size_t)ox * ldy + g * 8, acc);
  }
}
// ---- dilation-1 depthwise 3x3 with a rolling three-row window in shared memory --------------------------------
// block = (image, 32-pixel column strip, 64-channel chunk) walking down all rows: every input element is read
// from global memory once (plus the 2-pixel halo), the nine taps come from shared memory (conflict-free 16-byte
// reads), the thread's 9 x 8 filter taps (forward / data gradient) or 9 x 8 accumulators (weight gradient) live
// in registers.  256 threads = 32 pixels x 8 channel groups.
constexpr int DWR_XT = 32, DWR_CB = 64;
struct DwRows {
  const bf16* x; int ldx;
  int N, H, W, C;
};
__device__ __forceinline__ void dwr_load_row(bf16 (*ring)[DWR_XT + 2][DWR_CB], const DwRows& a, int n, int row, int x0,
                                             int c0) {
  // row `row` of the strip (with halo) -> ring[(row + 3) % 3]; out-of-image -> zeros
  bf16(*dst)[DWR_CB] = ring[(row + 3) % 3];
  for (int i = threadIdx.x; i < (DWR_XT + 2) * (DWR_CB / 8); i += 256) {
    const int px = i / (DWR_CB / 8), g = i % (DWR_CB / 8);
    const int ix = x0 - 1 + px, c = c0 + g * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row >= 0 && row < a.H && ix >= 0 && ix < a.W && c < a.C)
      v = *reinterpret_cast<const uint4*>(a.x + (((size_t)n * a.H + row) * a.W + ix) * a.ldx + c);
    *reinterpret_cast<uint4*>(&dst[px][g * 8]) = v;
  }
}
// forward (flip = 0) / data gradient (flip = 1: the filter is rotated by 180 degrees)
__global__ void __launch_bounds__(256) dw3x3_rows_kernel(const DwRows a, const float* w, int flip, bf16* y, int ldy) {
  __shared__ __align__(16) bf16 ring[3][DWR_XT + 2][DWR_CB];
  const int n = blockIdx.z, x0 = blockIdx.x * DWR_XT, c0 = blockIdx.y * DWR_CB;
  const int px = threadIdx.x >> 3, g = threadIdx.x & 7;
  const int c = c0 + g * 8, ox = x0 + px;
  const bool active = c < a.C && ox < a.W;
  float wt[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const V8 v = c < a.C ? ld8f(w + (size_t)(flip ? 8 - t : t) * a.C + c) : V8{};
#pragma unroll
    for (int i = 0; i < 8; ++i) wt[t][i] = v.v[i];
  }
  dwr_load_row(ring, a, n, -1, x0, c0);
  dwr_load_row(ring, a, n, 0, x0, c0);
  for (int oy = 0; oy < a.H; ++oy) {
    dwr_load_row(ring, a, n, oy + 1, x0, c0);
    __syncthreads();
    if (active) {
      float acc[8] = {};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const bf16(*row)[DWR_CB] = ring[(oy + r - 1 + 3) % 3];
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const V8 v = ld8(&row[px + s2][g * 8]);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(wt[r * 3 + s2][i], v.v[i], acc[i]);
        }
      }
      V8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = acc[i];
      st8(y + (((size_t)n * a.H + oy) * a.W + ox) * ldy + c, o);
    }
    __syncthreads();
  }
}
// weight gradient: dw[t][c] += sum over the strip of dy[p][c] * x[p + tap t][c]
__global__ void __launch_bounds__(256) dw3x3_rows_wgrad_kernel(const DwRows a, const bf16* dy, int lddy, float* dw) {
  __shared__ __align__(16) bf16 ring[3][DWR_XT + 2][DWR_CB];
  __shared__ float red[9][DWR_CB];
  const int n = blockIdx.z, x0 = blockIdx.x * DWR_XT, c0 = blockIdx.y * DWR_CB;
  const int px = threadIdx.x >> 3, g = threadIdx.x & 7;
  const int c = c0 + g * 8, ox = x0 + px;
  const bool active = c < a.C && ox < a.W;
  float acc[9][8] = {};
  for (int i = threadIdx.x; i < 9 * DWR_CB; i += 256) (&red[0][0])[i] = 0.f;
  dwr_load_row(ring, a, n, -1, x0, c0);
  dwr_load_row(ring, a, n, 0, x0, c0);
  for (int oy = 0; oy < a.H; ++oy) {
    dwr_load_row(ring, a, n, oy + 1, x0, c0);
    V8 d = {};
    if (active) d = ld8(dy + (((size_t)n * a.H + oy) * a.W + ox) * lddy + c);
    __syncthreads();
    if (active) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const bf16(*row)[DWR_CB] = ring[(oy + r - 1 + 3) % 3];
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const V8 v = ld8(&row[px + s2][g * 8]);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[r * 3 + s2][i] = fmaf(d.v[i], v.v[i], acc[r * 3 + s2][i]);
        }
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&red[t][g * 8 + i], acc[t][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * DWR_CB; i += 256) {
    const int t = i / DWR_CB, cc = c0 + i % DWR_CB;
    if (cc < a.C) atomicAdd(&dw[(size_t)t * a.C + cc], (&red[0][0])[i]);
  }
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
static int tuning_dw_rows() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AADG_DW_ROWS"); v = e ? atoi(e) : 1; }
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

int aadg_bn_finalize(const float* sum, const float* sumsq, const float* gamma, const float* beta, int c, float count,
                     float eps, float momentum, float* mean, float* invstd, float* scale, float* shift,
                     float* run_mean, float* run_var, void* stream) {
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, gamma, beta, c, count, eps, momentum,
                                                                        mean, invstd, scale, shift, run_mean, run_var);
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

int aadg_bn_backward(const void* dy, int lddy, const void* x, int ldx, const void* y, int ldy, const float* mean,
                     const float* invstd, const float* gamma, const float* shift, long long pixels, int c, int flags,
                     unsigned long long seed, float* dgamma, float* dbeta, void* dx, int lddx, void* dres, int lddr,
                     int dres_accumulate, void* stream) {
  NN_REQ_C(c);
  AADG_REQUIRE(pixels > 0 && pixels < (1ll << 31), "bad pixel count");
  cudaStream_t st = (cudaStream_t)stream;
  AADG_CUDA_TRY(cudaMemsetAsync(dgamma, 0, sizeof(float) * c, st));
  AADG_CUDA_TRY(cudaMemsetAsync(dbeta, 0, sizeof(float) * c, st));
  const dim3 blk = reduce_block(c);
  AADG_REQUIRE(!(flags & 1) || (flags & 4) || y, "ReLU mask needs y / the bit mask (or flag 4 to recompute it from x)");
  AADG_REQUIRE(!(flags & 4) || shift, "flag 4 needs the forward shift vector");
  const int mask = (flags & 4) ? 4 : (flags & 8) ? 8 : (flags & 1) ? 1 : 0;
  const int grid = stream_blocks(pixels, blk, 2);
#define AADG_BN_BWD(MASK)                                                                                             \
  {                                                                                                                   \
    bn_bwd_reduce_kernel<MASK><<<grid, blk, 0, st>>>((const bf16*)dy, lddy, (const bf16*)x, ldx, (const bf16*)y, ldy,  \
                                                     mean, invstd, gamma, shift, pixels, c, flags, seed, dgamma, dbeta); \
    bn_bwd_apply_kernel<MASK><<<grid, blk, 0, st>>>((const bf16*)dy, lddy, (const bf16*)x, ldx, (const bf16*)y, ldy,   \
                                                    mean, invstd, gamma, shift, dgamma, dbeta, pixels, c, flags, seed, \
                                                    (bf16*)dx, lddx, (bf16*)dres, lddr, dres_accumulate);              \
  }
  if (mask == 4) AADG_BN_BWD(4) else if (mask == 8) AADG_BN_BWD(8) else if (mask == 1) AADG_BN_BWD(1) else AADG_BN_BWD(0)
#undef AADG_BN_BWD
  return check_launch("bn_backward");
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
  maxpool_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
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
  upsample_fwd_kernel<<<grid_for((long long)n * ho * wo * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
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
int aadg_f32_to_bf16(const float* x, void* y, long long count, float scale, void* stream) {
  f32_to_bf16_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, count, scale);
  return check_launch("f32_to_bf16");
}

/* depthwise 3x3, stride 1, padding = dilation. direction 0: forward, 1: data gradient. w fp32 [9][c] */
int aadg_dwconv3x3(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int direction, void* y,
                   int ldy, void* stream) {
  NN_REQ_C(c);
  AADG_REQUIRE(h <= 65535 && n <= 65535, "image too tall / batch too large for the depthwise grid");
  const size_t smem = 0;
  if (dil == 1 && tuning_dw_rows()) {
    DwRows a{(const bf16*)x, ldx, n, h, w, c};
    dim3 grid((w + DWR_XT - 1) / DWR_XT, (c + DWR_CB - 1) / DWR_CB, n);
    dw3x3_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, wgt, direction ? 1 : 0, (bf16*)y, ldy);
  } else {
    dim3 grid(std::max(1, std::min((w * (c / 8) + 255) / 256, 64)), h, n);
    dw3x3_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, n, h, w, c, ldx, wgt, dil,
                                                           direction ? -1 : 1, (bf16*)y, ldy);
  }
  return check_launch("dwconv3x3");
}
int aadg_dwconv3x3_wgrad(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int lddy, int dil, float* dw,
                         void* stream) {
  NN_REQ_C(c);
  const long long pixels = (long long)n * h * w;
  AADG_REQUIRE(pixels < (1ll << 31), "too many pixels");
  if (dil == 1 && tuning_dw_rows()) {
    DwRows a{(const bf16*)x, ldx, n, h, w, c};
    dim3 grid((w + DWR_XT - 1) / DWR_XT, (c + DWR_CB - 1) / DWR_CB, n);
    dw3x3_rows_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, (const bf16*)dy, lddy, dw);
    return check_launch("dwconv3x3 rows wgrad");
  }
  const dim3 blk = reduce_block(c);
  const int blocks = (int)std::min<long long>((pixels + blk.y * 16 - 1) / (blk.y * 16), 148 * 6);
  const size_t smem = (size_t)3 * c * sizeof(float);
  dim3 grid(std::max(blocks, 1), 3);
  dw3x3_wgrad_kernel<<<grid, blk, smem, (cudaStream_t)stream>>>((const bf16*)x, n, h, w, c, ldx, (const bf16*)dy, lddy,
                                                               dil, dw);
  return check_launch("dwconv3x3 wgrad");
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

int aadg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long count, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  AADG_REQUIRE(step >= 1, "step counts from 1");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, count, lr, beta1,
                                                                beta2, eps, bc1, bc2, weight_decay);
  return check_launch("adam");
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
