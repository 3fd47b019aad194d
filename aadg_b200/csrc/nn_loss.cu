// Segmentation head and loss for sm_100a: 1x1 head convolution to `classes` logits at decoder
// resolution, then ONE kernel for  UpsamplingBilinear2d(x4) -> sigmoid -> BCELoss(mean)  plus the
// per-sample hard-Dice counts, and its gather-form backward.  Replaces smp's SegmentationHead and the
// loss/metric lines of the reference step (search_dg.py:140-142,164-165,170; losses.py:21-23), without
// materialising the [B*D*M, classes, H, W] logits, probabilities and their gradients in HBM.
#include <cuda_bf16.h>
#include <algorithm>

#include "common.cuh"

namespace aadg {
namespace nnl {

typedef __nv_bfloat16 bf16;
constexpr int MAXK = 2;

// z[p][k] = b[k] + sum_c w[k][c] * a[p][c] ; one warp per pixel, HEAD_U pixels per trip with the loads issued
// first; for C <= 256 (one 16-byte load per lane) the lane's 8 x K weights stay in registers
constexpr int HEAD_U = 4;
template <bool SMALL_C>
__global__ void __launch_bounds__(256) head_fwd_kernel(const bf16* __restrict__ a, long long P, int C, int lda,
                                                       const float* __restrict__ w, const float* __restrict__ b, int K,
                                                       float* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float wk[MAXK][8] = {};
  if (SMALL_C && lane * 8 < C)
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) wk[k][i] = w[k * C + lane * 8 + i];
  const float b0 = b[0], b1 = K > 1 ? b[1] : 0.f;
  for (long long p0 = warp; p0 < P; p0 += HEAD_U * nwarps) {
    float acc[HEAD_U][MAXK] = {};
    if (SMALL_C) {
      uint4 u[HEAD_U];
#pragma unroll
      for (int j = 0; j < HEAD_U; ++j) {
        const long long p = p0 + j * nwarps;
        u[j] = (p < P && lane * 8 < C) ? __ldg(reinterpret_cast<const uint4*>(a + p * lda + lane * 8)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < HEAD_U; ++j) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
#pragma unroll
          for (int k = 0; k < MAXK; ++k) acc[j][k] = fmaf(f.x, wk[k][2 * i], fmaf(f.y, wk[k][2 * i + 1], acc[j][k]));
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < HEAD_U; ++j) {
        const long long p = p0 + j * nwarps;
        if (p >= P) break;
        for (int c = lane * 8; c < C; c += 256) {
          const uint4 u = *reinterpret_cast<const uint4*>(a + p * lda + c);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
#pragma unroll
            for (int k = 0; k < MAXK; ++k)
              if (k < K) acc[j][k] = fmaf(f.x, w[k * C + c + 2 * i], fmaf(f.y, w[k * C + c + 2 * i + 1], acc[j][k]));
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < HEAD_U; ++j) {
      const long long p = p0 + j * nwarps;
      const float t0 = warp_sum(acc[j][0]), t1 = warp_sum(acc[j][1]);
      if (lane == 0 && p < P) {
        z[p * K] = t0 + b0;
        if (K > 1) z[p * K + 1] = t1 + b1;
      }
    }
  }
}

__device__ __forceinline__ void src_index(int o, float scale, int in, int& i0, int& i1, float& lam) {
  const float s = scale * (float)o;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = min(i0 + 1, in - 1);
  lam = s - (float)i0;
}
__device__ __forceinline__ float bilerp(const float* z, int W, int K, int k, int y0, int y1, int x0, int x1, float ly,
                                        float lx) {
  const float v00 = z[((size_t)y0 * W + x0) * K + k], v01 = z[((size_t)y0 * W + x1) * K + k];
  const float v10 = z[((size_t)y1 * W + x0) * K + k], v11 = z[((size_t)y1 * W + x1) * K + k];
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

// z low-res fp32 [N,h,w,K]; target fp32 [N,K,H,W]; loss_sum double; counts int [N][K][3] = TP, FP, FN
// (prediction = p > thr); logits_out optional fp32 [N,K,H,W]
__global__ void loss_fwd_kernel(const float* z, int N, int h, int w, int K, const float* target, int H, int W,
                                float thr, double* loss_sum, int* counts, float* logits_out) {
  const int n = blockIdx.z, k = blockIdx.y;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const float* zn = z + (size_t)n * h * w * K;
  const float* tn = target + ((size_t)n * K + k) * H * W;
  float lsum = 0.f;
  int tp = 0, fp = 0, fn = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int oy = i / W, ox = i % W;
    int y0, y1, x0, x1; float ly, lx;
    src_index(oy, sy, h, y0, y1, ly);
    src_index(ox, sx, w, x0, x1, lx);
    const float zu = bilerp(zn, w, K, k, y0, y1, x0, x1, ly, lx);
    if (logits_out) logits_out[((size_t)n * K + k) * H * W + i] = zu;
    const float p = sigmoidf_(zu), t = tn[i];
    // torch.nn.BCELoss: log terms clamped at -100
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    lsum -= t * lp + (1.f - t) * l1p;
    const bool pred = p > thr, pos = t > 0.5f;
    tp += pred && pos; fp += pred && !pos; fn += !pred && pos;
  }
  lsum = warp_sum(lsum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o); fp += __shfl_xor_sync(0xffffffffu, fp, o);
    fn += __shfl_xor_sync(0xffffffffu, fn, o);
  }
  // one set of atomics per CTA (147 456 warps hammering a single double was most of this kernel's time)
  __shared__ float s_l[8];
  __shared__ int s_c[8][3];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_l[wid] = lsum; s_c[wid][0] = tp; s_c[wid][1] = fp; s_c[wid][2] = fn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0;
    int c0 = 0, c1 = 0, c2 = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { l += (double)s_l[i]; c0 += s_c[i][0]; c1 += s_c[i][1]; c2 += s_c[i][2]; }
    atomicAdd(loss_sum, l);
    int* c = counts + ((size_t)n * K + k) * 3;
    if (c0) atomicAdd(c, c0);
    if (c1) atomicAdd(c + 1, c1);
    if (c2) atomicAdd(c + 2, c2);
  }
}

// dz[n,iy,ix,k] = sum over outputs sampling (iy,ix) of weight * (p - t) * sat * gscale
__global__ void loss_bwd_kernel(const float* z, int N, int h, int w, int K, const float* target, int H, int W,
                                float gscale, float* dz) {
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const long long total = (long long)N * h * w * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    long long q = e / K;
    const int ix = (int)(q % w); q /= w;
    const int iy = (int)(q % h);
    const int n = (int)(q / h);
    const float* zn = z + (size_t)n * h * w * K;
    const float* tn = target + ((size_t)n * K + k) * H * W;
    const int oy_lo = sy > 0.f ? max(0, (int)floorf((float)(iy - 1) / sy)) : 0;
    const int oy_hi = sy > 0.f ? min(H - 1, (int)ceilf((float)(iy + 1) / sy)) : H - 1;
    const int ox_lo = sx > 0.f ? max(0, (int)floorf((float)(ix - 1) / sx)) : 0;
    const int ox_hi = sx > 0.f ? min(W - 1, (int)ceilf((float)(ix + 1) / sx)) : W - 1;
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1; float ly;
      src_index(oy, sy, h, y0, y1, ly);
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1; float lx;
        src_index(ox, sx, w, x0, x1, lx);
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        const float zu = bilerp(zn, w, K, k, y0, y1, x0, x1, ly, lx);
        const float p = sigmoidf_(zu), t = tn[(size_t)oy * W + ox];
        // BCELoss backward divides by max(p(1-p), 1e-12); sigmoid backward multiplies by p(1-p)
        const float pq = p * (1.f - p);
        acc = fmaf(wy * wx, (p - t) * (pq / fmaxf(pq, 1e-12f)), acc);
      }
    }
    dz[e] = acc * gscale;
  }
}

// transpose of the align_corners bilinear up-sampling for an ARBITRARY incoming gradient: dz[n,iy,ix,k] = sum over the
// outputs (oy,ox) that sample cell (iy,ix) of weight * dlogits[n,k,oy,ox] (the autograd surface: the caller computed
// its own loss on the full-resolution logits, search_dg.py:140-142,170)
__global__ void upsample_logits_bwd_kernel(const float* dlogits, int N, int h, int w, int K, int H, int W, float* dz) {
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const long long total = (long long)N * h * w * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    long long q = e / K;
    const int ix = (int)(q % w); q /= w;
    const int iy = (int)(q % h);
    const int n = (int)(q / h);
    const float* gn = dlogits + ((size_t)n * K + k) * H * W;
    const int oy_lo = sy > 0.f ? max(0, (int)floorf((float)(iy - 1) / sy)) : 0;
    const int oy_hi = sy > 0.f ? min(H - 1, (int)ceilf((float)(iy + 1) / sy)) : H - 1;
    const int ox_lo = sx > 0.f ? max(0, (int)floorf((float)(ix - 1) / sx)) : 0;
    const int ox_hi = sx > 0.f ? min(W - 1, (int)ceilf((float)(ix + 1) / sx)) : W - 1;
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1; float ly;
      src_index(oy, sy, h, y0, y1, ly);
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1; float lx;
        src_index(ox, sx, w, x0, x1, lx);
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        acc = fmaf(wy * wx, gn[(size_t)oy * W + ox], acc);
      }
    }
    dz[e] = acc;
  }
}

// Scatter form of the same gradient for the up-sampling factor smp uses (H = 4h, W = 4w): a CTA owns a 64 x 64
// block of output pixels of one (image, class); a thread owns a 4 x 4 patch (four float4 target loads), evaluates
// the up-sampled logit from a shared-memory copy of the low-resolution logits, folds its 16 gradients into the
// <= 3 x 3 low-resolution cells they touch in registers, and adds those to a shared-memory tile; the tile leaves
// with one global atomic per cell (cells on a block border are shared with the neighbouring CTA; dz is zeroed first).
constexpr int LB_T = 64, LB_Z = 20;
__global__ void __launch_bounds__(256) loss_bwd_x4_kernel(const float* __restrict__ z, int h, int w, int K,
                                                          const float* __restrict__ target, int H, int W, float gscale,
                                                          float* dz) {
  __shared__ float sz[LB_Z][LB_Z], sacc[LB_Z][LB_Z];
  const int n = blockIdx.z, k = blockIdx.y;
  const int tiles_x = (W + LB_T - 1) / LB_T;
  const int X0 = (blockIdx.x % tiles_x) * LB_T, Y0 = (blockIdx.x / tiles_x) * LB_T;
  const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
  int yb, xb, t1; float tl;
  src_index(Y0, sy, h, yb, t1, tl);
  src_index(X0, sx, w, xb, t1, tl);
  const float* zn = z + (size_t)n * h * w * K + k;
  for (int i = threadIdx.x; i < LB_Z * LB_Z; i += 256) {
    const int r = i / LB_Z, c = i - r * LB_Z;
    const int yy = min(yb + r, h - 1), xx = min(xb + c, w - 1);
    sz[r][c] = zn[((size_t)yy * w + xx) * K];
    sacc[r][c] = 0.f;
  }
  const int oy0 = Y0 + (threadIdx.x >> 4) * 4, ox0 = X0 + (threadIdx.x & 15) * 4;
  const bool active = oy0 < H && ox0 < W;      // H, W multiples of 4: a patch is entirely inside or outside
  const float* tn = target + ((size_t)n * K + k) * H * W;
  float4 tv[4];
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) tv[j] = __ldg(reinterpret_cast<const float4*>(tn + (size_t)(oy0 + j) * W + ox0));
  }
  __syncthreads();
  if (active) {
    int py, px;
    src_index(oy0, sy, h, py, t1, tl);
    src_index(ox0, sx, w, px, t1, tl);
    int x0[4], x1[4]; float lx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) src_index(ox0 + i, sx, w, x0[i], x1[i], lx[i]);
    float a[3][3] = {};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int y0, y1; float ly;
      src_index(oy0 + j, sy, h, y0, y1, ly);
      const float tj[4] = {tv[j].x, tv[j].y, tv[j].z, tv[j].w};
      float row[3] = {};       // this output row's gradients folded onto the three low-resolution columns
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v00 = sz[y0 - yb][x0[i] - xb], v01 = sz[y0 - yb][x1[i] - xb];
        const float v10 = sz[y1 - yb][x0[i] - xb], v11 = sz[y1 - yb][x1[i] - xb];
        const float zu = (1.f - ly) * ((1.f - lx[i]) * v00 + lx[i] * v01) + ly * ((1.f - lx[i]) * v10 + lx[i] * v11);
        const float pr = sigmoidf_(zu);
        const float pq = pr * (1.f - pr);
        const float gr = (pr - tj[i]) * (pq / fmaxf(pq, 1e-12f));
        const int c0 = x0[i] - px, c1 = x1[i] - px;
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] += ((c == c0 ? 1.f - lx[i] : 0.f) + (c == c1 ? lx[i] : 0.f)) * gr;
      }
      const int r0 = y0 - py, r1 = y1 - py;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float wr = (r == r0 ? 1.f - ly : 0.f) + (r == r1 ? ly : 0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) a[r][c] = fmaf(wr, row[c], a[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (a[r][c] != 0.f) atomicAdd(&sacc[py - yb + r][px - xb + c], a[r][c]);
  }
  __syncthreads();
  float* dzn = dz + (size_t)n * h * w * K + k;
  for (int i = threadIdx.x; i < LB_Z * LB_Z; i += 256) {
    const int r = i / LB_Z, c = i - r * LB_Z;
    const float v = sacc[r][c];
    if (v != 0.f && yb + r < h && xb + c < w) atomicAdd(&dzn[((size_t)(yb + r) * w + xb + c) * K], v * gscale);
  }
}

// da[p][c] = sum_k dz[p][k] w[k][c] (bf16) ; dw[k][c] += sum_p dz[p][k] a[p][c] ; db[k] += sum_p dz[p][k]
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dz, const bf16* __restrict__ a, long long P,
                                                       int C, int lda, const float* w, int K, bf16* __restrict__ da, int ldda,
                                                       float* dw, float* db) {
  const int G = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y;
  float acc[MAXK][8] = {};
  float accb[MAXK] = {};
  float wk[MAXK][8] = {};
  if (tx < G) {
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < K) {
#pragma unroll
        for (int i = 0; i < 8; ++i) wk[k][i] = w[k * C + tx * 8 + i];
      }
  }
  if (tx < G) {
    constexpr int U = 4;
    const long long step = (long long)gridDim.x * blockDim.y;
    for (long long p0 = (long long)blockIdx.x * blockDim.y + ty; p0 < P; p0 += U * step) {
      uint4 u[U];
      float d[U][MAXK];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long p = p0 + j * step;
        u[j] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < MAXK; ++k) d[j][k] = 0.f;
        if (p < P) {
          u[j] = __ldg(reinterpret_cast<const uint4*>(a + p * lda + tx * 8));
#pragma unroll
          for (int k = 0; k < MAXK; ++k)
            if (k < K) d[j][k] = __ldg(dz + p * K + k);
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long p = p0 + j * step;
        if (p >= P) break;
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u[j]);
        float av[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(hh[i]); av[2 * i] = f.x; av[2 * i + 1] = f.y; }
        float o[8] = {};
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
          if (tx == 0) accb[k] += d[j][k];
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc[k][i] = fmaf(d[j][k], av[i], acc[k][i]); o[i] = fmaf(d[j][k], wk[k][i], o[i]); }
        }
        uint4 ou;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ou);
#pragma unroll
        for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(da + p * ldda + tx * 8) = ou;
      }
    }
  }
  __shared__ float s[MAXK][2048];
  __shared__ float sb[MAXK];
  for (int i = ty * blockDim.x + tx; i < MAXK * 2048; i += blockDim.x * blockDim.y) (&s[0][0])[i] = 0.f;
  if (tx == 0 && ty == 0) for (int k = 0; k < MAXK; ++k) sb[k] = 0.f;
  __syncthreads();
  if (tx < G) {
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < K) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&s[k][tx * 8 + i], acc[k][i]);
        if (tx == 0) atomicAdd(&sb[k], accb[k]);
      }
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    for (int i = ty * blockDim.x + tx; i < C; i += blockDim.x * blockDim.y) atomicAdd(&dw[k * C + i], s[k][i]);
    if (tx == 0 && ty == 0) atomicAdd(&db[k], sb[k]);
  }
}

// ---- 3x3 head (smp Unet SegmentationHead: Conv2d(16, classes, 3, padding=1)) --------------------------------
// z[n,y,x,k] = b[k] + sum_{r,s,c} w[k][r*3+s][c] * a[n,y+r-1,x+s-1,c] ; C <= 64, one thread per pixel
__global__ void head3x3_fwd_kernel(const bf16* a, int N, int H, int W, int C, int lda, const float* w, const float* b,
                                   int K, float* z) {
  extern __shared__ float sw[];          // [K][9][C]
  for (int i = threadIdx.x; i < K * 9 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long P = (long long)N * H * W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % W), y = (int)((p / W) % H);
    float acc[MAXK] = {};
    for (int r = 0; r < 3; ++r) {
      const int iy = y + r - 1;
      if (iy < 0 || iy >= H) continue;
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ix = x + s2 - 1;
        if (ix < 0 || ix >= W) continue;
        const bf16* ap = a + (p + (long long)(r - 1) * W + (s2 - 1)) * lda;
        for (int c = 0; c < C; c += 8) {
          const uint4 u = *reinterpret_cast<const uint4*>(ap + c);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h2[i]);
#pragma unroll
            for (int k = 0; k < MAXK; ++k)
              if (k < K) {
                const float* wk = sw + (k * 9 + r * 3 + s2) * C + c + 2 * i;
                acc[k] = fmaf(f.x, wk[0], fmaf(f.y, wk[1], acc[k]));
              }
          }
        }
      }
    }
    for (int k = 0; k < K; ++k) z[p * K + k] = acc[k] + b[k];
  }
}
// da[p][c] = sum_{r,s,k} dz[p - (r-1,s-1)][k] * w[k][r*3+s][c]
__global__ void head3x3_dgrad_kernel(const float* dz, int N, int H, int W, int C, const float* w, int K, bf16* da,
                                     int ldda) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < K * 9 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long P = (long long)N * H * W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % W), y = (int)((p / W) % H);
    float acc[64];
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int r = 0; r < 3; ++r) {
      const int oy = y - (r - 1);
      if (oy < 0 || oy >= H) continue;
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ox = x - (s2 - 1);
        if (ox < 0 || ox >= W) continue;
        const float* dp = dz + (p - (long long)(r - 1) * W - (s2 - 1)) * K;
        for (int k = 0; k < K; ++k) {
          const float d = dp[k];
          const float* wk = sw + (k * 9 + r * 3 + s2) * C;
          for (int c = 0; c < C; ++c) acc[c] = fmaf(d, wk[c], acc[c]);
        }
      }
    }
    for (int c = 0; c < C; c += 2)
      *reinterpret_cast<__nv_bfloat162*>(da + p * ldda + c) = __floats2bfloat162_rn(acc[c], acc[c + 1]);
  }
}
// dw[k][tap][c] += sum_p dz[p][k] * a[p + tap][c] ; db[k] += sum_p dz[p][k] ; grid (blocks, 9 taps)
__global__ void head3x3_wgrad_kernel(const float* dz, const bf16* a, int N, int H, int W, int C, int lda, int K,
                                     float* dw, float* db) {
  const int tap = blockIdx.y, r = tap / 3, s2 = tap % 3;
  const long long P = (long long)N * H * W;
  float acc[MAXK][64];
  float accb[MAXK] = {};
  for (int k = 0; k < MAXK; ++k) for (int c = 0; c < C; ++c) acc[k][c] = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % W), y = (int)((p / W) % H);
    float d[MAXK];
    for (int k = 0; k < K; ++k) { d[k] = dz[p * K + k]; if (tap == 4) accb[k] += d[k]; }
    const int iy = y + r - 1, ix = x + s2 - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const bf16* ap = a + (p + (long long)(r - 1) * W + (s2 - 1)) * lda;
    for (int c = 0; c < C; c += 2) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ap + c));
      for (int k = 0; k < K; ++k) { acc[k][c] = fmaf(d[k], f.x, acc[k][c]); acc[k][c + 1] = fmaf(d[k], f.y, acc[k][c + 1]); }
    }
  }
  for (int k = 0; k < K; ++k) {
    for (int c = 0; c < C; ++c) {
      const float t = warp_sum(acc[k][c]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&dw[(k * 9 + tap) * C + c], t);
    }
    if (tap == 4) {
      const float t = warp_sum(accb[k]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&db[k], t);
    }
  }
}

}  // namespace nnl
}  // namespace aadg

using namespace aadg;
using namespace aadg::nnl;

extern "C" {

/* logits at decoder resolution: z fp32 [pixels][classes] = bias + a bf16 [pixels][ld] . w fp32 [classes][c] */
int aadg_seg_head_fwd(const void* a, long long pixels, int c, int lda, const float* w, const float* bias, int classes,
                      float* z, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && c % 8 == 0 && c > 0, "classes must be 1..%d, channels a multiple of 8", MAXK);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((pixels * 32 + 255) / 256, 148 * 8));
  if (c <= 256)
    head_fwd_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)a, pixels, c, lda, w, bias, classes, z);
  else
    head_fwd_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)a, pixels, c, lda, w, bias, classes, z);
  return check_launch("seg_head_fwd");
}

/* upsample(align_corners=True) -> sigmoid -> BCELoss sum (double, accumulated) and per-(sample,class)
 * TP/FP/FN counts (int32 [n][classes][3], accumulated) at threshold thr; logits_out optional fp32
 * [n,classes,H,W].  target fp32 [n,classes,H,W]. */
int aadg_seg_loss_fwd(const float* z, int n, int h, int w, int classes, const float* target, int H, int W, float thr,
                      double* loss_sum, int* counts, float* logits_out, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && n > 0, "bad sizes");
  dim3 grid(std::min((H * W + 255) / 256, 64), classes, n);
  loss_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, n, h, w, classes, target, H, W, thr, loss_sum, counts, logits_out);
  return check_launch("seg_loss_fwd");
}

/* dz fp32 [n,h,w,classes] = d(grad_scale * BCE sum)/dz */
int aadg_seg_loss_bwd(const float* z, int n, int h, int w, int classes, const float* target, int H, int W,
                      float grad_scale, float* dz, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && n > 0, "bad sizes");
  const long long total = (long long)n * h * w * classes;
  if (H == 4 * h && W == 4 * w && h > 1 && w > 1 && n <= 65535 && ((uintptr_t)target & 15) == 0) {
    AADG_CUDA_TRY(cudaMemsetAsync(dz, 0, sizeof(float) * total, (cudaStream_t)stream));
    dim3 grid(((W + LB_T - 1) / LB_T) * ((H + LB_T - 1) / LB_T), classes, n);
    loss_bwd_x4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, h, w, classes, target, H, W, grad_scale, dz);
    return check_launch("seg_loss_bwd x4");
  }
  const int blocks = (int)std::min<long long>((total + 127) / 128, 148 * 32);
  loss_bwd_kernel<<<std::max(blocks, 1), 128, 0, (cudaStream_t)stream>>>(z, n, h, w, classes, target, H, W, grad_scale, dz);
  return check_launch("seg_loss_bwd");
}

/* dz fp32 [n,h,w,classes] = transpose of the head's bilinear (align_corners=True) up-sampling applied to an arbitrary
 * gradient dlogits fp32 [n,classes,H,W] */
int aadg_upsample_logits_bwd(const float* dlogits, int n, int h, int w, int classes, int H, int W, float* dz, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && n > 0 && h > 0 && w > 0 && H >= h && W >= w, "bad sizes");
  AADG_REQUIRE(dlogits && dz, "null pointer");
  const long long total = (long long)n * h * w * classes;
  const int blocks = (int)std::min<long long>((total + 127) / 128, 148 * 32);
  upsample_logits_bwd_kernel<<<std::max(blocks, 1), 128, 0, (cudaStream_t)stream>>>(dlogits, n, h, w, classes, H, W, dz);
  return check_launch("upsample_logits_bwd");
}

/* da bf16 [pixels][ldda] = dz . w ; dw fp32 [classes][c] += dz^T a ; db fp32 [classes] += sum dz */
int aadg_seg_head_bwd(const float* dz, const void* a, long long pixels, int c, int lda, const float* w, int classes,
                      void* da, int ldda, float* dw, float* db, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && c % 8 == 0 && c > 0 && c <= 2048, "bad sizes");
  int tx = 1;
  while (tx < (c >> 3)) tx <<= 1;
  tx = std::min(tx, 256);
  dim3 blk(tx, 256 / tx);
  const int blocks = (int)std::min<long long>((pixels + blk.y * 4 - 1) / (blk.y * 4), 148 * 4);
  head_bwd_kernel<<<std::max(blocks, 1), blk, 0, (cudaStream_t)stream>>>(dz, (const bf16*)a, pixels, c, lda, w, classes,
                                                                        (bf16*)da, ldda, dw, db);
  return check_launch("seg_head_bwd");
}

/* 3x3 segmentation head (padding 1) at full resolution: z fp32 [n,h,w,classes]; w fp32 [classes][9][c], c <= 64 */
int aadg_seg_head3x3_fwd(const void* a, int n, int h, int w, int c, int lda, const float* wgt, const float* bias,
                         int classes, float* z, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && c % 8 == 0 && c > 0 && c <= 64, "classes <= %d, channels <= 64", MAXK);
  const long long P = (long long)n * h * w;
  const int blocks = (int)std::min<long long>((P + 127) / 128, 148 * 16);
  head3x3_fwd_kernel<<<blocks, 128, classes * 9 * c * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)a, n, h, w, c, lda, wgt, bias, classes, z);
  return check_launch("head3x3 fwd");
}
int aadg_seg_head3x3_bwd(const float* dz, const void* a, int n, int h, int w, int c, int lda, const float* wgt,
                         int classes, void* da, int ldda, float* dw, float* db, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && c % 8 == 0 && c > 0 && c <= 64, "classes <= %d, channels <= 64", MAXK);
  const long long P = (long long)n * h * w;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<long long>((P + 127) / 128, 148 * 16);
  head3x3_dgrad_kernel<<<blocks, 128, classes * 9 * c * sizeof(float), st>>>(dz, n, h, w, c, wgt, classes, (bf16*)da, ldda);
  dim3 grid((int)std::min<long long>((P + 127) / 128, 148 * 2), 9);
  head3x3_wgrad_kernel<<<grid, 128, 0, st>>>(dz, (const bf16*)a, n, h, w, c, lda, classes, dw, db);
  return check_launch("head3x3 bwd");
}

}  // extern "C"
