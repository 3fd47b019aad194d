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

// z[p][k] = b[k] + sum_c w[k][c] * a[p][c] ; one warp per pixel
__global__ void head_fwd_kernel(const bf16* a, long long P, int C, int lda, const float* w, const float* b, int K,
                                float* z) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < P; p += nwarps) {
    float acc[MAXK] = {};
    for (int c = lane * 8; c < C; c += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(a + p * lda + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
          if (k < K) acc[k] = fmaf(f.x, w[k * C + c + 2 * i], fmaf(f.y, w[k * C + c + 2 * i + 1], acc[k]));
      }
    }
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
      if (k < K) {
        const float t = warp_sum(acc[k]);
        if (lane == 0) z[p * K + k] = t + b[k];
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
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(loss_sum, (double)lsum);
    int* c = counts + ((size_t)n * K + k) * 3;
    if (tp) atomicAdd(c, tp);
    if (fp) atomicAdd(c + 1, fp);
    if (fn) atomicAdd(c + 2, fn);
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

// da[p][c] = sum_k dz[p][k] w[k][c] (bf16) ; dw[k][c] += sum_p dz[p][k] a[p][c] ; db[k] += sum_p dz[p][k]
__global__ void head_bwd_kernel(const float* dz, const bf16* a, long long P, int C, int lda, const float* w, int K,
                                bf16* da, int ldda, float* dw, float* db) {
  const int G = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y;
  float acc[MAXK][8] = {};
  float accb[MAXK] = {};
  float wk[MAXK][8] = {};
  if (tx < G)
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) wk[k][i] = w[k * C + tx * 8 + i];
  if (tx < G)
    for (long long p = (long long)blockIdx.x * blockDim.y + ty; p < P; p += (long long)gridDim.x * blockDim.y) {
      const uint4 u = *reinterpret_cast<const uint4*>(a + p * lda + tx * 8);
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
      float av[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(hh[i]); av[2 * i] = f.x; av[2 * i + 1] = f.y; }
      float o[8] = {};
#pragma unroll
      for (int k = 0; k < MAXK; ++k) {
        if (k < K) {
          const float d = dz[p * K + k];
          if (tx == 0) accb[k] += d;
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc[k][i] = fmaf(d, av[i], acc[k][i]); o[i] = fmaf(d, wk[k][i], o[i]); }
        }
      }
      uint4 ou;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ou);
#pragma unroll
      for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
      *reinterpret_cast<uint4*>(da + p * ldda + tx * 8) = ou;
    }
  __shared__ float s[MAXK][2048];
  __shared__ float sb[MAXK];
  for (int i = ty * blockDim.x + tx; i < MAXK * 2048; i += blockDim.x * blockDim.y) (&s[0][0])[i] = 0.f;
  if (tx == 0 && ty == 0) for (int k = 0; k < MAXK; ++k) sb[k] = 0.f;
  __syncthreads();
  if (tx < G)
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s[k][tx * 8 + i], acc[k][i]);
      if (tx == 0) atomicAdd(&sb[k], accb[k]);
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
  const int blocks = (int)std::min<long long>((pixels * 32 + 255) / 256, 148 * 16);
  head_fwd_kernel<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, pixels, c, lda, w, bias, classes, z);
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
  const int blocks = (int)std::min<long long>((total + 127) / 128, 148 * 32);
  loss_bwd_kernel<<<std::max(blocks, 1), 128, 0, (cudaStream_t)stream>>>(z, n, h, w, classes, target, H, W, grad_scale, dz);
  return check_launch("seg_loss_bwd");
}

/* da bf16 [pixels][ldda] = dz . w ; dw fp32 [classes][c] += dz^T a ; db fp32 [classes] += sum dz */
int aadg_seg_head_bwd(const float* dz, const void* a, long long pixels, int c, int lda, const float* w, int classes,
                      void* da, int ldda, float* dw, float* db, void* stream) {
  AADG_REQUIRE(classes >= 1 && classes <= MAXK && c % 8 == 0 && c > 0 && c <= 2048, "bad sizes");
  int tx = 1;
  while (tx < (c >> 3)) tx <<= 1;
  tx = std::min(tx, 256);
  dim3 blk(tx, 256 / tx);
  const int blocks = (int)std::min<long long>((pixels + blk.y * 8 - 1) / (blk.y * 8), 148 * 4);
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
