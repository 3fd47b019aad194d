// Float tensor augmentation bank for sm_100a: the 19 operations of the reference's data/operations.py /
// data/functional.py (a copy of Faster-AutoAugment's differentiable bank; dead code there, named by the
// north star) on float32 [B,3,H,W] images in [0,1], forward values only:
//     out = clamp( mask_b * op(x, mag_b) + (1 - mask_b) * x , 0, 1 )        (operations.py:73-100)
// One launch per op for the whole batch (the reference spends ~15 ATen launches per op); statistics ops
// (Contrast mean, AutoContrast min/max, Equalize histogram) add one reduction pass + a 256-entry table.
// HBM-bound: 12 B read + 12 B written per pixel, 16-byte accesses.
#include <algorithm>

#include "common.cuh"

namespace aadg {
namespace f32 {

enum Op {
  SHEAR_X = 0, SHEAR_Y, TRANSLATE_X, TRANSLATE_Y, HFLIP, VFLIP, ROTATE, INVERT, SOLARIZE, POSTERIZE, GRAY, CONTRAST,
  AUTO_CONTRAST, SATURATE, BRIGHTNESS, HUE, SAMPLE_PAIRING, EQUALIZE, SHARPNESS, OP_COUNT
};

struct PlaneStat {          // one per (sample, channel)
  int lo, hi;               // ordered-int min / max of clamp(x)*255
  unsigned int hist[256];   // histogram of (int)(clamp(x)*255)
  float lut[256];           // AutoContrast / Equalize table (already divided by 255)
};
struct SampleStat { double gray_sum; double pad_; };

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
// functional.py:85-87 (0.110 for blue, sic), evaluated left to right without contraction
__device__ __forceinline__ float gray_of(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.110f, b));
}

__global__ void init_stats_kernel(PlaneStat* ps, SampleStat* ss, int planes, int samples) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < planes * 256) ps[i / 256].hist[i % 256] = 0;
  if (i < planes) { ps[i].lo = 0x7fffffff; ps[i].hi = (int)0x80000000; }
  if (i < samples) ss[i].gray_sum = 0.0;
}

// grid (chunks, B): min/max/histogram per plane and the grey sum per sample
__global__ void stats_kernel(const float* x, int HW, PlaneStat* ps, SampleStat* ss, int want_hist) {
  const int b = blockIdx.y;
  __shared__ unsigned int sh[3][256];
  __shared__ double sred[8];
  if (want_hist) for (int i = threadIdx.x; i < 768; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const float* xb = x + (size_t)b * 3 * HW;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  double gs = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float r = xb[i], g = xb[HW + i], bl = xb[2 * HW + i];
    gs += (double)gray_of(__fmul_rn(r, 255.f), __fmul_rn(g, 255.f), __fmul_rn(bl, 255.f));
    const float v[3] = {__fmul_rn(clamp01(r), 255.f), __fmul_rn(clamp01(g), 255.f), __fmul_rn(clamp01(bl), 255.f)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = fminf(lo[c], v[c]); hi[c] = fmaxf(hi[c], v[c]);
      if (want_hist) atomicAdd(&sh[c][(int)v[c]], 1u);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    lo[c] = -warp_max(-lo[c]); hi[c] = warp_max(hi[c]);
    if ((threadIdx.x & 31) == 0 && lo[c] <= hi[c]) {
      atomicMin(&ps[b * 3 + c].lo, f2ord(lo[c]));
      atomicMax(&ps[b * 3 + c].hi, f2ord(hi[c]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = gs;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sred[i];
    atomicAdd(&ss[b].gray_sum, t);
  }
  if (want_hist)
    for (int i = threadIdx.x; i < 768; i += blockDim.x)
      if ((&sh[0][0])[i]) atomicAdd(&ps[b * 3 + i / 256].hist[i % 256], (&sh[0][0])[i]);
}

// one CTA of 256 threads per plane
__global__ void lut_kernel(PlaneStat* ps, int op) {
  PlaneStat& p = ps[blockIdx.x];
  const int i = threadIdx.x;
  if (op == AUTO_CONTRAST) {
    // functional.py:196-207: floor((i - min) * (255 / (max - min + 0.1))) / 255
    const float mn = ord2f(p.lo), mx = ord2f(p.hi);
    const float scale = __fdiv_rn(255.f, __fadd_rn(__fsub_rn(mx, mn), 0.1f));
    p.lut[i] = __fdiv_rn(floorf(__fmul_rn(__fsub_rn((float)i, mn), scale)), 255.f);
    return;
  }
  // functional.py:242-262 (Pillow's equalize in float arithmetic)
  __shared__ float cdf[256];
  cdf[i] = (float)p.hist[i];
  __syncthreads();
  if (i == 0) { float run = 0.f; for (int k = 0; k < 256; ++k) { run = __fadd_rn(run, cdf[k]); cdf[k] = run; } }
  __syncthreads();
  const float step = floorf(__fdiv_rn(__fsub_rn(cdf[255], (float)p.hist[255]), 255.f));
  const float ex = (i == 0 ? 0.f : cdf[i - 1]) + floorf(__fdiv_rn(step, 2.f));
  p.lut[i] = __fdiv_rn(floorf(__fdiv_rn(ex, __fadd_rn(step, 0.1f))), 255.f);
}

struct OpArgs {
  const float* x; float* out;
  const float* mag; const float* mask; const int* perm;
  const PlaneStat* ps; const SampleStat* ss;
  int B, H, W, op;
};

__device__ __forceinline__ void rgb2hsv(float r, float g, float b, float& h, float& s, float& v) {
  const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b)), d = mx - mn;
  v = mx;
  s = mx > 0.f ? d / mx : 0.f;
  if (d <= 0.f) { h = 0.f; return; }
  float hh;
  if (mx == r) { hh = (g - b) / d; hh = hh - 6.f * floorf(hh / 6.f); }
  else if (mx == g) hh = (b - r) / d + 2.f;
  else hh = (r - g) / d + 4.f;
  hh = hh / 6.f;
  h = hh - floorf(hh);
}
__device__ __forceinline__ void hsv2rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float h6 = h * 6.f;
  const float fi = floorf(h6);
  const float f = h6 - fi;
  const float p = v * (1.f - s), q = v * (1.f - f * s), t = v * (1.f - (1.f - f) * s);
  int i = (int)fi % 6;
  if (i < 0) i += 6;
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// value of op(x)[b, :, y, xx] for one pixel
__device__ __forceinline__ void eval_op(const OpArgs& a, int b, int y, int xx, float m, float& r, float& g, float& bl) {
  const int H = a.H, W = a.W;
  const size_t HW = (size_t)H * W;
  const float* xb = a.x + (size_t)b * 3 * HW;
  const size_t i = (size_t)y * W + xx;
  switch (a.op) {
    case HFLIP: { const size_t j = (size_t)y * W + (W - 1 - xx); r = xb[j]; g = xb[HW + j]; bl = xb[2 * HW + j]; return; }
    case VFLIP: { const size_t j = (size_t)(H - 1 - y) * W + xx; r = xb[j]; g = xb[HW + j]; bl = xb[2 * HW + j]; return; }
    case SHEAR_X: case SHEAR_Y: case TRANSLATE_X: case TRANSLATE_Y: case ROTATE: {
      // forward pixel-space matrix M (dst = M src), sampled at M^-1 p with zero padding
      double m00 = 1, m01 = 0, m02 = 0, m10 = 0, m11 = 1, m12 = 0;
      const double mg = (double)m;
      if (a.op == SHEAR_X) m01 = mg;
      else if (a.op == SHEAR_Y) m10 = mg;
      else if (a.op == TRANSLATE_X) m02 = mg * W;
      else if (a.op == TRANSLATE_Y) m12 = mg * H;
      else {
        const double ang = mg * 0.017453292519943295, c = cos(ang), s = sin(ang);
        const double cx = (W - 1) * 0.5, cy = (H - 1) * 0.5;
        m00 = c; m01 = s; m02 = (1 - c) * cx - s * cy;
        m10 = -s; m11 = c; m12 = s * cx + (1 - c) * cy;
      }
      const double det = m00 * m11 - m01 * m10;
      const double i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
      const double i02 = -(i00 * m02 + i01 * m12), i12 = -(i10 * m02 + i11 * m12);
      const float sx = (float)(i00 * xx + i01 * y + i02), sy = (float)(i10 * xx + i11 * y + i12);
      const float fx0 = floorf(sx), fy0 = floorf(sy);
      const int x0 = (int)fx0, y0 = (int)fy0;
      const float fx = sx - fx0, fy = sy - fy0;
      r = g = bl = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int xi = x0 + dx, yi = y0 + dy;
          if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
          const float wgt = (dy ? fy : 1.f - fy) * (dx ? fx : 1.f - fx);
          const size_t j = (size_t)yi * W + xi;
          r += wgt * xb[j]; g += wgt * xb[HW + j]; bl += wgt * xb[2 * HW + j];
        }
      return;
    }
    default: break;
  }
  const float xr = xb[i], xg = xb[HW + i], xbv = xb[2 * HW + i];
  switch (a.op) {
    case INVERT: r = 1.f - xr; g = 1.f - xg; bl = 1.f - xbv; break;
    case SOLARIZE: r = xr < m ? xr : 1.f - xr; g = xg < m ? xg : 1.f - xg; bl = xbv < m ? xbv : 1.f - xbv; break;
    case POSTERIZE:
      r = __fdiv_rn((float)(long long)__fmul_rn(xr, 255.f), 255.f);
      g = __fdiv_rn((float)(long long)__fmul_rn(xg, 255.f), 255.f);
      bl = __fdiv_rn((float)(long long)__fmul_rn(xbv, 255.f), 255.f);
      break;
    case GRAY: r = g = bl = gray_of(xr, xg, xbv); break;
    case CONTRAST: {
      const float mean = __fdiv_rn(floorf((float)(a.ss[b].gray_sum / (double)HW) + 0.5f), 255.f);
      const float al = __fsub_rn(1.f, m);
      r = clamp01(__fadd_rn(mean, __fmul_rn(al, __fsub_rn(xr, mean))));
      g = clamp01(__fadd_rn(mean, __fmul_rn(al, __fsub_rn(xg, mean))));
      bl = clamp01(__fadd_rn(mean, __fmul_rn(al, __fsub_rn(xbv, mean))));
      break;
    }
    case AUTO_CONTRAST: case EQUALIZE: {
      const PlaneStat* p = a.ps + b * 3;
      r = p[0].lut[(int)__fmul_rn(clamp01(xr), 255.f)];
      g = p[1].lut[(int)__fmul_rn(clamp01(xg), 255.f)];
      bl = p[2].lut[(int)__fmul_rn(clamp01(xbv), 255.f)];
      break;
    }
    case SATURATE: {
      const float gr = gray_of(xr, xg, xbv), al = __fsub_rn(1.f, m);
      r = clamp01(__fadd_rn(gr, __fmul_rn(al, __fsub_rn(xr, gr))));
      g = clamp01(__fadd_rn(gr, __fmul_rn(al, __fsub_rn(xg, gr))));
      bl = clamp01(__fadd_rn(gr, __fmul_rn(al, __fsub_rn(xbv, gr))));
      break;
    }
    case BRIGHTNESS: {
      const float al = __fsub_rn(1.f, m);
      r = clamp01(__fmul_rn(al, xr)); g = clamp01(__fmul_rn(al, xg)); bl = clamp01(__fmul_rn(al, xbv));
      break;
    }
    case HUE: {
      float h, s, v;
      rgb2hsv(xr, xg, xbv, h, s, v);
      h = h + m;
      h = h - floorf(h);
      hsv2rgb(h, s, v, r, g, bl);
      break;
    }
    case SAMPLE_PAIRING: {
      const float* xo = a.x + (size_t)a.perm[b] * 3 * HW;
      const float om = __fsub_rn(1.f, m);
      r = __fadd_rn(__fmul_rn(om, xr), __fmul_rn(m, xo[i]));
      g = __fadd_rn(__fmul_rn(om, xg), __fmul_rn(m, xo[HW + i]));
      bl = __fadd_rn(__fmul_rn(om, xbv), __fmul_rn(m, xo[2 * HW + i]));
      break;
    }
    case SHARPNESS: {
      // functional.py:98-106,266-271: reflect-pad 3x3 [[1,1,1],[1,5,1],[1,1,1]]/13, blend(img, blur, 1 - mag)
      const float k1 = __fdiv_rn(1.f, 13.f), k5 = __fdiv_rn(5.f, 13.f);
      float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const size_t j = (size_t)reflect(y + dy, H) * W + reflect(xx + dx, W);
          const float k = (dy == 0 && dx == 0) ? k5 : k1;
          acc[0] = __fadd_rn(acc[0], __fmul_rn(k, xb[j]));
          acc[1] = __fadd_rn(acc[1], __fmul_rn(k, xb[HW + j]));
          acc[2] = __fadd_rn(acc[2], __fmul_rn(k, xb[2 * HW + j]));
        }
      const float al = __fsub_rn(1.f, m);
      r = clamp01(__fadd_rn(acc[0], __fmul_rn(al, __fsub_rn(xr, acc[0]))));
      g = clamp01(__fadd_rn(acc[1], __fmul_rn(al, __fsub_rn(xg, acc[1]))));
      bl = clamp01(__fadd_rn(acc[2], __fmul_rn(al, __fsub_rn(xbv, acc[2]))));
      break;
    }
    default: r = xr; g = xg; bl = xbv;
  }
}

// grid (x chunks, H, B); 4 consecutive pixels per thread, float4 stores per plane
__global__ void __launch_bounds__(128) op_kernel(const OpArgs a) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int W = a.W;
  const size_t HW = (size_t)a.H * W;
  const float m = a.mag ? a.mag[b] : 0.f;
  const float mk = a.mask ? a.mask[b] : 1.f;
  const float omk = __fsub_rn(1.f, mk);
  const float* xb = a.x + (size_t)b * 3 * HW + (size_t)y * W;
  float* ob = a.out + (size_t)b * 3 * HW + (size_t)y * W;
  for (int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; x0 < W; x0 += gridDim.x * blockDim.x * 4) {
    float o[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = x0 + k;
      if (xx >= W) { o[0][k] = o[1][k] = o[2][k] = 0.f; continue; }
      float r, g, bl;
      eval_op(a, b, y, xx, m, r, g, bl);
      // tensor_function clamps the op's output, _Operation.forward mixes and clamps again
      r = clamp01(r); g = clamp01(g); bl = clamp01(bl);
      o[0][k] = clamp01(__fadd_rn(__fmul_rn(mk, r), __fmul_rn(omk, xb[xx])));
      o[1][k] = clamp01(__fadd_rn(__fmul_rn(mk, g), __fmul_rn(omk, xb[HW + xx])));
      o[2][k] = clamp01(__fadd_rn(__fmul_rn(mk, bl), __fmul_rn(omk, xb[2 * HW + xx])));
    }
    if (x0 + 3 < W && (W & 3) == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) __stcs((float4*)(ob + c * HW + x0), make_float4(o[c][0], o[c][1], o[c][2], o[c][3]));
    } else {
      for (int k = 0; k < 4 && x0 + k < W; ++k)
        for (int c = 0; c < 3; ++c) ob[c * HW + x0 + k] = o[c][k];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: gradients of  out = clamp(mk * clamp(op(x, m)) + (1 - mk) * x)  with respect to x, the per-sample
// magnitude m_b and the per-sample mask mk_b, as torch autograd computes them through the reference's code
// (operations.py:73-100, functional.py): clamp passes the gradient where lo <= v <= hi (inclusive), the straight
// -through estimators of functional.py:21-46 send the gradient of Solarize / Posterize to the MAGNITUDE only (summed)
// and of AutoContrast / Equalize straight to the image, Contrast's rounded mean and the look-up tables carry none.
// ---------------------------------------------------------------------------------------------------------------
struct BwdArgs {
  const float* x; const float* go; float* gx;
  const float* mag; const float* mask; const int* perm;
  const PlaneStat* ps; const SampleStat* ss;
  float* gmag; float* gmask;
  int B, H, W, op, scatter;
};

__device__ __forceinline__ float gate01(float v) { return (v >= 0.f && v <= 1.f) ? 1.f : 0.f; }

// forward-mode duals over the directions (x_r, x_g, x_b, m): used for Hue, whose Jacobian goes through HSV
struct Dual { float v; float d[4]; };
__device__ __forceinline__ Dual dconst(float v) { Dual r; r.v = v; r.d[0] = r.d[1] = r.d[2] = r.d[3] = 0.f; return r; }
__device__ __forceinline__ Dual dvar(float v, int k) { Dual r = dconst(v); r.d[k] = 1.f; return r; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { a.v += b.v; for (int i = 0; i < 4; ++i) a.d[i] += b.d[i]; return a; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { a.v -= b.v; for (int i = 0; i < 4; ++i) a.d[i] -= b.d[i]; return a; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) {
  Dual r; r.v = a.v * b.v;
  for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  Dual r; r.v = a.v / b.v;
  for (int i = 0; i < 4; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
__device__ __forceinline__ Dual dshift(Dual a, float c) { a.v += c; return a; }      // + constant (floor terms)
__device__ __forceinline__ void hue_dual(float xr, float xg, float xb, float m, Dual out[3]) {
  const Dual r = dvar(xr, 0), g = dvar(xg, 1), b = dvar(xb, 2), mm = dvar(m, 3);
  const Dual mx = (xr >= xg && xr >= xb) ? r : (xg >= xb ? g : b);
  const Dual mn = (xr <= xg && xr <= xb) ? r : (xg <= xb ? g : b);
  const Dual d = mx - mn;
  const Dual v = mx;
  const Dual s = mx.v > 0.f ? d / mx : dconst(0.f);
  Dual h = dconst(0.f);
  if (d.v > 0.f) {
    Dual hh;
    if (mx.v == xr) { hh = (g - b) / d; hh = dshift(hh, -6.f * floorf(hh.v / 6.f)); }
    else if (mx.v == xg) hh = dshift((b - r) / d, 2.f);
    else hh = dshift((r - g) / d, 4.f);
    hh = hh * dconst(1.f / 6.f);
    h = dshift(hh, -floorf(hh.v));
  }
  h = h + mm;
  h = dshift(h, -floorf(h.v));
  const Dual h6 = h * dconst(6.f);
  const float fi = floorf(h6.v);
  const Dual f = dshift(h6, -fi);
  const Dual one = dconst(1.f);
  const Dual p = v * (one - s), q = v * (one - f * s), t = v * (one - (one - f) * s);
  int i = (int)fi % 6;
  if (i < 0) i += 6;
  switch (i) {
    case 0: out[0] = v; out[1] = t; out[2] = p; break;
    case 1: out[0] = q; out[1] = v; out[2] = p; break;
    case 2: out[0] = p; out[1] = v; out[2] = t; break;
    case 3: out[0] = p; out[1] = q; out[2] = v; break;
    case 4: out[0] = t; out[1] = p; out[2] = v; break;
    default: out[0] = v; out[1] = p; out[2] = q; break;
  }
}

// grid (x chunks, H, B), one pixel per thread-iteration
__global__ void __launch_bounds__(128) op_bwd_kernel(const BwdArgs a) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int H = a.H, W = a.W;
  const size_t HW = (size_t)H * W;
  const float m = a.mag ? a.mag[b] : 0.f;
  const float mk = a.mask ? a.mask[b] : 1.f;
  const float omk = __fsub_rn(1.f, mk);
  const float* xb = a.x + (size_t)b * 3 * HW;
  const float* gob = a.go + (size_t)b * 3 * HW;
  float* gxb = a.gx + (size_t)b * 3 * HW;
  OpArgs fa{};
  fa.x = a.x; fa.mag = a.mag; fa.mask = a.mask; fa.perm = a.perm; fa.ps = a.ps; fa.ss = a.ss; fa.B = a.B; fa.H = H; fa.W = W;
  fa.op = a.op;
  float acc_mag = 0.f, acc_mask = 0.f;
  for (int xx = blockIdx.x * blockDim.x + threadIdx.x; xx < W; xx += gridDim.x * blockDim.x) {
    const size_t i = (size_t)y * W + xx;
    const float xv[3] = {xb[i], xb[HW + i], xb[2 * HW + i]};
    float yv[3];
    eval_op(fa, b, y, xx, m, yv[0], yv[1], yv[2]);      // the forward value (blend ops: already clamped)
    float gy[3], gdir[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float yc = clamp01(yv[c]);
      const float pre = __fadd_rn(__fmul_rn(mk, yc), __fmul_rn(omk, xv[c]));
      const float g1 = gob[c * HW + i] * gate01(pre);
      acc_mask += g1 * (yc - xv[c]);
      gdir[c] = g1 * omk;
      gy[c] = g1 * mk * gate01(yv[c]);                     // tensor_function's clamp of the op output
    }
    float gxo[3] = {0.f, 0.f, 0.f};                        // the op's contribution to d/dx at THIS pixel
    switch (a.op) {
      case HFLIP: case VFLIP: {
        const size_t j = a.op == HFLIP ? (size_t)y * W + (W - 1 - xx) : (size_t)(H - 1 - y) * W + xx;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(gxb + c * HW + j, gy[c]);
        break;
      }
      case SHEAR_X: case SHEAR_Y: case TRANSLATE_X: case TRANSLATE_Y: case ROTATE: {
        double m00 = 1, m01 = 0, m02 = 0, m10 = 0, m11 = 1, m12 = 0;
        const double mg = (double)m;
        if (a.op == SHEAR_X) m01 = mg;
        else if (a.op == SHEAR_Y) m10 = mg;
        else if (a.op == TRANSLATE_X) m02 = mg * W;
        else if (a.op == TRANSLATE_Y) m12 = mg * H;
        else {
          const double ang = mg * 0.017453292519943295, c = cos(ang), s = sin(ang);
          const double cx = (W - 1) * 0.5, cy = (H - 1) * 0.5;
          m00 = c; m01 = s; m02 = (1 - c) * cx - s * cy;
          m10 = -s; m11 = c; m12 = s * cx + (1 - c) * cy;
        }
        const double det = m00 * m11 - m01 * m10;
        const double i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
        const double i02 = -(i00 * m02 + i01 * m12), i12 = -(i10 * m02 + i11 * m12);
        const float sx = (float)(i00 * xx + i01 * y + i02), sy = (float)(i10 * xx + i11 * y + i12);
        const float fx0 = floorf(sx), fy0 = floorf(sy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const float fx = sx - fx0, fy = sy - fy0;
        // d(source coordinate)/d(magnitude) of the INVERSE map
        float dsx = 0.f, dsy = 0.f;
        if (a.op == SHEAR_X) dsx = -(float)y;
        else if (a.op == SHEAR_Y) dsy = -(float)xx;
        else if (a.op == TRANSLATE_X) dsx = -(float)W;
        else if (a.op == TRANSLATE_Y) dsy = -(float)H;
        else { dsx = -(sy - (H - 1) * 0.5f) * 0.017453292519943295f; dsy = (sx - (W - 1) * 0.5f) * 0.017453292519943295f; }
        float v[3][2][2];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int xi = x0 + dx, yi = y0 + dy;
            const bool in = xi >= 0 && xi < W && yi >= 0 && yi < H;
            const size_t j = in ? (size_t)yi * W + xi : 0;
            const float wgt = (dy ? fy : 1.f - fy) * (dx ? fx : 1.f - fx);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              v[c][dy][dx] = in ? xb[c * HW + j] : 0.f;
              if (in && gy[c] != 0.f) atomicAdd(gxb + c * HW + j, wgt * gy[c]);
            }
          }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float dIdx = (1.f - fy) * (v[c][0][1] - v[c][0][0]) + fy * (v[c][1][1] - v[c][1][0]);
          const float dIdy = (1.f - fx) * (v[c][1][0] - v[c][0][0]) + fx * (v[c][1][1] - v[c][0][1]);
          acc_mag += gy[c] * (dIdx * dsx + dIdy * dsy);
        }
        break;
      }
      case INVERT: for (int c = 0; c < 3; ++c) gxo[c] = -gy[c]; break;
      case SOLARIZE: case POSTERIZE: acc_mag += gy[0] + gy[1] + gy[2]; break;                 // STE: magnitude only
      case AUTO_CONTRAST: case EQUALIZE: for (int c = 0; c < 3; ++c) gxo[c] = gy[c]; break;     // STE: straight to the image
      case GRAY: {
        const float t = gy[0] + gy[1] + gy[2];
        gxo[0] = 0.299f * t; gxo[1] = 0.587f * t; gxo[2] = 0.110f * t;
        break;
      }
      case CONTRAST: {
        const float mean = __fdiv_rn(floorf((float)(a.ss[b].gray_sum / (double)HW) + 0.5f), 255.f);
        const float al = __fsub_rn(1.f, m);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float t = __fadd_rn(mean, __fmul_rn(al, __fsub_rn(xv[c], mean)));
          const float gt = gy[c] * gate01(t);
          gxo[c] = gt * al;
          acc_mag -= gt * (xv[c] - mean);
        }
        break;
      }
      case SATURATE: {
        const float gr = gray_of(xv[0], xv[1], xv[2]), al = __fsub_rn(1.f, m);
        float gsum = 0.f, gt[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float t = __fadd_rn(gr, __fmul_rn(al, __fsub_rn(xv[c], gr)));
          gt[c] = gy[c] * gate01(t);
          gsum += gt[c];
          acc_mag -= gt[c] * (xv[c] - gr);
        }
        const float wc[3] = {0.299f, 0.587f, 0.110f};
#pragma unroll
        for (int c = 0; c < 3; ++c) gxo[c] = gt[c] * al + wc[c] * (1.f - al) * gsum;
        break;
      }
      case BRIGHTNESS: {
        const float al = __fsub_rn(1.f, m);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float gt = gy[c] * gate01(__fmul_rn(al, xv[c]));
          gxo[c] = gt * al;
          acc_mag -= gt * xv[c];
        }
        break;
      }
      case HUE: {
        Dual o[3];
        hue_dual(xv[0], xv[1], xv[2], m, o);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          gxo[0] += gy[c] * o[c].d[0]; gxo[1] += gy[c] * o[c].d[1]; gxo[2] += gy[c] * o[c].d[2];
          acc_mag += gy[c] * o[c].d[3];
        }
        break;
      }
      case SAMPLE_PAIRING: {
        const float* xo = a.x + (size_t)a.perm[b] * 3 * HW;
        float* gxp = a.gx + (size_t)a.perm[b] * 3 * HW;
        const float om = __fsub_rn(1.f, m);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          gxo[c] = gy[c] * om;
          atomicAdd(gxp + c * HW + i, gy[c] * m);
          acc_mag += gy[c] * (xo[c * HW + i] - xv[c]);
        }
        break;
      }
      case SHARPNESS: {
        const float k1 = __fdiv_rn(1.f, 13.f), k5 = __fdiv_rn(5.f, 13.f);
        const float al = __fsub_rn(1.f, m);
        float blur[3] = {0.f, 0.f, 0.f};
        size_t js[9];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const size_t j = (size_t)reflect(y + dy, H) * W + reflect(xx + dx, W);
            js[(dy + 1) * 3 + dx + 1] = j;
            const float k = (dy == 0 && dx == 0) ? k5 : k1;
#pragma unroll
            for (int c = 0; c < 3; ++c) blur[c] = __fadd_rn(blur[c], __fmul_rn(k, xb[c * HW + j]));
          }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float t = __fadd_rn(blur[c], __fmul_rn(al, __fsub_rn(xv[c], blur[c])));
          const float gt = gy[c] * gate01(t);
          gxo[c] = gt * al;
          acc_mag -= gt * (xv[c] - blur[c]);
          const float gb = gt * (1.f - al);
          if (gb != 0.f)
#pragma unroll
            for (int q = 0; q < 9; ++q) atomicAdd(gxb + c * HW + js[q], gb * (q == 4 ? k5 : k1));
        }
        break;
      }
      default: break;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = gdir[c] + gxo[c];
      if (a.scatter) { if (t != 0.f) atomicAdd(gxb + c * HW + i, t); }
      else gxb[c * HW + i] = t;
    }
  }
  acc_mag = warp_sum(acc_mag);
  acc_mask = warp_sum(acc_mask);
  if ((threadIdx.x & 31) == 0) {
    if (a.gmag && acc_mag != 0.f) atomicAdd(a.gmag + b, acc_mag);
    if (a.gmask && acc_mask != 0.f) atomicAdd(a.gmask + b, acc_mask);
  }
}

}  // namespace f32
}  // namespace aadg

using namespace aadg;
using namespace aadg::f32;

extern "C" {

size_t aadg_f32_workspace_bytes(int batch) {
  if (batch <= 0) return 0;
  return align_up(sizeof(PlaneStat) * 3 * (size_t)batch, 256) + align_up(sizeof(SampleStat) * (size_t)batch, 256);
}

/* out = clamp(mask*op(x, mag) + (1-mask)*x, 0, 1); x, out float32 [batch,3,h,w] (out != x); mag, mask float32
 * [batch] on the device (NULL: mag 0, mask 1); perm int32 [batch] (SamplePairing only); op = index in the
 * reference's data/operations.py __all__ (ShearX=0 ... Sharpness=18). */
int aadg_f32_op(int op, const float* x, int batch, int h, int w, const float* mag, const float* mask,
                const int32_t* perm, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  AADG_REQUIRE(op >= 0 && op < OP_COUNT, "unknown op %d", op);
  AADG_REQUIRE(batch > 0 && h > 0 && w > 0 && x && out && x != out, "bad arguments");
  AADG_REQUIRE(op != SAMPLE_PAIRING || perm, "SamplePairing needs a permutation");
  AADG_REQUIRE(h <= 65535 && batch <= 65535, "image too tall / batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  OpArgs a{};
  a.x = x; a.out = out; a.mag = mag; a.mask = mask; a.perm = perm; a.B = batch; a.H = h; a.W = w; a.op = op;
  if (op == CONTRAST || op == AUTO_CONTRAST || op == EQUALIZE) {
    if (!workspace || workspace_bytes < aadg_f32_workspace_bytes(batch)) {
      set_error("workspace too small");
      return AADG_ENOSPC;
    }
    PlaneStat* ps = (PlaneStat*)workspace;
    SampleStat* ss = (SampleStat*)((char*)workspace + align_up(sizeof(PlaneStat) * 3 * (size_t)batch, 256));
    const int planes = 3 * batch;
    init_stats_kernel<<<(planes * 256 + 255) / 256, 256, 0, st>>>(ps, ss, planes, batch);
    dim3 sg(std::max(1, std::min((h * w + 255) / 256, 64)), batch);
    stats_kernel<<<sg, 256, 0, st>>>(x, h * w, ps, ss, op == EQUALIZE);
    if (op != CONTRAST) lut_kernel<<<planes, 256, 0, st>>>(ps, op);
    a.ps = ps; a.ss = ss;
  }
  dim3 grid(std::max(1, std::min((w / 4 + 127) / 128, 8)), h, batch);
  op_kernel<<<grid, 128, 0, st>>>(a);
  return check_launch("f32 op kernel");
}

/* Backward of aadg_f32_op: given go = d(loss)/d(out) (float32 [batch,3,h,w]) writes gx = d/dx (same shape) and
 * ACCUMULATES nothing: gmag / gmask (float32 [batch], may be NULL) receive d/d(mag_b) and d/d(mask_b).  Gradient
 * semantics are torch autograd's through data/operations.py:73-100 and data/functional.py (inclusive clamp gates, the
 * straight-through estimators of functional.py:21-46).  Same workspace as the forward call. */
int aadg_f32_op_backward(int op, const float* x, const float* go, int batch, int h, int w, const float* mag,
                         const float* mask, const int32_t* perm, float* gx, float* gmag, float* gmask, void* workspace,
                         size_t workspace_bytes, void* stream) {
  AADG_REQUIRE(op >= 0 && op < OP_COUNT, "unknown op %d", op);
  AADG_REQUIRE(batch > 0 && h > 0 && w > 0 && x && go && gx && gx != x && gx != go, "bad arguments");
  AADG_REQUIRE(op != SAMPLE_PAIRING || perm, "SamplePairing needs a permutation");
  AADG_REQUIRE(h <= 65535 && batch <= 65535, "image too tall / batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  BwdArgs a{};
  a.x = x; a.go = go; a.gx = gx; a.mag = mag; a.mask = mask; a.perm = perm; a.gmag = gmag; a.gmask = gmask;
  a.B = batch; a.H = h; a.W = w; a.op = op;
  a.scatter = (op <= ROTATE && op != INVERT) || op == SAMPLE_PAIRING || op == SHARPNESS;
  if (op == CONTRAST || op == AUTO_CONTRAST || op == EQUALIZE) {
    if (!workspace || workspace_bytes < aadg_f32_workspace_bytes(batch)) {
      set_error("workspace too small");
      return AADG_ENOSPC;
    }
    PlaneStat* ps = (PlaneStat*)workspace;
    SampleStat* ss = (SampleStat*)((char*)workspace + align_up(sizeof(PlaneStat) * 3 * (size_t)batch, 256));
    const int planes = 3 * batch;
    init_stats_kernel<<<(planes * 256 + 255) / 256, 256, 0, st>>>(ps, ss, planes, batch);
    dim3 sg(std::max(1, std::min((h * w + 255) / 256, 64)), batch);
    stats_kernel<<<sg, 256, 0, st>>>(x, h * w, ps, ss, op == EQUALIZE);
    if (op != CONTRAST) lut_kernel<<<planes, 256, 0, st>>>(ps, op);
    a.ps = ps; a.ss = ss;
  }
  if (a.scatter) AADG_CUDA_TRY(cudaMemsetAsync(gx, 0, sizeof(float) * 3 * (size_t)batch * h * w, st));
  if (gmag) AADG_CUDA_TRY(cudaMemsetAsync(gmag, 0, sizeof(float) * batch, st));
  if (gmask) AADG_CUDA_TRY(cudaMemsetAsync(gmask, 0, sizeof(float) * batch, st));
  dim3 grid(std::max(1, std::min((w + 127) / 128, 8)), h, batch);
  op_bwd_kernel<<<grid, 128, 0, st>>>(a);
  return check_launch("f32 op backward kernel");
}

}  // extern "C"
