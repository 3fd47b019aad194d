// uint8 augmentation bank for sm_100a.
//
// Replaces, for whole batches on the device, what the reference does per image on the CPU through
// Pillow: data/basic.py:70-167 (ten live ops), :12-67,82 (geometric ops), data/policy.py:15-61
// (sub-policy application) and the Normalize_dg / ToTensor epilogue of data/transform.py:138-236.
//
// Design (HBM-bound byte work, no tensor cores):
//   * every output image is one row of the decision table; the host part of this file compiles the
//     row into a short "program" of steps (LUT / COLOR / CUTOUT / SHARP / AFFINE / FLIP);
//   * ops that are per-channel functions of the byte value (Invert, Solarize, Posterize,
//     Brightness, Contrast, AutoContrast, Equalize) become 3x256 look-up tables built on the device;
//     the three statistics ops take their histogram from a per-SOURCE histogram pass whenever the
//     ops before them are look-up tables too (the histogram is pushed through the tables instead
//     of re-reading the image), so the common case reads each source once for statistics;
//   * one tile kernel evaluates a whole program per pixel: image rows are staged in shared memory
//     with 1-D TMA bulk copies (cp.async.bulk -> mbarrier), the 3x3 SMOOTH stencil of Sharpness runs
//     on the staged tile, and the epilogue writes either uint8 HWC or the normalised float32 CHW
//     tensor the model consumes (x/127.5-1) with 16-byte stores.
//   * float arithmetic that must match Pillow bit for bit (blend, SMOOTH, autocontrast) uses
//     explicit round-to-nearest intrinsics so nothing is contracted into an FMA.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace aadg {
namespace u8 {

constexpr int TW = 128;          // tile width in pixels
constexpr int TH = 8;            // tile height in pixels
constexpr int NT = 256;          // threads per CTA: 32 threads x 4 pixels per tile row
constexpr int TILES_PER_CTA = 8;  // consecutive tiles walked by one CTA (amortises the table set-up)
constexpr int RS = 432;           // shared-memory bytes per staged row: (TW+2)*3 + 2x15 alignment slack, 16-B multiple
static_assert(RS % 16 == 0 && RS >= (TW + 2) * 3 + 30, "row stride must keep 16-byte alignment for bulk copies");

enum StepKind { K_LUT = 0, K_COLOR = 1, K_CUTOUT = 2, K_SHARP = 3, K_AFFINE = 4, K_FLIP = 5 };
enum LutKind { L_INVERT = 0, L_SOLARIZE, L_POSTERIZE, L_BRIGHT, L_CONTRAST, L_AUTOCONTRAST, L_EQUALIZE };

struct DevStep {
  int kind;
  int lut_kind;
  float f;      // blend factor
  int p[6];     // CUTOUT: x0,y0,x1,y1 | AFFINE: a0..a5 (16.16) | LUT: p[0]=param, p[1]=stat slot, p[2]=derive
};
struct DevRow {
  int src;
  int n_steps;
  DevStep s[AADG_MAX_OPS];
};
struct PassItem {
  int row;
  int s0, s1;      // steps [s0, s1) are evaluated
  int base;        // -1: source image rows[row].src; else scratch image index
  int out;         // output image index (scratch slot or final row) / statistics slot
  int sharp;       // index of the SHARP step inside [s0,s1) or -1
  int gather;      // 1 if [s0,s1) contains AFFINE/FLIP (then sharp == -1 or gathers precede it)
  int pad_;        // 1: every step in [s0,s1) is a look-up table (fast path)
};
struct Stat {      // one statistics slot
  unsigned int hist[3][256];
  unsigned long long luma_sum;
  unsigned long long pad_;
};

// ---- Pillow arithmetic ---------------------------------------------------------------------------
__device__ __forceinline__ int luma_u8(int r, int g, int b) {
  return (19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16;   // Convert.c rgb2l
}
// exact (float)v for |v| < 2^22 on the fp32 / integer pipes (the conversion unit runs at a fraction of their rate):
// 0x4B400000 is 1.5 * 2^23, whose ulp is 1, so adding v to its bit pattern adds v to its value
__device__ __forceinline__ float i2f_small(int v) {
  return __fsub_rn(__int_as_float(0x4B400000 + v), 12582912.0f);
}
// ImagingBlend: a + alpha*(b-a) in float32, mul then add (no FMA), truncated; clipped outside [0,1].  For alpha inside
// [0,1] the clip is the identity on uint8 inputs (|fl(alpha*(b-a))| <= |b-a| by monotonicity of rounding), so it is
// applied unconditionally: two min/max instead of a data-dependent branch.
__device__ __forceinline__ int blend_u8(int a, int b, float alpha, bool /*inside01*/) {
  float t = __fadd_rn(i2f_small(a), __fmul_rn(alpha, i2f_small(b - a)));
  t = fminf(fmaxf(t, 0.f), 255.f);
  return (int)t;
}

struct Ctx {
  const DevRow* row;
  const uint8_t* luts;   // this row's tables [MAX_OPS][3][256]
};

// pointwise steps [s0,s1) on one pixel at coordinate (x,y) of the level it lives on
__device__ __forceinline__ void apply_point(const DevStep& st, const uint8_t* lut, int x, int y,
                                            int& r, int& g, int& b) {
  if (st.kind == K_LUT) {
    r = lut[r]; g = lut[256 + g]; b = lut[512 + b];
  } else if (st.kind == K_COLOR) {
    const bool in01 = st.f >= 0.f && st.f <= 1.f;
    const int l = luma_u8(r, g, b);
    r = blend_u8(l, r, st.f, in01); g = blend_u8(l, g, st.f, in01); b = blend_u8(l, b, st.f, in01);
  } else if (st.kind == K_CUTOUT) {
    if (x >= st.p[0] && x <= st.p[2] && y >= st.p[1] && y <= st.p[3]) r = g = b = 127;
  }
}

// steps [k0,k1) on ONE pixel packed as r | g << 8 | b << 16, OUT OF LINE: the stencil kernel applies the steps around a
// Sharpness from ten places inside a three-way unrolled row loop, and inlining the step interpreter there (times the
// compiler's own unrolling of the step loop) grew the kernel to 175 KB of code -- it then stalled on instruction fetch
// more than on anything else (ncu: stall_no_instruction 5.8 warps per issue, profiles/r02_ncu_aug_kernels.csv)
__device__ __noinline__ uint32_t apply_steps_rgb(const DevRow* row, const uint8_t* luts, int k0, int k1, int x, int y,
                                                 uint32_t rgb) {
  int r = rgb & 255, g = (rgb >> 8) & 255, b = (rgb >> 16) & 255;
#pragma unroll 1
  for (int k = k0; k < k1; ++k) apply_point(row->s[k], luts + k * 768, x, y, r, g, b);
  return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// the same for a QUAD (four consecutive pixels of one image row = three 32-bit words, bytes r0 g0 b0 r1 | g1 b1 r2 g2 |
// b2 r3 g3 b3): one call and one decode of every step per four pixels
struct Quad { uint32_t w0, w1, w2; };
__device__ __noinline__ Quad apply_steps_quad(const DevRow* row, const uint8_t* luts, int k0, int k1, int x0, int y, Quad q) {
  int r[4], g[4], b[4];
  r[0] = q.w0 & 255; g[0] = (q.w0 >> 8) & 255; b[0] = (q.w0 >> 16) & 255; r[1] = q.w0 >> 24;
  g[1] = q.w1 & 255; b[1] = (q.w1 >> 8) & 255; r[2] = (q.w1 >> 16) & 255; g[2] = q.w1 >> 24;
  b[2] = q.w2 & 255; r[3] = (q.w2 >> 8) & 255; g[3] = (q.w2 >> 16) & 255; b[3] = q.w2 >> 24;
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    const DevStep& st = row->s[k];
    if (st.kind == K_LUT) {
      const uint8_t* lut = luts + k * 768;
#pragma unroll
      for (int i = 0; i < 4; ++i) { r[i] = lut[r[i]]; g[i] = lut[256 + g[i]]; b[i] = lut[512 + b[i]]; }
    } else if (st.kind == K_COLOR) {
      const float f = st.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int l = luma_u8(r[i], g[i], b[i]);
        r[i] = blend_u8(l, r[i], f, false); g[i] = blend_u8(l, g[i], f, false); b[i] = blend_u8(l, b[i], f, false);
      }
    } else if (st.kind == K_CUTOUT) {
      if (y >= st.p[1] && y <= st.p[3]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x0 + i >= st.p[0] && x0 + i <= st.p[2]) r[i] = g[i] = b[i] = 127;
      }
    }
  }
  Quad o;
  o.w0 = r[0] | (g[0] << 8) | (b[0] << 16) | (r[1] << 24);
  o.w1 = g[1] | (b[1] << 8) | (r[2] << 16) | (g[2] << 24);
  o.w2 = b[2] | (r[3] << 8) | (g[3] << 16) | (b[3] << 24);
  return o;
}

// Pull evaluation of steps [s0,s1) (no SHARP inside) at output pixel (x,y): walks the gathers
// backwards to find the source pixel, then applies the pointwise steps forwards.
__device__ __forceinline__ void eval_gather(const DevRow& row, const uint8_t* luts, int s0, int s1,
                                            const uint8_t* base, int W, int H, int x, int y,
                                            int& r, int& g, int& b) {
  int cx[AADG_MAX_OPS + 1], cy[AADG_MAX_OPS + 1];
  cx[s1 - s0] = x; cy[s1 - s0] = y;
  int dead = -1;   // highest step index whose gather fell outside (value is 0 after that step)
#pragma unroll
  for (int k = AADG_MAX_OPS - 1; k >= 0; --k) {
    if (k >= s1 - s0) continue;
    const DevStep& st = row.s[s0 + k];
    int px = cx[k + 1], py = cy[k + 1];
    if (dead < 0) {
      if (st.kind == K_AFFINE) {
        // Geometry.c affine_fixed + nearest: xx = a2 + y*a1 + x*a0 (16.16), sample [yy>>16][xx>>16]
        long long xx = (long long)st.p[2] + (long long)py * st.p[1] + (long long)px * st.p[0];
        long long yy = (long long)st.p[5] + (long long)py * st.p[4] + (long long)px * st.p[3];
        long long xi = xx >> 16, yi = yy >> 16;
        if (xi < 0 || xi >= W || yi < 0 || yi >= H) { dead = k; xi = 0; yi = 0; }
        px = (int)xi; py = (int)yi;
      } else if (st.kind == K_FLIP) {
        px = W - 1 - px;
      }
    }
    cx[k] = px; cy[k] = py;
  }
  if (dead < 0) {
    const uint8_t* p = base + ((size_t)cy[0] * W + cx[0]) * 3;
    r = p[0]; g = p[1]; b = p[2];
  } else {
    r = g = b = 0;
  }
#pragma unroll
  for (int k = 0; k < AADG_MAX_OPS; ++k) {
    if (k >= s1 - s0 || k <= dead) continue;
    apply_point(row.s[s0 + k], luts + (size_t)(s0 + k) * 768, cx[k + 1], cy[k + 1], r, g, b);
  }
}

// ImageFilter.SMOOTH through ImagingFilter3x3 (Filter.c): float32, starts at 0.5, rows y+1, y, y-1,
// each row (a*k0 + b*k1) + c*k2, then clip8 by truncation.
__device__ __forceinline__ int smooth_1ch(const uint8_t* up, const uint8_t* mid, const uint8_t* dn,
                                          float k1, float k5) {
  // up = row y-1, mid = row y, dn = row y+1; pointers at the centre pixel's channel byte
  float ss = 0.5f;
  float t = __fadd_rn(__fadd_rn(__fmul_rn((float)dn[-3], k1), __fmul_rn((float)dn[0], k1)),
                      __fmul_rn((float)dn[3], k1));
  ss = __fadd_rn(ss, t);
  t = __fadd_rn(__fadd_rn(__fmul_rn((float)mid[-3], k1), __fmul_rn((float)mid[0], k5)),
                __fmul_rn((float)mid[3], k1));
  ss = __fadd_rn(ss, t);
  t = __fadd_rn(__fadd_rn(__fmul_rn((float)up[-3], k1), __fmul_rn((float)up[0], k1)),
                __fmul_rn((float)up[3], k1));
  ss = __fadd_rn(ss, t);
  return ss <= 0.f ? 0 : (ss >= 255.f ? 255 : (int)ss);
}

enum Mode { MODE_STATS = 0, MODE_U8 = 1, MODE_F32 = 2 };

struct PassArgs {
  const DevRow* rows;
  const PassItem* items;
  const uint8_t* luts;        // [n_rows][MAX_OPS][768]
  const uint8_t* src;         // [n_src][H][W][3]
  const uint8_t* scratch;     // [*][H][W][3]
  uint8_t* out_u8;            // MODE_U8: [*][H][W][3]
  float* out_f32;             // MODE_F32: [*][3][H][W]
  Stat* stats;                // MODE_STATS
  int H, W;
  int aligned;                // rows can be staged with 16-byte bulk copies
};

template <int MODE>
__global__ void __launch_bounds__(NT) pass_kernel(const PassArgs a) {
  __shared__ __align__(128) uint8_t tile[(TH + 2) * RS];
  __shared__ __align__(16) uint8_t s_luts[AADG_MAX_OPS * 768];
  __shared__ float s_norm[256];
  __shared__ __align__(16) uint8_t s_comp[768];                    // composition of an all-LUT program
  __shared__ float s_fcomp[MODE == MODE_F32 ? 768 : 1];           // ... already normalised (MODE_F32)
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned int s_hist[MODE == MODE_STATS ? (NT / 32) * 768 : 1];
  __shared__ DevRow s_row;

  const PassItem it = a.items[blockIdx.y];
  const int W = a.W, H = a.H;
  const int tiles_x = (W + TW - 1) / TW;
  const int n_tiles = tiles_x * ((H + TH - 1) / TH);
  const int tid = threadIdx.x;

  if (tid < (int)(sizeof(DevRow) / 4)) ((int*)&s_row)[tid] = ((const int*)&a.rows[it.row])[tid];
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  // per-CTA set-up, amortised over TILES_PER_CTA tiles: tables, normalisation map, histograms
  {
    const uint4* g = (const uint4*)(a.luts + (size_t)it.row * AADG_MAX_OPS * 768);
    uint4* s = (uint4*)s_luts;
    for (int i = tid; i < AADG_MAX_OPS * 768 / 16; i += NT) s[i] = g[i];
    if (MODE == MODE_F32) s_norm[tid] = __fsub_rn(__fdiv_rn((float)tid, 127.5f), 1.0f);
    if (MODE == MODE_STATS)
      for (int i = tid; i < (NT / 32) * 768; i += NT) s_hist[i] = 0;
  }
  __syncthreads();
  const DevRow& row = s_row;
  // fast path: a program made only of look-up tables collapses into ONE table per channel (7 of the 10
  // searchable ops are tables), so the per-pixel work is three shared-memory reads
  const bool fast = it.pad_ != 0;
  if (fast) {
    for (int i = tid; i < 768; i += NT) {
      const int c = i >> 8;
      int v = i & 255;
      for (int k = it.s0; k < it.s1; ++k) v = s_luts[k * 768 + c * 256 + v];
      s_comp[i] = (uint8_t)v;
      if (MODE == MODE_F32) s_fcomp[i] = __fsub_rn(__fdiv_rn((float)v, 127.5f), 1.0f);
    }
    __syncthreads();
  }
  const uint8_t* base = it.base < 0 ? a.src + (size_t)row.src * H * W * 3
                                    : a.scratch + (size_t)it.base * H * W * 3;
  const int halo = it.sharp >= 0 ? 1 : 0;
  unsigned int lsum = 0;
  uint32_t bar_phase = 0;
  const int t_end = min(n_tiles, (int)(blockIdx.x + 1) * TILES_PER_CTA);
  for (int t = blockIdx.x * TILES_PER_CTA; t < t_end; ++t) {
  const int x0 = (t % tiles_x) * TW;
  const int y0 = (t / tiles_x) * TH;
  const int ya = max(y0 - halo, 0), yb = min(y0 + TH + halo, H);   // staged image rows [ya,yb)
  const int bx0 = max((x0 - halo) * 3, 0), bx1 = min((x0 + TW + halo) * 3, W * 3);
  int sbase;   // byte column of the image row that sits at tile row offset 0

  if (!it.gather) {
    if (a.aligned) {
      sbase = bx0 & ~15;
      const int send = min(W * 3, (bx1 + 15) & ~15);
      if (tid == 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)((send - sbase) * (yb - ya)));
        for (int yy = ya; yy < yb; ++yy)
          tma_bulk_g2s(tile + (yy - (y0 - 1)) * RS, base + (size_t)yy * W * 3 + sbase,
                       (uint32_t)(send - sbase), &bar);
      }
    } else {
      sbase = bx0;
      const int nb = bx1 - bx0;
      for (int i = tid; i < nb * (yb - ya); i += NT) {
        const int yy = ya + i / nb, xb = i % nb;
        tile[(yy - (y0 - 1)) * RS + xb] = base[(size_t)yy * W * 3 + bx0 + xb];
      }
    }
  } else {
    sbase = bx0;
  }
  if (it.gather || !a.aligned) __syncthreads();    // generic byte copies / nothing staged yet

  const int pre_end = it.sharp >= 0 ? it.sharp : it.s1;   // steps [s0,pre_end) before the stencil
  if (it.gather) {
    // every staged pixel is produced by a pull through the gathers (+ pointwise steps up to pre_end)
    const int nx = (bx1 - bx0) / 3;
    for (int i = tid; i < nx * (yb - ya); i += NT) {
      const int yy = ya + i / nx, xx = bx0 / 3 + i % nx;
      int r, g, b;
      eval_gather(row, s_luts, it.s0, pre_end, base, W, H, xx, yy, r, g, b);
      uint8_t* p = tile + (yy - (y0 - 1)) * RS + (xx * 3 - sbase);
      p[0] = (uint8_t)r; p[1] = (uint8_t)g; p[2] = (uint8_t)b;
    }
    __syncthreads();
  } else {
    if (a.aligned) { mbar_wait(&bar, bar_phase); bar_phase ^= 1; }
    if (it.sharp >= 0 && pre_end > it.s0) {
      // pointwise prefix applied in place on tile + halo before the stencil reads neighbours
      const int nx = (bx1 - bx0) / 3;
      for (int i = tid; i < nx * (yb - ya); i += NT) {
        const int yy = ya + i / nx, xx = bx0 / 3 + i % nx;
        uint8_t* p = tile + (yy - (y0 - 1)) * RS + (xx * 3 - sbase);
        int r = p[0], g = p[1], b = p[2];
        for (int k = it.s0; k < pre_end; ++k) apply_point(row.s[k], s_luts + k * 768, xx, yy, r, g, b);
        p[0] = (uint8_t)r; p[1] = (uint8_t)g; p[2] = (uint8_t)b;
      }
      __syncthreads();
    } else if (!a.aligned) {
      // generic loads were already followed by the barrier above
    }
  }

  // ---- compute phase: 4 consecutive pixels per thread -------------------------------------------
  const int ty = tid >> 5, y = y0 + ty;
  const int xq = x0 + (tid & 31) * 4;
  int vr[4], vg[4], vb[4];
  const bool row_ok = y < H;
  const float k1 = __fdiv_rn(1.0f, 13.0f), k5 = __fdiv_rn(5.0f, 13.0f);
  const bool pre_applied = it.gather || it.sharp >= 0;   // steps < pre_end already in the tile
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = xq + i;
    int r = 0, g = 0, b = 0;
    if (row_ok && x < W) {
      const uint8_t* p = tile + (ty + 1) * RS + (x * 3 - sbase);
      r = p[0]; g = p[1]; b = p[2];
      if (fast) { vr[i] = r; vg[i] = g; vb[i] = b; continue; }     // tables applied at the store
      int k = pre_applied ? pre_end : it.s0;
      if (it.sharp >= 0) {
        const DevStep& st = row.s[it.sharp];
        if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) {
          const bool in01 = st.f >= 0.f && st.f <= 1.f;
          const int dr = smooth_1ch(p - RS, p, p + RS, k1, k5);
          const int dg = smooth_1ch(p - RS + 1, p + 1, p + RS + 1, k1, k5);
          const int db = smooth_1ch(p - RS + 2, p + 2, p + RS + 2, k1, k5);
          r = blend_u8(dr, r, st.f, in01); g = blend_u8(dg, g, st.f, in01); b = blend_u8(db, b, st.f, in01);
        }
        k = it.sharp + 1;
      }
      for (; k < it.s1; ++k) apply_point(row.s[k], s_luts + k * 768, x, y, r, g, b);
    }
    vr[i] = r; vg[i] = g; vb[i] = b;
  }

  if (fast && MODE != MODE_F32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { vr[i] = s_comp[vr[i]]; vg[i] = s_comp[256 + vg[i]]; vb[i] = s_comp[512 + vb[i]]; }
  }
  if (MODE == MODE_STATS) {
    unsigned int* hh = s_hist + (tid >> 5) * 768;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (row_ok && xq + i < W) {
        atomicAdd(&hh[vr[i]], 1u); atomicAdd(&hh[256 + vg[i]], 1u); atomicAdd(&hh[512 + vb[i]], 1u);
        lsum += luma_u8(vr[i], vg[i], vb[i]);
      }
  } else if (MODE == MODE_U8) {
    if (row_ok) {
      uint8_t* o = a.out_u8 + ((size_t)it.out * H + y) * W * 3 + (size_t)xq * 3;
      if (xq + 3 < W && (W & 3) == 0) {
        uint32_t w0 = vr[0] | (vg[0] << 8) | (vb[0] << 16) | (vr[1] << 24);
        uint32_t w1 = vg[1] | (vb[1] << 8) | (vr[2] << 16) | (vg[2] << 24);
        uint32_t w2 = vb[2] | (vr[3] << 8) | (vg[3] << 16) | (vb[3] << 24);
        uint32_t* ow = (uint32_t*)o;
        ow[0] = w0; ow[1] = w1; ow[2] = w2;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (xq + i < W) { o[i * 3] = (uint8_t)vr[i]; o[i * 3 + 1] = (uint8_t)vg[i]; o[i * 3 + 2] = (uint8_t)vb[i]; }
      }
    }
  } else {
    if (row_ok) {
      const size_t plane = (size_t)H * W;
      float* o = a.out_f32 + (size_t)it.out * 3 * plane + (size_t)y * W + xq;
      const float* tr = fast ? s_fcomp : s_norm;
      const float* tg = fast ? s_fcomp + 256 : s_norm;
      const float* tb = fast ? s_fcomp + 512 : s_norm;
      if (xq + 3 < W && (W & 3) == 0) {
        __stcs((float4*)o, make_float4(tr[vr[0]], tr[vr[1]], tr[vr[2]], tr[vr[3]]));
        __stcs((float4*)(o + plane), make_float4(tg[vg[0]], tg[vg[1]], tg[vg[2]], tg[vg[3]]));
        __stcs((float4*)(o + 2 * plane), make_float4(tb[vb[0]], tb[vb[1]], tb[vb[2]], tb[vb[3]]));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (xq + i < W) { o[i] = tr[vr[i]]; o[plane + i] = tg[vg[i]]; o[2 * plane + i] = tb[vb[i]]; }
      }
    }
  }
  fence_proxy_async();  // generic accesses to the tile are ordered before the next iteration's bulk copies
  __syncthreads();      // the staged tile is reused by the next iteration
  }  // tile loop

  if (MODE == MODE_STATS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    Stat* st = a.stats + it.out;
    for (int i = tid; i < 768; i += NT) {
      unsigned int v = 0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) v += s_hist[w * 768 + i];
      if (v) atomicAdd(&st->hist[0][0] + i, v);
    }
    if ((tid & 31) == 0 && lsum) atomicAdd(&st->luma_sum, (unsigned long long)lsum);
  }
}

// ---- streaming kernel: programs made only of pointwise steps (tables, Color, Cutout) ---------------------------
// 9 of the 10 searchable ops are pointwise, so most rows never need a staged tile.  A thread owns pixel QUADS: four
// consecutive pixels = 12 bytes = three 32-bit loads, and a warp's three load instructions cover 384 contiguous bytes;
// the results leave as one 16-byte store per channel plane (MODE_F32: a warp writes 512 contiguous bytes per plane) or
// three 32-bit stores (MODE_U8).  SQ quads per thread are loaded before the first use, so a CTA keeps
// SQ * 12 B * 256 threads of reads in flight without any shared-memory staging or barrier in the loop.
// Requires W % 4 == 0 (a quad never straddles a row) and 4-byte aligned images; other shapes use pass_kernel.
constexpr int SQ = 4;             // quads per thread per iteration

struct StreamArgs {
  const DevRow* rows;
  const PassItem* items;
  const uint8_t* luts;
  const uint8_t* src;
  const uint8_t* scratch;
  uint8_t* out_u8;
  float* out_f32;
  Stat* stats;
  int H, W;
  int quads_per_cta;              // multiple of NT * SQ
};

template <int MODE>
__global__ void __launch_bounds__(NT) stream_kernel(const StreamArgs a) {
  __shared__ __align__(16) uint8_t s_luts[AADG_MAX_OPS * 768];
  __shared__ float s_f[768];                                       // normalised value of (composite) table entry
  __shared__ __align__(16) uint8_t s_comp[768];
  __shared__ unsigned int s_hist[MODE == MODE_STATS ? (NT / 32) * 768 : 1];
  __shared__ DevRow s_row;

  const PassItem it = a.items[blockIdx.y];
  const int W = a.W, H = a.H;
  const int tid = threadIdx.x;
  if (tid < (int)(sizeof(DevRow) / 4)) ((int*)&s_row)[tid] = ((const int*)&a.rows[it.row])[tid];
  const bool fast = it.pad_ != 0;
  {
    const uint4* g = (const uint4*)(a.luts + (size_t)it.row * AADG_MAX_OPS * 768);
    uint4* sl = (uint4*)s_luts;
    for (int i = tid; i < AADG_MAX_OPS * 768 / 16; i += NT) sl[i] = g[i];
    if (MODE == MODE_STATS)
      for (int i = tid; i < (NT / 32) * 768; i += NT) s_hist[i] = 0;
  }
  __syncthreads();
  for (int i = tid; i < 768; i += NT) {
    const int c = i >> 8;
    int v = i & 255;
    if (fast)
      for (int k = it.s0; k < it.s1; ++k) v = s_luts[k * 768 + c * 256 + v];
    s_comp[i] = (uint8_t)v;
    s_f[i] = __fsub_rn(__fdiv_rn((float)v, 127.5f), 1.0f);          // Normalize_dg: x / 127.5 - 1
  }
  __syncthreads();
  const DevRow& row = s_row;
  const size_t plane = (size_t)H * W;
  const uint8_t* base = it.base < 0 ? a.src + (size_t)row.src * plane * 3 : a.scratch + (size_t)it.base * plane * 3;
  const uint32_t* in = (const uint32_t*)base;
  const int n_quads = (int)(plane >> 2);
  const int q_begin = blockIdx.x * a.quads_per_cta;
  const int q_end = min(n_quads, q_begin + a.quads_per_cta);
  const int wq = W >> 2;                       // quads per image row
  unsigned int lsum = 0;

  for (int q0 = q_begin + tid; q0 < q_end; q0 += NT * SQ) {
    uint32_t w[SQ][3];
#pragma unroll
    for (int u = 0; u < SQ; ++u) {
      const int q = q0 + u * NT;
      if (q < q_end) {
        const uint32_t* p = in + (size_t)q * 3;
        w[u][0] = __ldg(p); w[u][1] = __ldg(p + 1); w[u][2] = __ldg(p + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < SQ; ++u) {
      const int q = q0 + u * NT;
      if (q >= q_end) break;
      // bytes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
      if (!fast) {          // the step interpreter stays out of line (code size: see apply_steps_rgb)
        const int y = q / wq, x0 = (q - y * wq) << 2;
        const Quad o = apply_steps_quad(&s_row, s_luts, it.s0, it.s1, x0, y, Quad{w[u][0], w[u][1], w[u][2]});
        w[u][0] = o.w0; w[u][1] = o.w1; w[u][2] = o.w2;
      }
      int vr[4], vg[4], vb[4];
      vr[0] = w[u][0] & 255; vg[0] = (w[u][0] >> 8) & 255; vb[0] = (w[u][0] >> 16) & 255; vr[1] = w[u][0] >> 24;
      vg[1] = w[u][1] & 255; vb[1] = (w[u][1] >> 8) & 255; vr[2] = (w[u][1] >> 16) & 255; vg[2] = w[u][1] >> 24;
      vb[2] = w[u][2] & 255; vr[3] = (w[u][2] >> 8) & 255; vg[3] = (w[u][2] >> 16) & 255; vb[3] = w[u][2] >> 24;
      if (MODE == MODE_STATS) {
        unsigned int* hh = s_hist + (tid >> 5) * 768;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = fast ? s_comp[vr[i]] : vr[i], g = fast ? s_comp[256 + vg[i]] : vg[i],
                    b = fast ? s_comp[512 + vb[i]] : vb[i];
          atomicAdd(&hh[r], 1u); atomicAdd(&hh[256 + g], 1u); atomicAdd(&hh[512 + b], 1u);
          lsum += luma_u8(r, g, b);
        }
      } else if (MODE == MODE_U8) {
        if (fast) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { vr[i] = s_comp[vr[i]]; vg[i] = s_comp[256 + vg[i]]; vb[i] = s_comp[512 + vb[i]]; }
        }
        uint32_t* o = (uint32_t*)(a.out_u8 + (size_t)it.out * plane * 3) + (size_t)q * 3;
        o[0] = vr[0] | (vg[0] << 8) | (vb[0] << 16) | (vr[1] << 24);
        o[1] = vg[1] | (vb[1] << 8) | (vr[2] << 16) | (vg[2] << 24);
        o[2] = vb[2] | (vr[3] << 8) | (vg[3] << 16) | (vb[3] << 24);
      } else {
        // fast: s_f holds the normalised composite table per channel; otherwise entries 0..255 of every channel block
        // are the plain normalisation map (s_comp is the identity then)
        const float* tr = s_f;
        const float* tg = s_f + 256;
        const float* tb = s_f + 512;
        float* o = a.out_f32 + (size_t)it.out * 3 * plane + ((size_t)q << 2);
        __stcs((float4*)o, make_float4(tr[vr[0]], tr[vr[1]], tr[vr[2]], tr[vr[3]]));
        __stcs((float4*)(o + plane), make_float4(tg[vg[0]], tg[vg[1]], tg[vg[2]], tg[vg[3]]));
        __stcs((float4*)(o + 2 * plane), make_float4(tb[vb[0]], tb[vb[1]], tb[vb[2]], tb[vb[3]]));
      }
    }
  }

  if (MODE == MODE_STATS) {
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    Stat* st = a.stats + it.out;
    for (int i = tid; i < 768; i += NT) {
      unsigned int v = 0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) v += s_hist[w * 768 + i];
      if (v) atomicAdd(&st->hist[0][0] + i, v);
    }
    if ((tid & 31) == 0 && lsum) atomicAdd(&st->luma_sum, (unsigned long long)lsum);
  }
}

// ---- streaming stencil kernel: [pointwise steps] Sharpness [pointwise steps] -------------------------------------------
// Same quad ownership as stream_kernel, for programs that contain ONE Sharpness step and no gather.  A WARP owns a strip
// of 128 pixels x `rows_per_band` image rows and walks DOWN it: per image row a thread loads the five 32-bit words that
// hold its four pixels plus one neighbour on each side (a warp's loads cover 384 contiguous bytes + 2 words, issued one
// row ahead of their use), and keeps three rows of horizontal 3-sums in registers.  No shared-memory tile, no barrier
// in the loop; every image byte is loaded (rows+2)/rows times per strip.
//
// Pillow's SMOOTH (ImagingFilter3x3, Filter.c) is float32: ss = 0.5 + sum k_i*x_i with k = {1,1,1,1,5,1,1,1,1}/13 in a
// fixed order, then clip8 by truncation.  It equals the INTEGER expression d = (S + 5c + 6) / 13 (S = the 8 neighbours,
// c = the centre) for every input, which is what the kernel evaluates, 2 x 16-bit lanes per register:
//   * the exact value ss* = (2(S + 5c) + 13) / 26 has an ODD numerator, so it is never an integer and sits at least
//     1/26 = 0.038 away from one;
//   * the float evaluation is off by < 1.1e-4 (nine products with relative error <= 2^-23 on values <= 98.1, nine sums
//     with half-ulp <= 2^-17 on partial sums < 256), so trunc(ss) = floor(ss*) = (2(S+5c)+13) div 26, and because the
//     numerator is odd that is (2(S+5c)+12) div 26 = (S + 5c + 6) div 13; 0 <= d <= 255 (clip8 never clips);
//   * n div 13 = (n * 5042) >> 16 for n <= 3321 (5042 = ceil(2^16/13), error n*0.77/65536 < 1/13).
// The blend d + f*(x - d) keeps Pillow's float32 operation order (ImagingBlend); its int->float conversions use the
// 1.5*2^23 trick (exact for |v| < 2^22) so that they run on the fp32 / integer pipes, not on the conversion unit.
struct SharpRow {            // one image row as a thread keeps it: 12 horizontal 3-sums and its 12 own bytes,
  uint32_t He[3], Ho[3];     // two zero-extended 16-bit lanes per register: word m of the quad, even / odd byte
  uint32_t Ce[3], Co[3];     // lane lo <-> byte j = 4m (+1 odd), lane hi <-> byte j = 4m + 2 (+1 odd); j = 3*pixel + channel
};

template <int MODE>
__global__ void __launch_bounds__(NT, 3) stencil_kernel(const StreamArgs a) {
  __shared__ __align__(16) uint8_t s_luts[AADG_MAX_OPS * 768];
  __shared__ float s_f[256];
  __shared__ unsigned int s_hist[MODE == MODE_STATS ? (NT / 32) * 768 : 1];
  __shared__ DevRow s_row;

  const PassItem it = a.items[blockIdx.y];
  const int W = a.W, H = a.H;
  const int tid = threadIdx.x;
  if (tid < (int)(sizeof(DevRow) / 4)) ((int*)&s_row)[tid] = ((const int*)&a.rows[it.row])[tid];
  {
    const uint4* g = (const uint4*)(a.luts + (size_t)it.row * AADG_MAX_OPS * 768);
    uint4* sl = (uint4*)s_luts;
    for (int i = tid; i < AADG_MAX_OPS * 768 / 16; i += NT) sl[i] = g[i];
    s_f[tid] = __fsub_rn(__fdiv_rn((float)tid, 127.5f), 1.0f);
    if (MODE == MODE_STATS)
      for (int i = tid; i < (NT / 32) * 768; i += NT) s_hist[i] = 0;
  }
  __syncthreads();
  const DevRow& row = s_row;
  const size_t plane = (size_t)H * W;
  const uint8_t* base = it.base < 0 ? a.src + (size_t)row.src * plane * 3 : a.scratch + (size_t)it.base * plane * 3;
  const int wq = W >> 2;                        // quads per row
  const int row_words = wq * 3;
  const int RB = a.quads_per_cta;               // here: image rows per band
  const int strips = (wq + 31) >> 5, bands = (H + RB - 1) / RB;
  const float sf = row.s[it.sharp].f;
  const bool has_pre = it.sharp > it.s0, has_post = it.sharp + 1 < it.s1;
  unsigned int lsum = 0;

  const int unit = blockIdx.x * (NT / 32) + (tid >> 5);          // (strip, band) of this warp
  const int qx = (unit % strips) * 32 + (tid & 31);
  const int band = unit / strips;
  if (band < bands && qx < wq) {
    const int x0 = qx << 2;
    const int y_begin = band * RB, y_end = min(H, y_begin + RB);
    const uint32_t* colp = (const uint32_t*)base + qx * 3;
    const bool has_left = qx > 0, has_right = qx + 1 < wq;
    const bool xin0 = x0 >= 1, xin3 = x0 + 3 <= W - 2;            // pixels 1, 2 of a quad are never on the x border

    uint32_t w[5];     // raw words of the next row to convert: loaded one iteration ahead (software prefetch)
    // image rows are clamped: rows outside the image are only ever neighbours of pass-through border rows
    auto fetch = [&](int yy) {
      const uint32_t* p = colp + (size_t)min(max(yy, 0), H - 1) * row_words;
      w[0] = has_left ? __ldg(p - 1) : 0u;
      w[1] = __ldg(p); w[2] = __ldg(p + 1); w[3] = __ldg(p + 2);
      w[4] = has_right ? __ldg(p + 3) : 0u;
    };
    // the 20 bytes in w (image row yy; own byte j is byte 4 + j, its x neighbours are bytes j + 1 and j + 7)
    auto convert = [&](int yy, SharpRow& R) {
      if (has_pre) {          // steps before the stencil on all six pixels, written back into w: the own quad (bytes
        const int yc = min(max(yy, 0), H - 1);       // 4..15 = w[1..3]) in one call, the two neighbours one by one
        const Quad o = apply_steps_quad(&s_row, s_luts, it.s0, it.sharp, x0, yc, Quad{w[1], w[2], w[3]});
        w[1] = o.w0; w[2] = o.w1; w[3] = o.w2;
        const uint32_t lft = apply_steps_rgb(&s_row, s_luts, it.s0, it.sharp, x0 - 1, yc, w[0] >> 8);          // bytes 1..3
        w[0] = (w[0] & 0xFFu) | (lft << 8);
        const uint32_t rgt = apply_steps_rgb(&s_row, s_luts, it.s0, it.sharp, x0 + 4, yc, w[4] & 0xFFFFFFu);   // bytes 16..18
        w[4] = (w[4] & 0xFF000000u) | rgt;
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const uint32_t xl = __funnelshift_r(w[m], w[m + 1], 8);          // bytes j + 1  (left neighbours)
        const uint32_t xc = w[m + 1];                                     // bytes j + 4  (own)
        const uint32_t xr = __funnelshift_r(w[m + 1], w[m + 2], 24);     // bytes j + 7  (right neighbours)
        R.Ce[m] = __byte_perm(xc, 0u, 0x4240);
        R.Co[m] = __byte_perm(xc, 0u, 0x4341);
        R.He[m] = __byte_perm(xl, 0u, 0x4240) + R.Ce[m] + __byte_perm(xr, 0u, 0x4240);
        R.Ho[m] = __byte_perm(xl, 0u, 0x4341) + R.Co[m] + __byte_perm(xr, 0u, 0x4341);
      }
    };
    auto output = [&](int y, const SharpRow& up, const SharpRow& mid, const SharpRow& dn) {
      int v[12];
      const bool row_in = y >= 1 && y <= H - 2;
#pragma unroll
      for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int par = 0; par < 2; ++par) {
          const uint32_t C = par ? mid.Co[m] : mid.Ce[m];
          const uint32_t V = par ? up.Ho[m] + mid.Ho[m] + dn.Ho[m] : up.He[m] + mid.He[m] + dn.He[m];
          const uint32_t M = V + (C << 2) + 0x00060006u;                  // S + 5c + 6 per lane (<= 3321)
#pragma unroll
          for (int lane = 0; lane < 2; ++lane) {
            const int j = 4 * m + par + 2 * lane;                           // byte index = 3 * pixel + channel
            const int x = lane ? (int)(C >> 16) : (int)(C & 0xFFFFu);
            const int d = lane ? (int)__umulhi(M & 0xFFFF0000u, 5042u) : (int)(((M & 0xFFFFu) * 5042u) >> 16);
            // ImagingBlend: d + f * (x - d), clipped when f is outside [0, 1] (inside it the clip is the identity)
            float t = __fadd_rn(i2f_small(d), __fmul_rn(sf, i2f_small(x - d)));
            t = fminf(fmaxf(t, 0.f), 255.f);
            const int px = j / 3;
            const bool inside = row_in && (px == 0 ? xin0 : (px == 3 ? xin3 : true));
            v[j] = inside ? (int)t : x;
          }
        }
      if (has_post) {
        Quad o;
        o.w0 = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
        o.w1 = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
        o.w2 = v[8] | (v[9] << 8) | (v[10] << 16) | (v[11] << 24);
        o = apply_steps_quad(&s_row, s_luts, it.sharp + 1, it.s1, x0, y, o);
#pragma unroll
        for (int j = 0; j < 12; ++j) v[j] = ((j < 4 ? o.w0 : (j < 8 ? o.w1 : o.w2)) >> (8 * (j & 3))) & 255;
      }
      int vr[4], vg[4], vb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { vr[i] = v[3 * i]; vg[i] = v[3 * i + 1]; vb[i] = v[3 * i + 2]; }
      const size_t q = (size_t)y * wq + qx;
      if (MODE == MODE_STATS) {
        unsigned int* hh = s_hist + (tid >> 5) * 768;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          atomicAdd(&hh[vr[i]], 1u); atomicAdd(&hh[256 + vg[i]], 1u); atomicAdd(&hh[512 + vb[i]], 1u);
          lsum += luma_u8(vr[i], vg[i], vb[i]);
        }
      } else if (MODE == MODE_U8) {
        uint32_t* o = (uint32_t*)(a.out_u8 + (size_t)it.out * plane * 3) + q * 3;
        o[0] = vr[0] | (vg[0] << 8) | (vb[0] << 16) | (vr[1] << 24);
        o[1] = vg[1] | (vb[1] << 8) | (vr[2] << 16) | (vg[2] << 24);
        o[2] = vb[2] | (vr[3] << 8) | (vg[3] << 16) | (vb[3] << 24);
      } else {
        float* o = a.out_f32 + (size_t)it.out * 3 * plane + (q << 2);
        __stcs((float4*)o, make_float4(s_f[vr[0]], s_f[vr[1]], s_f[vr[2]], s_f[vr[3]]));
        __stcs((float4*)(o + plane), make_float4(s_f[vg[0]], s_f[vg[1]], s_f[vg[2]], s_f[vg[3]]));
        __stcs((float4*)(o + 2 * plane), make_float4(s_f[vb[0]], s_f[vb[1]], s_f[vb[2]], s_f[vb[3]]));
      }
    };
    // one image row per step: convert the prefetched row y+1, prefetch row y+2, emit row y.  The three row records
    // rotate roles by NAME (three steps per loop trip), not by copying registers.
    auto step = [&](int y, const SharpRow& up, const SharpRow& mid, SharpRow& dn) {
      convert(y + 1, dn);
      fetch(y + 2);
      output(y, up, mid, dn);
    };
    SharpRow R0, R1, R2;
    fetch(y_begin - 1); convert(y_begin - 1, R0);
    fetch(y_begin); convert(y_begin, R1);
    fetch(y_begin + 1);
    for (int y = y_begin; y < y_end; y += 3) {
      step(y, R0, R1, R2);
      if (y + 1 >= y_end) break;
      step(y + 1, R1, R2, R0);
      if (y + 2 >= y_end) break;
      step(y + 2, R2, R0, R1);
    }
  }

  if (MODE == MODE_STATS) {
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    Stat* st = a.stats + it.out;
    for (int i = tid; i < 768; i += NT) {
      unsigned int v = 0;
#pragma unroll
      for (int w2 = 0; w2 < NT / 32; ++w2) v += s_hist[w2 * 768 + i];
      if (v) atomicAdd(&st->hist[0][0] + i, v);
    }
    if ((tid & 31) == 0 && lsum) atomicAdd(&st->luma_sum, (unsigned long long)lsum);
  }
}

// ---- look-up table construction ------------------------------------------------------------------
// One CTA of 256 threads per (row, step): thread i owns entry i of the three channel tables.
struct LutItem { int row; int step; };

__global__ void __launch_bounds__(256) lut_kernel(const DevRow* rows, const LutItem* items,
                                                  uint8_t* luts, const Stat* stats, int n_pix) {
  __shared__ unsigned int h[3][256];
  __shared__ unsigned int cum[3][256];
  __shared__ int s_lo[3], s_hi[3], s_nnz[3];
  const LutItem it = items[blockIdx.x];
  const DevRow& row = rows[it.row];
  const DevStep st = row.s[it.step];
  const int i = threadIdx.x;
  uint8_t* out = luts + ((size_t)it.row * AADG_MAX_OPS + it.step) * 768;
  if (st.kind != K_LUT) return;

  if (st.lut_kind == L_INVERT) {
    out[i] = out[256 + i] = out[512 + i] = (uint8_t)(255 - i);
    return;
  }
  if (st.lut_kind == L_SOLARIZE) {          // ImageOps.solarize: i < threshold ? i : 255-i
    const uint8_t v = (uint8_t)(i < st.p[0] ? i : 255 - i);
    out[i] = out[256 + i] = out[512 + i] = v;
    return;
  }
  if (st.lut_kind == L_POSTERIZE) {         // ImageOps.posterize: i & mask
    out[i] = out[256 + i] = out[512 + i] = (uint8_t)(i & st.p[0]);
    return;
  }
  const bool in01 = st.f >= 0.f && st.f <= 1.f;
  if (st.lut_kind == L_BRIGHT) {            // ImageEnhance.Brightness: blend(0, x, f)
    out[i] = out[256 + i] = out[512 + i] = (uint8_t)blend_u8(0, i, st.f, in01);
    return;
  }
  const Stat& sg = stats[st.p[1]];
  if (st.lut_kind == L_CONTRAST) {
    // ImageEnhance.Contrast: mean = int(ImageStat.Stat(L).mean[0] + 0.5); blend(mean, x, f)
    const double mean = __dadd_rn(__ddiv_rn((double)sg.luma_sum, (double)n_pix), 0.5);
    const int m = (int)mean;
    out[i] = out[256 + i] = out[512 + i] = (uint8_t)blend_u8(m, i, st.f, in01);
    return;
  }
  // histogram ops: fetch (or derive through the earlier tables) the histogram of the current image
  for (int c = 0; c < 3; ++c) { h[c][i] = 0; }
  if (i < 3) { s_lo[i] = 256; s_hi[i] = -1; s_nnz[i] = 0; }
  __syncthreads();
  if (st.p[2]) {
    for (int c = 0; c < 3; ++c) {
      int v = i;
      for (int k = 0; k < it.step; ++k)
        v = luts[((size_t)it.row * AADG_MAX_OPS + k) * 768 + c * 256 + v];
      const unsigned int cnt = sg.hist[c][i];
      if (cnt) atomicAdd(&h[c][v], cnt);
    }
  } else {
    for (int c = 0; c < 3; ++c) h[c][i] = sg.hist[c][i];
  }
  __syncthreads();
  for (int c = 0; c < 3; ++c)
    if (h[c][i]) { atomicMin(&s_lo[c], i); atomicMax(&s_hi[c], i); atomicAdd(&s_nnz[c], 1); }
  // exclusive prefix sums of each channel histogram: warp c scans channel c, 8 bins per lane
  {
    const int w = i >> 5, lane = i & 31;
    if (w < 3) {
      unsigned int loc[8], run = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = run; run += h[w][lane * 8 + j]; }
      unsigned int inc = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      const unsigned int excl = inc - run;
#pragma unroll
      for (int j = 0; j < 8; ++j) cum[w][lane * 8 + j] = excl + loc[j];
    }
  }
  __syncthreads();
  for (int c = 0; c < 3; ++c) {
    int v = i;
    if (st.lut_kind == L_AUTOCONTRAST) {
      // ImageOps.autocontrast(cutoff=0): ix = int(ix*scale + offset) in doubles, clipped
      const int lo = s_lo[c], hi = s_hi[c];
      if (hi > lo) {
        const double scale = __ddiv_rn(255.0, (double)(hi - lo));
        const double offset = __dmul_rn(-(double)lo, scale);
        const double t = __dadd_rn(__dmul_rn((double)i, scale), offset);
        const int q = (int)t;
        v = q < 0 ? 0 : (q > 255 ? 255 : q);
      }
    } else {
      // ImageOps.equalize: step = (sum(nonzero) - last nonzero) // 255; lut[i] = (step//2 + cum[i]) // step
      if (s_nnz[c] > 1) {
        const unsigned int total = cum[c][255] + h[c][255];
        const unsigned int step = (total - h[c][s_hi[c]]) / 255u;
        if (step) {
          const unsigned int q = (step / 2 + cum[c][i]) / step;
          v = q > 255u ? 255 : (int)q;
        }
      }
    }
    out[c * 256 + i] = (uint8_t)v;
  }
}

// ---- masks and labels ----------------------------------------------------------------------------
// data/basic.py edits the mask in Cutout (fill 0) and the geometric ops (same warp as the image).
__global__ void mask_kernel(const DevRow* rows, const uint8_t* src_masks, uint8_t* out, int n_rows,
                            int H, int W) {
  const int r = blockIdx.y;
  const DevRow row = rows[r];
  const uint8_t* base = src_masks + (size_t)row.src * H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    int x = i % W, y = i / W;
    int cx[AADG_MAX_OPS + 1], cy[AADG_MAX_OPS + 1];
    cx[row.n_steps] = x; cy[row.n_steps] = y;
    int dead = -1;
    for (int k = row.n_steps - 1; k >= 0; --k) {
      const DevStep& st = row.s[k];
      int px = cx[k + 1], py = cy[k + 1];
      if (dead < 0) {
        if (st.kind == K_AFFINE) {
          long long xx = (long long)st.p[2] + (long long)py * st.p[1] + (long long)px * st.p[0];
          long long yy = (long long)st.p[5] + (long long)py * st.p[4] + (long long)px * st.p[3];
          long long xi = xx >> 16, yi = yy >> 16;
          if (xi < 0 || xi >= W || yi < 0 || yi >= H) { dead = k; xi = 0; yi = 0; }
          px = (int)xi; py = (int)yi;
        }
        // Flip: the reference mirrors the image only (data/basic.py:82-83)
      }
      cx[k] = px; cy[k] = py;
    }
    int v = dead < 0 ? base[(size_t)cy[0] * W + cx[0]] : 0;
    for (int k = dead + 1; k < row.n_steps; ++k) {
      const DevStep& st = row.s[k];
      if (st.kind == K_CUTOUT && cx[k + 1] >= st.p[0] && cx[k + 1] <= st.p[2] &&
          cy[k + 1] >= st.p[1] && cy[k + 1] <= st.p[3])
        v = 0;
    }
    out[(size_t)r * H * W + i] = (uint8_t)v;
  }
}

// Normalize_dg mask branch + to_multilabel + ToTensor (data/transform.py:153-172,244-249,217-236):
// optic: >200 -> [0,0]; (50,201) -> [0,1]; else [1,1];  vessel: != 0 -> [1].
__global__ void label_kernel(const DevRow* rows, const uint8_t* src_masks, float* out, int H, int W,
                             int dataset) {
  const int r = blockIdx.y;
  const int src = rows[r].src;
  const size_t plane = (size_t)H * W;
  const uint8_t* m = src_masks + (size_t)src * plane;
  const int nch = dataset == AADG_DATASET_OPTIC ? 2 : 1;
  float* o = out + (size_t)r * nch * plane;
  const size_t n4 = plane / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const uchar4 v = ((const uchar4*)m)[i];
    const int vv[4] = {v.x, v.y, v.z, v.w};
    float c0[4], c1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (dataset == AADG_DATASET_OPTIC) {
        const bool bg = vv[j] > 200, ring = vv[j] > 50 && vv[j] < 201;
        c0[j] = (!bg && !ring) ? 1.f : 0.f;
        c1[j] = bg ? 0.f : 1.f;
      } else {
        c0[j] = vv[j] != 0 ? 1.f : 0.f;
      }
    }
    __stcs((float4*)o + i, make_float4(c0[0], c0[1], c0[2], c0[3]));
    if (nch == 2) __stcs((float4*)(o + plane) + i, make_float4(c1[0], c1[1], c1[2], c1[3]));
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < plane; i += blockDim.x) {
      const int v = m[i];
      if (dataset == AADG_DATASET_OPTIC) {
        const bool bg = v > 200, ring = v > 50 && v < 201;
        o[i] = (!bg && !ring) ? 1.f : 0.f;
        o[plane + i] = bg ? 0.f : 1.f;
      } else {
        o[i] = v != 0 ? 1.f : 0.f;
      }
    }
}

// ---- host: compile rows into programs and schedule the passes ---------------------------------------
struct Plan {
  std::vector<DevRow> rows;
  std::vector<PassItem> src_stats;                    // one per source that needs statistics
  std::vector<PassItem> mat[AADG_MAX_OPS + 1];        // materialise before step k
  std::vector<PassItem> stat[AADG_MAX_OPS];           // statistics pass feeding step k
  std::vector<LutItem> lut[AADG_MAX_OPS];             // tables of step k
  std::vector<PassItem> fin;                          // final pass per row that waits for statistics / a materialised image
  std::vector<PassItem> fin_ind;                      // final pass per INDEPENDENT row (no statistics op, one pass)
  std::vector<LutItem> lut_ind;                       // tables of the independent rows (any step: none needs statistics)
  int n_stat_slots = 0;
  int n_scratch = 0;
};

static bool is_stat_op(int op) {
  return op == AADG_OP_AUTOCONTRAST || op == AADG_OP_EQUALIZE || op == AADG_OP_CONTRAST;
}

static int compile(const aadg_aug_row_t* in, int n_rows, int n_src, Plan& pl) {
  pl.rows.resize(n_rows);
  std::vector<int> src_slot(n_src, -1);
  pl.n_stat_slots = 0;
  auto slot_for_src = [&](int s) {
    if (src_slot[s] < 0) {
      src_slot[s] = pl.n_stat_slots++;
      PassItem it{};  // row filled below: any row with this src works as a carrier
      it.row = -1; it.s0 = 0; it.s1 = 0; it.base = -1; it.out = src_slot[s]; it.sharp = -1; it.gather = 0;
      it.pad_ = s;
      pl.src_stats.push_back(it);
    }
    return src_slot[s];
  };
  for (int r = 0; r < n_rows; ++r) {
    const aadg_aug_row_t& a = in[r];
    AADG_REQUIRE(a.src >= 0 && a.src < n_src, "row %d: src %d out of range [0,%d)", r, a.src, n_src);
    AADG_REQUIRE(a.n_ops >= 0 && a.n_ops <= AADG_MAX_OPS, "row %d: n_ops %d out of range", r, a.n_ops);
    DevRow& d = pl.rows[r];
    memset(&d, 0, sizeof(d));
    d.src = a.src;
    d.n_steps = a.n_ops;
    int seg = 0, base = -1;
    bool sharp_seen = false, all_lut = true;
    int sharp_idx = -1;
    bool gather_seen = false;
    auto item = [&](int s1) {
      PassItem it{};
      it.row = r; it.s0 = seg; it.s1 = s1; it.base = base; it.sharp = -1; it.gather = 0;
      it.pad_ = 1;        // all steps are look-up tables (also true for an empty program)
      for (int k = seg; k < s1; ++k) {
        if (d.s[k].kind == K_SHARP) it.sharp = k;
        if (d.s[k].kind == K_AFFINE || d.s[k].kind == K_FLIP) it.gather = 1;
        if (d.s[k].kind != K_LUT) it.pad_ = 0;
      }
      return it;
    };
    for (int k = 0; k < a.n_ops; ++k) {
      const int op = a.op[k];
      AADG_REQUIRE(op >= 0 && op < AADG_OP_COUNT, "row %d: unknown op id %d", r, op);
      DevStep& st = d.s[k];
      const bool barrier = op == AADG_OP_SHARPNESS || (op >= AADG_OP_SHEAR_X && op <= AADG_OP_FLIP);
      if (barrier && sharp_seen) {
        // a stencil's output is needed at arbitrary neighbours/positions: materialise it first
        PassItem it = item(k);
        it.out = pl.n_scratch++;
        pl.mat[k].push_back(it);
        base = it.out; seg = k; sharp_seen = false; sharp_idx = -1; gather_seen = false;
      }
      (void)sharp_idx; (void)gather_seen;
      switch (op) {
        case AADG_OP_INVERT: st.kind = K_LUT; st.lut_kind = L_INVERT; break;
        case AADG_OP_SOLARIZE: st.kind = K_LUT; st.lut_kind = L_SOLARIZE; st.p[0] = a.iparam[k][0]; break;
        case AADG_OP_POSTERIZE: st.kind = K_LUT; st.lut_kind = L_POSTERIZE; st.p[0] = a.iparam[k][0]; break;
        case AADG_OP_BRIGHTNESS: st.kind = K_LUT; st.lut_kind = L_BRIGHT; st.f = a.fparam[k]; break;
        case AADG_OP_CONTRAST: st.kind = K_LUT; st.lut_kind = L_CONTRAST; st.f = a.fparam[k]; break;
        case AADG_OP_AUTOCONTRAST: st.kind = K_LUT; st.lut_kind = L_AUTOCONTRAST; break;
        case AADG_OP_EQUALIZE: st.kind = K_LUT; st.lut_kind = L_EQUALIZE; break;
        case AADG_OP_COLOR: st.kind = K_COLOR; st.f = a.fparam[k]; break;
        case AADG_OP_SHARPNESS: st.kind = K_SHARP; st.f = a.fparam[k]; sharp_seen = true; break;
        case AADG_OP_CUTOUT:
          st.kind = K_CUTOUT;
          for (int j = 0; j < 4; ++j) st.p[j] = a.iparam[k][j];
          break;
        case AADG_OP_FLIP: st.kind = K_FLIP; break;
        default:
          st.kind = K_AFFINE;
          for (int j = 0; j < 6; ++j) st.p[j] = a.iparam[k][j];
      }
      if (is_stat_op(op)) {
        if (k == 0) {
          st.p[1] = slot_for_src(a.src); st.p[2] = 0;
        } else if (all_lut && op != AADG_OP_CONTRAST) {
          st.p[1] = slot_for_src(a.src); st.p[2] = 1;      // histogram pushed through the tables
        } else {
          PassItem it = item(k);
          it.out = -1 - (r * AADG_MAX_OPS + k);             // own slot, numbered after the sources
          pl.stat[k].push_back(it);
          st.p[1] = it.out; st.p[2] = 0;
        }
      }
      if (st.kind != K_LUT) all_lut = false;
    }
    // an INDEPENDENT row needs no statistics and no materialised intermediate: its tables and its single pass can run
    // while the other rows' statistics -> tables chain is still in flight
    bool independent = base < 0;
    for (int k = 0; k < a.n_ops; ++k) independent = independent && !is_stat_op(a.op[k]);
    for (int k = 0; k < a.n_ops; ++k)
      if (d.s[k].kind == K_LUT) (independent ? pl.lut_ind : pl.lut[k]).push_back(LutItem{r, k});
    PassItem it = item(a.n_ops);
    it.out = r;
    (independent ? pl.fin_ind : pl.fin).push_back(it);
  }
  // own statistics slots follow the per-source ones
  int next = pl.n_stat_slots;
  for (int k = 0; k < AADG_MAX_OPS; ++k)
    for (PassItem& it : pl.stat[k]) {
      const int key = -1 - it.out;
      it.out = next++;
      pl.rows[key / AADG_MAX_OPS].s[key % AADG_MAX_OPS].p[1] = it.out;
    }
  pl.n_stat_slots = next;
  // carriers for source statistics: any row of that source; steps [0,0) of it
  for (PassItem& it : pl.src_stats) {
    const int s = it.pad_;
    for (int r = 0; r < n_rows; ++r)
      if (pl.rows[r].src == s) { it.row = r; break; }
    it.pad_ = 1;   // empty program: identity table, fast path
  }
  return AADG_OK;
}

struct Layout {
  size_t rows, items, lut_items, luts, stats, scratch, mask_dummy, total;
  size_t n_items, n_lut_items;
};

static Layout layout(int n_rows, int n_src, int H, int W) {
  Layout L{};
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = align_up(off, 256); off = o + b; return o; };
  L.n_items = (size_t)n_src + (size_t)n_rows * (2 * AADG_MAX_OPS + 2);
  L.n_lut_items = (size_t)n_rows * AADG_MAX_OPS;
  L.rows = take(sizeof(DevRow) * n_rows);
  L.items = take(sizeof(PassItem) * L.n_items);
  L.lut_items = take(sizeof(LutItem) * L.n_lut_items);
  L.luts = take((size_t)n_rows * AADG_MAX_OPS * 768);
  L.stats = take(sizeof(Stat) * ((size_t)n_src + (size_t)n_rows * AADG_MAX_OPS));
  L.scratch = take((size_t)n_rows * (AADG_MAX_OPS - 1) * H * W * 3);
  L.total = align_up(off, 256);
  return L;
}

static int num_sms_u8() {
  static int v = 0;
  if (!v) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

// items [0, n) are pointwise-only programs on a quad-aligned shape: the streaming kernel
template <int MODE>
static int launch_stream(const PassArgs& base, const PassItem* d_items, int n, cudaStream_t st) {
  if (n == 0) return AADG_OK;
  StreamArgs a{};
  a.rows = base.rows; a.luts = base.luts; a.src = base.src; a.scratch = base.scratch;
  a.out_u8 = base.out_u8; a.out_f32 = base.out_f32; a.stats = base.stats; a.H = base.H; a.W = base.W;
  const long long n_quads = (long long)base.H * base.W / 4;
  // CTAs per image: enough CTAs for ~6 waves of 8 resident CTAs per SM when the batch is small, at least one
  // iteration's worth of quads each, at most 64 iterations (amortises the per-CTA table set-up)
  const int unit = NT * SQ;
  // a statistics pass pays a per-CTA histogram set-up and a 768-counter flush: far fewer, longer CTAs
  long long want_ctas = ((MODE == MODE_STATS ? 4LL * 2 : 8LL * 6) * num_sms_u8() + n - 1) / n;
  long long per = (n_quads + want_ctas - 1) / want_ctas;
  per = std::max<long long>(unit, std::min<long long>(per, 64LL * unit));
  per = (per + unit - 1) / unit * unit;
  a.quads_per_cta = (int)per;
  const int chunks = (int)((n_quads + per - 1) / per);
  for (int done = 0; done < n; done += 65535) {
    a.items = d_items + done;
    dim3 grid(chunks, std::min(n - done, 65535));
    stream_kernel<MODE><<<grid, NT, 0, st>>>(a);
  }
  return check_launch("aug_u8 stream kernel");
}

// items [0, n): one Sharpness step, no gather, quad-aligned shape: the streaming stencil kernel
template <int MODE>
static int launch_stencil(const PassArgs& base, const PassItem* d_items, int n, cudaStream_t st) {
  if (n == 0) return AADG_OK;
  StreamArgs a{};
  a.rows = base.rows; a.luts = base.luts; a.src = base.src; a.scratch = base.scratch;
  a.out_u8 = base.out_u8; a.out_f32 = base.out_f32; a.stats = base.stats; a.H = base.H; a.W = base.W;
  const int wq = base.W / 4;
  const int strips = (wq + 31) / 32;
  // rows per warp-strip: 32 (6 % halo re-reads) when the batch alone fills the GPU, fewer rows (more warps) when it
  // does not: aim at ~6 waves of the 3 resident CTAs per SM (2 waves for a statistics pass, whose per-CTA histogram
  // set-up and flush cost as much as a few hundred pixels per thread), never below 8 rows
  const long long waves = MODE == MODE_STATS ? 2 : 6;
  const long long want_warps = (3LL * waves * num_sms_u8() * (NT / 32) + n - 1) / n;
  long long rb = ((long long)strips * base.H + want_warps - 1) / want_warps;
  rb = std::max<long long>(8, std::min<long long>(rb, 32));
  a.quads_per_cta = (int)rb;                                                         // image rows per band here
  const long long units = (long long)strips * ((base.H + rb - 1) / rb);
  const int chunks = (int)((units + NT / 32 - 1) / (NT / 32));
  for (int done = 0; done < n; done += 65535) {
    a.items = d_items + done;
    dim3 grid(chunks, std::min(n - done, 65535));
    stencil_kernel<MODE><<<grid, NT, 0, st>>>(a);
  }
  return check_launch("aug_u8 stencil kernel");
}

template <int MODE>
static int launch_pass(const PassArgs& base, const PassItem* d_items, int n, cudaStream_t st) {
  if (n == 0) return AADG_OK;
  const int tiles = ((base.W + TW - 1) / TW) * ((base.H + TH - 1) / TH);
  for (int done = 0; done < n; done += 65535) {
    PassArgs a = base;
    a.items = d_items + done;
    dim3 grid((tiles + TILES_PER_CTA - 1) / TILES_PER_CTA, std::min(n - done, 65535));
    pass_kernel<MODE><<<grid, NT, 0, st>>>(a);
  }
  return check_launch("aug_u8 pass kernel");
}

// auxiliary stream (+ fork / join events) per device for the independent rows' chain
struct AuxStream { cudaStream_t stream; cudaEvent_t fork, join; };
static AuxStream* aux_stream() {
  static AuxStream table[64];
  static bool made[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!made[dev]) {
    AuxStream x{};
    if (cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    table[dev] = x;
    made[dev] = true;
  }
  return &table[dev];
}
static bool overlap_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AADG_U8_OVERLAP"); v = e ? atoi(e) : 1; }
  return v != 0;
}

// mode: 0 = uint8 HWC images (+ optional masks), 1 = float32 CHW images (+ optional labels)
static int run(const uint8_t* src_images, const uint8_t* src_masks, const aadg_aug_row_t* rows,
               int n_rows, int n_src, int H, int W, int mode, int dataset, uint8_t* out_u8,
               uint8_t* out_masks, float* out_f32, float* out_labels, void* ws, size_t ws_bytes,
               cudaStream_t st) {
  AADG_REQUIRE(n_rows >= 0 && n_src > 0 && H > 0 && W > 0, "bad sizes n_rows=%d n_src=%d H=%d W=%d",
               n_rows, n_src, H, W);
  if (n_rows == 0) return AADG_OK;
  AADG_REQUIRE(src_images && rows, "null src_images / rows");
  AADG_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  Plan pl;
  int rc = compile(rows, n_rows, n_src, pl);
  if (rc) return rc;
  const Layout L = layout(n_rows, n_src, H, W);
  // scratch is sized for the worst case; the rest of the workspace is always needed
  const size_t need = L.scratch + (size_t)pl.n_scratch * H * W * 3;
  if (ws_bytes < need || !ws) {
    set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    return AADG_ENOSPC;
  }
  char* w = (char*)ws;
  // pointwise-only programs go to the streaming kernel (quad-aligned shapes): put them first in every list
  const bool can_stream = (W % 4 == 0) && (((uintptr_t)src_images & 3) == 0) && (!out_u8 || ((uintptr_t)out_u8 & 3) == 0) &&
                          (!out_f32 || ((uintptr_t)out_f32 & 15) == 0);
  // class 0: pointwise only (stream_kernel), 1: one Sharpness, no gather (stencil_kernel), 2: the tile kernel
  struct Split { int stream, stencil; };
  auto split = [&](std::vector<PassItem>& v) -> Split {
    if (!can_stream) return Split{0, 0};
    auto cls = [](const PassItem& it) { return it.gather ? 2 : (it.sharp >= 0 ? 1 : 0); };
    std::stable_sort(v.begin(), v.end(), [&](const PassItem& x, const PassItem& y) { return cls(x) < cls(y); });
    Split r{0, 0};
    for (const PassItem& it : v) { r.stream += cls(it) == 0; r.stencil += cls(it) == 1; }
    return r;
  };
  const Split ns_src = split(pl.src_stats), ns_fin = split(pl.fin), ns_ind = split(pl.fin_ind);
  Split ns_mat[AADG_MAX_OPS], ns_stat[AADG_MAX_OPS];
  for (int k = 0; k < AADG_MAX_OPS; ++k) { ns_mat[k] = split(pl.mat[k]); ns_stat[k] = split(pl.stat[k]); }
  // one host blob -> one copy: rows, then every launch's item list back to back
  std::vector<PassItem> items;
  std::vector<LutItem> litems;
  size_t o_src = items.size();
  items.insert(items.end(), pl.src_stats.begin(), pl.src_stats.end());
  size_t o_mat[AADG_MAX_OPS + 1], o_stat[AADG_MAX_OPS], o_lut[AADG_MAX_OPS];
  for (int k = 0; k < AADG_MAX_OPS; ++k) {
    o_mat[k] = items.size(); items.insert(items.end(), pl.mat[k].begin(), pl.mat[k].end());
    o_stat[k] = items.size(); items.insert(items.end(), pl.stat[k].begin(), pl.stat[k].end());
    o_lut[k] = litems.size(); litems.insert(litems.end(), pl.lut[k].begin(), pl.lut[k].end());
  }
  const size_t o_fin = items.size();
  items.insert(items.end(), pl.fin.begin(), pl.fin.end());
  const size_t o_ind = items.size();
  items.insert(items.end(), pl.fin_ind.begin(), pl.fin_ind.end());
  const size_t o_lut_ind = litems.size();
  litems.insert(litems.end(), pl.lut_ind.begin(), pl.lut_ind.end());
  AADG_REQUIRE(items.size() <= L.n_items && litems.size() <= L.n_lut_items, "internal: plan overflow");

  AADG_CUDA_TRY(cudaMemcpyAsync(w + L.rows, pl.rows.data(), sizeof(DevRow) * n_rows, cudaMemcpyHostToDevice, st));
  AADG_CUDA_TRY(cudaMemcpyAsync(w + L.items, items.data(), sizeof(PassItem) * items.size(), cudaMemcpyHostToDevice, st));
  if (!litems.empty())
    AADG_CUDA_TRY(cudaMemcpyAsync(w + L.lut_items, litems.data(), sizeof(LutItem) * litems.size(), cudaMemcpyHostToDevice, st));
  if (pl.n_stat_slots)
    AADG_CUDA_TRY(cudaMemsetAsync(w + L.stats, 0, sizeof(Stat) * pl.n_stat_slots, st));
  // the kernels stage a row's whole table block [MAX_OPS][768] in shared memory, including the slots of steps that are
  // not tables (never indexed): keep those bytes defined
  AADG_CUDA_TRY(cudaMemsetAsync(w + L.luts, 0, (size_t)n_rows * AADG_MAX_OPS * 768, st));

  PassArgs a{};
  a.rows = (const DevRow*)(w + L.rows);
  a.luts = (const uint8_t*)(w + L.luts);
  a.src = src_images;
  a.scratch = (const uint8_t*)(w + L.scratch);
  a.stats = (Stat*)(w + L.stats);
  a.H = H; a.W = W;
  a.aligned = ((W * 3) % 16 == 0) && (((uintptr_t)src_images & 15) == 0);
  const PassItem* d_items = (const PassItem*)(w + L.items);
  const LutItem* d_litems = (const LutItem*)(w + L.lut_items);
  AADG_REQUIRE(mode != 0 || out_u8, "null out_u8");
  AADG_REQUIRE(!(out_masks || out_labels) || src_masks, "masks / labels requested without src_masks");
  AADG_REQUIRE(!out_labels || ((((uintptr_t)src_masks & 3) == 0) && ((size_t)H * W) % 4 == 0),
               "label path needs 4-byte aligned masks and H*W %% 4 == 0");

  // a list = [pointwise | stencil | tiled]: up to three launches
#define AADG_U8_LAUNCH(MODE, ARGS, OFF, NS, N, STREAM)                                                      \
  {                                                                                                         \
    rc = launch_stream<MODE>(ARGS, d_items + (OFF), (NS).stream, STREAM);                                   \
    if (rc) return rc;                                                                                      \
    rc = launch_stencil<MODE>(ARGS, d_items + (OFF) + (NS).stream, (NS).stencil, STREAM);                   \
    if (rc) return rc;                                                                                      \
    rc = launch_pass<MODE>(ARGS, d_items + (OFF) + (NS).stream + (NS).stencil, (N) - (NS).stream - (NS).stencil, STREAM); \
    if (rc) return rc;                                                                                      \
  }
  auto final_pass = [&](size_t off, const Split& ns, int n, cudaStream_t s_) -> int {
    if (n == 0) return AADG_OK;
    PassArgs af = a;
    if (mode == 0) {
      af.out_u8 = out_u8;
      AADG_U8_LAUNCH(MODE_U8, af, off, ns, n, s_)
    } else if (out_f32) {
      af.out_f32 = out_f32;
      AADG_U8_LAUNCH(MODE_F32, af, off, ns, n, s_)
    }
    return AADG_OK;
  };

  // Two dependency chains share nothing but the uploaded tables above:
  //   (A) rows with a statistics op or a materialised intermediate: source histograms -> [materialise -> statistics ->
  //       tables] per step -> final pass.  Small, latency-bound launches (shared-memory atomics, a few images each).
  //   (B) INDEPENDENT rows (about half of a search batch): tables -> final pass, plus the masks / labels of every row.
  //       Large, HBM-bound launches.
  // (B) runs on an auxiliary stream forked from and joined back into the caller's stream, so the bandwidth-bound bulk
  // overlaps the latency-bound chain instead of queueing behind it (capture-safe: plain event fork / join).
  const bool have_b = !pl.fin_ind.empty() || out_masks || out_labels;
  const bool have_a = !pl.fin.empty() || !pl.src_stats.empty();
  cudaStream_t sb = st;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (have_a && have_b && overlap_enabled()) {
    AuxStream* ax = aux_stream();
    if (ax) {
      sb = ax->stream; ev_fork = ax->fork; ev_join = ax->join;
      AADG_CUDA_TRY(cudaEventRecord(ev_fork, st));
      AADG_CUDA_TRY(cudaStreamWaitEvent(sb, ev_fork, 0));
    }
  }
  // ---- chain (B) ----
  if (!pl.lut_ind.empty()) {
    lut_kernel<<<(unsigned)pl.lut_ind.size(), 256, 0, sb>>>(a.rows, d_litems + o_lut_ind, (uint8_t*)(w + L.luts), a.stats, H * W);
    rc = check_launch("aug_u8 lut kernel");
    if (rc) return rc;
  }
  rc = final_pass(o_ind, ns_ind, (int)pl.fin_ind.size(), sb);
  if (rc) return rc;
  if (mode == 0 && out_masks) {
    dim3 grid(std::min((H * W + 255) / 256, 1024), n_rows);
    mask_kernel<<<grid, 256, 0, sb>>>(a.rows, src_masks, out_masks, n_rows, H, W);
    rc = check_launch("aug_u8 mask kernel");
    if (rc) return rc;
  }
  if (mode != 0 && out_labels) {
    dim3 grid(std::min((H * W / 4 + 255) / 256, 512), n_rows);
    label_kernel<<<grid, 256, 0, sb>>>(a.rows, src_masks, out_labels, H, W, dataset);
    rc = check_launch("aug_u8 label kernel");
    if (rc) return rc;
  }
  // ---- chain (A) ----
  AADG_U8_LAUNCH(MODE_STATS, a, o_src, ns_src, (int)pl.src_stats.size(), st)
  for (int k = 0; k < AADG_MAX_OPS; ++k) {
    PassArgs am = a;
    am.out_u8 = (uint8_t*)(w + L.scratch);
    AADG_U8_LAUNCH(MODE_U8, am, o_mat[k], ns_mat[k], (int)pl.mat[k].size(), st)
    AADG_U8_LAUNCH(MODE_STATS, a, o_stat[k], ns_stat[k], (int)pl.stat[k].size(), st)
    if (!pl.lut[k].empty()) {
      lut_kernel<<<(unsigned)pl.lut[k].size(), 256, 0, st>>>(a.rows, d_litems + o_lut[k], (uint8_t*)(w + L.luts),
                                                            a.stats, H * W);
      rc = check_launch("aug_u8 lut kernel");
      if (rc) return rc;
    }
  }
  rc = final_pass(o_fin, ns_fin, (int)pl.fin.size(), st);
  if (rc) return rc;
  if (ev_join) {
    AADG_CUDA_TRY(cudaEventRecord(ev_join, sb));
    AADG_CUDA_TRY(cudaStreamWaitEvent(st, ev_join, 0));
  }
  return rc;
}

}  // namespace u8
}  // namespace aadg

extern "C" {

size_t aadg_u8_workspace_bytes(int n_rows, int n_src, int height, int width) {
  if (n_rows <= 0 || n_src <= 0 || height <= 0 || width <= 0) return 0;
  return aadg::u8::layout(n_rows, n_src, height, width).total;
}

int aadg_u8_apply_policy(const uint8_t* src_images, const uint8_t* src_masks,
                         const aadg_aug_row_t* rows, int n_rows, int n_src, int height, int width,
                         uint8_t* out_u8, uint8_t* out_masks, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return aadg::u8::run(src_images, src_masks, rows, n_rows, n_src, height, width, 0, 0, out_u8,
                       out_masks, nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

int aadg_u8_policy_normalize(const uint8_t* src_images, const uint8_t* src_masks,
                             const aadg_aug_row_t* rows, int n_rows, int n_src, int height,
                             int width, int dataset, float* out_images, float* out_labels,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (dataset != AADG_DATASET_OPTIC && dataset != AADG_DATASET_VESSEL) {
    aadg::set_error("unknown dataset %d", dataset);
    return AADG_EINVAL;
  }
  return aadg::u8::run(src_images, src_masks, rows, n_rows, n_src, height, width, 1, dataset,
                       nullptr, nullptr, out_images, out_labels, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

}  // extern "C"
