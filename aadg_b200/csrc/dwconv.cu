// Depthwise 3x3 convolution (stride 1, "same" padding = dilation): forward, data gradient and weight
// gradient on TMA-staged shared-memory tiles (sm_100a).  These are the depthwise halves of smp's
// SeparableConv2d in the ASPP branches (dilation 12/24/36 on the stride-16 map) and in the decoder
// (dilation 1 on the stride-4 map) behind `model(input)` / `seg_loss.backward()` (search_dg.py:132,171).
//
// HBM-bound: every input element should cross HBM once.  One 4-D tensor-map TMA box brings a
// (channels x pixels x rows) tile -- halo included, image borders zero-filled by the TMA -- into shared
// memory in a single bulk copy, so the SM has the whole tile in flight with one instruction and several
// resident CTAs overlap their copy and compute phases.
//   ROLL   (dilation 1): tile = 64 channels x 32 x 8 output pixels (+1 halo).  A thread owns one
//          (column, 8-channel group) and walks down the rows with three rolling partial sums, so each
//          staged element is read from shared memory three times instead of nine.
//   DIRECT (dilation > 1, whole image <= 1024 pixels): tile = 32 channels x the whole image, no halo;
//          taps that leave the image are skipped.
// Anything else falls back to the generic kernel in nn_elem.cu.
// Weight gradient: the same tiles, persistent CTAs (a channel chunk x a strided set of tiles) keeping the
// 9 x 8 partial sums in registers, one shared-memory + one global atomic reduction per CTA.
#include "tc_common.cuh"

#include <algorithm>

namespace aadg {
namespace nn {
int dwconv3x3_generic(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int direction,
                      void* y, int ldy, cudaStream_t st);
int dwconv3x3_wgrad_generic(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int lddy, int dil,
                            float* dw, cudaStream_t st);
int dwconv3x3_strided(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int stride, int direction,
                      void* y, int ho, int wo, int ldy, cudaStream_t st);
int dwconv3x3_strided_wgrad(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int ho, int wo, int lddy,
                            int dil, int stride, float* dw, cudaStream_t st);
}  // namespace nn

namespace dw {

typedef __nv_bfloat16 bf16;

constexpr int R_TW = 32, R_TH = 8, R_CB = 64;
constexpr int R_BW = R_TW + 2, R_BH = R_TH + 2;
constexpr int R_TILE_BYTES = R_BW * R_BH * R_CB * 2;     // 43 520
constexpr int D_CB = 32, D_MAX_PIX = 1024;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
  return u;
}
__device__ __forceinline__ void load_w8(const float* w, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(w)), b = __ldg(reinterpret_cast<const float4*>(w + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ uint8_t* align128(uint8_t* p) { return align_smem(p, 128u); }

struct Args {
  int N, H, W, C;
  int tiles_x, tiles_y;
  int dil;
  const float* w;       // [9][C] fp32
  int flip;             // data gradient: filter rotated by 180 degrees
  int accumulate;       // y += result (gradient fan-in of parallel branches)
  bf16* y; int ldy;     // output (forward / data gradient)
  const bf16* dy; int lddy;   // weight gradient: output gradient
  float* dwgt;          // weight gradient accumulator [9][C]
};

// ---- dilation 1, forward / data gradient ---------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) roll_kernel(const __grid_constant__ CUtensorMap tmX, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tile = align128(smem_raw);
  uint64_t* bar = (uint64_t*)(tile + R_TILE_BYTES);
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int x0 = tx * R_TW, y0 = ty * R_TH, c0 = blockIdx.y * R_CB, n = blockIdx.z;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, R_TILE_BYTES);
    tc::tma_load_4d(tile, &tmX, bar, c0, x0 - 1, y0 - 1, n);
  }
  const int g = threadIdx.x & 7, px = threadIdx.x >> 3;
  const int c = c0 + g * 8, ox = x0 + px;
  float wt[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (c < a.C) load_w8(a.w + (size_t)(a.flip ? 8 - t : t) * a.C + c, wt[t]);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) wt[t][i] = 0.f;
    }
  }
  mbar_wait(bar, 0);
  const bool col_ok = c < a.C && ox < a.W;
  const uint4* col = reinterpret_cast<const uint4*>(tile) + px * 8 + g;    // box row stride = R_BW * 8 uint4
  float mid[8], old[8];
#pragma unroll
  for (int i = 0; i < R_BH; ++i) {
    float xl[8], xc[8], xr[8];
    unpack8(col[(i * R_BW + 0) * 8], xl);
    unpack8(col[(i * R_BW + 1) * 8], xc);
    unpack8(col[(i * R_BW + 2) * 8], xr);
    if (i >= 2) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        old[e] = fmaf(wt[8][e], xr[e], fmaf(wt[7][e], xc[e], fmaf(wt[6][e], xl[e], old[e])));
      const int oy = y0 + i - 2;
      if (col_ok && oy < a.H) {
        uint4* dst = reinterpret_cast<uint4*>(a.y + (((size_t)n * a.H + oy) * a.W + ox) * a.ldy + c);
        if (a.accumulate) {
          float prev[8];
          unpack8(*dst, prev);
#pragma unroll
          for (int e = 0; e < 8; ++e) old[e] += prev[e];
        }
        *dst = pack8(old);
      }
    }
    if (i >= 1 && i <= R_TH) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        old[e] = fmaf(wt[5][e], xr[e], fmaf(wt[4][e], xc[e], fmaf(wt[3][e], xl[e], mid[e])));
    }
    if (i < R_TH) {
#pragma unroll
      for (int e = 0; e < 8; ++e) mid[e] = fmaf(wt[2][e], xr[e], fmaf(wt[1][e], xc[e], wt[0][e] * xl[e]));
    }
  }
}

// ---- dilation 1, weight gradient ---------------------------------------------------------------------------
// grid (channel chunks, workers), one CTA per SM; a worker walks tiles w, w + workers, ... of its chunk with the
// x tile (halo included) and the dy tile of the NEXT step already in flight (two shared-memory stages), so the
// threads only read shared memory and accumulate.  Zero fill outside the image / beyond C makes bounds checks
// unnecessary: those products are zero.
constexpr int R_DY_BYTES = R_TW * R_TH * R_CB * 2;        // 32 768
constexpr int R_WG_STAGE = R_TILE_BYTES + R_DY_BYTES;
__global__ void __launch_bounds__(256, 1) roll_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,
                                                            const __grid_constant__ CUtensorMap tmDY, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stages = align128(smem_raw);
  uint64_t* bar = (uint64_t*)(stages + 2 * R_WG_STAGE);
  float* red = (float*)(bar + 2);                       // [9][R_CB]
  const int c0 = blockIdx.x * R_CB;
  const int per_img = a.tiles_x * a.tiles_y, n_tiles = per_img * a.N;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 9 * R_CB; i += 256) red[i] = 0.f;
  __syncthreads();
  auto issue = [&](int t, int s) {
    const int n = t / per_img, r = t - n * per_img;
    const int x0 = (r % a.tiles_x) * R_TW, y0 = (r / a.tiles_x) * R_TH;
    uint8_t* dst = stages + s * R_WG_STAGE;
    mbar_arrive_expect_tx(&bar[s], R_WG_STAGE);
    tc::tma_load_4d(dst, &tmX, &bar[s], c0, x0 - 1, y0 - 1, n);
    tc::tma_load_4d(dst + R_TILE_BYTES, &tmDY, &bar[s], c0, x0, y0, n);
  };
  const int g = threadIdx.x & 7, px = threadIdx.x >> 3;
  float acc[9][8] = {};
  if (threadIdx.x == 0 && (int)blockIdx.y < n_tiles) issue(blockIdx.y, 0);
  int it = 0;
  for (int t = blockIdx.y; t < n_tiles; t += gridDim.y, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && t + (int)gridDim.y < n_tiles) issue(t + gridDim.y, s ^ 1);
    mbar_wait(&bar[s], (it >> 1) & 1);
    const uint4* col = reinterpret_cast<const uint4*>(stages + s * R_WG_STAGE) + px * 8 + g;
    const uint4* dcol = reinterpret_cast<const uint4*>(stages + s * R_WG_STAGE + R_TILE_BYTES) + px * 8 + g;
#pragma unroll
    for (int i = 0; i < R_BH; ++i) {
      float xl[8], xc[8], xr[8];
      unpack8(col[(i * R_BW + 0) * 8], xl);
      unpack8(col[(i * R_BW + 1) * 8], xc);
      unpack8(col[(i * R_BW + 2) * 8], xr);
      // box row i is filter row 0 of output row i, filter row 1 of output row i-1, filter row 2 of output row i-2
#pragma unroll
      for (int fr = 0; fr < 3; ++fr) {
        const int orow = i - fr;
        if (orow >= 0 && orow < R_TH) {
          float d[8];
          unpack8(dcol[orow * R_TW * 8], d);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[fr * 3 + 0][e] = fmaf(d[e], xl[e], acc[fr * 3 + 0][e]);
            acc[fr * 3 + 1][e] = fmaf(d[e], xc[e], acc[fr * 3 + 1][e]);
            acc[fr * 3 + 2][e] = fmaf(d[e], xr[e], acc[fr * 3 + 2][e]);
          }
        }
      }
    }
    __syncthreads();     // every thread is done with this stage before the copy after next overwrites it
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&red[t * R_CB + g * 8 + e], acc[t][e]);
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * R_CB; i += 256) {
    const int t = i / R_CB, cc = c0 + i % R_CB;
    if (cc < a.C) atomicAdd(&a.dwgt[(size_t)t * a.C + cc], red[i]);
  }
}

// ---- dilation > 1 on a small map: the whole image of a 32-channel chunk in shared memory ---------------------------
__global__ void __launch_bounds__(256) direct_kernel(const __grid_constant__ CUtensorMap tmX, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tile = align128(smem_raw);
  const int HW = a.H * a.W;
  uint64_t* bar = (uint64_t*)(tile + (size_t)HW * D_CB * 2);
  const int c0 = blockIdx.x * D_CB, n = blockIdx.y;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)(HW * D_CB * 2));
    tc::tma_load_4d(tile, &tmX, bar, c0, 0, 0, n);
  }
  const int g = threadIdx.x & 3;
  const int c = c0 + g * 8;
  float wt[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (c < a.C) load_w8(a.w + (size_t)(a.flip ? 8 - t : t) * a.C + c, wt[t]);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) wt[t][i] = 0.f;
    }
  }
  mbar_wait(bar, 0);
  const uint4* img = reinterpret_cast<const uint4*>(tile) + g;             // pixel stride = 4 uint4
  for (int p = threadIdx.x >> 2; p < HW; p += 64) {
    const int oy = p / a.W, ox = p - oy * a.W;
    float acc[8] = {};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy + (r - 1) * a.dil;
      if (iy < 0 || iy >= a.H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = ox + (s - 1) * a.dil;
        if (ix < 0 || ix >= a.W) continue;
        float v[8];
        unpack8(img[(iy * a.W + ix) * 4], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt[r * 3 + s][e], v[e], acc[e]);
      }
    }
    if (c < a.C) {
      uint4* dst = reinterpret_cast<uint4*>(a.y + ((size_t)n * HW + p) * a.ldy + c);
      if (a.accumulate) {
        float prev[8];
        unpack8(*dst, prev);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += prev[e];
      }
      *dst = pack8(acc);
    }
  }
}

// grid (channel chunks, workers); a worker walks images w, w + workers, ...
__global__ void __launch_bounds__(256) direct_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tile = align128(smem_raw);
  const int HW = a.H * a.W;
  uint64_t* bar = (uint64_t*)(tile + (size_t)HW * D_CB * 2);
  float* red = (float*)(bar + 2);                       // [9][D_CB]
  const int c0 = blockIdx.x * D_CB;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 9 * D_CB; i += 256) red[i] = 0.f;
  __syncthreads();
  const int g = threadIdx.x & 3;
  const int c = c0 + g * 8;
  const uint4* img = reinterpret_cast<const uint4*>(tile) + g;
  float acc[9][8] = {};
  uint32_t phase = 0;
  for (int n = blockIdx.y; n < a.N; n += gridDim.y) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, (uint32_t)(HW * D_CB * 2));
      tc::tma_load_4d(tile, &tmX, bar, c0, 0, 0, n);
    }
    const bf16* dyn = a.dy + (size_t)n * HW * a.lddy + c;
    int p = threadIdx.x >> 2;
    uint4 dnext = make_uint4(0, 0, 0, 0);
    if (p < HW && c < a.C) dnext = __ldg(reinterpret_cast<const uint4*>(dyn + (size_t)p * a.lddy));
    mbar_wait(bar, phase);
    phase ^= 1;
    for (; p < HW; p += 64) {
      float d[8];
      unpack8(dnext, d);
      if (p + 64 < HW && c < a.C) dnext = __ldg(reinterpret_cast<const uint4*>(dyn + (size_t)(p + 64) * a.lddy));
      const int oy = p / a.W, ox = p - oy * a.W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = oy + (r - 1) * a.dil;
        if (iy < 0 || iy >= a.H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ix = ox + (s - 1) * a.dil;
          if (ix < 0 || ix >= a.W) continue;
          float v[8];
          unpack8(img[(iy * a.W + ix) * 4], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[r * 3 + s][e] = fmaf(d[e], v[e], acc[r * 3 + s][e]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&red[t * D_CB + g * 8 + e], acc[t][e]);
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * D_CB; i += 256) {
    const int t = i / D_CB, cc = c0 + i % D_CB;
    if (cc < a.C) atomicAdd(&a.dwgt[(size_t)t * a.C + cc], red[i]);
  }
}

// ---- dilation > 1, persistent + pipelined (round 2) ----------------------------------------------------------------
// direct_kernel / direct_wgrad_kernel serialise copy -> compute per CTA and rely on three co-resident CTAs to overlap
// them: 0.53 ms per launch on the 144 x 32 x 32 x 2048 ASPP tensor against 0.19 ms of HBM time.  Here ONE CTA of 512
// threads per SM walks a contiguous range of (channel chunk, image) items with a ring of `stages` whole-image tiles:
// the TMA copies of the next items are in flight while the current one is computed, the filter taps are re-read only
// when the chunk changes (items are ordered image-fastest), and the weight gradient keeps its 9 x 8 partial sums in
// registers across all images of a chunk.
constexpr int DP_THREADS = 512;

// item k of CTA b is item number b + k * gridDim.x, channel chunk FASTEST (chunk = item % chunks, image = item / chunks):
// at any moment the CTAs of the grid read adjacent 64-byte channel slices of the same pixel rows, i.e. whole DRAM pages
// (a contiguous range of images per CTA scattered every CTA's 64-byte reads over its own pages: 2x slower).  The host
// sizes the grid as a multiple of `chunks` when it can, so a CTA keeps its chunk (taps loaded once, one flush).

__global__ void __launch_bounds__(DP_THREADS, 1) direct_p_kernel(const __grid_constant__ CUtensorMap tmX, const Args a,
                                                                 int chunks, int stages, int tile_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = align128(smem_raw);
  uint64_t* full = (uint64_t*)(tiles + (size_t)stages * tile_bytes);
  const int HW = a.H * a.W;
  const long long total = (long long)chunks * a.N;
  const long long G = gridDim.x;
  if ((long long)blockIdx.x >= total) return;
  const int n_mine = (int)((total - blockIdx.x + G - 1) / G);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int k, int s) {
    const long long item = (long long)blockIdx.x + (long long)k * G;
    const int n = (int)(item / chunks), chunk = (int)(item - (long long)n * chunks);
    mbar_arrive_expect_tx(&full[s], (uint32_t)(HW * D_CB * 2));
    tc::tma_load_4d(tiles + (size_t)s * tile_bytes, &tmX, &full[s], chunk * D_CB, 0, 0, n);
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < stages && s < n_mine; ++s) issue(s, s);
  const int g = threadIdx.x & 3;
  int cur_chunk = -1;
  float wt[9][8];
  for (int k = 0; k < n_mine; ++k) {
    const int s = k % stages;
    const uint32_t phase = (uint32_t)(k / stages) & 1u;
    const long long item = (long long)blockIdx.x + (long long)k * G;
    const int n = (int)(item / chunks), chunk = (int)(item - (long long)n * chunks);
    const int c = chunk * D_CB + g * 8;
    if (chunk != cur_chunk) {
      cur_chunk = chunk;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (c < a.C) load_w8(a.w + (size_t)(a.flip ? 8 - t : t) * a.C + c, wt[t]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) wt[t][i] = 0.f;
        }
      }
    }
    mbar_wait(&full[s], phase);
    const uint4* img = reinterpret_cast<const uint4*>(tiles + (size_t)s * tile_bytes) + g;     // pixel stride = 4 uint4
    for (int p = threadIdx.x >> 2; p < HW; p += DP_THREADS / 4) {
      const int oy = p / a.W, ox = p - oy * a.W;
      float acc[8] = {};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = oy + (r - 1) * a.dil;
        if (iy < 0 || iy >= a.H) continue;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int ix = ox + (q - 1) * a.dil;
          if (ix < 0 || ix >= a.W) continue;
          float v[8];
          unpack8(img[(iy * a.W + ix) * 4], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt[r * 3 + q][e], v[e], acc[e]);
        }
      }
      if (c < a.C) {
        uint4* dst = reinterpret_cast<uint4*>(a.y + ((size_t)n * HW + p) * a.ldy + c);
        if (a.accumulate) {
          float prev[8];
          unpack8(*dst, prev);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += prev[e];
        }
        *dst = pack8(acc);
      }
    }
    fence_proxy_async();      // the generic reads of this stage are ordered before the bulk copy that refills it
    __syncthreads();
    if (threadIdx.x == 0 && k + stages < n_mine) issue(k + stages, s);
  }
}

__global__ void __launch_bounds__(DP_THREADS, 1) direct_p_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const Args a,
                                                                       int chunks, int stages, int tile_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = align128(smem_raw);
  uint64_t* full = (uint64_t*)(tiles + (size_t)stages * tile_bytes);
  float* red = (float*)(full + stages);                   // [9][D_CB]
  const int HW = a.H * a.W;
  const long long total = (long long)chunks * a.N;
  const long long G = gridDim.x;
  if ((long long)blockIdx.x >= total) return;
  const int n_mine = (int)((total - blockIdx.x + G - 1) / G);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 9 * D_CB; i += DP_THREADS) red[i] = 0.f;
  __syncthreads();
  auto issue = [&](int k, int s) {
    const long long item = (long long)blockIdx.x + (long long)k * G;
    const int n = (int)(item / chunks), chunk = (int)(item - (long long)n * chunks);
    mbar_arrive_expect_tx(&full[s], (uint32_t)(HW * D_CB * 2));
    tc::tma_load_4d(tiles + (size_t)s * tile_bytes, &tmX, &full[s], chunk * D_CB, 0, 0, n);
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < stages && s < n_mine; ++s) issue(s, s);
  const int g = threadIdx.x & 3;
  float acc[9][8] = {};
  // flush the partial sums of channel chunk `chunk`: registers -> shared (atomics) -> one global atomic per entry
  auto flush = [&](int chunk) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) { atomicAdd(&red[t * D_CB + g * 8 + e], acc[t][e]); acc[t][e] = 0.f; }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * D_CB; i += DP_THREADS) {
      const int t = i / D_CB, cc = chunk * D_CB + i % D_CB;
      if (cc < a.C) atomicAdd(&a.dwgt[(size_t)t * a.C + cc], red[i]);
      red[i] = 0.f;
    }
    __syncthreads();
  };
  int cur_chunk = (int)(blockIdx.x % chunks);
  for (int k = 0; k < n_mine; ++k) {
    const int s = k % stages;
    const uint32_t phase = (uint32_t)(k / stages) & 1u;
    const long long item = (long long)blockIdx.x + (long long)k * G;
    const int n = (int)(item / chunks), chunk = (int)(item - (long long)n * chunks);
    if (chunk != cur_chunk) { flush(cur_chunk); cur_chunk = chunk; }
    const int c = chunk * D_CB + g * 8;
    const bf16* dyn = a.dy + (size_t)n * HW * a.lddy + c;
    int p = threadIdx.x >> 2;
    uint4 dnext = make_uint4(0, 0, 0, 0);
    if (p < HW && c < a.C) dnext = __ldg(reinterpret_cast<const uint4*>(dyn + (size_t)p * a.lddy));
    mbar_wait(&full[s], phase);
    const uint4* img = reinterpret_cast<const uint4*>(tiles + (size_t)s * tile_bytes) + g;
    for (; p < HW; p += DP_THREADS / 4) {
      float d[8];
      unpack8(dnext, d);
      if (p + DP_THREADS / 4 < HW && c < a.C)
        dnext = __ldg(reinterpret_cast<const uint4*>(dyn + (size_t)(p + DP_THREADS / 4) * a.lddy));
      const int oy = p / a.W, ox = p - oy * a.W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = oy + (r - 1) * a.dil;
        if (iy < 0 || iy >= a.H) continue;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int ix = ox + (q - 1) * a.dil;
          if (ix < 0 || ix >= a.W) continue;
          float v[8];
          unpack8(img[(iy * a.W + ix) * 4], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[r * 3 + q][e] = fmaf(d[e], v[e], acc[r * 3 + q][e]);
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0 && k + stages < n_mine) issue(k + stages, s);
  }
  flush(cur_chunk);
}

static int tuning_direct_p() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AADG_DW_DIRECT_P"); v = e ? atoi(e) : 1; }
  return v;
}
// stages x whole-image tiles within ~200 KB of shared memory (at least 2, at most 6)
static int direct_p_stages(int tile_bytes) { return std::max(2, std::min(6, (200 * 1024) / tile_bytes)); }

static int num_sms() {
  static int v = 0;
  if (!v) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}
static int tuning_tma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AADG_DW_TMA"); v = e ? atoi(e) : 1; }
  return v;
}

// tensor map over x [N,H,W,ldx] (C channels used), plain row-major box (cb, bw, bh, 1)
static int make_x_map(CUtensorMap* m, const void* x, int n, int h, int w, int c, int ldx, int cb, int bw, int bh) {
  const long long dims[4] = {c, w, h, n};
  const long long strides[3] = {ldx, (long long)w * ldx, (long long)h * w * ldx};
  const int box[4] = {cb, bw, bh, 1};
  return tc::make_map_bf16(m, x, 4, dims, strides, box, nullptr, false);
}

template <class Kern>
static int set_smem(Kern k, int bytes) {
  AADG_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return AADG_OK;
}

}  // namespace dw
}  // namespace aadg

using namespace aadg;
using namespace aadg::dw;

#define DW_REQ_C(C) AADG_REQUIRE((C) > 0 && (C) % 8 == 0 && (C) <= 2048, "channels %d must be a multiple of 8 and <= 2048", (C))

extern "C" {

/* depthwise 3x3, stride 1, padding = dilation. direction bit 0: 0 = forward, 1 = data gradient; bit 1 (2): y += result.
 * w fp32 [9][c] */
int aadg_dwconv3x3(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int direction, void* y,
                   int ldy, void* stream) {
  DW_REQ_C(c);
  AADG_REQUIRE(h > 0 && w > 0 && n > 0 && dil >= 1, "bad depthwise geometry");
  AADG_REQUIRE(h <= 65535 && n <= 65535, "image too tall / batch too large for the depthwise grid");
  AADG_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0,
               "depthwise tensors must be 16-byte aligned with channel strides that are multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  Args a{};
  a.N = n; a.H = h; a.W = w; a.C = c; a.dil = dil; a.w = wgt; a.flip = (direction & 1) ? 1 : 0;
  a.accumulate = (direction & 2) ? 1 : 0;
  a.y = (bf16*)y; a.ldy = ldy;
  if (tuning_tma() && dil == 1) {
    CUtensorMap m;
    int rc = make_x_map(&m, x, n, h, w, c, ldx, R_CB, R_BW, R_BH);
    if (rc) return rc;
    a.tiles_x = (w + R_TW - 1) / R_TW; a.tiles_y = (h + R_TH - 1) / R_TH;
    const int smem = R_TILE_BYTES + 128 + 64;
    static bool set = false;
    if (!set) { rc = set_smem(roll_kernel, smem); if (rc) return rc; set = true; }
    dim3 grid(a.tiles_x * a.tiles_y, (c + R_CB - 1) / R_CB, n);
    roll_kernel<<<grid, 256, smem, st>>>(m, a);
    return check_launch("dwconv3x3 roll");
  }
  if (tuning_tma() && h * w <= D_MAX_PIX && h <= 256 && w <= 256) {
    CUtensorMap m;
    int rc = make_x_map(&m, x, n, h, w, c, ldx, D_CB, w, h);
    if (rc) return rc;
    if (tuning_direct_p()) {
      const int tile_bytes = h * w * D_CB * 2, stages = direct_p_stages(tile_bytes);
      const int chunks = (c + D_CB - 1) / D_CB;
      const int smem_p = stages * tile_bytes + 128 + 8 * stages + 64;
      static bool set_p = false;
      if (!set_p) { rc = set_smem(direct_p_kernel, 204 * 1024); if (rc) return rc; set_p = true; }
      const long long total = (long long)chunks * n;
      long long g = std::min<long long>(total, num_sms());
      if (chunks <= g) g = g / chunks * chunks;           // a CTA keeps its channel chunk
      dim3 grid((unsigned)std::max<long long>(1, g));
      direct_p_kernel<<<grid, DP_THREADS, smem_p, st>>>(m, a, chunks, stages, tile_bytes);
      return check_launch("dwconv3x3 direct (persistent)");
    }
    const int smem = h * w * D_CB * 2 + 128 + 64;
    static bool set = false;
    if (!set) { rc = set_smem(direct_kernel, D_MAX_PIX * D_CB * 2 + 128 + 64); if (rc) return rc; set = true; }
    dim3 grid((c + D_CB - 1) / D_CB, n);
    direct_kernel<<<grid, 256, smem, st>>>(m, a);
    return check_launch("dwconv3x3 direct");
  }
  return nn::dwconv3x3_generic(x, n, h, w, c, ldx, wgt, dil, direction, y, ldy, st);
}

/* dw[t][c] += sum over pixels of dy[p][c] * x[p + tap t][c] (fp32, accumulated: zero dw first for a fresh gradient) */
int aadg_dwconv3x3_wgrad(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int lddy, int dil, float* dwgt,
                         void* stream) {
  DW_REQ_C(c);
  AADG_REQUIRE(h > 0 && w > 0 && n > 0 && dil >= 1, "bad depthwise geometry");
  AADG_REQUIRE((long long)n * h * w < (1ll << 31), "too many pixels");
  AADG_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0,
               "depthwise tensors must be 16-byte aligned with channel strides that are multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  Args a{};
  a.N = n; a.H = h; a.W = w; a.C = c; a.dil = dil; a.dy = (const bf16*)dy; a.lddy = lddy; a.dwgt = dwgt;
  if (tuning_tma() && dil == 1) {
    CUtensorMap m;
    int rc = make_x_map(&m, x, n, h, w, c, ldx, R_CB, R_BW, R_BH);
    if (rc) return rc;
    CUtensorMap mdy;
    rc = make_x_map(&mdy, dy, n, h, w, c, lddy, R_CB, R_TW, R_TH);
    if (rc) return rc;
    a.tiles_x = (w + R_TW - 1) / R_TW; a.tiles_y = (h + R_TH - 1) / R_TH;
    const int smem = 2 * R_WG_STAGE + 128 + 64 + 9 * R_CB * 4;
    static bool set = false;
    if (!set) { rc = set_smem(roll_wgrad_kernel, smem); if (rc) return rc; set = true; }
    const int chunks = (c + R_CB - 1) / R_CB;
    const long long n_tiles = (long long)a.tiles_x * a.tiles_y * n;
    const int workers = (int)std::max<long long>(1, std::min<long long>(n_tiles, std::max(1, num_sms() / chunks)));
    dim3 grid(chunks, workers);
    roll_wgrad_kernel<<<grid, 256, smem, st>>>(m, mdy, a);
    return check_launch("dwconv3x3 roll wgrad");
  }
  if (tuning_tma() && h * w <= D_MAX_PIX && h <= 256 && w <= 256) {
    CUtensorMap m;
    int rc = make_x_map(&m, x, n, h, w, c, ldx, D_CB, w, h);
    if (rc) return rc;
    if (tuning_direct_p()) {
      const int tile_bytes = h * w * D_CB * 2, stages = direct_p_stages(tile_bytes);
      const int chunks = (c + D_CB - 1) / D_CB;
      const int smem_p = stages * tile_bytes + 128 + 8 * stages + 9 * D_CB * 4 + 64;
      static bool set_p = false;
      if (!set_p) { rc = set_smem(direct_p_wgrad_kernel, 204 * 1024); if (rc) return rc; set_p = true; }
      const long long total = (long long)chunks * n;
      long long g = std::min<long long>(total, num_sms());
      if (chunks <= g) g = g / chunks * chunks;           // a CTA keeps its channel chunk: one flush
      dim3 grid((unsigned)std::max<long long>(1, g));
      direct_p_wgrad_kernel<<<grid, DP_THREADS, smem_p, st>>>(m, a, chunks, stages, tile_bytes);
      return check_launch("dwconv3x3 direct wgrad (persistent)");
    }
    const int smem = h * w * D_CB * 2 + 128 + 64 + 9 * D_CB * 4;
    static bool set = false;
    if (!set) { rc = set_smem(direct_wgrad_kernel, D_MAX_PIX * D_CB * 2 + 128 + 64 + 9 * D_CB * 4); if (rc) return rc; set = true; }
    const int chunks = (c + D_CB - 1) / D_CB;
    const int workers = std::max(1, std::min(n, (num_sms() * 3 + chunks - 1) / chunks));
    dim3 grid(chunks, workers);
    direct_wgrad_kernel<<<grid, 256, smem, st>>>(m, a);
    return check_launch("dwconv3x3 direct wgrad");
  }
  return nn::dwconv3x3_wgrad_generic(x, n, h, w, c, ldx, dy, lddy, dil, dwgt, st);
}

/* strided depthwise 3x3 (padding = dilation, output ho x wo = floor((h-1)/stride)+1 ...): the down-sampling blocks of
 * MobileNetV2.  direction 0: y [n,ho,wo] <- x [n,h,w]; direction 1 (data gradient): `x` is dy [n,ho,wo] with stride
 * ldx, `y` is dx [n,h,w] with stride ldy.  stride 1 is forwarded to aadg_dwconv3x3. */
int aadg_dwconv3x3_strided(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int stride,
                           int direction, void* y, int ho, int wo, int ldy, void* stream) {
  DW_REQ_C(c);
  AADG_REQUIRE(stride >= 1 && dil >= 1 && h > 0 && w > 0 && n > 0, "bad depthwise geometry");
  AADG_REQUIRE(ho == (h - 1) / stride + 1 && wo == (w - 1) / stride + 1, "output size mismatch: expected %dx%d", (h - 1) / stride + 1,
               (w - 1) / stride + 1);
  if (stride == 1) return aadg_dwconv3x3(x, n, h, w, c, ldx, wgt, dil, direction, y, ldy, stream);
  AADG_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0,
               "depthwise tensors must be 16-byte aligned with channel strides that are multiples of 8");
  return nn::dwconv3x3_strided(x, n, h, w, c, ldx, wgt, dil, stride, direction, y, ho, wo, ldy, (cudaStream_t)stream);
}
int aadg_dwconv3x3_strided_wgrad(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int ho, int wo, int lddy,
                                 int dil, int stride, float* dwgt, void* stream) {
  DW_REQ_C(c);
  AADG_REQUIRE(stride >= 1 && dil >= 1 && h > 0 && w > 0 && n > 0, "bad depthwise geometry");
  AADG_REQUIRE(ho == (h - 1) / stride + 1 && wo == (w - 1) / stride + 1, "output size mismatch");
  if (stride == 1) return aadg_dwconv3x3_wgrad(x, n, h, w, c, ldx, dy, lddy, dil, dwgt, stream);
  AADG_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0,
               "depthwise tensors must be 16-byte aligned with channel strides that are multiples of 8");
  return nn::dwconv3x3_strided_wgrad(x, n, h, w, c, ldx, dy, ho, wo, lddy, dil, stride, dwgt, (cudaStream_t)stream);
}

}  // extern "C"
