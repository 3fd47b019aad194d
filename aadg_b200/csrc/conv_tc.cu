// Convolution forward / data-gradient / weight-gradient as implicit GEMMs on tcgen05 + TMEM, operands
// staged by tensor-map TMA (sm_100a only).
//
// Replaces the cuDNN convolutions behind `model(input)` / `seg_loss.backward()` of the reference
// (search_dg.py:132,171; models/__init__.py:17-23 -> segmentation_models_pytorch DeepLabV3+/UNet).
//
// Data layout: activations bf16 NHWC (channel stride `ld` may exceed the channel count so that a
// tensor can be a channel slice of a concat buffer), weights bf16 [tap][Cout][Cin] (forward / wgrad)
// and [tap][Cin][Cout] (data gradient), fp32 accumulation in tensor memory.
//
//   fprop / dgrad ("igemm"): D[128 pixels, BN channels] = sum over taps and 64-channel blocks of
//     A_tap[128 px, 64 ch] * W_tap[BN, 64 ch]^T.  The A tile of a tap is ONE 4-D TMA box
//     (64 ch, TW, TH, TN) of the NHWC input shifted by the tap offset; image borders are the TMA's
//     zero fill, strided convolutions use the box's element strides, dilation is just the offset.
//     Both operands are K-major with the 128-byte swizzle.  Warp 0 = TMA producer, warp 1 = MMA
//     issuer (one elected thread), warps 2-5 = epilogue (TMEM -> registers -> bf16 NHWC stores).
//   wgrad: D[128 Cout, BN Cin] = sum over pixels dY[px, Cout]^T X[px + tap, Cin]: the pixel axis is K,
//     so both operands are MN-major tiles (64-channel x 64-pixel boxes); split over pixel ranges
//     across CTAs, fp32 `red.add` into dW.
#include "tc_common.cuh"

#include <stdlib.h>
#include <algorithm>
#include <mutex>

namespace aadg {
namespace tc {

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int make_map_bf16(CUtensorMap* map, const void* base, int rank, const long long* dims,
                  const long long* strides_elems, const int* box, const int* elem_strides, bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return AADG_ECUDA; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    es[i] = (cuuint32_t)(elem_strides ? elem_strides[i] : 1);
    if (i > 0) gs[i - 1] = (cuuint64_t)strides_elems[i - 1] * 2;
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %lld %lld %lld %lld box %d %d %d %d", (int)r, rank,
              dims[0], rank > 1 ? dims[1] : 0, rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0, box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return AADG_ECUDA;
  }
  return AADG_OK;
}

constexpr int MAX_TAPS = 49;
struct Taps {
  int n;
  short dy[MAX_TAPS], dx[MAX_TAPS], w[MAX_TAPS];
};

struct IgemmArgs {
  int lg_tw, lg_th;             // log2 tile extents in x, y (the rest of the 128 rows spans images)
  int tiles_x, tiles_y, tiles_n;
  int Wsub, Hsub, Nimg;         // logical output grid of this launch
  int in_step;                  // input coordinate = output coordinate * in_step + tap offset
  int k_blocks;                 // ceil(Cin / 64)
  __nv_bfloat16* out;
  int ldc, c_off, Cout;         // output channel stride / offset / valid channel count
  int Hout, Wout;               // physical output image size
  int o_step, o_y0, o_x0;       // physical pixel = logical pixel * o_step + (o_y0, o_x0)
  int accumulate;               // out += result
  Taps taps;
};

constexpr int IG_THREADS = 192;
constexpr int A_BYTES = 128 * 128;      // 128 pixel rows x 64 bf16

template <int BN, int STAGES>
constexpr int igemm_smem_bytes() { return STAGES * (A_BYTES + BN * 128) + 1024 + 256; }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(IG_THREADS) igemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const IgemmArgs a) {
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { prefetch_map(&tmA); prefetch_map(&tmB); }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> logical output origin
  int t = blockIdx.x;
  const int tile_x = t % a.tiles_x; t /= a.tiles_x;
  const int tile_y = t % a.tiles_y; t /= a.tiles_y;
  const int tile_n = t;
  const int lg_tn = 7 - a.lg_tw - a.lg_th;
  const int sx0 = tile_x << a.lg_tw, sy0 = tile_y << a.lg_th, n_img0 = tile_n << lg_tn;
  const int n0 = blockIdx.y * BN;
  const int total_k = a.taps.n * a.k_blocks;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      for (int tap = 0; tap < a.taps.n; ++tap) {
        const int ix = sx0 * a.in_step + a.taps.dx[tap], iy = sy0 * a.in_step + a.taps.dy[tap];
        const int wi = a.taps.w[tap];
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_4d(sa, &tmA, &full[stage], kb * 64, ix, iy, n_img0);
          tma_load_3d(sa + A_BYTES, &tmB, &full[stage], kb * 64, n0, wi);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      int stage = 0, phase = 0;
      for (int k = 0; k < total_k; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_sdesc(sa, 16, 1024);
        const uint64_t bdesc = make_sdesc(sa + A_BYTES, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzled row
          umma_bf16(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) != 0);
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 ; row m of the tile = lane of TMEM
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int tw = m & ((1 << a.lg_tw) - 1);
    const int th = (m >> a.lg_tw) & ((1 << a.lg_th) - 1);
    const int tn = m >> (a.lg_tw + a.lg_th);
    const int sx = sx0 + tw, sy = sy0 + th, img = n_img0 + tn;
    const bool valid = sx < a.Wsub && sy < a.Hsub && img < a.Nimg;
    const size_t pix = ((size_t)img * a.Hout + (size_t)(sy * a.o_step + a.o_y0)) * a.Wout + (sx * a.o_step + a.o_x0);
    __nv_bfloat16* orow = a.out + pix * a.ldc + a.c_off + n0;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (n0 + c + g * 8 < a.Cout) {
            uint4* dst = (uint4*)(orow + c + g * 8);
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[g * 8 + e]);
            if (a.accumulate) {
              const uint4 old = *dst;
              const __nv_bfloat162* ob = (const __nv_bfloat162*)&old;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 o2 = __bfloat1622float2(ob[e]);
                f[2 * e] += o2.x; f[2 * e + 1] += o2.y;
              }
            }
            uint4 v;
            v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
            v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
            *dst = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- persistent variant: tile loop per CTA, two TMEM accumulators, TMA-store epilogue -----------------------
// One CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the MMA issuer fills accumulator
// (t & 1) while the epilogue warps drain the other one, 64 output channels at a time: TMEM -> registers -> bf16 ->
// 128B-swizzled staging slot (a ring of two 16 KB slots) -> one TMA tensor store (or reduce-add) per slot, so the
// output leaves the SM as full 128-byte lines and the epilogue overlaps the next tile's main loop.
// STATS: the batch-norm statistics of the output (per-channel sum and sum of squares of the bf16 values that were
// stored) are accumulated from the staged slot into per-CTA shared-memory arrays and flushed with one atomic per
// channel per CTA -- the separate statistics pass over the tensor (one full HBM read) disappears.
struct IgemmPArgs {
  int lg_tw, lg_th;
  int tiles_x, tiles_y, tiles_n, cout_tiles;
  int in_step, k_blocks;
  int accumulate;
  int Wout, Hout, Nimg, Cout;
  float* stat_sum; float* stat_sq;
  // HALO variant (k_blocks == 1): the distinct input rows of the filter are staged once per tile with their
  // horizontal halo and every tap reads its shifted window out of them
  int halo_rows, halo_row_bytes, halo_tx_bytes, halo_dx0, halo_k16, halo_base_offset_mode;
  short halo_dy[4];
  short halo_row_of_tap[MAX_TAPS];
  Taps taps;
};
constexpr int SLOT_BYTES = 128 * 128;     // 128 rows x 64 bf16
template <int BN, int STAGES>
constexpr int igemm_p_smem_bytes(int stat_channels) {
  return STAGES * (A_BYTES + BN * 128) + 2 * SLOT_BYTES + 4 * stat_channels * 4 + 1024 + 256;
}

constexpr int STAT_THREADS = 256;                      // eight warps that only accumulate the statistics
constexpr int IGP_THREADS_STATS = IG_THREADS + STAT_THREADS;
constexpr int STAT_BAR = 128 + STAT_THREADS;           // epilogue + statistics threads on the slot barriers
// HALO (3x3-style filters on maps at least 65 pixels wide, Cin <= 64, BN = 64): a tile is 128 consecutive pixels of
// one image row.  Instead of one shifted A tile per tap (R*S re-reads of the input through L2), the R distinct input
// rows are staged once per tile as (128 + span) x 64-channel boxes and tap (r, s) is the 128-row window that starts s
// pixels into row r: a UMMA descriptor whose start address is 128*s bytes into the swizzled box (the 128-byte swizzle
// is a function of the shared-memory address, so TMA and MMA agree on any 128-byte-aligned window).  The R*S weight
// tiles of the CTA's output-channel block stay resident in shared memory; blockIdx.y = output-channel block.
template <int BN, int STAGES, bool STATS, bool HALO = false>
__global__ void __launch_bounds__(STATS ? IGP_THREADS_STATS : IG_THREADS, 1) igemm_p_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmC,
                                                                const IgemmPArgs a) {
  constexpr int B_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // HALO: [resident weights: taps x BN x 128 B][stages: rows x row_bytes]; otherwise [stages: A tile + B tile]
  const int STAGE_BYTES = HALO ? a.halo_rows * a.halo_row_bytes : A_BYTES + B_BYTES;
  uint8_t* wres = smem_base;
  uint8_t* smem = smem_base + (HALO ? a.taps.n * B_BYTES : 0);
  uint8_t* ring = smem + STAGES * STAGE_BYTES;
  const int stat_c = STATS ? ((a.Cout + 63) & ~63) : 0;
  float* s_sum = (float*)(ring + 2 * SLOT_BYTES);       // [2 row halves][stat_c] sums, then the same for the squares
  float* s_sq = s_sum + 2 * stat_c;
  uint64_t* full = (uint64_t*)(s_sq + 2 * stat_c);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* wbar = acc_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(wbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (STATS)
    for (int i = threadIdx.x; i < 4 * stat_c; i += blockDim.x) s_sum[i] = 0.f;
  if (warp == 0 && lane == 0) { prefetch_map(&tmA); prefetch_map(&tmB); prefetch_map(&tmC); }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int lg_tn = 7 - a.lg_tw - a.lg_th;
  const int n_tiles = a.tiles_x * a.tiles_y * a.tiles_n * (HALO ? 1 : a.cout_tiles);
  const int total_k = a.taps.n * a.k_blocks;

  auto decode = [&](int t, int& sx0, int& sy0, int& img0, int& n0) {
    if (HALO) n0 = blockIdx.y * BN;
    else { n0 = (t % a.cout_tiles) * BN; t /= a.cout_tiles; }
    sx0 = (t % a.tiles_x) << a.lg_tw; t /= a.tiles_x;
    sy0 = (t % a.tiles_y) << a.lg_th; t /= a.tiles_y;
    img0 = t << lg_tn;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      if (HALO) {       // the CTA's weight tiles, once
        mbar_arrive_expect_tx(wbar, (uint32_t)(a.taps.n * B_BYTES));
        for (int tap = 0; tap < a.taps.n; ++tap)
          tma_load_3d(wres + tap * B_BYTES, &tmB, wbar, 0, blockIdx.y * BN, a.taps.w[tap]);
      }
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int sx0, sy0, img0, n0;
        decode(t, sx0, sy0, img0, n0);
        if (HALO) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], (uint32_t)a.halo_tx_bytes);     // boxes are (128 + span) rows, slots are padded
          uint8_t* sa = smem + stage * STAGE_BYTES;
          for (int r = 0; r < a.halo_rows; ++r)
            tma_load_4d(sa + r * a.halo_row_bytes, &tmA, &full[stage], 0, sx0 + a.halo_dx0, sy0 + a.halo_dy[r], img0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
        for (int tap = 0; tap < a.taps.n; ++tap) {
          const int ix = sx0 * a.in_step + a.taps.dx[tap], iy = sy0 * a.in_step + a.taps.dy[tap];
          const int wi = a.taps.w[tap];
          for (int kb = 0; kb < a.k_blocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            uint8_t* sa = smem + stage * STAGE_BYTES;
            tma_load_4d(sa, &tmA, &full[stage], kb * 64, ix, iy, img0);
            tma_load_3d(sa + A_BYTES, &tmB, &full[stage], kb * 64, n0, wi);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      int stage = 0, phase = 0, it = 0;
      if (HALO) { mbar_wait(wbar, 0); tc_fence_after(); }
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        if (HALO) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t wa = smem_u32(wres);
          for (int tap = 0; tap < a.taps.n; ++tap) {
            const uint32_t start = sa + a.halo_row_of_tap[tap] * a.halo_row_bytes + (a.taps.dx[tap] - a.halo_dx0) * 128;
            const uint64_t adesc = make_sdesc(start, 16, 1024, a.halo_base_offset_mode ? (start >> 7) : 0);
            const uint64_t bdesc = make_sdesc(wa + tap * B_BYTES, 16, 1024);
            for (int kk = 0; kk < a.halo_k16; ++kk)
              umma_bf16(acc, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (tap | kk) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          umma_commit(&acc_full[buf]);
          continue;
        }
        for (int k = 0; k < total_k; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_sdesc(sa, 16, 1024);
          const uint64_t bdesc = make_sdesc(sa + A_BYTES, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(acc, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) != 0);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const bool leader = threadIdx.x == 64;     // first epilogue thread issues the TMA stores
    const int tw = m & ((1 << a.lg_tw) - 1);
    const int th = (m >> a.lg_tw) & ((1 << a.lg_th) - 1);
    const int tn = m >> (a.lg_tw + a.lg_th);
    int it = 0, slot = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      int sx0, sy0, img0, n0;
      decode(t, sx0, sy0, img0, n0);
      // rows of a partial tile that fall outside the tensor: the TMA store clips them, the statistics must not see them
      const bool row_ok = !STATS || (sx0 + tw < a.Wout && sy0 + th < a.Hout && img0 + tn < a.Nimg);
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < BN / 64; ++h, ++slot) {
        uint8_t* so = ring + (slot & 1) * SLOT_BYTES;
        // this slot was handed to the TMA two stores ago: that store must have finished reading it (and, with
        // STATS, the statistics warps must be done with it: barrier 4 + slot, 256 threads)
        if (leader) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        if (STATS && slot >= 2) named_bar_sync(4 + (slot & 1), STAT_BAR);
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + h * 64;
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 32, r + 32);
        tmem_ld_wait();
        if (h == BN / 64 - 1) {
          // accumulator drained: hand it back to the MMA issuer (one arrival per epilogue warp)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        uint8_t* row = so + m * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1]));
          v.y = pack_bf16x2(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3]));
          v.z = pack_bf16x2(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5]));
          v.w = pack_bf16x2(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7]));
          if (!row_ok) v = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(row + ((j ^ (m & 7)) << 4)) = v;
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (leader) {
          if (a.accumulate) tma_reduce_add_4d(&tmC, so, n0 + h * 64, sx0, sy0, img0);
          else tma_store_4d(&tmC, so, n0 + h * 64, sx0, sy0, img0);
          tma_store_commit();
        }
        if (STATS) named_bar_arrive(2 + (slot & 1), STAT_BAR);  // slot staged: the statistics warps may read it
      }
    }
    if (leader) tma_store_wait<0>();
  } else if (STATS) {
    // statistics warps: warp w owns channels [16(w&3), +16) and rows [64(w>>2), +64) of every staged 64-channel slot
    // (bf16 values exactly as stored).  Lane = (column pair cp8, row group rsub): 16 rows x 2 channels each, the four
    // row groups folded with two shuffle steps, then lanes 0..7 add into the CTA's shared accumulators of their row
    // half -- always the same thread for a given (half, channel), so no atomics.  The groups walk their rows with a
    // stagger of two so that the 128-byte swizzle sends the four concurrent rows to different banks.
    // Barriers 2/3 = slot 0/1 staged, 4/5 = slot 0/1 consumed.
    const int et = threadIdx.x - IG_THREADS;
    const int sw = et >> 5;
    const int cpair = (sw & 3) * 8 + (lane & 7), rsub = lane >> 3, half = sw >> 2;
    float* my_sum = s_sum + half * stat_c;
    float* my_sq = s_sq + half * stat_c;
    const int total_slots = (int)(BN / 64) * ((n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
    int slot = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int sx0, sy0, img0, n0;
      decode(t, sx0, sy0, img0, n0);
#pragma unroll 1
      for (int h = 0; h < BN / 64; ++h, ++slot) {
        const uint8_t* so = ring + (slot & 1) * SLOT_BYTES + (cpair & 3) * 4;
        named_bar_sync(2 + (slot & 1), STAT_BAR);
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int rr = half * 64 + rsub * 16 + ((i + 2 * rsub) & 15);
          const uint32_t u = *reinterpret_cast<const uint32_t*>(so + rr * 128 + (((cpair >> 2) ^ (rr & 7)) << 4));
          const float v0 = __uint_as_float(u << 16), v1 = __uint_as_float(u & 0xffff0000u);
          s0 += v0; s1 += v1;
          q0 = fmaf(v0, v0, q0); q1 = fmaf(v1, v1, q1);
        }
        // done reading: the epilogue warps may overwrite the slot (only waited for when the slot is used again)
        if (slot + 2 < total_slots) named_bar_arrive(4 + (slot & 1), STAT_BAR);
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        }
        const int c = n0 + h * 64 + cpair * 2;
        if (rsub == 0 && c < a.Cout) {
          my_sum[c] += s0; my_sum[c + 1] += s1;
          my_sq[c] += q0; my_sq[c + 1] += q1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  if (STATS) {
    for (int i = threadIdx.x; i < a.Cout; i += blockDim.x) {
      atomicAdd(&a.stat_sum[i], s_sum[i] + s_sum[stat_c + i]);
      atomicAdd(&a.stat_sq[i], s_sq[i] + s_sq[stat_c + i]);
    }
  }
}

// ---- cosine-cost Gram matrix for the Sinkhorn set-up -------------------------------------------------------
// C[i][j] = 1 - <A_i, B_j> / (na_i nb_j) with A [n,K], B [m,K] bf16 K-major (K = 6d: the bf16x3 split of the
// fp32 features arranged so that one dot product sums the six significant cross terms), fp32 accumulation in
// TMEM, fp32 output; optionally also the transpose.  Same pipeline as igemm_kernel, 2-D tensor maps.
struct GramArgs {
  int n, m, k_blocks;
  const float* na; const float* nb;
  float* c; long long ldc;
  float* ct; long long ldct;
};
constexpr int GRAM_STAGES = 3;
constexpr int gram_smem_bytes() { return GRAM_STAGES * (A_BYTES + 128 * 128) + 1024 + 256; }

__global__ void __launch_bounds__(IG_THREADS) gram_cost_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const GramArgs a) {
  constexpr int BN = 128, STAGES = GRAM_STAGES;
  constexpr int STAGE_BYTES = A_BYTES + BN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int i0 = blockIdx.y * 128, j0 = blockIdx.x * BN;     // x fastest: CTAs sharing the A rows run together
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        tma_load_2d(sa, &tmA, &full[stage], kb * 64, i0);
        tma_load_2d(sa + A_BYTES, &tmB, &full[stage], kb * 64, j0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      int stage = 0, phase = 0;
      for (int k = 0; k < a.k_blocks; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_sdesc(sa, 16, 1024), bdesc = make_sdesc(sa + A_BYTES, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) != 0);
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;
    const bool row_ok = i < a.n;
    const float ni = row_ok ? a.na[i] : 1.f;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = j0 + c + e;
        const float nj = j < a.m ? a.nb[j] : 1.f;
        v[e] = __fsub_rn(1.0f, __fdiv_rn(__uint_as_float(r[e]), __fmul_rn(ni, nj)));
      }
      if (row_ok) {
        float* crow = a.c + (long long)i * a.ldc + j0 + c;
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          if (j0 + c + e < a.m) *reinterpret_cast<float4*>(crow + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
      if (a.ct) {
        // transposed copy: for a fixed column the 32 lanes (consecutive rows i) write 128 contiguous bytes
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = j0 + c + e;
          if (row_ok && j < a.m) a.ct[(long long)j * a.ldct + i] = v[e];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

int gram_cost(const void* A, int n, const void* B, int m, int k, const float* na, const float* nb, float* c,
              long long ldc, float* ct, long long ldct, cudaStream_t st) {
  AADG_REQUIRE(k % 8 == 0 && ldc % 4 == 0, "gram_cost: k %% 8 and ldc %% 4 must be 0");
  CUtensorMap mA, mB;
  {
    const long long dims[2] = {k, n};
    const long long strides[1] = {k};
    const int box[2] = {64, 128};
    int rc = make_map_bf16(&mA, A, 2, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  {
    const long long dims[2] = {k, m};
    const long long strides[1] = {k};
    const int box[2] = {64, 128};
    int rc = make_map_bf16(&mB, B, 2, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  GramArgs a{};
  a.n = n; a.m = m; a.k_blocks = (k + 63) / 64; a.na = na; a.nb = nb; a.c = c; a.ldc = ldc; a.ct = ct; a.ldct = ldct;
  static bool attr_set = false;
  if (!attr_set) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(gram_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_smem_bytes()));
    attr_set = true;
  }
  dim3 grid((m + 127) / 128, (n + 127) / 128);
  AADG_REQUIRE(grid.y <= 65535, "too many rows for one launch");
  gram_cost_kernel<<<grid, IG_THREADS, gram_smem_bytes(), st>>>(mA, mB, a);
  return check_launch("gram cost kernel");
}

// ---- wgrad --------------------------------------------------------------------------------------------
struct WgradArgs {
  int lg_tw, lg_th;            // log2 extents of the 64-pixel K tile in x, y
  int tiles_x, tiles_y, tiles_n;
  int in_step;                 // X coordinate = dY coordinate * in_step + tap offset
  int ksplit;                  // pixel tiles are dealt round-robin to ksplit CTAs
  int Cout, Cin;
  float* dw;                   // [taps][Cout][Cin] fp32, accumulated with red.add
  Taps taps;
};
constexpr int WG_PIX = 64;
constexpr int WG_A_BYTES = 2 * WG_PIX * 128;     // two 64-channel blocks of Cout

template <int BN, int STAGES>
constexpr int wgrad_smem_bytes() { return STAGES * (WG_A_BYTES + (BN / 64) * WG_PIX * 128) + 1024 + 256; }

template <int BN, int STAGES>
__global__ void __launch_bounds__(IG_THREADS) wgrad_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                           const __grid_constant__ CUtensorMap tmX,
                                                           const WgradArgs a) {
  constexpr int B_BYTES = (BN / 64) * WG_PIX * 128;
  constexpr int STAGE_BYTES = WG_A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  // Cout <= 64: the second 64-channel half of the dY operand is all padding -- zero it once instead of having the
  // TMA zero-fill a fully out-of-bounds box on every step (a third of the copy work of the small-channel layers)
  const bool half_a = a.Cout <= 64;
  if (half_a) {
    for (int i = threadIdx.x; i < STAGES * (WG_PIX * 128 / 16); i += IG_THREADS) {
      const int st = i / (WG_PIX * 128 / 16), o = i % (WG_PIX * 128 / 16);
      reinterpret_cast<uint4*>(smem + st * STAGE_BYTES + WG_PIX * 128)[o] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { prefetch_map(&tmDY); prefetch_map(&tmX); }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int co0 = blockIdx.x * 128, ci0 = blockIdx.y * BN;
  // taps of one pixel range are adjacent in launch order: the dY / X slices of that range stay in L2 while all
  // R*S taps consume them (tap-major order re-read both tensors from HBM once per tap)
  const int n_taps = a.taps.n;
  const int tap = blockIdx.z % n_taps, split = blockIdx.z / n_taps;
  const int n_tiles = a.tiles_x * a.tiles_y * a.tiles_n;
  const int my_tiles = split < n_tiles ? (n_tiles - split + a.ksplit - 1) / a.ksplit : 0;
  const int lg_tn = 6 - a.lg_tw - a.lg_th;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      const int dx = a.taps.dx[tap], dy = a.taps.dy[tap];
      for (int i = 0; i < my_tiles; ++i) {
        int t = split + i * a.ksplit;
        const int tx = t % a.tiles_x; t /= a.tiles_x;
        const int ty = t % a.tiles_y; t /= a.tiles_y;
        const int ox = tx << a.lg_tw, oy = ty << a.lg_th, img = t << lg_tn;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], half_a ? STAGE_BYTES - WG_PIX * 128 : STAGE_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        tma_load_4d(sa, &tmDY, &full[stage], co0, ox, oy, img);
        if (!half_a) tma_load_4d(sa + WG_PIX * 128, &tmDY, &full[stage], co0 + 64, ox, oy, img);
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_load_4d(sa + WG_A_BYTES + b * WG_PIX * 128, &tmX, &full[stage], ci0 + b * 64,
                      ox * a.in_step + dx, oy * a.in_step + dy, img);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
      int stage = 0, phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < WG_PIX / 16; ++kk) {   // 16 pixels = two 8-row groups = 2048 bytes per step
          const uint64_t adesc = make_sdesc(sa + kk * 2048, WG_PIX * 128, 1024);
          const uint64_t bdesc = make_sdesc(sa + WG_A_BYTES + kk * 2048, WG_PIX * 128, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (i | kk) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (my_tiles > 0) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    float* drow = a.dw + ((size_t)tap * a.Cout + co) * a.Cin + ci0;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      tmem_ld_wait();
      if (co < a.Cout) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (ci0 + c + g * 4 < a.Cin) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c + g * 4),
                         "f"(__uint_as_float(r[g * 4])), "f"(__uint_as_float(r[g * 4 + 1])),
                         "f"(__uint_as_float(r[g * 4 + 2])), "f"(__uint_as_float(r[g * 4 + 3]))
                         : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ---- wgrad, halo-reuse variant ---------------------------------------------------------------------------
// For 3x3-style filters with Cin <= 64 on maps wider than 64 pixels (stride 1): a K tile is 128 consecutive pixels of one
// image row.  dY is staged once per tile and the distinct input rows of the filter once each with their horizontal
// halo; tap (r, s) is the MN-major B window that starts s pixels into row r (same address-based swizzle argument as
// the forward halo kernel), accumulated into its own 64 TMEM columns.  512 columns hold 8 taps, so a 3x3 filter
// takes two passes (blockIdx.y) of 5 + 4 taps; every pass stages only the rows its taps touch.  Compared with one
// CTA per tap this cuts the TMA boxes per 128 pixels from 4 per tap to 1 + rows per pass.
struct WgradHaloArgs {
  int tiles_x, H, N, ksplit;
  int Cout, Cin;
  float* dw;
  int row_bytes, box_bytes, dx0, taps_per_pass;
  short row_of_tap[MAX_TAPS];
  short row_dy[4];
  Taps taps;
};
constexpr int WH_PIX = 128;
constexpr int WH_A_BYTES = 2 * WH_PIX * 128;
constexpr int WH_STAGES = 2;

__global__ void __launch_bounds__(IG_THREADS, 1) wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                                   const __grid_constant__ CUtensorMap tmX,
                                                                   const WgradHaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int STAGE_BYTES = WH_A_BYTES + 3 * a.row_bytes;
  uint64_t* full = (uint64_t*)(smem + WH_STAGES * STAGE_BYTES);
  uint64_t* empty = full + WH_STAGES;
  uint64_t* acc_full = empty + WH_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * 128;
  const int tap0 = blockIdx.y * a.taps_per_pass, tap1 = min(a.taps.n, tap0 + a.taps_per_pass);
  const int r_lo = a.row_of_tap[tap0], n_rows = a.row_of_tap[tap1 - 1] - r_lo + 1;
  const int split = blockIdx.z;
  const int n_tiles = a.tiles_x * a.H * a.N;
  const int my_tiles = split < n_tiles ? (n_tiles - split + a.ksplit - 1) / a.ksplit : 0;
  const bool half_a = a.Cout - co0 <= 64;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WH_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (half_a) {      // the second 64-channel half of dY is padding: zero it once
    for (int i = threadIdx.x; i < WH_STAGES * (WH_PIX * 128 / 16); i += IG_THREADS) {
      const int st = i / (WH_PIX * 128 / 16), o = i % (WH_PIX * 128 / 16);
      reinterpret_cast<uint4*>(smem + st * STAGE_BYTES + WH_PIX * 128)[o] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { prefetch_map(&tmDY); prefetch_map(&tmX); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      const uint32_t tx_bytes = (uint32_t)((half_a ? WH_PIX * 128 : WH_A_BYTES) + n_rows * a.box_bytes);
      for (int i = 0; i < my_tiles; ++i) {
        int t = split + i * a.ksplit;
        const int ox = (t % a.tiles_x) * WH_PIX; t /= a.tiles_x;
        const int oy = t % a.H, img = t / a.H;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], tx_bytes);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        tma_load_4d(sa, &tmDY, &full[stage], co0, ox, oy, img);
        if (!half_a) tma_load_4d(sa + WH_PIX * 128, &tmDY, &full[stage], co0 + 64, ox, oy, img);
        for (int r = 0; r < n_rows; ++r)
          tma_load_4d(sa + WH_A_BYTES + r * a.row_bytes, &tmX, &full[stage], 0, ox + a.dx0, oy + a.row_dy[r_lo + r], img);
        if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      int stage = 0, phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        for (int tap = tap0; tap < tap1; ++tap) {
          const uint32_t xb = sa + WH_A_BYTES + (a.row_of_tap[tap] - r_lo) * a.row_bytes + (a.taps.dx[tap] - a.dx0) * 128;
#pragma unroll
          for (int kk = 0; kk < WH_PIX / 16; ++kk) {
            const uint64_t adesc = make_sdesc(sa + kk * 2048, WH_PIX * 128, 1024);
            const uint64_t bdesc = make_sdesc(xb + kk * 2048, WH_PIX * 128, 1024);
            umma_bf16(tmem_base + (tap - tap0) * 64, adesc, bdesc, idesc, (i | kk) != 0);
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (my_tiles > 0) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int tap = tap0; tap < tap1; ++tap) {
      float* drow = a.dw + ((size_t)a.taps.w[tap] * a.Cout + co) * a.Cin;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (tap - tap0) * 64 + c, r);
        tmem_ld_wait();
        if (co < a.Cout) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (c + g * 4 < a.Cin) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c + g * 4),
                           "f"(__uint_as_float(r[g * 4])), "f"(__uint_as_float(r[g * 4 + 1])),
                           "f"(__uint_as_float(r[g * 4 + 2])), "f"(__uint_as_float(r[g * 4 + 3]))
                           : "memory");
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- host ---------------------------------------------------------------------------------------------
static int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// split 2^total rows of a tile between x, y (and images): x first, then y
static void pick_tile(int W, int H, int total_lg, int* lg_tw, int* lg_th) {
  *lg_tw = std::min(ilog2_ceil(W), total_lg);
  *lg_th = std::min(ilog2_ceil(H), total_lg - *lg_tw);
}

struct ConvGeom {
  int N, H, W, Cin, ldx;        // input
  int Ho, Wo, Cout, ldy;        // output
  int R, S, stride, pad, dil;
};

static int tuning_stages() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AADG_CONV_STAGES");
    v = e ? atoi(e) : 2;
    if (v < 2 || v > 4) v = 2;
  }
  return v;
}

static int tuning_persistent() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AADG_CONV_PERSISTENT");
    v = e ? atoi(e) : 1;
  }
  return v;
}
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

template <int BN, int STAGES>
static int launch_igemm_t(const CUtensorMap& mA, const CUtensorMap& mB, const IgemmArgs& args, int cout_tiles,
                          cudaStream_t st) {
  const int smem = igemm_smem_bytes<BN, STAGES>();
  static bool attr_set = false;
  if (!attr_set) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(igemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(args.tiles_x * args.tiles_y * args.tiles_n, cout_tiles);
  igemm_kernel<BN, STAGES><<<grid, IG_THREADS, smem, st>>>(mA, mB, args);
  return check_launch("igemm kernel");
}

static int tuning_bn256() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AADG_CONV_BN256");
    v = e ? atoi(e) : 1;
  }
  return v;
}

template <int BN, int STAGES, bool STATS>
static int launch_igemm_p_t(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mC, const IgemmPArgs& pa,
                            int grid, cudaStream_t st) {
  const int stat_c = STATS ? ((pa.Cout + 63) & ~63) : 0;
  const int smem = igemm_p_smem_bytes<BN, STAGES>(stat_c);
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(igemm_p_kernel<BN, STAGES, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_bytes = smem;
  }
  igemm_p_kernel<BN, STAGES, STATS><<<grid, STATS ? IGP_THREADS_STATS : IG_THREADS, smem, st>>>(mA, mB, mC, pa);
  return check_launch("igemm persistent kernel");
}

static int tuning_halo() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AADG_CONV_HALO");      // 0 = off, 1 = on (descriptor base offset 0), 2 = on with base offset
    v = e ? atoi(e) : 1;
  }
  return v;
}

template <bool STATS>
static int launch_igemm_halo_t(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mC, const IgemmPArgs& pa,
                               dim3 grid, int smem, cudaStream_t st) {
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(igemm_p_kernel<64, 2, STATS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_bytes = smem;
  }
  igemm_p_kernel<64, 2, STATS, true><<<grid, STATS ? IGP_THREADS_STATS : IG_THREADS, smem, st>>>(mA, mB, mC, pa);
  return check_launch("igemm halo kernel");
}

// One implicit-GEMM launch.  in: bf16 [Nimg, Hin, Win, ld_in] (Cin channels used); wgt: bf16 [n_w_taps][Cn][Cin];
// logical output grid Wsub x Hsub mapped to physical pixels by (o_step, o_y0, o_x0).
static int launch_igemm(const void* in, int Nimg, int Hin, int Win, int Cin, int ld_in, int in_step,
                        const void* wgt, int n_w_taps, int Cn, const Taps& taps, int Wsub, int Hsub, void* out,
                        int Hout, int Wout, int ldc, int c_off, int o_step, int o_y0, int o_x0, int accumulate,
                        cudaStream_t st, float* stat_sum = nullptr, float* stat_sq = nullptr) {
  AADG_REQUIRE(ld_in % 8 == 0 && ldc % 8 == 0 && c_off % 8 == 0 && Cn % 8 == 0 && Cin % 8 == 0,
               "channel counts/strides must be multiples of 8 (Cin %d ld_in %d Cout %d ldc %d off %d)", Cin, ld_in,
               Cn, ldc, c_off);
  AADG_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)wgt & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "tensors must be 16-byte aligned");
  IgemmArgs a{};
  pick_tile(Wsub, Hsub, 7, &a.lg_tw, &a.lg_th);
  const int lg_tn = 7 - a.lg_tw - a.lg_th;
  a.tiles_x = (Wsub + (1 << a.lg_tw) - 1) >> a.lg_tw;
  a.tiles_y = (Hsub + (1 << a.lg_th) - 1) >> a.lg_th;
  a.tiles_n = (Nimg + (1 << lg_tn) - 1) >> lg_tn;
  a.Wsub = Wsub; a.Hsub = Hsub; a.Nimg = Nimg;
  a.in_step = in_step;
  a.k_blocks = (Cin + 63) / 64;
  a.out = (__nv_bfloat16*)out;
  a.ldc = ldc; a.c_off = c_off; a.Cout = Cn;
  a.Hout = Hout; a.Wout = Wout;
  a.o_step = o_step; a.o_y0 = o_y0; a.o_x0 = o_x0;
  a.accumulate = accumulate;
  a.taps = taps;
  AADG_REQUIRE((1 << a.lg_tw) * in_step <= 256 && (1 << a.lg_th) * in_step <= 256, "tile box too large");

  CUtensorMap mA, mB;
  {
    const long long dims[4] = {Cin, Win, Hin, Nimg};
    const long long strides[3] = {ld_in, (long long)Win * ld_in, (long long)Hin * Win * ld_in};
    const int box[4] = {64, (1 << a.lg_tw) * in_step, (1 << a.lg_th) * in_step, 1 << lg_tn};
    const int es[4] = {1, in_step, in_step, 1};
    int rc = make_map_bf16(&mA, in, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  const bool persistent = o_step == 1 && o_y0 == 0 && o_x0 == 0 && Wsub == Wout && Hsub == Hout && tuning_persistent();
  AADG_REQUIRE(!stat_sum || persistent, "fused statistics need the persistent kernel (dense stride-1 output grid)");
  AADG_REQUIRE(!stat_sum || Cn <= 2048, "fused statistics support at most 2048 output channels");
  // 256-wide tiles halve the A-operand traffic per flop; worth it once there are enough 256-column tiles to fill
  // the SMs and the main loop is long enough to hide the 4-slot epilogue
  const long long pix_tiles = (long long)a.tiles_x * a.tiles_y * a.tiles_n;
  const bool wide = persistent && tuning_bn256() && Cn % 256 == 0 && a.k_blocks * taps.n >= 4 &&
                    pix_tiles * (Cn / 256) >= num_sms();
  const int bn = Cn <= 64 ? 64 : (wide ? 256 : 128);
  {
    const long long dims[3] = {Cin, Cn, n_w_taps};
    const long long strides[2] = {Cin, (long long)Cn * Cin};
    const int box[3] = {64, bn, 1};
    int rc = make_map_bf16(&mB, wgt, 3, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  const int cout_tiles = (Cn + bn - 1) / bn;
  // HALO variant: several taps, one 64-channel k-block, tiles = 128 pixels of one image row
  if (persistent && tuning_halo() && taps.n > 1 && a.k_blocks == 1 && in_step == 1 && a.lg_tw == 7 && a.lg_th == 0) {
    IgemmPArgs pa{};
    int dx0 = 1 << 20, dx1 = -(1 << 20);
    for (int t = 0; t < taps.n; ++t) {
      dx0 = std::min<int>(dx0, taps.dx[t]); dx1 = std::max<int>(dx1, taps.dx[t]);
      int r = 0;
      while (r < pa.halo_rows && pa.halo_dy[r] != taps.dy[t]) ++r;
      if (r == pa.halo_rows) {
        if (pa.halo_rows == 4) { pa.halo_rows = 99; break; }
        pa.halo_dy[pa.halo_rows++] = taps.dy[t];
      }
      pa.halo_row_of_tap[t] = (short)r;
    }
    const int span = dx1 - dx0;
    const int row_bytes = (int)align_up((size_t)(128 + span) * 128, 1024);
    const int halo_cout_tiles = (Cn + 63) / 64;
    const int stat_c = stat_sum ? ((Cn + 63) & ~63) : 0;
    const int smem = taps.n * 64 * 128 + 2 * pa.halo_rows * row_bytes + 2 * SLOT_BYTES + 4 * stat_c * 4 + 1024 + 256;
    if (pa.halo_rows <= 4 && 128 + span <= 256 && smem <= 227 * 1024) {
      CUtensorMap hA, hB, mC;
      {
        const long long dims[4] = {Cin, Win, Hin, Nimg};
        const long long strides[3] = {ld_in, (long long)Win * ld_in, (long long)Hin * Win * ld_in};
        const int box[4] = {64, 128 + span, 1, 1};
        int rc = make_map_bf16(&hA, in, 4, dims, strides, box, nullptr);
        if (rc) return rc;
      }
      {
        const long long dims[3] = {Cin, Cn, n_w_taps};
        const long long strides[2] = {Cin, (long long)Cn * Cin};
        const int box[3] = {64, 64, 1};
        int rc = make_map_bf16(&hB, wgt, 3, dims, strides, box, nullptr);
        if (rc) return rc;
      }
      {
        const long long dims[4] = {Cn, Wout, Hout, Nimg};
        const long long strides[3] = {ldc, (long long)Wout * ldc, (long long)Hout * Wout * ldc};
        const int box[4] = {64, 128, 1, 1};
        int rc = make_map_bf16(&mC, (const __nv_bfloat16*)out + c_off, 4, dims, strides, box, nullptr);
        if (rc) return rc;
      }
      pa.lg_tw = 7; pa.lg_th = 0;
      pa.tiles_x = a.tiles_x; pa.tiles_y = a.tiles_y; pa.tiles_n = a.tiles_n; pa.cout_tiles = halo_cout_tiles;
      pa.in_step = 1; pa.k_blocks = 1; pa.accumulate = accumulate;
      pa.Wout = Wout; pa.Hout = Hout; pa.Nimg = Nimg; pa.Cout = Cn;
      pa.stat_sum = stat_sum; pa.stat_sq = stat_sq;
      pa.halo_row_bytes = row_bytes; pa.halo_dx0 = dx0; pa.halo_k16 = (Cin + 15) / 16;
      pa.halo_tx_bytes = pa.halo_rows * (128 + span) * 128;
      pa.halo_base_offset_mode = tuning_halo() == 2 ? 1 : 0;
      pa.taps = taps;
      const long long n_tiles = (long long)pa.tiles_x * pa.tiles_y * pa.tiles_n;
      dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(n_tiles, num_sms() / halo_cout_tiles)), halo_cout_tiles);
      if (stat_sum) return launch_igemm_halo_t<true>(hA, hB, mC, pa, grid, smem, st);
      return launch_igemm_halo_t<false>(hA, hB, mC, pa, grid, smem, st);
    }
  }
  if (persistent) {
    // persistent kernel with TMA-store epilogue: output = channel slice [c_off, c_off + Cn) of the NHWC tensor
    CUtensorMap mC;
    const long long dims[4] = {Cn, Wout, Hout, Nimg};
    const long long strides[3] = {ldc, (long long)Wout * ldc, (long long)Hout * Wout * ldc};
    const int box[4] = {64, 1 << a.lg_tw, 1 << a.lg_th, 1 << lg_tn};
    int rc = make_map_bf16(&mC, (const __nv_bfloat16*)out + c_off, 4, dims, strides, box, nullptr);
    if (rc) return rc;
    IgemmPArgs pa{};
    pa.lg_tw = a.lg_tw; pa.lg_th = a.lg_th;
    pa.tiles_x = a.tiles_x; pa.tiles_y = a.tiles_y; pa.tiles_n = a.tiles_n; pa.cout_tiles = cout_tiles;
    pa.in_step = in_step; pa.k_blocks = a.k_blocks; pa.accumulate = accumulate;
    pa.Wout = Wout; pa.Hout = Hout; pa.Nimg = Nimg; pa.Cout = Cn;
    pa.stat_sum = stat_sum; pa.stat_sq = stat_sq;
    pa.taps = taps;
    const int n_tiles = pa.tiles_x * pa.tiles_y * pa.tiles_n * cout_tiles;
    const int grid = std::min(n_tiles, num_sms());
    if (stat_sum) {
      if (bn == 64) return launch_igemm_p_t<64, 6, true>(mA, mB, mC, pa, grid, st);
      if (bn == 128) return launch_igemm_p_t<128, 4, true>(mA, mB, mC, pa, grid, st);
      return launch_igemm_p_t<256, 3, true>(mA, mB, mC, pa, grid, st);
    }
    if (bn == 64) return launch_igemm_p_t<64, 6, false>(mA, mB, mC, pa, grid, st);
    if (bn == 128) return launch_igemm_p_t<128, 4, false>(mA, mB, mC, pa, grid, st);
    return launch_igemm_p_t<256, 4, false>(mA, mB, mC, pa, grid, st);
  }
  // fewer stages = less shared memory = more CTAs per SM: the per-CTA latencies (TMEM allocation, first TMA
  // round trip, epilogue) of one CTA overlap the main loop of its neighbours
  const int stages = tuning_stages();
  if (bn == 64) {
    if (stages == 2) return launch_igemm_t<64, 2>(mA, mB, a, cout_tiles, st);
    if (stages == 3) return launch_igemm_t<64, 3>(mA, mB, a, cout_tiles, st);
    return launch_igemm_t<64, 4>(mA, mB, a, cout_tiles, st);
  }
  if (stages == 2) return launch_igemm_t<128, 2>(mA, mB, a, cout_tiles, st);
  if (stages == 3) return launch_igemm_t<128, 3>(mA, mB, a, cout_tiles, st);
  return launch_igemm_t<128, 4>(mA, mB, a, cout_tiles, st);
}

template <int BN, int STAGES>
static int launch_wgrad_t(const CUtensorMap& mDY, const CUtensorMap& mX, const WgradArgs& args, dim3 grid,
                          cudaStream_t st) {
  const int smem = wgrad_smem_bytes<BN, STAGES>();
  static bool attr_set = false;
  if (!attr_set) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  wgrad_kernel<BN, STAGES><<<grid, IG_THREADS, smem, st>>>(mDY, mX, args);
  return check_launch("wgrad kernel");
}

}  // namespace tc
}  // namespace aadg

using namespace aadg;
using namespace aadg::tc;

static int check_geom(const ConvGeom& g) {
  AADG_REQUIRE(g.N > 0 && g.H > 0 && g.W > 0 && g.Cin > 0 && g.Cout > 0, "bad tensor sizes");
  AADG_REQUIRE(g.R > 0 && g.S > 0 && g.R * g.S <= MAX_TAPS, "filter %dx%d not supported (max %d taps)", g.R, g.S, MAX_TAPS);
  AADG_REQUIRE(g.stride == 1 || g.stride == 2, "stride %d not supported", g.stride);
  AADG_REQUIRE(g.dil >= 1 && g.pad >= 0, "bad dilation/padding");
  const int ho = (g.H + 2 * g.pad - g.dil * (g.R - 1) - 1) / g.stride + 1;
  const int wo = (g.W + 2 * g.pad - g.dil * (g.S - 1) - 1) / g.stride + 1;
  AADG_REQUIRE(ho == g.Ho && wo == g.Wo, "output size mismatch: expected %dx%d, got %dx%d", ho, wo, g.Ho, g.Wo);
  return AADG_OK;
}

extern "C" {

/* y[n,oy,ox,c_off+co] (+)= sum_{r,s,ci} x[n, oy*stride-pad+r*dil, ox*stride-pad+s*dil, ci] * w[r*S+s][co][ci] */
int aadg_conv_fprop_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* wgt, int cout, int r,
                         int s, int stride, int pad, int dil, void* y, int ho, int wo, int ldy, int y_c_off,
                         int accumulate, void* stream) {
  ConvGeom g{n, h, w, cin, ldx, ho, wo, cout, ldy, r, s, stride, pad, dil};
  int rc = check_geom(g);
  if (rc) return rc;
  Taps taps{};
  taps.n = r * s;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < s; ++j) {
      taps.dy[i * s + j] = (short)(i * dil - pad);
      taps.dx[i * s + j] = (short)(j * dil - pad);
      taps.w[i * s + j] = (short)(i * s + j);
    }
  return launch_igemm(x, n, h, w, cin, ldx, stride, wgt, r * s, cout, taps, wo, ho, y, ho, wo, ldy, y_c_off, 1, 0, 0,
                      accumulate, (cudaStream_t)stream);
}

/* the same forward convolution, and the batch-norm statistics of its output in the same pass: stat_sum[co] and
 * stat_sq[co] (fp32, `cout` entries each, ACCUMULATED: zero them first) receive the per-channel sum and sum of
 * squares of the bf16 values written to y.  Replaces aadg_bn_stats over y (one full read of the tensor). */
int aadg_conv_fprop_stats_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* wgt, int cout, int r,
                               int s, int stride, int pad, int dil, void* y, int ho, int wo, int ldy, int y_c_off,
                               float* stat_sum, float* stat_sq, void* stream) {
  ConvGeom g{n, h, w, cin, ldx, ho, wo, cout, ldy, r, s, stride, pad, dil};
  int rc = check_geom(g);
  if (rc) return rc;
  AADG_REQUIRE(stat_sum && stat_sq, "statistics buffers are required");
  Taps taps{};
  taps.n = r * s;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < s; ++j) {
      taps.dy[i * s + j] = (short)(i * dil - pad);
      taps.dx[i * s + j] = (short)(j * dil - pad);
      taps.w[i * s + j] = (short)(i * s + j);
    }
  return launch_igemm(x, n, h, w, cin, ldx, stride, wgt, r * s, cout, taps, wo, ho, y, ho, wo, ldy, y_c_off, 1, 0, 0, 0,
                      (cudaStream_t)stream, stat_sum, stat_sq);
}

/* dx[n,iy,ix,c_off+ci] (+)= sum over (r,s,co) with iy = oy*stride-pad+r*dil of dy[n,oy,ox,co] * wt[r*S+s][ci][co]
 * (wt = the forward weights with the two channel axes swapped).  Geometry arguments are the FORWARD
 * convolution's. */
int aadg_conv_dgrad_bf16(const void* dy, int n, int ho, int wo, int cout, int lddy, const void* wgt_t, int cin,
                         int r, int s, int stride, int pad, int dil, void* dx, int h, int w, int lddx,
                         int dx_c_off, int accumulate, void* stream) {
  ConvGeom g{n, h, w, cin, lddx, ho, wo, cout, lddy, r, s, stride, pad, dil};
  int rc = check_geom(g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate && stride > 1) {
    // parity classes that no filter tap reaches (e.g. the odd pixels of a strided 1x1) get zero gradient
    bool any_empty = false;
    for (int py = 0; py < stride; ++py)
      for (int px = 0; px < stride; ++px) {
        int cnt = 0;
        for (int i = 0; i < r; ++i)
          for (int j = 0; j < s; ++j)
            if ((py + pad - i * dil) % stride == 0 && (px + pad - j * dil) % stride == 0) ++cnt;
        any_empty |= cnt == 0;
      }
    if (any_empty) {
      AADG_REQUIRE(cin == lddx && dx_c_off == 0, "zero-fill of untouched pixels needs a dense dx tensor");
      AADG_CUDA_TRY(cudaMemsetAsync(dx, 0, (size_t)n * h * w * lddx * 2, st));
    }
  }
  for (int py = 0; py < stride; ++py)
    for (int px = 0; px < stride; ++px) {
      const int hsub = (h - py + stride - 1) / stride, wsub = (w - px + stride - 1) / stride;
      if (hsub <= 0 || wsub <= 0) continue;
      Taps taps{};
      for (int i = 0; i < r; ++i) {
        const int ty = py + pad - i * dil;
        if (ty % stride) continue;
        for (int j = 0; j < s; ++j) {
          const int tx = px + pad - j * dil;
          if (tx % stride) continue;
          // floor division is exact here (ty, tx are multiples of stride, possibly negative)
          taps.dy[taps.n] = (short)(ty / stride);
          taps.dx[taps.n] = (short)(tx / stride);
          taps.w[taps.n] = (short)(i * s + j);
          ++taps.n;
        }
      }
      if (taps.n == 0) continue;   // untouched pixels were zeroed above
      rc = launch_igemm(dy, n, ho, wo, cout, lddy, 1, wgt_t, r * s, cin, taps, wsub, hsub, dx, h, w, lddx, dx_c_off,
                        stride, py, px, accumulate, st);
      if (rc) return rc;
    }
  return AADG_OK;
}

/* dw[r*S+s][co][ci] += sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*stride-pad+r*dil, ox*stride-pad+s*dil, ci]
 * (fp32, accumulated: zero dw first for a fresh gradient). */
static int wgrad_impl(const void* x, int n, int h, int w, int cin, int ldx, const void* dy, int ho, int wo, int cout,
                      int lddy, int r, int s, int stride, int pad, int dil, float* dw, void* stream);

int aadg_conv_wgrad_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* dy, int ho, int wo,
                         int cout, int lddy, int r, int s, int stride, int pad, int dil, float* dw, void* stream) {
  ConvGeom g{n, h, w, cin, ldx, ho, wo, cout, lddy, r, s, stride, pad, dil};
  int rc = check_geom(g);
  if (rc) return rc;
  return wgrad_impl(x, n, h, w, cin, ldx, dy, ho, wo, cout, lddy, r, s, stride, pad, dil, dw, stream);
}

/* "Window" convolutions: a VALID (no padding), stride-1 convolution whose input pixel pitch `ldx` may be SMALLER than
 * `cin`, i.e. every input "pixel" is a window of cin consecutive elements that overlaps its neighbours (the tensor map
 * simply has a dimension whose stride is shorter than the extent of the dimension below it).  The caller chooses the
 * output extent ho <= h - r + 1, wo <= w - s + 1 and guarantees that the windows it reaches stay inside the allocation
 * (windows of the last pixels run past the tensor's nominal end by cin - ldx elements).  Used by the ResNet stem: the
 * 7x7 / stride-2 convolution of models/__init__.py:17-23 (smp encoder conv1) over the space-to-depth image of
 * aadg_stem_s2d is r = 4, s = 1, cin = 64, ldx = 16.  stat_sum / stat_sq optional (fused batch-norm statistics). */
int aadg_conv_fprop_windows_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* wgt, int cout, int r,
                                 int s, void* y, int ho, int wo, int ldy, float* stat_sum, float* stat_sq, void* stream) {
  AADG_REQUIRE(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && r > 0 && s > 0 && r * s <= MAX_TAPS, "bad window-conv sizes");
  AADG_REQUIRE(ho > 0 && wo > 0 && ho <= h - r + 1 && wo <= w - s + 1, "output %dx%d does not fit a valid %dx%d filter on %dx%d",
               ho, wo, r, s, h, w);
  AADG_REQUIRE((stat_sum == nullptr) == (stat_sq == nullptr), "pass both statistics buffers or neither");
  Taps taps{};
  taps.n = r * s;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < s; ++j) {
      taps.dy[i * s + j] = (short)i;
      taps.dx[i * s + j] = (short)j;
      taps.w[i * s + j] = (short)(i * s + j);
    }
  return launch_igemm(x, n, h, w, cin, ldx, 1, wgt, r * s, cout, taps, wo, ho, y, ho, wo, ldy, 0, 1, 0, 0, 0,
                      (cudaStream_t)stream, stat_sum, stat_sq);
}

/* dw[t][co][k] += sum over output pixels of dy[n,oy,ox,co] * x_window[n, oy + i, ox + j][k], t = i*s + j: the weight
 * gradient of aadg_conv_fprop_windows_bf16 (same geometry contract) */
int aadg_conv_wgrad_windows_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* dy, int ho, int wo,
                                 int cout, int lddy, int r, int s, float* dw, void* stream) {
  AADG_REQUIRE(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && r > 0 && s > 0 && r * s <= MAX_TAPS, "bad window-conv sizes");
  AADG_REQUIRE(ho > 0 && wo > 0 && ho <= h - r + 1 && wo <= w - s + 1, "output %dx%d does not fit a valid %dx%d filter on %dx%d",
               ho, wo, r, s, h, w);
  return wgrad_impl(x, n, h, w, cin, ldx, dy, ho, wo, cout, lddy, r, s, 1, 0, 1, dw, stream);
}

}  // extern "C"

static int wgrad_impl(const void* x, int n, int h, int w, int cin, int ldx, const void* dy, int ho, int wo, int cout,
                      int lddy, int r, int s, int stride, int pad, int dil, float* dw, void* stream) {
  int rc = AADG_OK;
  AADG_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && cin % 8 == 0 && cout % 8 == 0, "channels must be multiples of 8");
  AADG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dw & 15) == 0 && cin % 4 == 0,
               "tensors must be 16-byte aligned");
  WgradArgs a{};
  pick_tile(wo, ho, 6, &a.lg_tw, &a.lg_th);
  const int lg_tn = 6 - a.lg_tw - a.lg_th;
  a.tiles_x = (wo + (1 << a.lg_tw) - 1) >> a.lg_tw;
  a.tiles_y = (ho + (1 << a.lg_th) - 1) >> a.lg_th;
  a.tiles_n = (n + (1 << lg_tn) - 1) >> lg_tn;
  a.in_step = stride;
  a.Cout = cout; a.Cin = cin; a.dw = dw;
  a.taps.n = r * s;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < s; ++j) {
      a.taps.dy[i * s + j] = (short)(i * dil - pad);
      a.taps.dx[i * s + j] = (short)(j * dil - pad);
      a.taps.w[i * s + j] = (short)(i * s + j);
    }
  // halo-reuse variant: several taps, Cin <= 64, stride 1, rows of 128 pixels
  {
    static int halo_mode = -1;
    if (halo_mode < 0) { const char* e = getenv("AADG_WGRAD_HALO"); halo_mode = e ? atoi(e) : 1; }
    if (halo_mode && stride == 1 && r * s > 1 && cin <= 64 && wo >= 65) {
      WgradHaloArgs ha{};
      int n_rows = 0, dx0 = 1 << 20, dx1 = -(1 << 20);
      bool ok = true;
      for (int t = 0; t < a.taps.n; ++t) {
        dx0 = std::min<int>(dx0, a.taps.dx[t]); dx1 = std::max<int>(dx1, a.taps.dx[t]);
        int q = 0;
        while (q < n_rows && ha.row_dy[q] != a.taps.dy[t]) ++q;
        if (q == n_rows) {
          if (n_rows == 4) { ok = false; break; }
          ha.row_dy[n_rows++] = a.taps.dy[t];
        }
        ha.row_of_tap[t] = (short)q;
      }
      const int span = dx1 - dx0;
      // 512 TMEM columns = 8 taps x 64 per pass, taps balanced over the passes (3x3 -> 5 + 4); a pass may touch at most
      // 3 distinct rows (its shared-memory stage holds three): more passes until that holds (4x1 -> 2 + 2)
      int passes = (a.taps.n + 7) / 8, tpp = a.taps.n;
      for (; ok && passes <= a.taps.n; ++passes) {
        tpp = (a.taps.n + passes - 1) / passes;
        bool fits = true;
        for (int p = 0; p * tpp < a.taps.n; ++p) {
          const int t0 = p * tpp, t1 = std::min(a.taps.n, t0 + tpp);
          if (ha.row_of_tap[t1 - 1] - ha.row_of_tap[t0] + 1 > 3 || ha.row_of_tap[t1 - 1] < ha.row_of_tap[t0]) fits = false;
        }
        if (fits) break;
      }
      if (passes > a.taps.n) ok = false;
      passes = (a.taps.n + tpp - 1) / tpp;
      const int row_bytes = (int)align_up((size_t)(WH_PIX + span) * 128, 1024);
      const int smem = WH_STAGES * (WH_A_BYTES + 3 * row_bytes) + 1024 + 256;
      if (ok && WH_PIX + span <= 256 && smem <= 227 * 1024) {
        CUtensorMap hDY, hX;
        {
          const long long dims[4] = {cout, wo, ho, n};
          const long long strides[3] = {lddy, (long long)wo * lddy, (long long)ho * wo * lddy};
          const int box[4] = {64, WH_PIX, 1, 1};
          rc = make_map_bf16(&hDY, dy, 4, dims, strides, box, nullptr);
          if (rc) return rc;
        }
        {
          const long long dims[4] = {cin, w, h, n};
          const long long strides[3] = {ldx, (long long)w * ldx, (long long)h * w * ldx};
          const int box[4] = {64, WH_PIX + span, 1, 1};
          rc = make_map_bf16(&hX, x, 4, dims, strides, box, nullptr);
          if (rc) return rc;
        }
        ha.tiles_x = (wo + WH_PIX - 1) / WH_PIX; ha.H = ho; ha.N = n;
        ha.Cout = cout; ha.Cin = cin; ha.dw = dw;
        ha.row_bytes = row_bytes; ha.box_bytes = (WH_PIX + span) * 128; ha.dx0 = dx0; ha.taps_per_pass = tpp;
        ha.taps = a.taps;
        const long long tiles = (long long)ha.tiles_x * ho * n;
        const int base = ((cout + 127) / 128) * passes;
        ha.ksplit = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(tiles, 65535), (148 + base - 1) / base));
        static int attr = 0;
        if (smem > attr) {
          AADG_CUDA_TRY(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          attr = smem;
        }
        dim3 grid((cout + 127) / 128, passes, ha.ksplit);
        wgrad_halo_kernel<<<grid, IG_THREADS, smem, (cudaStream_t)stream>>>(hDY, hX, ha);
        return check_launch("wgrad halo kernel");
      }
    }
  }
  // 256-wide Cin tiles halve the dY traffic per flop (one CTA per SM pair of stages in flight instead of four small ones)
  static int wide_mode = -1;
  if (wide_mode < 0) { const char* e = getenv("AADG_WGRAD_BN256"); wide_mode = e ? atoi(e) : 0; }   // measured neutral on the ResNet-50 step: off by default
  const int n_tiles = a.tiles_x * a.tiles_y * a.tiles_n;
  const bool wide = wide_mode > 0 && cin % 256 == 0 && n_tiles >= 64;
  const int bn = cin <= 64 ? 64 : (wide ? 256 : 128);
  const int base_ctas = ((cout + 127) / 128) * ((cin + bn - 1) / bn) * r * s;
  int ksplit = (148 * (wide ? 2 : 4) + base_ctas - 1) / base_ctas;
  ksplit = std::max(1, std::min(ksplit, (n_tiles + 7) / 8));
  ksplit = std::min(ksplit, 65535 / (r * s));
  a.ksplit = ksplit;
  CUtensorMap mDY, mX;
  {
    const long long dims[4] = {cout, wo, ho, n};
    const long long strides[3] = {lddy, (long long)wo * lddy, (long long)ho * wo * lddy};
    const int box[4] = {64, 1 << a.lg_tw, 1 << a.lg_th, 1 << lg_tn};
    rc = make_map_bf16(&mDY, dy, 4, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  {
    const long long dims[4] = {cin, w, h, n};
    const long long strides[3] = {ldx, (long long)w * ldx, (long long)h * w * ldx};
    const int box[4] = {64, (1 << a.lg_tw) * stride, (1 << a.lg_th) * stride, 1 << lg_tn};
    const int es[4] = {1, stride, stride, 1};
    AADG_REQUIRE(box[1] <= 256 && box[2] <= 256, "tile box too large");
    rc = make_map_bf16(&mX, x, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  dim3 grid((cout + 127) / 128, (cin + bn - 1) / bn, r * s * ksplit);
  const int stages = tuning_stages();
  if (bn == 256) {
    if (wide_mode == 3) return launch_wgrad_t<256, 3>(mDY, mX, a, grid, (cudaStream_t)stream);
    return launch_wgrad_t<256, 2>(mDY, mX, a, grid, (cudaStream_t)stream);
  }
  if (bn == 64) {
    if (stages == 2) return launch_wgrad_t<64, 2>(mDY, mX, a, grid, (cudaStream_t)stream);
    if (stages == 3) return launch_wgrad_t<64, 3>(mDY, mX, a, grid, (cudaStream_t)stream);
    return launch_wgrad_t<64, 4>(mDY, mX, a, grid, (cudaStream_t)stream);
  }
  if (stages == 2) return launch_wgrad_t<128, 2>(mDY, mX, a, grid, (cudaStream_t)stream);
  if (stages == 3) return launch_wgrad_t<128, 3>(mDY, mX, a, grid, (cudaStream_t)stream);
  return launch_wgrad_t<128, 4>(mDY, mX, a, grid, (cudaStream_t)stream);
}
