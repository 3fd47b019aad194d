// 95th-percentile Hausdorff distance of binary masks on the GPU (sm_100a): the validation metric of the
// reference, which calls medpy.metric.binary.hd95 per image and class on the CPU (search_dg.py:246-260,
// train_dg.py validate(): scipy binary_erosion + distance_transform_edt + numpy.percentile dominate validate()).
//
//   surface(M)  = M xor erode(M)      (4-neighbour cross, pixels outside the image count as background)
//   d(p, S)     = exact Euclidean distance from pixel p to the nearest pixel of surface S
//   hd95(A, B)  = numpy.percentile( {d(p, surface B) : p in surface A} U {d(p, surface A) : p in surface B}, 95 )
//
// Exact integer arithmetic up to the final square roots: kernel 1 scans every column once for the surface flags
// and the vertical distance g to the nearest surface pixel of that column; kernel 2 evaluates, for every surface
// pixel of the other mask, d^2 = min_x' ((x - x')^2 + g(x', y)^2) walking outwards from x until (x - x')^2 can no
// longer win; kernel 3 selects the two order statistics numpy's linear interpolation needs with a 3-digit radix
// select over the integer d^2 and reproduces numpy's lerp in float64 -- bit-identical to scipy/numpy results.
#include "common.cuh"

#include <algorithm>

namespace aadg {
namespace hd {

constexpr unsigned short NO_SURF = 0xFFFFu;

struct Layout {
  unsigned short* g;      // [pairs][2][h][w] vertical distance to the nearest surface pixel in the column
  int* d2;                // [pairs][2*h*w] squared surface distances (both directions)
  int* count;             // [pairs] entries in d2
  int* fg;                // [pairs][2] foreground pixel counts
};

__device__ __forceinline__ bool fg_at(const unsigned char* m, int h, int w, int y, int x) {
  return y >= 0 && y < h && x >= 0 && x < w && m[(size_t)y * w + x] != 0;
}

// one thread per (pair, mask kind, column)
__global__ void __launch_bounds__(128) columns_kernel(const unsigned char* result, const unsigned char* reference, int pairs,
                                                      int h, int w, Layout L) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int kind = blockIdx.y, p = blockIdx.z;
  if (x >= w) return;
  const unsigned char* m = (kind ? reference : result) + (size_t)p * h * w;
  unsigned short* g = L.g + ((size_t)(p * 2 + kind) * h) * w;
  int d = -1, nfg = 0;
  bool up = false, cur = fg_at(m, h, w, 0, x);
  for (int y = 0; y < h; ++y) {
    const bool down = fg_at(m, h, w, y + 1, x);
    const bool surf = cur && !(up && down && fg_at(m, h, w, y, x - 1) && fg_at(m, h, w, y, x + 1));
    nfg += cur;
    if (surf) d = 0; else if (d >= 0) ++d;
    g[(size_t)y * w + x] = d >= 0 ? (unsigned short)min(d, 0xFFFE) : NO_SURF;
    up = cur; cur = down;
  }
  d = -1;
  for (int y = h - 1; y >= 0; --y) {
    const unsigned short v = g[(size_t)y * w + x];
    if (v == 0) d = 0; else if (d >= 0) ++d;
    if (d >= 0 && (v == NO_SURF || d < v)) g[(size_t)y * w + x] = (unsigned short)min(d, 0xFFFE);
  }
  if (nfg) atomicAdd(&L.fg[p * 2 + kind], nfg);
}

// one thread per (pair, direction, pixel): surface pixels of mask `dir` measure their distance to the other surface
__global__ void __launch_bounds__(256) distances_kernel(int pairs, int h, int w, Layout L) {
  const int p = blockIdx.z, dir = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)h * w) return;
  const int y = (int)(i / w), x = (int)(i - (long long)y * w);
  const unsigned short* ga = L.g + ((size_t)(p * 2 + dir) * h) * w;
  const unsigned short* gb = L.g + ((size_t)(p * 2 + (dir ^ 1)) * h) * w + (size_t)y * w;
  if (ga[(size_t)y * w + x] != 0) return;
  long long best = -1;
  for (int dx = 0; dx < w; ++dx) {
    const long long dx2 = (long long)dx * dx;
    if (best >= 0 && dx2 >= best) break;
    if (x - dx < 0 && x + dx >= w) break;
    if (x - dx >= 0) {
      const unsigned short v = gb[x - dx];
      if (v != NO_SURF) { const long long c = dx2 + (long long)v * v; if (best < 0 || c < best) best = c; }
    }
    if (dx && x + dx < w) {
      const unsigned short v = gb[x + dx];
      if (v != NO_SURF) { const long long c = dx2 + (long long)v * v; if (best < 0 || c < best) best = c; }
    }
  }
  if (best < 0) return;          // the other mask has no surface at all: reported through the status word
  const int slot = atomicAdd(&L.count[p], 1);
  L.d2[(size_t)p * 2 * h * w + slot] = (int)best;
}

// k-th smallest (0-based) of n non-negative ints < 2^21 with a 7+7+7-bit radix select; one CTA
__device__ int radix_select(const int* v, int n, int k, int* hist) {
  int prefix = 0, mask = 0;
  for (int shift = 14; shift >= 0; shift -= 7) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int e = v[i];
      if ((e & mask) == prefix) atomicAdd(&hist[(e >> shift) & 127], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, digit = 0;
      for (; digit < 128; ++digit) {
        if (acc + hist[digit] > k) break;
        acc += hist[digit];
      }
      hist[128] = digit; hist[129] = acc;
    }
    __syncthreads();
    const int digit = hist[128];
    k -= hist[129];
    prefix |= digit << shift;
    mask |= 127 << shift;
    __syncthreads();
  }
  return prefix;
}

// numpy.percentile(values, 95) (method "linear") over sqrt(d2); status: 0 ok, 1 result empty, 2 reference empty
__global__ void __launch_bounds__(256) percentile_kernel(int h, int w, Layout L, double q, double* out, int* status) {
  __shared__ int hist[130];
  const int p = blockIdx.x;
  const int n = L.count[p];
  const int st = L.fg[p * 2] == 0 ? 1 : (L.fg[p * 2 + 1] == 0 ? 2 : 0);
  if (st || n == 0) {
    if (threadIdx.x == 0) { status[p] = st ? st : 2; out[p] = __longlong_as_double(0x7ff8000000000000ll); }
    return;
  }
  const int* v = L.d2 + (size_t)p * 2 * h * w;
  // numpy's "linear" method: virtual index (n - 1) * q, then _get_indexes / _get_gamma / _lerp
  // (explicit rounding steps below: a fused multiply-add would round differently from numpy's separate operations)
  const double vi = __dmul_rn((double)(n - 1), q);
  long long lo = (long long)floor(vi), hi = lo + 1;
  double t = vi - floor(vi);
  if (vi >= (double)(n - 1)) { lo = hi = n - 1; }
  if (vi < 0.0) { lo = hi = 0; }
  const int a2 = radix_select(v, n, (int)lo, hist);
  const int b2 = hi == lo ? a2 : radix_select(v, n, (int)hi, hist);
  if (threadIdx.x == 0) {
    const double a = sqrt((double)a2), b = sqrt((double)b2);
    const double diff = __dsub_rn(b, a);
    double r = __dadd_rn(a, __dmul_rn(diff, t));
    if (t >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
    out[p] = r;
    status[p] = 0;
  }
}

static size_t g_bytes(int pairs, int h, int w) { return align_up((size_t)pairs * 2 * h * w * sizeof(unsigned short), 256); }
static size_t d2_bytes(int pairs, int h, int w) { return align_up((size_t)pairs * 2 * h * w * sizeof(int), 256); }

}  // namespace hd
}  // namespace aadg

using namespace aadg;
using namespace aadg::hd;

extern "C" {

size_t aadg_hd95_workspace_bytes(int n_pairs, int h, int w) {
  if (n_pairs <= 0 || h <= 0 || w <= 0) return 0;
  return g_bytes(n_pairs, h, w) + d2_bytes(n_pairs, h, w) + align_up((size_t)n_pairs * 3 * sizeof(int), 256) + 256;
}

/* result, reference: uint8 [n_pairs][h][w] (non-zero = foreground); out float64 [n_pairs] = the
 * `percentile`-th percentile (95 for hd95) of the symmetric surface distances in pixels; status int32 [n_pairs]:
 * 0 ok, 1 = `result` has no foreground, 2 = `reference` has no foreground (medpy raises in both cases; out = NaN). */
int aadg_hd95(const unsigned char* result, const unsigned char* reference, int n_pairs, int h, int w, double percentile,
              double* out, int* status, void* workspace, size_t workspace_bytes, void* stream) {
  AADG_REQUIRE(n_pairs > 0 && h > 0 && w > 0 && h <= 32768 && w <= 32768 && n_pairs <= 65535, "bad sizes");
  AADG_REQUIRE((long long)h * w * 2 < (1ll << 31) && (long long)(h - 1) * (h - 1) + (long long)(w - 1) * (w - 1) < (1 << 21),
               "image too large for the 21-bit squared distances (max about 1024 x 1024)");
  AADG_REQUIRE(percentile >= 0.0 && percentile <= 100.0, "percentile must be in [0, 100]");
  if (workspace_bytes < aadg_hd95_workspace_bytes(n_pairs, h, w)) {
    set_error("hd95: workspace too small (%zu < %zu)", workspace_bytes, aadg_hd95_workspace_bytes(n_pairs, h, w));
    return AADG_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  Layout L;
  L.g = (unsigned short*)base; base += g_bytes(n_pairs, h, w);
  L.d2 = (int*)base; base += d2_bytes(n_pairs, h, w);
  L.count = (int*)base;
  L.fg = L.count + n_pairs;
  AADG_CUDA_TRY(cudaMemsetAsync(L.count, 0, (size_t)n_pairs * 3 * sizeof(int), st));
  columns_kernel<<<dim3((w + 127) / 128, 2, n_pairs), 128, 0, st>>>(result, reference, n_pairs, h, w, L);
  distances_kernel<<<dim3((unsigned)(((long long)h * w + 255) / 256), 2, n_pairs), 256, 0, st>>>(n_pairs, h, w, L);
  percentile_kernel<<<n_pairs, 256, 0, st>>>(h, w, L, percentile / 100.0, out, status);
  return check_launch("hd95");
}

}  // extern "C"
