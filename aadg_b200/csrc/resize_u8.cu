// DGRandomScaleCrop + Normalize_dg + ToTensor on the device, bit-exact with Pillow.
//
// Replaces data/transform.py:97-135 (Image.resize BILINEAR for the image / NEAREST for the ORIGINAL mask,
// ImageOps.expand zero padding, crop) and :138-236 (x/127.5-1, mask -> multilabel, HWC -> CHW) for every
// augmented copy.  Pillow's BILINEAR resize is a separable convolution with 22-bit fixed-point weights,
// horizontal pass first, uint8 intermediate (Resample.c); its NEAREST resize accumulates the source
// coordinate in double (Geometry.c).  Both are restated here so that outputs match the reference bit for bit
// given the same decisions (scale_w, scale_h, pad, crop_x, crop_y of aadg_aug_row_t).
//
//   coef_kernel  : per (row, axis, output index) the tap window and integer weights, in IEEE double with
//                  explicit round-to-nearest operations (no FMA contraction), like the C code;
//   crop_kernel  : one thread per output pixel of the crop: <= KMAX x KMAX taps of uint8 source reads
//                  (L1/L2 resident), float32 CHW image + label stores.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace aadg {
namespace rs {

constexpr int KMAX = 8;                 // taps per axis: bilinear support 1 * max(scale, 1) -> covers down to 1/3.5
constexpr int PRECISION_BITS = 32 - 8 - 2;

struct RowGeom {
  int img;                 // image index in `images`
  int src;                 // source index (mask)
  int do_scale, sw, sh;    // scaled size (== W, H when !do_scale)
  int pad, cx, cy;
};
struct Axis {              // one output index of one axis
  int lo, cnt;
  int k[KMAX];
};

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter
__global__ void coef_kernel(const RowGeom* rows, int n_rows, int in_w, int in_h, int max_w, int max_h, Axis* xa,
                            Axis* ya, int* xn, int* yn) {
  const int r = blockIdx.y;
  const RowGeom g = rows[r];
  if (!g.do_scale) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int axis = 0; axis < 2; ++axis) {
    const int in = axis ? in_h : in_w, out = axis ? g.sh : g.sw;
    if (i >= out) continue;
    Axis a;
    const double scale = __ddiv_rn((double)in, (double)out);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;            // bilinear support = 1.0
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(__dadd_rn((double)i, 0.5), scale);
    int lo = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (hi > in) hi = in;
    int n = hi - lo;
    if (n > KMAX) n = KMAX;
    double w[KMAX];
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
      if (t < 0.0) t = -t;
      w[x] = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
      ww = __dadd_rn(ww, w[x]);
    }
    for (int x = 0; x < KMAX; ++x) {
      int kk = 0;
      if (x < n) {
        double v = w[x];
        if (ww != 0.0) v = __ddiv_rn(v, ww);
        kk = (int)__dadd_rn(0.5, __dmul_rn(v, (double)(1 << PRECISION_BITS)));
      }
      a.k[x] = kk;
    }
    a.lo = lo; a.cnt = n;
    (axis ? ya : xa)[(size_t)r * (axis ? max_h : max_w) + i] = a;
  }
  // Geometry.c ImagingScaleAffine (NEAREST): source index = (int)o, o accumulated in double
  if (blockIdx.x == 0 && threadIdx.x < 2) {
    const int axis = threadIdx.x;
    const int in = axis ? in_h : in_w, out = axis ? g.sh : g.sw;
    int* dst = (axis ? yn : xn) + (size_t)r * (axis ? max_h : max_w);
    const double scale = __ddiv_rn((double)in, (double)out);
    double o = __dmul_rn(scale, 0.5);
    for (int k = 0; k < out; ++k) {
      const int idx = o >= 0.0 ? (int)o : -1;
      dst[k] = idx < in ? idx : -1;
      o = __dadd_rn(o, scale);
    }
  }
}

__device__ __forceinline__ int clip8(long long acc) {
  const long long v = acc >> PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : (int)v);
}

struct CropArgs {
  const uint8_t* images; const uint8_t* masks;
  const RowGeom* rows; const Axis* xa; const Axis* ya; const int* xn; const int* yn;
  float* out_images; float* out_labels;
  int H, W, cw, ch, max_w, max_h, dataset;
};

// grid (x chunks, crop rows, n_rows)
__global__ void __launch_bounds__(128) crop_kernel(const CropArgs a) {
  const int r = blockIdx.z, oy = blockIdx.y;
  const RowGeom g = a.rows[r];
  const int W = a.W, H = a.H;
  const uint8_t* img = a.images + (size_t)g.img * H * W * 3;
  const uint8_t* msk = a.masks ? a.masks + (size_t)g.src * H * W : nullptr;
  const int py = g.cy + oy - g.pad;
  const bool row_in = py >= 0 && py < g.sh;
  const size_t plane = (size_t)a.ch * a.cw;
  const int nlab = a.dataset == AADG_DATASET_OPTIC ? 2 : 1;
  const bool need_h = g.do_scale && g.sw != W, need_v = g.do_scale && g.sh != H;
  for (int ox = blockIdx.x * blockDim.x + threadIdx.x; ox < a.cw; ox += gridDim.x * blockDim.x) {
    const int px = g.cx + ox - g.pad;
    int v[3] = {0, 0, 0};
    int m = 0;
    if (row_in && px >= 0 && px < g.sw) {
      // vertical window over horizontally resampled rows
      int ylo = py, ycnt = 1;
      const int* ky = nullptr;
      if (need_v) { const Axis& ay = a.ya[(size_t)r * a.max_h + py]; ylo = ay.lo; ycnt = ay.cnt; ky = ay.k; }
      int xlo = px, xcnt = 1;
      const int* kx = nullptr;
      if (need_h) { const Axis& ax = a.xa[(size_t)r * a.max_w + px]; xlo = ax.lo; xcnt = ax.cnt; kx = ax.k; }
      long long accv[3] = {1ll << (PRECISION_BITS - 1), 1ll << (PRECISION_BITS - 1), 1ll << (PRECISION_BITS - 1)};
      for (int t = 0; t < ycnt; ++t) {
        const uint8_t* srow = img + (size_t)(ylo + t) * W * 3;
        int hval[3];
        if (need_h) {
          long long acch[3] = {1ll << (PRECISION_BITS - 1), 1ll << (PRECISION_BITS - 1), 1ll << (PRECISION_BITS - 1)};
          for (int k = 0; k < xcnt; ++k) {
            const uint8_t* p = srow + (size_t)(xlo + k) * 3;
            acch[0] += (long long)kx[k] * p[0]; acch[1] += (long long)kx[k] * p[1]; acch[2] += (long long)kx[k] * p[2];
          }
          hval[0] = clip8(acch[0]); hval[1] = clip8(acch[1]); hval[2] = clip8(acch[2]);
        } else {
          const uint8_t* p = srow + (size_t)xlo * 3;
          hval[0] = p[0]; hval[1] = p[1]; hval[2] = p[2];
        }
        if (need_v) {
          accv[0] += (long long)ky[t] * hval[0]; accv[1] += (long long)ky[t] * hval[1]; accv[2] += (long long)ky[t] * hval[2];
        } else {
          v[0] = hval[0]; v[1] = hval[1]; v[2] = hval[2];
        }
      }
      if (need_v) { v[0] = clip8(accv[0]); v[1] = clip8(accv[1]); v[2] = clip8(accv[2]); }
      if (msk) {
        int my = py, mx = px;
        if (g.do_scale) { my = a.yn[(size_t)r * a.max_h + py]; mx = a.xn[(size_t)r * a.max_w + px]; }
        m = (my >= 0 && mx >= 0) ? msk[(size_t)my * W + mx] : 0;
      }
    }
    const size_t o = (size_t)oy * a.cw + ox;
    if (a.out_images) {
      float* oi = a.out_images + (size_t)r * 3 * plane + o;
#pragma unroll
      for (int c = 0; c < 3; ++c) oi[c * plane] = __fsub_rn(__fdiv_rn((float)v[c], 127.5f), 1.0f);
    }
    if (a.out_labels && msk) {
      float* ol = a.out_labels + (size_t)r * nlab * plane + o;
      if (a.dataset == AADG_DATASET_OPTIC) {
        const bool bg = m > 200, ring = m > 50 && m < 201;
        ol[0] = (!bg && !ring) ? 1.f : 0.f;
        ol[plane] = bg ? 0.f : 1.f;
      } else {
        ol[0] = m != 0 ? 1.f : 0.f;
      }
    }
  }
}

struct Layout { size_t rows, xa, ya, xn, yn, total; };
static Layout layout(int n_rows, int max_w, int max_h) {
  Layout L{};
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = align_up(off, 256); off = o + b; return o; };
  L.rows = take(sizeof(RowGeom) * n_rows);
  L.xa = take(sizeof(Axis) * (size_t)n_rows * max_w);
  L.ya = take(sizeof(Axis) * (size_t)n_rows * max_h);
  L.xn = take(sizeof(int) * (size_t)n_rows * max_w);
  L.yn = take(sizeof(int) * (size_t)n_rows * max_h);
  L.total = align_up(off, 256);
  return L;
}

}  // namespace rs
}  // namespace aadg

using namespace aadg;
using namespace aadg::rs;

extern "C" {

size_t aadg_u8_scale_crop_workspace_bytes(int n_rows, int max_scale_w, int max_scale_h) {
  if (n_rows <= 0 || max_scale_w <= 0 || max_scale_h <= 0) return 0;
  return layout(n_rows, max_scale_w, max_scale_h).total;
}

int aadg_u8_scale_crop_normalize(const uint8_t* images, int image_by_row, const uint8_t* masks,
                                 const aadg_aug_row_t* rows, int n_rows, int n_src, int height, int width,
                                 int crop_w, int crop_h, int dataset, float* out_images, float* out_labels,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  AADG_REQUIRE(n_rows >= 0 && n_src > 0 && height > 0 && width > 0 && crop_w > 0 && crop_h > 0, "bad sizes");
  AADG_REQUIRE(dataset == AADG_DATASET_OPTIC || dataset == AADG_DATASET_VESSEL, "unknown dataset %d", dataset);
  if (n_rows == 0) return AADG_OK;
  AADG_REQUIRE(images && rows, "null images / rows");
  AADG_REQUIRE(crop_h <= 65535 && n_rows <= 65535, "crop too tall / too many rows");
  std::vector<RowGeom> hr(n_rows);
  int max_w = 1, max_h = 1;
  for (int r = 0; r < n_rows; ++r) {
    const aadg_aug_row_t& a = rows[r];
    AADG_REQUIRE(a.src >= 0 && a.src < n_src, "row %d: src %d out of range", r, a.src);
    RowGeom& g = hr[r];
    g.img = image_by_row ? r : a.src;
    g.src = a.src;
    g.do_scale = a.do_scale ? 1 : 0;
    g.sw = g.do_scale ? a.scale_w : width;
    g.sh = g.do_scale ? a.scale_h : height;
    AADG_REQUIRE(g.sw > 0 && g.sh > 0, "row %d: bad scaled size %dx%d", r, g.sw, g.sh);
    AADG_REQUIRE((double)width / g.sw <= 3.5 && (double)height / g.sh <= 3.5, "row %d: down-scaling beyond 3.5x", r);
    g.pad = a.pad; g.cx = a.crop_x; g.cy = a.crop_y;
    AADG_REQUIRE(g.pad >= 0 && g.cx >= 0 && g.cy >= 0 && g.cx + crop_w <= g.sw + 2 * g.pad &&
                     g.cy + crop_h <= g.sh + 2 * g.pad,
                 "row %d: crop window outside the padded image", r);
    max_w = std::max(max_w, g.sw); max_h = std::max(max_h, g.sh);
  }
  const Layout L = layout(n_rows, max_w, max_h);
  if (!workspace || workspace_bytes < L.total) {
    set_error("workspace too small: need %zu bytes, got %zu", L.total, workspace_bytes);
    return AADG_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  AADG_CUDA_TRY(cudaMemcpyAsync(w + L.rows, hr.data(), sizeof(RowGeom) * n_rows, cudaMemcpyHostToDevice, st));
  dim3 cg((std::max(max_w, max_h) + 127) / 128, n_rows);
  coef_kernel<<<cg, 128, 0, st>>>((const RowGeom*)(w + L.rows), n_rows, width, height, max_w, max_h, (Axis*)(w + L.xa),
                                  (Axis*)(w + L.ya), (int*)(w + L.xn), (int*)(w + L.yn));
  CropArgs a{};
  a.images = images; a.masks = masks; a.rows = (const RowGeom*)(w + L.rows);
  a.xa = (const Axis*)(w + L.xa); a.ya = (const Axis*)(w + L.ya); a.xn = (const int*)(w + L.xn); a.yn = (const int*)(w + L.yn);
  a.out_images = out_images; a.out_labels = out_labels;
  a.H = height; a.W = width; a.cw = crop_w; a.ch = crop_h; a.max_w = max_w; a.max_h = max_h; a.dataset = dataset;
  dim3 grid(std::max(1, std::min((crop_w + 127) / 128, 8)), crop_h, n_rows);
  crop_kernel<<<grid, 128, 0, st>>>(a);
  return check_launch("scale/crop kernels");
}

}  // extern "C"
