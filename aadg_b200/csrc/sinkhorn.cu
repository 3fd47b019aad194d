// Debiased Sinkhorn divergence with the cosine cost, forward only, for sm_100a.
//
// Replaces geomloss 0.2.4 SamplesLoss("sinkhorn", cost="( IntCst(1) - (X | Y) / ( Norm2(X) * Norm2(Y) ) )",
// backend="online") as called at search_dg.py:116,158-160 / search_dg_2d.py:116,159-161 (about fifty
// KeOps LogSumExp launches and one host sync per call, eighteen calls per step) and the reward
// assembly around it (search_dg.py:150-162).  Algorithm: SURVEY.md App. A.1 / oracle/sinkhorn.py.
//
// Two regimes:
//   * small clouds (<= 64 points): one CTA per problem runs the whole epsilon schedule out of
//     shared memory — cost matrices, diameter, schedule, every soft-min with warp-shuffle
//     reductions — so a training step's 18 divergences are ONE launch with no host sync; the
//     reward variant also does the `[j::M]` / argmax-by-domain split and `rewards[j] += ...`.
//   * large clouds: the four cost matrices (xx, yy, xy and its transpose) are materialised once in
//     fp32 and every epsilon iteration streams them once (HBM-bound, 4*N*M*4 bytes): each soft-min
//     is a column-direction online log-sum-exp (coalesced 16-byte loads, per-thread running
//     max/sum, no cross-thread traffic), split over row bands and merged by a small combine kernel.
#include "tc_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace aadg {
namespace sk {

constexpr double BLUR = 0.05;
constexpr int MAX_EPS = 96;
constexpr int SMALL_MAX = 64;      // points per cloud handled by the shared-memory kernel
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// geomloss epsilon_schedule(p=2, diameter, blur=.05, scaling=.5):
//   [diam^2] + [exp(e) for e in np.arange(2 ln diam, 2 ln blur, 2 ln .5)] + [blur^2]
// np.arange(double): n = ceil((stop-start)/step), value_i = start + i*((start+step)-start).
__host__ __device__ inline int eps_schedule(double diam, double* out) {
  const double start = 2.0 * log(diam), stop = 2.0 * log(BLUR), step = 2.0 * log(0.5);
  double len = ceil((stop - start) / step);
  int n = len > 0 ? (int)len : 0;
  if (n > MAX_EPS - 2) n = MAX_EPS - 2;
  const double delta = (start + step) - start;
  int k = 0;
  out[k++] = diam * diam;
  for (int i = 0; i < n; ++i) out[k++] = exp(i == 0 ? start : start + i * delta);
  out[k++] = BLUR * BLUR;
  return k;
}

// ---------------------------------------------------------------------------------------------------
// small problems: one CTA each
// ---------------------------------------------------------------------------------------------------
struct SmallProblem {
  int x_off, x_n, y_off, y_n;   // rows of the point matrix
};

constexpr int SMALL_THREADS = 512;

// soft-min of one output element: -eps * LSE_k( h[k] - C(k) * P ), C(k) = cost[k*stride] ; whole warp
__device__ __forceinline__ float softmin_warp(const float* cost, int stride, const float* h, int n,
                                              float P, float neg_eps, int lane) {
  float v0 = -INFINITY, v1 = -INFINITY;
  if (lane < n) v0 = __fsub_rn(h[lane], __fmul_rn(cost[lane * stride], P));
  if (lane + 32 < n) v1 = __fsub_rn(h[lane + 32], __fmul_rn(cost[(lane + 32) * stride], P));
  const float m = warp_max(fmaxf(v0, v1));
  float s = 0.f;
  if (lane < n) s += expf(v0 - m);
  if (lane + 32 < n) s += expf(v1 - m);
  s = warp_sum(s);
  return neg_eps * (m + logf(s));
}

struct SmallSmem {
  float* X; float* Y;            // [n][d+1]
  float* Cxx; float* Cyy; float* Cxy;   // [N][N], [M][M], [N][M]
  float* nx; float* ny;          // norms
  float* ax; float* by; float* ay; float* bx;     // potentials
  float* hax; float* hby; float* hay; float* hbx; // log-weight + potential / eps
  float* red;                    // [64]
  double* eps;                   // [MAX_EPS]
  int* idx;                      // [2*SMALL_MAX] gathered row indices
  int* misc;                     // n_eps, counts
};

__device__ __forceinline__ SmallSmem carve(unsigned char* base, int d) {
  SmallSmem s;
  double* dp = (double*)base;
  s.eps = dp; dp += MAX_EPS;
  float* p = (float*)dp;
  const int ld = d + 1;
  s.X = p; p += SMALL_MAX * ld;
  s.Y = p; p += SMALL_MAX * ld;
  s.Cxx = p; p += SMALL_MAX * SMALL_MAX;
  s.Cyy = p; p += SMALL_MAX * SMALL_MAX;
  s.Cxy = p; p += SMALL_MAX * SMALL_MAX;
  s.nx = p; p += SMALL_MAX; s.ny = p; p += SMALL_MAX;
  s.ax = p; p += SMALL_MAX; s.by = p; p += SMALL_MAX; s.ay = p; p += SMALL_MAX; s.bx = p; p += SMALL_MAX;
  s.hax = p; p += SMALL_MAX; s.hby = p; p += SMALL_MAX; s.hay = p; p += SMALL_MAX; s.hbx = p; p += SMALL_MAX;
  s.red = p; p += 64;
  s.idx = (int*)p; p += 2 * SMALL_MAX;
  s.misc = (int*)p; p += 8;
  return s;
}
static size_t small_smem_bytes(int d) {
  return sizeof(double) * MAX_EPS +
         sizeof(float) * ((size_t)2 * SMALL_MAX * (d + 1) + 3 * SMALL_MAX * SMALL_MAX + 10 * SMALL_MAX + 64 +
                          2 * SMALL_MAX + 8);
}

// The whole divergence for the two clouds already gathered in shared memory (s.X [N][d+1], s.Y [M][d+1]).
// All threads call; the result is returned by every thread.
__device__ float small_divergence(const SmallSmem& s, int N, int M, int d) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SMALL_THREADS / 32;
  const int ld = d + 1;
  // norms (KeOps Norm2 = sqrt(sum of squares), fp32)
  for (int r = warp; r < N + M; r += nwarps) {
    const float* p = r < N ? s.X + r * ld : s.Y + (r - N) * ld;
    float a = 0.f;
    for (int k = lane; k < d; k += 32) a = fmaf(p[k], p[k], a);
    a = warp_sum(a);
    if (lane == 0) (r < N ? s.nx[r] : s.ny[r - N]) = sqrtf(a);
  }
  // diameter: || max(max x, max y) - min(min x, min y) ||_2 over coordinates (geomloss max_diameter)
  float part = 0.f;
  for (int k = tid; k < d; k += SMALL_THREADS) {
    float lo = INFINITY, hi = -INFINITY;
    for (int r = 0; r < N; ++r) { const float v = s.X[r * ld + k]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
    for (int r = 0; r < M; ++r) { const float v = s.Y[r * ld + k]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
    part = fmaf(hi - lo, hi - lo, part);
  }
  part = warp_sum(part);
  if (tid < 64) s.red[tid] = 0.f;
  __syncthreads();
  if (lane == 0) s.red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += s.red[w];
    const double diam = (double)sqrtf(t);
    s.misc[0] = eps_schedule(diam, s.eps);
  }
  // cost matrices: 1 - <x,y> / (|x| |y|), one thread per entry, sequential fp32 dot like KeOps
  const int total = N * N + M * M + N * M;
  for (int e = tid; e < total; e += SMALL_THREADS) {
    const float *p, *q; float np_, nq; float* out;
    if (e < N * N) { const int i = e / N, j = e % N; p = s.X + i * ld; q = s.X + j * ld; np_ = s.nx[i]; nq = s.nx[j]; out = s.Cxx + e; }
    else if (e < N * N + M * M) { const int f = e - N * N; const int i = f / M, j = f % M; p = s.Y + i * ld; q = s.Y + j * ld; np_ = s.ny[i]; nq = s.ny[j]; out = s.Cyy + f; }
    else { const int f = e - N * N - M * M; const int i = f / M, j = f % M; p = s.X + i * ld; q = s.Y + j * ld; np_ = s.nx[i]; nq = s.ny[j]; out = s.Cxy + f; }
    float a = 0.f;
    for (int k = 0; k < d; ++k) a = fmaf(p[k], q[k], a);
    *out = __fsub_rn(1.0f, __fdiv_rn(a, __fmul_rn(np_, nq)));
  }
  __syncthreads();
  const int n_eps = s.misc[0];
  const float alog = logf(__fdiv_rn(1.0f, (float)N)), blog = logf(__fdiv_rn(1.0f, (float)M));

  // iteration -1 is the initialisation at eps_s[0] with h = log weights; iterations 0..n_eps-1 the
  // averaged descent; iteration n_eps the final extrapolation at the last eps (no averaging).
  for (int it = -1; it <= n_eps; ++it) {
    const double eps_d = s.eps[it < 0 ? 0 : (it >= n_eps ? n_eps - 1 : it)];
    const float P = (float)(1.0 / eps_d);            // torch.Tensor([1/eps]).type_as(x)
    const float inv_eps = __fdiv_rn(1.0f, (float)eps_d);   // tensor / python scalar = tensor * (1/eps)
    const float neg_eps = -(float)eps_d;
    for (int i = tid; i < N; i += SMALL_THREADS) {
      s.hax[i] = it < 0 ? alog : __fadd_rn(alog, __fmul_rn(s.ax[i], inv_eps));
      s.hay[i] = it < 0 ? alog : __fadd_rn(alog, __fmul_rn(s.bx[i], inv_eps));   // reduces over x with b_x
    }
    for (int j = tid; j < M; j += SMALL_THREADS) {
      s.hby[j] = it < 0 ? blog : __fadd_rn(blog, __fmul_rn(s.by[j], inv_eps));
      s.hbx[j] = it < 0 ? blog : __fadd_rn(blog, __fmul_rn(s.ay[j], inv_eps));   // reduces over y with a_y
    }
    __syncthreads();
    const bool avg = it >= 0 && it < n_eps;
    for (int r = warp; r < 2 * (N + M); r += nwarps) {
      float v; float* dst;
      if (r < N) { v = softmin_warp(s.Cxx + r * N, 1, s.hax, N, P, neg_eps, lane); dst = s.ax + r; }              // a_x
      else if (r < N + M) { const int j = r - N; v = softmin_warp(s.Cyy + j * M, 1, s.hby, M, P, neg_eps, lane); dst = s.by + j; }   // b_y
      else if (r < N + 2 * M) { const int j = r - N - M; v = softmin_warp(s.Cxy + j, M, s.hay, N, P, neg_eps, lane); dst = s.ay + j; }  // a_y: over x
      else { const int i = r - N - 2 * M; v = softmin_warp(s.Cxy + i * M, 1, s.hbx, M, P, neg_eps, lane); dst = s.bx + i; }          // b_x: over y
      if (lane == 0) *dst = avg ? __fmul_rn(0.5f, __fadd_rn(*dst, v)) : v;
    }
    __syncthreads();
  }
  // sinkhorn_cost: <alpha, b_x - a_x> + <beta, a_y - b_y>
  float acc = 0.f;
  if (warp == 0) {
    const float wa = __fdiv_rn(1.0f, (float)N), wb = __fdiv_rn(1.0f, (float)M);
    float t0 = 0.f, t1 = 0.f;
    for (int i = lane; i < N; i += 32) t0 = fmaf(wa, s.bx[i] - s.ax[i], t0);
    for (int j = lane; j < M; j += 32) t1 = fmaf(wb, s.ay[j] - s.by[j], t1);
    acc = warp_sum(t0) + warp_sum(t1);
    if (lane == 0) s.red[0] = acc;
  }
  __syncthreads();
  acc = s.red[0];
  __syncthreads();
  return acc;
}

__device__ __forceinline__ void gather_cloud(float* dst, const float* pts, const int* idx, int off, int n,
                                             int d, int row_stride) {
  // rows idx[k] (or off+k when idx is null) of pts [*, d] -> dst [n][d+1]
  for (int e = threadIdx.x; e < n * d; e += SMALL_THREADS) {
    const int r = e / d, k = e % d;
    const int src = idx ? idx[r] : off + r;
    dst[r * (d + 1) + k] = pts[(size_t)src * row_stride + k];
  }
}

__global__ void __launch_bounds__(SMALL_THREADS) small_kernel(const float* pts, const SmallProblem* probs,
                                                              int d, float* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SmallSmem s = carve(smem_raw, d);
  const SmallProblem p = probs[blockIdx.x];
  gather_cloud(s.X, pts, nullptr, p.x_off, p.x_n, d, d);
  gather_cloud(s.Y, pts, nullptr, p.y_off, p.y_n, d, d);
  __syncthreads();
  const float v = small_divergence(s, p.x_n, p.y_n, d);
  if (threadIdx.x == 0) out[blockIdx.x] = v;
}

// search_dg.py:150-162 in one launch: CTA (pair, j) splits rows j::M by argmax(dc) into the domain
// clouds, computes the pair's divergence, and the last CTA of policy j adds (d12 + d13) + d23
// (the reference's summation order) to rewards[j].
__global__ void __launch_bounds__(SMALL_THREADS) reward_kernel(const float* feat, const float* dc, int n, int d,
                                                               int n_dom, int M, float* rewards,
                                                               float* pair_values, int* counters,
                                                               int* status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SmallSmem s = carve(smem_raw, d);
  const int n_pairs = n_dom * (n_dom - 1) / 2;
  const int pair = blockIdx.x, j = blockIdx.y;
  // call order of the reference for three domains: (1,2), (2,3), (1,3); otherwise lexicographic
  int da = 0, db = 1;
  if (n_dom == 3) { da = pair == 2 ? 0 : pair; db = pair == 0 ? 1 : 2; }
  else { int k = 0; for (int a = 0; a < n_dom; ++a) for (int b = a + 1; b < n_dom; ++b) { if (k == pair) { da = a; db = b; } ++k; } }
  if (threadIdx.x == 0) {
    int na = 0, nb = 0;
    for (int r = j; r < n; r += M) {
      const float* q = dc + (size_t)r * n_dom;
      int best = 0;
      for (int k = 1; k < n_dom; ++k) if (q[k] > q[best]) best = k;
      if (best == da) { if (na < SMALL_MAX) s.idx[na] = r; ++na; }
      if (best == db) { if (nb < SMALL_MAX) s.idx[SMALL_MAX + nb] = r; ++nb; }
    }
    s.misc[1] = na; s.misc[2] = nb;
  }
  __syncthreads();
  const int na = s.misc[1], nb = s.misc[2];
  float v;
  if (na > SMALL_MAX || nb > SMALL_MAX || na == 0 || nb == 0) {
    if (threadIdx.x == 0) atomicExch(status, na == 0 || nb == 0 ? 2 : 1);
    v = __int_as_float(0x7fc00000);
  } else {
    gather_cloud(s.X, feat, s.idx, 0, na, d, d);
    gather_cloud(s.Y, feat, s.idx + SMALL_MAX, 0, nb, d, d);
    __syncthreads();
    v = small_divergence(s, na, nb, d);
  }
  if (threadIdx.x == 0) {
    pair_values[j * n_pairs + pair] = v;
    __threadfence();
    const int done = atomicAdd(&counters[j], 1);
    if (done == n_pairs - 1) {
      __threadfence();
      volatile float* pv = pair_values + j * n_pairs;
      float t;
      if (n_dom == 3) t = __fadd_rn(__fadd_rn(pv[0], pv[2]), pv[1]);     // (d12 + d13) + d23
      else { t = 0.f; for (int k = 0; k < n_pairs; ++k) t = __fadd_rn(t, pv[k]); }
      rewards[j] = __fadd_rn(rewards[j], t);
      counters[j] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// large problems
// ---------------------------------------------------------------------------------------------------
// per-row L2 norms; one warp per row
__global__ void norms_kernel(const float* x, int n, int d, float* out) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  float a = 0.f;
  for (int k = lane; k < d; k += 32) { const float v = x[(size_t)r * d + k]; a = fmaf(v, v, a); }
  a = warp_sum(a);
  if (lane == 0) out[r] = sqrtf(a);
}

// coordinate-wise min/max over the rows of x and y: grid (ceil(d/32), splits); atomics on ordered ints
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__global__ void minmax_kernel(const float* x, int n, int d, int* lo, int* hi) {
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rstep = gridDim.y * (blockDim.x >> 5);
  float l = INFINITY, h = -INFINITY;
  if (k < d)
    for (int r = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += rstep) {
      const float v = x[(size_t)r * d + k];
      l = fminf(l, v); h = fmaxf(h, v);
    }
  if (k < d && l <= h) { atomicMin(&lo[k], f2ord(l)); atomicMax(&hi[k], f2ord(h)); }
}
__global__ void diameter_kernel(const int* lo, const int* hi, int d, float* out) {
  float a = 0.f;
  for (int k = threadIdx.x; k < d; k += 32) { const float t = ord2f(hi[k]) - ord2f(lo[k]); a = fmaf(t, t, a); }
  a = warp_sum(a);
  if (threadIdx.x == 0) *out = sqrtf(a);
}

// C[i][j] = 1 - <a_i, b_j> / (|a_i| |b_j|); fp32 CUDA-core tile GEMM, 64x64 tile, 4x4 per thread.
// Also writes the transpose when ct != nullptr.
constexpr int CT = 64, CK = 16;
__global__ void __launch_bounds__(256) cost_kernel(const float* a, const float* na, int n, const float* b,
                                                   const float* nb, int m, int d, float* c, int ldc,
                                                   float* ct, int ldct) {
  __shared__ float sa[CK][CT + 4], sb[CK][CT + 4];
  __shared__ float st[CT][CT + 1];
  const int i0 = blockIdx.y * CT, j0 = blockIdx.x * CT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < d; k0 += CK) {
    for (int e = threadIdx.x; e < CT * CK; e += 256) {
      const int r = e / CK, k = e % CK;
      sa[k][r] = (i0 + r < n && k0 + k < d) ? a[(size_t)(i0 + r) * d + k0 + k] : 0.f;
      sb[k][r] = (j0 + r < m && k0 + k < d) ? b[(size_t)(j0 + r) * d + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { av[u] = sa[k][ty * 4 + u]; bv[u] = sb[k][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = j0 + tx * 4 + v;
      float val = 0.f;
      if (i < n && j < m) val = __fsub_rn(1.0f, __fdiv_rn(acc[u][v], __fmul_rn(na[i], nb[j])));
      st[ty * 4 + u][tx * 4 + v] = val;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < CT * CT; e += 256) {
    const int ii = e / CT, jj = e % CT;
    if (i0 + ii < n && j0 + jj < m) c[(size_t)(i0 + ii) * ldc + j0 + jj] = st[ii][jj];
  }
  if (ct) {
    for (int e = threadIdx.x; e < CT * CT; e += 256) {
      const int jj = e / CT, ii = e % CT;     // ct[j][i], i fastest
      if (j0 + jj < m && i0 + ii < n) ct[(size_t)(j0 + jj) * ldct + i0 + ii] = st[ii][jj];
    }
  }
}

// fp32 -> three bf16 terms (hi + mid + lo reproduces x to ~2^-24), written in the two arrangements whose dot
// product sums the six significant cross terms:  A = [hi, hi, mid, hi, lo, mid],  B = [hi, mid, hi, lo, hi, mid]
__global__ void split3_kernel(const float* x, long long total, int d, __nv_bfloat16* arr_a, __nv_bfloat16* arr_b) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / d;
    const int k = (int)(e - r * d);
    const float v = x[e];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    __nv_bfloat16* pa = arr_a + r * 6 * d + k;
    __nv_bfloat16* pb = arr_b + r * 6 * d + k;
    pa[0] = hi; pa[d] = hi; pa[2 * d] = mid; pa[3 * d] = hi; pa[4 * d] = lo; pa[5 * d] = mid;
    pb[0] = hi; pb[d] = mid; pb[2 * d] = hi; pb[3 * d] = lo; pb[4 * d] = hi; pb[5 * d] = mid;
  }
}

// One soft-min family: out[c] = LSE_r( h[r] - P * C[r][c] ) over the rows of a row band.
struct LseMat {
  const float* C;     // [rows][ld]
  const float* h;     // [rows], already in log2 units: (log w + pot/eps) * log2(e)
  float2* part;       // [splits][cols] running (max, sum) in log2 units
  int rows, cols, ld, splits;
};
struct LseArgs {
  LseMat m[4];
  float P2;           // P * log2(e)
};

constexpr int LSE_THREADS = 256;
constexpr int LSE_RB = 8;   // rows per register block

__global__ void __launch_bounds__(LSE_THREADS) lse_kernel(const LseArgs a) {
  const LseMat& M = a.m[blockIdx.z];
  const int c0 = (blockIdx.x * LSE_THREADS + threadIdx.x) * 4;
  if (blockIdx.y >= M.splits || blockIdx.x * LSE_THREADS * 4 >= M.cols) return;
  const int band = (M.rows + M.splits - 1) / M.splits;
  const int r0 = blockIdx.y * band, r1 = min(r0 + band, M.rows);
  __shared__ float sh[LSE_RB * 64];
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, s[4] = {0.f, 0.f, 0.f, 0.f};
  const bool active = c0 < M.cols;      // cols is padded to a multiple of 4 through ld
  const float nP = -a.P2;
  for (int rb = r0; rb < r1; rb += LSE_RB * 64) {
    // stage h for up to 512 rows
    __syncthreads();
    for (int i = threadIdx.x; i < LSE_RB * 64 && rb + i < r1; i += LSE_THREADS) sh[i] = M.h[rb + i];
    __syncthreads();
    const int rend = min(rb + LSE_RB * 64, r1);
    if (active) {
      for (int r = rb; r < rend; r += LSE_RB) {
        float4 v[LSE_RB];
#pragma unroll
        for (int u = 0; u < LSE_RB; ++u) {
          const int rr = r + u;
          if (rr < rend) v[u] = __ldcs((const float4*)(M.C + (size_t)rr * M.ld + c0));
          else v[u] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        }
        float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int u = 0; u < LSE_RB; ++u) {
          const float hh = (r + u < rend) ? sh[r + u - rb] : -INFINITY;
          v[u].x = fmaf(v[u].x, nP, hh); v[u].y = fmaf(v[u].y, nP, hh);
          v[u].z = fmaf(v[u].z, nP, hh); v[u].w = fmaf(v[u].w, nP, hh);
          if (r + u >= rend) v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          bm[0] = fmaxf(bm[0], v[u].x); bm[1] = fmaxf(bm[1], v[u].y);
          bm[2] = fmaxf(bm[2], v[u].z); bm[3] = fmaxf(bm[3], v[u].w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (bm[q] > m[q]) { s[q] *= exp2f(m[q] - bm[q]); m[q] = bm[q]; }
#pragma unroll
        for (int u = 0; u < LSE_RB; ++u) {
          s[0] += exp2f(v[u].x - m[0]); s[1] += exp2f(v[u].y - m[1]);
          s[2] += exp2f(v[u].z - m[2]); s[3] += exp2f(v[u].w - m[3]);
        }
      }
    }
  }
  if (active) {
    float2* p = M.part + (size_t)blockIdx.y * M.cols + c0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (c0 + q < M.cols) p[q] = make_float2(m[q], s[q]);
  }
}

// merge the band partials, update the potentials and prepare the next iteration's h vectors
struct CombineVec {
  const float2* part;   // [splits][n]
  float* pot;           // potential being updated, [n]
  float* h_next;        // h vector this potential feeds next, [n]
  int n, splits;
  float logw;           // log weight of the measure h_next reduces over
};
struct CombineArgs {
  CombineVec v[4];
  float neg_eps;        // -(float)eps of this iteration
  float inv_eps_next;   // 1/(float)eps of the next iteration
  int average;          // 1: pot = .5*(pot + new); 0: pot = new
};
__global__ void combine_kernel(const CombineArgs a) {
  const CombineVec& V = a.v[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V.n) return;
  float m = -INFINITY, s = 0.f;
  for (int k = 0; k < V.splits; ++k) {
    const float2 p = V.part[(size_t)k * V.n + i];
    if (p.x > m) { s = s * exp2f(m - p.x) + p.y; m = p.x; }
    else if (p.x > -INFINITY) s += p.y * exp2f(p.x - m);
  }
  const float lse = (m + log2f(s)) * LN2;
  const float nv = a.neg_eps * lse;
  const float pot = a.average ? 0.5f * (V.pot[i] + nv) : nv;
  V.pot[i] = pot;
  V.h_next[i] = (V.logw + pot * a.inv_eps_next) * LOG2E;
}
__global__ void fill_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// <alpha, b_x - a_x> + <beta, a_y - b_y>, one CTA
__global__ void final_kernel(const float* ax, const float* bx, int n, const float* ay, const float* by,
                             int m, float* out) {
  __shared__ float red[32];
  float t = 0.f;
  const float wa = 1.0f / (float)n, wb = 1.0f / (float)m;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t = fmaf(wa, bx[i] - ax[i], t);
  for (int j = threadIdx.x; j < m; j += blockDim.x) t = fmaf(wb, ay[j] - by[j], t);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) *out = t;
  }
}

struct LargeLayout {
  size_t cxx, cyy, cxy, cyx, nx, ny, lo, hi, diam, pot, h, part, arr, total;
  int ldn, ldm, splits_n, splits_m;
};
static int pick_splits(int rows, int cols) {
  const int col_tiles = (cols + LSE_THREADS * 4 - 1) / (LSE_THREADS * 4);
  int s = (148 * 4 + col_tiles * 4 - 1) / (col_tiles * 4);     // ~4 CTAs per SM over the 4 matrices
  s = std::max(1, std::min(s, (rows + 63) / 64));
  return s;
}
static bool use_tc(int dim) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AADG_SINKHORN_TC"); v = e ? atoi(e) : 1; }
  return v && dim % 8 == 0;
}
static LargeLayout large_layout(int n, int m, int dim = 0) {
  LargeLayout L{};
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = align_up(off, 256); off = o + b; return o; };
  L.ldn = (n + 3) & ~3; L.ldm = (m + 3) & ~3;
  L.splits_n = pick_splits(n, std::max(n, m));   // bands over x rows
  L.splits_m = pick_splits(m, std::max(n, m));   // bands over y rows
  L.cxx = take(sizeof(float) * (size_t)n * L.ldn);
  L.cyy = take(sizeof(float) * (size_t)m * L.ldm);
  L.cxy = take(sizeof(float) * (size_t)n * L.ldm);
  L.cyx = take(sizeof(float) * (size_t)m * L.ldn);
  L.nx = take(sizeof(float) * n); L.ny = take(sizeof(float) * m);
  L.lo = take(sizeof(int) * 4096); L.hi = take(sizeof(int) * 4096);
  L.diam = take(256);
  L.pot = take(sizeof(float) * 2 * ((size_t)n + m));      // a_x, b_x [n]; b_y, a_y [m]
  L.h = take(sizeof(float) * 2 * ((size_t)n + m));        // hax, hay [n]; hby, hbx [m]
  const int smax = std::max(L.splits_n, L.splits_m);
  L.part = take(sizeof(float2) * (size_t)smax * 2 * ((size_t)n + m));
  L.arr = take(dim > 0 && use_tc(dim) ? sizeof(__nv_bfloat16) * 2 * 6 * (size_t)dim * ((size_t)n + m) : 0);
  L.total = align_up(off, 256);
  return L;
}

}  // namespace sk
}  // namespace aadg

using namespace aadg;
using namespace aadg::sk;

extern "C" {

/* points per cloud the single-launch shared-memory path accepts */
int aadg_sinkhorn_small_max_points(void) { return SMALL_MAX; }

int aadg_sinkhorn_small_batched(const float* points, const int32_t* problems, int n_problems, int dim,
                                float* out, void* stream) {
  AADG_REQUIRE(n_problems >= 0 && dim > 0 && dim <= 512, "bad n_problems=%d / dim=%d (dim <= 512)", n_problems, dim);
  if (n_problems == 0) return AADG_OK;
  AADG_REQUIRE(points && problems && out, "null pointer");
  const size_t smem = small_smem_bytes(dim);
  AADG_REQUIRE(smem <= 227 * 1024, "dim %d needs %zu bytes of shared memory", dim, smem);
  AADG_CUDA_TRY(cudaFuncSetAttribute(small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  small_kernel<<<n_problems, SMALL_THREADS, smem, (cudaStream_t)stream>>>(points, (const SmallProblem*)problems, dim, out);
  return check_launch("sinkhorn small kernel");
}

size_t aadg_sinkhorn_rewards_workspace_bytes(int n_policies, int n_domains) {
  if (n_policies <= 0 || n_domains < 2) return 0;
  return align_up(sizeof(int) * (size_t)(n_policies + 1), 256);
}

int aadg_sinkhorn_diversity_rewards(const float* features, const float* domain_code, int n_rows, int dim,
                                    int n_domains, int n_policies, float* rewards, float* pair_values,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  AADG_REQUIRE(n_rows > 0 && dim > 0 && dim <= 512 && n_domains >= 2 && n_domains <= 8 && n_policies > 0,
               "bad sizes n_rows=%d dim=%d n_domains=%d n_policies=%d", n_rows, dim, n_domains, n_policies);
  AADG_REQUIRE(features && domain_code && rewards && pair_values, "null pointer");
  if (workspace_bytes < aadg_sinkhorn_rewards_workspace_bytes(n_policies, n_domains) || !workspace) {
    set_error("workspace too small");
    return AADG_ENOSPC;
  }
  const size_t smem = small_smem_bytes(dim);
  AADG_REQUIRE(smem <= 227 * 1024, "dim %d needs %zu bytes of shared memory", dim, smem);
  cudaStream_t st = (cudaStream_t)stream;
  // counters [n_policies] must be zero on entry; the kernel leaves them zero.  status at the end.
  int* counters = (int*)workspace;
  AADG_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(int) * (n_policies + 1), st));
  AADG_CUDA_TRY(cudaFuncSetAttribute(reward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_domains * (n_domains - 1) / 2, n_policies);
  reward_kernel<<<grid, SMALL_THREADS, smem, st>>>(features, domain_code, n_rows, dim, n_domains, n_policies,
                                                  rewards, pair_values, counters, counters + n_policies);
  return check_launch("sinkhorn reward kernel");
}

size_t aadg_sinkhorn_large_workspace_bytes(int n, int m, int dim) {
  if (n <= 0 || m <= 0 || dim <= 0) return 0;
  return large_layout(n, m, dim).total;
}

/* One divergence on big clouds.  Synchronises `stream` once (the diameter, like the reference's
 * .item()) unless diameter > 0 is supplied.  n_iterations_out (HOST, optional) receives the number
 * of epsilon steps (soft-min sweeps = n + 2). */
static int sinkhorn_large_impl(const float* x, int n, const float* y, int m, int dim, float diameter, float* out,
                               int* n_iterations_out, void* workspace, size_t workspace_bytes, void* stream,
                               int cost_only);

int aadg_sinkhorn_large(const float* x, int n, const float* y, int m, int dim, float diameter, float* out,
                        int* n_iterations_out, void* workspace, size_t workspace_bytes, void* stream) {
  return sinkhorn_large_impl(x, n, y, m, dim, diameter, out, n_iterations_out, workspace, workspace_bytes, stream, 0);
}

/* Only the set-up part of aadg_sinkhorn_large (norms, diameter, the four cost matrices): lets a benchmark
 * separate the one-off cost build from the epsilon iterations. */
int aadg_sinkhorn_large_setup(const float* x, int n, const float* y, int m, int dim, float diameter,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return sinkhorn_large_impl(x, n, y, m, dim, diameter, nullptr, nullptr, workspace, workspace_bytes, stream, 1);
}

static int sinkhorn_large_impl(const float* x, int n, const float* y, int m, int dim, float diameter, float* out,
                               int* n_iterations_out, void* workspace, size_t workspace_bytes, void* stream,
                               int cost_only) {
  AADG_REQUIRE(n > 0 && m > 0 && dim > 0 && dim <= 4096, "bad sizes n=%d m=%d dim=%d", n, m, dim);
  AADG_REQUIRE(x && y && (out || cost_only), "null pointer");
  const LargeLayout L = large_layout(n, m, dim);
  if (!workspace || workspace_bytes < L.total) {
    set_error("workspace too small: need %zu bytes, got %zu", L.total, workspace_bytes);
    return AADG_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  float *cxx = (float*)(w + L.cxx), *cyy = (float*)(w + L.cyy), *cxy = (float*)(w + L.cxy), *cyx = (float*)(w + L.cyx);
  float *nx = (float*)(w + L.nx), *ny = (float*)(w + L.ny);
  float* pot = (float*)(w + L.pot);
  float *ax = pot, *bx = pot + n, *by = pot + 2 * (size_t)n, *ay = by + m;
  float* hv = (float*)(w + L.h);
  float *hax = hv, *hay = hv + n, *hby = hv + 2 * (size_t)n, *hbx = hby + m;
  float2* part = (float2*)(w + L.part);
  const int smax = std::max(L.splits_n, L.splits_m);
  float2 *p_ax = part, *p_bx = part + (size_t)smax * n, *p_by = part + (size_t)smax * 2 * n,
         *p_ay = p_by + (size_t)smax * m;

  norms_kernel<<<(n + 7) / 8, 256, 0, st>>>(x, n, dim, nx);
  norms_kernel<<<(m + 7) / 8, 256, 0, st>>>(y, m, dim, ny);
  double diam = diameter;
  if (!(diameter > 0.f)) {
    int* lo = (int*)(w + L.lo); int* hi = (int*)(w + L.hi);
    AADG_CUDA_TRY(cudaMemsetAsync(lo, 0x7f, sizeof(int) * dim, st));   // large positive ordered ints
    AADG_CUDA_TRY(cudaMemsetAsync(hi, 0x80, sizeof(int) * dim, st));   // large negative
    dim3 g((dim + 31) / 32, 64);
    minmax_kernel<<<g, 256, 0, st>>>(x, n, dim, lo, hi);
    minmax_kernel<<<g, 256, 0, st>>>(y, m, dim, lo, hi);
    diameter_kernel<<<1, 32, 0, st>>>(lo, hi, dim, (float*)(w + L.diam));
    float hd = 0.f;
    AADG_CUDA_TRY(cudaMemcpyAsync(&hd, w + L.diam, sizeof(float), cudaMemcpyDeviceToHost, st));
    AADG_CUDA_TRY(cudaStreamSynchronize(st));
    diam = hd;
  }
  AADG_REQUIRE(diam > 0 && std::isfinite(diam), "degenerate clouds: diameter %g", diam);
  double eps[MAX_EPS];
  const int n_eps = eps_schedule(diam, eps);
  if (n_iterations_out) *n_iterations_out = n_eps;

  if (use_tc(dim) && n >= 128 && m >= 128) {
    // tensor-core cost build: bf16x3 split of the features, fp32 accumulation (tcgen05), ~fp32 accuracy
    __nv_bfloat16* xa = (__nv_bfloat16*)(w + L.arr);
    __nv_bfloat16* xb = xa + (size_t)6 * dim * n;
    __nv_bfloat16* ya = xb + (size_t)6 * dim * n;
    __nv_bfloat16* yb = ya + (size_t)6 * dim * m;
    split3_kernel<<<148 * 8, 256, 0, st>>>(x, (long long)n * dim, dim, xa, xb);
    split3_kernel<<<148 * 8, 256, 0, st>>>(y, (long long)m * dim, dim, ya, yb);
    int rc2 = tc::gram_cost(xa, n, xb, n, 6 * dim, nx, nx, cxx, L.ldn, nullptr, 0, st);
    if (!rc2) rc2 = tc::gram_cost(ya, m, yb, m, 6 * dim, ny, ny, cyy, L.ldm, nullptr, 0, st);
    if (!rc2) rc2 = tc::gram_cost(xa, n, yb, m, 6 * dim, nx, ny, cxy, L.ldm, cyx, L.ldn, st);
    if (rc2) return rc2;
  } else {
    dim3 g((n + CT - 1) / CT, (n + CT - 1) / CT);
    cost_kernel<<<g, 256, 0, st>>>(x, nx, n, x, nx, n, dim, cxx, L.ldn, nullptr, 0);
    dim3 g2((m + CT - 1) / CT, (m + CT - 1) / CT);
    cost_kernel<<<g2, 256, 0, st>>>(y, ny, m, y, ny, m, dim, cyy, L.ldm, nullptr, 0);
    dim3 g3((m + CT - 1) / CT, (n + CT - 1) / CT);
    cost_kernel<<<g3, 256, 0, st>>>(x, nx, n, y, ny, m, dim, cxy, L.ldm, cyx, L.ldn);
  }
  int rc = check_launch("sinkhorn cost kernels");
  if (rc || cost_only) return rc;
  const float alog = logf(1.0f / (float)n), blog = logf(1.0f / (float)m);
  fill_kernel<<<(2 * n + 255) / 256, 256, 0, st>>>(hax, 2 * n, alog * LOG2E);
  fill_kernel<<<(2 * m + 255) / 256, 256, 0, st>>>(hby, 2 * m, blog * LOG2E);

  LseArgs la{};
  // a_x: rows x, cols x, C_xx, h = hax;  b_y: C_yy, hby;  a_y: rows x, cols y, C_xy, h = hay;
  // b_x: rows y, cols x, C_yx, h = hbx
  la.m[0] = LseMat{cxx, hax, p_ax, n, n, L.ldn, L.splits_n};
  la.m[1] = LseMat{cyy, hby, p_by, m, m, L.ldm, L.splits_m};
  la.m[2] = LseMat{cxy, hay, p_ay, n, m, L.ldm, L.splits_n};
  la.m[3] = LseMat{cyx, hbx, p_bx, m, n, L.ldn, L.splits_m};
  CombineArgs ca{};
  ca.v[0] = CombineVec{p_ax, ax, hax, n, L.splits_n, alog};   // a_x feeds hax (reduces over x)
  ca.v[1] = CombineVec{p_by, by, hby, m, L.splits_m, blog};
  ca.v[2] = CombineVec{p_ay, ay, hbx, m, L.splits_n, blog};   // a_y feeds hbx (reduces over y)
  ca.v[3] = CombineVec{p_bx, bx, hay, n, L.splits_m, alog};   // b_x feeds hay (reduces over x)
  const int cmax = std::max(n, m);
  dim3 lgrid((cmax + LSE_THREADS * 4 - 1) / (LSE_THREADS * 4), smax, 4);
  dim3 cgrid((cmax + 255) / 256, 4);
  for (int it = -1; it <= n_eps; ++it) {
    const double e = eps[it < 0 ? 0 : (it >= n_eps ? n_eps - 1 : it)];
    const double e_next = eps[it + 1 < 0 ? 0 : (it + 1 >= n_eps ? n_eps - 1 : it + 1)];
    la.P2 = (float)(1.0 / e) * LOG2E;
    lse_kernel<<<lgrid, LSE_THREADS, 0, st>>>(la);
    ca.neg_eps = -(float)e;
    ca.inv_eps_next = 1.0f / (float)e_next;
    ca.average = (it >= 0 && it < n_eps) ? 1 : 0;
    combine_kernel<<<cgrid, 256, 0, st>>>(ca);
  }
  final_kernel<<<1, 1024, 0, st>>>(ax, bx, n, ay, by, m, out);
  return check_launch("sinkhorn large kernels");
}

}  // extern "C"
