// The policy controller (models/controller.py:9-145) as two kernels instead of ~40 LSTMCell / Linear / softmax /
// multinomial launches per call: one launch walks all Q*L*2 decisions of every policy in the batch (sampling or
// re-evaluating given actions), one launch runs the whole back-propagation through time of
// sum_t log pi(a_t) for the PPO update (losses.py:127-157).  Trivial FLOPs, pure latency on the reference path.
//
//   per sub-policy q (state reset, controller.py:81):  for j < L:
//       h,c = LSTMCell(x, (h,c));  op  ~ softmax(C*tanh(W_op h + b_op)/T);   x = embedding[op]
//       h,c = LSTMCell(x, (h,c));  mag ~ softmax(C*tanh(W_mag h + b_mag)/T); x = embedding[n_ops + mag]
//
// Sampling is counter based: decision t of batch row m uses Philox4x32-10(key = seed, counter = (m, t, call lo,
// call hi)) -> u in [0,1) with 24 bits, action = first k with cumsum(p)[k] > u.  One CTA per batch row.
#include "common.cuh"

namespace aadg {
namespace ctl {

constexpr int MAX_H = 128, MAX_E = 64, MAX_V = 32, MAX_STEPS = 64, THREADS = 128;

struct Params {
  const float *emb, *w_ih, *w_hh, *b_ih, *b_hh, *w_op, *b_op, *w_mag, *b_mag;
};
struct Grads {
  float *emb, *w_ih, *w_hh, *b_ih, *b_hh, *w_op, *b_op, *w_mag, *b_mag;
};
struct Dims {
  int n_ops, n_mags, Q, L, E, H;
  float C, T;
};

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
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// logits of one head from h (shared) -> softmax(C*tanh(.)/T) probabilities and tanh values, by the first warp
__device__ void head_probs(const float* w, const float* b, int V, int H, const float* h, float C, float T, float* th,
                           float* p, float* logp) {
  const int k = threadIdx.x;
  if (k < 32) {
    float z = -INFINITY;
    if (k < V) {
      float acc = b[k];
      for (int j = 0; j < H; ++j) acc = fmaf(w[k * H + j], h[j], acc);
      const float t = tanhf(acc);
      th[k] = t;
      z = C * t / T;
    }
    const float mx = warp_max(z);
    const float e = k < V ? expf(z - mx) : 0.f;
    const float s = warp_sum(e);
    if (k < V) {
      logp[k] = (z - mx) - logf(s);
      p[k] = e / s;
    }
  }
}

// saved activations per (row, step): i, f, g, o, c, h  (6*H floats)
__global__ void __launch_bounds__(THREADS) walk_kernel(const Params P, const Dims D, int mode, unsigned long long seed,
                                                       unsigned long long call, long long* policies, float* step_logp,
                                                       float* step_ent, float* step_probs, float* saved) {
  __shared__ float x[MAX_E], h[MAX_H], c[MAX_H], gates[4 * MAX_H], th[MAX_V], p[MAX_V], logp[MAX_V];
  __shared__ int s_action;
  const int m = blockIdx.x, tid = threadIdx.x;
  const int steps = D.Q * D.L * 2, H = D.H, E = D.E;
  const int VM = max(D.n_ops, D.n_mags);
  for (int t = 0; t < steps; ++t) {
    const int in_chain = t % (2 * D.L);
    if (in_chain == 0) {           // controller.py:81: fresh input and state for every sub-policy
      for (int i = tid; i < E; i += THREADS) x[i] = 0.f;
      for (int i = tid; i < H; i += THREADS) { h[i] = 0.f; c[i] = 0.f; }
    }
    __syncthreads();
    if (tid < H) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int r = g * H + tid;
        float acc = P.b_ih[r] + P.b_hh[r];
        const float* wi = P.w_ih + (size_t)r * E;
        for (int k = 0; k < E; ++k) acc = fmaf(wi[k], x[k], acc);
        const float* wh = P.w_hh + (size_t)r * H;
        for (int k = 0; k < H; ++k) acc = fmaf(wh[k], h[k], acc);
        gates[r] = acc;
      }
    }
    __syncthreads();
    if (tid < H) {
      const float ig = sigmoidf_(gates[tid]), fg = sigmoidf_(gates[H + tid]);
      const float gg = tanhf(gates[2 * H + tid]), og = sigmoidf_(gates[3 * H + tid]);
      const float cn = fg * c[tid] + ig * gg;
      const float hn = og * tanhf(cn);
      c[tid] = cn; h[tid] = hn;
      if (saved) {
        float* s = saved + ((size_t)m * steps + t) * 6 * H;
        s[tid] = ig; s[H + tid] = fg; s[2 * H + tid] = gg; s[3 * H + tid] = og; s[4 * H + tid] = cn; s[5 * H + tid] = hn;
      }
    }
    __syncthreads();
    const bool is_mag = t & 1;
    const int V = is_mag ? D.n_mags : D.n_ops;
    head_probs(is_mag ? P.w_mag : P.w_op, is_mag ? P.b_mag : P.b_op, V, H, h, D.C, D.T, th, p, logp);
    __syncthreads();
    if (tid == 0) {
      int a;
      if (mode == 0) {
        const uint4 r = philox(make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)),
                               make_uint4((unsigned int)m, (unsigned int)t, (unsigned int)call, (unsigned int)(call >> 32)));
        const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);
        float cum = 0.f;
        a = V - 1;
        for (int k = 0; k < V; ++k) {
          cum += p[k];
          if (cum > u) { a = k; break; }
        }
        policies[(size_t)m * steps + t] = a;
      } else {
        a = (int)policies[(size_t)m * steps + t];
        a = min(max(a, 0), V - 1);
      }
      s_action = a;
      float ent = 0.f;
      for (int k = 0; k < V; ++k) ent -= logp[k] * p[k];
      step_logp[(size_t)m * steps + t] = logp[a];
      step_ent[(size_t)m * steps + t] = ent;
    }
    if (step_probs && tid < VM) step_probs[((size_t)m * steps + t) * VM + tid] = tid < V ? p[tid] : 0.f;
    __syncthreads();
    const float* row = P.emb + (size_t)((is_mag ? D.n_ops : 0) + s_action) * E;
    for (int i = tid; i < E; i += THREADS) x[i] = row[i];
  }
}

// d(sum_t logp)/d(parameters) * grad[m], accumulated into G; one CTA per batch row, back-propagation through time
// inside every sub-policy chain.  da[steps][4H] (gate pre-activation gradients) lives in dynamic shared memory.
__global__ void __launch_bounds__(THREADS) backward_kernel(const Params P, const Dims D, const long long* policies,
                                                           const float* saved, const float* grad, Grads G) {
  extern __shared__ float da_all[];                      // [steps][4H]
  __shared__ float h[MAX_H], dh[MAX_H], dc[MAX_H], dh_prev[MAX_H], th[MAX_V], p[MAX_V], logp[MAX_V], dl[MAX_V];
  const int m = blockIdx.x, tid = threadIdx.x;
  const int steps = D.Q * D.L * 2, H = D.H, E = D.E, chain = 2 * D.L;
  const float gm = grad[m];
  const float* sv = saved + (size_t)m * steps * 6 * H;
  const long long* pol = policies + (size_t)m * steps;
  for (int t = steps - 1; t >= 0; --t) {
    const int in_chain = t % chain;
    if (in_chain == chain - 1)                           // last step of a chain: nothing flows in from the future
      for (int i = tid; i < H; i += THREADS) { dh[i] = 0.f; dc[i] = 0.f; }
    const float* s = sv + (size_t)t * 6 * H;
    for (int i = tid; i < H; i += THREADS) h[i] = s[5 * H + i];
    __syncthreads();
    const bool is_mag = t & 1;
    const int V = is_mag ? D.n_mags : D.n_ops;
    const float* w = is_mag ? P.w_mag : P.w_op;
    head_probs(w, is_mag ? P.b_mag : P.b_op, V, H, h, D.C, D.T, th, p, logp);
    __syncthreads();
    const int a = min(max((int)pol[t], 0), V - 1);
    if (tid < V) {
      const float dz = gm * ((tid == a ? 1.f : 0.f) - p[tid]);
      const float d = dz * (D.C / D.T) * (1.f - th[tid] * th[tid]);
      dl[tid] = d;
      atomicAdd((is_mag ? G.b_mag : G.b_op) + tid, d);
    }
    __syncthreads();
    float* gw = is_mag ? G.w_mag : G.w_op;
    for (int i = tid; i < V * H; i += THREADS) atomicAdd(gw + i, dl[i / H] * h[i % H]);
    if (tid < H) {
      float acc = dh[tid];
      for (int k = 0; k < V; ++k) acc = fmaf(w[k * H + tid], dl[k], acc);
      const float ig = s[tid], fg = s[H + tid], gg = s[2 * H + tid], og = s[3 * H + tid], cn = s[4 * H + tid];
      const float cprev = in_chain == 0 ? 0.f : sv[(size_t)(t - 1) * 6 * H + 4 * H + tid];
      const float tc = tanhf(cn);
      const float dcn = acc * og * (1.f - tc * tc) + dc[tid];
      float* da = da_all + (size_t)t * 4 * H;
      da[tid] = dcn * gg * ig * (1.f - ig);
      da[H + tid] = dcn * cprev * fg * (1.f - fg);
      da[2 * H + tid] = dcn * ig * (1.f - gg * gg);
      da[3 * H + tid] = acc * tc * og * (1.f - og);
      dc[tid] = dcn * fg;
    }
    __syncthreads();
    const float* da = da_all + (size_t)t * 4 * H;
    if (tid < H) {                                        // gradient reaching the previous hidden state
      float acc = 0.f;
      for (int r = 0; r < 4 * H; ++r) acc = fmaf(P.w_hh[(size_t)r * H + tid], da[r], acc);
      dh_prev[tid] = acc;
    }
    if (in_chain != 0 && tid < E) {                       // gradient of this step's input = an embedding row
      float acc = 0.f;
      for (int r = 0; r < 4 * H; ++r) acc = fmaf(P.w_ih[(size_t)r * E + tid], da[r], acc);
      const int prev_a = (int)pol[t - 1];
      const int rowi = ((t - 1) & 1 ? D.n_ops : 0) + prev_a;
      atomicAdd(G.emb + (size_t)rowi * E + tid, acc);
    }
    __syncthreads();
    for (int i = tid; i < H; i += THREADS) dh[i] = dh_prev[i];
    __syncthreads();
  }
  // weight gradients: sum over steps of da_t (x) [x_t ; h_{t-1}]  (chain starts have x = 0, h = 0)
  for (int i = tid; i < 4 * H * (E + H); i += THREADS) {
    const int r = i / (E + H), k = i - r * (E + H);
    float acc = 0.f;
    for (int t = 0; t < steps; ++t) {
      if (t % chain == 0) continue;
      float in;
      if (k < E) {
        const int rowi = ((t - 1) & 1 ? D.n_ops : 0) + (int)pol[t - 1];
        in = P.emb[(size_t)rowi * E + k];
      } else {
        in = sv[(size_t)(t - 1) * 6 * H + 5 * H + (k - E)];
      }
      acc = fmaf(da_all[(size_t)t * 4 * H + r], in, acc);
    }
    if (k < E) atomicAdd(G.w_ih + (size_t)r * E + k, acc);
    else atomicAdd(G.w_hh + (size_t)r * H + (k - E), acc);
  }
  for (int r = tid; r < 4 * H; r += THREADS) {
    float acc = 0.f;
    for (int t = 0; t < steps; ++t) acc += da_all[(size_t)t * 4 * H + r];
    atomicAdd(G.b_ih + r, acc);
    atomicAdd(G.b_hh + r, acc);
  }
}

static int check_dims(const Dims& d, int batch) {
  AADG_REQUIRE(batch > 0 && d.Q > 0 && d.L > 0 && d.Q * d.L * 2 <= MAX_STEPS, "controller: Q*L*2 must be 1..%d", MAX_STEPS);
  AADG_REQUIRE(d.H > 0 && d.H <= MAX_H && d.E > 0 && d.E <= MAX_E, "controller: hidden <= %d, embedding <= %d", MAX_H, MAX_E);
  AADG_REQUIRE(d.n_ops > 0 && d.n_ops <= MAX_V && d.n_mags > 0 && d.n_mags <= MAX_V, "controller: at most %d choices per head", MAX_V);
  AADG_REQUIRE(d.T > 0.f, "controller: temperature must be positive");
  return AADG_OK;
}

}  // namespace ctl
}  // namespace aadg

using namespace aadg;
using namespace aadg::ctl;

extern "C" {

/* One launch for Controller.sample (mode 0: policies written) / Controller.evaluate (mode 1: policies read),
 * models/controller.py:73-145.  params: the nine parameter tensors in module order (embedding.weight [V,E],
 * lstm.weight_ih [4H,E], lstm.weight_hh [4H,H], lstm.bias_ih, lstm.bias_hh [4H], outop.weight [n_ops,H], outop.bias,
 * outmag.weight [n_mags,H], outmag.bias).  Outputs per (row, decision): step_log_prob, step_entropy [batch, Q*L*2];
 * step_probs [batch, Q*L*2, max(n_ops,n_mags)] (may be NULL); saved [batch, Q*L*2, 6H] activations for
 * aadg_controller_backward (may be NULL).  (seed, call) key the Philox stream of mode 0. */
int aadg_controller_walk(const float* const* params, int n_ops, int n_mags, int q, int l, int e, int h, float c, float t,
                         int batch, int mode, unsigned long long seed, unsigned long long call, long long* policies,
                         float* step_log_prob, float* step_entropy, float* step_probs, float* saved, void* stream) {
  Dims d{n_ops, n_mags, q, l, e, h, c, t};
  int rc = check_dims(d, batch);
  if (rc) return rc;
  AADG_REQUIRE(params && policies && step_log_prob && step_entropy, "controller: null buffer");
  Params P{params[0], params[1], params[2], params[3], params[4], params[5], params[6], params[7], params[8]};
  walk_kernel<<<batch, THREADS, 0, (cudaStream_t)stream>>>(P, d, mode, seed, call, policies, step_log_prob, step_entropy,
                                                         step_probs, saved);
  return check_launch("controller walk");
}

/* grads[i] += d( sum_m grad_log_prob[m] * sum_t log pi(a_t | m) ) / d params[i]  (fp32, accumulated: zero first);
 * `saved` and `policies` from the aadg_controller_walk call being differentiated. */
int aadg_controller_backward(const float* const* params, int n_ops, int n_mags, int q, int l, int e, int h, float c,
                             float t, int batch, const long long* policies, const float* saved,
                             const float* grad_log_prob, float* const* grads, void* stream) {
  Dims d{n_ops, n_mags, q, l, e, h, c, t};
  int rc = check_dims(d, batch);
  if (rc) return rc;
  AADG_REQUIRE(params && grads && policies && saved && grad_log_prob, "controller: null buffer");
  Params P{params[0], params[1], params[2], params[3], params[4], params[5], params[6], params[7], params[8]};
  Grads G{grads[0], grads[1], grads[2], grads[3], grads[4], grads[5], grads[6], grads[7], grads[8]};
  const size_t smem = (size_t)q * l * 2 * 4 * h * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    AADG_CUDA_TRY(cudaFuncSetAttribute(backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  backward_kernel<<<batch, THREADS, smem, (cudaStream_t)stream>>>(P, d, policies, saved, grad_log_prob, G);
  return check_launch("controller backward");
}

}  // extern "C"
