// Ensemble training step on the device (include/simba_b200.h "ensemble training step"):
// replaces MlpEnsemble.training_step / validation_step / fit's inner loop
// (simba/models/mlp_ensemble.py:134-187). fp32 throughout — the reference trains in fp32 and the
// planner's fp32 and bf16 weight images are both derived from these master weights.
//
// Shape of the work: E members x batch 64 x a 4x128 MLP is 28 MFLOP per member-step, far below
// what one launch can hide, so the step is latency-bound. The design goal is therefore few, wide
// launches that stay inside one CUDA graph:
//   forward   L+1 launches   Y = act([X, 1] . theta_l)            grid (N/32, rows/64, E)
//   nll       1 launch       loss, d(mu), d(raw var), lr_t         grid (rows/64, E)
//   backward  L launches     dZ_{l-1} = relu'(H) * (dZ_l . W_l^T)  grid (K/32, rows/64, E)
//   update    1 launch       dtheta = [H, 1]^T . dZ for EVERY layer, clip, Adam, in one grid
// Parameters of train layer l are stored as one [(K_l + 1) x N_l] row-major block (Keras kernel
// [in, out] followed by the bias row), so "bias" is just the row that multiplies the constant 1
// and the weight-gradient GEMM produces the bias gradient as its last row. The Gaussian head's
// two Dense layers (mlp_ensemble.py:28-30) are one block with N = 2 * O (mu columns, then var).
// Reductions are in a fixed order: results are bit-reproducible run to run.
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "internal.h"

using namespace simba;

namespace {

constexpr int kTileM = 64;     // rows per CTA in forward / backward
constexpr int kTileN = 32;     // output columns per CTA
constexpr int kTileK = 32;     // contraction chunk staged in shared memory
constexpr int kThreads = 256;

struct TrainState {
  int iterations;        // optimizer.iterations
  int fit_step;          // step index inside the running fit()
  float lr_t;            // lr(iterations) * sqrt(1 - beta2^t) / (1 - beta1^t) of the current step
  float loss;
  unsigned nll_ticket;
  unsigned upd_ticket;
};

struct FitDesc {
  const float* inputs;
  const float* targets;
  const int* batch_index;   // [steps, E, bmax]
  const int* batch_rows;    // [steps] or null
  float* losses;            // [steps] or null
  int bmax;
};

struct OptParams {
  float lr0, beta1, beta2, epsilon, clipvalue;
  int schedule, steps_per_epoch, train_epochs;
};

__device__ __forceinline__ int resolve_rows(const FitDesc* desc, const TrainState* st, int rows_fixed) {
  if (desc != nullptr && desc->batch_rows != nullptr) return desc->batch_rows[st->fit_step];
  return rows_fixed;
}

// ---------------------------------------------------------------------------------------------
// fit(): batch gather (mlp_ensemble.py:175-176 `train_inputs[shuffles_per_mlp]`)
// ---------------------------------------------------------------------------------------------
__global__ void set_fit_desc_kernel(FitDesc d, FitDesc* out, TrainState* st) {
  *out = d;
  st->fit_step = 0;
}

__global__ void gather_batch_kernel(const FitDesc* desc, const TrainState* st, int rows_fixed,
                                    int in_dim, int out_dim, float* x, float* y,
                                    int64_t x_estride, int64_t y_estride) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const int e = blockIdx.y;
  const int width = in_dim + out_dim;
  const int* idx = desc->batch_index + ((int64_t)st->fit_step * gridDim.y + e) * desc->bmax;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * width; i += gridDim.x * blockDim.x) {
    const int r = i / width, c = i - r * width;
    const int64_t src = idx[r];
    if (c < in_dim)
      x[e * x_estride + (int64_t)r * in_dim + c] = desc->inputs[src * in_dim + c];
    else
      y[e * y_estride + (int64_t)r * out_dim + (c - in_dim)] = desc->targets[src * out_dim + (c - in_dim)];
  }
}

// ---------------------------------------------------------------------------------------------
// forward: Y = act(X . W + b) — BaseLayer.call / GaussianHead.call (mlp_ensemble.py:17-22, :32-34)
// ---------------------------------------------------------------------------------------------
struct LayerArgs {
  const float* in;        // [E][rows][K]   (e-stride may be 0: validation shares its rows)
  int64_t in_estride;
  const float* theta;     // [E][pn]
  int64_t pn;
  int off, K, N;
  float* out;             // forward: [E][rows][N]; backward: dZ_{l-1} [E][rows][K]
  int64_t out_estride;
  const float* dz;        // backward only: dZ_l [E][rows][N]
  int64_t dz_estride;
  int relu;
};

__global__ void __launch_bounds__(kThreads)
train_forward_kernel(LayerArgs a, const FitDesc* desc, const TrainState* st, int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const int e = blockIdx.z, r0 = blockIdx.y * kTileM, n0 = blockIdx.x * kTileN;
  if (r0 >= rows) return;
  __shared__ float Xs[kTileM][kTileK + 1];
  __shared__ __align__(16) float Ws[kTileK][kTileN];
  const float* X = a.in + e * a.in_estride;
  const float* W = a.theta + e * a.pn + a.off;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  float acc[2][4] = {};
  for (int k0 = 0; k0 < a.K; k0 += kTileK) {
    for (int i = tid; i < kTileM * kTileK; i += kThreads) {
      const int r = i >> 5, k = i & 31;
      Xs[r][k] = (r0 + r < rows && k0 + k < a.K) ? X[(int64_t)(r0 + r) * a.K + k0 + k] : 0.0f;
    }
    for (int i = tid; i < kTileK * kTileN; i += kThreads) {
      const int k = i >> 5, n = i & 31;
      Ws[k][n] = (k0 + k < a.K && n0 + n < a.N) ? W[(int64_t)(k0 + k) * a.N + n0 + n] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTileK; ++k) {
      const float x0 = Xs[ty * 2][k], x1 = Xs[ty * 2 + 1][k];
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      acc[0][0] = fmaf(x0, w.x, acc[0][0]); acc[0][1] = fmaf(x0, w.y, acc[0][1]);
      acc[0][2] = fmaf(x0, w.z, acc[0][2]); acc[0][3] = fmaf(x0, w.w, acc[0][3]);
      acc[1][0] = fmaf(x1, w.x, acc[1][0]); acc[1][1] = fmaf(x1, w.y, acc[1][1]);
      acc[1][2] = fmaf(x1, w.z, acc[1][2]); acc[1][3] = fmaf(x1, w.w, acc[1][3]);
    }
    __syncthreads();
  }
  const float* bias = W + (int64_t)a.K * a.N;
  float* Y = a.out + e * a.out_estride;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + ty * 2 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j] + bias[n];
      if (a.relu) v = fmaxf(v, 0.0f);
      Y[(int64_t)r * a.N + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// negative_log_likelihood (mlp_ensemble.py:64-67) and its gradient w.r.t. the head pre-activations
// ---------------------------------------------------------------------------------------------
struct NllArgs {
  const float* raw;        // [E][rows][2 O]: mu, then the var head's pre-activation
  int64_t raw_estride;
  const float* y;          // [E][rows][O]
  int64_t y_estride;
  float* d_raw;            // [E][rows][2 O] or null (validation)
  int64_t d_estride;
  float* partial;          // [E][tiles][2]
  int tiles_cap;
  int out_dim, ensemble;
  int train;               // 1: last CTA finalises the loss and the step's lr_t
  float* out_loss;         // device [1] or null
};

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  // fixed-order tree: shuffles inside the warp, then warp 0 adds the 8 warp sums in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float s = 0.0f;
  if (threadIdx.x == 0)
    for (int i = 0; i < kThreads / 32; ++i) s += scratch[i];
  return s;   // valid in thread 0
}

__device__ float schedule_lr(const OptParams& o, int iterations) {
  if (!o.schedule) return o.lr0;
  const float epochs = floorf((float)(iterations / o.steps_per_epoch));
  return fmaxf(o.lr0 * (1.0f - epochs / (float)o.train_epochs), 0.0f);
}

__global__ void __launch_bounds__(kThreads)
train_nll_kernel(NllArgs a, OptParams opt, const FitDesc* desc, TrainState* st, int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const int e = blockIdx.y, r0 = blockIdx.x * kTileM;
  const int O = a.out_dim;
  __shared__ float scratch[kThreads / 32];
  __shared__ bool last;
  float s_log = 0.0f, s_sq = 0.0f;
  if (r0 < rows) {
    const float c = 1.0f / ((float)rows * (float)O * (float)a.ensemble);
    const int nr = min(kTileM, rows - r0);
    const float* raw = a.raw + e * a.raw_estride + (int64_t)r0 * 2 * O;
    const float* y = a.y + e * a.y_estride + (int64_t)r0 * O;
    float* d = a.d_raw ? a.d_raw + e * a.d_estride + (int64_t)r0 * 2 * O : nullptr;
    for (int i = threadIdx.x; i < nr * O; i += kThreads) {
      const int r = i / O, o = i - r * O;
      const float mu = raw[(int64_t)r * 2 * O + o];
      const float pre = raw[(int64_t)r * 2 * O + O + o];
      const float var = softplus_tf(pre) + 1e-4f;
      const float diff = mu - y[(int64_t)r * O + o];
      const float inv = 1.0f / var;
      s_log += logf(6.28318530717958647692f * var);
      s_sq += diff * diff * inv;
      if (d) {
        d[(int64_t)r * 2 * O + o] = c * diff * inv;
        const float sig = 1.0f / (1.0f + expf(-pre));
        d[(int64_t)r * 2 * O + O + o] = 0.5f * c * (inv - diff * diff * inv * inv) * sig;
      }
    }
  }
  const float t_log = block_sum(s_log, scratch);
  const float t_sq = block_sum(s_sq, scratch);
  if (threadIdx.x == 0) {
    a.partial[((int64_t)e * a.tiles_cap + blockIdx.x) * 2 + 0] = t_log;
    a.partial[((int64_t)e * a.tiles_cap + blockIdx.x) * 2 + 1] = t_sq;
  }
  if (!a.train) return;
  // the last CTA to arrive adds the partials in (member, tile) order and prepares the update
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y;
    last = (atomicAdd(&st->nll_ticket, 1u) == total - 1);
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  const int tiles = (rows + kTileM - 1) / kTileM;
  const float denom = (float)rows * (float)O;
  float loss = 0.0f;
  for (int m = 0; m < a.ensemble; ++m) {
    float sl = 0.0f, sq = 0.0f;
    for (int t = 0; t < tiles; ++t) {
      sl += __ldcg(&a.partial[((int64_t)m * a.tiles_cap + t) * 2 + 0]);
      sq += __ldcg(&a.partial[((int64_t)m * a.tiles_cap + t) * 2 + 1]);
    }
    loss += (0.5f * (sl / denom) + 0.5f * (sq / denom)) / (float)a.ensemble;
  }
  const int it = st->iterations;
  const float t = (float)(it + 1);
  st->lr_t = schedule_lr(opt, it) * sqrtf(1.0f - powf(opt.beta2, t)) / (1.0f - powf(opt.beta1, t));
  st->loss = loss;
  if (a.out_loss) *a.out_loss = loss;
  if (desc != nullptr && desc->losses != nullptr) desc->losses[st->fit_step] = loss;
  st->nll_ticket = 0;
}

// validation_step (mlp_ensemble.py:148-156): chunk partials -> running sums -> loss
__global__ void val_accumulate_kernel(const float* partial, int tiles_cap, int tiles, int ensemble,
                                      double* acc, int reset) {
  const int e = threadIdx.x;
  if (e >= ensemble) return;
  double sl = reset ? 0.0 : acc[e * 2], sq = reset ? 0.0 : acc[e * 2 + 1];
  for (int t = 0; t < tiles; ++t) {
    sl += partial[((int64_t)e * tiles_cap + t) * 2 + 0];
    sq += partial[((int64_t)e * tiles_cap + t) * 2 + 1];
  }
  acc[e * 2] = sl;
  acc[e * 2 + 1] = sq;
}

__global__ void val_finalize_kernel(const double* acc, int ensemble, double denom, float* out_loss) {
  double loss = 0.0;
  for (int e = 0; e < ensemble; ++e)
    loss += (0.5 * acc[e * 2] / denom + 0.5 * acc[e * 2 + 1] / denom) / ensemble;
  *out_loss = (float)loss;
}

// ---------------------------------------------------------------------------------------------
// backward through one Dense + ReLU: dZ_{l-1} = (H_{l-1} > 0) * (dZ_l . W_l^T)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
train_backward_kernel(LayerArgs a, const FitDesc* desc, const TrainState* st, int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const int e = blockIdx.z, r0 = blockIdx.y * kTileM, k0 = blockIdx.x * kTileN;
  if (r0 >= rows) return;
  __shared__ float Zs[kTileM][kTileK + 1];
  __shared__ float Ws[kTileN][kTileK + 1];
  const float* dZ = a.dz + e * a.dz_estride;
  const float* W = a.theta + e * a.pn + a.off;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  float acc[2][4] = {};
  for (int n0 = 0; n0 < a.N; n0 += kTileK) {
    for (int i = tid; i < kTileM * kTileK; i += kThreads) {
      const int r = i >> 5, n = i & 31;
      Zs[r][n] = (r0 + r < rows && n0 + n < a.N) ? dZ[(int64_t)(r0 + r) * a.N + n0 + n] : 0.0f;
    }
    for (int i = tid; i < kTileN * kTileK; i += kThreads) {
      const int k = i >> 5, n = i & 31;
      Ws[k][n] = (k0 + k < a.K && n0 + n < a.N) ? W[(int64_t)(k0 + k) * a.N + n0 + n] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < kTileK; ++n) {
      const float z0 = Zs[ty * 2][n], z1 = Zs[ty * 2 + 1][n];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = Ws[tx * 4 + j][n];
        acc[0][j] = fmaf(z0, w, acc[0][j]);
        acc[1][j] = fmaf(z1, w, acc[1][j]);
      }
    }
    __syncthreads();
  }
  const float* H = a.in + e * a.in_estride;      // H_{l-1} after ReLU: > 0 <=> the unit was active
  float* out = a.out + e * a.out_estride;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + ty * 2 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= a.K) continue;
      out[(int64_t)r * a.K + k] = H[(int64_t)r * a.K + k] > 0.0f ? acc[i][j] : 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient + clip + Adam for every layer in one grid
// ---------------------------------------------------------------------------------------------
struct UpdTile { int layer, k0, n0; };
struct UpdLayer {
  const float* h;       // input of the layer [E][rows][K]
  int64_t h_estride;
  const float* dz;      // [E][rows][N]
  int64_t dz_estride;
  int off, K, N;
};
constexpr int kMaxTrainLayers = 18;
struct UpdArgs {
  UpdLayer layers[kMaxTrainLayers];
  const UpdTile* tiles;
  float* theta;
  float* m;
  float* v;
  float* grad;
  int64_t pn;
};

__global__ void __launch_bounds__(kThreads)
train_update_kernel(UpdArgs a, OptParams opt, const FitDesc* desc, TrainState* st, int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const UpdTile tile = a.tiles[blockIdx.x];
  const UpdLayer& L = a.layers[tile.layer];
  const int e = blockIdx.y;
  __shared__ float Hs[kTileK][kTileK + 1];
  __shared__ __align__(16) float Zs[kTileK][kTileN];
  const float* H = L.h + e * L.h_estride;
  const float* dZ = L.dz + e * L.dz_estride;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  float acc[4] = {};
  for (int r0 = 0; r0 < rows; r0 += kTileK) {
    for (int i = tid; i < kTileK * kTileK; i += kThreads) {
      const int r = i >> 5, k = i & 31;
      float h = 0.0f;
      if (r0 + r < rows) {
        if (tile.k0 + k < L.K) h = H[(int64_t)(r0 + r) * L.K + tile.k0 + k];
        else if (tile.k0 + k == L.K) h = 1.0f;          // the bias row
      }
      Hs[r][k] = h;
    }
    for (int i = tid; i < kTileK * kTileN; i += kThreads) {
      const int r = i >> 5, n = i & 31;
      Zs[r][n] = (r0 + r < rows && tile.n0 + n < L.N) ? dZ[(int64_t)(r0 + r) * L.N + tile.n0 + n] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kTileK; ++r) {
      const float h = Hs[r][ty];
      const float4 z = *reinterpret_cast<const float4*>(&Zs[r][tx * 4]);
      acc[0] = fmaf(h, z.x, acc[0]); acc[1] = fmaf(h, z.y, acc[1]);
      acc[2] = fmaf(h, z.z, acc[2]); acc[3] = fmaf(h, z.w, acc[3]);
    }
    __syncthreads();
  }
  const float lr_t = st->lr_t;
  const int k = tile.k0 + ty;
  if (k <= L.K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tile.n0 + tx * 4 + j;
      if (n >= L.N) continue;
      const int64_t p = e * a.pn + L.off + (int64_t)k * L.N + n;
      float g = acc[j];
      a.grad[p] = g;
      if (opt.clipvalue > 0.0f) g = fminf(fmaxf(g, -opt.clipvalue), opt.clipvalue);
      const float m = a.m[p] + (1.0f - opt.beta1) * (g - a.m[p]);
      const float v = a.v[p] + (1.0f - opt.beta2) * (g * g - a.v[p]);
      a.m[p] = m;
      a.v[p] = v;
      a.theta[p] -= lr_t * m / (sqrtf(v) + opt.epsilon);
    }
  }
  // the last CTA advances optimizer.iterations and fit's step counter
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(&st->upd_ticket, 1u) == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    st->iterations += 1;
    st->fit_step += 1;
    st->upd_ticket = 0;
  }
}

}   // namespace

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct simba_trainer {
  simba_model_t* model = nullptr;
  simba_trainer_config_t cfg{};
  int E = 0, L = 0, U = 0, O = 0, IN = 0;
  int cap_rows = 0;
  int64_t pn = 0;
  std::vector<int> off, K, N;        // train layers 0..L (L = merged head)
  float *theta = nullptr, *m = nullptr, *v = nullptr, *grad = nullptr;
  std::vector<float*> act;           // act[l]: input of train layer l (act[0] = gathered x)
  std::vector<float*> dz;            // dz[l]: gradient w.r.t. the pre-activation of train layer l
  float* raw = nullptr;
  float* ybuf = nullptr;
  float* partial = nullptr;
  int tiles_cap = 0;
  double* val_acc = nullptr;
  TrainState* state = nullptr;
  FitDesc* desc = nullptr;
  UpdTile* upd_tiles = nullptr;
  int n_upd_tiles = 0;
  cudaGraphExec_t fit_graph = nullptr;
  cudaStream_t graph_stream = nullptr;
  int launches_per_step = 0;
};

static OptParams opt_params(const simba_trainer_t* t) {
  OptParams o;
  o.lr0 = t->cfg.learning_rate;
  o.beta1 = t->cfg.beta1;
  o.beta2 = t->cfg.beta2;
  o.epsilon = t->cfg.epsilon;
  o.clipvalue = t->cfg.clipvalue;
  o.schedule = t->cfg.lr_schedule;
  o.steps_per_epoch = t->cfg.steps_per_epoch;
  o.train_epochs = t->cfg.train_epochs;
  return o;
}

// Keras arrays of one member -> the [(K+1) x N] blocks of theta (host)
static int pull_member(simba_trainer_t* t, int e, float* theta_e) {
  std::vector<float> kern, bias, kern2, bias2;
  for (int l = 0; l < t->L; ++l) {
    kern.resize((size_t)t->K[l] * t->N[l]);
    bias.resize(t->N[l]);
    int rc = simba_model_get_layer(t->model, e, l, kern.data(), bias.data());
    if (rc) return rc;
    float* dst = theta_e + t->off[l];
    memcpy(dst, kern.data(), kern.size() * sizeof(float));
    memcpy(dst + kern.size(), bias.data(), bias.size() * sizeof(float));
  }
  const int U = t->U, O = t->O;
  kern.resize((size_t)U * O); bias.resize(O); kern2.resize((size_t)U * O); bias2.resize(O);
  int rc = simba_model_get_layer(t->model, e, t->L, kern.data(), bias.data());
  if (rc) return rc;
  rc = simba_model_get_layer(t->model, e, t->L + 1, kern2.data(), bias2.data());
  if (rc) return rc;
  float* dst = theta_e + t->off[t->L];
  for (int k = 0; k < U; ++k)
    for (int o = 0; o < O; ++o) {
      dst[(size_t)k * 2 * O + o] = kern[(size_t)k * O + o];
      dst[(size_t)k * 2 * O + O + o] = kern2[(size_t)k * O + o];
    }
  for (int o = 0; o < O; ++o) {
    dst[(size_t)U * 2 * O + o] = bias[o];
    dst[(size_t)U * 2 * O + O + o] = bias2[o];
  }
  return SIMBA_OK;
}

// one block of theta (host) -> Keras kernel / bias of `layer` in [0, L + 2)
static void split_layer(const simba_trainer_t* t, const float* theta_e, int layer, float* kernel,
                        float* bias) {
  if (layer < t->L) {
    const float* src = theta_e + t->off[layer];
    const size_t nk = (size_t)t->K[layer] * t->N[layer];
    memcpy(kernel, src, nk * sizeof(float));
    memcpy(bias, src + nk, t->N[layer] * sizeof(float));
    return;
  }
  const int U = t->U, O = t->O, c0 = layer == t->L ? 0 : O;
  const float* src = theta_e + t->off[t->L];
  for (int k = 0; k < U; ++k)
    for (int o = 0; o < O; ++o) kernel[(size_t)k * O + o] = src[(size_t)k * 2 * O + c0 + o];
  for (int o = 0; o < O; ++o) bias[o] = src[(size_t)U * 2 * O + c0 + o];
}

extern "C" int simba_trainer_destroy(simba_trainer_t* t) {
  if (!t) return SIMBA_OK;
  if (t->fit_graph) cudaGraphExecDestroy(t->fit_graph);
  if (t->graph_stream) cudaStreamDestroy(t->graph_stream);
  cudaFree(t->theta); cudaFree(t->m); cudaFree(t->v); cudaFree(t->grad);
  for (float* p : t->act) cudaFree(p);
  for (float* p : t->dz) cudaFree(p);
  cudaFree(t->raw); cudaFree(t->ybuf); cudaFree(t->partial); cudaFree(t->val_acc);
  cudaFree(t->state); cudaFree(t->desc); cudaFree(t->upd_tiles);
  delete t;
  return SIMBA_OK;
}

extern "C" int simba_trainer_create(simba_model_t* model, const simba_trainer_config_t* cfg,
                                    simba_trainer_t** out) {
  if (!model || !cfg || !out) return set_error(SIMBA_ERR_BAD_CONFIG, "null argument");
  if (cfg->batch_size < 1 || cfg->batch_size > (1 << 20))
    return set_error(SIMBA_ERR_BAD_CONFIG, "batch_size %d out of range", cfg->batch_size);
  if (cfg->lr_schedule && (cfg->steps_per_epoch < 1 || cfg->train_epochs < 1))
    return set_error(SIMBA_ERR_BAD_CONFIG, "lr_schedule needs steps_per_epoch >= 1 and train_epochs >= 1");
  if (!(cfg->beta1 >= 0.0f && cfg->beta1 < 1.0f && cfg->beta2 >= 0.0f && cfg->beta2 < 1.0f) ||
      !(cfg->epsilon > 0.0f) || !(cfg->learning_rate >= 0.0f))
    return set_error(SIMBA_ERR_BAD_CONFIG, "Adam hyper-parameters out of range");
  int rc = simba_device_check();
  if (rc) return rc;
  const simba_model_config_t* mc = model_config(model);
  if (mc->n_layers + 1 > kMaxTrainLayers)
    return set_error(SIMBA_ERR_UNSUPPORTED, "n_layers %d > %d", mc->n_layers, kMaxTrainLayers - 1);
  auto* t = new simba_trainer();
  t->model = model;
  t->cfg = *cfg;
  t->E = mc->ensemble_size; t->L = mc->n_layers; t->U = mc->units; t->O = mc->obs_dim;
  t->IN = mc->obs_dim + mc->act_dim;
  t->cap_rows = cfg->max_eval_rows > cfg->batch_size ? cfg->max_eval_rows : cfg->batch_size;
  int64_t off = 0;
  for (int l = 0; l <= t->L; ++l) {
    const int K = l == 0 ? t->IN : t->U;
    const int N = l < t->L ? t->U : 2 * t->O;
    t->K.push_back(K); t->N.push_back(N); t->off.push_back((int)off);
    off += (int64_t)(K + 1) * N;
  }
  t->pn = off;
  const int E = t->E;
  std::vector<float> host((size_t)E * t->pn);
  for (int e = 0; e < E; ++e) {
    rc = pull_member(t, e, host.data() + (size_t)e * t->pn);
    if (rc) { delete t; return rc; }
  }
#define TRY_OR_FREE(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      simba_trainer_destroy(t);                                                                 \
      return set_error(SIMBA_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));         \
    }                                                                                           \
  } while (0)
  const size_t pbytes = (size_t)E * t->pn * sizeof(float);
  TRY_OR_FREE(cudaMalloc(&t->theta, pbytes));
  TRY_OR_FREE(cudaMalloc(&t->m, pbytes));
  TRY_OR_FREE(cudaMalloc(&t->v, pbytes));
  TRY_OR_FREE(cudaMalloc(&t->grad, pbytes));
  TRY_OR_FREE(cudaMemcpy(t->theta, host.data(), pbytes, cudaMemcpyHostToDevice));
  TRY_OR_FREE(cudaMemset(t->m, 0, pbytes));
  TRY_OR_FREE(cudaMemset(t->v, 0, pbytes));
  TRY_OR_FREE(cudaMemset(t->grad, 0, pbytes));
  const size_t R = t->cap_rows;
  t->act.assign(t->L + 1, nullptr);
  t->dz.assign(t->L + 1, nullptr);
  for (int l = 0; l <= t->L; ++l) {
    TRY_OR_FREE(cudaMalloc(&t->act[l], (size_t)E * R * t->K[l] * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&t->dz[l], (size_t)E * cfg->batch_size * t->N[l] * sizeof(float)));
  }
  TRY_OR_FREE(cudaMalloc(&t->raw, (size_t)E * R * 2 * t->O * sizeof(float)));
  TRY_OR_FREE(cudaMalloc(&t->ybuf, (size_t)E * cfg->batch_size * t->O * sizeof(float)));
  t->tiles_cap = (int)((R + kTileM - 1) / kTileM);
  TRY_OR_FREE(cudaMalloc(&t->partial, (size_t)E * t->tiles_cap * 2 * sizeof(float)));
  TRY_OR_FREE(cudaMalloc(&t->val_acc, (size_t)E * 2 * sizeof(double)));
  TRY_OR_FREE(cudaMalloc(&t->state, sizeof(TrainState)));
  TRY_OR_FREE(cudaMemset(t->state, 0, sizeof(TrainState)));
  TRY_OR_FREE(cudaMalloc(&t->desc, sizeof(FitDesc)));
  TRY_OR_FREE(cudaMemset(t->desc, 0, sizeof(FitDesc)));
  std::vector<UpdTile> tiles;
  for (int l = 0; l <= t->L; ++l)
    for (int k0 = 0; k0 <= t->K[l]; k0 += kTileK)
      for (int n0 = 0; n0 < t->N[l]; n0 += kTileN) tiles.push_back({l, k0, n0});
  t->n_upd_tiles = (int)tiles.size();
  TRY_OR_FREE(cudaMalloc(&t->upd_tiles, tiles.size() * sizeof(UpdTile)));
  TRY_OR_FREE(cudaMemcpy(t->upd_tiles, tiles.data(), tiles.size() * sizeof(UpdTile),
                         cudaMemcpyHostToDevice));
#undef TRY_OR_FREE
  t->launches_per_step = (t->L + 1) + 1 + t->L + 1;
  *out = t;
  return SIMBA_OK;
}

// forward through all train layers; x may be shared by the members (estride 0)
static int enqueue_forward(simba_trainer_t* t, const float* x, int64_t x_estride, int grid_rows,
                           int rows_fixed, const FitDesc* desc, cudaStream_t s) {
  const int64_t R = t->cap_rows;
  for (int l = 0; l <= t->L; ++l) {
    LayerArgs a{};
    a.in = l == 0 ? x : t->act[l];
    a.in_estride = l == 0 ? x_estride : R * t->K[l];
    a.theta = t->theta; a.pn = t->pn; a.off = t->off[l]; a.K = t->K[l]; a.N = t->N[l];
    a.out = l < t->L ? t->act[l + 1] : t->raw;
    a.out_estride = R * t->N[l];
    a.relu = l < t->L ? 1 : 0;
    dim3 grid((a.N + kTileN - 1) / kTileN, (grid_rows + kTileM - 1) / kTileM, t->E);
    train_forward_kernel<<<grid, kThreads, 0, s>>>(a, desc, t->state, rows_fixed);
  }
  SIMBA_CUDA_TRY(cudaGetLastError());
  return SIMBA_OK;
}

static int enqueue_step(simba_trainer_t* t, const float* x, int64_t x_estride, const float* y,
                        int64_t y_estride, int rows_fixed, const FitDesc* desc, float* out_loss,
                        cudaStream_t s) {
  const int64_t R = t->cap_rows, B = t->cfg.batch_size;
  const int grid_rows = desc ? t->cfg.batch_size : rows_fixed;
  const OptParams opt = opt_params(t);
  int rc = enqueue_forward(t, x, x_estride, grid_rows, rows_fixed, desc, s);
  if (rc) return rc;
  const int row_tiles = (grid_rows + kTileM - 1) / kTileM;
  NllArgs n{};
  n.raw = t->raw; n.raw_estride = R * 2 * t->O;
  n.y = y; n.y_estride = y_estride;
  n.d_raw = t->dz[t->L]; n.d_estride = B * 2 * t->O;
  n.partial = t->partial; n.tiles_cap = t->tiles_cap;
  n.out_dim = t->O; n.ensemble = t->E; n.train = 1; n.out_loss = out_loss;
  train_nll_kernel<<<dim3(row_tiles, t->E), kThreads, 0, s>>>(n, opt, desc, t->state, rows_fixed);
  for (int l = t->L; l >= 1; --l) {
    LayerArgs a{};
    a.in = t->act[l]; a.in_estride = R * t->K[l];       // H_{l-1}: the (ReLU) input of layer l
    a.theta = t->theta; a.pn = t->pn; a.off = t->off[l]; a.K = t->K[l]; a.N = t->N[l];
    a.dz = t->dz[l]; a.dz_estride = B * t->N[l];
    a.out = t->dz[l - 1]; a.out_estride = B * t->N[l - 1];   // N_{l-1} == K_l
    dim3 grid((a.K + kTileN - 1) / kTileN, row_tiles, t->E);
    train_backward_kernel<<<grid, kThreads, 0, s>>>(a, desc, t->state, rows_fixed);
  }
  UpdArgs u{};
  for (int l = 0; l <= t->L; ++l) {
    u.layers[l].h = l == 0 ? x : t->act[l];
    u.layers[l].h_estride = l == 0 ? x_estride : R * t->K[l];
    u.layers[l].dz = t->dz[l];
    u.layers[l].dz_estride = B * t->N[l];
    u.layers[l].off = t->off[l]; u.layers[l].K = t->K[l]; u.layers[l].N = t->N[l];
  }
  u.tiles = t->upd_tiles; u.theta = t->theta; u.m = t->m; u.v = t->v; u.grad = t->grad; u.pn = t->pn;
  train_update_kernel<<<dim3(t->n_upd_tiles, t->E), kThreads, 0, s>>>(u, opt, desc, t->state,
                                                                      rows_fixed);
  SIMBA_CUDA_TRY(cudaGetLastError());
  return SIMBA_OK;
}

extern "C" int simba_trainer_step(simba_trainer_t* t, const float* x, const float* y, int32_t rows,
                                  float* out_loss, void* stream) {
  if (!t || !x || !y) return set_error(SIMBA_ERR_BAD_CONFIG, "null argument");
  if (rows < 1 || rows > t->cfg.batch_size)
    return set_error(SIMBA_ERR_SHAPE, "rows %d outside [1, batch_size %d]", rows, t->cfg.batch_size);
  return enqueue_step(t, x, (int64_t)rows * t->IN, y, (int64_t)rows * t->O, rows, nullptr, out_loss,
                      (cudaStream_t)stream);
}

extern "C" int simba_trainer_fit(simba_trainer_t* t, const float* inputs, const float* targets,
                                 int64_t n, const int32_t* batch_index, const int32_t* batch_rows,
                                 int32_t steps, float* out_losses, void* stream) {
  if (!t || !inputs || !targets || !batch_index)
    return set_error(SIMBA_ERR_BAD_CONFIG, "null argument");
  if (n < 1 || steps < 0) return set_error(SIMBA_ERR_SHAPE, "n %lld / steps %d", (long long)n, steps);
  cudaStream_t s = (cudaStream_t)stream;
  const int B = t->cfg.batch_size;
  FitDesc d{inputs, targets, batch_index, batch_rows, out_losses, B};
  set_fit_desc_kernel<<<1, 1, 0, s>>>(d, t->desc, t->state);
  SIMBA_CUDA_TRY(cudaGetLastError());
  if (!t->fit_graph) {
    // every pointer the step reads is either owned by the handle or reached through *desc, so one
    // captured step serves every fit() call
    if (!t->graph_stream) SIMBA_CUDA_TRY(cudaStreamCreateWithFlags(&t->graph_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    SIMBA_CUDA_TRY(cudaStreamBeginCapture(t->graph_stream, cudaStreamCaptureModeThreadLocal));
    const int width = t->IN + t->O;
    gather_batch_kernel<<<dim3((B * width + 255) / 256, t->E), 256, 0, t->graph_stream>>>(
        t->desc, t->state, B, t->IN, t->O, t->act[0], t->ybuf, (int64_t)t->cap_rows * t->IN,
        (int64_t)B * t->O);
    int rc = enqueue_step(t, t->act[0], (int64_t)t->cap_rows * t->IN, t->ybuf, (int64_t)B * t->O, B,
                          t->desc, nullptr, t->graph_stream);
    cudaError_t ce = cudaStreamEndCapture(t->graph_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    SIMBA_CUDA_TRY(ce);
    ce = cudaGraphInstantiate(&t->fit_graph, graph, 0);
    cudaGraphDestroy(graph);
    SIMBA_CUDA_TRY(ce);
  }
  for (int i = 0; i < steps; ++i) SIMBA_CUDA_TRY(cudaGraphLaunch(t->fit_graph, s));
  return SIMBA_OK;
}

extern "C" int simba_trainer_validation(simba_trainer_t* t, const float* x, const float* y,
                                        int64_t rows, float* out_loss, void* stream) {
  if (!t || !x || !y || !out_loss) return set_error(SIMBA_ERR_BAD_CONFIG, "null argument");
  if (rows < 1) return set_error(SIMBA_ERR_SHAPE, "rows %lld", (long long)rows);
  cudaStream_t s = (cudaStream_t)stream;
  const OptParams opt = opt_params(t);
  for (int64_t r0 = 0; r0 < rows; r0 += t->cap_rows) {
    const int nr = (int)((rows - r0) < t->cap_rows ? (rows - r0) : t->cap_rows);
    int rc = enqueue_forward(t, x + r0 * t->IN, 0, nr, nr, nullptr, s);
    if (rc) return rc;
    const int tiles = (nr + kTileM - 1) / kTileM;
    NllArgs n{};
    n.raw = t->raw; n.raw_estride = (int64_t)t->cap_rows * 2 * t->O;
    n.y = y + r0 * t->O; n.y_estride = 0;
    n.d_raw = nullptr;
    n.partial = t->partial; n.tiles_cap = t->tiles_cap;
    n.out_dim = t->O; n.ensemble = t->E; n.train = 0; n.out_loss = nullptr;
    train_nll_kernel<<<dim3(tiles, t->E), kThreads, 0, s>>>(n, opt, nullptr, t->state, nr);
    val_accumulate_kernel<<<1, 32, 0, s>>>(t->partial, t->tiles_cap, tiles, t->E, t->val_acc,
                                           r0 == 0 ? 1 : 0);
  }
  val_finalize_kernel<<<1, 1, 0, s>>>(t->val_acc, t->E, (double)rows * t->O, out_loss);
  SIMBA_CUDA_TRY(cudaGetLastError());
  return SIMBA_OK;
}

extern "C" int simba_trainer_sync_model(simba_trainer_t* t, void* stream) {
  if (!t) return set_error(SIMBA_ERR_BAD_CONFIG, "null argument");
  std::vector<float> host((size_t)t->E * t->pn);
  SIMBA_CUDA_TRY(cudaMemcpyAsync(host.data(), t->theta, host.size() * sizeof(float),
                                 cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SIMBA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  std::vector<float> kern, bias;
  for (int e = 0; e < t->E; ++e)
    for (int layer = 0; layer < t->L + 2; ++layer) {
      const int K = layer == 0 ? t->IN : t->U, N = layer < t->L ? t->U : t->O;
      kern.resize((size_t)K * N);
      bias.resize(N);
      split_layer(t, host.data() + (size_t)e * t->pn, layer, kern.data(), bias.data());
      int rc = simba_model_set_layer(t->model, e, layer, kern.data(), bias.data());
      if (rc) return rc;
    }
  return simba_model_commit(t->model);
}

extern "C" int simba_trainer_get(simba_trainer_t* t, int32_t which, int32_t member, int32_t layer,
                                 float* kernel_out, float* bias_out) {
  if (!t || !kernel_out || !bias_out) return set_error(SIMBA_ERR_BAD_CONFIG, "null argument");
  if (which < 0 || which > 3 || member < 0 || member >= t->E || layer < 0 || layer >= t->L + 2)
    return set_error(SIMBA_ERR_BAD_CONFIG, "which %d / member %d / layer %d out of range", which,
                     member, layer);
  const float* src = which == 0 ? t->theta : which == 1 ? t->grad : which == 2 ? t->m : t->v;
  std::vector<float> host(t->pn);
  SIMBA_CUDA_TRY(cudaDeviceSynchronize());
  SIMBA_CUDA_TRY(cudaMemcpy(host.data(), src + (size_t)member * t->pn, t->pn * sizeof(float),
                            cudaMemcpyDeviceToHost));
  split_layer(t, host.data(), layer, kernel_out, bias_out);
  return SIMBA_OK;
}

extern "C" int64_t simba_trainer_iterations(simba_trainer_t* t) {
  if (!t) return -1;
  TrainState st;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(&st, t->state, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return st.iterations;
}

extern "C" int simba_trainer_launches_per_step(simba_trainer_t* t) {
  return t ? t->launches_per_step : 0;
}
