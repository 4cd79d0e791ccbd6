// Ensemble training step on the device (include/simba_b200.h "ensemble training step"):
// replaces MlpEnsemble.training_step / validation_step / fit's inner loop
// (simba/models/mlp_ensemble.py:134-187). fp32 throughout — the reference trains in fp32 and the
// planner's fp32 and bf16 weight images are both derived from these master weights.
//
// Shape of the work: E members x batch 64 x a 4x128 MLP is 28 MFLOP per member-step, far below
// what one launch can hide, so the step is latency-bound. The design goal is therefore few, wide
// launches that stay inside one CUDA graph:
//   forward   L launches     H_l = relu([H_{l-1}, 1] . theta_l)    grid (N/32, rows/16, E)
//   head+nll  1 launch       mu, var, loss, d(mu), d(raw var), lr_t (mu / var stay in registers)
//   backward  L launches     dZ_{l-1} = relu'(H) * (dZ_l . W_l^T)  grid (K/32, rows/16, E)
//   update    1 launch       dtheta = [H, 1]^T . dZ for EVERY layer, clip, Adam, in one grid
// i.e. 2L + 2 launches per step; fit()'s batch gather is folded into the kernels that read the
// batch. Every CTA stages its whole (<= 128-deep) contraction with one wave of loads, so it pays
// the memory latency once, and the tiles are small so that one layer spreads over ~80 SMs.
// Parameters of train layer l are stored as one [(K_l + 1) x N_l] row-major block (Keras kernel
// [in, out] followed by the bias row), so "bias" is just the row that multiplies the constant 1
// and the weight-gradient GEMM produces the bias gradient as its last row. The Gaussian head's
// two Dense layers (mlp_ensemble.py:28-30) are one block with N = 2 * O whose columns interleave
// (mu_0, var_0, mu_1, var_1, ...). Reductions are in a fixed order: results are bit-reproducible.
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "internal.h"

using namespace simba;

namespace {

constexpr int kRowsF = 16;     // rows per CTA in forward / backward (more, smaller CTAs: the step is
                               // latency-bound, so the work is spread over as many SMs as possible)
constexpr int kTileN = 32;     // output columns per CTA
constexpr int kChunk = 128;    // contraction chunk staged in shared memory in ONE load phase
constexpr int kThreadsF = 128;
constexpr int kRowsU = 64;     // batch rows per chunk in the weight-gradient kernel
constexpr int kThreadsU = 256;

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

// fit(): `train_inputs[shuffles_per_mlp]` (mlp_ensemble.py:175-176) is never materialised — the
// kernels that read the batch follow the index
__device__ __forceinline__ const int* batch_rows_of(const FitDesc* desc, const TrainState* st, int e,
                                                    int ensemble) {
  return desc->batch_index + ((int64_t)st->fit_step * ensemble + e) * desc->bmax;
}

__global__ void set_fit_desc_kernel(FitDesc d, FitDesc* out, TrainState* st) {
  *out = d;
  st->fit_step = 0;
}

// ---------------------------------------------------------------------------------------------
// forward: Y = act(X . W + b) — BaseLayer.call / GaussianHead.call (mlp_ensemble.py:17-22, :32-34)
// ---------------------------------------------------------------------------------------------
struct LayerArgs {
  const float* in;        // [E][rows][K]   (e-stride may be 0: validation shares its rows)
  int64_t in_estride;
  int gather_in;          // 1: rows of `in` are desc->inputs[batch_index[...]] (layer 0 inside fit)
  const float* theta;     // [E][pn]
  int64_t pn;
  int off, K, N;
  float* out;             // forward: [E][rows][N]; backward: dZ_{l-1} [E][rows][K]
  int64_t out_estride;
  const float* dz;        // backward only: dZ_l [E][rows][N]
  int64_t dz_estride;
  int ensemble;
};

// acc[j] = sum_k X[r0 + ty][k] * W[k][n0 + tx * 4 + j]; all of a <= 128-deep contraction is staged
// by one wave of loads, so a CTA pays the memory latency once
__device__ __forceinline__ void forward_tile(const LayerArgs& a, const FitDesc* desc,
                                             const TrainState* st, int rows, int e, int r0, int n0,
                                             float (&Xs)[kRowsF][kChunk + 1],
                                             float (&Ws)[kChunk][kTileN], float (&acc)[4]) {
  const float* W = a.theta + e * a.pn + a.off;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const float* X = a.in + e * a.in_estride;
  const int* gidx = a.gather_in ? batch_rows_of(desc, st, e, a.ensemble) : nullptr;
  const bool vec = (a.N & 3) == 0;      // then every block offset and row of W is 16-byte aligned
  for (int k0 = 0; k0 < a.K; k0 += kChunk) {
    const int kc = min(kChunk, a.K - k0);
    if (k0) __syncthreads();
    // all global loads of the chunk are issued into registers before the first shared store, so
    // the CTA waits for memory once (the step is latency-bound, not bandwidth-bound)
    float xv[kRowsF * kChunk / kThreadsF];
#pragma unroll
    for (int j = 0; j < kRowsF * kChunk / kThreadsF; ++j) {
      const int i = tid + j * kThreadsF, r = i >> 7, k = i & (kChunk - 1);
      xv[j] = 0.0f;
      if (r0 + r < rows && k < kc) {
        const float* row = gidx ? desc->inputs + (int64_t)gidx[r0 + r] * a.K : X + (int64_t)(r0 + r) * a.K;
        xv[j] = __ldg(row + k0 + k);
      }
    }
    if (vec) {
      float4 wv[kChunk * kTileN / 4 / kThreadsF];
#pragma unroll
      for (int j = 0; j < kChunk * kTileN / 4 / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF, k = i >> 3, n = (i & 7) * 4;
        wv[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (k < kc && n0 + n < a.N)
          wv[j] = __ldcg(reinterpret_cast<const float4*>(W + (int64_t)(k0 + k) * a.N + n0 + n));
      }
#pragma unroll
      for (int j = 0; j < kChunk * kTileN / 4 / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF;
        *reinterpret_cast<float4*>(&Ws[i >> 3][(i & 7) * 4]) = wv[j];
      }
    } else {
      float wv[kChunk * kTileN / kThreadsF];
#pragma unroll
      for (int j = 0; j < kChunk * kTileN / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF, k = i >> 5, n = i & 31;
        wv[j] = (k < kc && n0 + n < a.N) ? __ldcg(W + (int64_t)(k0 + k) * a.N + n0 + n) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < kChunk * kTileN / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF;
        Ws[i >> 5][i & 31] = wv[j];
      }
    }
#pragma unroll
    for (int j = 0; j < kRowsF * kChunk / kThreadsF; ++j) {
      const int i = tid + j * kThreadsF;
      Xs[i >> 7][i & (kChunk - 1)] = xv[j];
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kc; ++k) {
      const float x = Xs[ty][k];
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      acc[0] = fmaf(x, w.x, acc[0]); acc[1] = fmaf(x, w.y, acc[1]);
      acc[2] = fmaf(x, w.z, acc[2]); acc[3] = fmaf(x, w.w, acc[3]);
    }
  }
}

__global__ void __launch_bounds__(kThreadsF)
train_forward_kernel(LayerArgs a, const FitDesc* desc, const TrainState* st, int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const int e = blockIdx.z, r0 = blockIdx.y * kRowsF, n0 = blockIdx.x * kTileN;
  if (r0 >= rows) return;
  __shared__ float Xs[kRowsF][kChunk + 1];
  __shared__ __align__(16) float Ws[kChunk][kTileN];
  float acc[4] = {};
  forward_tile(a, desc, st, rows, e, r0, n0, Xs, Ws, acc);
  const int tx = threadIdx.x & 7, r = r0 + (threadIdx.x >> 3);
  if (r >= rows) return;
  const float* bias = a.theta + e * a.pn + a.off + (int64_t)a.K * a.N;
  float* Y = a.out + e * a.out_estride + (int64_t)r * a.N;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    if (n < a.N) Y[n] = fmaxf(acc[j] + bias[n], 0.0f);
  }
}

// ---------------------------------------------------------------------------------------------
// Gaussian head + negative_log_likelihood (mlp_ensemble.py:28-34, :64-67) + its gradient w.r.t.
// the head pre-activations, fused: the head block stores (mu_o, var_o) in adjacent columns, so a
// thread's four accumulators are two complete outputs and mu / var never go to memory.
// ---------------------------------------------------------------------------------------------
struct NllArgs {
  const float* y;          // [E][rows][O] (ignored when gather_y)
  int64_t y_estride;
  int gather_y;            // 1: rows of y are desc->targets[batch_index[...]]
  float* d_raw;            // [E][rows][2 O] (interleaved like the head block) or null (validation)
  int64_t d_estride;
  float* partial;          // [E][tiles_cap][2]
  int tiles_cap;
  int out_dim;
  int train;               // 1: last CTA finalises the loss and the step's lr_t
  float* out_loss;         // device [1] or null
};

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  // fixed-order tree: shuffles inside the warp, then thread 0 adds the warp sums in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float s = 0.0f;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += scratch[i];
  return s;   // valid in thread 0
}

__device__ float schedule_lr(const OptParams& o, int iterations) {
  if (!o.schedule) return o.lr0;
  const float epochs = floorf((float)(iterations / o.steps_per_epoch));
  return fmaxf(o.lr0 * (1.0f - epochs / (float)o.train_epochs), 0.0f);
}

__global__ void __launch_bounds__(kThreadsF)
train_head_nll_kernel(LayerArgs a, NllArgs n, OptParams opt, const FitDesc* desc, TrainState* st,
                      int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const int e = blockIdx.z, r0 = blockIdx.y * kRowsF, n0 = blockIdx.x * kTileN;
  const int O = n.out_dim;
  __shared__ float Xs[kRowsF][kChunk + 1];
  __shared__ __align__(16) float Ws[kChunk][kTileN];
  __shared__ float scratch[kThreadsF / 32];
  __shared__ bool last;
  float s_log = 0.0f, s_sq = 0.0f;
  if (r0 < rows) {
    float acc[4] = {};
    forward_tile(a, desc, st, rows, e, r0, n0, Xs, Ws, acc);
    const int tx = threadIdx.x & 7, r = r0 + (threadIdx.x >> 3);
    if (r < rows) {
      const float c = 1.0f / ((float)rows * (float)O * (float)a.ensemble);
      const float* bias = a.theta + e * a.pn + a.off + (int64_t)a.K * a.N;
      const float* y = n.gather_y
          ? desc->targets + (int64_t)batch_rows_of(desc, st, e, a.ensemble)[r] * O
          : n.y + e * n.y_estride + (int64_t)r * O;
      float* d = n.d_raw ? n.d_raw + e * n.d_estride + (int64_t)r * 2 * O : nullptr;
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int col = n0 + tx * 4 + j;
        if (col >= a.N) continue;
        const int o = col >> 1;
        const float mu = acc[j] + bias[col];
        const float pre = acc[j + 1] + bias[col + 1];
        const float var = softplus_tf(pre) + 1e-4f;
        const float diff = mu - y[o];
        const float inv = 1.0f / var;
        s_log += logf(6.28318530717958647692f * var);
        s_sq += diff * diff * inv;
        if (d) {
          d[col] = c * diff * inv;
          const float sig = 1.0f / (1.0f + expf(-pre));
          d[col + 1] = 0.5f * c * (inv - diff * diff * inv * inv) * sig;
        }
      }
    }
  }
  const float t_log = block_sum(s_log, scratch);
  const float t_sq = block_sum(s_sq, scratch);
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) {
    n.partial[((int64_t)e * n.tiles_cap + tile) * 2 + 0] = t_log;
    n.partial[((int64_t)e * n.tiles_cap + tile) * 2 + 1] = t_sq;
  }
  if (!n.train) return;
  // the last CTA to arrive adds the partials in a fixed order and prepares the update
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    last = (atomicAdd(&st->nll_ticket, 1u) == total - 1);
  }
  __syncthreads();
  if (!last || threadIdx.x >= 32) return;
  __threadfence();
  const int lane = threadIdx.x;
  const int tiles = ((rows + kRowsF - 1) / kRowsF) * gridDim.x;
  const float denom = (float)rows * (float)O;
  float loss = 0.0f;
  for (int m = 0; m < a.ensemble; ++m) {
    float sl = 0.0f, sq = 0.0f;
    for (int t = lane; t < tiles; t += 32) {
      sl += __ldcg(&n.partial[((int64_t)m * n.tiles_cap + t) * 2 + 0]);
      sq += __ldcg(&n.partial[((int64_t)m * n.tiles_cap + t) * 2 + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sl += __shfl_down_sync(0xffffffffu, sl, o);
      sq += __shfl_down_sync(0xffffffffu, sq, o);
    }
    loss += (0.5f * (sl / denom) + 0.5f * (sq / denom)) / (float)a.ensemble;
  }
  if (lane != 0) return;
  const int it = st->iterations;
  const float t = (float)(it + 1);
  st->lr_t = schedule_lr(opt, it) * sqrtf(1.0f - powf(opt.beta2, t)) / (1.0f - powf(opt.beta1, t));
  st->loss = loss;
  if (n.out_loss) *n.out_loss = loss;
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
__global__ void __launch_bounds__(kThreadsF)
train_backward_kernel(LayerArgs a, const FitDesc* desc, const TrainState* st, int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const int e = blockIdx.z, r0 = blockIdx.y * kRowsF, k0 = blockIdx.x * kTileN;
  if (r0 >= rows) return;
  __shared__ float Zs[kRowsF][kChunk + 1];
  __shared__ float Ws[kTileN][kChunk + 1];
  const float* dZ = a.dz + e * a.dz_estride;
  const float* W = a.theta + e * a.pn + a.off;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  float acc[4] = {};
  for (int n0 = 0; n0 < a.N; n0 += kChunk) {
    const int nc = min(kChunk, a.N - n0);
    if (n0) __syncthreads();
    float zv[kRowsF * kChunk / kThreadsF];
#pragma unroll
    for (int j = 0; j < kRowsF * kChunk / kThreadsF; ++j) {
      const int i = tid + j * kThreadsF, r = i >> 7, n = i & (kChunk - 1);
      zv[j] = (r0 + r < rows && n < nc) ? __ldcg(dZ + (int64_t)(r0 + r) * a.N + n0 + n) : 0.0f;
    }
    if ((a.N & 3) == 0) {
      float4 wv[kTileN * kChunk / 4 / kThreadsF];
#pragma unroll
      for (int j = 0; j < kTileN * kChunk / 4 / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF, k = i >> 5, n = (i & 31) * 4;
        wv[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (k0 + k < a.K && n < nc)
          wv[j] = __ldcg(reinterpret_cast<const float4*>(W + (int64_t)(k0 + k) * a.N + n0 + n));
      }
#pragma unroll
      for (int j = 0; j < kTileN * kChunk / 4 / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF, k = i >> 5, n = (i & 31) * 4;
        Ws[k][n] = wv[j].x; Ws[k][n + 1] = wv[j].y; Ws[k][n + 2] = wv[j].z; Ws[k][n + 3] = wv[j].w;
      }
    } else {
      float wv[kTileN * kChunk / kThreadsF];
#pragma unroll
      for (int j = 0; j < kTileN * kChunk / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF, k = i >> 7, n = i & (kChunk - 1);
        wv[j] = (k0 + k < a.K && n < nc) ? __ldcg(W + (int64_t)(k0 + k) * a.N + n0 + n) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < kTileN * kChunk / kThreadsF; ++j) {
        const int i = tid + j * kThreadsF;
        Ws[i >> 7][i & (kChunk - 1)] = wv[j];
      }
    }
#pragma unroll
    for (int j = 0; j < kRowsF * kChunk / kThreadsF; ++j) {
      const int i = tid + j * kThreadsF;
      Zs[i >> 7][i & (kChunk - 1)] = zv[j];
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < nc; ++n) {
      const float z = Zs[ty][n];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(z, Ws[tx * 4 + j][n], acc[j]);
    }
  }
  const int r = r0 + ty;
  if (r >= rows) return;
  const float* H = a.in + e * a.in_estride + (int64_t)r * a.K;   // post-ReLU: > 0 <=> unit active
  float* out = a.out + e * a.out_estride + (int64_t)r * a.K;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + tx * 4 + j;
    if (k < a.K) out[k] = H[k] > 0.0f ? acc[j] : 0.0f;
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
  int gather_h;         // layer 0 inside fit: rows follow the batch index
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
  int ensemble;
};

__global__ void __launch_bounds__(kThreadsU)
train_update_kernel(UpdArgs a, OptParams opt, const FitDesc* desc, TrainState* st, int rows_fixed) {
  const int rows = resolve_rows(desc, st, rows_fixed);
  const UpdTile tile = a.tiles[blockIdx.x];
  const UpdLayer& L = a.layers[tile.layer];
  const int e = blockIdx.y;
  __shared__ float Hs[kRowsU][kTileN + 1];
  __shared__ __align__(16) float Zs[kRowsU][kTileN];
  const float* H = L.h + e * L.h_estride;
  const float* dZ = L.dz + e * L.dz_estride;
  const int* gidx = L.gather_h ? batch_rows_of(desc, st, e, a.ensemble) : nullptr;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  // the Adam operands are fetched while the gradient tile is being computed
  const int k = tile.k0 + ty;
  float th[4], mo[4], ve[4];
  const int64_t p0 = e * a.pn + L.off + (int64_t)k * L.N + tile.n0 + tx * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool ok = k <= L.K && tile.n0 + tx * 4 + j < L.N;
    th[j] = ok ? a.theta[p0 + j] : 0.0f;
    mo[j] = ok ? a.m[p0 + j] : 0.0f;
    ve[j] = ok ? a.v[p0 + j] : 0.0f;
  }
  float acc[4] = {};
  for (int r0 = 0; r0 < rows; r0 += kRowsU) {
    if (r0) __syncthreads();
    float hv[kRowsU * kTileN / kThreadsU], zv[kRowsU * kTileN / kThreadsU];
#pragma unroll
    for (int j = 0; j < kRowsU * kTileN / kThreadsU; ++j) {
      const int i = tid + j * kThreadsU, r = i >> 5, kk = i & 31;
      float h = 0.0f, z = 0.0f;
      if (r0 + r < rows) {
        if (tile.k0 + kk < L.K) {
          const float* row = gidx ? desc->inputs + (int64_t)gidx[r0 + r] * L.K : H + (int64_t)(r0 + r) * L.K;
          h = __ldcg(row + tile.k0 + kk);
        } else if (tile.k0 + kk == L.K) {
          h = 1.0f;                                   // the bias row
        }
        if (tile.n0 + kk < L.N) z = __ldcg(dZ + (int64_t)(r0 + r) * L.N + tile.n0 + kk);
      }
      hv[j] = h;
      zv[j] = z;
    }
#pragma unroll
    for (int j = 0; j < kRowsU * kTileN / kThreadsU; ++j) {
      const int i = tid + j * kThreadsU;
      Hs[i >> 5][i & 31] = hv[j];
      Zs[i >> 5][i & 31] = zv[j];
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kRowsU; ++r) {
      const float h = Hs[r][ty];
      const float4 z = *reinterpret_cast<const float4*>(&Zs[r][tx * 4]);
      acc[0] = fmaf(h, z.x, acc[0]); acc[1] = fmaf(h, z.y, acc[1]);
      acc[2] = fmaf(h, z.z, acc[2]); acc[3] = fmaf(h, z.w, acc[3]);
    }
  }
  const float lr_t = st->lr_t;
  if (k <= L.K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (tile.n0 + tx * 4 + j >= L.N) continue;
      float g = acc[j];
      a.grad[p0 + j] = g;
      if (opt.clipvalue > 0.0f) g = fminf(fmaxf(g, -opt.clipvalue), opt.clipvalue);
      const float m = mo[j] + (1.0f - opt.beta1) * (g - mo[j]);
      const float v = ve[j] + (1.0f - opt.beta2) * (g * g - ve[j]);
      a.m[p0 + j] = m;
      a.v[p0 + j] = v;
      a.theta[p0 + j] = th[j] - lr_t * m / (sqrtf(v) + opt.epsilon);
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
      dst[(size_t)k * 2 * O + 2 * o] = kern[(size_t)k * O + o];
      dst[(size_t)k * 2 * O + 2 * o + 1] = kern2[(size_t)k * O + o];
    }
  for (int o = 0; o < O; ++o) {
    dst[(size_t)U * 2 * O + 2 * o] = bias[o];
    dst[(size_t)U * 2 * O + 2 * o + 1] = bias2[o];
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
  const int U = t->U, O = t->O, c0 = layer == t->L ? 0 : 1;
  const float* src = theta_e + t->off[t->L];
  for (int k = 0; k < U; ++k)
    for (int o = 0; o < O; ++o) kernel[(size_t)k * O + o] = src[(size_t)k * 2 * O + 2 * o + c0];
  for (int o = 0; o < O; ++o) bias[o] = src[(size_t)U * 2 * O + 2 * o + c0];
}

extern "C" int simba_trainer_destroy(simba_trainer_t* t) {
  if (!t) return SIMBA_OK;
  if (t->fit_graph) cudaGraphExecDestroy(t->fit_graph);
  if (t->graph_stream) cudaStreamDestroy(t->graph_stream);
  cudaFree(t->theta); cudaFree(t->m); cudaFree(t->v); cudaFree(t->grad);
  for (float* p : t->act) cudaFree(p);
  for (float* p : t->dz) cudaFree(p);
  cudaFree(t->partial); cudaFree(t->val_acc);
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
  t->tiles_cap = (int)((R + kRowsF - 1) / kRowsF) * ((2 * t->O + kTileN - 1) / kTileN);
  TRY_OR_FREE(cudaMalloc(&t->partial, (size_t)E * t->tiles_cap * 2 * sizeof(float)));
  TRY_OR_FREE(cudaMalloc(&t->val_acc, (size_t)E * 2 * sizeof(double)));
  TRY_OR_FREE(cudaMalloc(&t->state, sizeof(TrainState)));
  TRY_OR_FREE(cudaMemset(t->state, 0, sizeof(TrainState)));
  TRY_OR_FREE(cudaMalloc(&t->desc, sizeof(FitDesc)));
  TRY_OR_FREE(cudaMemset(t->desc, 0, sizeof(FitDesc)));
  std::vector<UpdTile> tiles;
  for (int l = 0; l <= t->L; ++l)
    for (int k0 = 0; k0 <= t->K[l]; k0 += kTileN)
      for (int n0 = 0; n0 < t->N[l]; n0 += kTileN) tiles.push_back({l, k0, n0});
  t->n_upd_tiles = (int)tiles.size();
  TRY_OR_FREE(cudaMalloc(&t->upd_tiles, tiles.size() * sizeof(UpdTile)));
  TRY_OR_FREE(cudaMemcpy(t->upd_tiles, tiles.data(), tiles.size() * sizeof(UpdTile),
                         cudaMemcpyHostToDevice));
#undef TRY_OR_FREE
  t->launches_per_step = 2 * t->L + 2;
  *out = t;
  return SIMBA_OK;
}

// one launch per hidden layer, then the head fused with the likelihood. x may be shared by the
// members (estride 0, validation) or gathered through the fit descriptor (gather = 1).
static int enqueue_forward(simba_trainer_t* t, const float* x, int64_t x_estride, const float* y,
                           int64_t y_estride, int gather, int grid_rows, int rows_fixed,
                           const FitDesc* desc, int train, float* out_loss, cudaStream_t s) {
  const int64_t R = t->cap_rows;
  const int row_tiles = (grid_rows + kRowsF - 1) / kRowsF;
  for (int l = 0; l <= t->L; ++l) {
    LayerArgs a{};
    a.in = l == 0 ? x : t->act[l];
    a.in_estride = l == 0 ? x_estride : R * t->K[l];
    a.gather_in = l == 0 ? gather : 0;
    a.theta = t->theta; a.pn = t->pn; a.off = t->off[l]; a.K = t->K[l]; a.N = t->N[l];
    a.ensemble = t->E;
    dim3 grid((a.N + kTileN - 1) / kTileN, row_tiles, t->E);
    if (l < t->L) {
      a.out = t->act[l + 1];
      a.out_estride = R * t->N[l];
      train_forward_kernel<<<grid, kThreadsF, 0, s>>>(a, desc, t->state, rows_fixed);
    } else {
      NllArgs n{};
      n.y = y; n.y_estride = y_estride; n.gather_y = gather;
      n.d_raw = train ? t->dz[t->L] : nullptr;
      n.d_estride = (int64_t)t->cfg.batch_size * 2 * t->O;
      n.partial = t->partial; n.tiles_cap = t->tiles_cap;
      n.out_dim = t->O; n.train = train; n.out_loss = out_loss;
      train_head_nll_kernel<<<grid, kThreadsF, 0, s>>>(a, n, opt_params(t), desc, t->state, rows_fixed);
    }
  }
  SIMBA_CUDA_TRY(cudaGetLastError());
  return SIMBA_OK;
}

static int enqueue_step(simba_trainer_t* t, const float* x, int64_t x_estride, const float* y,
                        int64_t y_estride, int rows_fixed, const FitDesc* desc, float* out_loss,
                        cudaStream_t s) {
  const int64_t R = t->cap_rows, B = t->cfg.batch_size;
  const int gather = desc ? 1 : 0;
  const int grid_rows = desc ? t->cfg.batch_size : rows_fixed;
  int rc = enqueue_forward(t, x, x_estride, y, y_estride, gather, grid_rows, rows_fixed, desc, 1,
                           out_loss, s);
  if (rc) return rc;
  const int row_tiles = (grid_rows + kRowsF - 1) / kRowsF;
  for (int l = t->L; l >= 1; --l) {
    LayerArgs a{};
    a.in = t->act[l]; a.in_estride = R * t->K[l];       // H_{l-1}: the (ReLU) input of layer l
    a.theta = t->theta; a.pn = t->pn; a.off = t->off[l]; a.K = t->K[l]; a.N = t->N[l];
    a.dz = t->dz[l]; a.dz_estride = B * t->N[l];
    a.out = t->dz[l - 1]; a.out_estride = B * t->N[l - 1];   // N_{l-1} == K_l
    a.ensemble = t->E;
    dim3 grid((a.K + kTileN - 1) / kTileN, row_tiles, t->E);
    train_backward_kernel<<<grid, kThreadsF, 0, s>>>(a, desc, t->state, rows_fixed);
  }
  UpdArgs u{};
  for (int l = 0; l <= t->L; ++l) {
    u.layers[l].h = l == 0 ? x : t->act[l];
    u.layers[l].h_estride = l == 0 ? x_estride : R * t->K[l];
    u.layers[l].gather_h = l == 0 ? gather : 0;
    u.layers[l].dz = t->dz[l];
    u.layers[l].dz_estride = B * t->N[l];
    u.layers[l].off = t->off[l]; u.layers[l].K = t->K[l]; u.layers[l].N = t->N[l];
  }
  u.tiles = t->upd_tiles; u.theta = t->theta; u.m = t->m; u.v = t->v; u.grad = t->grad; u.pn = t->pn;
  u.ensemble = t->E;
  train_update_kernel<<<dim3(t->n_upd_tiles, t->E), kThreadsU, 0, s>>>(u, opt_params(t), desc,
                                                                        t->state, rows_fixed);
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
    int rc = enqueue_step(t, nullptr, 0, nullptr, 0, B, t->desc, nullptr, t->graph_stream);
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
  for (int64_t r0 = 0; r0 < rows; r0 += t->cap_rows) {
    const int nr = (int)((rows - r0) < t->cap_rows ? (rows - r0) : t->cap_rows);
    int rc = enqueue_forward(t, x + r0 * t->IN, 0, y + r0 * t->O, 0, 0, nr, nr, nullptr, 0, nullptr, s);
    if (rc) return rc;
    const int tiles = ((nr + kRowsF - 1) / kRowsF) * ((2 * t->O + kTileN - 1) / kTileN);
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
