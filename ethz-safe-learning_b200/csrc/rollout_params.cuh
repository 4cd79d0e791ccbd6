// Launch parameters shared by the fp32 and bf16-tcgen05 rollout kernels.
#pragma once
#include "common.cuh"

namespace simba {

constexpr int kF32TileRows = 32;     // rows per CTA of the fp32 kernel
constexpr int kF32ChunkCols = 128;   // weight chunk = [16 k-rows][128 cols] fp32
constexpr int kF32ChunkRows = 16;
constexpr int kTcTileRows = 128;     // rows per tile of the tcgen05 kernel (UMMA_M)

struct RolloutParams {
  RowGeom g;
  simba_scorer_t scorer;
  const Tile* tiles;
  int32_t n_tiles;
  int32_t L, U;
  // fp32 image: per member, chunks [n_chunks][16][128] in consumption order; biases padded per layer
  const float* w_f32;
  const float* bias_f32;
  int32_t n_chunks, bias_stride, act_rows;
  // bf16 image: per member, per layer pre-swizzled UMMA B tiles (see pack_bf16 in api.cu)
  const void* w_bf16;
  int64_t w_bf16_member_bytes;
  // wide models (128 < units <= 440): per member the per-step stream of UMMA weight tiles in
  // consumption order, biases in the K padding (see rollout_tc_wide.cu)
  const void* w_wide;
  int64_t w_wide_member_bytes;
  const void* bias_k16;       // [E][L+1][4096 B] bias K-blocks (see simba_model_commit)
  const float* bias_tc;       // [E][L+1][128] fp32 (heads: mu bias at [0, O), var bias at [64, 64+O))
  int32_t tc_tiles_per_cta;   // 1 (latency: small populations) or 2 (MMA / epilogue ping-pong)
  int32_t n_sms;              // SMs of the planner's device (grid of the persistent tcgen05 kernels)
  int32_t tc_pair;            // one-tile variant only: a cluster of two CTAs per tile splits the head pass
  int32_t pdl;                // launch with programmatic stream serialization (fused plan path)
  float tc_scale_a[64];       // bf16 path: x_scaled[k] = fma(x[k], a[k], b[k]); zero beyond O + A
  float tc_scale_b[64];
  // scaler (transition_model.py:79-87): x_scaled = (x - smin) / sdelta, sdelta already holds 1.01
  // where max - min < 1e-5
  const float* smin;
  const float* sdelta;
  const float* sinv;          // 1 / sdelta (bf16 path multiplies)
  int32_t scale_on;
  // inputs
  const float* states;        // [S, O] (or per row when state_per_row)
  int64_t state_stride;
  int32_t state_per_row;
  const float* actions;       // [S, N, H, A] : offset = (s * N + i) * action_stride + t * A + a
  int64_t action_stride;
  const float* eps;           // [S, H, P*N, O] or null -> Philox
  uint64_t seed;
  const uint64_t* seed_ptr;   // if not null the seed is read from device memory (graph replay)
  int32_t iteration;
  int32_t sampling_propagation;
  int32_t objective;
  const int32_t* active;      // [S] or null
  // outputs
  float* row_return;
  uint64_t* row_costmask;
  float* row_costsum;
  float* traj_out;            // [rows, H+1, O] or null
  float* mu_out;              // [rows, O] or null (with var_out; sample_out optional)
  float* var_out;
  float* sample_out;
  long long* timeline;        // clock64 stamps of CTA 0 (-DSIMBA_TC_TIMELINE builds only, else null)
};

size_t rollout_f32_smem_bytes(const RolloutParams& prm);
cudaError_t launch_rollout_f32(const RolloutParams& prm, int n_tiles, cudaStream_t stream);
cudaError_t launch_rollout_tc(const RolloutParams& prm, int n_tiles, cudaStream_t stream);
bool rollout_tc_supported(int O, int A, int L, int U, int H);
bool rollout_tc_two_tiles_fit(int L, int n_constraints);
bool rollout_tc_pair_fits(int L, int n_constraints);
cudaError_t launch_rollout_tc_wide(const RolloutParams& prm, int n_tiles, cudaStream_t stream);
bool rollout_tc_wide_supported(int O, int A, int L, int U, int H);
bool rollout_tc_wide_fits(int L, int U, int n_constraints);
int rollout_tc_wide_units(int U);                  // U rounded up to a multiple of 16
int64_t rollout_tc_wide_member_bytes(int L, int U);   // U = padded units

}  // namespace simba
