// bf16 tcgen05 fused rollout + scoring kernel (sm_100a) — the throughput path of the planner.
//
// Same contract as rollout_f32.cu (one launch per CEM iteration replaces cem_mpc.py:49-55,
// transition_model.py:64-87, mlp_ensemble.py:122-132,189-193, safety_gym.py:110-166 and the per-row
// part of mpc_policy.py:30-37 / safe_cem_mpc.py:82-93), with the per-layer GEMMs on the 5th-gen
// tensor cores:
//
//   * one CTA = one ensemble member x NTILES row tiles of 128 rollouts; the member's whole bf16
//     weight set (144 KB for 4x128) is staged ONCE in shared memory by the TMA engine
//     (cp.async.bulk of pre-swizzled UMMA images) and reused for all H steps x (L+1) layers;
//   * activations are the A operand and live in TENSOR MEMORY (TS-mode tcgen05.mma: A from TMEM,
//     B = weights from shared memory), so an MMA streams only the weight tile from SMEM. Q epilogue
//     threads share one rollout row (Q = 4: latency configuration for small populations, Q = 2
//     with two tiles per CTA: throughput configuration): each reads its column slice of the fp32
//     accumulator row from TMEM (tcgen05.ld 32x32b), applies ReLU, packs bf16 and writes the next
//     layer's A operand back into the tile's A columns (tcgen05.st); the thread count per SM
//     sub-partition (4 warps) is what hides the MUFU / TMEM latencies. The bias of every layer is
//     one more K = 16 MMA (constant ones tile x [bf16(b), bf16(b - hi)] block, both from SMEM);
//   * per tile and layer: the tile's threads wait for their A-operand stores (tcgen05.wait::st),
//     fence and meet at a named barrier (bar.sync, tile threads only); one elected thread of the
//     tile then issues tcgen05.mma (M=128, N=128, K=16 per instruction, fp32 accumulate in TMEM)
//     and tcgen05.commit onto the tile's "accumulator ready" mbarrier, on which the tile's threads
//     wait. There is no separate MMA warp to wake up; with two tiles per CTA the tensor pipe works
//     on one tile while the other tile's threads run their epilogue;
//   * the rollout state s_t (fp32) lives in spare TMEM columns next to the accumulators, so it costs
//     no registers between steps. The Gaussian-head epilogue is ONE pass over 16-wide column chunks:
//     load mu / raw-var accumulators and the state chunk from TMEM, softplus, sqrt, Philox4x32-10 +
//     Box-Muller, residual update, store the state back (tcgen05.st), and in the same registers
//     pack the scaled bf16 input of the NEXT step into the A tile and take the partial lidar minima.
//     The per-row min over slices owned by different threads goes through a small shared-memory
//     exchange and one named barrier per step and tile.
#include <cuda_bf16.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "rollout_params.cuh"
#include "tc_ptx.cuh"

namespace simba {

// Debug timeline (compile with -DSIMBA_TC_TIMELINE): two threads of CTA 0 stamp clock64() at phase
// boundaries into prm.traj_out (a scratch buffer handed in by tools/tc_timeline.py).
#ifdef SIMBA_TC_TIMELINE
#define TL(ev)                                                                                   \
  do {                                                                                           \
    if (tl_who >= 0 && blockIdx.x == 0)                                                          \
      reinterpret_cast<long long*>(prm.traj_out)[(tl_who * 64 + tl_t) * 64 + (ev)] = clock64();  \
  } while (0)
#else
#define TL(ev) do { } while (0)
#endif

namespace {

constexpr int kU = 128;                 // hidden width this kernel covers
constexpr int kMaxO = 60;               // observation dims (two Gaussian heads padded to 2 x 64 columns)
constexpr int kAtomBytes = 128 * 128;   // one 64-wide K atom of a 128-row tile (bf16, SW128)
constexpr int kParts = 1 + SIMBA_MAX_CONSTRAINTS;   // goal + constrained lidar partial minima

struct TileInfo {
  int32_t member, k0, count, valid;
};

}  // namespace

// NTILES row tiles per CTA, Q epilogue threads per rollout row.
template <int NTILES, int Q>
__global__ void __launch_bounds__(NTILES* Q * 128, 1) rollout_tc_kernel(const RolloutParams prm) {
  constexpr int kTileThreads = Q * 128;
  constexpr int kEpiThreads = NTILES * kTileThreads;
  constexpr int OW = 64 / Q;          // head outputs (= state dims = layer-0 K elements) per thread
  constexpr int CW = OW >= 16 ? 16 : OW;   // state / head columns handled per chunk (16 or 8)
  constexpr int NSUB = OW / CW;
  constexpr int HW = (128 / Q) >= 32 ? 32 : (128 / Q);   // accumulator columns per hidden-layer chunk
  constexpr int HC = (128 / Q) / HW;                      // such chunks per thread
  // TMEM columns per tile: 128 fp32 accumulator columns, 64 fp32 state columns and 64 columns that
  // hold the bf16 A operand (128 K-elements, two per 32-bit column). Layout:
  // [NTILES x 128 acc][NTILES x 64 state][NTILES x 64 A]  ->  256 / 512 columns (power of two)
  constexpr int kTmemCols = NTILES * 256;
  // Throughput variant (two tiles, 512 threads at the 128-register cap): the step's noise, the
  // per-row running objective and the Philox key are parked in shared memory to free registers.
  // Latency variant (one tile): they stay in registers — no spills there, and the extra shared
  // memory round trips would sit on the critical path.
  constexpr bool kPark = NTILES > 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const RowGeom& g = prm.g;
  const int L = prm.L;
  const int H = g.H;
  const int O = g.O, A = g.A;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- shared memory carve-up (base is 1024-aligned: required by SWIZZLE_128B) -------------------
  const uint32_t w_bytes = kAtomBytes + (uint32_t)L * 2 * kAtomBytes;     // layer 0: 1 atom; others: 2
  uint8_t* w_smem = smem_raw;
  uint8_t* biask_smem = w_smem + w_bytes;                                  // [(L+1)][4096]: bias K-blocks (B operand)
  uint8_t* ones_smem = biask_smem + (L + 1) * 4096;                        // [4096]: A operand of the bias K-step
  float* scale_smem = reinterpret_cast<float*>(ones_smem + 4096);          // [2][64]: a, b of x*a+b
  float* pen_smem = scale_smem + 128;                                      // [kParts][64]: 0 in slice, +inf outside
  const int nparts = 1 + prm.scorer.n_constraints;                         // goal + constrained lidars
  float* part_smem = pen_smem + kParts * 64;                               // [NTILES][Q][nparts][128]
  // N(0,1) draws of the current step, packed bf16x2, [OW / 2][threads] (thread-minor: conflict-free).
  // They are produced during the hidden layers and consumed by the head pass; parking them here
  // instead of in OW registers leaves the head pass room to keep more TMEM loads in flight.
  uint32_t* noise_smem = reinterpret_cast<uint32_t*>(part_smem + NTILES * Q * nparts * 128);
  // running objective of every rollout row (RowScore fields, field-major): only one of the Q threads of a
  // row scores, and only once per step, so the seven words live here instead of in registers
  uint32_t* rs_smem = noise_smem + (NTILES > 1 ? (OW / 2) * kEpiThreads : 0);   // [8][NTILES * 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rs_smem + (NTILES > 1 ? 8 * NTILES * 128 : 0));
  // (both parking areas exist only in the two-tile variant, see kPark below)
  // bars[0] = weights landed; bars[1 + j] = accumulator ready (tile j)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * NTILES);
  TileInfo* tinfo = reinterpret_cast<TileInfo*>(tmem_slot + 2);

  const uint32_t bar_w = smem_u32(&bars[0]);
  uint32_t bar_acc[NTILES];
#pragma unroll
  for (int j = 0; j < NTILES; ++j) bar_acc[j] = smem_u32(&bars[1 + j]);

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
#pragma unroll
    for (int j = 0; j < NTILES; ++j) mbar_init(bar_acc[j], 1);      // tcgen05.commit
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  // everything above is independent of the previous kernel (the CEM update that wrote the actions
  // and the active flags); from here on its results are needed
  pdl_wait_prior_grid();
  __shared__ uint64_t seed_sh;          // Philox key (kPark: read where it is used, not held in two registers)
  if (threadIdx.x == 0) seed_sh = prm.seed_ptr ? *prm.seed_ptr : prm.seed;
  if (threadIdx.x < NTILES) {
    const int ti = blockIdx.x * NTILES + threadIdx.x;
    TileInfo info{0, 0, 0, 0};
    if (ti < prm.n_tiles) {
      const Tile t = prm.tiles[ti];
      info.member = t.member; info.k0 = t.k0; info.count = t.count;
      bool any = t.count > 0;
      if (any && prm.active != nullptr) {                     // cem_mpc.py:66-67 early exit
        const int m = g.rows_per_state[t.member];
        any = false;
        for (int s = t.k0 / m; s <= (t.k0 + t.count - 1) / m; ++s) any = any || prm.active[s] != 0;
      }
      info.valid = any ? 1 : 0;
    }
    tinfo[threadIdx.x] = info;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  bool any_valid = false;
  int member = 0;
#pragma unroll
  for (int j = 0; j < NTILES; ++j)
    if (tinfo[j].valid) { any_valid = true; member = tinfo[j].member; }

  // (A dedicated MMA warp fed by per-warp mbarrier arrivals, and setmaxnreg register re-balancing,
  //  were the first designs: the polling warp cost issue slots and one extra wake-up per layer, and
  //  setmaxnreg faulted at run time on sm_100a / CUDA 12.9. The tile-local barrier + elected issuer
  //  below needs neither.)
  {
    if (any_valid) {
      // weights of this member: one TMA bulk copy per layer, all onto bars[0]
      if (threadIdx.x == 0) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(prm.w_bf16) +
                              (size_t)member * prm.w_bf16_member_bytes;
        const uint32_t bk_bytes = (uint32_t)(L + 1) * 4096u;
        mbar_expect_tx(bar_w, w_bytes + bk_bytes);
        uint32_t off = 0;
        for (int l = 0; l <= L; ++l) {
          const uint32_t nb = (l == 0) ? kAtomBytes : 2 * kAtomBytes;
          bulk_g2s(smem_u32(w_smem + off), wsrc + off, nb, bar_w);
          off += nb;
        }
        bulk_g2s(smem_u32(biask_smem), reinterpret_cast<const uint8_t*>(prm.bias_k16) + (size_t)member * bk_bytes,
                 bk_bytes, bar_w);
      }
      // ============ epilogue warps: Q threads per rollout row, each owns a column slice ============
      const int j = warp / (4 * Q);                       // tile of this warp
      const int wl = warp - j * 4 * Q;                    // warp within the tile
      const int cgp = wl >> 2;                            // column group in [0, Q)
      const int r = (wl & 3) * 32 + lane;                 // row in tile == TMEM lane
      const TileInfo ti = tinfo[j];
      // biases + scaler of this member -> shared (all epilogue threads of the CTA cooperate)
      {
        // A operand of the bias K-step: ones[m][0] = ones[m][1] = 1, rest 0 (same core-matrix layout)
        for (int i = threadIdx.x; i < 1024; i += kEpiThreads) {
          const int byte = i * 4;                           // [16 groups][2 K halves][8 rows][16 B]
          const bool first = ((byte & 255) < 128) && ((byte & 15) == 0);   // K half 0, elements 0 and 1
          reinterpret_cast<uint32_t*>(ones_smem)[i] = first ? 0x3F803F80u : 0u;
        }
        fence_proxy_async();                                // generic-proxy writes -> visible to the MMA
        for (int i = threadIdx.x; i < 64; i += kEpiThreads) {
          scale_smem[i] = prm.tc_scale_a[i];
          scale_smem[64 + i] = prm.tc_scale_b[i];
          const simba_scorer_t& scc = prm.scorer;
          pen_smem[i] = (i >= scc.goal_begin && i < scc.goal_end) ? 0.0f : INFINITY;
          for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q)
            pen_smem[(1 + q) * 64 + i] =
                (q < scc.n_constraints && i >= scc.con_begin[q] && i < scc.con_end[q]) ? 0.0f : INFINITY;
        }
        named_bar_sync<kEpiThreads>(1);
      }
      if (ti.valid) {
        const bool row_ok = r < ti.count;
        const RowId id = decode_row(g, ti.member, ti.k0 + (row_ok ? r : 0));
        const uint32_t t_lane = tmem_base + ((uint32_t)((wl & 3) * 32) << 16) + (uint32_t)j * 128;
        const float* act_ptr = prm.actions + ((int64_t)id.s * g.N + id.i_global) * prm.action_stride;
        const bool done_first = objective_done_first(prm.objective);
        const simba_scorer_t& sc = prm.scorer;
        const int o_base = cgp * OW;                      // first state dim / head output of this thread
        float* part = part_smem + ((j * Q + cgp) * nparts) * 128 + r;         // [nparts] stride 128
        const float* part_row = part_smem + (j * Q * nparts) * 128 + r;       // group 0 base of this row
        const float D = sc.lidar_max_dist;

        const uint32_t t_state = tmem_base + ((uint32_t)((wl & 3) * 32) << 16) +
                                 (uint32_t)(NTILES * 128 + j * 64);            // this row's state columns
        const uint32_t t_a = tmem_base + ((uint32_t)((wl & 3) * 32) << 16) +
                             (uint32_t)(NTILES * 192 + j * 64);                // this row's A-operand columns
        const float* s0_ptr = prm.states + (prm.state_per_row ? id.r_global : (int64_t)id.s) * prm.state_stride;

        // The tile's A operand for `layer` is complete once every thread of the tile has passed the
        // named barrier below (each fenced its own writes to the async proxy first); the tile's
        // elected thread then issues that layer's MMAs and commits them onto the tile's mbarrier.
#ifdef SIMBA_TC_TIMELINE
        const int tl_who = (j == 0 && lane == 0) ? (wl == 0 ? 0 : (wl == 4 * Q - 1 ? 1 : -1)) : -1;
        int tl_t = 0;
#endif
        const bool issuer = (wl == 0) && (lane == 0);
        auto tile_sync_and_issue = [&](int layer) {
          tmem_st_wait();                                     // this thread's A-operand stores have landed
          TL(48 + layer);
          tc_fence_before();
          named_bar_sync<kTileThreads>(2 + j);
          TL(54 + layer);
          if (issuer) {
            tc_fence_after();
            const int ksteps = (layer == 0) ? 4 : 8;          // K = 64 or 128, UMMA_K = 16
            const uint32_t woff = (layer == 0) ? 0u : (uint32_t)(kAtomBytes + (layer - 1) * 2 * kAtomBytes);
            const uint32_t a_base = tmem_base + (uint32_t)(NTILES * 192 + j * 64);   // lane 0, A columns
            const uint32_t b_base = smem_u32(w_smem + woff);
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t koff = (uint32_t)(k >> 2) * kAtomBytes + (uint32_t)(k & 3) * 32;
              umma_bf16_ts(tmem_base + j * 128, a_base + (uint32_t)k * 8,           // 16 bf16 = 8 columns
                           umma_desc_sw128(b_base + koff), kIdesc, k > 0 ? 1u : 0u);
            }
            // + bias: ones[128 x 16] * biasK[16 x 128] (bf16 hi + lo rows), both operands from SMEM
            umma_bf16(tmem_base + j * 128, umma_desc_k16_noswizzle(smem_u32(ones_smem)),
                      umma_desc_k16_noswizzle(smem_u32(biask_smem + layer * 4096)), kIdesc, 1u);
            umma_commit(bar_acc[j]);
          }
        };

        // Accumulator-ready wait: one warp of the tile polls the mbarrier, the other warps block at a
        // named barrier (BAR.SYNC waits in hardware and costs no issue slots, unlike a try_wait loop).
        uint32_t ph = 0;
        auto wait_accumulator = [&]() {
          if (NTILES == 1) {
            // latency configuration (one tile, SM otherwise idle): every warp polls, which saves the
            // barrier hop after the poller wakes up
            mbar_wait(bar_acc[j], ph);
          } else {
            if (wl == 0) mbar_wait(bar_acc[j], ph);
            named_bar_sync<kTileThreads>(2 + NTILES + j);
          }
          ph ^= 1;
          tc_fence_after();
        };

        // Which of this thread's 16-wide chunks intersect the goal slice / constrained slices
        // (warp-uniform bit masks, bit = sub-chunk), so that chunks outside every lidar do no scoring.
        uint32_t has_goal = 0, has_con = 0;
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          const int lo = o_base + sub * CW, hi = lo + CW;
          if (sc.goal_dist_index >= 0 ? (sc.goal_dist_index >= lo && sc.goal_dist_index < hi)
                                      : (sc.goal_begin < hi && sc.goal_end > lo)) has_goal |= 1u << sub;
          for (int q = 0; q < sc.n_constraints; ++q)
            if (sc.con_begin[q] < hi && sc.con_end[q] > lo) has_con |= 1u << (sub * SIMBA_MAX_CONSTRAINTS + q);
        }

        // the thread that owns the action columns of the layer-0 input prefetches a_{t+1} one step
        // ahead, so the L2 latency is off the critical path (A <= 4 on this path)
        const bool owns_actions = (o_base + OW > O) && (o_base < O + A);
        float act_pf[4] = {0.f, 0.f, 0.f, 0.f};
        auto prefetch_actions = [&](int tn) {
          if (owns_actions && row_ok && tn < H) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (a < A) act_pf[a] = act_ptr[tn * A + a];
          }
        };

        // N(0,1) draws of this thread's OW outputs for the current step. They depend only on
        // (seed, iteration, t, row, o), not on the network, so they are produced one Philox block
        // (8 normals) at a time right after this tile's MMAs have been issued for a hidden layer,
        // i.e. in the shadow of the tensor-core latency instead of inside the head epilogue.
        uint32_t* my_noise = noise_smem + threadIdx.x;      // element pair i at my_noise[i * kEpiThreads]
        float e_pre[kPark ? 1 : OW];                        // the same draws in registers (latency variant)
        const uint64_t seed_reg = kPark ? 0ull : (prm.seed_ptr ? *prm.seed_ptr : prm.seed);
        if constexpr (kPark) {
#pragma unroll
          for (int i = 0; i < OW / 2; ++i) my_noise[i * kEpiThreads] = 0u;
        } else {
#pragma unroll
          for (int i = 0; i < OW; ++i) e_pre[i] = 0.0f;
        }
        auto put_noise = [&](int idx, float a, float b) {   // elements idx, idx + 1 (idx even)
          const uint32_t pk = pack_bf16(a, b);             // both variants use the bf16-rounded draws,
          if constexpr (kPark) {                               // so their rows stay bit-identical
            my_noise[(idx / 2) * kEpiThreads] = pk;
          } else {
            e_pre[idx] = __uint_as_float(pk << 16);
            e_pre[idx + 1] = __uint_as_float(pk & 0xffff0000u);
          }
        };
        auto make_noise = [&](int call, int t) {              // call in [0, OW / 8)
#pragma unroll
          for (int c = 0; c < OW / 8; ++c) {
            if (c != call) continue;
            const int o0 = o_base + c * 8;
            if (o0 >= O) continue;
            if (prm.eps != nullptr) {
              const float* ep = prm.eps + (((int64_t)id.s * H + t) * ((int64_t)g.P * g.N) + id.r_global) * O;
#pragma unroll
              for (int q = 0; q < 8; q += 2)
                put_noise(c * 8 + q, (o0 + q < O) ? ep[o0 + q] : 0.0f, (o0 + q + 1 < O) ? ep[o0 + q + 1] : 0.0f);
            } else {
              float z[8];
#ifdef ABL_NO_PHILOX
#pragma unroll
              for (int q = 0; q < 8; ++q) z[q] = 0.3f + 0.01f * (float)(q + t);
#else
              philox_noise8<true>(kPark ? seed_sh : seed_reg, (uint32_t)id.s, (uint32_t)prm.iteration, (uint32_t)t,
                                  (uint32_t)id.r_global, (uint32_t)(o0 >> 3), z);
#endif
#pragma unroll
              for (int q = 0; q < 8; q += 2) put_noise(c * 8 + q, z[q], z[q + 1]);
            }
          }
        };

        // One pass over this thread's OW state dims in 16-wide chunks. kFirst: load s_0 from global
        // memory; otherwise apply the Gaussian-head update of step t (mlp_ensemble.py:189-193,
        // transition_model.py:75). Either way: store the state to TMEM, write the scaled bf16 input
        // of step t_next into the layer-0 A tile (transition_model.py:72,79-87) and publish the
        // partial lidar minima (closest_distance, safety_gym.py:188-192).
        auto state_pass = [&](auto first_tag, auto sample_tag, int t, int t_next) {
          constexpr bool kFirst = decltype(first_tag)::value;
          constexpr bool kSample = decltype(sample_tag)::value;
          float gmin = INFINITY;
          float cmin[SIMBA_MAX_CONSTRAINTS];
#pragma unroll
          for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q) cmin[q] = INFINITY;
#pragma unroll 1
          for (int sub = 0; sub < NSUB; ++sub) {
            const int oc = o_base + sub * CW;               // first state dim of this chunk
            const bool full = oc + CW <= O;                 // warp-uniform: no padding / action columns
            float sv[CW];
            if (kFirst) {
#pragma unroll
              for (int i = 0; i < CW; ++i) sv[i] = (row_ok && oc + i < O) ? s0_ptr[oc + i] : 0.0f;
            } else {
              uint32_t vm[CW], vv[CW], st[CW];
              tmem_ld<CW>(t_lane + oc, vm);
              if (kSample) tmem_ld<CW>(t_lane + 64 + oc, vv);
              tmem_ld<CW>(t_state + oc, st);
              tmem_ld_wait();
              if (sub == 0) TL(44);
#pragma unroll
              for (int jb = 0; jb < CW / 8; ++jb) {             // one Philox NOISE block = 8 outputs
                const int o0 = oc + jb * 8;
                float d[8];                                  // mu (bias folded into the GEMM)
#pragma unroll
                for (int q = 0; q < 8; ++q) d[q] = __uint_as_float(vm[jb * 8 + q]);
                if (kSample && o0 < O) {
                  float eps8[8];
#pragma unroll
                  for (int q = 0; q < 8; q += 2) {
                    if constexpr (kPark) {
                      const uint32_t pk = my_noise[((sub * CW + jb * 8 + q) / 2) * kEpiThreads];
                      eps8[q] = __uint_as_float(pk << 16);           // bf16 -> fp32: low half first
                      eps8[q + 1] = __uint_as_float(pk & 0xffff0000u);
                    } else {
                      eps8[q] = e_pre[sub * CW + jb * 8 + q];        // NSUB == 1 here: constant index
                      eps8[q + 1] = e_pre[sub * CW + jb * 8 + q + 1];
                    }
                  }
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
#ifdef ABL_NO_SOFTPLUS
                    const float var = __uint_as_float(vv[jb * 8 + q]) * 1e-6f + 3e-4f;
                    d[q] = fmaf(var, eps8[q], d[q]);
#else
                    const float var = softplus_fast(__uint_as_float(vv[jb * 8 + q])) + 1e-4f;
                    d[q] = fmaf(sqrt_approx(var), eps8[q], d[q]);
#endif
                  }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float old = __uint_as_float(st[jb * 8 + q]);
                  sv[jb * 8 + q] = (full || o0 + q < O) ? old + d[q] : old;
                }
              }
            }
            {
              uint32_t st[CW];
#pragma unroll
              for (int i = 0; i < CW; ++i) st[i] = __float_as_uint(sv[i]);
              tmem_st<CW>(t_state + oc, st);
              if (sub == 0) TL(45);
            }
            // ---- partial lidar minima, only for chunks that intersect a slice -----------------
#ifdef ABL_NO_SCORE
            const uint32_t con_bits = 0;
            if (false) {
#else
            const uint32_t con_bits = (has_con >> (sub * SIMBA_MAX_CONSTRAINTS)) & 0xFu;
            if (((has_goal >> sub) & 1u) | con_bits) {
#endif
              if (sc.goal_dist_index >= 0) {
#pragma unroll
                for (int i = 0; i < CW; ++i)
                  if (oc + i == sc.goal_dist_index) gmin = fmaxf(sv[i], 0.0f);  // safety_gym.py:172-174
              }
              float v[CW];
#pragma unroll
              for (int i = 0; i < CW; ++i) {
                const float w = __fsub_rn(D, __fmul_rn(D, __fsub_rn(1.0f, sv[i])));
                v[i] = fminf(fmaxf(w, 0.0f), D);
              }
              // pen[k][o] = 0 inside slice k, +inf outside: min(v + pen) is the slice minimum
              if (((has_goal >> sub) & 1u) && sc.goal_dist_index < 0) {
#pragma unroll
                for (int i4 = 0; i4 < CW / 4; ++i4) {
                  const float4 pn = *reinterpret_cast<const float4*>(pen_smem + oc + i4 * 4);
                  gmin = fminf(gmin, fminf(fminf(v[i4 * 4] + pn.x, v[i4 * 4 + 1] + pn.y),
                                           fminf(v[i4 * 4 + 2] + pn.z, v[i4 * 4 + 3] + pn.w)));
                }
              }
#pragma unroll
              for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q) {
                if ((con_bits >> q) & 1u) {
#pragma unroll
                  for (int i4 = 0; i4 < CW / 4; ++i4) {
                    const float4 pn = *reinterpret_cast<const float4*>(pen_smem + (1 + q) * 64 + oc + i4 * 4);
                    cmin[q] = fminf(cmin[q], fminf(fminf(v[i4 * 4] + pn.x, v[i4 * 4 + 1] + pn.y),
                                                   fminf(v[i4 * 4 + 2] + pn.z, v[i4 * 4 + 3] + pn.w)));
                  }
                }
              }
            }
            // ---- scaled bf16 input of the next step ---------------------------------------------
            if (t_next < H) {
              float x[CW];
              if (full) {
#pragma unroll
                for (int i4 = 0; i4 < CW / 4; ++i4) {
                  const float4 sa = *reinterpret_cast<const float4*>(scale_smem + oc + i4 * 4);
                  const float4 sb = *reinterpret_cast<const float4*>(scale_smem + 64 + oc + i4 * 4);
                  x[i4 * 4 + 0] = fmaf(sv[i4 * 4 + 0], sa.x, sb.x);
                  x[i4 * 4 + 1] = fmaf(sv[i4 * 4 + 1], sa.y, sb.y);
                  x[i4 * 4 + 2] = fmaf(sv[i4 * 4 + 2], sa.z, sb.z);
                  x[i4 * 4 + 3] = fmaf(sv[i4 * 4 + 3], sa.w, sb.w);
                }
              } else {
#pragma unroll
                for (int i = 0; i < CW; ++i) {
                  const int o = oc + i;
                  float xin = sv[i];                                             // zero beyond O
                  if (o >= O && o < O + A) xin = act_pf[(o - O) & 3];            // prefetched a_{t_next}
                  x[i] = fmaf(xin, scale_smem[o], scale_smem[64 + o]);           // padded k: a = b = 0
                }
              }
              uint32_t pk[CW / 2];
#pragma unroll
              for (int c = 0; c < CW / 2; ++c) pk[c] = pack_bf16(x[2 * c], x[2 * c + 1]);
              tmem_st<CW / 2>(t_a + oc / 2, pk);            // K elements [oc, oc + CW) of the layer-0 input
            }
          }
          TL(46);
          tmem_st_wait();
          TL(47);
          part[0] = gmin;
#pragma unroll
          for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q)
            if (q < sc.n_constraints) part[(1 + q) * 128] = cmin[q];
        };
        // combine the Q partials of this row (group-0 thread only): goal distance and cost(s)
        auto combine = [&](float& dist, float& cost) {
          float gmin = INFINITY;
#pragma unroll
          for (int c = 0; c < Q; ++c) gmin = fminf(gmin, part_row[(c * nparts) * 128]);
          float cst = 0.0f;
          for (int q = 0; q < sc.n_constraints; ++q) {
            float m = INFINITY;
#pragma unroll
            for (int c = 0; c < Q; ++c) m = fminf(m, part_row[(c * nparts + 1 + q) * 128]);
            cst += (m <= sc.con_size[q]) ? 1.0f : 0.0f;
          }
          dist = gmin;
          cost = sc.constrain_indicator ? (cst > 0.0f ? 1.0f : 0.0f) : cst;
        };

        constexpr int kRsStride = NTILES * 128;
        uint32_t* my_rs = rs_smem + j * 128 + r;            // cum, costsum, cmask lo / hi, dist, cost, done
        RowScore rs_reg;                                    // the same in registers (latency variant)
        rs_reg.cum = 0.0f; rs_reg.costsum = 0.0f; rs_reg.cmask = 0ull; rs_reg.done = false;
        rs_reg.dist = 0.0f; rs_reg.cost = 0.0f;
        prefetch_actions(0);
        state_pass(std::true_type{}, std::false_type{}, 0, 0);
        if (issuer) mbar_wait(bar_w, 0);                 // weights have landed before the first MMA
        tile_sync_and_issue(0);                          // also orders the partials for combine()
        if (cgp == 0) {
          float d0, c0;
          combine(d0, c0);
          if constexpr (kPark) {
            my_rs[0] = 0u; my_rs[kRsStride] = 0u; my_rs[2 * kRsStride] = 0u; my_rs[3 * kRsStride] = 0u;
            my_rs[4 * kRsStride] = __float_as_uint(d0);
            my_rs[5 * kRsStride] = __float_as_uint(c0);
            my_rs[6 * kRsStride] = 0u;
            my_rs[7 * kRsStride] = (uint32_t)id.out;        // output slot of this row, needed again at the end
          } else {
            rs_reg.dist = d0;
            rs_reg.cost = c0;
          }
        }
        // (the partials are next written after the tile has passed L more tile barriers, which the
        //  group-0 threads reading here reach only after combine())

        for (int t = 0; t < H; ++t) {
#ifdef SIMBA_TC_TIMELINE
          tl_t = t;
#endif
          TL(0);
          prefetch_actions(t + 1);
          // ---- hidden layers: TMEM -> +bias, ReLU, bf16 -> next A operand -----------------------
          for (int l = 0; l < L; ++l) {
            wait_accumulator();
            TL(1 + l * 4);
#pragma unroll
            for (int cc = 0; cc < HC; ++cc) {
              const int col0 = (cgp * HC + cc) * HW;       // first accumulator column of this chunk
              uint32_t v[HW];
              tmem_ld<HW>(t_lane + col0, v);
              tmem_ld_wait();
              uint32_t pk[HW / 2];                         // bias is already in the accumulator
#pragma unroll
              for (int q = 0; q < HW / 2; ++q)
                pk[q] = pack_relu_bf16(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
              tmem_st<HW / 2>(t_a + col0 / 2, pk);          // K elements [col0, col0 + HW) of this row
            }
            TL(2 + l * 4);
            tile_sync_and_issue(l + 1);
            if (prm.sampling_propagation) {
              // spread the OW / 8 Philox blocks over the hidden layers (the last layer takes the rest)
#pragma unroll
              for (int c = 0; c < OW / 8; ++c)
                if ((c < L - 1 ? c : L - 1) == l) make_noise(c, t);
            }
            TL(3 + l * 4);
          }

          // ---- Gaussian heads + state update + next input + partial minima (one fused pass) --------
          wait_accumulator();
          TL(40);
          if (prm.sampling_propagation) state_pass(std::false_type{}, std::true_type{}, t, t + 1);
          else state_pass(std::false_type{}, std::false_type{}, t, t + 1);

          // ---- next step's layer-0 MMA goes out first; then scoring of (s_t, s_{t+1}):
          //      safety_gym.py:110-166, per-row objective ------------------------------------------------
          TL(41);
          if (t + 1 < H) {
            tile_sync_and_issue(0);
            TL(42);
          } else {
            named_bar_sync<kTileThreads>(2 + j);
          }
          if (cgp == 0) {
            float next_dist, next_cost;
            combine(next_dist, next_cost);
            RowScore rs = rs_reg;
            if constexpr (kPark) {
              rs.cum = __uint_as_float(my_rs[0]);
              rs.costsum = __uint_as_float(my_rs[kRsStride]);
              rs.cmask = (uint64_t)my_rs[2 * kRsStride] | ((uint64_t)my_rs[3 * kRsStride] << 32);
              rs.dist = __uint_as_float(my_rs[4 * kRsStride]);
              rs.cost = __uint_as_float(my_rs[5 * kRsStride]);
              rs.done = my_rs[6 * kRsStride] != 0u;
            }
            const bool goal = rs.dist <= sc.goal_threshold;
            const float rew = step_reward(sc, rs.dist, next_dist, goal);
            if (done_first) {                                  // safe_cem_mpc.py:87-93
              rs.done = rs.done || goal;
              if (!rs.done && rs.cost > 0.0f) rs.cmask |= (1ull << t);
              rs.cum += rs.done ? 0.0f : rew;
            } else {                                           // mpc_policy.py:35-37
              rs.cum += rs.done ? 0.0f : rew;
              if (!rs.done && rs.cost > 0.0f) rs.cmask |= (1ull << t);
              rs.done = rs.done || goal;
            }
            rs.costsum += rs.cost;
            if constexpr (kPark) {
              my_rs[0] = __float_as_uint(rs.cum);
              my_rs[kRsStride] = __float_as_uint(rs.costsum);
              my_rs[2 * kRsStride] = (uint32_t)rs.cmask;
              my_rs[3 * kRsStride] = (uint32_t)(rs.cmask >> 32);
              my_rs[4 * kRsStride] = __float_as_uint(next_dist);
              my_rs[5 * kRsStride] = __float_as_uint(next_cost);
              my_rs[6 * kRsStride] = rs.done ? 1u : 0u;
            } else {
              rs.dist = next_dist;
              rs.cost = next_cost;
              rs_reg = rs;
            }
          }
          TL(43);
        }
        if (cgp == 0 && row_ok && prm.row_return != nullptr) {
          if constexpr (kPark) {
            const uint32_t out_slot = my_rs[7 * kRsStride];
            prm.row_return[out_slot] = __uint_as_float(my_rs[0]);
            prm.row_costmask[out_slot] = (uint64_t)my_rs[2 * kRsStride] | ((uint64_t)my_rs[3 * kRsStride] << 32);
            prm.row_costsum[out_slot] = __uint_as_float(my_rs[kRsStride]);
          } else {
            prm.row_return[id.out] = rs_reg.cum;
            prm.row_costmask[id.out] = rs_reg.cmask;
            prm.row_costsum[id.out] = rs_reg.costsum;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- host side ------------------------------------------------------------------------------------
static size_t tc_smem_bytes(int L, int ntiles, int q, int nparts) {
  size_t b = (size_t)kAtomBytes + (size_t)L * 2 * kAtomBytes;      // weights
  b += (size_t)(L + 1) * 4096 + 4096;                              // bias K-blocks + ones tile
  b += 128 * sizeof(float);                                        // scaler
  b += kParts * 64 * sizeof(float);                                // slice penalty table
  b += (size_t)ntiles * q * nparts * 128 * sizeof(float);          // partial minima exchange
  if (ntiles > 1) {                                                // parking areas of the two-tile variant
    b += (size_t)(64 / q / 2) * (ntiles * q * 128) * sizeof(uint32_t);   // bf16x2 noise of the current step
    b += (size_t)8 * ntiles * 128 * sizeof(uint32_t);                    // per-row running objective
  }
  b += (1 + 2 * ntiles) * sizeof(uint64_t) + 2 * sizeof(uint32_t) + ntiles * sizeof(TileInfo);
  return b + 1024;                                                 // alignment slack
}

constexpr size_t kMaxSmem = 227 * 1024;

bool rollout_tc_supported(int O, int A, int L, int U, int H) {
  // narrower hidden layers run zero-padded to 128 units (simba_model_commit pads the images); the
  // member's whole weight set must fit in shared memory next to the one-tile variant's buffers
  return U >= 1 && U <= kU && O >= 1 && O <= kMaxO && O + A <= 64 && A <= 4 && L >= 1 && H >= 1 && H <= 64 &&
         tc_smem_bytes(L, 1, 4, kParts) + 1024 <= kMaxSmem;
}

// whether the two-tile throughput variant (and its shared-memory parking areas) fits for this depth
bool rollout_tc_two_tiles_fit(int L, int n_constraints) {
  return tc_smem_bytes(L, 2, 2, 1 + n_constraints) + 1024 <= kMaxSmem;
}

template <int NTILES, int Q>
static cudaError_t launch_variant(const RolloutParams& prm, int n_tiles, cudaStream_t stream) {
  const size_t smem = tc_smem_bytes(prm.L, NTILES, Q, 1 + prm.scorer.n_constraints);
  // set on every launch: the attribute is per device and per function, and launches happen only at
  // graph capture or in the non-graph entry points, never on the replayed hot path
  {
    cudaError_t e = cudaFuncSetAttribute(rollout_tc_kernel<NTILES, Q>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int grid = (n_tiles + NTILES - 1) / NTILES;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTILES * Q * 128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = prm.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rollout_tc_kernel<NTILES, Q>, prm);
}

cudaError_t launch_rollout_tc(const RolloutParams& prm, int n_tiles, cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  if (prm.tc_tiles_per_cta == 2) return launch_variant<2, 2>(prm, n_tiles, stream);
  return launch_variant<1, 4>(prm, n_tiles, stream);   // (<1,8>, 1024 threads, measured 15 % slower)
}

}  // namespace simba
