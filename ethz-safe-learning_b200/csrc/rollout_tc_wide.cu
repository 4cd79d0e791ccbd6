// bf16 tcgen05 rollout + scoring kernel for WIDE ensembles (128 < units <= 416, e.g. the 4x400
// model of BASELINE configs[4]) — same contract as rollout_tc.cu / rollout_f32.cu.
//
// A member's weights (1.2 MB for 4x400) no longer fit in shared memory, and a 400-column fp32
// accumulator fills tensor memory, so the kernel is a warp-specialised streaming pipeline:
//
//   * producer warp: streams the member's pre-swizzled UMMA weight tiles ([units n x 64 k] bf16,
//     51 KB) from L2 through a two-stage TMA ring (cp.async.bulk + full / empty mbarriers). The
//     tile sequence of a rollout step (layer 0, (L-1) x KA hidden K-blocks, KA head K-blocks) is
//     the same for every step, so the producer just cycles through it, decoupled from the math;
//   * MMA warp (one elected thread): per layer waits for "A ready", then per K-block waits for the
//     stage, issues 4 x tcgen05.mma (K = 16) per N half (N = 208 + 192 for 400 units; N = 128 for
//     the heads) with both operands in shared memory, and tcgen05.commit's the stage back to the
//     producer; after the last K-block it commits "accumulator ready";
//   * 4 x 128 epilogue threads (Q = 4 per rollout row, as in rollout_tc.cu): drain the fp32
//     accumulator row from TMEM, ReLU, pack bf16 and store 16-byte chunks into the 128B-swizzled
//     K-major A tile in shared memory (7 atoms of [128 x 64] for 400 units), fence to the async
//     proxy, meet at a named barrier; one thread arrives on "A ready". The Gaussian-head / state /
//     scoring pass is the one of rollout_tc.cu (state in spare TMEM columns, Philox noise in the
//     shadow of the hidden layers), writing the next step's scaled input into atom 0.
//   * biases ride in the K padding: the A tile holds constant ones at k = K_real, K_real + 1 and
//     the weight tiles hold bf16(b) and bf16(b - hi) in those two k rows, so no epilogue touches
//     a bias (K_real = 62 -> 64 for layer 0, 400 -> 448 for the others).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "rollout_params.cuh"
#include "tc_ptx.cuh"

namespace simba {

// Debug timeline (compile with -DSIMBA_TC_TIMELINE): epilogue thread 0 and the last epilogue warp's
// lane 0 of CTA 0 stamp clock64() at phase boundaries into prm.timeline (tools/tc_timeline.py c5).
#ifdef SIMBA_TC_TIMELINE
#define TLW(ev)                                                                                  \
  do {                                                                                           \
    if (tl_who >= 0 && blockIdx.x == 0)                                                          \
      prm.timeline[(tl_who * 64 + tl_t) * 64 + (ev)] = clock64();  \
  } while (0)
#else
#define TLW(ev) do { } while (0)
#endif

namespace {

constexpr int kQ = 4;                         // epilogue threads per rollout row
constexpr int kEpiThreads = kQ * 128;
constexpr int kWideThreads = kEpiThreads + 64;   // + producer warp + MMA warp
constexpr int kAtomBytes = 128 * 128;         // one 64-wide K atom of a 128-row tile (bf16, SW128)
constexpr int kStages = 2;
constexpr int kParts = 1 + SIMBA_MAX_CONSTRAINTS;
constexpr int kStateCol = 448;                // TMEM: accumulator columns [0, units), state [448, 512)
constexpr int kTmemCols = 512;
constexpr int OW = 64 / kQ;                   // head outputs / state dims / layer-0 K elements per thread
constexpr int CW = 16;

constexpr int kHeadGroup = 3;                 // head K-blocks (16 KB each) fetched per TMA tile

struct WideShape {
  int U, KA, L;
  uint32_t hid_tile_bytes;      // [U x 64] bf16
  uint32_t head_tile_bytes;     // [128 x 64] bf16 (one K-block; kHeadGroup of them travel together)
  int head_tiles;               // ceil(KA / kHeadGroup)
  int tiles_per_step;
};

__host__ __device__ inline WideShape wide_shape(int U, int L) {
  WideShape s;
  s.U = U; s.L = L;
  s.KA = (U + 2 + 63) / 64;
  s.hid_tile_bytes = (uint32_t)U * 128u;
  s.head_tile_bytes = 128u * 128u;
  s.head_tiles = (s.KA + kHeadGroup - 1) / kHeadGroup;
  s.tiles_per_step = 1 + (L - 1) * s.KA + s.head_tiles;
  return s;
}

__host__ __device__ inline uint32_t wide_stage_bytes(const WideShape& s) {
  const uint32_t head = (uint32_t)kHeadGroup * s.head_tile_bytes;
  const uint32_t big = s.hid_tile_bytes > head ? s.hid_tile_bytes : head;
  return (big + 1023u) & ~1023u;
}

// source offset / size of tile `i` of the per-step stream
__device__ __forceinline__ void wide_tile(const WideShape& s, int i, uint32_t& off, uint32_t& bytes) {
  const int n_hid = 1 + (s.L - 1) * s.KA;
  if (i < n_hid) { off = (uint32_t)i * s.hid_tile_bytes; bytes = s.hid_tile_bytes; return; }
  // the heads' K-blocks are small (16 KB), so kHeadGroup consecutive ones travel as one tile: with
  // one block per tile the two-stage ring was latency-bound there (4.9 k cycles for 2.6 k of MMAs)
  const int ht = i - n_hid;
  const int blocks = min(kHeadGroup, s.KA - ht * kHeadGroup);
  off = (uint32_t)n_hid * s.hid_tile_bytes + (uint32_t)(ht * kHeadGroup) * s.head_tile_bytes;
  bytes = (uint32_t)blocks * s.head_tile_bytes;
}

struct TileInfoW {
  int32_t member, k0, count, valid;
};

}  // namespace

__global__ void __launch_bounds__(kWideThreads, 1) rollout_tc_wide_kernel(const RolloutParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const RowGeom& g = prm.g;
  const int L = prm.L, U = prm.U;
  const int H = g.H;
  const int O = g.O, A = g.A;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WideShape ws = wide_shape(U, L);
  const int KA = ws.KA;

  // ---- shared memory carve-up (base is 1024-aligned: required by SWIZZLE_128B) -------------------
  uint8_t* a_smem = smem_raw;                                              // [KA][128 x 64] bf16 SW128
  uint8_t* b_smem = a_smem + (size_t)KA * kAtomBytes;                      // [kStages][U x 64] bf16 SW128
  const uint32_t stage_bytes = wide_stage_bytes(ws);
  float* scale_smem = reinterpret_cast<float*>(b_smem + (size_t)kStages * stage_bytes);   // [2][64]
  float* pen_smem = scale_smem + 128;                                      // [kParts][64]
  const int nparts = 1 + prm.scorer.n_constraints;
  float* part_smem = pen_smem + kParts * 64;                               // [kQ][nparts][128]
  // running objective of every rollout row (RowScore fields, field-major), see rollout_tc.cu
  uint32_t* rs_smem = reinterpret_cast<uint32_t*>(part_smem + kQ * nparts * 128);   // [8][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rs_smem + 8 * 128);
  // bars: [0,1] full, [2,3] empty, [4] A ready, [5] accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  TileInfoW* tinfo = reinterpret_cast<TileInfoW*>(tmem_slot + 2);

  const uint32_t bar_full[kStages] = {smem_u32(&bars[0]), smem_u32(&bars[1])};
  const uint32_t bar_empty[kStages] = {smem_u32(&bars[2]), smem_u32(&bars[3])};
  const uint32_t bar_a = smem_u32(&bars[4]);
  const uint32_t bar_acc = smem_u32(&bars[5]);

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full[s], 1); mbar_init(bar_empty[s], 1); }
    mbar_init(bar_a, 1);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  pdl_wait_prior_grid();
  if (threadIdx.x == 0) {
    TileInfoW info{0, 0, 0, 0};
    if ((int)blockIdx.x < prm.n_tiles) {
      const Tile t = prm.tiles[blockIdx.x];
      info.member = t.member; info.k0 = t.k0; info.count = t.count;
      bool any = t.count > 0;
      if (any && prm.active != nullptr) {                     // cem_mpc.py:66-67 early exit
        const int m = g.rows_per_state[t.member];
        any = false;
        for (int s = t.k0 / m; s <= (t.k0 + t.count - 1) / m; ++s) any = any || prm.active[s] != 0;
      }
      info.valid = any ? 1 : 0;
    }
    tinfo[0] = info;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const TileInfoW ti = tinfo[0];

  if (ti.valid) {
    if (warp == kEpiThreads / 32) {
      // ===================== producer warp: weight tiles through the TMA ring =====================
      if (lane == 0) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(prm.w_wide) +
                              (size_t)ti.member * prm.w_wide_member_bytes;
        const int total = H * ws.tiles_per_step;
        int in_step = 0;
        for (int i = 0; i < total; ++i) {
          const int s = i % kStages;
          if (i >= kStages) mbar_wait(bar_empty[s], ((i / kStages) - 1) & 1);
          uint32_t off, bytes;
          wide_tile(ws, in_step, off, bytes);
          mbar_expect_tx(bar_full[s], bytes);
          bulk_g2s(smem_u32(b_smem + (size_t)s * stage_bytes), wsrc + off, bytes, bar_full[s]);
          if (++in_step == ws.tiles_per_step) in_step = 0;
        }
      }
    } else if (warp == kEpiThreads / 32 + 1) {
      // ===================== MMA warp: one elected thread issues every tcgen05.mma ================
      if (lane == 0) {
        const int N1 = ((U / 2 + 15) / 16) * 16, N2 = U - N1;      // two N halves, multiples of 16
        const uint32_t idesc1 = umma_idesc_n((uint32_t)N1), idesc2 = umma_idesc_n((uint32_t)N2);
        const uint32_t idesc_head = umma_idesc_n(128u);
        const uint32_t a_base = smem_u32(a_smem);
        int tile = 0;                                              // running index into the ring
        uint32_t a_phase = 0;
        for (int t = 0; t < H; ++t) {
          for (int layer = 0; layer <= L; ++layer) {
            mbar_wait(bar_a, a_phase);                             // this layer's A tile is complete
            a_phase ^= 1;
            tc_fence_after();
            if (layer < L) {
              const int kblocks = layer == 0 ? 1 : KA;
              for (int kb = 0; kb < kblocks; ++kb, ++tile) {
                const int s = tile % kStages;
                mbar_wait(bar_full[s], (tile / kStages) & 1);
                tc_fence_after();
                const uint32_t b_base = smem_u32(b_smem + (size_t)s * stage_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {                      // UMMA_K = 16 -> 32 bytes along K
                  const uint64_t a_desc = umma_desc_sw128(a_base + (uint32_t)kb * kAtomBytes + (uint32_t)k * 32);
                  const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                  umma_bf16(tmem_base, a_desc, umma_desc_sw128(b_base + (uint32_t)k * 32), idesc1, acc);
                  umma_bf16(tmem_base + (uint32_t)N1, a_desc,
                            umma_desc_sw128(b_base + (uint32_t)N1 * 128u + (uint32_t)k * 32), idesc2, acc);
                }
                umma_commit(bar_empty[s]);                         // stage free once these MMAs retire
              }
            } else {
              for (int ht = 0; ht < ws.head_tiles; ++ht, ++tile) {
                const int s = tile % kStages;
                mbar_wait(bar_full[s], (tile / kStages) & 1);
                tc_fence_after();
                const uint32_t b_base = smem_u32(b_smem + (size_t)s * stage_bytes);
                const int blocks = min(kHeadGroup, KA - ht * kHeadGroup);
                for (int j = 0; j < blocks; ++j) {
                  const int kb = ht * kHeadGroup + j;
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, umma_desc_sw128(a_base + (uint32_t)kb * kAtomBytes + (uint32_t)k * 32),
                              umma_desc_sw128(b_base + (uint32_t)j * ws.head_tile_bytes + (uint32_t)k * 32),
                              idesc_head, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty[s]);
              }
            }
            umma_commit(bar_acc);                                  // accumulator of this layer ready
          }
        }
      }
    } else {
      // ============ epilogue warps: kQ threads per rollout row, each owns a column slice ============
      const int wl = warp;                                  // warp within the tile
      const int cgp = wl >> 2;                              // column group in [0, kQ)
      const int r = (wl & 3) * 32 + lane;                   // row in tile == TMEM lane
      // scaler / penalty tables, zeroed A tile with the constant ones of the bias rows
      {
        const int IN = O + A;
        for (int i = threadIdx.x; i < 64; i += kEpiThreads) {
          const bool one = (i == IN || i == IN + 1);          // layer-0 bias rows (k = IN, IN + 1)
          scale_smem[i] = one ? 0.0f : prm.tc_scale_a[i];
          scale_smem[64 + i] = one ? 1.0f : prm.tc_scale_b[i];
          const simba_scorer_t& scc = prm.scorer;
          pen_smem[i] = (i >= scc.goal_begin && i < scc.goal_end) ? 0.0f : INFINITY;
          for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q)
            pen_smem[(1 + q) * 64 + i] =
                (q < scc.n_constraints && i >= scc.con_begin[q] && i < scc.con_end[q]) ? 0.0f : INFINITY;
        }
        uint32_t* a32 = reinterpret_cast<uint32_t*>(a_smem);
        for (int i = threadIdx.x; i < KA * kAtomBytes / 4; i += kEpiThreads) a32[i] = 0u;
        named_bar_sync<kEpiThreads>(1);
        // ones at k = U, U + 1 of every row (same 16-byte chunk; U is a multiple of 8)
        if (threadIdx.x < 128) {
          const int row = threadIdx.x, c8 = U >> 3, atom = c8 >> 3, cin = c8 & 7;
          uint32_t* p = reinterpret_cast<uint32_t*>(a_smem + (size_t)atom * kAtomBytes + (row >> 3) * 1024 +
                                                    (row & 7) * 128 + ((cin ^ (row & 7)) * 16));
          p[0] = 0x3F803F80u;                                 // bf16 1.0, 1.0
        }
        fence_proxy_async();
        named_bar_sync<kEpiThreads>(1);
      }
      const bool row_ok = r < ti.count;
      const RowId id = decode_row(g, ti.member, ti.k0 + (row_ok ? r : 0));
      const uint64_t seed = prm.seed_ptr ? *prm.seed_ptr : prm.seed;
      const uint32_t t_lane = tmem_base + ((uint32_t)((wl & 3) * 32) << 16);
      const uint32_t t_state = t_lane + (uint32_t)kStateCol;
      const float* act_ptr = prm.actions + ((int64_t)id.s * g.N + id.i_global) * prm.action_stride;
      const bool done_first = objective_done_first(prm.objective);
      const simba_scorer_t& sc = prm.scorer;
      const int o_base = cgp * OW;
      float* part = part_smem + (cgp * nparts) * 128 + r;
      const float* part_row = part_smem + r;
      const float D = sc.lidar_max_dist;
      const float* s0_ptr = prm.states + (prm.state_per_row ? id.r_global : (int64_t)id.s) * prm.state_stride;
      const uint32_t a_row = smem_u32(a_smem) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
      const uint32_t rsw = (uint32_t)(r & 7);

      // 16-byte chunk `c8` (k = 8 c8 .. 8 c8 + 7) of this row in the swizzled A tile
      auto a_chunk = [&](int c8) -> uint32_t {
        return a_row + (uint32_t)(c8 >> 3) * kAtomBytes + ((((uint32_t)c8 & 7u) ^ rsw) << 4);
      };
      // this layer's A tile is complete: fence, meet, one thread tells the MMA warp
      auto publish_a = [&]() {
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync<kEpiThreads>(2);
        if (threadIdx.x == 0) mbar_arrive(bar_a);
      };
      uint32_t acc_ph = 0;
      auto wait_accumulator = [&]() {
        mbar_wait(bar_acc, acc_ph);
        acc_ph ^= 1;
        tc_fence_after();
      };

      uint32_t has_goal = 0, has_con = 0;
      {
        const int lo = o_base, hi = lo + CW;
        if (sc.goal_dist_index >= 0 ? (sc.goal_dist_index >= lo && sc.goal_dist_index < hi)
                                    : (sc.goal_begin < hi && sc.goal_end > lo)) has_goal = 1u;
        for (int q = 0; q < sc.n_constraints; ++q)
          if (sc.con_begin[q] < hi && sc.con_end[q] > lo) has_con |= 1u << q;
      }
      const bool owns_actions = (o_base + OW > O) && (o_base < O + A);
      float act_pf[4] = {0.f, 0.f, 0.f, 0.f};
      auto prefetch_actions = [&](int tn) {
        if (owns_actions && row_ok && tn < H) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (a < A) act_pf[a] = act_ptr[tn * A + a];
        }
      };
      float e_pre[OW];
#pragma unroll
      for (int i = 0; i < OW; ++i) e_pre[i] = 0.0f;
      auto make_noise = [&](int call, int t) {              // call in [0, OW / 8)
#pragma unroll
        for (int c = 0; c < OW / 8; ++c) {
          if (c != call) continue;
          const int o0 = o_base + c * 8;
          if (o0 >= O) continue;
          if (prm.eps != nullptr) {
            const float* ep = prm.eps + (((int64_t)id.s * H + t) * ((int64_t)g.P * g.N) + id.r_global) * O;
#pragma unroll
            for (int q = 0; q < 8; ++q) e_pre[c * 8 + q] = (o0 + q < O) ? ep[o0 + q] : 0.0f;
          } else {
            float z[8];
            philox_noise8<true>(seed, (uint32_t)id.s, (uint32_t)prm.iteration, (uint32_t)t,
                                (uint32_t)id.r_global, (uint32_t)(o0 >> 3), z);
#pragma unroll
            for (int q = 0; q < 8; ++q) e_pre[c * 8 + q] = z[q];
          }
        }
      };

      // Head / state pass over this thread's 16 state dims (see rollout_tc.cu state_pass): kFirst
      // loads s_0; otherwise s += mu (+ sqrt(var) * eps). Stores the state to TMEM, writes the
      // scaled bf16 input of step t_next into atom 0 of the A tile and publishes the lidar minima.
      auto state_pass = [&](auto first_tag, auto sample_tag, int t_next) {
        constexpr bool kFirst = decltype(first_tag)::value;
        constexpr bool kSample = decltype(sample_tag)::value;
        float gmin = INFINITY;
        float cmin[SIMBA_MAX_CONSTRAINTS];
#pragma unroll
        for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q) cmin[q] = INFINITY;
        const int oc = o_base;
        const bool full = oc + CW <= O;
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
#pragma unroll
          for (int jb = 0; jb < CW / 8; ++jb) {
            const int o0 = oc + jb * 8;
            float d[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) d[q] = __uint_as_float(vm[jb * 8 + q]);
            if (kSample && o0 < O) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float var = softplus_fast(__uint_as_float(vv[jb * 8 + q])) + 1e-4f;
                d[q] = fmaf(sqrt_approx(var), e_pre[jb * 8 + q], d[q]);
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
        }
        if (has_goal | has_con) {
          if (sc.goal_dist_index >= 0) {
#pragma unroll
            for (int i = 0; i < CW; ++i)
              if (oc + i == sc.goal_dist_index) gmin = fmaxf(sv[i], 0.0f);   // safety_gym.py:172-174
          }
          float v[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const float w = __fsub_rn(D, __fmul_rn(D, __fsub_rn(1.0f, sv[i])));
            v[i] = fminf(fmaxf(w, 0.0f), D);
          }
          if (has_goal && sc.goal_dist_index < 0) {
#pragma unroll
            for (int i4 = 0; i4 < CW / 4; ++i4) {
              const float4 pn = *reinterpret_cast<const float4*>(pen_smem + oc + i4 * 4);
              gmin = fminf(gmin, fminf(fminf(v[i4 * 4] + pn.x, v[i4 * 4 + 1] + pn.y),
                                       fminf(v[i4 * 4 + 2] + pn.z, v[i4 * 4 + 3] + pn.w)));
            }
          }
#pragma unroll
          for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q) {
            if ((has_con >> q) & 1u) {
#pragma unroll
              for (int i4 = 0; i4 < CW / 4; ++i4) {
                const float4 pn = *reinterpret_cast<const float4*>(pen_smem + (1 + q) * 64 + oc + i4 * 4);
                cmin[q] = fminf(cmin[q], fminf(fminf(v[i4 * 4] + pn.x, v[i4 * 4 + 1] + pn.y),
                                               fminf(v[i4 * 4 + 2] + pn.z, v[i4 * 4 + 3] + pn.w)));
              }
            }
          }
        }
        if (t_next < H) {
          float x[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const int o = oc + i;
            float xin = sv[i];                                             // zero beyond O
            if (o >= O && o < O + A) xin = act_pf[(o - O) & 3];            // prefetched a_{t_next}
            x[i] = fmaf(xin, scale_smem[o], scale_smem[64 + o]);           // bias rows: a = 0, b = 1
          }
#pragma unroll
          for (int c = 0; c < CW / 8; ++c)
            st_shared_v4(a_chunk((oc >> 3) + c), pack_bf16(x[8 * c], x[8 * c + 1]),
                         pack_bf16(x[8 * c + 2], x[8 * c + 3]), pack_bf16(x[8 * c + 4], x[8 * c + 5]),
                         pack_bf16(x[8 * c + 6], x[8 * c + 7]));
        }
        tmem_st_wait();
        part[0] = gmin;
#pragma unroll
        for (int q = 0; q < SIMBA_MAX_CONSTRAINTS; ++q)
          if (q < sc.n_constraints) part[(1 + q) * 128] = cmin[q];
      };
      auto combine = [&](float& dist, float& cost) {
        float gmin = INFINITY;
#pragma unroll
        for (int c = 0; c < kQ; ++c) gmin = fminf(gmin, part_row[(c * nparts) * 128]);
        float cst = 0.0f;
        for (int q = 0; q < sc.n_constraints; ++q) {
          float m = INFINITY;
#pragma unroll
          for (int c = 0; c < kQ; ++c) m = fminf(m, part_row[(c * nparts + 1 + q) * 128]);
          cst += (m <= sc.con_size[q]) ? 1.0f : 0.0f;
        }
        dist = gmin;
        cost = sc.constrain_indicator ? (cst > 0.0f ? 1.0f : 0.0f) : cst;
      };

      uint32_t* my_rs = rs_smem + r;                      // cum, costsum, cmask lo / hi, dist, cost, done
      prefetch_actions(0);
      state_pass(std::true_type{}, std::false_type{}, 0);
      publish_a();                                       // layer-0 input of step 0 (also orders the partials)
      if (cgp == 0) {
        float d0, c0;
        combine(d0, c0);
        my_rs[0] = 0u; my_rs[128] = 0u; my_rs[256] = 0u; my_rs[384] = 0u;
        my_rs[512] = __float_as_uint(d0);
        my_rs[640] = __float_as_uint(c0);
        my_rs[768] = 0u;
      }

#ifdef SIMBA_TC_TIMELINE
      const int tl_who = (lane == 0) ? (wl == 0 ? 0 : (wl == 4 * kQ - 1 ? 1 : -1)) : -1;
      int tl_t = 0;
#endif
      const int n_chunks = U >> 3;                       // 16-byte chunks per hidden activation row
      const int my_c0 = cgp * (n_chunks / kQ) + min(cgp, n_chunks % kQ);
      const int my_c1 = my_c0 + n_chunks / kQ + (cgp < n_chunks % kQ ? 1 : 0);
      for (int t = 0; t < H; ++t) {
#ifdef SIMBA_TC_TIMELINE
        tl_t = t;
#endif
        TLW(0);
        prefetch_actions(t + 1);
        // ---- hidden layers: TMEM -> ReLU -> bf16 -> swizzled A tile (bias is in the accumulator) ----
        for (int l = 0; l < L; ++l) {
          wait_accumulator();
          TLW(1 + l * 4);
          // this column group's contiguous run of 16-byte chunks, eight chunks (64 columns = two
          // 32-column TMEM loads) in flight at a time
          for (int c0 = my_c0; c0 < my_c1; c0 += 8) {
            uint32_t v[2][32];
            const bool two = c0 + 8 <= my_c1, one = c0 + 4 <= my_c1;
            if (one) {
              tmem_ld<32>(t_lane + (uint32_t)c0 * 8, v[0]);
              if (two) tmem_ld<32>(t_lane + (uint32_t)(c0 + 4) * 8, v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h == 1 && !two) break;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  st_shared_v4(a_chunk(c0 + h * 4 + u),
                               pack_relu_bf16(__uint_as_float(v[h][u * 8 + 0]), __uint_as_float(v[h][u * 8 + 1])),
                               pack_relu_bf16(__uint_as_float(v[h][u * 8 + 2]), __uint_as_float(v[h][u * 8 + 3])),
                               pack_relu_bf16(__uint_as_float(v[h][u * 8 + 4]), __uint_as_float(v[h][u * 8 + 5])),
                               pack_relu_bf16(__uint_as_float(v[h][u * 8 + 6]), __uint_as_float(v[h][u * 8 + 7])));
              }
            }
            // tail: up to three single chunks
            const int done = two ? 8 : (one ? 4 : 0);
            for (int c = c0 + done; c < my_c1 && c < c0 + 8; ++c) {
              uint32_t w[8];
              tmem_ld<8>(t_lane + (uint32_t)c * 8, w);
              tmem_ld_wait();
              st_shared_v4(a_chunk(c), pack_relu_bf16(__uint_as_float(w[0]), __uint_as_float(w[1])),
                           pack_relu_bf16(__uint_as_float(w[2]), __uint_as_float(w[3])),
                           pack_relu_bf16(__uint_as_float(w[4]), __uint_as_float(w[5])),
                           pack_relu_bf16(__uint_as_float(w[6]), __uint_as_float(w[7])));
            }
          }
          TLW(2 + l * 4);
          publish_a();
          TLW(3 + l * 4);
          if (prm.sampling_propagation) {
#pragma unroll
            for (int c = 0; c < OW / 8; ++c)
              if ((c < L - 1 ? c : L - 1) == l) make_noise(c, t);
          }
        }
        // ---- Gaussian heads + state update + next input + partial minima ---------------------------
        wait_accumulator();
        TLW(40);
        if (prm.sampling_propagation) state_pass(std::false_type{}, std::true_type{}, t + 1);
        else state_pass(std::false_type{}, std::false_type{}, t + 1);
        TLW(41);
        if (t + 1 < H) publish_a();
        else named_bar_sync<kEpiThreads>(2);
        TLW(42);
        if (cgp == 0) {
          float next_dist, next_cost;
          combine(next_dist, next_cost);
          RowScore rs;
          rs.cum = __uint_as_float(my_rs[0]);
          rs.costsum = __uint_as_float(my_rs[128]);
          rs.cmask = (uint64_t)my_rs[256] | ((uint64_t)my_rs[384] << 32);
          rs.dist = __uint_as_float(my_rs[512]);
          rs.cost = __uint_as_float(my_rs[640]);
          rs.done = my_rs[768] != 0u;
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
          my_rs[0] = __float_as_uint(rs.cum);
          my_rs[128] = __float_as_uint(rs.costsum);
          my_rs[256] = (uint32_t)rs.cmask;
          my_rs[384] = (uint32_t)(rs.cmask >> 32);
          my_rs[512] = __float_as_uint(next_dist);
          my_rs[640] = __float_as_uint(next_cost);
          my_rs[768] = rs.done ? 1u : 0u;
        }
        TLW(43);
      }
      if (cgp == 0 && row_ok && prm.row_return != nullptr) {
        prm.row_return[id.out] = __uint_as_float(my_rs[0]);
        prm.row_costmask[id.out] = (uint64_t)my_rs[256] | ((uint64_t)my_rs[384] << 32);
        prm.row_costsum[id.out] = __uint_as_float(my_rs[128]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- host side ------------------------------------------------------------------------------------
static size_t wide_smem_bytes(int L, int U, int nparts);

// widths that are not a multiple of 16 run zero-padded to the next one (rollout_tc_wide_units)
int rollout_tc_wide_units(int U) { return (U + 15) / 16 * 16; }

// shape check used when the weight images are packed (no scorer yet: one constraint slot assumed)
bool rollout_tc_wide_supported(int O, int A, int L, int U, int H) {
  return U > 128 && rollout_tc_wide_units(U) <= 448 - 2 && O >= 1 && O <= 60 && O + A + 2 <= 64 && A <= 4 &&
         L >= 1 && L <= 6 && H >= 1 && H <= 64 && wide_smem_bytes(L, rollout_tc_wide_units(U), 1) <= 227 * 1024;
}

// the A tile, the two-stage weight ring and the per-constraint exchange buffers must share 227 KB
bool rollout_tc_wide_fits(int L, int U, int n_constraints) {
  return wide_smem_bytes(L, rollout_tc_wide_units(U), 1 + n_constraints) <= 227 * 1024;
}

int64_t rollout_tc_wide_member_bytes(int L, int U) {
  const WideShape ws = wide_shape(U, L);
  return (int64_t)(1 + (L - 1) * ws.KA) * ws.hid_tile_bytes + (int64_t)ws.KA * ws.head_tile_bytes;
}

static size_t wide_smem_bytes(int L, int U, int nparts) {
  const WideShape ws = wide_shape(U, L);
  size_t b = (size_t)ws.KA * kAtomBytes;
  b += (size_t)kStages * wide_stage_bytes(ws);
  b += 128 * sizeof(float) + kParts * 64 * sizeof(float);
  b += (size_t)kQ * nparts * 128 * sizeof(float);
  b += (size_t)8 * 128 * sizeof(uint32_t);                          // per-row running objective
  b += 6 * sizeof(uint64_t) + 2 * sizeof(uint32_t) + sizeof(TileInfoW);
  return b + 1024;                                                 // alignment slack
}

cudaError_t launch_rollout_tc_wide(const RolloutParams& prm, int n_tiles, cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  const size_t smem = wide_smem_bytes(prm.L, prm.U, 1 + prm.scorer.n_constraints);
  // set on every launch: the attribute is per device and per function, and launches happen only at
  // graph capture or in the non-graph entry points, never on the replayed hot path
  {
    cudaError_t e = cudaFuncSetAttribute(rollout_tc_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_tiles);
  cfg.blockDim = dim3(kWideThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = prm.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rollout_tc_wide_kernel, prm);
}

}  // namespace simba
