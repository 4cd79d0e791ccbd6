// bf16 tcgen05 fused rollout + scoring kernel (sm_100a) — the throughput path of the planner.
//
// Same contract as rollout_f32.cu (one launch per CEM iteration replaces cem_mpc.py:49-55,
// transition_model.py:64-87, mlp_ensemble.py:122-132,189-193, safety_gym.py:110-166 and the per-row
// part of mpc_policy.py:30-37 / safe_cem_mpc.py:82-93), with the per-layer GEMMs on the 5th-gen
// tensor cores:
//
//   * one CTA = one ensemble member x NTILES row tiles of 128 rollouts; the member's whole bf16
//     weight set (144 KB for 4x128) is staged ONCE in shared memory by the TMA engine
//     (cp.async.bulk, pre-swizzled UMMA images) and reused for all H steps x (L+1) layers;
//   * activations are the A operand: each epilogue thread owns one rollout row, reads its fp32
//     accumulator row from TMEM (tcgen05.ld 32x32b), applies bias + ReLU, packs bf16 and stores the
//     row straight into the 128B-swizzled K-major A tile of the next layer;
//   * one elected thread issues tcgen05.mma (M=128, N=128, K=16 per instruction, fp32 accumulate in
//     TMEM) and tcgen05.commit; two mbarriers per tile (A ready / accumulator ready) are the only
//     synchronisation, so with NTILES=2 the MMAs of one tile overlap the epilogue of the other;
//   * the Gaussian head epilogue does softplus, sqrt, Philox4x32-10 + Box-Muller, the residual
//     state update, and the lidar reward / hazard cost on registers; states never leave the SM.
#include <cuda_bf16.h>

#include "common.cuh"
#include "rollout_params.cuh"

namespace simba {

namespace {

constexpr int kU = 128;          // hidden width this kernel covers
constexpr int kMaxO = 60;        // observation dims held in registers
constexpr int kAtomBytes = 128 * 128;   // one 64-wide K atom of a 128-row tile (bf16, SW128)

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must fault the kernel, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, single CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i gets row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// relu(a), relu(b) -> packed bf16x2 (a in the low half)
__device__ __forceinline__ uint32_t pack_relu_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (dense 128 B rows)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// softplus on the MUFU path, accurate for very negative inputs (var head biases of trained models)
__device__ __forceinline__ float softplus_fast(float x) {
  const float e = __expf(fminf(x, 30.0f));
  const float small = e * (1.0f - e * (0.5f - e * 0.33333334f));
  const float big = __logf(1.0f + e);
  const float sp = e < 0.03f ? small : big;
  return x > 15.0f ? x : sp;
}

struct TileInfo {
  int32_t member, k0, count, valid;
};

}  // namespace

// kPG1: compile-time PointGoal1 layout (O=60, A=2, goal_lidar [3,19), one constrained lidar
// [22,38)) so that the lidar reductions index the register-resident state statically. The generic
// instantiation handles any O <= 60, O + A <= 64 with predicated (slower) scoring.
template <int NTILES, bool kPG1>
__global__ void __launch_bounds__(NTILES * 128 + 32, 1) rollout_tc_kernel(const RolloutParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const RowGeom& g = prm.g;
  const int L = prm.L;
  const int H = g.H;
  const int O = kPG1 ? 60 : g.O;
  const int A = kPG1 ? 2 : g.A;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kCtrlWarp = NTILES * 4;
  const bool is_ctrl = warp == kCtrlWarp;

  // ---- shared memory carve-up (base is 1024-aligned: required by SWIZZLE_128B) -------------------
  const uint32_t w_bytes = kAtomBytes + (uint32_t)L * 2 * kAtomBytes;     // layer 0: 1 atom; others: 2
  uint8_t* w_smem = smem_raw;
  uint8_t* a_smem = w_smem + w_bytes;                                      // [NTILES][2 atoms]
  float* bias_smem = reinterpret_cast<float*>(a_smem + NTILES * 2 * kAtomBytes);   // [(L+1)][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_smem + (L + 1) * 128);
  // bars[0] = weights landed; bars[1 + j] = A ready (tile j); bars[1 + NTILES + j] = accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * NTILES);
  TileInfo* tinfo = reinterpret_cast<TileInfo*>(tmem_slot + 2);

  const uint32_t bar_w = smem_u32(&bars[0]);
  uint32_t bar_a[NTILES], bar_acc[NTILES];
#pragma unroll
  for (int j = 0; j < NTILES; ++j) {
    bar_a[j] = smem_u32(&bars[1 + j]);
    bar_acc[j] = smem_u32(&bars[1 + NTILES + j]);
  }

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
#pragma unroll
    for (int j = 0; j < NTILES; ++j) {
      mbar_init(bar_a[j], 4);        // one arrival per epilogue warp of the tile
      mbar_init(bar_acc[j], 1);      // tcgen05.commit
    }
    fence_barrier_init();
  }
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
  if (is_ctrl) tmem_alloc(smem_u32(tmem_slot), NTILES * 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  bool any_valid = false;
  int member = 0;
#pragma unroll
  for (int j = 0; j < NTILES; ++j)
    if (tinfo[j].valid) { any_valid = true; member = tinfo[j].member; }

  if (any_valid) {
    if (is_ctrl) {
      // =========================== control warp: TMA + MMA issue ===============================
      if (lane == 0) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(prm.w_bf16) +
                              (size_t)member * prm.w_bf16_member_bytes;
        mbar_expect_tx(bar_w, w_bytes);
        uint32_t off = 0;
        for (int l = 0; l <= L; ++l) {
          const uint32_t nb = (l == 0) ? kAtomBytes : 2 * kAtomBytes;
          bulk_g2s(smem_u32(w_smem + off), wsrc + off, nb, bar_w);
          off += nb;
        }
        mbar_wait(bar_w, 0);
        uint32_t ph[NTILES];
#pragma unroll
        for (int j = 0; j < NTILES; ++j) ph[j] = 0;
        for (int t = 0; t < H; ++t) {
          uint32_t woff = 0;
          for (int l = 0; l <= L; ++l) {
            const int ksteps = (l == 0) ? 4 : 8;             // K = 64 or 128, UMMA_K = 16
#pragma unroll
            for (int j = 0; j < NTILES; ++j) {
              if (!tinfo[j].valid) continue;
              mbar_wait(bar_a[j], ph[j]);
              ph[j] ^= 1;
              tc_fence_after();
              const uint32_t a_base = smem_u32(a_smem + j * 2 * kAtomBytes);
              const uint32_t b_base = smem_u32(w_smem + woff);
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t koff = (uint32_t)(k >> 2) * kAtomBytes + (uint32_t)(k & 3) * 32;
                umma_bf16(tmem_base + j * 128, umma_desc_sw128(a_base + koff),
                          umma_desc_sw128(b_base + koff), kIdesc, k > 0 ? 1u : 0u);
              }
              umma_commit(bar_acc[j]);
            }
            woff += (l == 0) ? kAtomBytes : 2 * kAtomBytes;
          }
        }
      }
      __syncwarp();
    } else {
      // =========================== epilogue warps: one thread = one rollout row ==================
      const int j = warp >> 2;                 // tile of this warp
      const int r = threadIdx.x - j * 128;     // row in tile == TMEM lane
      const TileInfo ti = tinfo[j];
      // biases of this member -> shared (all epilogue threads of the CTA cooperate)
      {
        const float* bsrc = prm.bias_tc + (size_t)member * (L + 1) * 128;
        for (int i = threadIdx.x; i < (L + 1) * 128; i += NTILES * 128) bias_smem[i] = bsrc[i];
        asm volatile("bar.sync 1, %0;" ::"n"(NTILES * 128));
      }
      if (ti.valid) {
        const bool row_ok = r < ti.count;
        const RowId id = decode_row(g, ti.member, ti.k0 + (row_ok ? r : 0));
        const uint64_t seed = prm.seed_ptr ? *prm.seed_ptr : prm.seed;
        const uint32_t a_tile = smem_u32(a_smem + j * 2 * kAtomBytes);
        const uint32_t a_row = a_tile + (uint32_t)r * 128;
        const uint32_t swz = (uint32_t)(r & 7);
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)j * 128;
        const float* act_ptr = prm.actions + ((int64_t)id.s * g.N + id.i_global) * prm.action_stride;
        const float* sc_a = prm.tc_scale_a;     // x_scaled = fma(x, sc_a, sc_b)
        const float* sc_b = prm.tc_scale_b;
        const bool done_first = objective_done_first(prm.objective);
        const simba_scorer_t& sc = prm.scorer;

        float s[kMaxO];
        {
          const float* sp = prm.states + (prm.state_per_row ? id.r_global : (int64_t)id.s) * prm.state_stride;
#pragma unroll
          for (int o = 0; o < kMaxO; ++o) s[o] = (row_ok && o < O) ? sp[o] : 0.0f;
        }
        // static-index state access for the scorer
        auto closest = [&](int begin, int end, float D) {
          float best = INFINITY;
#pragma unroll
          for (int o = 0; o < kMaxO; ++o) {
            if (kPG1 ? (o >= begin && o < end) : true) {
              float v = __fsub_rn(D, __fmul_rn(D, __fsub_rn(1.0f, s[o])));
              v = fminf(fmaxf(v, 0.0f), D);
              if (kPG1) best = fminf(best, v);
              else best = (o >= begin && o < end) ? fminf(best, v) : best;
            }
          }
          return best;
        };
        auto goal_dist = [&]() {
          if (kPG1) return closest(3, 19, sc.lidar_max_dist);
          if (sc.goal_dist_index >= 0) {
            float v = 0.0f;
#pragma unroll
            for (int o = 0; o < kMaxO; ++o) v = (o == sc.goal_dist_index) ? s[o] : v;
            return fmaxf(v, 0.0f);
          }
          return closest(sc.goal_begin, sc.goal_end, sc.lidar_max_dist);
        };
        auto cost_now = [&]() {
          if (kPG1) return closest(22, 38, sc.lidar_max_dist) <= sc.con_size[0] ? 1.0f : 0.0f;
          float c = 0.0f;
          for (int q = 0; q < sc.n_constraints; ++q)
            c += closest(sc.con_begin[q], sc.con_end[q], sc.lidar_max_dist) <= sc.con_size[q] ? 1.0f : 0.0f;
          return sc.constrain_indicator ? (c > 0.0f ? 1.0f : 0.0f) : c;
        };
        RowScore rs;
        rs.cum = 0.0f; rs.costsum = 0.0f; rs.cmask = 0ull; rs.done = false;
        rs.dist = goal_dist();
        rs.cost = cost_now();

        uint32_t ph = 0;
        for (int t = 0; t < H; ++t) {
          // ---- layer-0 A operand: bf16(scale([s_t, a_t])), 64 K-elements = 8 swizzled 16B chunks --
          {
            float x[64];
#pragma unroll
            for (int k = 0; k < 64; ++k) {
              float v = 0.0f;
              if (k < kMaxO) v = s[k];
              if (!kPG1) {
                if (k >= O && k < O + A) v = row_ok ? act_ptr[t * A + (k - O)] : 0.0f;
              } else if (k >= 60 && k < 62) {
                v = row_ok ? act_ptr[t * 2 + (k - 60)] : 0.0f;
              }
              x[k] = fmaf(v, sc_a[k], sc_b[k]);      // padded k: sc_a = sc_b = 0
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
              st_shared_v4(a_row + (((uint32_t)c ^ swz) << 4), pack_bf16(x[8 * c], x[8 * c + 1]),
                           pack_bf16(x[8 * c + 2], x[8 * c + 3]), pack_bf16(x[8 * c + 4], x[8 * c + 5]),
                           pack_bf16(x[8 * c + 6], x[8 * c + 7]));
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_a[j]);

          // ---- hidden layers: TMEM -> +bias, ReLU, bf16 -> next A operand -----------------------
          for (int l = 0; l < L; ++l) {
            mbar_wait(bar_acc[j], ph);
            ph ^= 1;
            tc_fence_after();
            const float* bl = bias_smem + l * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t v[32];
              tmem_ld32(t_lane + c * 32, v);
              tmem_ld_wait();
              const uint32_t atom = a_row + (uint32_t)(c >> 1) * kAtomBytes;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 b0 = *reinterpret_cast<const float4*>(bl + c * 32 + q * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(bl + c * 32 + q * 8 + 4);
                const uint32_t p0 = pack_relu_bf16(__uint_as_float(v[q * 8 + 0]) + b0.x,
                                                   __uint_as_float(v[q * 8 + 1]) + b0.y);
                const uint32_t p1 = pack_relu_bf16(__uint_as_float(v[q * 8 + 2]) + b0.z,
                                                   __uint_as_float(v[q * 8 + 3]) + b0.w);
                const uint32_t p2 = pack_relu_bf16(__uint_as_float(v[q * 8 + 4]) + b1.x,
                                                   __uint_as_float(v[q * 8 + 5]) + b1.y);
                const uint32_t p3 = pack_relu_bf16(__uint_as_float(v[q * 8 + 6]) + b1.z,
                                                   __uint_as_float(v[q * 8 + 7]) + b1.w);
                const uint32_t chunk = (uint32_t)((c & 1) * 4 + q);
                st_shared_v4(atom + ((chunk ^ swz) << 4), p0, p1, p2, p3);
              }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a[j]);
          }

          // ---- Gaussian heads: mu cols [0, 64), raw var cols [64, 128) ----------------------------
          mbar_wait(bar_acc[j], ph);
          ph ^= 1;
          tc_fence_after();
          {
            const float* bh = bias_smem + L * 128;
#pragma unroll
            for (int hhalf = 0; hhalf < 2; ++hhalf) {
              uint32_t vm[32], vv[32];
              tmem_ld32(t_lane + hhalf * 32, vm);
              tmem_ld32(t_lane + 64 + hhalf * 32, vv);
              tmem_ld_wait();
#pragma unroll
              for (int jb = 0; jb < 8; ++jb) {
                const int o0 = hhalf * 32 + jb * 4;
                if (o0 >= kMaxO) continue;
                float e4[4] = {0.f, 0.f, 0.f, 0.f};
                if (prm.sampling_propagation) {
                  if (prm.eps != nullptr) {
                    const float* ep = prm.eps + (((int64_t)id.s * H + t) * ((int64_t)g.P * g.N) + id.r_global) * O;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                      if (o0 + q < O) e4[q] = ep[o0 + q];
                  } else {
                    const float4 z = philox_normals<true>(seed, kStreamNoise, (uint32_t)id.s,
                                                          (uint32_t)prm.iteration, (uint32_t)t,
                                                          (uint32_t)id.r_global, (uint32_t)(o0 >> 2));
                    e4[0] = z.x; e4[1] = z.y; e4[2] = z.z; e4[3] = z.w;
                  }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int o = o0 + q;
                  if (o >= kMaxO) continue;
                  const float mu = __uint_as_float(vm[jb * 4 + q]) + bh[o];
                  float d = mu;
                  if (prm.sampling_propagation) {
                    const float raw = __uint_as_float(vv[jb * 4 + q]) + bh[64 + o];
                    const float var = softplus_fast(raw) + 1e-4f;
                    d = fmaf(sqrtf(var), e4[q], mu);
                  }
                  if (kPG1 || o < O) s[o] += d;
                }
              }
            }
          }
          // the accumulator has been consumed; the next step's first MMA may overwrite it once this
          // thread's next bar_a arrival (after the A-operand write above) is observed.

          // ---- scoring of (s_t, s_{t+1}): safety_gym.py:110-166, per-row objective ------------------
          {
            const float next_dist = goal_dist();
            const float next_cost = cost_now();
            const bool goal = rs.dist <= sc.goal_threshold;
            const float rew = step_reward(sc, rs.dist, next_dist, goal);
            if (done_first) {
              rs.done = rs.done || goal;
              if (!rs.done && rs.cost > 0.0f) rs.cmask |= (1ull << t);
              rs.cum += rs.done ? 0.0f : rew;
            } else {
              rs.cum += rs.done ? 0.0f : rew;
              if (!rs.done && rs.cost > 0.0f) rs.cmask |= (1ull << t);
              rs.done = rs.done || goal;
            }
            rs.costsum += rs.cost;
            rs.dist = next_dist;
            rs.cost = next_cost;
          }
        }
        if (row_ok && prm.row_return != nullptr) {
          prm.row_return[id.out] = rs.cum;
          prm.row_costmask[id.out] = rs.cmask;
          prm.row_costsum[id.out] = rs.costsum;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (is_ctrl) tmem_dealloc(tmem_base, NTILES * 128);
}

// ---- host side ------------------------------------------------------------------------------------
bool rollout_tc_supported(int O, int A, int L, int U, int H) {
  return U == kU && O >= 1 && O <= kMaxO && O + A <= 64 && L >= 1 && L <= 6 && H >= 1 && H <= 64;
}

static size_t tc_smem_bytes(int L, int ntiles) {
  size_t b = (size_t)kAtomBytes + (size_t)L * 2 * kAtomBytes;      // weights
  b += (size_t)ntiles * 2 * kAtomBytes;                            // A operands
  b += (size_t)(L + 1) * 128 * sizeof(float);                      // biases
  b += (1 + 2 * ntiles) * sizeof(uint64_t) + 2 * sizeof(uint32_t) + ntiles * sizeof(TileInfo);
  return b + 1024;                                                 // alignment slack
}

template <int NTILES, bool kPG1>
static cudaError_t launch_variant(const RolloutParams& prm, int n_tiles, cudaStream_t stream) {
  const size_t smem = tc_smem_bytes(prm.L, NTILES);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(rollout_tc_kernel<NTILES, kPG1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int grid = (n_tiles + NTILES - 1) / NTILES;
  rollout_tc_kernel<NTILES, kPG1><<<grid, NTILES * 128 + 32, smem, stream>>>(prm);
  return cudaGetLastError();
}

cudaError_t launch_rollout_tc(const RolloutParams& prm, int n_tiles, cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  const simba_scorer_t& sc = prm.scorer;
  const bool pg1 = prm.g.O == 60 && prm.g.A == 2 && sc.goal_dist_index < 0 && sc.goal_begin == 3 &&
                   sc.goal_end == 19 && sc.n_constraints == 1 && sc.con_begin[0] == 22 &&
                   sc.con_end[0] == 38;
  const bool two = prm.tc_tiles_per_cta == 2;
  if (pg1) return two ? launch_variant<2, true>(prm, n_tiles, stream) : launch_variant<1, true>(prm, n_tiles, stream);
  return two ? launch_variant<2, false>(prm, n_tiles, stream) : launch_variant<1, false>(prm, n_tiles, stream);
}

}  // namespace simba
