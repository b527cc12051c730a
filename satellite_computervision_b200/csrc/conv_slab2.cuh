// CTA-pair form of the persistent weight-stationary slab kernel (conv_umma.cuh: conv_slab_kernel) for the Cout = 64
// layers of the 192 x 192 level: Conv2D 3x3 + BN + ReLU (+ MaxPool + skip BN), utils/model_tools.py:178-186, :281,
// :307-317.
//
// Two CTAs of a cluster (the two SMs of a TPC) issue ONE tcgen05.mma.cta_group::2 per K = 16 step: M = 256 = the
// 8 x 16 pixel tile of each CTA, N = 64.  Every CTA loads its own halo slabs (A: 128 rows) and HALF of the weights
// (B: output channels [32 r, 32 r + 32) of cluster rank r); the tensor cores exchange the B halves between the SMs.
// What this buys on these layers:
//   * the resident weight set per SM halves (Cin = 128: 147 KB -> 74 KB), so decoder_1/conv0 gets a slab ring that
//     covers more than one tile and four accumulators instead of two (its single-CTA form runs with two slabs = one
//     tile in flight: tensor pipe 72 % busy);
//   * the per-SM operand fetch of a K = 16 step drops from 4 KB + 2 KB to 4 KB + 1 KB (measured
//     tools/microbench/umma_rate_2cta.cu: 53.9 -> 49.8 cycles).
// The K loop is conv_slab_kernel's (chunk, tap, k), so a pixel's result is bit-identical to the single-CTA kernels.
//
// Protocol (all barriers exist at the same offsets in both CTAs; L = leader = cluster rank 0):
//   w_full     [L]     1 arrive.expect_tx by L's producer; both CTAs' weight TMA loads complete_tx on it
//   slab_full  [L]     per slot: 1 arrive.expect_tx (2 slabs' bytes) by L's producer; both CTAs' slab loads complete_tx on it
//   slab_empty [own]   per slot: one multicast tcgen05.commit from L's issuer (the UMMAs that read the slot are done)
//   acc_full   [own]   per accumulator: one multicast tcgen05.commit (the tile's UMMAs are done)
//   acc_empty  [L]     per accumulator: one arrive per epilogue warp of BOTH CTAs (4 local + 4 remote)
// Only L's issuer warps issue UMMAs; the follower's issuer warps idle.  Every wait is watchdogged (ptx.cuh).
#pragma once
#include "conv_umma.cuh"

namespace scv {

constexpr int kSlab2Issuers = 2;
__host__ __device__ constexpr int slab2_threads(int nacc) { return 32 * (1 + kSlab2Issuers) + 128 * nacc; }
constexpr int kSlab2BN = 64;   // output channels of the pair's accumulator
constexpr int kSlab2BH = 32;   // weight rows (output channels) resident per CTA

__host__ __device__ inline size_t slab2_weight_bytes(int cin) { return static_cast<size_t>(9) * cin * kSlab2BH * 2; }
__host__ __device__ inline size_t slab2_smem_bytes(int KC, int cin, int nslab, int epi, int nacc) {
  size_t s = 1024 + slab2_weight_bytes(cin) + static_cast<size_t>(nslab) * slab_stride_bytes(KC, 9) + slab_stage_bytes(epi, nacc);
  s += (1 + 2 * nslab + 2 * 4) * 8 + 16;
  s += kSlab2BN * 4;
  if (epi == EPI_POOL_SKIP) s += 2 * kSlab2BN * 4;
  return s + 64;
}

// ---- cluster / cta_group::2 PTX --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t slab2_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t slab2_mapa(uint32_t laddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void slab2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by the leader CTA only
__device__ __forceinline__ void umma_pair_bf16_lo_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi_a,
                                                      uint32_t desc_hi_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .b64 da, db;\n"
      ".reg .pred p;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi_a), "r"(desc_hi_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once all UMMAs issued so far have completed
__device__ __forceinline__ void umma_pair_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// TMA loads into THIS CTA's shared memory whose bytes are counted on a barrier of the pair given as a shared::cluster
// address (the leader's)
__device__ __forceinline__ void tma_pair_load_2d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_pair_load_4d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on a barrier of the pair given as a shared::cluster address (CTA-scope release: the data it orders is TMEM,
// fenced with tcgen05.fence::before_thread_sync -- no MEMBAR.ALL.GPU as with .release.cluster)
__device__ __forceinline__ void mbar_arrive_pair(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

// KC: channel chunk (32 or 64); EPI: EPI_STORE or EPI_POOL_SKIP; NACC accumulators == epilogue warpgroups.
template <int KC, int EPI, int NACC>
__global__ void __launch_bounds__(slab2_threads(NACC), 1)
    conv_slab2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPool,
                      const ConvParams p) {
  constexpr int BN = kSlab2BN, BH = kSlab2BH;
  constexpr int ROWB = KC * 2;
  constexpr int WT_BYTES = BH * KC * 2;  // this CTA's half of one (chunk, tap) weight tile
  constexpr int SW = 10, SH = 18;
  constexpr int SLAB_BYTES = SW * SH * ROWB;
  constexpr int SLAB_STRIDE = (SLAB_BYTES + 1023) & ~1023;
  constexpr uint32_t IDESC = umma_idesc_bf16(256, BN);
  static_assert(NACC * BN <= 512 && (NACC == 2 || NACC == 4), "TMEM columns");
  static_assert(EPI == EPI_STORE || EPI == EPI_POOL_SKIP, "epilogues with bf16 outputs");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int chunks = p.Cin / KC;
  const int nwt = 9 * chunks;
  const int nslab = p.nslab;
  uint8_t* w_smem = base;
  uint8_t* slabs = base + static_cast<size_t>(nwt) * WT_BYTES;  // WT_BYTES is a multiple of 1024 for KC >= 32... (32*32*2 = 2048)
  uint8_t* staging = slabs + static_cast<size_t>(nslab) * SLAB_STRIDE;
  constexpr int STAGE_TOTAL = 4 * NACC * slab_stage_warp_bytes(EPI);
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STAGE_TOTAL);
  uint64_t* w_full = bars;
  uint64_t* slab_full = bars + 1;
  uint64_t* slab_empty = slab_full + nslab;
  uint64_t* acc_full = slab_empty + nslab;
  uint64_t* acc_empty = acc_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
  float* s_extra = s_bias + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = slab2_ctarank();
  const bool leader = rank == 0;
  // tile m of this CTA in round i: 2 * (pair + i * npairs) + rank; num_m_tiles is even, so both CTAs of a pair always
  // have the same number of tiles
  const int m_first = static_cast<int>(blockIdx.x);
  const int m_stride = static_cast<int>(gridDim.x);
  const int tiles_xy = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if constexpr (EPI == EPI_POOL_SKIP) tma_prefetch_desc(&tmPool);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int s = 0; s < nslab; ++s) {
        mbar_init(&slab_full[s], 1);
        mbar_init(&slab_empty[s], 1);
      }
      for (int a = 0; a < NACC; ++a) {
        mbar_init(&acc_full[a], 1);
        mbar_init(&acc_empty[a], 8);  // one arrive per epilogue warp of the accumulator's group, both CTAs
      }
      *abort_flag = 0;
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, NACC * BN);
    tmem_relinquish_pair();
  }
  constexpr int kFirstEpiWarp = 1 + kSlab2Issuers;
  if (warp >= kFirstEpiWarp)
    load_epilogue_consts<BN, EPI>(p, threadIdx.x - 32 * kFirstEpiWarp, 128 * NACC, 0, s_bias, s_extra);
  tc_fence_before();
  slab2_cluster_sync();  // both CTAs' barriers and TMEM exist before any remote arrive / paired UMMA (also a CTA barrier)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t w_full_L = slab2_mapa(smem_u32(w_full), 0);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own slabs + own half of the weights =====================
    if (elect_one()) {
      if (leader) mbar_arrive_expect_tx(w_full, 2u * static_cast<uint32_t>(nwt) * WT_BYTES);
      for (int ch = 0; ch < chunks; ++ch)
        for (int tap = 0; tap < 9; ++tap)  // smem order [chunk][tap]; global K index = tap*Cin + ch*KC
          tma_pair_load_2d(w_smem + static_cast<size_t>(ch * 9 + tap) * WT_BYTES, &tmB, w_full_L, tap * p.Cin + ch * KC,
                           static_cast<int>(rank) * BH);
    }
    __syncwarp();
    uint32_t it = 0;
    bool run = true;
    for (int m = m_first; run && m < p.num_m_tiles; m += m_stride) {
      const int n = m / tiles_xy;
      const int rem = m - n * tiles_xy;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      for (int ch = 0; ch < chunks; ++ch, ++it) {
        const uint32_t s = it % nslab;
        const bool ok = mbar_wait(&slab_empty[s], ((it / nslab) & 1) ^ 1, abort_flag, p.watchdog_ns);
        if (!__all_sync(0xffffffffu, ok)) {
          run = false;
          break;
        }
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&slab_full[s], 2u * SLAB_BYTES);
          tma_pair_load_4d(slabs + static_cast<size_t>(s) * SLAB_STRIDE, &tmA, slab2_mapa(smem_u32(&slab_full[s]), 0), ch * KC,
                           tx * 8 - 1, ty * 16 - 1, n + p.n_in_off);
        }
        __syncwarp();
      }
    }
  } else if (warp < kFirstEpiWarp) {
    // ===================== MMA issuers (leader only): issuer j takes the pair's tiles t == j (mod 2) =====================
    const int nissue = p.n_issuers;
    bool run = leader && (warp - 1) < nissue && __all_sync(0xffffffffu, mbar_wait(w_full, 0, abort_flag, p.watchdog_ns));
    uint32_t t = warp - 1;
    const uint32_t w_addr = smem_u32(w_smem);
    for (int m = m_first + static_cast<int>(t) * m_stride; run && m < p.num_m_tiles; m += m_stride * nissue, t += nissue) {
      uint32_t it = t * chunks;
      const uint32_t a = t % NACC;
      const bool ok = mbar_wait(&acc_empty[a], ((t / NACC) & 1) ^ 1, abort_flag, p.watchdog_ns);
      if (!__all_sync(0xffffffffu, ok)) break;
      tc_fence_after();
      const uint32_t tacc = tmem_base + a * BN;
      for (int ch = 0; ch < chunks; ++ch, ++it) {
        const uint32_t s = it % nslab;
        const bool ok2 = mbar_wait(&slab_full[s], (it / nslab) & 1, abort_flag, p.watchdog_ns);
        if (!__all_sync(0xffffffffu, ok2)) {
          run = false;
          break;
        }
        tc_fence_after();
        const uint32_t slab_addr = smem_u32(slabs + static_cast<size_t>(s) * SLAB_STRIDE);
        const uint64_t da0 = umma_smem_desc_sbo(slab_addr, ROWB, SW * ROWB);
        const uint64_t db0 = umma_smem_desc(w_addr + static_cast<uint32_t>(ch * 9) * WT_BYTES, ROWB);
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              const uint32_t da = desc_lo(da0) + static_cast<uint32_t>(((dy * SW + dx) * ROWB + k * 32) >> 4);
              const uint32_t db = desc_lo(db0) + static_cast<uint32_t>((tap * WT_BYTES + k * 32) >> 4);
              umma_pair_bf16_lo_acc(tacc, da, db, desc_hi(da0), desc_hi(db0), IDESC, (tap | k) != 0 ? 1u : (ch != 0 ? 1u : 0u));
            }
          }
          umma_pair_commit(&slab_empty[s]);
        }
        __syncwarp();
      }
      if (run && elect_one()) umma_pair_commit(&acc_full[a]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue (both CTAs): warpgroup g handles this CTA's tiles t == g (mod NACC) =====================
    const int g = (warp - kFirstEpiWarp) >> 2;
    const int q = warp & 3;
    uint32_t t = g;
    const uint32_t acc_empty_L = slab2_mapa(smem_u32(&acc_empty[g]), 0);
    for (int m = m_first + g * m_stride; m < p.num_m_tiles; m += m_stride * NACC, t += NACC) {
      const int n = m / tiles_xy;
      const int rem = m - n * tiles_xy;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const bool ready = mbar_wait(&acc_full[g], (t / NACC) & 1, abort_flag, p.watchdog_ns);
      if (!__all_sync(0xffffffffu, ready)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + g * BN + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_slab<BN, EPI>(p, &tmOut, &tmPool, taddr, lane, q, tx * 8, ty * 16, n, 0, s_bias, s_extra,
                             staging + static_cast<size_t>(warp - kFirstEpiWarp) * slab_stage_warp_bytes(EPI), [&] {
                               tc_fence_before();  // every lane's tcgen05.ld has completed (wait::ld is warp-wide)
                               __syncwarp();
                               if (lane == 0) mbar_arrive_pair(acc_empty_L);
                             });
    }
    if (lane == 0) bulk_wait_read<0>();  // staging must stay valid until the last stores have read it
  }

  tc_fence_before();
  slab2_cluster_sync();  // neither CTA frees TMEM / exits while the pair's UMMAs, commits or remote arrives are in flight
  if (warp == 1) {
    tmem_dealloc_pair(tmem_base, NACC * BN);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

cudaError_t conv_slab2_launch(const ConvLaunch& L, cudaStream_t stream);
cudaError_t conv_slab2_init_attributes();
int conv_slab2_max_pairs(size_t smem, int nacc);

}  // namespace scv
