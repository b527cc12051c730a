// Halo-slab 3x3 convolution with STREAMED weights, for the mid-resolution layers whose weights do not fit
// shared memory (Cout = 128 at 96x96: 9*Cin*128*2 B = 147..295 KB) -- utils/model_tools.py:178-186 (+ :281-286
// max-pool, :307-309 skip BN) as in the other conv kernels.
//
// The one-tile-per-CTA / persistent-tile kernels fetch a 16 KB pixel tile AND a 16 KB weight tile from L2 for
// every (tap, 64-channel chunk): 32 KB per four 128x128x16 UMMAs, ~120 B/clk/SM at full tensor rate, which the
// L2 -> SM path cannot deliver to 148 SMs (measured: decoder_2/conv0 sits at ~15 TB/s of L2 traffic and
// 0.97 PFLOP/s whichever of the two runs it).  Here, per pass of TWO 8x16-pixel M tiles and per channel chunk:
//   * each tile's (8+2)x(16+2) halo slab is loaded ONCE and the nine taps are nine shifted UMMA descriptors
//     into it (as in conv_slab_kernel): pixel traffic 144 KB -> 23 KB per tile and chunk;
//   * the nine 128-channel weight tiles stream through a small ring and each is used by BOTH tiles:
//     weight traffic 147 KB -> 74 KB per tile and chunk.
// ~3x less L2 traffic per FLOP.  Four TMEM accumulators: the pass in flight uses two, the epilogue warpgroups
// drain the other two.  Per-pixel summation order is (chunk, tap, k) -- identical to the other kernels.
#pragma once
#include "conv_umma.cuh"

namespace scv {

// warp 0: TMA; warps 1..kSlabwIssuers: MMA issuers taking turns tap by tap (token passing as in conv_rows_kernel:
// one issuer's barrier waits / descriptor arithmetic / commits overlap the other's UMMAs, issue order stays
// sequential); then 4 epilogue warpgroups.
constexpr int kSlabwIssuers = 2;
constexpr int kSlabwFirstEpiWarp = 1 + kSlabwIssuers;
constexpr int kSlabwThreads = 32 * kSlabwFirstEpiWarp + 128 * 4;
constexpr int kSlabwSlabs = 4;               // slab ring (two passes x two tiles ... of one chunk each)

__host__ __device__ inline size_t slabw_smem_bytes(int KC, int BN, int nb, int epi) {
  size_t s = 1024 + static_cast<size_t>(kSlabwSlabs) * slab_stride_bytes(KC, 9) + static_cast<size_t>(nb) * BN * KC * 2 +
             slab_stage_bytes(epi, 4);
  s += (2 * kSlabwSlabs + 2 * nb + 8 + kSlabwIssuers) * 8 + 16;
  s += BN * 4;
  if (epi == EPI_POOL_SKIP) s += 2 * BN * 4;
  return s + 64;
}

template <int KC, int BN, int EPI>
__global__ void __launch_bounds__(kSlabwThreads, 1)
    conv_slabw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPool,
                      const ConvParams p) {
  constexpr int ROWB = KC * 2;
  constexpr int WT_BYTES = BN * KC * 2;
  constexpr int SW = 10, SH = 18;
  constexpr int SLAB_BYTES = SW * SH * ROWB;
  constexpr int SLAB_STRIDE = (SLAB_BYTES + 1023) & ~1023;
  constexpr int NS = kSlabwSlabs;
  constexpr int NACC = 4;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, BN);
  static_assert(NACC * BN <= 512, "four accumulators must fit TMEM");
  static_assert(EPI == EPI_STORE || EPI == EPI_POOL_SKIP, "bf16 tensor outputs only");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int chunks = p.Cin / KC;
  const int nb = p.nstage;  // weight ring depth
  uint8_t* slabs = base;
  uint8_t* wring = slabs + static_cast<size_t>(NS) * SLAB_STRIDE;
  uint8_t* staging = wring + static_cast<size_t>(nb) * WT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + slab_stage_bytes(EPI, NACC));
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = slab_full + NS;
  uint64_t* b_full = slab_empty + NS;
  uint64_t* b_empty = b_full + nb;
  uint64_t* acc_full = b_empty + nb;
  uint64_t* acc_empty = acc_full + NACC;
  uint64_t* turn = acc_empty + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + kSlabwIssuers);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
  float* s_extra = s_bias + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles_n;
  const int m_first = blockIdx.x / p.n_tiles_n;
  const int m_stride = gridDim.x / p.n_tiles_n;
  const int nb0 = n_tile * BN;
  const int tiles_xy = p.tiles_x * p.tiles_y;
  // local tiles of this CTA: m = m_first + t * m_stride, t = 0 .. ntl-1; passes take tiles (2k, 2k+1)
  const int ntl = m_first < p.num_m_tiles ? (p.num_m_tiles - m_first + m_stride - 1) / m_stride : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if constexpr (EPI == EPI_POOL_SKIP) tma_prefetch_desc(&tmPool);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) {
        mbar_init(&slab_full[s], 1);
        mbar_init(&slab_empty[s], 1);
      }
      for (int s = 0; s < nb; ++s) {
        mbar_init(&b_full[s], 1);
        mbar_init(&b_empty[s], 1);
      }
      for (int a = 0; a < NACC; ++a) {
        mbar_init(&acc_full[a], 1);
        mbar_init(&acc_empty[a], 128);
      }
      for (int i = 0; i < kSlabwIssuers; ++i) mbar_init(&turn[i], 1);
      *abort_flag = 0;
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, NACC * BN);
    tmem_relinquish();
  }
  if (warp >= kSlabwFirstEpiWarp)
    load_epilogue_consts<BN, EPI>(p, threadIdx.x - 32 * kSlabwFirstEpiWarp, 128 * NACC, nb0, s_bias, s_extra);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();  // programmatic dependent launch: see launch_pdl()
  grid_dep_wait();    // nothing above reads an activation; everything below may

  if (warp == 0) {
    // ===================== TMA producer: per pass and chunk -- two slabs, then the nine weight tiles ==========
    uint32_t ss = 0, sph = 1, bs = 0, bph = 1;
    bool run = true;
    for (int t = 0; run && t < ntl; t += 2) {
      const int nt = (ntl - t) < 2 ? (ntl - t) : 2;
      for (int ch = 0; run && ch < chunks; ++ch) {
        for (int i = 0; i < nt; ++i) {
          const int m = m_first + (t + i) * m_stride;
          const int n = m / tiles_xy;
          const int rem = m - n * tiles_xy;
          const int ty = rem / p.tiles_x;
          const int tx = rem - ty * p.tiles_x;
          const bool ok = mbar_wait(&slab_empty[ss], sph, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok)) {
            run = false;
            break;
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(&slab_full[ss], SLAB_BYTES);
            tma_load_4d(slabs + static_cast<size_t>(ss) * SLAB_STRIDE, &tmA, &slab_full[ss], ch * KC, tx * 8 - 1,
                        ty * 16 - 1, n + p.n_in_off);
          }
          __syncwarp();
          if (++ss == NS) ss = 0, sph ^= 1;
        }
        for (int tap = 0; run && tap < 9; ++tap) {
          const bool ok = mbar_wait(&b_empty[bs], bph, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok)) {
            run = false;
            break;
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[bs], WT_BYTES);
            tma_load_2d(wring + static_cast<size_t>(bs) * WT_BYTES, &tmB, &b_full[bs], tap * p.Cin + ch * KC, nb0);
          }
          __syncwarp();
          if (++bs == static_cast<uint32_t>(nb)) bs = 0, bph ^= 1;
        }
      }
    }
  } else if (warp < kSlabwFirstEpiWarp) {
    // ===================== MMA issuers: issuer k takes the CTA's (pass, chunk, tap) steps k, k+ni, ... ==========
    const int ni = p.n_issuers;  // 1 or 2
    const int me = warp - 1;
    uint32_t ss = 0, sph = 0, bs = 0, bph = 0, nth = 0;
    int turn_of = 0;
    bool run = me < ni;
    for (int t = 0; run && t < ntl; t += 2) {
      const int nt = (ntl - t) < 2 ? (ntl - t) : 2;
      uint32_t tacc[2];
      for (int i = 0; i < nt; ++i) tacc[i] = tmem_base + ((t + i) % NACC) * BN;
      for (int ch = 0; run && ch < chunks; ++ch) {
        // this chunk's slabs: every issuer tracks the ring; it waits for them at its first step of the chunk
        uint32_t sslot[2], sphase[2];
        for (int i = 0; i < nt; ++i) {
          sslot[i] = ss;
          sphase[i] = sph;
          if (++ss == NS) ss = 0, sph ^= 1;
        }
        bool have_slabs = false;
        uint64_t da0[2] = {0, 0};
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, ++turn_of) {
          if (turn_of == ni) turn_of = 0;
          const uint32_t bsc = bs, bphc = bph;
          if (++bs == static_cast<uint32_t>(nb)) bs = 0, bph ^= 1;
          if (turn_of != me) continue;
          // ---- everything that does not need the token
          if (ch == 0 && tap == 0) {  // first step of the pass: the tiles' accumulators must have been drained
            for (int i = 0; i < nt; ++i) {
              const uint32_t tt = t + i;
              const bool ok = mbar_wait(&acc_empty[tt % NACC], ((tt / NACC) & 1) ^ 1, abort_flag, p.watchdog_ns);
              if (!__all_sync(0xffffffffu, ok)) run = false;
            }
          }
          if (!have_slabs) {
            for (int i = 0; i < nt; ++i) {
              const bool ok = mbar_wait(&slab_full[sslot[i]], sphase[i], abort_flag, p.watchdog_ns);
              if (!__all_sync(0xffffffffu, ok)) run = false;
              da0[i] = umma_smem_desc_sbo(smem_u32(slabs + static_cast<size_t>(sslot[i]) * SLAB_STRIDE), ROWB, SW * ROWB);
            }
            have_slabs = true;
          }
          {
            const bool ok = mbar_wait(&b_full[bsc], bphc, abort_flag, p.watchdog_ns);
            if (!__all_sync(0xffffffffu, ok)) run = false;
          }
          if (!run) break;
          const uint64_t db0 = umma_smem_desc(smem_u32(wring + static_cast<size_t>(bsc) * WT_BYTES), ROWB);
          const int dy = tap / 3, dx = tap - 3 * dy;
          const uint32_t a_off = static_cast<uint32_t>(((dy * SW + dx) * ROWB) >> 4);
          // ---- the turn
          if (ni > 1) {
            const uint32_t par = me == 0 ? ((nth & 1) ^ 1) : (nth & 1);
            const bool ok3 = mbar_wait(&turn[me], par, abort_flag, p.watchdog_ns);
            if (!__all_sync(0xffffffffu, ok3)) {
              run = false;
              break;
            }
          }
          tc_fence_after();
          if (elect_one()) {
            for (int i = 0; i < nt; ++i) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma_bf16_lo_acc(tacc[i], desc_lo(da0[i]) + a_off + 2 * k, desc_lo(db0) + 2 * k, desc_hi(da0[i]), desc_hi(db0), IDESC,
                                 (ch | tap | k) != 0 ? 1u : 0u);
            }
            if (ni > 1) mbar_arrive(&turn[me + 1 == ni ? 0 : me + 1]);  // hand the token on before the commits
            umma_commit(&b_empty[bsc]);
            if (tap == 8) {  // in-order pipe: this step's MMAs retire after every earlier read of the slabs
              for (int i = 0; i < nt; ++i) umma_commit(&slab_empty[sslot[i]]);
              if (ch == chunks - 1)
                for (int i = 0; i < nt; ++i) umma_commit(&acc_full[(t + i) % NACC]);
            }
          }
          __syncwarp();
          ++nth;
        }
      }
    }
  } else {
    // ===================== epilogue: warpgroup g handles local tiles t == g (mod 4) =====================
    const int g = (warp - kSlabwFirstEpiWarp) >> 2;
    const int q = warp & 3;
    for (int t = g; t < ntl; t += NACC) {
      const int m = m_first + t * m_stride;
      const int n = m / tiles_xy;
      const int rem = m - n * tiles_xy;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const bool ready = mbar_wait(&acc_full[g], (static_cast<uint32_t>(t) / NACC) & 1, abort_flag, p.watchdog_ns);
      if (!__all_sync(0xffffffffu, ready)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + g * BN + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_slab<BN, EPI>(p, &tmOut, &tmPool, taddr, lane, q, tx * 8, ty * 16, n, nb0, s_bias, s_extra,
                             staging + static_cast<size_t>(warp - kSlabwFirstEpiWarp) * slab_stage_warp_bytes(EPI), [&] {
                               tc_fence_before();
                               mbar_arrive(&acc_empty[g]);
                             });
    }
    if (lane == 0) bulk_wait_read<0>();  // staging must stay valid until the last stores have read it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, NACC * BN);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

cudaError_t conv_slabw_launch(const ConvLaunch& L, cudaStream_t stream);
cudaError_t conv_slabw_init_attributes();

}  // namespace scv
