// Two 3x3 convolutions of the 384-wide decoder tail fused into ONE row-streaming kernel:
//   decoder_0/conv0 (64 -> 32, BN + ReLU)  ->  decoder_0/conv1 (32 -> 32, BN + ReLU) + 1x1 head
// (utils/model_tools.py:178-186 twice, :405 / :443), so that the 32-channel intermediate never goes to HBM.
//
// Why: both layers are HBM-bound as separate kernels (profiles/r02_store_path.md: decoder_0/conv0 moves 1.77 GB per
// 63 chips at 6.0 TB/s, and its 9.4 MB-per-chip output is written only to be read back by the next launch).
//
// How the halo is handled without recomputing anything: a CLUSTER of three CTAs works on the same image rows, one
// 128-pixel strip each (3 x 128 = the 384-pixel row).  Every CTA runs the row kernel's machinery twice:
//   conv 1: TMA slabs (two input rows) -> tap-packed UMMAs (N = 96) -> row accumulators in TMEM columns [0, 256)
//           -> epilogue A: bias + ReLU -> bf16 -> a ring of shared-memory slabs laid out exactly like a TMA-written
//              SWIZZLE_64B slab (2 rows x 130 pixels x 32 channels); the strip's first / last pixel are ALSO written
//              into the neighbour CTA's slab as its right / left halo pixel (st.shared::cluster), zeros at the image
//              border;
//   conv 2: the same UMMAs reading those slabs -> row accumulators in TMEM columns [256, 512) -> epilogue B: bias +
//           ReLU + 1x1 head in fp32 -> logits.
// Conv 1 runs one output row ahead and behind of conv 2's range inside a segment (rows outside the image are written
// as zeros = conv 2's 'same' padding).  Summation orders are those of conv_rows_kernel, so the logits are bit-identical
// to the two-launch path.
//
// Ring protocol (slot s, use k): conv-1 epilogue threads wait slot_free[s], which counts one arrival per CTA whose
// conv-2 UMMAs read a copy of the data written for use k-1 -- this CTA and its neighbour(s): each conv-2 issuer
// multicasts ONE tcgen05.commit to the barrier at that offset in all of them.  ring_full[s] counts 128 own arrivals +
// one per halo pixel column = 130.  The halo pixel of a side with a neighbour arrives as an async-proxy bulk copy
// (cp.async.bulk shared::cta -> shared::cluster, 2 rows x 64 B, complete_tx on THIS CTA's ring_full[s]); the edge thread
// of that side contributes the matching arrive.expect_tx.  (The first version used st.shared::cluster + fence.proxy.async
// + mbarrier.arrive.release.cluster: SASS shows a MEMBAR.ALL.GPU for each of the two -- 1 400-2 400 cycles per row pair on
// the path conv 2 waits for.  -DSCV_F2_HALO_GENERIC keeps that version for comparison.)
#pragma once
#include "conv_rows.cuh"

namespace scv {

#ifndef SCV_F2_NI1
#define SCV_F2_NI1 2
#endif
#ifndef SCV_F2_NS16
#define SCV_F2_NS16 6
#endif
#ifndef SCV_F2_NR16
#define SCV_F2_NR16 5
#endif
constexpr int kF2Issuers1 = SCV_F2_NI1, kF2Issuers2 = 2;                   // issuer warps per conv (taking turns: a single issuer
                                                                  // spends > 1000 cycles per row on waits and commits)
// Epilogue warpgroups (four per kernel, 768 threads).  The decoder tail (fused head) gives three to conv 1 -- whose
// proxy fence + arrive is what conv 2 waits for -- and one to the light head epilogue (13.5 ms with 2 + 2 -> 13.3 with
// 3 + 1); the encoder pair, whose second epilogue pools, applies the skip affine and stores two tensors, splits 2 + 2
// (19.4 ms with 1 + 3 -> 14.1).  The epilogues work on 8 channels at a time (61-72 registers), so six groups (1024
// threads, -DSCV_F2_GROUPS_POOL=6) fit without spills -- measured 13.9 vs 14.0 ms: warps are not what it lacks.
#ifndef SCV_F2_G1_HEAD
#define SCV_F2_G1_HEAD 3
#endif
#ifndef SCV_F2_G1_POOL
#define SCV_F2_G1_POOL 2
#endif
#ifndef SCV_F2_GROUPS_POOL
#define SCV_F2_GROUPS_POOL 4
#endif
// staging tiles per epilogue-B warp (EPI_POOL_SKIP): with 2 a warp fills one tile while the TMA stores of the previous
// pair still read the other (cp.async.bulk.wait_group.read 1)
#ifndef SCV_F2_NSTAGE
#define SCV_F2_NSTAGE 1
#endif
__host__ __device__ constexpr int f2_groups(int epi2) { return epi2 == EPI_POOL_SKIP ? SCV_F2_GROUPS_POOL : 4; }
__host__ __device__ constexpr int f2_groups1(int epi2) { return epi2 == EPI_POOL_SKIP ? SCV_F2_G1_POOL : SCV_F2_G1_HEAD; }
__host__ __device__ constexpr int f2_groups2(int epi2) { return f2_groups(epi2) - f2_groups1(epi2); }
constexpr int kF2Issuer2Warp = 1 + kF2Issuers1;                   // warp 0: TMA, 1..2: conv-1 issuers, 3..4: conv-2 issuers
constexpr int kF2FirstEpi1 = 8;                                   // warps 5..7 idle (TMEM lane quadrant == warp & 3)
__host__ __device__ constexpr int f2_threads(int epi2) { return 32 * (kF2FirstEpi1 + 4 * f2_groups(epi2)); }
constexpr int kF2R = 8;                                           // row accumulators per conv: 8 x 32 columns = 256
constexpr int kF2RP = kF2R / 2;
// conv-1 input slabs (two rows of 130 px x KC1 channels) and conv-1 -> conv-2 slabs (two rows of 130 px x 32 channels);
// the 64-channel decoder tail fills shared memory with 3 + 4, the 16-channel encoder pair has room for more
__host__ __device__ constexpr int f2_in_slabs(int kc1) { return kc1 == 16 ? SCV_F2_NS16 : 3; }
__host__ __device__ constexpr int f2_ring(int kc1) { return kc1 == 16 ? SCV_F2_NR16 : 4; }
constexpr int kF2Cluster = 3;                                     // strips per image row
static_assert(SCV_F2_NSTAGE == 1 || SCV_F2_NSTAGE == 2, "one or two staging tiles per epilogue-B warp");
static_assert(kF2FirstEpi1 % 4 == 0, "epilogue warps must start on a TMEM quadrant boundary");
static_assert(kF2RP >= 4 && kF2RP >= kF2Issuers1 && kF2RP >= kF2Issuers2, "accumulator reuse distance (<= 4 groups per conv)");
static_assert(kF2Issuer2Warp + kF2Issuers2 <= kF2FirstEpi1 && kF2Issuers1 <= 3, "warp layout / ring depth");

__host__ __device__ inline size_t fused_smem_bytes(int KC1, int epi2, int ncls) {
  size_t s = 1024 + static_cast<size_t>(9) * 32 * KC1 * 2 + static_cast<size_t>(9) * 32 * 64 +
             static_cast<size_t>(f2_in_slabs(KC1)) * rows_slab_stride(KC1) + static_cast<size_t>(f2_ring(KC1)) * rows_slab_stride(32) +
             static_cast<size_t>(4 * f2_groups2(epi2)) * SCV_F2_NSTAGE * rows_stage_warp_bytes(epi2);
  s += (1 + 2 * f2_in_slabs(KC1) + 4 * kF2RP + kF2Issuers1 + kF2Issuers2 + 2 * f2_ring(KC1)) * 8 + 16;
  s += (32 + 32 + (epi2 == EPI_HEAD ? 32 * ncls + ncls : 64)) * 4;
  return s + 64;
}

// ---- cluster PTX ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t laddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster128(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// async-proxy copy of `bytes` (multiple of 16) from this CTA's shared memory into a peer CTA's, completing on the PEER's
// mbarrier (complete_tx): the whole transfer stays in the async proxy, no generic-proxy remote store / cluster fence
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t peer_dst, uint32_t src, uint32_t bytes, uint32_t peer_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(peer_dst),
               "r"(src), "r"(bytes), "r"(peer_bar)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit that arrives on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// mbar_wait with cluster-scope acquire (data written by another CTA of the cluster)
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, volatile int* abort_flag, uint64_t budget_ns) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (true) {
    if (mbar_try_wait_cluster(bar, parity)) return true;
    if ((++spins & 0x3f) == 0) {
      if (*abort_flag) return false;
      if (globaltimer_ns() - t0 > budget_ns) {
        *abort_flag = 1;
        return false;
      }
    }
  }
}

// rows_piece with an explicit ring size (R row accumulators of COUT columns starting at `base`)
template <int COUT, int R>
__device__ __forceinline__ RowsPiece rows_piece_r(uint32_t base, int j, int npairs, uint32_t opc) {
  RowsPiece r;
  const int i_lo = j >= 2 ? j - 2 : 0;
  const int i_hi = j < 2 * npairs ? j : 2 * npairs - 1;
  const int nblk = i_hi - i_lo + 1;
  const uint32_t slot_lo = (2 * opc + i_lo) % R;
  const int n1 = nblk < static_cast<int>(R - slot_lo) ? nblk : static_cast<int>(R - slot_lo);
  const int n2 = nblk - n1;
  r.d1 = base + slot_lo * COUT;
  r.id1 = umma_idesc_bf16(128, n1 * COUT);
  r.d2 = base;
  r.id2 = n2 > 0 ? umma_idesc_bf16(128, n2 * COUT) : 0u;
  r.b_off = static_cast<uint32_t>(2 - (j - i_lo));
  r.b_wrap = static_cast<uint32_t>(n1);
  return r;
}

// KC1: channel chunk of conv 1 (== its padded Cin: one chunk; 16 = the 8-channel first layer read through 16-channel
// TMA boxes).  Conv 2 is 32 -> 32; EPI2 = EPI_HEAD (decoder tail: fused 1x1 head -> logits) or EPI_POOL_SKIP (encoder
// pair: 2x2 max-pool + skip affine, two TMA-stored outputs).
template <int KC1, int EPI2>
__global__ void __launch_bounds__(f2_threads(EPI2), 1)
    conv_fused2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                       const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
                       const __grid_constant__ CUtensorMap tmPool, const ConvParams p) {
  constexpr int COUT = 32;
  constexpr int NG1 = f2_groups1(EPI2), NG2 = f2_groups2(EPI2);
  constexpr int kF2FirstEpi2 = kF2FirstEpi1 + 4 * NG1;
  constexpr int STAGE_W = rows_stage_warp_bytes(EPI2);
  static_assert(EPI2 == EPI_HEAD || EPI2 == EPI_POOL_SKIP, "second epilogue: fused head or pool + skip");
  constexpr int ROWB1 = KC1 * 2, WT1 = COUT * ROWB1;        // conv-1 weight tile per (kx, ky)
  constexpr int ROWB2 = 64, WT2 = COUT * ROWB2;             // conv-2: 32 input channels
  constexpr int ROW1 = kRowsSlabPx * ROWB1, SLAB1 = 2 * ROW1, STRIDE1 = rows_slab_stride(KC1);
  constexpr int ROW2 = kRowsSlabPx * ROWB2, STRIDE2 = rows_slab_stride(32);
  constexpr int NS1 = f2_in_slabs(KC1), NR = f2_ring(KC1), R = kF2R, RP = kF2RP, NI1 = kF2Issuers1, NI2 = kF2Issuers2;
  constexpr uint32_t ACC2 = R * COUT;                       // first TMEM column of conv 2's accumulators

  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = base;                                     // 1024-aligned slabs first
  uint8_t* slabs = ring + static_cast<size_t>(NR) * STRIDE2;
  uint8_t* w1 = slabs + static_cast<size_t>(NS1) * STRIDE1;  // [kx][ky = 2,1,0][32 rows][KC1]
  uint8_t* w2 = w1 + 9 * WT1;
  uint8_t* staging = w2 + 9 * WT2;  // epilogue B's TMA-store staging tiles (EPI_POOL_SKIP)
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 4 * NG2 * SCV_F2_NSTAGE * STAGE_W);
  uint64_t* w_full = bars;
  uint64_t* in_full = bars + 1;
  uint64_t* in_empty = in_full + NS1;
  uint64_t* acc1_full = in_empty + NS1;
  uint64_t* acc1_empty = acc1_full + RP;
  uint64_t* acc2_full = acc1_empty + RP;
  uint64_t* acc2_empty = acc2_full + RP;
  uint64_t* turn = acc2_empty + RP;
  uint64_t* turn2 = turn + NI1;
  uint64_t* ring_full = turn2 + NI2;
  uint64_t* slot_free = ring_full + NR;     // slot s may be rewritten here AND in the neighbours (multicast conv-2 commits)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_free + NR);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_bias1 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
  float* s_bias2 = s_bias1 + COUT;
  float* s_head = s_bias2 + COUT;  // EPI_HEAD: [32][ncls] + [ncls]; EPI_POOL_SKIP: skip scale [32] + shift [32]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H2 = p.H >> 1;
  const uint32_t xs = cluster_ctarank();                    // this CTA's strip
  const int ncl = gridDim.x / kF2Cluster, cl = blockIdx.x / kF2Cluster;
  const long long P = static_cast<long long>(p.N) * H2;     // (image, output row pair) units, shared by the cluster
  const long long p0 = P * cl / ncl, p1 = P * (cl + 1) / ncl;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmB2);
    if constexpr (EPI2 == EPI_POOL_SKIP) {
      tma_prefetch_desc(&tmOut);
      tma_prefetch_desc(&tmPool);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int s = 0; s < NS1; ++s) {
        mbar_init(&in_full[s], 1);
        mbar_init(&in_empty[s], 1);
      }
      for (int a = 0; a < RP; ++a) {
        mbar_init(&acc1_full[a], 1);
        mbar_init(&acc1_empty[a], 128);
        mbar_init(&acc2_full[a], 1);
        mbar_init(&acc2_empty[a], 128);
      }
      for (int i = 0; i < NI1; ++i) mbar_init(&turn[i], 1);
      for (int i = 0; i < NI2; ++i) mbar_init(&turn2[i], 1);
      for (int s = 0; s < NR; ++s) {
        mbar_init(&ring_full[s], 130);
        mbar_init(&slot_free[s], 1 + (xs > 0 ? 1 : 0) + (xs + 1 < kF2Cluster ? 1 : 0));
      }
      *abort_flag = 0;
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp >= kF2FirstEpi1) {
    const int t = threadIdx.x - 32 * kF2FirstEpi1, nt = f2_threads(EPI2) - 32 * kF2FirstEpi1;
    for (int i = t; i < COUT; i += nt) {
      s_bias1[i] = p.bias[i];
      s_bias2[i] = p.bias2[i];
    }
    if constexpr (EPI2 == EPI_HEAD) {
      for (int i = t; i < COUT * p.ncls; i += nt) s_head[i] = p.head_w[i];
      for (int i = t; i < p.ncls; i += nt) s_head[COUT * p.ncls + i] = p.head_b[i];
    } else {
      for (int i = t; i < COUT; i += nt) {
        s_head[i] = p.skip_s[i];
        s_head[COUT + i] = p.skip_t[i];
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // every CTA's barriers exist before any remote arrive (also a CTA-wide barrier)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kF2FirstEpi1 && warp < kF2FirstEpi1 + 4) {  // one warp per TMEM lane quadrant zeroes all 512 columns
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 512; c += 32) tmem_st32_zero(t0 + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===================== TMA producer: conv-1 input row pairs (y0-2+2v, y0-1+2v), v = 0 .. npairs+1 ===========
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, 9 * WT1 + 9 * WT2);
      for (int kx = 0; kx < 3; ++kx)
        for (int b = 0; b < 3; ++b) {  // block b holds ky = 2 - b; global K index = tap * Cin, tap = ky * 3 + kx
          tma_load_2d(w1 + static_cast<size_t>(kx * 3 + b) * WT1, &tmB1, w_full, ((2 - b) * 3 + kx) * KC1, 0);
          tma_load_2d(w2 + static_cast<size_t>(kx * 3 + b) * WT2, &tmB2, w_full, ((2 - b) * 3 + kx) * 32, 0);
        }
    }
    __syncwarp();
    uint32_t s = 0, ph = 1;
    long long pc = p0;
    RowSeg sg;
    bool run = true;
    while (run && rows_next_seg(pc, p1, 1, H2, sg)) {
      for (int v = 0; run && v <= sg.npairs + 1; ++v) {
        const bool ok = mbar_wait(&in_empty[s], ph, abort_flag, p.watchdog_ns);
        if (!__all_sync(0xffffffffu, ok)) {
          run = false;
          break;
        }
        if (elect_one()) {
          mbar_arrive_expect_tx(&in_full[s], SLAB1);
          // rows < 0 and >= H, pixels -1 and W are out of bounds: zero filled == the tile's own 'same' padding
          tma_load_4d(slabs + static_cast<size_t>(s) * STRIDE1, &tmA, &in_full[s], 0, static_cast<int>(xs) * kRowsPx - 1,
                      sg.y0 - 2 + 2 * v, sg.n + p.n_in_off);
        }
        __syncwarp();
        if (++s == NS1) s = 0, ph ^= 1;
      }
    }
  } else if (warp < kF2Issuer2Warp) {
    // ===================== conv-1 issuers: issuer k takes the CTA's input pairs k, k+ni, ... =====================
    // conv 1 covers output rows [y0-1, y0+2*npairs+1): npairs+1 output pairs, npairs+2 input pairs per segment
    const int me = warp - 1;
    bool run = __all_sync(0xffffffffu, mbar_wait(w_full, 0, abort_flag, p.watchdog_ns));
    tc_fence_after();
    const uint32_t w_desc = static_cast<uint32_t>(umma_smem_desc(smem_u32(w1), ROWB1) & 0xffffffffu);
    const uint32_t desc_hi = static_cast<uint32_t>(umma_smem_desc(0, ROWB1) >> 32);
    const uint32_t slab0_desc = static_cast<uint32_t>(umma_smem_desc(smem_u32(slabs), ROWB1) & 0xffffffffu);
    uint32_t nth = 0, t = 0, opc = 0;
    long long pc = p0;
    RowSeg sg;
    while (run && rows_next_seg(pc, p1, 1, H2, sg)) {
      const int np1 = sg.npairs + 1;  // conv-1 output pairs of this segment
      for (int v = 0; run && v <= np1; ++v, ++t) {
        if (static_cast<int>(t % NI1) != me) continue;
        const uint32_t s0 = t % NS1, ph0 = (t / NS1) & 1;
        if (v < np1) {  // output pair v enters the ring: its accumulators must have been drained and zeroed
          const uint32_t op = opc + v;
          const bool ok = mbar_wait(&acc1_empty[op % RP], ((op / RP) & 1) ^ 1, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok)) {
            run = false;
            break;
          }
        }
        const RowsPiece pc0 = rows_piece_r<COUT, R>(tmem_base, 2 * v, np1, opc);
        const RowsPiece pc1 = rows_piece_r<COUT, R>(tmem_base, 2 * v + 1, np1, opc);
        const bool ok2 = mbar_wait(&in_full[s0], ph0, abort_flag, p.watchdog_ns);
        if (!__all_sync(0xffffffffu, ok2)) {
          run = false;
          break;
        }
        if (!(me == 0 && nth == 0)) {  // issuer 0 starts with the token
          const uint32_t par = me == 0 ? ((nth - 1) & 1) : (nth & 1);
          const bool ok3 = mbar_wait(&turn[me], par, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok3)) {
            run = false;
            break;
          }
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t da0 = slab0_desc + s0 * static_cast<uint32_t>(STRIDE1 >> 4);
          rows_issue_row<KC1, COUT>(da0, w_desc, desc_hi, pc0.d1, pc0.id1, pc0.d2, pc0.id2, pc0.b_off, pc0.b_wrap);
          rows_issue_row<KC1, COUT>(da0 + (ROW1 >> 4), w_desc, desc_hi, pc1.d1, pc1.id1, pc1.d2, pc1.id2, pc1.b_off, pc1.b_wrap);
          mbar_arrive(&turn[me + 1 == NI1 ? 0 : me + 1]);  // hand the token on before the (slow) commits
          umma_commit(&in_empty[s0]);
          if (v >= 1) umma_commit(&acc1_full[(opc + v - 1) % RP]);  // output pair v-1 is complete
        }
        __syncwarp();
        ++nth;
      }
      opc += np1;
    }
  } else if (warp < kF2Issuer2Warp + NI2) {
    // ===================== conv-2 issuers (taking turns): input pair w == conv-1 output pair w of the segment ====
    const int me = warp - kF2Issuer2Warp;
    bool run = __all_sync(0xffffffffu, mbar_wait(w_full, 0, abort_flag, p.watchdog_ns));
    tc_fence_after();
    const uint32_t w_desc = static_cast<uint32_t>(umma_smem_desc(smem_u32(w2), ROWB2) & 0xffffffffu);
    const uint32_t desc_hi = static_cast<uint32_t>(umma_smem_desc(0, ROWB2) >> 32);
    const uint32_t ring0_desc = static_cast<uint32_t>(umma_smem_desc(smem_u32(ring), ROWB2) & 0xffffffffu);
    const uint16_t mask = static_cast<uint16_t>((1u << xs) | (xs > 0 ? 1u << (xs - 1) : 0u) | (xs + 1 < kF2Cluster ? 1u << (xs + 1) : 0u));
    uint32_t ip = 0, opc = 0, nth = 0;  // running input-pair (== ring use) and output-pair counters, own turns
    long long pc = p0;
    RowSeg sg;
    while (run && rows_next_seg(pc, p1, 1, H2, sg)) {
      for (int w = 0; run && w <= sg.npairs; ++w, ++ip) {
        if (static_cast<int>(ip % NI2) != me) continue;
        const uint32_t rs = ip % NR, rph = (ip / NR) & 1;
        if (w < sg.npairs) {
          const uint32_t op = opc + w;
          const bool ok = mbar_wait(&acc2_empty[op % RP], ((op / RP) & 1) ^ 1, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok)) {
            run = false;
            break;
          }
        }
        const RowsPiece pc0 = rows_piece_r<COUT, R>(tmem_base + ACC2, 2 * w, sg.npairs, opc);
        const RowsPiece pc1 = rows_piece_r<COUT, R>(tmem_base + ACC2, 2 * w + 1, sg.npairs, opc);
#ifdef SCV_F2_HALO_GENERIC
        const bool ok2 = mbar_wait_cluster(&ring_full[rs], rph, abort_flag, p.watchdog_ns);
#else
        const bool ok2 = mbar_wait(&ring_full[rs], rph, abort_flag, p.watchdog_ns);
#endif
        if (!__all_sync(0xffffffffu, ok2)) {
          run = false;
          break;
        }
        if (!(me == 0 && nth == 0)) {  // issuer 0 starts with the token
          const uint32_t par = me == 0 ? ((nth - 1) & 1) : (nth & 1);
          const bool ok3 = mbar_wait(&turn2[me], par, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok3)) {
            run = false;
            break;
          }
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t da0 = ring0_desc + rs * static_cast<uint32_t>(STRIDE2 >> 4);
          rows_issue_row<32, COUT>(da0, w_desc, desc_hi, pc0.d1, pc0.id1, pc0.d2, pc0.id2, pc0.b_off, pc0.b_wrap);
          rows_issue_row<32, COUT>(da0 + (ROW2 >> 4), w_desc, desc_hi, pc1.d1, pc1.id1, pc1.d2, pc1.id2, pc1.b_off, pc1.b_wrap);
          mbar_arrive(&turn2[me + 1 == NI2 ? 0 : me + 1]);  // hand the token on before the commits
          umma_commit_multicast(&slot_free[rs], mask);      // the slot may be refilled: here and (halo pixels) next door
          if (w >= 1) umma_commit(&acc2_full[(opc + w - 1) % RP]);
        }
        __syncwarp();
        ++nth;
      }
      opc += sg.npairs;
    }
  } else if (warp < kF2FirstEpi1) {
    // idle warps (keep the epilogue warps on TMEM quadrant boundaries)
  } else if (warp < kF2FirstEpi2) {
    // ===================== epilogue A: conv-1 accumulators -> bf16 slabs for conv 2 (+ halo exchange) ============
    const int ew = warp - kF2FirstEpi1;
    const int g = ew >> 2;
    const int q = warp & 3;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const bool edge_l = q == 0 && lane == 0;    // owns strip pixel 0   -> left neighbour's halo pixel 129
    const bool edge_r = q == 3 && lane == 31;   // owns strip pixel 127 -> right neighbour's halo pixel 0
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t px_off = static_cast<uint32_t>(q * 32 + lane + 1) * ROWB2;  // this lane's pixel inside a slab row
    uint32_t op1 = 0;  // running conv-1 output pair (== ring use) counter
    long long pc = p0;
    RowSeg sg;
    bool run = true;
    long long ta_acc = 0, ta_free = 0, ta_ld = 0, ta_math = 0, ta_fence = 0, ta_begin = ROWS_CLOCK();
    uint32_t ta_n = 0;
    while (run && rows_next_seg(pc, p1, 1, H2, sg)) {
      for (int u = 0; u <= sg.npairs; ++u, ++op1) {
        if (op1 % NG1 != static_cast<uint32_t>(g)) continue;
        const uint32_t slot = op1 % RP, rs = op1 % NR, use = op1 / NR;
        const long long c0 = ROWS_CLOCK();
        bool ok = mbar_wait(&acc1_full[slot], (op1 / RP) & 1, abort_flag, p.watchdog_ns);
        const long long c1 = ROWS_CLOCK();
        ok = ok && mbar_wait(&slot_free[rs], (use & 1) ^ 1, abort_flag, p.watchdog_ns);
        const long long c2 = ROWS_CLOCK();
        ta_acc += c1 - c0;
        ta_free += c2 - c1;
        ++ta_n;
        if (!__all_sync(0xffffffffu, ok)) {
          run = false;
          break;
        }
        tc_fence_after();
        const uint32_t taddr = tq + slot * (2 * COUT);
        const int ya = sg.y0 - 1 + 2 * u;                       // rows ya, ya + 1; outside the image -> zeros (padding)
        const bool in_a = ya >= 0 && ya < p.H, in_b = ya + 1 >= 0 && ya + 1 < p.H;
        const uint32_t slab = ring_u32 + rs * STRIDE2;
        const uint32_t nb_l = (edge_l && xs > 0) ? mapa_shared(slab, xs - 1) : 0u;
        const uint32_t nb_r = (edge_r && xs + 1 < kF2Cluster) ? mapa_shared(slab, xs + 1) : 0u;
        // 8 channels = one 16-byte chunk of a pixel row per step (keeps the epilogue within 64 registers, which is what
        // lets six epilogue warpgroups share the register file)
        // a slab is laid out like a TMA-written SWIZZLE_64B box: 16-byte chunk c of the pixel row at byte offset o
        // sits at o + ((c ^ ((o >> 7) & 3)) << 4)  (slabs are 1024-byte aligned)
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t r0[8], r1[8];
          const long long c3 = ROWS_CLOCK();
          tmem_ld8(taddr + c8 * 8, r0);
          tmem_ld8(taddr + COUT + c8 * 8, r1);
          tmem_ld_wait();
          if (c8 == 3) {  // everything read: zero both accumulators and hand them back
#pragma unroll
            for (int c = 0; c < 2 * COUT; c += 32) tmem_st32_zero(taddr + c);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&acc1_empty[slot]);
          }
          ta_ld += ROWS_CLOCK() - c3;
          float bias[8];
          lds8(s_bias1 + c8 * 8, bias);
          uint32_t pk0[4], pk1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pk0[j] = pack_bf16x2(__uint_as_float(r0[2 * j]) + bias[2 * j], __uint_as_float(r0[2 * j + 1]) + bias[2 * j + 1]);
            pk1[j] = pack_bf16x2(__uint_as_float(r1[2 * j]) + bias[2 * j], __uint_as_float(r1[2 * j + 1]) + bias[2 * j + 1]);
            if (p.relu) {
              pk0[j] = max_bf16x2(pk0[j], 0u);
              pk1[j] = max_bf16x2(pk1[j], 0u);
            }
            if (!in_a) pk0[j] = 0u;
            if (!in_b) pk1[j] = 0u;
          }
          auto put_px = [&](uint32_t slab_addr, uint32_t off, const uint32_t (&v)[4], bool remote) {
            const uint32_t a0 = slab_addr + off + ((static_cast<uint32_t>(c8) ^ ((off >> 7) & 3)) << 4);
            if (remote) st_cluster128(a0, v[0], v[1], v[2], v[3]);
            else sts128(a0, v[0], v[1], v[2], v[3]);
          };
          put_px(slab, px_off, pk0, false);
          put_px(slab, ROW2 + px_off, pk1, false);
          const uint32_t zero4[4] = {0u, 0u, 0u, 0u};
#ifdef SCV_F2_HALO_GENERIC
          if (edge_l) {  // strip pixel 0: the left neighbour's halo pixel 129, or this CTA's own (zero) halo pixel 0
            if (xs > 0) {
              put_px(nb_l, (kRowsSlabPx - 1) * ROWB2, pk0, true);
              put_px(nb_l, ROW2 + (kRowsSlabPx - 1) * ROWB2, pk1, true);
            } else {
              put_px(slab, 0, zero4, false);
              put_px(slab, ROW2, zero4, false);
            }
          }
          if (edge_r) {
            if (xs + 1 < kF2Cluster) {
              put_px(nb_r, 0, pk0, true);
              put_px(nb_r, ROW2, pk1, true);
            } else {
              put_px(slab, (kRowsSlabPx - 1) * ROWB2, zero4, false);
              put_px(slab, ROW2 + (kRowsSlabPx - 1) * ROWB2, zero4, false);
            }
          }
#else
          if (edge_l && xs == 0) {  // image border: this CTA's own halo pixel 0 is zero padding
            put_px(slab, 0, zero4, false);
            put_px(slab, ROW2, zero4, false);
          }
          if (edge_r && xs + 1 == kF2Cluster) {
            put_px(slab, (kRowsSlabPx - 1) * ROWB2, zero4, false);
            put_px(slab, ROW2 + (kRowsSlabPx - 1) * ROWB2, zero4, false);
          }
#endif
        }
        const long long c4 = ROWS_CLOCK();
#ifdef SCV_F2_HALO_GENERIC
        // generic-proxy stores (local and remote) -> visible to the UMMAs that read the slab
        fence_proxy_async_all();
        mbar_arrive(&ring_full[rs]);
        if (edge_l) {
          if (xs > 0) mbar_arrive_remote(mapa_shared(smem_u32(&ring_full[rs]), xs - 1));
          else mbar_arrive(&ring_full[rs]);
        }
        if (edge_r) {
          if (xs + 1 < kF2Cluster) mbar_arrive_remote(mapa_shared(smem_u32(&ring_full[rs]), xs + 1));
          else mbar_arrive(&ring_full[rs]);
        }
#else
        // own generic-proxy stores -> visible to the async proxy (the UMMAs that read the slab, and the bulk copies below)
        fence_proxy_async();
        mbar_arrive(&ring_full[rs]);
        // Halo exchange.  Strip pixel 0 (slab pixel 1) is the left neighbour's halo pixel 129, strip pixel 127 (slab
        // pixel 128) the right neighbour's halo pixel 0.  Source and destination pixel are 128 pixels = 8 KB apart, so
        // they share the SWIZZLE_64B phase and each 64-byte pixel row copies verbatim.  The copies complete_tx on the
        // NEIGHBOUR's ring_full[rs]; the bytes flowing INTO this CTA's halo are expected by this side's edge thread.
        if (edge_l) {
          if (xs > 0) {
            const uint32_t bar_l = mapa_shared(smem_u32(&ring_full[rs]), xs - 1);
            bulk_copy_to_peer(nb_l + (kRowsSlabPx - 1) * ROWB2, slab + ROWB2, ROWB2, bar_l);
            bulk_copy_to_peer(nb_l + ROW2 + (kRowsSlabPx - 1) * ROWB2, slab + ROW2 + ROWB2, ROWB2, bar_l);
            mbar_arrive_expect_tx(&ring_full[rs], 2 * ROWB2);
          } else {
            mbar_arrive(&ring_full[rs]);
          }
        }
        if (edge_r) {
          if (xs + 1 < kF2Cluster) {
            const uint32_t bar_r = mapa_shared(smem_u32(&ring_full[rs]), xs + 1);
            bulk_copy_to_peer(nb_r, slab + (kRowsSlabPx - 2) * ROWB2, ROWB2, bar_r);
            bulk_copy_to_peer(nb_r + ROW2, slab + ROW2 + (kRowsSlabPx - 2) * ROWB2, ROWB2, bar_r);
            mbar_arrive_expect_tx(&ring_full[rs], 2 * ROWB2);
          } else {
            mbar_arrive(&ring_full[rs]);
          }
        }
#endif
        ta_fence += ROWS_CLOCK() - c4;
        ta_math += c4 - c2;
      }
    }
#ifdef SCV_ROWS_PROF
    if ((p.dbg & 32) && blockIdx.x < 3 && lane == 0 && q == 0)
      printf("[fused prof] cta %d epilogue A group %d: total %lld cyc over %u pairs: wait acc1_full %lld, wait slot_free %lld, tmem ld+zero %lld, math+stores (incl. ld) %lld, fence+arrive %lld\n",
             blockIdx.x, g, ROWS_CLOCK() - ta_begin, ta_n, ta_acc, ta_free, ta_ld, ta_math, ta_fence);
#endif
  } else {
    // ===================== epilogue B: conv-2 accumulators -> fused head (logits) | pool + skip (TMA stores) ======
    const int ew = warp - kF2FirstEpi2;
    const int g = ew >> 2;
    const int q = warp & 3;
    const uint32_t tq = tmem_base + ACC2 + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* stage_base = staging + static_cast<size_t>(ew) * (SCV_F2_NSTAGE * STAGE_W);
    uint32_t stage_i = 0;
    const uint32_t phase = (lane >> 1) & 3;  // SWIZZLE_64B phase of staging rows `lane` and `32 + lane`
    uint32_t opc = 0;
    long long pc = p0;
    RowSeg sg;
    bool run = true;
    long long tb_wait = 0, tb_rd = 0, tb_math = 0, tb_store = 0, tb_begin = ROWS_CLOCK();
    uint32_t tb_n = 0;
    while (run && rows_next_seg(pc, p1, 1, H2, sg)) {
      for (int u = 0; u < sg.npairs; ++u) {
        const uint32_t op = opc + u;
        if (op % NG2 != static_cast<uint32_t>(g)) continue;
        const uint32_t slot = op % RP;
        const long long c0 = ROWS_CLOCK();
        const bool ready = mbar_wait(&acc2_full[slot], (op / RP) & 1, abort_flag, p.watchdog_ns);
        tb_wait += ROWS_CLOCK() - c0;
        ++tb_n;
        if (!__all_sync(0xffffffffu, ready)) {
          run = false;
          break;
        }
        tc_fence_after();
        const uint32_t taddr = tq + slot * (2 * COUT);
        const int xw = static_cast<int>(xs) * kRowsPx + q * 32;  // first pixel of this warp
        const int y = sg.y0 + 2 * u;
        if constexpr (EPI2 == EPI_HEAD) {
          epilogue_head<COUT>(p, taddr, 0, 0, xw + lane, y, sg.n, true, 0, s_bias2, s_head);
          epilogue_head<COUT>(p, taddr + COUT, 0, 0, xw + lane, y + 1, sg.n, true, 0, s_bias2, s_head);
#pragma unroll
          for (int c = 0; c < 2 * COUT; c += 32) tmem_st32_zero(taddr + c);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&acc2_empty[slot]);
        } else {
          // the row kernel's pooling epilogue (conv_rows.cuh), 16 channels at a time
          const long long d0 = ROWS_CLOCK();
          uint8_t* stage = stage_base + stage_i * STAGE_W;
          if (SCV_F2_NSTAGE > 1) stage_i ^= 1u;
          if (lane == 0) bulk_wait_read<SCV_F2_NSTAGE - 1>();  // the stores that last read THIS tile have finished reading it
          __syncwarp();
          const long long d1 = ROWS_CLOCK();
          tb_rd += d1 - d0;
          const uint32_t row0 = smem_u32(stage) + lane * 64;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {  // 8 channels = one 16-byte chunk of every staging row per step
            const int col = c8 * 8;
            uint32_t r0[8], r1[8];
            tmem_ld8(taddr + col, r0);
            tmem_ld8(taddr + COUT + col, r1);
            tmem_ld_wait();
            if (c8 == 3) {
#pragma unroll
              for (int c = 0; c < 2 * COUT; c += 32) tmem_st32_zero(taddr + c);
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(&acc2_empty[slot]);
            }
            float bias[8], sc[8], sh[8];
            lds8(s_bias2 + col, bias);
            lds8(s_head + col, sc);
            lds8(s_head + COUT + col, sh);
            uint32_t pk0[4], pk1[4], pm[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a0 = __uint_as_float(r0[2 * j]) + bias[2 * j], a1 = __uint_as_float(r0[2 * j + 1]) + bias[2 * j + 1];
              float c0 = __uint_as_float(r1[2 * j]) + bias[2 * j], c1 = __uint_as_float(r1[2 * j + 1]) + bias[2 * j + 1];
              if (p.relu) a0 = fmaxf(a0, 0.f), a1 = fmaxf(a1, 0.f), c0 = fmaxf(c0, 0.f), c1 = fmaxf(c1, 0.f);
              const uint32_t m = max_bf16x2(pack_bf16x2(a0, a1), pack_bf16x2(c0, c1));
              pm[j] = max_bf16x2(m, __shfl_xor_sync(0xffffffffu, m, 1));
              pk0[j] = pack_bf16x2(fmaxf(fmaf(a0, sc[2 * j], sh[2 * j]), 0.f), fmaxf(fmaf(a1, sc[2 * j + 1], sh[2 * j + 1]), 0.f));
              pk1[j] = pack_bf16x2(fmaxf(fmaf(c0, sc[2 * j], sh[2 * j]), 0.f), fmaxf(fmaf(c1, sc[2 * j + 1], sh[2 * j + 1]), 0.f));
            }
            if (!(lane & 1)) {
              const uint32_t pp = lane >> 1;  // pooled pixel of this warp
              const uint32_t prow = smem_u32(stage) + 2 * 32 * 64 + pp * 64;
              sts128(prow + ((static_cast<uint32_t>(c8) ^ ((pp >> 1) & 3)) << 4), pm[0], pm[1], pm[2], pm[3]);
            }
            sts128(row0 + ((static_cast<uint32_t>(c8) ^ phase) << 4), pk0[0], pk0[1], pk0[2], pk0[3]);
            sts128(row0 + 32 * 64 + ((static_cast<uint32_t>(c8) ^ phase) << 4), pk1[0], pk1[1], pk1[2], pk1[3]);
          }
          const long long d2 = ROWS_CLOCK();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (p.out != nullptr) tma_store_4d(&tmOut, stage, p.out_choff, xw, y, sg.n);
            tma_store_4d(&tmPool, stage + 2 * 32 * 64, 0, xw >> 1, y >> 1, sg.n);
            bulk_commit();
          }
          tb_math += d2 - d1;
          tb_store += ROWS_CLOCK() - d2;
        }
      }
      opc += sg.npairs;
    }
    if constexpr (EPI2 == EPI_POOL_SKIP) {
      if (lane == 0) bulk_wait_read<0>();  // staging must stay valid until the last stores have read it
    }
#ifdef SCV_ROWS_PROF
    if ((p.dbg & 32) && blockIdx.x < 3 && lane == 0 && q == 0)
      printf("[fused prof] cta %d epilogue B group %d: total %lld cyc over %u pairs: wait acc2_full %lld, wait staging read %lld, ld+math+sts %lld, fence+store issue %lld\n",
             blockIdx.x, g, ROWS_CLOCK() - tb_begin, tb_n, tb_wait, tb_rd, tb_math, tb_store);
#endif
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while a neighbour may still store into its slabs / arrive on its barriers
  if (warp == 1) {
    tmem_dealloc(tmem_base, 512);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

}  // namespace scv
