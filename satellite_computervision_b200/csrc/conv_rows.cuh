// Row-streaming, tap-packed 3x3 convolution for the high-resolution / low-channel layers
// (Cout = 32 or 64, image width a multiple of 128): utils/model_tools.py:178-186 (Conv2D 'same' + BN + ReLU),
// :281-286 (MaxPooling2D), :307-309 (skip half of the post-concat BN + ReLU), :405 / :443 (1x1 head).
//
// Why a third kernel.  With both operands in shared memory a 128 x N x 16 UMMA is bound by the
// shared-memory operand fetch (~115 B/clk, tools/microbench/umma_rate.cu): the 4 KB pixel operand is
// re-read for every tap, so at N = Cout = 32 the tensor pipe cannot exceed 35 % (59 % at N = 64).
// Here the three VERTICAL taps share one pixel operand:
//   * an M tile is 128 consecutive pixels of ONE image row; a CTA streams down a 128-pixel column strip;
//   * input row r contributes to output rows r-1, r, r+1 (ky = 2, 1, 0) at the SAME lanes, so one UMMA with
//     B = [W(ky=2,kx) | W(ky=1,kx) | W(ky=0,kx)] (N = 3*Cout) accumulates into the three row accumulators,
//     which sit in adjacent TMEM columns (a ring of 512/Cout row accumulators);
//   * the horizontal taps are three descriptors into the same 130-pixel halo row (start address + kx rows).
// Pixel-operand reads drop 3x: N = 96 runs at ~62 cycles per 3 taps instead of 3 x 46.
// Accumulators are zeroed by the epilogue warps (tcgen05.st) when they release them, so every UMMA
// accumulates; a pixel's summation order is (ky, channel chunk, kx, k) -- identical to the other two
// kernels whenever Cin fits one chunk.
//
// Everything is scheduled in ROW PAIRS (one TMA box = two input rows, one accumulator hand-off = two output
// rows) and the UMMAs are issued by two warps taking turns: measured (SCV_ROWS_DBG=32), the barrier waits,
// descriptor arithmetic and tcgen05.commit of a single issuer cost ~1100 cycles per row during which the
// shallow UMMA queue runs dry.  An issuer does all of that for its pair while the other one issues, then
// waits for the turn token; issue order -- and with it the summation order -- stays strictly sequential.
#pragma once
#include <cstdio>

#include "conv_umma.cuh"

// Per-role cycle accounting (printf from CTA 0 when SCV_ROWS_DBG & 32): compile with -DSCV_ROWS_PROF.
#ifdef SCV_ROWS_PROF
#define ROWS_CLOCK() clock64()
#else
#define ROWS_CLOCK() 0ll
#endif

namespace scv {

#ifndef SCV_ROWS_EPI_GROUPS
#define SCV_ROWS_EPI_GROUPS 4
#endif
#ifndef SCV_ROWS_ISSUERS
#define SCV_ROWS_ISSUERS 3
#endif
constexpr int kRowsEpiGroups = SCV_ROWS_EPI_GROUPS;           // epilogue warpgroups (one output row pair each) at Cout 64
#ifndef SCV_ROWS_EPI_GROUPS32
#define SCV_ROWS_EPI_GROUPS32 4
#endif
// Cout 32 has 8 accumulator pairs and, with the 16-channel epilogue (<= 80 registers), room for a fifth epilogue group
// (768 threads).  Measured with 5: enc0.c2 8.02 vs 7.89 ms, dec0.c1 8.80 vs 8.67 ms -- no gain, because those two layers
// already move 5.0 / 6.0 TB/s through HBM (ncu: 1.34 GB in 268 us, 1.77 GB in 296 us per 63 chips), i.e. they sit at the
// roof of a write-heavy / read-heavy stream.  Left at 4.
// (the fused-head epilogue would spill at 80 registers and stays at 4 groups in any case; EPI_HEAD == 3)
__host__ __device__ constexpr int rows_epi_groups(int cout, int epi) {
  return (cout == 32 && epi != 3) ? SCV_ROWS_EPI_GROUPS32 : kRowsEpiGroups;
}
constexpr int kRowsIssuers = SCV_ROWS_ISSUERS;                // MMA issuer warps taking turns (max)
constexpr int kRowsFirstEpiWarp = 1 + kRowsIssuers;
__host__ __device__ constexpr int rows_threads(int cout, int epi) { return 32 * kRowsFirstEpiWarp + 128 * rows_epi_groups(cout, epi); }
constexpr int kRowsPx = 128;                                  // strip width == UMMA M
constexpr int kRowsSlabPx = kRowsPx + 2;                      // input row with its two halo pixels

// one slab = two input rows of one channel chunk
__host__ __device__ constexpr int rows_slab_stride(int KC) { return (2 * kRowsSlabPx * KC * 2 + 1023) & ~1023; }
// per epilogue warp: 2 rows x 32 pixels x 32 channels (+ 16 pooled pixels for the pooling epilogue)
__host__ __device__ constexpr int rows_stage_warp_bytes(int epi) {
  return epi == EPI_HEAD ? 0 : (2 * 32 * 64 + (epi == EPI_POOL_SKIP ? 16 * 64 : 0));
}
__host__ __device__ inline size_t rows_smem_bytes(int KC, int COUT, int cin, int nslab, int epi, int ncls) {
  size_t s = 1024 + static_cast<size_t>(9) * cin * COUT * 2 + static_cast<size_t>(nslab) * rows_slab_stride(KC) +
             static_cast<size_t>(4 * rows_epi_groups(COUT, epi)) * rows_stage_warp_bytes(epi);
  s += (1 + 2 * nslab + 2 * (256 / COUT) + kRowsIssuers) * 8 + 16;
  s += COUT * 4;
  if (epi == EPI_POOL_SKIP) s += 2 * COUT * 4;
  if (epi == EPI_HEAD) s += (COUT * ncls + ncls) * 4;
  return s + 64;
}

__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A CTA owns the row pairs [p0, p1) of the flattened (image, strip, row pair) index space; a segment is a
// maximal run inside one (image, strip).
struct RowSeg {
  int n, xs, y0, npairs;
};
__device__ __forceinline__ bool rows_next_seg(long long& pc, long long p1, int XS, int H2, RowSeg& s) {
  if (pc >= p1) return false;
  const long long strip = pc / H2;
  const int y2 = static_cast<int>(pc - strip * H2);
  const long long left = p1 - pc;
  s.npairs = left < static_cast<long long>(H2 - y2) ? static_cast<int>(left) : (H2 - y2);
  s.n = static_cast<int>(strip / XS);
  s.xs = static_cast<int>(strip - static_cast<long long>(s.n) * XS);
  s.y0 = 2 * y2;
  pc += s.npairs;
  return true;
}

// +bias, ReLU of 32 accumulator columns
__device__ __forceinline__ void rows_bias_relu(const uint32_t (&raw)[32], const float* s_bias, int relu, float (&v)[32]) {
  lds32(s_bias, v);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] += __uint_as_float(raw[j]);
    if (relu) v[j] = fmaxf(v[j], 0.f);
  }
}

// skip half of the decoder's post-concat BatchNorm + ReLU: v = relu(s*v + t)
__device__ __forceinline__ void rows_skip_affine(const float* s_scale, const float* s_shift, float (&v)[32]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float sc[16], sh[16];
    lds16(s_scale + 16 * h, sc);
    lds16(s_shift + 16 * h, sh);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[16 * h + j] = fmaxf(fmaf(v[16 * h + j], sc[j], sh[j]), 0.f);
  }
}

// Where the UMMAs of one input row land: up to three consecutive row accumulators, split in two pieces when
// they wrap around the ring.
struct RowsPiece {
  uint32_t d1, id1, d2, id2;  // TMEM column address + instruction descriptor of each piece (id2 == 0: no wrap)
  uint32_t b_off;             // first weight block (ky = 2 - block)
  uint32_t b_wrap;            // weight blocks consumed by the first piece
};

// input row j (local to the segment, j = 0 is image row y0-1) feeds output rows i = j-2 (ky 2), j-1 (ky 1),
// j (ky 0), clipped to the segment's 2*npairs output rows; `opc` = output pairs of this CTA before the segment
template <int COUT>
__device__ __forceinline__ RowsPiece rows_piece(uint32_t tmem_base, int j, int npairs, uint32_t opc) {
  constexpr int R = 512 / COUT;
  RowsPiece r;
  const int i_lo = j >= 2 ? j - 2 : 0;
  const int i_hi = j < 2 * npairs ? j : 2 * npairs - 1;
  const int nblk = i_hi - i_lo + 1;
  const uint32_t slot_lo = (2 * opc + i_lo) % R;
  const int n1 = nblk < static_cast<int>(R - slot_lo) ? nblk : static_cast<int>(R - slot_lo);
  const int n2 = nblk - n1;
  r.d1 = tmem_base + slot_lo * COUT;
  r.id1 = umma_idesc_bf16(128, n1 * COUT);
  r.d2 = tmem_base;
  r.id2 = n2 > 0 ? umma_idesc_bf16(128, n2 * COUT) : 0u;
  r.b_off = static_cast<uint32_t>(2 - (j - i_lo));  // in weight blocks; scaled by the caller
  r.b_wrap = static_cast<uint32_t>(n1);
  return r;
}

// The 3 * KC/16 UMMAs of one input row and one channel chunk (issued by one elected thread).  Descriptors are
// handled as 32-bit low words (14-bit address field + constants) with one shared high word.
template <int KC, int COUT>
__device__ __forceinline__ void rows_issue_row(uint32_t da0, uint32_t db0, uint32_t desc_hi, uint32_t d1, uint32_t id1,
                                               uint32_t d2, uint32_t id2, uint32_t b_off, uint32_t b_wrap) {
  constexpr int ROWB = KC * 2, WT_BLOCK = COUT * ROWB;
  const uint32_t dbr = db0 + b_off * (WT_BLOCK >> 4);
  if (id2 == 0) {  // common case: one piece
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int k = 0; k < KC / 16; ++k)
        umma_bf16_lo(d1, da0 + ((kx * ROWB + k * 32) >> 4), dbr + ((kx * 3 * WT_BLOCK + k * 32) >> 4), desc_hi, id1);
    }
  } else {         // the row accumulators wrap around the TMEM ring: two pieces per (kx, k)
    const uint32_t wrap = b_wrap * (WT_BLOCK >> 4);
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int k = 0; k < KC / 16; ++k) {
        const uint32_t da = da0 + ((kx * ROWB + k * 32) >> 4);
        const uint32_t db = dbr + ((kx * 3 * WT_BLOCK + k * 32) >> 4);
        umma_bf16_lo(d1, da, db, desc_hi, id1);
        umma_bf16_lo(d2, da, db + wrap, desc_hi, id2);
      }
    }
  }
}

template <int KC, int COUT, int EPI>
__global__ void __launch_bounds__(rows_threads(COUT, EPI), 1)
    conv_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPool,
                     const ConvParams p) {
  constexpr int ROWB = KC * 2;               // bytes per pixel row of a slab == swizzle span
  constexpr int WT_BLOCK = COUT * ROWB;      // one (chunk, kx, ky) weight tile
  constexpr int ROW_BYTES = kRowsSlabPx * ROWB;
  constexpr int SLAB_BYTES = 2 * ROW_BYTES;
  constexpr int SLAB_STRIDE = rows_slab_stride(KC);
  constexpr int R = 512 / COUT;              // row accumulators in TMEM
  constexpr int RP = R / 2;                  // ... handed over in pairs
  constexpr int NG = rows_epi_groups(COUT, EPI);
  constexpr int NI = kRowsIssuers;
  constexpr int STAGE_W = rows_stage_warp_bytes(EPI);
  static_assert(COUT == 32 || COUT == 64, "row kernel: Cout 32 or 64");
  // mbarrier parity waits are only meaningful one phase ahead: whoever waits for use k+1 of a barrier must not
  // start before use k has completed.  Waiters of one kind take their pairs in order and are at most NG (NI)
  // pairs apart, so a reuse distance of RP >= NG (NI) pairs is enough.
  static_assert(RP >= NG && RP >= NI, "accumulator reuse distance must cover the waiters in flight");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int chunks = p.Cin / KC;
  const int nslab = p.nslab;
  uint8_t* w_smem = base;  // [chunk][kx][ky = 2,1,0][COUT rows][KC]
  uint8_t* slabs = base + static_cast<size_t>(chunks) * 9 * WT_BLOCK;
  uint8_t* staging = slabs + static_cast<size_t>(nslab) * SLAB_STRIDE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 4 * NG * STAGE_W);
  uint64_t* w_full = bars;
  uint64_t* slab_full = bars + 1;
  uint64_t* slab_empty = slab_full + nslab;
  uint64_t* acc_full = slab_empty + nslab;   // per pair slot
  uint64_t* acc_empty = acc_full + RP;
  uint64_t* turn = acc_empty + RP;           // issue token, one per issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + NI);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
  float* s_extra = s_bias + COUT;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int XS = (p.W + kRowsPx - 1) / kRowsPx, H2 = p.H >> 1;  // the last strip may be partial: TMA zero-fills its
                                                                // loads and clips its stores at the image border
  const long long P = static_cast<long long>(p.N) * XS * H2;
  const long long p0 = P * blockIdx.x / gridDim.x, p1 = P * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI != EPI_HEAD) tma_prefetch_desc(&tmOut);
    if constexpr (EPI == EPI_POOL_SKIP) tma_prefetch_desc(&tmPool);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int s = 0; s < nslab; ++s) {
        mbar_init(&slab_full[s], 1);
        mbar_init(&slab_empty[s], 1);
      }
      for (int a = 0; a < RP; ++a) {
        mbar_init(&acc_full[a], 1);
        mbar_init(&acc_empty[a], 128);
      }
      for (int i = 0; i < NI; ++i) mbar_init(&turn[i], 1);
      *abort_flag = 0;
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp >= kRowsFirstEpiWarp)
    load_epilogue_consts<COUT, EPI>(p, threadIdx.x - 32 * kRowsFirstEpiWarp, 128 * NG, 0, s_bias, s_extra);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kRowsFirstEpiWarp && warp < kRowsFirstEpiWarp + 4) {  // one warp per TMEM lane quadrant zeroes all 512 columns
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 512; c += 32) tmem_st32_zero(t0 + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  grid_dep_launch();  // programmatic dependent launch: see launch_pdl()
  grid_dep_wait();    // nothing above reads an activation; everything below may

  if (warp == 0) {
    // ===================== TMA producer: input row pairs (y0-1+2v, y0+2v), v = 0 .. npairs =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(chunks) * 9 * WT_BLOCK);
      for (int ch = 0; ch < chunks; ++ch)
        for (int kx = 0; kx < 3; ++kx)
          for (int b = 0; b < 3; ++b)  // block b holds ky = 2 - b; global K index = tap*Cin + ch*KC, tap = ky*3 + kx
            tma_load_2d(w_smem + static_cast<size_t>((ch * 3 + kx) * 3 + b) * WT_BLOCK, &tmB, w_full,
                        ((2 - b) * 3 + kx) * p.Cin + ch * KC, 0);
    }
    __syncwarp();
    uint32_t s = 0, ph = 1;  // slab slot and the parity that means "free"
    long long pc = p0;
    RowSeg sg;
    bool run = true;
    long long t_w = 0, t_begin = ROWS_CLOCK();
    while (run && rows_next_seg(pc, p1, XS, H2, sg)) {
      for (int v = 0; run && v <= sg.npairs; ++v) {
        for (int ch = 0; ch < chunks; ++ch) {
          const long long tw0 = ROWS_CLOCK();
          const bool ok = mbar_wait(&slab_empty[s], ph, abort_flag, p.watchdog_ns);
          t_w += ROWS_CLOCK() - tw0;
          if (!__all_sync(0xffffffffu, ok)) {
            run = false;
            break;
          }
          if (p.dbg & 4) {  // timing experiment: no A loads (stale shared memory)
            if (elect_one()) mbar_arrive(&slab_full[s]);
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&slab_full[s], SLAB_BYTES);
            // rows -1 and H, pixels -1 and W are out of bounds: zero filled == the tile's own 'same' padding
            tma_load_4d(slabs + static_cast<size_t>(s) * SLAB_STRIDE, &tmA, &slab_full[s], ch * KC, sg.xs * kRowsPx - 1,
                        sg.y0 - 1 + 2 * v, sg.n + p.n_in_off);
          }
          __syncwarp();
          if (++s == static_cast<uint32_t>(nslab)) s = 0, ph ^= 1;
        }
      }
    }
#ifdef SCV_ROWS_PROF
    if ((p.dbg & 32) && blockIdx.x == 0 && lane == 0)
      printf("[rows prof] producer: total %lld cyc, wait slab_empty %lld\n", ROWS_CLOCK() - t_begin, t_w);
#endif
  } else if (warp < kRowsFirstEpiWarp) {
    // ===================== MMA issuers: issuer k takes the CTA's input pairs k, k+ni, ... =====================
    const int ni = p.n_issuers;  // <= min(kRowsIssuers, nslab / chunks, RP), see the host planner
    const int me = warp - 1;
    bool run = me < ni && __all_sync(0xffffffffu, mbar_wait(w_full, 0, abort_flag, p.watchdog_ns));
    tc_fence_after();
    const uint32_t w_desc = static_cast<uint32_t>(umma_smem_desc(smem_u32(w_smem), ROWB) & 0xffffffffu);
    const uint32_t desc_hi = static_cast<uint32_t>(umma_smem_desc(0, ROWB) >> 32);
    const uint32_t slab0_desc = static_cast<uint32_t>(umma_smem_desc(smem_u32(slabs), ROWB) & 0xffffffffu);
    uint32_t nth = 0;    // pairs this issuer has issued
    int turn_of = 0;     // t % ni without the division
    bool first_pair = true;
    uint32_t s0 = static_cast<uint32_t>(me * chunks) % nslab, ph0 = (static_cast<uint32_t>(me * chunks) / nslab) & 1;
    uint32_t t = 0;      // running input-pair counter of this CTA (all issuers count all pairs)
    uint32_t opc = 0;    // running output-pair counter at the start of the current segment
    long long t_wait = 0, t_turn = 0, t_issue = 0, t_mma = 0, t_hand = 0, t_begin = ROWS_CLOCK();
    long long pc = p0;
    RowSeg sg;
    while (run && rows_next_seg(pc, p1, XS, H2, sg)) {
      for (int v = 0; run && v <= sg.npairs; ++v, ++t, ++turn_of) {
        if (turn_of == ni) turn_of = 0;
        if (turn_of != me) continue;
        // slabs of this pair: the producer's sequence position is t*chunks
        if (!first_pair) {
          s0 += static_cast<uint32_t>(ni * chunks);
          while (s0 >= static_cast<uint32_t>(nslab)) s0 -= nslab, ph0 ^= 1;
        }
        first_pair = false;
        const long long tw0 = ROWS_CLOCK();
        // ---- (1) everything that does not need the turn token
        if (v < sg.npairs) {  // output pair v enters the ring: its accumulators must have been drained and zeroed
          const uint32_t op = opc + v;
          const bool ok = mbar_wait(&acc_empty[op % RP], ((op / RP) & 1) ^ 1, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok)) {
            run = false;
            break;
          }
        }
        // input row j = 2v + rr feeds output rows i = j-2 (ky 2), j-1 (ky 1), j (ky 0), clipped to the segment
        const RowsPiece pc0 = rows_piece<COUT>(tmem_base, 2 * v, sg.npairs, opc);
        const RowsPiece pc1 = rows_piece<COUT>(tmem_base, 2 * v + 1, sg.npairs, opc);
        // this pair's slabs: [s0, s0 + chunks) modulo nslab, parity ph0 (flips at the wrap)
        {
          uint32_t sc = s0, phc = ph0;
          for (int ch = 0; ch < chunks; ++ch) {
            const bool ok2 = mbar_wait(&slab_full[sc], phc, abort_flag, p.watchdog_ns);
            if (!__all_sync(0xffffffffu, ok2)) run = false;
            if (++sc == static_cast<uint32_t>(nslab)) sc = 0, phc ^= 1;
          }
        }
        if (!run) break;
        const long long tw1 = ROWS_CLOCK();
        // ---- (2) the turn: pair t may only be issued after pair t-1 (strict issue order == summation order)
        if (ni > 1) {
          const uint32_t par = me == 0 ? ((nth & 1) ^ 1) : (nth & 1);
          const bool ok3 = mbar_wait(&turn[me], par, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok3)) {
            run = false;
            break;
          }
        }
        tc_fence_after();
        const long long tw2 = ROWS_CLOCK();
#ifdef SCV_ROWS_PROF
        if (t > 0) t_hand += tw2 - *reinterpret_cast<volatile long long*>(s_extra + 64);
#endif
        // ---- (3) issue
        if (elect_one()) {
          uint32_t sc = s0;
          for (int ch = 0; ch < chunks; ++ch) {
            const uint32_t da0 = slab0_desc + sc * static_cast<uint32_t>(SLAB_STRIDE >> 4);
            if (++sc == static_cast<uint32_t>(nslab)) sc = 0;
            const uint32_t db0 = w_desc + static_cast<uint32_t>((ch * 9 * WT_BLOCK) >> 4);
            if (!(p.dbg & 8)) {  // (timing experiment: no MMAs)
              rows_issue_row<KC, COUT>(da0, db0, desc_hi, pc0.d1, pc0.id1, pc0.d2, pc0.id2, pc0.b_off, pc0.b_wrap);
              rows_issue_row<KC, COUT>(da0 + (ROW_BYTES >> 4), db0, desc_hi, pc1.d1, pc1.id1, pc1.d2, pc1.id2, pc1.b_off,
                                       pc1.b_wrap);
            }
          }
#ifdef SCV_ROWS_PROF
          t_mma += clock64() - tw2;
          *reinterpret_cast<volatile long long*>(s_extra + 64) = clock64();  // hand-off timestamp (profiling only)
#endif
          if (ni > 1) mbar_arrive(&turn[me + 1 == ni ? 0 : me + 1]);  // hand the token on before the (slow) commits
          sc = s0;
          for (int ch = 0; ch < chunks; ++ch) {
            umma_commit(&slab_empty[sc]);
            if (++sc == static_cast<uint32_t>(nslab)) sc = 0;
          }
          if (v >= 1) umma_commit(&acc_full[(opc + v - 1) % RP]);  // output pair v-1 is complete
        }
        __syncwarp();
        ++nth;
        const long long tw3 = ROWS_CLOCK();
        t_wait += tw1 - tw0;
        t_turn += tw2 - tw1;
        t_issue += tw3 - tw2;
      }
      opc += sg.npairs;
    }
#ifdef SCV_ROWS_PROF
    if ((p.dbg & 32) && blockIdx.x == 0 && lane == 0 && me < ni)
      printf("[rows prof] issuer %d: total %lld cyc, waits+setup %lld, turn %lld, issue+commit %lld (mma issue %lld), hand-off latency %lld, own pairs %u, out pairs %u\n", me,
             ROWS_CLOCK() - t_begin, t_wait, t_turn, t_issue, t_mma, t_hand, nth, opc);
#endif
  } else {
    // ===================== epilogue: group g takes output row pairs g, g+NG, ... =====================
    const int ew = warp - kRowsFirstEpiWarp;
    const int g = ew >> 2;
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* stage = staging + static_cast<size_t>(ew) * STAGE_W;
    const uint32_t phase = (lane >> 1) & 3;  // SWIZZLE_64B phase of staging rows `lane` and `32 + lane`
    uint32_t opc = 0;
    long long pc = p0;
    RowSeg sg;
    bool run = true;
    long long t_w = 0, t_st = 0, t_ld = 0, t_zero = 0, t_fence = 0, t_begin = ROWS_CLOCK();
    while (run && rows_next_seg(pc, p1, XS, H2, sg)) {
      for (int u = 0; u < sg.npairs; ++u) {
        const uint32_t op = opc + u;
        if (op % NG != static_cast<uint32_t>(g)) continue;
        const uint32_t slot = op % RP;
        const long long te0 = ROWS_CLOCK();
        const bool ready = mbar_wait(&acc_full[slot], (op / RP) & 1, abort_flag, p.watchdog_ns);
        t_w += ROWS_CLOCK() - te0;
        if (!__all_sync(0xffffffffu, ready)) {
          run = false;
          break;
        }
        tc_fence_after();
        const uint32_t taddr = tq + slot * (2 * COUT);  // row y at +0, row y+1 at +COUT
        const int xw = sg.xs * kRowsPx + q * 32;         // first pixel of this warp
        const int y = sg.y0 + 2 * u;
        if constexpr (EPI == EPI_HEAD) {
          if (!(p.dbg & 1)) {
            epilogue_head<COUT>(p, taddr, 0, 0, xw + lane, y, sg.n, xw + lane < p.W, 0, s_bias, s_extra);
            epilogue_head<COUT>(p, taddr + COUT, 0, 0, xw + lane, y + 1, sg.n, xw + lane < p.W, 0, s_bias, s_extra);
          }
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int c = 0; c < 2 * COUT; c += 32) tmem_st32_zero(taddr + c);
            tmem_st_wait();
          }
          tc_fence_before();
          mbar_arrive(&acc_empty[slot]);
        } else {
#pragma unroll
          for (int b = 0; b < COUT / 32; ++b) {  // one staging tile / TMA store per 32-channel block
            const long long ts0 = ROWS_CLOCK();
            if (lane == 0) bulk_wait_read<0>();  // the previous stores have finished reading the staging tiles
            __syncwarp();
            t_st += ROWS_CLOCK() - ts0;
            const uint32_t row0 = smem_u32(stage) + lane * 64;
            // 16 channels at a time: two row halves (2 x 16 accumulator columns) + their constants stay in registers;
            // the 32-column form spilled 116-144 B per thread in the pooling epilogue (ptxas -v) -- local-memory traffic
            // in the stage that bounds the K-short 384 x 384 layers (profiles/r02_store_path.md, ablation table)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int col = b * 32 + h * 16;
              uint32_t r0[16], r1[16];
              const long long tl0 = ROWS_CLOCK();
              tmem_ld16(taddr + col, r0);
              tmem_ld16(taddr + COUT + col, r1);
              tmem_ld_wait();
              const long long tl1 = ROWS_CLOCK();
              t_ld += tl1 - tl0;
              if (b == COUT / 32 - 1 && h == 1) {  // everything read: zero both accumulators and hand them back
                if (!(p.dbg & 2)) {
#pragma unroll
                  for (int c = 0; c < 2 * COUT; c += 32) tmem_st32_zero(taddr + c);
                  tmem_st_wait();
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[slot]);
              }
              t_zero += ROWS_CLOCK() - tl1;
              if (p.dbg & 1) continue;
              float bias[16];
              lds16(s_bias + col, bias);
              uint32_t pk0[8], pk1[8];  // packed bf16 pairs of rows y and y+1
              if constexpr (EPI == EPI_POOL_SKIP) {
                float sc[16], sh[16];
                lds16(s_extra + col, sc);
                lds16(s_extra + COUT + col, sh);
                uint32_t pm[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  // conv output relu(acc + bias) feeds the pool (bf16-rounded; rounding is monotonic, so max commutes
                  // with it) and, through the decoder's post-concat BN + ReLU, the skip half of the concat buffer
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
                  const uint32_t pph = (pp >> 1) & 3;
                  sts128(prow + ((static_cast<uint32_t>(2 * h) ^ pph) << 4), pm[0], pm[1], pm[2], pm[3]);
                  sts128(prow + ((static_cast<uint32_t>(2 * h + 1) ^ pph) << 4), pm[4], pm[5], pm[6], pm[7]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  pk0[j] = pack_bf16x2(__uint_as_float(r0[2 * j]) + bias[2 * j], __uint_as_float(r0[2 * j + 1]) + bias[2 * j + 1]);
                  pk1[j] = pack_bf16x2(__uint_as_float(r1[2 * j]) + bias[2 * j], __uint_as_float(r1[2 * j + 1]) + bias[2 * j + 1]);
                  if (p.relu) {  // relu(round(x)) == round(relu(x)): one packed max per pair instead of two fp32 ones
                    pk0[j] = max_bf16x2(pk0[j], 0u);
                    pk1[j] = max_bf16x2(pk1[j], 0u);
                  }
                }
              }
              sts128(row0 + ((static_cast<uint32_t>(2 * h) ^ phase) << 4), pk0[0], pk0[1], pk0[2], pk0[3]);
              sts128(row0 + ((static_cast<uint32_t>(2 * h + 1) ^ phase) << 4), pk0[4], pk0[5], pk0[6], pk0[7]);
              sts128(row0 + 32 * 64 + ((static_cast<uint32_t>(2 * h) ^ phase) << 4), pk1[0], pk1[1], pk1[2], pk1[3]);
              sts128(row0 + 32 * 64 + ((static_cast<uint32_t>(2 * h + 1) ^ phase) << 4), pk1[4], pk1[5], pk1[6], pk1[7]);
            }
            if (p.dbg & 1) continue;
            const long long tf0 = ROWS_CLOCK();
            fence_proxy_async();
            __syncwarp();
            t_fence += ROWS_CLOCK() - tf0;
            if (lane == 0 && !(p.dbg & 16) && xw < p.W) {  // (a warp wholly beyond a partial strip's edge stores nothing)
              if (p.out != nullptr) tma_store_4d(&tmOut, stage, p.out_choff + b * 32, xw, y, sg.n);
              if constexpr (EPI == EPI_POOL_SKIP) tma_store_4d(&tmPool, stage + 2 * 32 * 64, b * 32, xw >> 1, y >> 1, sg.n);
              bulk_commit();
            }
          }
        }
      }
      opc += sg.npairs;
    }
#ifdef SCV_ROWS_PROF
    if ((p.dbg & 32) && blockIdx.x == 0 && lane == 0 && q == 0)
      printf("[rows prof] epilogue group %d: total %lld cyc, wait acc_full %lld, wait staging free %lld, tmem ld %lld, zero+release %lld, fence %lld\n", g,
             ROWS_CLOCK() - t_begin, t_w, t_st, t_ld, t_zero, t_fence);
#endif
    if constexpr (EPI != EPI_HEAD) {
      if (lane == 0) bulk_wait_read<0>();  // staging must stay valid until the last stores have read it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, 512);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

}  // namespace scv
