// Implicit-GEMM 3x3 'same' convolution / 2x2-stride-2 transposed convolution on
// the 5th-gen tensor cores: TMA-fed, tcgen05.mma with the accumulator in TMEM,
// fused epilogues (bias+ReLU | +2x2 max-pool +skip affine | pixel-shuffle | 1x1 head).
//
// Replaces, per launch, the Keras op chain of utils/model_tools.py:178-186
// (Conv2D 'same' -> BatchNormalization -> ReLU; BN folded into W/bias at load),
// :281-286 (MaxPooling2D), :306-309 (Conv2DTranspose -> concatenate -> BN -> ReLU)
// and :405 / :443 (1x1 head conv; activation happens in the stitch kernel).
//
// GEMM view: M = pixels (128 per accumulator tile), N = output channels (BN per CTA),
// K = Cin * taps walked as (channel chunk of KC, tap, 16-wide k step) -- the same order in
// both kernels below, so a pixel's result does not depend on which kernel (or batch size)
// produced it.
//
//   conv_umma_kernel  one 128 x BN output tile per CTA; the A tile (a TW x TH x TN box of the
//                     NHWC activation tensor) is fetched by one 4-D TMA per (chunk, tap) with
//                     out-of-bounds zero fill supplying the per-tile 'same' padding.  Used for the
//                     deep layers (large Cin*Cout, few pixels).
//   conv_slab_kernel  persistent + weight-stationary, for the high-resolution layers (see below).
#pragma once
#include <cstdlib>

#include "ptx.cuh"

namespace scv {

enum { EPI_STORE = 0, EPI_POOL_SKIP = 1, EPI_CONVT = 2, EPI_HEAD = 3 };

struct ConvParams {
  int N, H, W;      // input spatial dims (== output dims for conv3x3; convT doubles H, W)
  int Cin;          // padded input channels, multiple of KC
  int ntaps;        // 9 (3x3) or 1 (GEMM / convT)
  int TW, TH, TN;   // M-tile box, TW*TH*TN == 128
  int tiles_x, tiles_y, tiles_n;
  int n_tiles_n;    // N-tiles (Ntotal / BN)
  int nstage;
  int relu;
  // main output (bf16 NHWC with channel pitch): EPI_STORE / EPI_CONVT / skip half of EPI_POOL_SKIP
  __nv_bfloat16* out;
  int out_pitch, out_choff;
  const float* bias;  // [Ntotal]
  const float* bias2; // fused two-conv kernel (conv_fused.cuh): bias of the second conv
  int Cout;           // EPI_CONVT: channels per (a,b) sub-pixel, Ntotal == 4*Cout
  // EPI_POOL_SKIP
  __nv_bfloat16* pool_out;
  int pool_pitch;
  const float* skip_s;  // [Ntotal] relu(s*v+t) goes to `out`
  const float* skip_t;
  // EPI_HEAD
  const float* head_w;  // [Ntotal][ncls]
  const float* head_b;  // [ncls]
  int ncls;
  float* logits;        // N*H*W*ncls fp32
  // persistent slab kernel
  int n_in_off;         // image offset added to the A-operand TMA coordinate (layer input is a slice of a larger tensor)
  int nslab;            // slab ring depth
  int dil;              // dilation of the 3x3 taps (tile kernels only; 1 everywhere but the siamese network's ASPP)
  int n_issuers;        // MMA issuer warps in use (1 or 2); 2 requires nslab % (2 * Cin/KC) == 0
  int num_m_tiles;      // tiles_x * tiles_y * N
  int dbg;              // timing experiments only (env SCV_ROWS_DBG; results are wrong when non-zero)
  int linear_out;       // experiment: the slab kernel's transposed-conv outputs leave the staging tile through ordinary
                        // coalesced stores instead of TMA stores (see epilogue_slab; no faster)
  // watchdog
  int* err;
  unsigned long long watchdog_ns;
};

// Programmatic dependent launch (experiment, SCV_PDL=1; off by default): the conv kernels can be launched with
// programmatic stream serialization so that a layer's prologue (barrier init, TMEM allocation, tensor-map prefetch,
// epilogue constants) runs while the previous layer drains; grid_dep_wait() sits right after the prologue, before the
// first read of an activation.  Measured on the full scene: 124.5-124.8 ms with it, 123.3-123.5 ms without
// (tools/r02_exp14.sh) -- with one 200 KB CTA per SM there is nothing for the early CTAs to overlap with.
inline bool pdl_enabled() {
  static const int on = [] {
    const char* s = getenv("SCV_PDL");
    return (s && *s) ? atoi(s) : 0;
  }();
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define SCV_LAUNCH_CHECK(expr)          \
  do {                                  \
    cudaError_t _le = (expr);           \
    if (_le != cudaSuccess) return _le; \
  } while (0)

constexpr int kConvThreads = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int kMaxHeadClasses = 16;

__host__ __device__ constexpr int conv_stage_bytes(int KC, int BN) { return 128 * KC * 2 + BN * KC * 2; }
__host__ __device__ inline size_t conv_smem_bytes(int KC, int BN, int nstage, int epi, int ncls) {
  size_t s = 1024 + static_cast<size_t>(nstage) * conv_stage_bytes(KC, BN);
  s += (2 * nstage + 1) * 8 + 16;  // barriers + tmem slot + abort flag
  s += BN * 4;                     // bias
  if (epi == EPI_POOL_SKIP) s += 2 * BN * 4;
  if (epi == EPI_HEAD) s += (BN * ncls + ncls) * 4;
  return s + 64;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ void pack16(const float (&v)[16], uint32_t (&pk)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
}
__device__ __forceinline__ void store_pk16(__nv_bfloat16* dst, const uint32_t (&pk)[8]) {
  uint4* p = reinterpret_cast<uint4*>(dst);
  p[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  p[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// 16 floats from 16-byte-aligned shared memory as 4 x LDS.128.  Explicit ld.shared: the constant area
// is carved out of the dynamic buffer with integer alignment arithmetic, after which the compiler no
// longer knows the address space and would emit (slower) generic loads.
__device__ __forceinline__ void lds16(const float* src, float (&dst)[16]) {
  const uint32_t a = smem_u32(src);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(dst[4 * j + 0]), "=f"(dst[4 * j + 1]), "=f"(dst[4 * j + 2]), "=f"(dst[4 * j + 3])
                 : "r"(a + 16 * j));
}

__device__ __forceinline__ void lds8(const float* src, float (&dst)[8]) {
  const uint32_t a = smem_u32(src);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(dst[4 * j + 0]), "=f"(dst[4 * j + 1]), "=f"(dst[4 * j + 2]), "=f"(dst[4 * j + 3])
                 : "r"(a + 16 * j));
}

__device__ __forceinline__ void lds32(const float* src, float (&dst)[32]) {
  const uint32_t a = smem_u32(src);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(dst[4 * j + 0]), "=f"(dst[4 * j + 1]), "=f"(dst[4 * j + 2]), "=f"(dst[4 * j + 3])
                 : "r"(a + 16 * j));
}

// Epilogue of one 128-pixel accumulator tile: this thread owns TMEM lane == pixel (x, y, n).
// `pair_w` is the lane distance of the vertical 2x2-pool partner (the M-tile's width in pixels).
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t taddr, int xx, int yy, int x, int y,
                                              int n, bool valid, int pair_w, int nb0, const float* s_bias,
                                              const float* s_extra) {
  const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
  float hacc[kMaxHeadClasses];
  if constexpr (EPI == EPI_HEAD) {
#pragma unroll
    for (int k = 0; k < kMaxHeadClasses; ++k) hacc[k] = 0.f;
  }
#pragma unroll 1
  for (int c = 0; c < BN; c += 16) {
    uint32_t raw[16];
    tmem_ld16(taddr + c, raw);
    tmem_ld_wait();
    float v[16], cst[16];
    lds16(s_bias + c, cst);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] = __uint_as_float(raw[j]) + cst[j];
      if (p.relu) v[j] = fmaxf(v[j], 0.f);
    }
    if constexpr (EPI == EPI_STORE) {
      uint32_t pk[8];
      pack16(v, pk);
      if (valid) store_pk16(p.out + pix * p.out_pitch + p.out_choff + nb0 + c, pk);
    } else if constexpr (EPI == EPI_POOL_SKIP) {
      // 2x2 max-pool on the bf16-rounded values (rounding is monotonic, so max commutes with it)
      uint32_t pk[8], m[8];
      pack16(v, pk);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t t = max_bf16x2(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 1));
        m[j] = max_bf16x2(t, __shfl_xor_sync(0xffffffffu, t, pair_w));
      }
      if (valid && !(xx & 1) && !(yy & 1)) {
        const size_t ppix = (static_cast<size_t>(n) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
        store_pk16(p.pool_out + ppix * p.pool_pitch + nb0 + c, m);
      }
      if (p.out != nullptr) {
        float sh[16];
        lds16(s_extra + c, cst);
        lds16(s_extra + BN + c, sh);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], cst[j], sh[j]), 0.f);
        pack16(v, pk);
        if (valid) store_pk16(p.out + pix * p.out_pitch + p.out_choff + nb0 + c, pk);
      }
    } else if constexpr (EPI == EPI_CONVT) {
      const int col = nb0 + c;
      const int g = col / p.Cout;
      const int o = col - g * p.Cout;
      const size_t opix =
          (static_cast<size_t>(n) * (2 * p.H) + (2 * y + (g >> 1))) * (2 * p.W) + (2 * x + (g & 1));
      uint32_t pk[8];
      pack16(v, pk);
      if (valid) store_pk16(p.out + opix * p.out_pitch + p.out_choff + o, pk);
    } else {  // EPI_HEAD
      if (p.ncls == 1) {  // sigmoid head: one dot product per pixel
        lds16(s_extra + c, cst);
        float a = hacc[0];
#pragma unroll
        for (int j = 0; j < 16; ++j) a = fmaf(v[j], cst[j], a);
        hacc[0] = a;
      } else
#pragma unroll
      for (int k = 0; k < kMaxHeadClasses; ++k) {
        if (k < p.ncls) {
          float a = hacc[k];
#pragma unroll
          for (int j = 0; j < 16; ++j) a = fmaf(v[j], s_extra[(c + j) * p.ncls + k], a);
          hacc[k] = a;
        }
      }
    }
  }
  if constexpr (EPI == EPI_HEAD) {
    if (valid) {
#pragma unroll
      for (int k = 0; k < kMaxHeadClasses; ++k)
        if (k < p.ncls) p.logits[pix * p.ncls + k] = hacc[k] + s_extra[BN * p.ncls + k];
    }
  }
}

// Slab-kernel epilogue for the bf16 tensor outputs (EPI_STORE / EPI_POOL_SKIP / EPI_CONVT): every epilogue
// warp owns 4 image rows x 8 pixels of the tile.  Results go, 32 channels at a time, through a per-warp
// swizzled shared-memory staging tile (32 pixel rows of 64 B) and leave with ONE TMA store per (warp,
// 32-channel block) -- fully coalesced global writes and no per-lane address arithmetic -- instead of
// 16-byte stores scattered at pixel pitch.  32-channel blocks never straddle two ConvT sub-pixels.
constexpr int kStageRowB = 64;                        // 32 bf16 channels
constexpr int kStageOutBytes = 32 * kStageRowB;       // 4 rows x 8 pixels
constexpr int kStagePoolBytes = 8 * kStageRowB;       // 2 rows x 4 pooled pixels
__host__ __device__ constexpr int slab_stage_warp_bytes(int epi) {
  return epi == EPI_HEAD ? 0 : (kStageOutBytes + (epi == EPI_POOL_SKIP ? kStagePoolBytes : 0));
}

// `release()` runs right after the LAST tcgen05.ld of the tile has completed: the accumulator is handed back to the
// issuer while this warp still does the last block's math, staging and store (early release; the kernels used to
// arrive only after the whole epilogue, i.e. also after the wait for the previous TMA store's read-out).
template <int BN, int EPI, typename Release>
__device__ __forceinline__ void epilogue_slab(const ConvParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmPool,
                                              uint32_t taddr, int lane, int q, int x0, int y0, int n, int nb0,
                                              const float* s_bias, const float* s_extra, uint8_t* stage, Release&& release) {
  const int xx = lane & 7, yl = lane >> 3;
  // swizzle phase of a 64-byte staging row = absolute smem address bits [7,9) (SWIZZLE_64B)
  const uint32_t phase = (lane >> 1) & 3;
  const int prow = (yl >> 1) * 4 + (xx >> 1);  // pooled pixel row in the pool staging tile
  const uint32_t pphase = (prow >> 1) & 3;
  const bool pool_leader = !(xx & 1) && !(yl & 1);
  const uint32_t out_row = smem_u32(stage) + lane * kStageRowB;
  const uint32_t pool_row = smem_u32(stage) + kStageOutBytes + prow * kStageRowB;
#pragma unroll 1
  for (int blk = 0; blk < BN / 32; ++blk) {
    // all loads of the block are issued before the first use (TMEM + shared-memory latency overlap)
    uint32_t pk[16], pm[16];
    const int col = blk * 32;
    uint32_t raw[32];
    tmem_ld32(taddr + col, raw);
    float v[32];
    lds32(s_bias + col, v);
    tmem_ld_wait();
    if (blk == BN / 32 - 1) release();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] += __uint_as_float(raw[j]);
      if (p.relu) v[j] = fmaxf(v[j], 0.f);
    }
    if constexpr (EPI == EPI_POOL_SKIP) {
      // 2x2 max-pool on the bf16-rounded values (rounding is monotonic, so max commutes with it)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t w = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        const uint32_t t = max_bf16x2(w, __shfl_xor_sync(0xffffffffu, w, 1));
        pm[j] = max_bf16x2(t, __shfl_xor_sync(0xffffffffu, t, 8));
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float sc[16], sh[16];
        lds16(s_extra + col + 16 * h, sc);
        lds16(s_extra + BN + col + 16 * h, sh);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 * h + j] = fmaxf(fmaf(v[16 * h + j], sc[j], sh[j]), 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    if constexpr (EPI == EPI_CONVT) {
      if (p.linear_out) {
        // Experiment (SCV_LSU_CONVT=1, off): transposed conv through the LSU instead of the TMA store this epilogue
        // waits for (ncu: 25 % of the stall samples on the staging-release wait).  Linear staging, read back 8 pixels
        // x 64 B per warp instruction, 128-bit stores.  Measured: dec0.up 6.2 vs 6.5 ms, dec1.up 4.6 vs 2.9 ms,
        // dec2.up 2.5 vs 1.7 ms -- the store path is not the limiter; these layers run at ~77 % of the write-heavy
        // HBM ceiling (tools/microbench/hbm_rw_mix.cu) and only fusing them into their consumer removes the traffic.
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c = (static_cast<uint32_t>(j) + phase) & 3;
          const uint32_t a0 = c == 0 ? pk[0] : (c == 1 ? pk[4] : (c == 2 ? pk[8] : pk[12]));
          const uint32_t a1 = c == 0 ? pk[1] : (c == 1 ? pk[5] : (c == 2 ? pk[9] : pk[13]));
          const uint32_t a2 = c == 0 ? pk[2] : (c == 1 ? pk[6] : (c == 2 ? pk[10] : pk[14]));
          const uint32_t a3 = c == 0 ? pk[3] : (c == 1 ? pk[7] : (c == 2 ? pk[11] : pk[15]));
          sts128(out_row + (c << 4), a0, a1, a2, a3);
        }
        __syncwarp();
        const int ch0 = nb0 + blk * 32;
        const int g = ch0 / p.Cout;  // sub-pixel (a, b) = (g >> 1, g & 1)
        const int sx = lane >> 2, c16 = (lane & 3) * 16;
        uint4 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = lds128(smem_u32(stage) + k * 512 + lane * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // staging row k * 8 + sx = tile pixel (x0 + sx, y0 + 4 q + k)
          const size_t oy = static_cast<size_t>(n) * (2 * p.H) + 2 * (y0 + q * 4 + k) + (g >> 1);
          const size_t ox = 2 * (x0 + sx) + (g & 1);
          uint8_t* gp = reinterpret_cast<uint8_t*>(p.out) + ((oy * (2 * p.W) + ox) * p.out_pitch + p.out_choff + (ch0 - g * p.Cout)) * 2 + c16;
          *reinterpret_cast<uint4*>(gp) = t[k];
        }
        continue;
      }
    }
    if (lane == 0) bulk_wait_read<0>();  // this warp's previous store has finished reading the staging tile
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(out_row + ((static_cast<uint32_t>(j) ^ phase) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    if constexpr (EPI == EPI_POOL_SKIP) {
      if (pool_leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128(pool_row + ((static_cast<uint32_t>(j) ^ pphase) << 4), pm[4 * j], pm[4 * j + 1], pm[4 * j + 2],
                 pm[4 * j + 3]);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const int ch0 = nb0 + blk * 32;
      if constexpr (EPI == EPI_CONVT) {
        const int g = ch0 / p.Cout;  // sub-pixel (a, b) = (g >> 1, g & 1); rows of (n, y) are merged in the map
        tma_store_5d(tmOut, stage, p.out_choff + (ch0 - g * p.Cout), g & 1, x0, g >> 1, n * p.H + y0 + q * 4);
      } else {
        tma_store_4d(tmOut, stage, p.out_choff + ch0, x0, y0 + q * 4, n);
        if constexpr (EPI == EPI_POOL_SKIP)
          tma_store_4d(tmPool, stage + kStageOutBytes, ch0, x0 >> 1, (y0 >> 1) + q * 2, n);
      }
      bulk_commit();
    }
  }
}

// Epilogue of the last decoder conv with the 1x1 head fused (EPI_HEAD), used by both kernels (same summation
// order, so logits do not depend on the kernel): 32 accumulator columns per
// step, all loads issued before first use; the common single-class (sigmoid) head is one fp32 dot product per
// pixel, other class counts take the generic path.
template <int BN>
__device__ __forceinline__ void epilogue_head(const ConvParams& p, uint32_t taddr, int xx, int yy, int x, int y,
                                                   int n, bool valid, int nb0, const float* s_bias,
                                                   const float* s_extra) {
  if (p.ncls != 1) {
    epilogue_tile<BN, EPI_HEAD>(p, taddr, xx, yy, x, y, n, valid, 8, nb0, s_bias, s_extra);
    return;
  }
  float acc = s_extra[BN];  // head bias
#pragma unroll 1
  for (int col = 0; col < BN; col += 32) {
    float part[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent FMA chains over the 32 columns of this step
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // 16 accumulator columns at a time (register pressure); the order of the sums is that
      uint32_t raw[16];            // of one 32-column pass
      tmem_ld16(taddr + col + 16 * h, raw);
      float v[16], w[16];
      lds16(s_bias + col + 16 * h, v);
      lds16(s_extra + col + 16 * h, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float t = v[j] + __uint_as_float(raw[j]);
        if (p.relu) t = fmaxf(t, 0.f);
        part[j & 3] = fmaf(t, w[j], part[j & 3]);
      }
    }
    acc += (part[0] + part[1]) + (part[2] + part[3]);
  }
  if (valid) p.logits[(static_cast<size_t>(n) * p.H + y) * p.W + x] = acc;
}

template <int BN, int EPI>
__device__ __forceinline__ void load_epilogue_consts(const ConvParams& p, int t, int nthreads, int nb0, float* s_bias,
                                                     float* s_extra) {
  for (int i = t; i < BN; i += nthreads) s_bias[i] = p.bias[nb0 + i];
  if constexpr (EPI == EPI_POOL_SKIP) {
    for (int i = t; i < BN; i += nthreads) {
      s_extra[i] = p.skip_s[nb0 + i];
      s_extra[BN + i] = p.skip_t[nb0 + i];
    }
  }
  if constexpr (EPI == EPI_HEAD) {
    for (int i = t; i < BN * p.ncls; i += nthreads) s_extra[i] = p.head_w[i];
    for (int i = t; i < p.ncls; i += nthreads) s_extra[BN * p.ncls + i] = p.head_b[i];
  }
}

// NOTE on the producer / MMA warps: the loops run warp-uniformly on all 32 lanes and only the
// instruction issue is predicated on elect.sync.  Guarding the whole loop with `lane == 0` makes
// every UTCHMMA / UTMALDG operand "divergent" for the compiler, which then wraps each issue in a
// ~30-instruction R2UR waterfall loop -- more cycles than a 128x32x16 MMA itself takes.

template <int KC, int BN, int EPI>
__global__ void __launch_bounds__(kConvThreads)
    conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const ConvParams p) {
  constexpr int A_BYTES = 128 * KC * 2;
  constexpr int B_BYTES = BN * KC * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ROW_BYTES = KC * 2;  // == swizzle span of both operands
  constexpr uint32_t IDESC = umma_idesc_bf16(128, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nstage = p.nstage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(nstage) * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + nstage;
  uint64_t* tmem_full_bar = empty_bar + nstage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
  float* s_extra = s_bias + BN;  // skip (s,t) or head (w,b)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which output tile
  const int n_tile = blockIdx.x % p.n_tiles_n;
  int m_tile = blockIdx.x / p.n_tiles_n;
  const int tx = m_tile % p.tiles_x;
  m_tile /= p.tiles_x;
  const int ty = m_tile % p.tiles_y;
  const int tn = m_tile / p.tiles_y;
  const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * p.TN;
  const int nb0 = n_tile * BN;

  const int chunks = p.Cin / KC;
  const int iters = p.ntaps * chunks;

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < nstage; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      *abort_flag = 0;
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  if (warp >= 2) load_epilogue_consts<BN, EPI>(p, threadIdx.x - 64, 128, nb0, s_bias, s_extra);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();  // programmatic dependent launch: see launch_pdl()
  grid_dep_wait();    // nothing above reads an activation; everything below may

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, elected issue) =====================
    int ch = 0, tap = 0;
    for (int it = 0; it < iters; ++it) {
      const int s = it % nstage;
      const uint32_t round = static_cast<uint32_t>(it / nstage);
      const bool ok = mbar_wait(&empty_bar[s], (round & 1) ^ 1, abort_flag, p.watchdog_ns);
      if (!__all_sync(0xffffffffu, ok)) break;
      int dy = 0, dx = 0;
      if (p.ntaps == 9) {
        dy = (tap / 3 - 1) * p.dil;
        dx = (tap % 3 - 1) * p.dil;
      }
      uint8_t* a_dst = tiles + static_cast<size_t>(s) * STAGE_BYTES;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_4d(a_dst, &tmA, &full_bar[s], ch * KC, x0 + dx, y0 + dy, n0 + p.n_in_off);
        tma_load_2d(a_dst + A_BYTES, &tmB, &full_bar[s], tap * p.Cin + ch * KC, nb0);
      }
      __syncwarp();
      if (++tap == p.ntaps) {
        tap = 0;
        ++ch;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected issue) =====================
    bool all_ok = true;
    for (int it = 0; it < iters; ++it) {
      const int s = it % nstage;
      const uint32_t round = static_cast<uint32_t>(it / nstage);
      const bool ok = mbar_wait(&full_bar[s], round & 1, abort_flag, p.watchdog_ns);
      if (!__all_sync(0xffffffffu, ok)) {
        all_ok = false;
        break;
      }
      tc_fence_after();
      const uint32_t a_addr = smem_u32(tiles + static_cast<size_t>(s) * STAGE_BYTES);
      const uint64_t da0 = umma_smem_desc(a_addr, ROW_BYTES);
      const uint64_t db0 = umma_smem_desc(a_addr + A_BYTES, ROW_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KC / 16; ++k)  // +32 B per k step == +2 in the descriptor's 16-byte address units
          umma_bf16_lo_acc(tmem_base, desc_lo(da0) + 2 * k, desc_lo(db0) + 2 * k, desc_hi(da0), desc_hi(db0), IDESC,
                           (it | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above retire
      }
      __syncwarp();
    }
    if (all_ok && elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // accumulator row == pixel within the M tile
    const int xx = r % p.TW;
    const int yy = (r / p.TW) % p.TH;
    const int nn = r / (p.TW * p.TH);
    const int x = x0 + xx, y = y0 + yy, n = n0 + nn;
    const bool valid = (x < p.W) && (y < p.H) && (n < p.N);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    const bool acc_ready = mbar_wait(tmem_full_bar, 0, abort_flag, p.watchdog_ns);
    if (__all_sync(0xffffffffu, acc_ready)) {  // warp-uniform: the epilogue uses .sync.aligned ops
      tc_fence_after();
      if constexpr (EPI == EPI_HEAD)
        epilogue_head<BN>(p, taddr, xx, yy, x, y, n, valid, nb0, s_bias, s_extra);
      else
        epilogue_tile<BN, EPI>(p, taddr, xx, yy, x, y, n, valid, p.TW, nb0, s_bias, s_extra);
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, BN);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent form of conv_umma_kernel for the deep / mid layers (weights too large to keep resident, K long):
// one CTA per SM walks the (M tile, N tile) work list with the SAME per-tile K order, so results are
// bit-identical to the one-tile-per-CTA kernel.  What it removes is the per-CTA fixed cost that kernel pays
// for every 128 x BN tile -- TMEM allocation, barrier init, a cold TMA pipeline, and an epilogue the tensor
// pipe waits for: here the TMA ring keeps running across tile boundaries and two TMEM accumulators
// (2 x BN <= 512 columns) alternate, so epilogue warpgroup g drains tile j while the issuer runs tile j+1.
// warp 0: TMA; warps 1..kPtileIssuers: MMA issuers taking turns stage by stage (an issuer's barrier waits,
// descriptor arithmetic and tcgen05.commit run while the others issue; the token keeps issue order -- and with
// it the summation order -- strictly sequential, as in conv_rows_kernel); then 2 epilogue warpgroups.
constexpr int kPtileIssuers = 2;
constexpr int kPtileFirstEpiWarp = 1 + kPtileIssuers;
constexpr int kPtileThreads = 32 * kPtileFirstEpiWarp + 256;
// floats of epilogue constants per warpgroup (bias + skip (s,t) / head (w,b)), 16-byte multiple
__host__ __device__ inline int kPtileConstStride(int BN, int epi, int ncls) {
  int n = BN;
  if (epi == EPI_POOL_SKIP) n += 2 * BN;
  if (epi == EPI_HEAD) n += BN * ncls + ncls;
  return (n + 3) & ~3;
}
__host__ __device__ inline size_t ptile_smem_bytes(int KC, int BN, int nstage, int epi, int ncls) {
  size_t s = 1024 + static_cast<size_t>(nstage) * conv_stage_bytes(KC, BN);
  s += (2 * nstage + 4 + kPtileIssuers) * 8 + 16;
  s += 2 * static_cast<size_t>(kPtileConstStride(BN, epi, ncls)) * 4;
  return s + 64;
}

template <int KC, int BN, int EPI>
__global__ void __launch_bounds__(kPtileThreads, 1)
    conv_ptile_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const ConvParams p) {
  constexpr int A_BYTES = 128 * KC * 2;
  constexpr int B_BYTES = BN * KC * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ROW_BYTES = KC * 2;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, BN);
  static_assert(2 * BN <= 512, "two accumulators must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nstage = p.nstage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(nstage) * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + nstage;
  uint64_t* acc_full = empty_bar + nstage;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* turn = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + kPtileIssuers);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunks = p.Cin / KC;
  const int iters = p.ntaps * chunks;
  const int total = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < nstage; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&acc_full[a], 1);
        mbar_init(&acc_empty[a], 128);
      }
      for (int i = 0; i < kPtileIssuers; ++i) mbar_init(&turn[i], 1);
      *abort_flag = 0;
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();  // programmatic dependent launch: see launch_pdl()
  grid_dep_wait();    // nothing above reads an activation; everything below may

  if (warp == 0) {
    // ===================== TMA producer: one uninterrupted stage stream over all tiles of this CTA ==========
    uint32_t s = 0, ph = 1;
    bool run = true;
    for (int w = blockIdx.x; run && w < total; w += gridDim.x) {
      const int n_tile = w % p.n_tiles_n;
      int m_tile = w / p.n_tiles_n;
      const int tx = m_tile % p.tiles_x;
      m_tile /= p.tiles_x;
      const int ty = m_tile % p.tiles_y;
      const int tn = m_tile / p.tiles_y;
      const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * p.TN, nb0 = n_tile * BN;
      int ch = 0, tap = 0;
      for (int it = 0; it < iters; ++it) {
        const bool ok = mbar_wait(&empty_bar[s], ph, abort_flag, p.watchdog_ns);
        if (!__all_sync(0xffffffffu, ok)) {
          run = false;
          break;
        }
        int dy = 0, dx = 0;
        if (p.ntaps == 9) {
          dy = (tap / 3 - 1) * p.dil;
          dx = (tap % 3 - 1) * p.dil;
        }
        uint8_t* a_dst = tiles + static_cast<size_t>(s) * STAGE_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_4d(a_dst, &tmA, &full_bar[s], ch * KC, x0 + dx, y0 + dy, n0 + p.n_in_off);
          tma_load_2d(a_dst + A_BYTES, &tmB, &full_bar[s], tap * p.Cin + ch * KC, nb0);
        }
        __syncwarp();
        if (++tap == p.ntaps) {
          tap = 0;
          ++ch;
        }
        if (++s == static_cast<uint32_t>(nstage)) s = 0, ph ^= 1;
      }
    }
  } else if (warp < kPtileFirstEpiWarp) {
    // ===================== MMA issuers: issuer k takes the CTA's stages k, k+ni, ... =====================
    const int ni = p.n_issuers;  // <= min(kPtileIssuers, nstage, 2 * iters): reuse distances cover the issuers in flight
    const int me = warp - 1;
    uint32_t s = 0, ph = 0, j = 0, nth = 0;
    int turn_of = 0;
    bool run = me < ni;
    for (int w = blockIdx.x; run && w < total; w += gridDim.x, ++j) {
      const uint32_t a = j & 1;
      const uint32_t tacc = tmem_base + a * BN;
      for (int it = 0; it < iters; ++it, ++turn_of) {
        if (turn_of == ni) turn_of = 0;
        const uint32_t sc = s, phc = ph;
        if (++s == static_cast<uint32_t>(nstage)) s = 0, ph ^= 1;
        if (turn_of != me) continue;
        // ---- everything that does not need the token
        if (it == 0) {  // first stage of a tile: its accumulator must have been drained
          const bool ok0 = mbar_wait(&acc_empty[a], ((j >> 1) & 1) ^ 1, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok0)) {
            run = false;
            break;
          }
        }
        const bool ok = mbar_wait(&full_bar[sc], phc, abort_flag, p.watchdog_ns);
        if (!__all_sync(0xffffffffu, ok)) {
          run = false;
          break;
        }
        const uint32_t a_addr = smem_u32(tiles + static_cast<size_t>(sc) * STAGE_BYTES);
        const uint64_t da0 = umma_smem_desc(a_addr, ROW_BYTES);
        const uint64_t db0 = umma_smem_desc(a_addr + A_BYTES, ROW_BYTES);
        // ---- the turn: stage g may only be issued after stage g-1
        if (ni > 1 && !(me == 0 && nth == 0)) {  // issuer 0 starts with the token; its k-th later turn waits for arrival k
          const uint32_t par = me == 0 ? ((nth - 1) & 1) : (nth & 1);
          const bool ok3 = mbar_wait(&turn[me], par, abort_flag, p.watchdog_ns);
          if (!__all_sync(0xffffffffu, ok3)) {
            run = false;
            break;
          }
        }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)  // +32 B per k step == +2 in the descriptor's 16-byte address units
            umma_bf16_lo_acc(tacc, desc_lo(da0) + 2 * k, desc_lo(db0) + 2 * k, desc_hi(da0), desc_hi(db0), IDESC,
                             (it | k) != 0 ? 1u : 0u);
          if (ni > 1) mbar_arrive(&turn[me + 1 == ni ? 0 : me + 1]);  // hand the token on before the commits
          umma_commit(&empty_bar[sc]);                                 // frees this smem stage once its MMAs retire
          if (it == iters - 1) umma_commit(&acc_full[a]);              // in-order pipe: the tile's last MMA retires last
        }
        __syncwarp();
        ++nth;
      }
    }
  } else {
    // ===================== epilogue: warpgroup g drains this CTA's tiles g, g+2, ... =====================
    const int g = (warp - kPtileFirstEpiWarp) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int xx = r % p.TW;
    const int yy = (r / p.TW) % p.TH;
    const int nn = r / (p.TW * p.TH);
    // each warpgroup keeps its own copy of the epilogue constants (the two groups may be on different N tiles)
    uint32_t j = g;
    int last_n_tile = -1;
    for (int w = blockIdx.x + g * gridDim.x; w < total; w += 2 * gridDim.x, j += 2) {
      const int n_tile = w % p.n_tiles_n;
      int m_tile = w / p.n_tiles_n;
      const int tx = m_tile % p.tiles_x;
      m_tile /= p.tiles_x;
      const int ty = m_tile % p.tiles_y;
      const int tn = m_tile / p.tiles_y;
      const int x = tx * p.TW + xx, y = ty * p.TH + yy, n = tn * p.TN + nn;
      const int nb0 = n_tile * BN;
      const bool valid = (x < p.W) && (y < p.H) && (n < p.N);
      float* c_bias = s_bias + g * kPtileConstStride(BN, EPI, p.ncls);
      float* c_extra = c_bias + BN;
      if (n_tile != last_n_tile) {  // this group's constants: named barrier over the group's 128 threads
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        load_epilogue_consts<BN, EPI>(p, (warp - kPtileFirstEpiWarp - 4 * g) * 32 + lane, 128, nb0, c_bias, c_extra);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        last_n_tile = n_tile;
      }
      const bool ready = mbar_wait(&acc_full[g], (j >> 1) & 1, abort_flag, p.watchdog_ns);
      if (!__all_sync(0xffffffffu, ready)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + g * BN + (static_cast<uint32_t>(q * 32) << 16);
      if constexpr (EPI == EPI_HEAD)
        epilogue_head<BN>(p, taddr, xx, yy, x, y, n, valid, nb0, c_bias, c_extra);
      else
        epilogue_tile<BN, EPI>(p, taddr, xx, yy, x, y, n, valid, p.TW, nb0, c_bias, c_extra);
      tc_fence_before();
      mbar_arrive(&acc_empty[g]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, 2 * BN);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent, weight-stationary variant for the high-resolution layers (small Cin*Cout, huge M).
//
// The tile kernel above re-fetches the A tile once per tap (9x L2->smem traffic) and spends a CTA
// launch + TMEM allocation per 128 pixels; at Cout = 32/64 that leaves the tensor pipe idle.  Here
//   * one CTA per SM loops over M tiles (8 wide x 16 tall pixels of one image),
//   * all taps of the layer's weights are TMA-loaded ONCE per CTA and stay in shared memory,
//   * per tile and channel chunk one halo SLAB ((8+2) x (16+2) pixels, zero-filled out of bounds) is
//     loaded; the 9 taps are 9 UMMA descriptors into the same slab: start address shifted by
//     (dy*SLAB_W + dx) pixel rows, 8-row group stride = SLAB_W pixel rows.  (Measured on B200:
//     the UMMA swizzle XOR is a function of the absolute shared-memory address, exactly like the
//     TMA write swizzle, so a descriptor may start at any pixel row of a TMA-written slab and use
//     any 16-byte-multiple group stride; the descriptor's base-offset field stays 0.)
//   * NACC TMEM accumulators and NACC epilogue warpgroups rotate over the tiles, so epilogues of
//     up to NACC-1 earlier tiles overlap the MMAs of the current one.
template <int NTAPS>
struct SlabGeom {
  static constexpr int HALO = NTAPS == 9 ? 1 : 0;
  static constexpr int W = 8 + 2 * HALO;
  static constexpr int H = 16 + 2 * HALO;
};
// NACC = number of TMEM accumulators == number of epilogue warpgroups (2 or 4; NACC*BN <= 512 columns).
__host__ __device__ constexpr bool slab_nacc_ok(int BN, int nacc) { return nacc * BN <= 512; }
// warp 0: TMA producer; warps 1..kSlabIssuers: MMA issuers (tiles alternate between them: the UMMA queue
// is only a few instructions deep, so a single issuer's per-tile barrier work would leave the tensor pipe
// idle between tiles); then 4 epilogue warps per accumulator.
// mbarrier parity waits are only meaningful one phase ahead, so every barrier must always be waited on by
// the SAME issuer in consecutive rounds: accumulators alternate (NACC even), and a slab slot returns to the
// same issuer only if nslab is a multiple of kSlabIssuers * chunks -- otherwise the host selects one issuer.
constexpr int kSlabIssuers = 2;
__host__ __device__ constexpr int slab_threads(int nacc) { return 32 * (1 + kSlabIssuers) + 128 * nacc; }

template <int KC, int BN, int EPI, int NTAPS, int NACC>
__global__ void __launch_bounds__(slab_threads(NACC), 1)
    conv_slab_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPool,
                     const ConvParams p) {
  constexpr int ROWB = KC * 2;             // bytes per pixel row of a slab == swizzle span
  constexpr int WT_BYTES = BN * KC * 2;    // one (chunk, tap) weight tile
  constexpr int SW = SlabGeom<NTAPS>::W, SH = SlabGeom<NTAPS>::H, HALO = SlabGeom<NTAPS>::HALO;
  constexpr int SLAB_BYTES = SW * SH * ROWB;
  constexpr int SLAB_STRIDE = (SLAB_BYTES + 1023) & ~1023;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // KC == 8 (first layer, 8 stored input channels): rows are 16 B, no swizzle; two taps form one K = 16 step
  // (the descriptor's leading-byte-offset is the distance between the two taps' windows), K = 9*8 -> 80.
  const int chunks = KC == 8 ? 1 : p.Cin / KC;
  const int nwt = KC == 8 ? 10 : NTAPS * chunks;
  const int nslab = p.nslab;
  uint8_t* w_smem = base;
  uint8_t* slabs = base + static_cast<size_t>(nwt) * WT_BYTES;
  uint8_t* staging = slabs + static_cast<size_t>(nslab) * SLAB_STRIDE;  // 1024-aligned
  constexpr int STAGE_TOTAL = 4 * NACC * slab_stage_warp_bytes(EPI);
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STAGE_TOTAL);
  uint64_t* w_full = bars;
  uint64_t* slab_full = bars + 1;
  uint64_t* slab_empty = slab_full + nslab;
  uint64_t* acc_full = slab_empty + nslab;
  uint64_t* acc_empty = acc_full + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NACC);
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
      for (int a = 0; a < NACC; ++a) {
        mbar_init(&acc_full[a], 1);
        mbar_init(&acc_empty[a], 128);
      }
      *abort_flag = 0;
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, NACC * BN);
    tmem_relinquish();
  }
  constexpr int kFirstEpiWarp = 1 + kSlabIssuers;
  if (warp >= kFirstEpiWarp)
    load_epilogue_consts<BN, EPI>(p, threadIdx.x - 32 * kFirstEpiWarp, 128 * NACC, nb0, s_bias, s_extra);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();  // programmatic dependent launch: see launch_pdl()
  grid_dep_wait();    // nothing above reads an activation; everything below may

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(nwt) * WT_BYTES);
      if constexpr (KC == 8) {
        for (int i = 0; i < 10; ++i)  // ten 8-wide K slices (nine taps + one all-zero slice), [BN rows][16 B] each
          tma_load_2d(w_smem + static_cast<size_t>(i) * WT_BYTES, &tmB, w_full, i * 8, nb0);
      } else {
        for (int ch = 0; ch < chunks; ++ch)
          for (int tap = 0; tap < NTAPS; ++tap)  // smem order [chunk][tap]; global K index = tap*Cin + ch*KC
            tma_load_2d(w_smem + static_cast<size_t>(ch * NTAPS + tap) * WT_BYTES, &tmB, w_full,
                        tap * p.Cin + ch * KC, nb0);
      }
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
          mbar_arrive_expect_tx(&slab_full[s], SLAB_BYTES);
          tma_load_4d(slabs + static_cast<size_t>(s) * SLAB_STRIDE, &tmA, &slab_full[s], ch * KC, tx * 8 - HALO,
                      ty * 16 - HALO, n + p.n_in_off);
        }
        __syncwarp();
      }
    }
  } else if (warp < kFirstEpiWarp) {
    // ===================== MMA issuers: issuer j takes local tiles t == j (mod kSlabIssuers) =====================
    const int nissue = p.n_issuers;
    bool run = (warp - 1) < nissue && __all_sync(0xffffffffu, mbar_wait(w_full, 0, abort_flag, p.watchdog_ns));
    uint32_t t = warp - 1;
    const uint32_t w_addr = smem_u32(w_smem);
    for (int m = m_first + static_cast<int>(t) * m_stride; run && m < p.num_m_tiles;
         m += m_stride * nissue, t += nissue) {
      uint32_t it = t * chunks;  // position of this tile's first slab in the producer's sequence
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
        if constexpr (KC == 8) {
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {  // tap pair (2j, 2j+1); the last pair re-reads tap 8 against zero weights
              const int t0 = 2 * j, t1 = j < 4 ? 2 * j + 1 : 8;
              const int o0 = (t0 / 3) * SW + t0 % 3, o1 = (t1 / 3) * SW + t1 % 3;  // window offsets in pixels
              const uint64_t da = umma_smem_desc_nosw(slab_addr + o0 * 16, (o1 - o0) * 16, SW * 16);
              const uint64_t db = umma_smem_desc_nosw(w_addr + t0 * WT_BYTES, WT_BYTES, 128);
              umma_bf16_lo_acc(tacc, desc_lo(da), desc_lo(db), desc_hi(da), desc_hi(db), IDESC, j != 0 ? 1u : 0u);
            }
            umma_commit(&slab_empty[s]);
          }
          __syncwarp();
          continue;
        }
        const uint64_t da0 = umma_smem_desc_sbo(slab_addr, ROWB, SW * ROWB);
        const uint64_t db0 = umma_smem_desc(w_addr + static_cast<uint32_t>(ch * NTAPS) * WT_BYTES, ROWB);
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < NTAPS; ++tap) {
            const int dy = NTAPS == 9 ? tap / 3 : 0;
            const int dx = NTAPS == 9 ? tap % 3 : 0;
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              // descriptor address fields are in 16-byte units: all offsets below are compile-time
              const uint32_t da = desc_lo(da0) + static_cast<uint32_t>(((dy * SW + dx) * ROWB + k * 32) >> 4);
              const uint32_t db = desc_lo(db0) + static_cast<uint32_t>((tap * WT_BYTES + k * 32) >> 4);
              umma_bf16_lo_acc(tacc, da, db, desc_hi(da0), desc_hi(db0), IDESC, (tap | k) != 0 ? 1u : (ch != 0 ? 1u : 0u));
            }
          }
          umma_commit(&slab_empty[s]);
        }
        __syncwarp();
      }
      if (run && elect_one()) umma_commit(&acc_full[a]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warpgroup g handles local tiles t == g (mod NACC) =====================
    const int g = (warp - kFirstEpiWarp) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int xx = r & 7, yy = r >> 3;
    uint32_t t = g;
    for (int m = m_first + g * m_stride; m < p.num_m_tiles; m += m_stride * NACC, t += NACC) {
      const int n = m / tiles_xy;
      const int rem = m - n * tiles_xy;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const bool ready = mbar_wait(&acc_full[g], (t / NACC) & 1, abort_flag, p.watchdog_ns);
      if (!__all_sync(0xffffffffu, ready)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + g * BN + (static_cast<uint32_t>(q * 32) << 16);
#if !defined(SCV_DBG_NO_EPILOGUE)  // timing experiment hook (results are wrong when defined)
      if constexpr (EPI == EPI_HEAD) {
        const int x = tx * 8 + xx, y = ty * 16 + yy;
        epilogue_head<BN>(p, taddr, xx, yy, x, y, n, (x < p.W) && (y < p.H), nb0, s_bias, s_extra);
      } else {
        epilogue_slab<BN, EPI>(p, &tmOut, &tmPool, taddr, lane, q, tx * 8, ty * 16, n, nb0, s_bias, s_extra,
                               staging + static_cast<size_t>(warp - kFirstEpiWarp) * slab_stage_warp_bytes(EPI), [&] {
                                 tc_fence_before();
                                 mbar_arrive(&acc_empty[g]);
                               });
      }
#endif
      if constexpr (EPI == EPI_HEAD) {
        tc_fence_before();
        mbar_arrive(&acc_empty[g]);
      }
#if defined(SCV_DBG_NO_EPILOGUE)
      else {
        tc_fence_before();
        mbar_arrive(&acc_empty[g]);
      }
#endif
    }
    if constexpr (EPI != EPI_HEAD) {
      if (lane == 0) bulk_wait_read<0>();  // staging must stay valid until the last stores have read it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, NACC * BN);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

__host__ __device__ inline int slab_stride_bytes(int KC, int ntaps) {
  const int sw = ntaps == 9 ? 10 : 8, sh = ntaps == 9 ? 18 : 16;
  return (sw * sh * KC * 2 + 1023) & ~1023;
}
__host__ __device__ inline size_t slab_stage_bytes(int epi, int nacc) {
  return static_cast<size_t>(4 * nacc) * slab_stage_warp_bytes(epi);
}
__host__ __device__ inline size_t slab_weight_bytes(int KC, int BN, int ntaps, int cin) {
  return KC == 8 ? static_cast<size_t>(10) * BN * 16 : static_cast<size_t>(ntaps) * cin * BN * 2;
}
__host__ __device__ inline size_t slab_smem_bytes(int KC, int BN, int ntaps, int cin, int nslab, int epi, int ncls,
                                                  int nacc) {
  size_t s = 1024 + slab_weight_bytes(KC, BN, ntaps, cin) +
             static_cast<size_t>(nslab) * slab_stride_bytes(KC, ntaps) + slab_stage_bytes(epi, nacc);
  s += (2 * nslab + 2 * 4 + 1) * 8 + 16;
  s += BN * 4;
  if (epi == EPI_POOL_SKIP) s += 2 * BN * 4;
  if (epi == EPI_HEAD) s += (BN * ncls + ncls) * 4;
  return s + 64;
}

// Host side -------------------------------------------------------------------
struct ConvLaunch {
  CUtensorMap tmA, tmB;
  CUtensorMap tmB2;           // fused two-conv kernel: weights of the second conv
  CUtensorMap tmOut, tmPool;  // slab kernel only: TMA-store maps of the bf16 outputs
  ConvParams p;
  int KC, BN, EPI;
  int slab;  // 6: conv_slab2_kernel (conv_slab2.cuh, CTA pairs), 5: conv_fused2_kernel (conv_fused.cuh; the next layer's launch is then -1 = folded into this one), 4: conv_slabw_kernel (conv_slabw.cuh), 3: conv_ptile_kernel (persistent tiles), 2: conv_rows_kernel (conv_rows.cuh), 1: conv_slab_kernel, 0: conv_umma_kernel
  int nacc;  // slab kernel: accumulators / epilogue warpgroups (2 or 4)
  int grid;
  size_t smem;
};

// Returns cudaSuccess or the launch error; cudaErrorInvalidValue for an unsupported (KC,BN,EPI).
cudaError_t conv_launch(const ConvLaunch& L, cudaStream_t stream);
// Sets the max-dynamic-smem attribute of every instantiation (once per device).
cudaError_t conv_init_attributes();
// conv_fused.cu
cudaError_t conv_fused_launch(const ConvLaunch& L, cudaStream_t stream);
cudaError_t conv_fused_init_attributes();
int conv_fused_max_clusters(size_t smem, int epi2);
// conv_rows.cu
cudaError_t conv_rows_launch(const ConvLaunch& L, cudaStream_t stream);
cudaError_t conv_rows_init_attributes();

}  // namespace scv
