// Implicit-GEMM 3x3 'same' convolution / 2x2-stride-2 transposed convolution on
// the 5th-gen tensor cores: TMA-fed, tcgen05.mma with the accumulator in TMEM,
// fused epilogues (bias+ReLU | +2x2 max-pool +skip affine | pixel-shuffle | 1x1 head).
//
// Replaces, per launch, the Keras op chain of utils/model_tools.py:178-186
// (Conv2D 'same' -> BatchNormalization -> ReLU; BN folded into W/bias at load),
// :281-286 (MaxPooling2D), :306-309 (Conv2DTranspose -> concatenate -> BN -> ReLU)
// and :405 / :443 (1x1 head conv; activation happens in the stitch kernel).
//
// GEMM view: M = pixels (128 per CTA: a TW x TH x TN box of the NHWC activation
// tensor, fetched by one 4-D TMA per (tap, channel chunk) with out-of-bounds zero
// fill supplying the per-tile 'same' padding), N = output channels (BN per CTA),
// K = taps * Cin walked as (tap, chunk of KC channels).
#pragma once
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
  // watchdog
  int* err;
  unsigned long long watchdog_ns;
};

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
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* dst, const float (&v)[16]) {
  uint4 lo, hi;
  lo.x = pack_bf16x2(v[0], v[1]);
  lo.y = pack_bf16x2(v[2], v[3]);
  lo.z = pack_bf16x2(v[4], v[5]);
  lo.w = pack_bf16x2(v[6], v[7]);
  hi.x = pack_bf16x2(v[8], v[9]);
  hi.y = pack_bf16x2(v[10], v[11]);
  hi.z = pack_bf16x2(v[12], v[13]);
  hi.w = pack_bf16x2(v[14], v[15]);
  uint4* p = reinterpret_cast<uint4*>(dst);
  p[0] = lo;
  p[1] = hi;
}

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
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
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
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    for (int i = t; i < BN; i += 128) s_bias[i] = p.bias[nb0 + i];
    if constexpr (EPI == EPI_POOL_SKIP) {
      for (int i = t; i < BN; i += 128) {
        s_extra[i] = p.skip_s[nb0 + i];
        s_extra[BN + i] = p.skip_t[nb0 + i];
      }
    }
    if constexpr (EPI == EPI_HEAD) {
      for (int i = t; i < BN * p.ncls; i += 128) s_extra[i] = p.head_w[i];
      for (int i = t; i < p.ncls; i += 128) s_extra[BN * p.ncls + i] = p.head_b[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % nstage;
        const uint32_t round = static_cast<uint32_t>(it / nstage);
        if (!mbar_wait(&empty_bar[s], (round & 1) ^ 1, abort_flag, p.watchdog_ns)) break;
        const int tap = it / chunks;
        const int c0 = (it - tap * chunks) * KC;
        int dy = 0, dx = 0;
        if (p.ntaps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
        uint8_t* a_dst = tiles + static_cast<size_t>(s) * STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_4d(a_dst, &tmA, &full_bar[s], c0, x0 + dx, y0 + dy, n0);
        tma_load_2d(b_dst, &tmB, &full_bar[s], tap * p.Cin + c0, nb0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      bool ok = true;
      for (int it = 0; it < iters; ++it) {
        const int s = it % nstage;
        const uint32_t round = static_cast<uint32_t>(it / nstage);
        if (!mbar_wait(&full_bar[s], round & 1, abort_flag, p.watchdog_ns)) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + static_cast<size_t>(s) * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
          const uint64_t da = umma_smem_desc(a_addr + k * 32, ROW_BYTES);
          const uint64_t db = umma_smem_desc(b_addr + k * 32, ROW_BYTES);
          umma_bf16(tmem_base, da, db, IDESC, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above retire
      }
      if (ok) umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // accumulator row == pixel within the M tile
    const int xx = r % p.TW;
    const int yy = (r / p.TW) % p.TH;
    const int nn = r / (p.TW * p.TH);
    const int x = x0 + xx, y = y0 + yy, n = n0 + nn;
    const bool valid = (x < p.W) && (y < p.H) && (n < p.N);
    const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    const bool acc_ready = mbar_wait(tmem_full_bar, 0, abort_flag, p.watchdog_ns);
    if (__all_sync(0xffffffffu, acc_ready)) {  // warp-uniform: the loop below uses .sync.aligned ops
      tc_fence_after();
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
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = __uint_as_float(raw[j]) + s_bias[c + j];
          if (p.relu) v[j] = fmaxf(v[j], 0.f);
        }
        if constexpr (EPI == EPI_STORE) {
          if (valid) store_bf16x16(p.out + pix * p.out_pitch + p.out_choff + nb0 + c, v);
        } else if constexpr (EPI == EPI_POOL_SKIP) {
          float m[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            m[j] = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, p.TW));
          }
          if (valid && !(xx & 1) && !(yy & 1)) {
            const size_t ppix = (static_cast<size_t>(n) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
            store_bf16x16(p.pool_out + ppix * p.pool_pitch + nb0 + c, m);
          }
          if (p.out != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], s_extra[c + j], s_extra[BN + c + j]), 0.f);
            if (valid) store_bf16x16(p.out + pix * p.out_pitch + p.out_choff + nb0 + c, v);
          }
        } else if constexpr (EPI == EPI_CONVT) {
          const int col = nb0 + c;
          const int g = col / p.Cout;
          const int o = col - g * p.Cout;
          const size_t opix =
              (static_cast<size_t>(n) * (2 * p.H) + (2 * y + (g >> 1))) * (2 * p.W) + (2 * x + (g & 1));
          if (valid) store_bf16x16(p.out + opix * p.out_pitch + p.out_choff + o, v);
        } else {  // EPI_HEAD
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
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, BN);
    if (lane == 0 && *abort_flag) atomicExch(p.err, 1);
  }
}

// Host side -------------------------------------------------------------------
struct ConvLaunch {
  CUtensorMap tmA, tmB;
  ConvParams p;
  int KC, BN, EPI;
  int grid;
  size_t smem;
};

// Returns cudaSuccess or the launch error; cudaErrorInvalidValue for an unsupported (KC,BN,EPI).
cudaError_t conv_launch(const ConvLaunch& L, cudaStream_t stream);
// Sets the max-dynamic-smem attribute of every instantiation (once per device).
cudaError_t conv_init_attributes();

}  // namespace scv
