#include "tile_kernels.cuh"

#include <cmath>
#include <cstdlib>

namespace scv {

namespace {

__host__ __device__ inline int dtype_size(int dt) {
  switch (dt) {
    case SCV_U8: return 1;
    case SCV_U16:
    case SCV_I16: return 2;
    case SCV_F32: return 4;
    default: return 8;
  }
}

__device__ __forceinline__ float load_elem(const uint8_t* p, int dt) {
  switch (dt) {
    case SCV_U8: return static_cast<float>(*p);
    case SCV_U16: return static_cast<float>(*reinterpret_cast<const uint16_t*>(p));
    case SCV_I16: return static_cast<float>(*reinterpret_cast<const int16_t*>(p));
    case SCV_F32: return *reinterpret_cast<const float*>(p);
    default: return static_cast<float>(*reinterpret_cast<const double*>(p));
  }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ int smem_row_stride(int row_bytes) { return (row_bytes + 16 + 15) & ~15; }

// Element -> fp32 without the (quarter-rate) I2F unit for the integer types: for 0 <= x < 2^23,
// float(x) == as_float(0x4B000000 | x) - 2^23 exactly.
template <int DT>
__device__ __forceinline__ float load_elem_t(const uint8_t* p) {
  if constexpr (DT == SCV_U8) return __uint_as_float(0x4B000000u | *p) - 8388608.0f;
  else if constexpr (DT == SCV_U16)
    return __uint_as_float(0x4B000000u | *reinterpret_cast<const uint16_t*>(p)) - 8388608.0f;
  else if constexpr (DT == SCV_I16)  // bias by 2^15 to make it unsigned, undo after the conversion
    return __uint_as_float(0x4B000000u | (*reinterpret_cast<const uint16_t*>(p) ^ 0x8000u)) - 8421376.0f;
  else if constexpr (DT == SCV_F32) return *reinterpret_cast<const float*>(p);
  else return static_cast<float>(*reinterpret_cast<const double*>(p));
}

// bf16(RN_f32((x - sub) / div)) without paying an IEEE division per element: q' = (x - sub) * RN(1/div) is
// within 2.5 fp32 ulps of the correctly rounded quotient, so both round to the same bf16 unless q' lies within
// a few ulps of a bf16 rounding boundary (low 16 bits near 0x8000) -- only then (about 1 value in 7000) is the
// exact division evaluated.  Results are bit-identical to dividing always.
__device__ __forceinline__ float exact_div_for_bf16(float t, float div, float rdiv) {
  const float q = __fmul_rn(t, rdiv);
  const uint32_t bits = __float_as_uint(q);
  const uint32_t expo = (bits >> 23) & 0xffu;
  if (((bits & 0xffffu) - 0x7ffcu) <= 8u || expo - 2u >= 252u) return __fdiv_rn(t, div);  // near boundary, tiny, inf/nan
  return q;
}

// Which (tile, first row) a block works on.  Legacy mapping: blockIdx.y = tile, blockIdx.x = row block.
// Ordered mapping (regular chip grid, p.order_kernel > 0): blockIdx.y = SOURCE row block m (in units of
// `rpb` mosaic rows from the first chip's top row), blockIdx.x = slot * tiles_per_row + column; slot s looks at
// tile row floor(m / kb) - s (kb = kernel / rpb).  Vertically adjacent chips share side - kernel rows: the two
// blocks that read the same mosaic rows are then only tiles_per_row blocks apart, so the second read hits L2
// instead of DRAM (ncu, round 1: 2.10 GB read for 1.42 GB of unique mosaic bytes).
__device__ __forceinline__ bool extract_block_unit(const ExtractParams& p, int rpb, int& tile, int& r0) {
  if (p.order_kernel <= 0) {
    tile = blockIdx.y;
    r0 = blockIdx.x * rpb;
    return r0 < p.side;
  }
  const int kb = p.order_kernel / rpb, sb = p.side / rpb;
  const int slot = blockIdx.x / p.tiles_per_row, col = blockIdx.x - slot * p.tiles_per_row;
  const int m = blockIdx.y;
  const int tr = m / kb - slot;
  if (tr < 0) return false;
  const int rb = m - tr * kb;
  if (rb >= sb) return false;
  tile = tr * p.tiles_per_row + col - p.order_skip;  // the launch holds chips [order_skip, order_skip + n_tiles) of the grid
  if (tile < 0 || tile >= p.n_tiles) return false;
  r0 = rb * rpb;
  return true;
}

// K1.  One block = `rows_per_block` rows of one chip.  Stage: the rows' bytes are fetched with 128-bit
// loads from the 16-byte-aligned span covering them (any element alignment of the chip origin is handled)
// into shared memory.  Compute: one thread per pixel reads its C bands from shared memory, normalises in
// fp32 exactly as the reference writes it (subtract, then IEEE divide), and writes cpad bf16 channels with
// 128-bit stores (a warp writes 32*cpad*2 contiguous bytes).  CT > 0 fixes the band count at compile time.
template <int DT, int CT>
__global__ void __launch_bounds__(256) extract_kernel(const ExtractParams p) {
  extern __shared__ __align__(16) uint8_t sm[];
  constexpr int esize = DT == SCV_U8 ? 1 : (DT == SCV_U16 || DT == SCV_I16) ? 2 : (DT == SCV_F32 ? 4 : 8);
  constexpr int NV = CT > 0 ? ((CT + 7) / 8) * 8 : SCV_MAX_BANDS;  // values held per pixel
  const int C = CT > 0 ? CT : p.C;
  int tile, r0;
  if (!extract_block_unit(p, p.rows_per_block, tile, r0)) return;
  const int2 org = p.origins[tile];
  const int row_bytes = p.side * C * esize;
  const int sstride = smem_row_stride(row_bytes);
  const int nrows = min(p.rows_per_block, p.side - r0);
  const uint8_t* src_end = p.src + p.src_bytes;
  const long long pitch = static_cast<long long>(p.W) * C * esize;
  const uint8_t* g_first = p.src + static_cast<long long>(org.y + r0 - p.src_row0) * pitch +
                           static_cast<long long>(org.x) * C * esize;

  for (int rr = 0; rr < nrows; ++rr) {
    const uint8_t* g = g_first + rr * pitch;
    const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15);
    const uint8_t* g0 = g - mis;
    const int nvec = (mis + row_bytes + 15) >> 4;
    uint8_t* srow = sm + rr * sstride;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      const uint8_t* gv = g0 + (v << 4);
      uint4 val;
      if (gv >= p.src && gv + 16 <= src_end) {
        val = __ldg(reinterpret_cast<const uint4*>(gv));
      } else {
        uint8_t tmp[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) tmp[b] = (gv + b >= p.src && gv + b < src_end) ? gv[b] : uint8_t(0);
        val = *reinterpret_cast<uint4*>(tmp);
      }
      *reinterpret_cast<uint4*>(srow + (v << 4)) = val;
    }
  }
  __syncthreads();

  const int nchunk = p.cpad >> 3;
  const int mode = p.norm_mode;
  const float* st = p.tile_stats != nullptr ? p.tile_stats + static_cast<size_t>(tile) * C * 2 : nullptr;  // SCV_NORM_TILE_*
  for (int rr = 0; rr < nrows; ++rr) {
    const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(g_first + rr * pitch) & 15);
    const uint8_t* srow = sm + rr * sstride + mis;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(tile) * p.side + (r0 + rr)) * p.side * p.cpad;
    for (int x = threadIdx.x; x < p.side; x += blockDim.x) {
      const uint8_t* s = srow + x * C * esize;
      float v[NV];
#pragma unroll
      for (int c = 0; c < NV; ++c) v[c] = (c < C) ? load_elem_t<DT>(s + c * esize) : 0.f;

      if (mode == SCV_NORM_PER_BAND) {
#pragma unroll
        for (int c = 0; c < NV; ++c)
          if (c < C) v[c] = exact_div_for_bf16(__fsub_rn(v[c], p.sub[c]), p.div[c], p.rdiv[c]);
      } else if (st != nullptr) {
#pragma unroll
        for (int c = 0; c < NV; ++c)
          if (c < C) v[c] = __fdiv_rn(__fsub_rn(v[c], st[2 * c]), st[2 * c + 1]);
      } else if (mode == SCV_NORM_PIXEL_MINMAX || mode == SCV_NORM_PIXEL_ZSCORE || mode == SCV_NORM_PIXEL_ZSCORE_SD) {
        // per pixel, independently per channel group (splits=); channels beyond the last group pass through
        const int ng = p.ngroups > 0 ? p.ngroups : 1;
        for (int g = 0; g < ng; ++g) {
          const int b0 = (p.ngroups > 0 && g > 0) ? p.group_end[g - 1] : 0;
          const int b1 = p.ngroups > 0 ? p.group_end[g] : C;
          if (mode == SCV_NORM_PIXEL_MINMAX) {
            float mn = INFINITY, mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1) {
                mn = fminf(mn, v[c]);
                mx = fmaxf(mx, v[c]);
              }
            const float den = __fadd_rn(__fsub_rn(mx, mn), p.div[0]);
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1) v[c] = __fdiv_rn(__fsub_rn(v[c], mn), den);
          } else if (mode == SCV_NORM_PIXEL_ZSCORE) {
            const float n = static_cast<float>(b1 - b0);
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1) sum = __fadd_rn(sum, v[c]);
            const float mean = __fdiv_rn(sum, n);
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1) {
                const float d = __fsub_rn(v[c], mean);
                ss = __fadd_rn(ss, __fmul_rn(d, d));
              }
            const float den = __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, n), p.div[0]));
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1) v[c] = __fdiv_rn(__fsub_rn(v[c], mean), den);
          } else {
            // normalize_dataArray: nanmean / nanstd (population) across the bands, (x - mean) / (sd + eps)
            float sum = 0.f, n = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1 && v[c] == v[c]) {
                sum = __fadd_rn(sum, v[c]);
                n += 1.f;
              }
            const float mean = __fdiv_rn(sum, n);  // all-NaN pixel: 0/0 = NaN, like np.nanmean
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1 && v[c] == v[c]) {
                const float d = __fsub_rn(v[c], mean);
                ss = __fadd_rn(ss, __fmul_rn(d, d));
              }
            const float den = __fadd_rn(__fsqrt_rn(__fdiv_rn(ss, n)), p.div[0]);
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (c >= b0 && c < b1) v[c] = __fdiv_rn(__fsub_rn(v[c], mean), den);
          }
        }
      }
      if (p.valid_y1 > p.valid_y0) {  // outside the valid window: exact zeros AFTER normalisation
        const int gy = org.y + r0 + rr, gx = org.x + x;
        if (gy < p.valid_y0 || gy >= p.valid_y1 || gx < p.valid_x0 || gx >= p.valid_x1) {
#pragma unroll
          for (int c = 0; c < NV; ++c) v[c] = 0.f;
        }
      }

      uint4* o4 = reinterpret_cast<uint4*>(orow + static_cast<size_t>(x) * p.cpad);
#pragma unroll
      for (int k = 0; k < NV / 8; ++k) {
        if (k < nchunk) {
          uint4 w;
          w.x = pack2(v[8 * k + 0], v[8 * k + 1]);
          w.y = pack2(v[8 * k + 2], v[8 * k + 3]);
          w.z = pack2(v[8 * k + 4], v[8 * k + 5]);
          w.w = pack2(v[8 * k + 6], v[8 * k + 7]);
          o4[k] = w;
        }
      }
      for (int k = NV / 8; k < nchunk; ++k) o4[k] = make_uint4(0, 0, 0, 0);
    }
  }
}

// K1, fast path for the BASELINE input: uint16 digital numbers, 6 bands, per-band constants, 8 stored channels.
// One warp per chip row, one pixel (12 B in, 16 B out) per lane and step: three 32-bit loads (12*x is always
// 4-byte aligned), u16 -> fp32 with one PRMT + one FADD each (0x4B00xxxx - 2^23), the reference's fp32
// subtract, and the division as multiply-by-reciprocal with the bf16 rounding-boundary test of
// exact_div_for_bf16 (integers cannot produce the tiny / non-finite cases, so that part of the test is gone).
// No shared memory, no block barrier: ~45 instructions per pixel instead of ~110, and every byte of a 128-byte
// line is consumed by the same warp within three consecutive instructions (L1 hits).
template <bool SUB>
__global__ void __launch_bounds__(256, 8) extract_u16x6_kernel(const ExtractParams p) {
  int tile, r0;
  if (!extract_block_unit(p, 8, tile, r0)) return;
  const int row = r0 + (threadIdx.x >> 5);
  if (row >= p.side) return;
  const int lane = threadIdx.x & 31;
  const int2 org = p.origins[tile];
  const uint32_t* src = reinterpret_cast<const uint32_t*>(
      p.src + (static_cast<long long>(org.y + row - p.src_row0) * p.W + org.x) * 12);
  uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(tile) * p.side + row) * p.side * 8);
  float sub[6], div[6], rdiv[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    sub[c] = p.sub[c];
    div[c] = p.div[c];
    rdiv[c] = p.rdiv[c];
  }
#pragma unroll 4
  for (int x = lane; x < p.side; x += 32) {
    uint32_t w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = __ldg(src + 3 * x + k);
    float q[6];
    bool slow = false;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      // bytes {lo, hi, 0x00, 0x4B}: as_float == 2^23 + value exactly
      const uint32_t bits = __byte_perm(w[c >> 1], 0x4B000000u, (c & 1) ? 0x7632 : 0x7610);
      float t = __uint_as_float(bits) - 8388608.0f;
      if (SUB) t = __fsub_rn(t, sub[c]);
      q[c] = __fmul_rn(t, rdiv[c]);
      slow |= ((__float_as_uint(q[c]) & 0xffffu) - 0x7ffcu) <= 8u;
    }
    if (slow) {  // some value sits within a few ulps of a bf16 rounding boundary: take the IEEE quotient
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (((__float_as_uint(q[c]) & 0xffffu) - 0x7ffcu) <= 8u) {
          const uint32_t bits = __byte_perm(w[c >> 1], 0x4B000000u, (c & 1) ? 0x7632 : 0x7610);
          float t = __uint_as_float(bits) - 8388608.0f;
          if (SUB) t = __fsub_rn(t, sub[c]);
          q[c] = __fdiv_rn(t, div[c]);
        }
      }
    }
    uint4 o;
    o.x = pack2(q[0], q[1]);
    o.y = pack2(q[2], q[3]);
    o.z = pack2(q[4], q[5]);
    o.w = 0u;
    __stcs(dst + x, o);  // streaming: the whole scene's tiles are written before the first conv reads them
  }
}

// Per-tile statistics for SCV_NORM_TILE_* (one block per tile): per band (axes=[0,1]) or, for the GLOBAL
// modes (axes=[0,1,2]), per channel group.  Two passes like tf.nn.moments: mean, then mean of squared
// differences.  Output: per band (sub, div) so the extract kernel applies (x - sub) / div either way.
__global__ void __launch_bounds__(256) tile_stats_kernel(const TileStatsParams p) {
  __shared__ float red[2][SCV_MAX_BANDS][8];
  __shared__ float band[2][SCV_MAX_BANDS];
  __shared__ float s_mean[SCV_MAX_BANDS];
  const int tile = blockIdx.x;
  const int2 org = p.origins[tile];
  const int esize = dtype_size(p.dtype);
  const int C = p.C;
  const int npix = p.side * p.side;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool zs = p.mode == SCV_NORM_TILE_ZSCORE || p.mode == SCV_NORM_TILE_GLOBAL_ZSCORE;
  const bool global = p.mode == SCV_NORM_TILE_GLOBAL_ZSCORE || p.mode == SCV_NORM_TILE_GLOBAL_MINMAX;

  float a[SCV_MAX_BANDS], b[SCV_MAX_BANDS];
  for (int pass = 0; pass < (zs ? 2 : 1); ++pass) {
#pragma unroll
    for (int c = 0; c < SCV_MAX_BANDS; ++c) {
      a[c] = zs ? 0.f : INFINITY;
      b[c] = zs ? 0.f : -INFINITY;
    }
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
      const int y = i / p.side, x = i - y * p.side;
      const uint8_t* s =
          p.src + ((static_cast<long long>(org.y + y - p.src_row0) * p.W + org.x + x) * C) * esize;
#pragma unroll
      for (int c = 0; c < SCV_MAX_BANDS; ++c)
        if (c < C) {
          const float v = load_elem(s + c * esize, p.dtype);
          if (zs) {
            if (pass == 0) a[c] += v;
            else {
              const float d = v - s_mean[c];
              a[c] += d * d;
            }
          } else {
            a[c] = fminf(a[c], v);
            b[c] = fmaxf(b[c], v);
          }
        }
    }
#pragma unroll
    for (int c = 0; c < SCV_MAX_BANDS; ++c) {
      if (c < C) {
        for (int o = 16; o > 0; o >>= 1) {
          const float oa = __shfl_xor_sync(0xffffffffu, a[c], o);
          const float ob = __shfl_xor_sync(0xffffffffu, b[c], o);
          a[c] = zs ? a[c] + oa : fminf(a[c], oa);
          b[c] = zs ? b[c] + ob : fmaxf(b[c], ob);
        }
        if (lane == 0) {
          red[0][c][warp] = a[c];
          red[1][c][warp] = b[c];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < C) {
      const int c = threadIdx.x;
      float ra = red[0][c][0], rb = red[1][c][0];
      for (int w = 1; w < 8; ++w) {
        ra = zs ? ra + red[0][c][w] : fminf(ra, red[0][c][w]);
        rb = zs ? rb + red[1][c][w] : fmaxf(rb, red[1][c][w]);
      }
      band[0][c] = ra;
      band[1][c] = rb;
    }
    __syncthreads();
    if (threadIdx.x < C) {
      const int c = threadIdx.x;
      float ra = band[0][c], rb = band[1][c];
      float count = static_cast<float>(npix);
      bool pass_through = false;
      if (global) {  // combine the bands of this band's group
        const int ng = p.ngroups > 0 ? p.ngroups : 1;
        int b0 = 0, b1 = p.ngroups > 0 ? 0 : C;
        pass_through = p.ngroups > 0;
        for (int g = 0; g < ng && p.ngroups > 0; ++g) {
          const int lo = g > 0 ? p.group_end[g - 1] : 0, hi = p.group_end[g];
          if (c >= lo && c < hi) {
            b0 = lo, b1 = hi;
            pass_through = false;
          }
        }
        if (!pass_through) {
          ra = band[0][b0];
          rb = band[1][b0];
          for (int k = b0 + 1; k < b1; ++k) {
            ra = zs ? ra + band[0][k] : fminf(ra, band[0][k]);
            rb = zs ? rb + band[1][k] : fmaxf(rb, band[1][k]);
          }
          count *= static_cast<float>(b1 - b0);
        }
      } else if (p.ngroups > 0 && c >= p.group_end[p.ngroups - 1]) {
        pass_through = true;
      }
      float* st = p.stats + (static_cast<size_t>(tile) * C + c) * 2;
      if (pass_through) {
        s_mean[c] = 0.f;
        st[0] = 0.f;
        st[1] = 1.f;
      } else if (zs) {
        if (pass == 0) s_mean[c] = ra / count;
        else {
          st[0] = s_mean[c];
          st[1] = sqrtf(ra / count + p.eps);
        }
      } else {
        st[0] = ra;
        st[1] = (rb - ra) + p.eps;
      }
    }
    __syncthreads();
  }
}

// sigmoid with the hardware exp2 / reciprocal approximations (relative error ~1e-6, far inside the 1e-2 budget):
// the full-precision expf + IEEE division made K4 ALU-bound under the power-capped clocks of a long step (ncu:
// sm__throughput 73 % next to 60 % DRAM).  Every head path uses this one function, so stitched rasters and
// whole-tile predictions stay bit-identical.
__device__ __forceinline__ float sigmoid_fast(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }

__device__ __forceinline__ void head_eval(const float* z, int ncls, int head, float thr, int out_channel,
                                          float& prob, int& cls) {
  if (head == SCV_HEAD_SIGMOID) {
    const float pr = sigmoid_fast(z[0]);
    prob = pr;
    cls = pr > thr ? 1 : 0;
  } else {
    float m = z[0];
    for (int k = 1; k < ncls; ++k) m = fmaxf(m, z[k]);
    float e[SCV_MAX_CLASSES];
    float sum = 0.f;
    for (int k = 0; k < ncls; ++k) {
      e[k] = expf(z[k] - m);
      sum += e[k];
    }
    float best = -1.f;
    int bi = 0;
    for (int k = 0; k < ncls; ++k) {
      const float pk = e[k] / sum;
      if (pk > best) {
        best = pk;
        bi = k;
      }
      if (k == out_channel) prob = pk;
    }
    cls = bi;
  }
}

template <typename OUT, bool ACC>
__device__ __forceinline__ void put_prob(void* base, size_t o, float pr) {
  OUT* q = reinterpret_cast<OUT*>(base) + o;
  *q = ACC ? static_cast<OUT>(*q + static_cast<OUT>(pr)) : static_cast<OUT>(pr);
}

// K4, vector path: sigmoid head (ncls == 1), PX (4 or 8) consecutive core pixels per thread: float4 logit
// loads, float4 / double2 probability stores, one 32- / 64-bit mask store; streaming (evict-first) accesses:
// nothing here is read again on the device.  OUT = float | double raster, ACC: += instead of = (:154).
template <int PX, typename OUT, bool ACC>
__global__ void __launch_bounds__(256) stitch_kernel_vec(const StitchParams p) {
  const int tile = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = p.kernel_w / PX;
  if (q >= per_row * p.kernel_h) return;
  const int row = q / per_row;
  const int col = (q - row * per_row) * PX;
  const int2 d = p.dst_origins[tile];
  const float* zsrc = p.logits + (static_cast<size_t>(tile) * p.side + (p.crop_y + row)) * p.side + p.crop_x + col;
  float z[PX], pr[PX];
#pragma unroll
  for (int k = 0; k < PX / 4; ++k) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(zsrc) + k);
    z[4 * k] = t.x, z[4 * k + 1] = t.y, z[4 * k + 2] = t.z, z[4 * k + 3] = t.w;
  }
#pragma unroll
  for (int k = 0; k < PX; ++k) pr[k] = sigmoid_fast(z[k]);
  const size_t o = static_cast<size_t>(d.y + row - p.dst_row0) * p.out_W + d.x + col;
  if (p.prob) {
    if constexpr (sizeof(OUT) == 4) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.prob) + o);
#pragma unroll
      for (int k = 0; k < PX / 4; ++k) {
        float4 v = make_float4(pr[4 * k], pr[4 * k + 1], pr[4 * k + 2], pr[4 * k + 3]);
        if (ACC) {
          const float4 t = dst[k];
          v.x += t.x, v.y += t.y, v.z += t.z, v.w += t.w;
        }
        __stcs(dst + k, v);
      }
    } else {
      double2* dst = reinterpret_cast<double2*>(reinterpret_cast<double*>(p.prob) + o);
#pragma unroll
      for (int k = 0; k < PX / 2; ++k) {
        double2 v = make_double2(static_cast<double>(pr[2 * k]), static_cast<double>(pr[2 * k + 1]));
        if (ACC) {
          const double2 t = dst[k];
          v.x += t.x, v.y += t.y;
        }
        __stcs(dst + k, v);
      }
    }
  }
  if (p.mask) {
    uint32_t m[PX / 4];
#pragma unroll
    for (int k = 0; k < PX / 4; ++k)
      m[k] = (pr[4 * k] > p.threshold ? 1u : 0u) | (pr[4 * k + 1] > p.threshold ? 1u << 8 : 0u) |
             (pr[4 * k + 2] > p.threshold ? 1u << 16 : 0u) | (pr[4 * k + 3] > p.threshold ? 1u << 24 : 0u);
    if constexpr (PX == 8) __stcs(reinterpret_cast<uint2*>(p.mask + o), make_uint2(m[0], m[1]));
    else __stcs(reinterpret_cast<uint32_t*>(p.mask + o), m[0]);
  }
}

// K4, general path: any head / class count / alignment / crop window, one pixel per thread.
template <typename OUT, bool ACC>
__global__ void __launch_bounds__(256) stitch_kernel_scalar(const StitchParams p) {
  const int tile = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.kernel_h * p.kernel_w) return;
  const int row = q / p.kernel_w;
  const int col = q - row * p.kernel_w;
  const int2 d = p.dst_origins[tile];
  const float* z =
      p.logits + ((static_cast<size_t>(tile) * p.side + (p.crop_y + row)) * p.side + p.crop_x + col) * p.ncls;
  float zz[SCV_MAX_CLASSES];
  for (int k = 0; k < p.ncls; ++k) zz[k] = __ldg(z + k);
  float prob = 0.f;
  int cls = 0;
  head_eval(zz, p.ncls, p.head, p.threshold, p.out_channel, prob, cls);
  const size_t o = static_cast<size_t>(d.y + row - p.dst_row0) * p.out_W + d.x + col;
  if (p.prob) put_prob<OUT, ACC>(p.prob, o, prob);
  if (p.mask) p.mask[o] = static_cast<uint8_t>(cls);
}

__global__ void __launch_bounds__(256) head_tiles_kernel(const HeadTilesParams p) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.npix) return;
  float zz[SCV_MAX_CLASSES];
  for (int k = 0; k < p.ncls; ++k) zz[k] = __ldg(p.logits + i * p.ncls + k);
  if (p.head == SCV_HEAD_SIGMOID) {
    const float pr = sigmoid_fast(zz[0]);
    if (p.probs) p.probs[i] = pr;
    if (p.classes) p.classes[i] = pr > p.threshold ? 1 : 0;
  } else {
    float m = zz[0];
    for (int k = 1; k < p.ncls; ++k) m = fmaxf(m, zz[k]);
    float sum = 0.f;
    for (int k = 0; k < p.ncls; ++k) {
      zz[k] = expf(zz[k] - m);
      sum += zz[k];
    }
    float best = -1.f;
    int bi = 0;
    for (int k = 0; k < p.ncls; ++k) {
      const float pk = zz[k] / sum;
      if (p.probs) p.probs[i * p.ncls + k] = pk;
      if (pk > best) {
        best = pk;
        bi = k;
      }
    }
    if (p.classes) p.classes[i] = bi;
  }
}

__global__ void widen_kernel(const __nv_bfloat16* src, float* dst, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}
__global__ void narrow_kernel(const float* src, __nv_bfloat16* dst, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

}  // namespace

size_t extract_smem_bytes(const ExtractParams& p) {
  const int row_bytes = p.side * p.C * dtype_size(p.dtype);
  return static_cast<size_t>(p.rows_per_block) * ((row_bytes + 16 + 15) & ~15);
}

// grid of an extract launch (see extract_block_unit)
static dim3 extract_grid(const ExtractParams& p, int rpb) {
  if (p.order_kernel <= 0) return dim3((p.side + rpb - 1) / rpb, p.n_tiles);
  const int kb = p.order_kernel / rpb, sb = p.side / rpb;
  const int n_tile_rows = (p.order_skip + p.n_tiles + p.tiles_per_row - 1) / p.tiles_per_row;
  return dim3(((sb + kb - 1) / kb) * p.tiles_per_row, (n_tile_rows - 1) * kb + sb);
}

template <int DT, int CT>
static cudaError_t launch_extract_tc(const ExtractParams& p, cudaStream_t s) {
  const size_t smem = extract_smem_bytes(p);
  static bool attr_done[64] = {};  // the attribute is per device (one engine per GPU, several per process)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e =
        cudaFuncSetAttribute(extract_kernel<DT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  extract_kernel<DT, CT><<<extract_grid(p, p.rows_per_block), 256, smem, s>>>(p);
  return cudaGetLastError();
}
template <int DT>
static cudaError_t launch_extract_t(const ExtractParams& p, cudaStream_t s) {
  switch (p.C) {  // the band counts of the reference's models: Sentinel-2 (6), NAIP (3 / 4)
    case 6: return launch_extract_tc<DT, 6>(p, s);
    case 4: return launch_extract_tc<DT, 4>(p, s);
    case 3: return launch_extract_tc<DT, 3>(p, s);
    default: return launch_extract_tc<DT, 0>(p, s);
  }
}

cudaError_t launch_extract(const ExtractParams& pin, cudaStream_t s) {
  if (pin.n_tiles <= 0) return cudaSuccess;
  ExtractParams p = pin;
  for (int c = 0; c < SCV_MAX_BANDS; ++c) p.rdiv[c] = 1.0f / p.div[c];  // fp32 RN reciprocal (host IEEE division)
  // source-row-ordered block mapping needs whole row blocks on both the chip side and the chip pitch
  const bool fast = p.dtype == SCV_U16 && p.C == 6 && p.cpad == 8 && p.norm_mode == SCV_NORM_PER_BAND &&
                    p.valid_y1 <= p.valid_y0 && (reinterpret_cast<uintptr_t>(p.src) & 3) == 0 && !getenv("SCV_K1_GENERIC");
  const int rpb = fast ? 8 : p.rows_per_block;
  if (p.order_kernel > 0 && (p.tiles_per_row <= 0 || p.order_skip < 0 || p.order_skip >= p.tiles_per_row || p.order_kernel % rpb || p.side % rpb ||
                             p.order_kernel > p.side || getenv("SCV_K1_UNORDERED")))
    p.order_kernel = 0;
  if (fast) {
    bool sane = true, sub = false;
    for (int c = 0; c < 6; ++c) {
      sane = sane && p.div[c] >= 1e-3f && p.div[c] <= 1e9f && fabsf(p.sub[c]) <= 1e9f;  // quotients stay normal and finite
      sub = sub || p.sub[c] != 0.f;
    }
    if (sane) {
      const dim3 grid = extract_grid(p, 8);
      if (sub) extract_u16x6_kernel<true><<<grid, 256, 0, s>>>(p);
      else extract_u16x6_kernel<false><<<grid, 256, 0, s>>>(p);
      return cudaGetLastError();
    }
    if (p.order_kernel > 0 && (p.order_kernel % p.rows_per_block || p.side % p.rows_per_block)) p.order_kernel = 0;
  }
  switch (p.dtype) {
    case SCV_U8: return launch_extract_t<SCV_U8>(p, s);
    case SCV_U16: return launch_extract_t<SCV_U16>(p, s);
    case SCV_I16: return launch_extract_t<SCV_I16>(p, s);
    case SCV_F32: return launch_extract_t<SCV_F32>(p, s);
    case SCV_F64: return launch_extract_t<SCV_F64>(p, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_tile_stats(const TileStatsParams& p, int n_tiles, cudaStream_t s) {
  if (n_tiles <= 0) return cudaSuccess;
  tile_stats_kernel<<<n_tiles, 256, 0, s>>>(p);
  return cudaGetLastError();
}

template <typename OUT, bool ACC>
static cudaError_t launch_stitch_t(const StitchParams& p, int n_tiles, int px, cudaStream_t s) {
  if (px == 8) {
    dim3 grid(((p.kernel_w >> 3) * p.kernel_h + 255) / 256, n_tiles);
    stitch_kernel_vec<8, OUT, ACC><<<grid, 256, 0, s>>>(p);
  } else if (px == 4) {
    dim3 grid(((p.kernel_w >> 2) * p.kernel_h + 255) / 256, n_tiles);
    stitch_kernel_vec<4, OUT, ACC><<<grid, 256, 0, s>>>(p);
  } else {
    dim3 grid((p.kernel_h * p.kernel_w + 255) / 256, n_tiles);
    stitch_kernel_scalar<OUT, ACC><<<grid, 256, 0, s>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_stitch(const StitchParams& p, int n_tiles, cudaStream_t s) {
  if (n_tiles <= 0) return cudaSuccess;
  // vector path: every address a thread touches must be aligned for its PX pixels.  The caller reports the
  // destination columns through force_scalar: bit 0 = some x origin is not a multiple of 4, bit 1 = not of 8.
  const int esz = p.prob_f64 ? 8 : 4;
  auto aligned = [&](int px) {
    return p.head == SCV_HEAD_SIGMOID && p.ncls == 1 && p.kernel_w % px == 0 && p.crop_x % 4 == 0 && p.side % 4 == 0 &&
           p.out_W % px == 0 && (reinterpret_cast<uintptr_t>(p.prob) % (px * esz > 16 ? 16 : px * esz)) == 0 &&
           (reinterpret_cast<uintptr_t>(p.mask) % px) == 0 && (reinterpret_cast<uintptr_t>(p.logits) & 15) == 0;
  };
  int px = 1;
  if (!(p.force_scalar & 1) && aligned(4)) px = 4;
  if (px == 4 && !(p.force_scalar & 2) && aligned(8)) px = 8;
  if (const char* o = getenv("SCV_K4_PX"))  // experiments: cap the pixels per thread
    if (atoi(o) > 0 && atoi(o) < px) px = atoi(o) >= 4 ? 4 : 1;
  if (p.prob_f64) return p.accumulate ? launch_stitch_t<double, true>(p, n_tiles, px, s) : launch_stitch_t<double, false>(p, n_tiles, px, s);
  return p.accumulate ? launch_stitch_t<float, true>(p, n_tiles, px, s) : launch_stitch_t<float, false>(p, n_tiles, px, s);
}

cudaError_t launch_head_tiles(const HeadTilesParams& p, cudaStream_t s) {
  if (p.npix <= 0) return cudaSuccess;
  head_tiles_kernel<<<static_cast<unsigned>((p.npix + 255) / 256), 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_widen(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  widen_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t launch_narrow(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  narrow_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, dst, n);
  return cudaGetLastError();
}

}  // namespace scv
