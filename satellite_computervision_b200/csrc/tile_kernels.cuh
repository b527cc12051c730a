// HBM-bound kernels either side of the network:
//   K1  extract_kernel   -- buffered chip gather + per-band / per-pixel normalise -> bf16 NHWC tiles
//                           (utils/prediction_tools.py:149 slice; utils/processing.py:225-322 normalisers)
//   K4  stitch_kernel    -- sigmoid / softmax head activation + threshold / argmax + crop of the
//                           buffer + scatter of the kept core into the output raster
//                           (utils/prediction_tools.py:154, :267, :346-349, :520; utils/model_tools.py:405-406, :443-445)
//       head_tiles_kernel -- same activation for whole tiles (keras model.predict output)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scv.h"

namespace scv {

struct ExtractParams {
  const uint8_t* src;        // first resident mosaic row
  unsigned long long src_bytes;  // resident bytes (bounds for the 128-bit loads)
  int dtype;                 // SCV_U8..SCV_F64
  int W, C;                  // mosaic width (pixels) and bands
  int src_row0;              // mosaic row index of src
  const int2* origins;       // per tile {.x = x, .y = y} of the chip's upper-left corner INCLUDING the buffer
  int n_tiles;
  int side;                  // chip side (kernel + buff)
  int cpad;                  // stored channels (>= C, zero padded)
  int rows_per_block;
  int norm_mode;
  float sub[SCV_MAX_BANDS];
  float div[SCV_MAX_BANDS];
  float rdiv[SCV_MAX_BANDS];  // filled by launch_extract: RN(1 / div)
  const float* tile_stats;   // SCV_NORM_TILE_*: per tile, per band (sub, div)
  __nv_bfloat16* out;        // n_tiles * side * side * cpad
};

struct TileStatsParams {
  const uint8_t* src;
  int dtype, W, C, src_row0;
  const int2* origins;
  int side;
  int mode;    // SCV_NORM_TILE_ZSCORE or SCV_NORM_TILE_MINMAX
  float eps;
  float* stats;  // n_tiles * C * 2 -> (sub, div)
};

struct StitchParams {
  const float* logits;  // n_tiles * side * side * ncls
  int side, ncls, head;
  float threshold;
  int out_channel;
  int crop;    // buff / 2
  int kernel;  // kept core side
  const int2* dst_origins;  // per tile {.x = x, .y = y} of the core's upper-left in output-raster coordinates
  int force_scalar;         // set when some dst x is not a multiple of 4 (vector path needs aligned stores)
  int dst_row0;             // raster row held at prob[0]
  int out_W;
  float* prob;    // may be null
  uint8_t* mask;  // may be null
};

struct HeadTilesParams {
  const float* logits;
  long long npix;
  int ncls, head;
  float threshold;
  float* probs;      // npix * ncls, may be null
  int32_t* classes;  // npix, may be null
};

size_t extract_smem_bytes(const ExtractParams& p);
cudaError_t launch_extract(const ExtractParams& p, cudaStream_t s);
cudaError_t launch_tile_stats(const TileStatsParams& p, int n_tiles, cudaStream_t s);
cudaError_t launch_stitch(const StitchParams& p, int n_tiles, cudaStream_t s);
cudaError_t launch_head_tiles(const HeadTilesParams& p, cudaStream_t s);
// bf16 -> fp32 widening (debug / tests)
cudaError_t launch_widen(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t s);
cudaError_t launch_narrow(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s);

}  // namespace scv
