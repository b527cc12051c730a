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
  // splits= of the data-derived per-pixel modes: channel groups [group_end[g-1], group_end[g]); channels
  // >= group_end[ngroups-1] pass through
  int ngroups;
  int group_end[SCV_MAX_BANDS];
  // pixels outside [valid_y0, valid_y1) x [valid_x0, valid_x1) (mosaic coordinates) become exact zeros after
  // normalisation; valid_y1 <= valid_y0: everything is valid
  int valid_y0, valid_y1, valid_x0, valid_x1;
  // > 0: the tiles are a regular row-major grid (tiles_per_row per tile row, origins `order_kernel` rows apart):
  // blocks are then ordered by SOURCE mosaic row, so the rows two vertically adjacent chips share are read
  // back to back and the second read hits L2
  // order_skip: chips missing at the start of the first tile row (a chip RANGE that starts mid-row)
  int order_kernel, tiles_per_row, order_skip;
};

struct TileStatsParams {
  const uint8_t* src;
  int dtype, W, C, src_row0;
  const int2* origins;
  int side;
  int mode;    // SCV_NORM_TILE_ZSCORE / _MINMAX (per band) or SCV_NORM_TILE_GLOBAL_ZSCORE / _MINMAX (per group)
  float eps;
  float* stats;  // n_tiles * C * 2 -> (sub, div)
  int ngroups;   // global modes: channel groups; bands beyond the last group get (0, 1)
  int group_end[SCV_MAX_BANDS];
};

struct StitchParams {
  const float* logits;  // n_tiles * side * side * ncls
  int side, ncls, head;
  float threshold;
  int out_channel;
  int crop_y, crop_x;        // first kept row / column of a tile (buff / 2)
  int kernel_h, kernel_w;    // kept core size
  const int2* dst_origins;  // per tile {.x = x, .y = y} of the core's upper-left in output-raster coordinates
  int force_scalar;         // bit 0: some dst x is not a multiple of 4, bit 1: not of 8 (vector paths need aligned stores)
  int dst_row0;             // raster row held at prob[0]
  int out_W;
  void* prob;     // float or double raster, may be null
  int prob_f64;   // prob is double (the reference's float64 template, utils/prediction_tools.py:769)
  int accumulate; // prob[o] += p (utils/prediction_tools.py:154) instead of prob[o] = p
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
