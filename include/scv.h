/* scv.h -- C-ABI of the B200-native tiled U-Net predict path (libscv.so).
 *
 * The reference (mjevans26/Satellite_ComputerVision) is pure Python and has no
 * FFI/plugin API of its own: its boundary is the Keras duck-type
 * `model.predict(...)` called from `utils/prediction_tools.py` (:152, :251,
 * :333, :515) and `utils/model_tools.py:1299`.  This header is the C-ABI a
 * drop-in for that path binds (ctypes stub: satellite_computervision_b200/_lib.py,
 * reference-side binding: INTEGRATION.md).  Each entry point cites the reference
 * interface it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions: every function returning int returns SCV_OK (0) or a negative
 * status; scv_last_error() gives the thread-local message.  The caller owns
 * every host/device buffer passed in for the duration of the call; the engine
 * owns its device memory, streams and tensor maps.  A handle is not
 * thread-safe; different handles are independent (one engine per GPU).
 * There is no CPU fallback: without a CUDA device every compute entry fails
 * with SCV_ERR_CUDA.
 */
#ifndef SCV_H_
#define SCV_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SCV_API __attribute__((visibility("default")))
#else
#define SCV_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SCV_OK 0
#define SCV_ERR_INVALID (-1) /* bad argument / shape precondition            */
#define SCV_ERR_CUDA (-2)    /* CUDA runtime/driver error, or no device       */
#define SCV_ERR_STATE (-3)   /* e.g. predict before set_weights               */
#define SCV_ERR_KERNEL (-4)  /* device-side watchdog tripped (pipeline stall) */

#define SCV_MAX_LEVELS 8
#define SCV_MAX_BANDS 16
#define SCV_MAX_CLASSES 16
#define SCV_MAX_LAYERS 64

typedef struct scv_engine scv_engine;

/* element type of mosaic / patch inputs */
enum { SCV_U8 = 0, SCV_U16 = 1, SCV_I16 = 2, SCV_F32 = 3, SCV_F64 = 4 };
/* heads: utils/model_tools.py:443-445 (sigmoid, strict >thr) and :405-406 (softmax, argmax) */
enum { SCV_HEAD_SIGMOID = 0, SCV_HEAD_SOFTMAX = 1 };
/* network families: the U-Net of get_unet_model / binary_unet (utils/model_tools.py:394-454) and the siamese U-Net with an
 * atrous spatial pyramid of make_siamese_unet (utils/model_tools.py:533-663) */
enum { SCV_ARCH_UNET = 0, SCV_ARCH_SIAMESE = 1 };
/* normaliser fused into the extract kernel */
enum {
  SCV_NORM_NONE = 0,
  /* y = (x - sub[c]) / div[c] in fp32.  Covers rescale_tensor(moments=)
   * (utils/processing.py:302-311: sub=min, div=(max-min)+eps), normalize_tensor(
   * moments=) (:252-262: sub=mean, div=sqrt(var+eps)) and the scalar rescale of
   * UNETDataGenerator (:551-552, :601, :613: sub=0, div=10000 | 255). */
  SCV_NORM_PER_BAND = 1,
  /* rescale_tensor default axes=[2] (utils/processing.py:281, :307-308): per pixel
   * (x - min_c x) / ((max_c x - min_c x) + eps); eps = div[0]. */
  SCV_NORM_PIXEL_MINMAX = 2,
  /* normalize_tensor axes=[2] (:225, :257-262): per pixel z-score across bands,
   * population variance, (x-mean)/sqrt(var+eps); eps = div[0]. */
  SCV_NORM_PIXEL_ZSCORE = 3,
  /* per-tile per-band z-score, solar notebook normalize(stacked,[0,1])
   * (notebooks/UNET_G4G_2019_solar.ipynb:808-820, :1541); eps = div[0]. */
  SCV_NORM_TILE_ZSCORE = 4,
  /* rescale_tensor axes=[0,1]: per-tile per-band min/max; eps = div[0]. */
  SCV_NORM_TILE_MINMAX = 5,
  /* pc_tools.normalize_dataArray(da, 'band') (utils/pc_tools.py:90-107, applied to the mosaic at
   * utils/prediction_tools.py:757-758): per pixel across bands, NaN-skipping mean and population
   * standard deviation, (x - mean) / (sd + eps); eps = div[0] (1e-6 in the reference).  A NaN band stays NaN. */
  SCV_NORM_PIXEL_ZSCORE_SD = 6,
  /* rescale_tensor axes=[0,1,2] (utils/processing.py:307-308): ONE min/max per tile (per channel group). */
  SCV_NORM_TILE_GLOBAL_MINMAX = 7,
  /* normalize_tensor axes=[0,1,2] (:257): ONE mean/variance per tile (per channel group). */
  SCV_NORM_TILE_GLOBAL_ZSCORE = 8
};

/* Architecture = arguments of model_tools.get_unet_model (utils/model_tools.py:394)
 * plus the variant switch of SURVEY Appendix A1. */
typedef struct {
  int device;                  /* CUDA ordinal                                             */
  int double_conv;             /* 1: two conv-BN-ReLU per encoder/centre block (notebook
                                  get_model, solar.ipynb:1162-1170); 0: one, as
                                  conv_block.call is written (model_tools.py:238-239)      */
  int nchannels;               /* input bands (<= SCV_MAX_BANDS)                            */
  int nclasses;                /* head width (1 for the sigmoid head)                      */
  int nlevels;                 /* len(filters)                                             */
  int filters[SCV_MAX_LEVELS]; /* multiples of 32; pooling factors are all 2               */
  int head;                    /* SCV_HEAD_*                                               */
  float threshold;             /* sigmoid class threshold (0.5; 0.9 for solar, :444)       */
  int max_batch;               /* tiles per device batch (0 = default 64)                  */
  int arch;                    /* SCV_ARCH_*.  SCV_ARCH_SIAMESE: make_siamese_unet(n_channels, filters, factors)
                                  (model_tools.py:638-663): `nchannels` is the band count of ONE image; every input
                                  (tiles, patches, mosaics) carries the two images stacked along the channel axis,
                                  [input_a | input_b] = 2 * nchannels bands; shared single-conv encoder blocks,
                                  ASPP (1x1 + 3x3 at dilation 3 / 6 / 12 -> 1x1) on both pooled images, decoder over
                                  concat([encoded_b, encoded_a, up]); sigmoid head, strict > threshold           */
} scv_config;

/* One Keras weight array, fp32, C-contiguous, Keras layout (Conv2D HWIO,
 * Conv2DTranspose (kh,kw,out,in), BN [gamma,beta,moving_mean,moving_variance]). */
typedef struct {
  const float* data;
  int ndim;
  int64_t shape[4];
} scv_tensor;

typedef struct {
  int mode; /* SCV_NORM_* */
  int nbands;
  float sub[SCV_MAX_BANDS];
  float div[SCV_MAX_BANDS];
  /* data-derived modes (every mode except NONE / PER_BAND): `splits=` of rescale_tensor / normalize_tensor
   * (utils/processing.py:314-318, :267-275) -- contiguous channel groups normalised independently.
   * ngroups == 0: one group of all nbands.  Channels beyond sum(group_size) pass through unchanged
   * (normalize_tensor :269-274; derived / one-hot planes appended by make_pred_dataset,
   * utils/prediction_tools.py:198-215). */
  int ngroups;
  int group_size[SCV_MAX_BANDS];
} scv_norm;

/* generate_chip_indices(arr, buff, kernel), utils/prediction_tools.py:87-109 */
typedef struct {
  int kernel; /* kept core side (256)                       */
  int buff;   /* TOTAL buffer (128): buff/2 trimmed per side */
} scv_tiling;

/* Options of the extended mosaic calls.  Zero-initialised == the plain calls on the whole chip list. */
typedef struct {
  int tile_begin, tile_end; /* chips [begin, end) of the row-major chip list of generate_chip_indices
                               (tile-balanced multi-GPU sharding); end <= 0: to the last chip */
  int out_channel;          /* probability channel stitched (0 as at prediction_tools.py:154; 1 as at :267) */
  int out_dtype;            /* element type of out_prob: 0 / SCV_F32 or SCV_F64 (the reference's template is
                               float64 zeros, utils/prediction_tools.py:769)                                 */
  int accumulate;           /* 1: out_prob[core] += p like predict_chips (:154); 0: assign                  */
  int valid[4];             /* y0, y1, x0, x1 (exclusive ends), mosaic coordinates: pixels outside this window
                               are fed to the network as exact zeros AFTER normalisation (dask map_overlap
                               boundary=0 pads the already normalised raster, utils/prediction_tools.py:823-829);
                               all zero: the whole mosaic is valid                                          */
} scv_mosaic_opts;

/* crop window of the patch-list stitchers (utils/prediction_tools.py:258-261, :340-343, :496-520): patch i keeps
 * rows [y0, y0+h) and columns [x0, x0+w) and lands at row (i / cols) * h, column (i % cols) * w. */
typedef struct {
  int y0, x0, h, w;
} scv_crop;

/* per-call device timings of the last predict call, milliseconds (CUDA events) */
typedef struct {
  float total_ms;
  float extract_ms; /* K1 gather+normalise, summed over batches */
  float network_ms; /* all conv / convT launches                */
  float stitch_ms;  /* K4 head activation + crop + stitch       */
  int n_batches;
  int n_tiles;
  int n_launches;                   /* kernels launched by the call              */
  int n_layers;                     /* conv/convT launches per batch             */
  float layer_ms[SCV_MAX_LAYERS];   /* per layer, summed over batches (only when
                                       scv_set_option("profile_layers",1))      */
  double layer_flops[SCV_MAX_LAYERS]; /* algorithmic FLOPs per tile of that layer */
  /* host-buffer mosaic calls only (0 otherwise): the exposed ends of the H2D / compute / D2H pipeline */
  float h2d_lead_ms;                /* first H2D copy enqueued -> first kernel starts */
  float d2h_tail_ms;                /* last kernel ends -> last D2H copy lands        */
} scv_times;

/* ---- lifecycle ------------------------------------------------------------ */
SCV_API const char* scv_version(void);
SCV_API const char* scv_last_error(void);
SCV_API int scv_device_count(void);

/* model_tools.get_unet_model(nclasses, nchannels, filters, ...) :394-415 /
 * binary_unet :417-454 -- builds the layer graph and allocates nothing big yet. */
SCV_API int scv_engine_create(const scv_config* cfg, scv_engine** out);
SCV_API void scv_engine_destroy(scv_engine* e);

/* Number and shapes of the arrays keras `model.get_weights()` returns for this
 * architecture, in order (so a host binding can validate before upload). */
SCV_API int scv_num_weights(const scv_config* cfg);
SCV_API int scv_weight_shape(const scv_config* cfg, int index, int* ndim, int64_t shape[4], char* name, int name_len);

/* keras Model.set_weights / load_weights (utils/model_tools.py:1162, :1200):
 * `tensors` is the get_weights()-ordered list.  The engine folds every
 * BatchNormalization (eps 1e-3, moving statistics) into the adjacent conv in
 * fp32, rounds once to bf16, re-tiles to the UMMA K-major layout and uploads. */
SCV_API int scv_engine_set_weights(scv_engine* e, const scv_tensor* tensors, int n);

/* Tuning / debug knobs: "profile_layers", "stages", "watchdog_ms". */
SCV_API int scv_set_option(scv_engine* e, const char* key, int value);

/* ---- predict: host buffers (the drop-in calls) ------------------------------ */

/* keras `model.predict(x)` on a batch of equally sized patches
 * (utils/prediction_tools.py:152, :251, :333, :515; utils/model_tools.py:1299).
 * nhwc: N*H*W*C elements of `dtype`; H, W multiples of 2^nlevels.
 * probs: N*H*W*nclasses fp32 (may be NULL); classes: N*H*W int32 (may be NULL). */
SCV_API int scv_predict_tiles(scv_engine* e, const void* nhwc, int dtype, int N, int H, int W, int C,
                      const scv_norm* norm, float* probs, int32_t* classes);

/* generate_chip_indices + predict_chips on a raster mosaic
 * (utils/prediction_tools.py:87-109, :133-156, :767-776) for the tile rows
 * [tile_row_begin, tile_row_end) of the chip grid (-1 end = all).
 * hwc: H*W*C mosaic.  out_prob (H*W fp32) and out_mask (H*W u8, may be NULL)
 * are full-size rasters; only the kept cores of the processed tile rows are
 * written (assigned), everything else is left untouched.  out_channel selects
 * the probability channel stitched (0 as at :154; 1 as at :267). */
SCV_API int scv_predict_mosaic(scv_engine* e, const void* hwc, int dtype, int H, int W, int C,
                       const scv_tiling* tiling, const scv_norm* norm, int tile_row_begin,
                       int tile_row_end, int out_channel, float* out_prob, uint8_t* out_mask);

/* Patch-list geometry (make_array_predictions / callback_predictions /
 * write_geotiff_predictions, utils/prediction_tools.py:245-373, :475-520):
 * N patches of (kernel+buff)^2, patch i kept core -> row i/cols, col i%cols of
 * the (rows*kernel, cols*kernel) output raster. */
SCV_API int scv_predict_patches(scv_engine* e, const void* nhwc, int dtype, int N, int H, int W, int C,
                        const scv_tiling* tiling, const scv_norm* norm, int cols, int out_channel,
                        float* out_prob, uint8_t* out_mask);

/* scv_predict_mosaic with chip-range sharding, float64 / accumulating output and a valid window (see
 * scv_mosaic_opts).  out_prob is H*W elements of opts->out_dtype.  Page-locked host buffers
 * (scv_host_alloc) give full H2D / compute / D2H overlap; pageable ones work through staged copies that are
 * interleaved with the kernel launches (scv_set_option "host_register" = 1 page-locks them for the call instead). */
SCV_API int scv_predict_mosaic_ex(scv_engine* e, const void* hwc, int dtype, int H, int W, int C,
                          const scv_tiling* tiling, const scv_norm* norm, const scv_mosaic_opts* opts,
                          void* out_prob, uint8_t* out_mask);

/* Streaming form of scv_predict_mosaic_ex for sustained multi-scene throughput (BASELINE configs[4]): returns
 * as soon as the scene's copies and kernels are enqueued; up to two scenes are in flight per engine, scene i+1's
 * H2D overlapping scene i's compute and scene i's D2H overlapping scene i+1's compute.  hwc / out_prob /
 * out_mask must be page-locked (scv_host_alloc) and stay valid until scv_stream_wait(ticket) returns. */
SCV_API int scv_stream_submit(scv_engine* e, const void* hwc, int dtype, int H, int W, int C,
                      const scv_tiling* tiling, const scv_norm* norm, const scv_mosaic_opts* opts, void* out_prob,
                      uint8_t* out_mask, int* ticket);
/* Blocks until scene `ticket` (ticket < 0: every submitted scene) is complete in host memory. */
SCV_API int scv_stream_wait(scv_engine* e, int ticket);

/* scv_predict_patches with an arbitrary crop window (non-square kernel_buffer, utils/prediction_tools.py:258-261). */
SCV_API int scv_predict_patches_ex(scv_engine* e, const void* nhwc, int dtype, int N, int H, int W, int C,
                           const scv_crop* crop, const scv_norm* norm, int cols, int out_channel, float* out_prob,
                           uint8_t* out_mask);

/* ---- predict: device-resident buffers (bench `value`, multi-GPU sharding) --- */

/* Same as scv_predict_mosaic but d_hwc / d_prob / d_mask are DEVICE pointers.
 * d_hwc points at mosaic row `src_row0` (only rows needed by the tile rows must
 * be resident); d_prob / d_mask point at output row `dst_row0`.
 * stream: a cudaStream_t to run on (NULL = the engine's own stream; the call
 * then synchronises before returning). */
SCV_API int scv_predict_mosaic_device(scv_engine* e, const void* d_hwc, int dtype, int H, int W, int C,
                              int src_row0, const scv_tiling* tiling, const scv_norm* norm,
                              int tile_row_begin, int tile_row_end, int out_channel, float* d_prob,
                              uint8_t* d_mask, int dst_row0, void* stream);

/* ... with scv_mosaic_opts (chip-range sharding; out_dtype / accumulate apply to d_prob). */
SCV_API int scv_predict_mosaic_device_ex(scv_engine* e, const void* d_hwc, int dtype, int H, int W, int C,
                                 int src_row0, const scv_tiling* tiling, const scv_norm* norm,
                                 const scv_mosaic_opts* opts, void* d_prob, uint8_t* d_mask, int dst_row0,
                                 void* stream);

SCV_API int scv_get_times(scv_engine* e, scv_times* out);
/* Reads (and clears) the engine's device-side error word: SCV_ERR_KERNEL if a watchdog tripped in any launch
 * since the last check -- for callers of the stream-taking device entry points, which return before the
 * kernels have run. */
SCV_API int scv_check(scv_engine* e);

/* pinned host memory for overlap-capable H2D/D2H (cudaHostAlloc) */
SCV_API void* scv_host_alloc(size_t bytes);
SCV_API void scv_host_free(void* p);

/* ---- kernel-level entry points used by the parity tests --------------------- */

/* One fused conv3x3('same') + bias + optional ReLU layer through the UMMA
 * implicit-GEMM kernel.  x: N*H*W*Cin fp32 host (rounded to bf16 on upload),
 * kernel HWIO fp32 (rounded to bf16), bias fp32.  y: N*H*W*Cout fp32 host
 * (bf16 results widened).  pooled (optional): N*(H/2)*(W/2)*Cout 2x2 max-pool of
 * the same outputs (fused epilogue). */
SCV_API int scv_debug_conv3x3(int device, const float* x, int N, int H, int W, int Cin, const float* kernel,
                      const float* bias, int Cout, int relu, float* y, float* pooled);

/* Conv2DTranspose(k=2,s=2) + bias + optional ReLU; kernel (2,2,Cout,Cin).
 * y: N*(2H)*(2W)*Cout. */
SCV_API int scv_debug_convT2x2(int device, const float* x, int N, int H, int W, int Cin, const float* kernel,
                       const float* bias, int Cout, int relu, float* y);

/* K1 alone: gather + normalise the chips of a mosaic (or pass tiles through)
 * into bf16 NHWC tiles (channels padded to Cpad, zeros); returned widened to
 * fp32: n_tiles*side*side*Cpad.  indices: n_tiles (y,x) pairs as produced by
 * generate_chip_indices. */
SCV_API int scv_debug_extract(int device, const void* hwc, int dtype, int H, int W, int C,
                      const scv_tiling* tiling, const scv_norm* norm, const int32_t* indices_yx,
                      int n_tiles, float* tiles_out, int* cpad_out);

#ifdef __cplusplus
}
#endif
#endif /* SCV_H_ */
