// libscv.so -- engine + C-ABI of the B200-native tiled U-Net predict path.
// See include/scv.h for the boundary and the reference call sites it replaces.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/scv.h"
#include "conv_rows.cuh"
#include "conv_fused.cuh"
#include "conv_slab2.cuh"
#include "conv_slabw.cuh"
#include "conv_umma.cuh"
#include "tile_kernels.cuh"

using namespace scv;

// ============================================================== error handling
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(SCV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define SCV_TRY(expr)       \
  do {                      \
    int _r = (expr);        \
    if (_r != SCV_OK) return _r; \
  } while (0)

// ============================================================== tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : (kc == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
}

// bf16 NHWC activation tensor (N,H,W,Cpitch) viewed as 4-D (C, W, H, N); box (KC, TW, TH, TN).
static int make_act_tmap(CUtensorMap* tm, const void* ptr, int N, int H, int W, int Cpitch, int KC, int TW, int TH,
                         int TN) {
  auto fn = get_encode_fn();
  if (!fn) return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)Cpitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Cpitch * 2, (cuuint64_t)W * Cpitch * 2, (cuuint64_t)H * W * Cpitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(KC), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled(act N=%d H=%d W=%d C=%d box=%d,%d,%d,%d) -> %d", N, H, W, Cpitch,
                KC, TW, TH, TN, (int)r);
  return SCV_OK;
}

// bf16 weights [Ntotal][Ktotal] (K contiguous) viewed as 2-D (K, N); box (KC, BN).
static int make_w_tmap(CUtensorMap* tm, const void* ptr, int Ktotal, int Ntotal, int KC, int BN) {
  auto fn = get_encode_fn();
  if (!fn) return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)Ktotal, (cuuint64_t)Ntotal};
  cuuint64_t strides[1] = {(cuuint64_t)Ktotal * 2};
  cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled(weights K=%d N=%d box=%d,%d) -> %d", Ktotal, Ntotal, KC, BN,
                (int)r);
  return SCV_OK;
}

// bf16 NHWC output tensor maps for the slab kernel's per-warp TMA stores (box = CB channels x 8 px x 4 rows).
static int make_store_tmap(CUtensorMap* tm, const void* ptr, int N, int H, int W, int Cpitch, int CB, int bw, int bh) {
  auto fn = get_encode_fn();
  if (!fn) return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)Cpitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Cpitch * 2, (cuuint64_t)W * Cpitch * 2, (cuuint64_t)H * W * Cpitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled(store map) -> %d", (int)r);
  return SCV_OK;
}
// Conv2DTranspose output (N, 2H, 2W, Cpitch) viewed as 5-D (C, b, x, a, n*H + y): pixel (2y+a, 2x+b).
static int make_convt_store_tmap(CUtensorMap* tm, const void* ptr, int N, int H, int W, int Cpitch, int CB) {
  auto fn = get_encode_fn();
  if (!fn) return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t px = (cuuint64_t)Cpitch * 2;
  cuuint64_t dims[5] = {(cuuint64_t)Cpitch, 2, (cuuint64_t)W, 2, (cuuint64_t)N * H};
  cuuint64_t strides[4] = {px, 2 * px, 2 * (cuuint64_t)W * px, 4 * (cuuint64_t)W * px};
  cuuint32_t box[5] = {(cuuint32_t)CB, 1, 8, 1, 4};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCV_ERR_CUDA, "cuTensorMapEncodeTiled(convT store map) -> %d", (int)r);
  return SCV_OK;
}

// ============================================================== architecture
static int pad_channels(int c) {
  if (c <= 8) return 8;  // one pixel = one 16-byte store; first layer runs the KC == 8 no-swizzle slab path
  if (c <= 16) return 16;
  if (c <= 32) return 32;
  return (c + 63) / 64 * 64;
}
// channel chunk of a layer's K loop; a 96-channel tensor (the siamese decoder entry at 32 filters) runs as three chunks of 32
static int kc_for(int cpad) { return cpad <= 8 ? 8 : (cpad <= 16 ? 16 : ((cpad <= 32 || cpad % 64) ? 32 : 64)); }
// K extent of a layer's weight matrix: taps * Cin_pad, except the 8-channel first layer (9 taps -> ten 8-wide slices)
static size_t k_total(int kc, int ntaps, int cin_pad) { return kc == 8 ? 80 : (size_t)ntaps * cin_pad; }
static int bn_for(int ntotal) {
  for (int bn : {256, 128, 64, 32})
    if (ntotal % bn == 0) return bn;
  return 0;
}
static void tile_box(int W, int* TW, int* TH, int* TN) {
  if (W % 16 == 0) {
    *TW = 16, *TH = 8, *TN = 1;
  } else if (W % 8 == 0 || W > 12) {
    *TW = 8, *TH = 8, *TN = 2;
  } else {
    *TW = 4, *TH = 4, *TN = 8;
  }
}

struct WeightSpec {
  std::string name;
  int ndim;
  int64_t shape[4];
};

static int validate_config(const scv_config* c) {
  if (!c) return fail(SCV_ERR_INVALID, "config is NULL");
  if (c->nlevels < 1 || c->nlevels > SCV_MAX_LEVELS) return fail(SCV_ERR_INVALID, "nlevels %d out of range", c->nlevels);
  if (c->nchannels < 1 || c->nchannels > SCV_MAX_BANDS)
    return fail(SCV_ERR_INVALID, "nchannels %d out of range [1,%d]", c->nchannels, SCV_MAX_BANDS);
  if (c->nclasses < 1 || c->nclasses > SCV_MAX_CLASSES)
    return fail(SCV_ERR_INVALID, "nclasses %d out of range [1,%d]", c->nclasses, SCV_MAX_CLASSES);
  if (c->head == SCV_HEAD_SIGMOID && c->nclasses != 1) return fail(SCV_ERR_INVALID, "sigmoid head needs nclasses == 1");
  if (c->head != SCV_HEAD_SIGMOID && c->head != SCV_HEAD_SOFTMAX) return fail(SCV_ERR_INVALID, "unknown head %d", c->head);
  for (int i = 0; i < c->nlevels; ++i) {
    const int f = c->filters[i];
    if (f < 32 || f % 32 != 0 || (f > 32 && f % 64 != 0))
      return fail(SCV_ERR_INVALID, "filters[%d]=%d unsupported: must be 32 or a multiple of 64", i, f);
  }
  if (c->filters[0] > 128) return fail(SCV_ERR_INVALID, "filters[0]=%d > 128 unsupported by the fused head", c->filters[0]);
  if (c->arch != SCV_ARCH_UNET && c->arch != SCV_ARCH_SIAMESE) return fail(SCV_ERR_INVALID, "unknown arch %d", c->arch);
  if (c->arch == SCV_ARCH_SIAMESE) {
    if (2 * c->nchannels > SCV_MAX_BANDS)
      return fail(SCV_ERR_INVALID, "siamese: 2 x nchannels = %d exceeds %d bands", 2 * c->nchannels, SCV_MAX_BANDS);
    if (c->double_conv) return fail(SCV_ERR_INVALID, "siamese: encoder blocks are single conv-BN-ReLU (conv_block.call as written)");
    if (c->head != SCV_HEAD_SIGMOID) return fail(SCV_ERR_INVALID, "siamese: the head is Conv2D(1, 1x1, sigmoid) (model_tools.py:659)");
  }
  return SCV_OK;
}
// bands of the input rasters / tiles: the siamese network takes its two images stacked along the channel axis
static int input_channels(const scv_config* c) { return c->arch == SCV_ARCH_SIAMESE ? 2 * c->nchannels : c->nchannels; }

static void build_specs_siamese(const scv_config* c, std::vector<WeightSpec>* out);
// keras model.get_weights() order (see oracle/unet.py weight_specs for the same walk)
static void build_specs(const scv_config* c, std::vector<WeightSpec>* out) {
  if (c->arch == SCV_ARCH_SIAMESE) return build_specs_siamese(c, out);
  auto add = [&](const std::string& n, std::initializer_list<int64_t> s) {
    WeightSpec w;
    w.name = n;
    w.ndim = (int)s.size();
    int i = 0;
    for (auto v : s) w.shape[i++] = v;
    for (; i < 4; ++i) w.shape[i] = 1;
    out->push_back(w);
  };
  auto conv = [&](const std::string& n, int cin, int cout, int k) {
    add(n + "/kernel", {k, k, cin, cout});
    add(n + "/bias", {cout});
  };
  auto bn = [&](const std::string& n, int ch) {
    add(n + "/gamma", {ch});
    add(n + "/beta", {ch});
    add(n + "/moving_mean", {ch});
    add(n + "/moving_variance", {ch});
  };
  const int nconv = c->double_conv ? 2 : 1;
  int cin = c->nchannels;
  for (int i = 0; i < c->nlevels; ++i) {
    for (int j = 0; j < nconv; ++j) {
      const std::string b = "encoder_" + std::to_string(i);
      conv(b + "/conv" + std::to_string(j), cin, c->filters[i], 3);
      bn(b + "/bn" + std::to_string(j), c->filters[i]);
      cin = c->filters[i];
    }
  }
  const int fc = c->filters[c->nlevels - 1] * 2;
  for (int j = 0; j < nconv; ++j) {
    conv("center/conv" + std::to_string(j), cin, fc, 3);
    bn("center/bn" + std::to_string(j), fc);
    cin = fc;
  }
  for (int i = c->nlevels - 1; i >= 0; --i) {
    const int f = c->filters[i];
    const std::string b = "decoder_" + std::to_string(i);
    add(b + "/up/kernel", {2, 2, f, cin});
    add(b + "/up/bias", {f});
    bn(b + "/bn_cat", 2 * f);
    conv(b + "/conv0", 2 * f, f, 3);
    bn(b + "/bn0", f);
    conv(b + "/conv1", f, f, 3);
    bn(b + "/bn1", f);
    cin = f;
  }
  conv("head", cin, c->nclasses, 1);
}

// make_siamese_unet / get_siamese_layers / DilatedSpatialPyramidPooling (utils/model_tools.py:533-663), per conv-BN unit
// in (kernel, bias, gamma, beta, moving_mean, moving_variance) order (oracle/siamese.py walks the same list; the Python
// layer maps tf.keras' trainable-first order of the ASPP layer onto it).
static void build_specs_siamese(const scv_config* c, std::vector<WeightSpec>* out) {
  auto add = [&](const std::string& n, std::initializer_list<int64_t> s) {
    WeightSpec w;
    w.name = n;
    w.ndim = (int)s.size();
    int i = 0;
    for (auto v : s) w.shape[i++] = v;
    for (; i < 4; ++i) w.shape[i] = 1;
    out->push_back(w);
  };
  auto conv = [&](const std::string& n, int cin, int cout, int k) {
    add(n + "/kernel", {k, k, cin, cout});
    add(n + "/bias", {cout});
  };
  auto bn = [&](const std::string& n, int ch) {
    add(n + "/gamma", {ch});
    add(n + "/beta", {ch});
    add(n + "/moving_mean", {ch});
    add(n + "/moving_variance", {ch});
  };
  const int L = c->nlevels;
  int cin = c->nchannels;
  for (int i = 0; i < L; ++i) {
    const std::string b = "encoder_" + std::to_string(i);
    conv(b + "/conv0", cin, c->filters[i], 3);
    bn(b + "/bn0", c->filters[i]);
    cin = c->filters[i];
  }
  const int nf = 2 * c->filters[L - 1];
  conv("ASPP/cba/conv", cin, nf, 1);
  bn("ASPP/cba/bn", nf);
  conv("ASPP/cba3/conv", 4 * nf, nf, 1);
  bn("ASPP/cba3/bn", nf);
  for (int r : {3, 6, 12}) {
    const std::string b = "ASPP/cba3_" + std::to_string(r);
    conv(b + "/conv", cin, nf, 3);
    bn(b + "/bn", nf);
  }
  cin = 2 * nf;
  for (int i = L - 1; i >= 0; --i) {
    const int f = c->filters[i];
    const std::string b = "decoder_" + std::to_string(i);
    add(b + "/up/kernel", {2, 2, f, cin});
    add(b + "/up/bias", {f});
    bn(b + "/bn_cat", 3 * f);
    conv(b + "/conv0", 3 * f, f, 3);
    bn(b + "/bn0", f);
    conv(b + "/conv1", f, f, 3);
    bn(b + "/bn1", f);
    cin = f;
  }
  conv("head", cin, c->nclasses, 1);
}

// ---- layer graph -----------------------------------------------------------
enum { L_CONV3 = 0, L_CONVT = 1, L_CONV1 = 2 };

struct BufDef {
  int level;     // resolution = tile >> level
  int channels;  // channel pitch
  bool f32;      // logits
  std::string name;
  int mult = 1;  // images per tile (2: the siamese network's A half then B half)
};

struct LayerDef {
  std::string name;
  int kind, epi;
  int level;            // input resolution level
  int cin_real, cin_pad;  // input channels consumed (pitch of in_buf == cin_pad)
  int cout;             // conv: output channels; convT: F (Ntotal = 4F)
  int ntotal;
  int KC, BN;
  int in_buf, out_buf, out_choff, pool_buf;
  // keras weight-list indices (-1 = none)
  int w_kernel, w_bias, w_bn;  // bn: index of gamma (gamma,beta,mean,var consecutive)
  int w_skip_bn;               // EPI_POOL_SKIP: decoder bn_cat gamma index (first F channels)
  int w_up_bn;                 // L_CONVT: decoder bn_cat gamma index (channels [F,2F))
  double flops_per_tile_at_unit;  // FLOPs per input pixel of this layer (multiply by h*w of its level)
  // siamese network (all zero / one for the plain U-Net)
  int dil = 1;         // dilation rate of a 3x3 conv (ASPP: 3, 6, 12)
  int cin_off = 0;     // first input channel the kernel applies to (the two images are stacked along the channel axis)
  int skip_off = 0;    // EPI_POOL_SKIP: first channel of this layer's slice of the decoder's bn_cat
  int up_off = -1;     // L_CONVT: first channel of the up slice of bn_cat (-1: cout, the plain U-Net's [F, 2F))
  int nmult = 1;       // images per tile this launch covers (2: both halves of a mult-2 buffer as one batch)
  int in_half = 0, out_half = 0, pool_half = 0;  // which half of a mult-2 buffer the launch reads / writes
  // device weights
  __nv_bfloat16* d_w = nullptr;
  float* d_bias = nullptr;
  float* d_skip_s = nullptr;
  float* d_skip_t = nullptr;
  CUtensorMap tmB;
  // 8-channel first layer, row-kernel form: the same weights as [Cout][tap * 16 + c] (channels 8..15 zero); the
  // activation keeps its 8-channel pitch and TMA zero-fills the upper half of every 16-channel box
  __nv_bfloat16* d_w16 = nullptr;
};

struct Arch {
  scv_config cfg;
  std::vector<WeightSpec> specs;
  std::map<std::string, int> spec_index;
  std::vector<BufDef> bufs;
  std::vector<LayerDef> layers;
  int x0_buf, logits_buf;
  int c0pad;
};

static int build_arch_siamese(const scv_config* c, Arch* a);
static int build_arch(const scv_config* c, Arch* a) {
  SCV_TRY(validate_config(c));
  a->cfg = *c;
  build_specs(c, &a->specs);
  for (size_t i = 0; i < a->specs.size(); ++i) a->spec_index[a->specs[i].name] = (int)i;
  if (c->arch == SCV_ARCH_SIAMESE) return build_arch_siamese(c, a);
  auto idx = [&](const std::string& n) { return a->spec_index.at(n); };
  auto add_buf = [&](int level, int ch, bool f32, const std::string& n) {
    a->bufs.push_back({level, ch, f32, n});
    return (int)a->bufs.size() - 1;
  };
  const int L = c->nlevels;
  const int nconv = c->double_conv ? 2 : 1;
  a->c0pad = pad_channels(c->nchannels);
  a->x0_buf = add_buf(0, a->c0pad, false, "x0");
  std::vector<int> t(L, -1), cat(L), pooled(L);
  for (int i = 0; i < L; ++i) {
    t[i] = add_buf(i, c->filters[i], false, "t" + std::to_string(i));  // enc conv0 out (variant A) / dec conv0 out
    cat[i] = add_buf(i, 2 * c->filters[i], false, "cat" + std::to_string(i));
    pooled[i] = add_buf(i + 1, c->filters[i], false, "pool" + std::to_string(i));
  }
  const int fc = c->filters[L - 1] * 2;
  const int ct1 = nconv == 2 ? add_buf(L, fc, false, "center0") : -1;
  const int ct2 = add_buf(L, fc, false, "center1");
  std::vector<int> d2(L, -1);
  for (int i = 1; i < L; ++i) d2[i] = add_buf(i, c->filters[i], false, "dec" + std::to_string(i));
  a->logits_buf = add_buf(0, c->nclasses, true, "logits");

  auto conv_layer = [&](const std::string& name, int level, int cin_real, int in_buf, int cout) {
    LayerDef l{};
    l.name = name;
    l.kind = L_CONV3;
    l.epi = EPI_STORE;
    l.level = level;
    l.cin_real = cin_real;
    l.cin_pad = a->bufs[in_buf].channels;
    l.cout = cout;
    l.ntotal = cout;
    l.KC = kc_for(l.cin_pad);
    l.BN = bn_for(cout);
    l.in_buf = in_buf;
    l.out_buf = -1;
    l.out_choff = 0;
    l.pool_buf = -1;
    l.w_kernel = idx(name + "/kernel");
    l.w_bias = idx(name + "/bias");
    l.w_bn = -1;
    l.w_skip_bn = l.w_up_bn = -1;
    l.flops_per_tile_at_unit = 2.0 * cin_real * cout * 9;
    return l;
  };

  int cur = a->x0_buf, cur_real = c->nchannels;
  for (int i = 0; i < L; ++i) {
    const std::string b = "encoder_" + std::to_string(i);
    for (int j = 0; j < nconv; ++j) {
      LayerDef l = conv_layer(b + "/conv" + std::to_string(j), i, cur_real, cur, c->filters[i]);
      l.w_bn = idx(b + "/bn" + std::to_string(j) + "/gamma");
      if (j == nconv - 1) {
        l.epi = EPI_POOL_SKIP;
        l.out_buf = cat[i];
        l.out_choff = 0;
        l.pool_buf = pooled[i];
        l.w_skip_bn = idx("decoder_" + std::to_string(i) + "/bn_cat/gamma");
      } else {
        l.out_buf = t[i];
      }
      a->layers.push_back(l);
      cur = (j == nconv - 1) ? pooled[i] : t[i];
      cur_real = c->filters[i];
    }
  }
  for (int j = 0; j < nconv; ++j) {
    LayerDef l = conv_layer("center/conv" + std::to_string(j), L, cur_real, cur, fc);
    l.w_bn = idx("center/bn" + std::to_string(j) + "/gamma");
    l.out_buf = (j == nconv - 1) ? ct2 : ct1;
    a->layers.push_back(l);
    cur = l.out_buf;
    cur_real = fc;
  }
  for (int i = L - 1; i >= 0; --i) {
    const int f = c->filters[i];
    const std::string b = "decoder_" + std::to_string(i);
    LayerDef u{};
    u.name = b + "/up";
    u.kind = L_CONVT;
    u.epi = EPI_CONVT;
    u.level = i + 1;
    u.cin_real = cur_real;
    u.cin_pad = a->bufs[cur].channels;
    u.cout = f;
    u.ntotal = 4 * f;
    u.KC = kc_for(u.cin_pad);
    u.BN = bn_for(u.ntotal);
    u.in_buf = cur;
    u.out_buf = cat[i];
    u.out_choff = f;
    u.pool_buf = -1;
    u.w_kernel = idx(b + "/up/kernel");
    u.w_bias = idx(b + "/up/bias");
    u.w_bn = -1;
    u.w_skip_bn = -1;
    u.w_up_bn = idx(b + "/bn_cat/gamma");
    u.flops_per_tile_at_unit = 2.0 * cur_real * f * 4;
    a->layers.push_back(u);

    LayerDef c0 = conv_layer(b + "/conv0", i, 2 * f, cat[i], f);
    c0.w_bn = idx(b + "/bn0/gamma");
    c0.out_buf = t[i];
    a->layers.push_back(c0);

    LayerDef c1 = conv_layer(b + "/conv1", i, f, t[i], f);
    c1.w_bn = idx(b + "/bn1/gamma");
    if (i == 0) {
      c1.epi = EPI_HEAD;
      c1.out_buf = a->logits_buf;
      c1.flops_per_tile_at_unit += 2.0 * f * c->nclasses;  // fused 1x1 head
    } else {
      c1.out_buf = d2[i];
    }
    a->layers.push_back(c1);
    cur = c1.out_buf;
    cur_real = f;
  }
  for (auto& l : a->layers) {
    if (l.BN == 0) return fail(SCV_ERR_INVALID, "layer %s: N=%d has no supported tile width", l.name.c_str(), l.ntotal);
    if (l.cin_pad % l.KC != 0) return fail(SCV_ERR_INVALID, "layer %s: Cin pad %d vs KC %d", l.name.c_str(), l.cin_pad, l.KC);
  }
  if ((int)a->layers.size() > SCV_MAX_LAYERS) return fail(SCV_ERR_INVALID, "too many layers");
  return SCV_OK;
}

// Siamese U-Net with an atrous pyramid between encoder and decoder: make_siamese_unet / get_siamese_layers /
// DilatedSpatialPyramidPooling (utils/model_tools.py:533-663).  The two images arrive stacked along the channel axis
// ([a | b], so every tiled-predict entry point works unchanged); the shared encoder runs once per image (the first
// conv reads the stacked tile with its kernel embedded at channel offset 0 / C), every later encoder conv reads the A /
// B half of a two-images-per-tile buffer.  Concatenations cost nothing: the encoder epilogues write relu(bn_cat(.)) of
// their output straight into channel slices [0, F) (image b) / [F, 2F) (image a) of the decoder-entry tensor, the
// transposed conv into [2F, 3F); the four ASPP branches write the four slices of one 4 nf-channel tensor.
static int build_arch_siamese(const scv_config* c, Arch* a) {
  auto idx = [&](const std::string& n) { return a->spec_index.at(n); };
  auto add_buf = [&](int level, int ch, bool f32, const std::string& n, int mult) {
    BufDef b{level, ch, f32, n};
    b.mult = mult;
    a->bufs.push_back(b);
    return (int)a->bufs.size() - 1;
  };
  const int L = c->nlevels, C = c->nchannels;
  a->c0pad = pad_channels(2 * C);
  a->x0_buf = add_buf(0, a->c0pad, false, "x0", 1);
  std::vector<int> cat(L), pooled(L), t(L), d2(L, -1);
  for (int i = 0; i < L; ++i) {
    // 3 F channels; F = 32 gives 96 = three 32-channel chunks (padding it to 128 cost the 384-pixel decoder conv its
    // row-kernel form: 30.9 ms in the N = 32 slab kernel against 13.4 ms; the 192-byte pixel pitch costs the two
    // encoder epilogues and the transposed conv that write into it 2 ms each, net -13 ms per scene)
    cat[i] = add_buf(i, 3 * c->filters[i] % 32 == 0 ? 3 * c->filters[i] : pad_channels(3 * c->filters[i]), false,
                     "cat" + std::to_string(i), 1);
    pooled[i] = add_buf(i + 1, c->filters[i], false, "pool" + std::to_string(i), 2);
    t[i] = add_buf(i, c->filters[i], false, "t" + std::to_string(i), 1);
    if (i > 0) d2[i] = add_buf(i, c->filters[i], false, "dec" + std::to_string(i), 1);
  }
  const int FL = c->filters[L - 1], nf = 2 * FL;
  const int aspp_cat = add_buf(L, 4 * nf, false, "aspp_cat", 2);
  const int squeezed = add_buf(L, 2 * nf, false, "squeezed", 1);
  a->logits_buf = add_buf(0, c->nclasses, true, "logits", 1);

  auto conv_layer = [&](const std::string& wname, const std::string& name, int kind, int level, int cin_real, int in_buf,
                        int cout) {
    LayerDef l{};
    l.name = name;
    l.kind = kind;
    l.epi = EPI_STORE;
    l.level = level;
    l.cin_real = cin_real;
    l.cin_pad = a->bufs[in_buf].channels;
    l.cout = cout;
    l.ntotal = cout;
    l.KC = kc_for(l.cin_pad);
    l.BN = bn_for(cout);
    l.in_buf = in_buf;
    l.out_buf = -1;
    l.out_choff = 0;
    l.pool_buf = -1;
    l.w_kernel = idx(wname + "/kernel");
    l.w_bias = idx(wname + "/bias");
    l.w_bn = -1;
    l.w_skip_bn = l.w_up_bn = -1;
    l.flops_per_tile_at_unit = 2.0 * cin_real * cout * (kind == L_CONV3 ? 9 : 1);
    return l;
  };

  // shared encoder: image a (first C channels) then image b; net[i] = concat([encoded_b, encoded_a]) (:604, :609)
  for (int i = 0; i < L; ++i) {
    const std::string b = "encoder_" + std::to_string(i);
    const int F = c->filters[i];
    for (int half = 0; half < 2; ++half) {
      LayerDef l = conv_layer(b + "/conv0", b + (half ? "/conv0[b]" : "/conv0[a]"), L_CONV3, i, i == 0 ? C : c->filters[i - 1],
                              i == 0 ? a->x0_buf : pooled[i - 1], F);
      l.w_bn = idx(b + "/bn0/gamma");
      l.epi = EPI_POOL_SKIP;
      l.out_buf = cat[i];
      l.out_choff = half ? 0 : F;
      l.skip_off = l.out_choff;
      l.w_skip_bn = idx("decoder_" + std::to_string(i) + "/bn_cat/gamma");
      l.pool_buf = pooled[i];
      l.pool_half = half;
      if (i == 0) l.cin_off = half ? C : 0;
      else l.in_half = half;
      a->layers.push_back(l);
    }
  }
  // ASPP on both pooled images as one batch (:561-573): 1x1 and three dilated 3x3 branches -> concat -> 1x1
  {
    LayerDef l = conv_layer("ASPP/cba/conv", "ASPP/cba", L_CONV1, L, FL, pooled[L - 1], nf);
    l.w_bn = idx("ASPP/cba/bn/gamma");
    l.out_buf = aspp_cat;
    l.nmult = 2;
    l.flops_per_tile_at_unit *= 2;
    a->layers.push_back(l);
    int k = 1;
    for (int r : {3, 6, 12}) {
      const std::string n = "ASPP/cba3_" + std::to_string(r);
      LayerDef d = conv_layer(n + "/conv", n, L_CONV3, L, FL, pooled[L - 1], nf);
      d.w_bn = idx(n + "/bn/gamma");
      d.dil = r;
      d.out_buf = aspp_cat;
      d.out_choff = k++ * nf;
      d.nmult = 2;
      d.flops_per_tile_at_unit *= 2;
      a->layers.push_back(d);
    }
    for (int half = 0; half < 2; ++half) {  // squeezed = concat([aspp_b, aspp_a]) (:620)
      LayerDef o = conv_layer("ASPP/cba3/conv", half ? "ASPP/cba3[b]" : "ASPP/cba3[a]", L_CONV1, L, 4 * nf, aspp_cat, nf);
      o.w_bn = idx("ASPP/cba3/bn/gamma");
      o.in_half = half;
      o.out_buf = squeezed;
      o.out_choff = half ? 0 : nf;
      a->layers.push_back(o);
    }
  }
  int cur = squeezed, cur_real = 2 * nf;
  for (int i = L - 1; i >= 0; --i) {
    const int f = c->filters[i];
    const std::string b = "decoder_" + std::to_string(i);
    LayerDef u{};
    u.name = b + "/up";
    u.kind = L_CONVT;
    u.epi = EPI_CONVT;
    u.level = i + 1;
    u.cin_real = cur_real;
    u.cin_pad = a->bufs[cur].channels;
    u.cout = f;
    u.ntotal = 4 * f;
    u.KC = kc_for(u.cin_pad);
    u.BN = bn_for(u.ntotal);
    u.in_buf = cur;
    u.out_buf = cat[i];
    u.out_choff = 2 * f;
    u.up_off = 2 * f;
    u.pool_buf = -1;
    u.w_kernel = idx(b + "/up/kernel");
    u.w_bias = idx(b + "/up/bias");
    u.w_bn = -1;
    u.w_skip_bn = -1;
    u.w_up_bn = idx(b + "/bn_cat/gamma");
    u.flops_per_tile_at_unit = 2.0 * cur_real * f * 4;
    a->layers.push_back(u);

    LayerDef c0 = conv_layer(b + "/conv0", b + "/conv0", L_CONV3, i, 3 * f, cat[i], f);
    c0.w_bn = idx(b + "/bn0/gamma");
    c0.out_buf = t[i];
    a->layers.push_back(c0);

    LayerDef c1 = conv_layer(b + "/conv1", b + "/conv1", L_CONV3, i, f, t[i], f);
    c1.w_bn = idx(b + "/bn1/gamma");
    if (i == 0) {
      c1.epi = EPI_HEAD;
      c1.out_buf = a->logits_buf;
      c1.flops_per_tile_at_unit += 2.0 * f * c->nclasses;
    } else {
      c1.out_buf = d2[i];
    }
    a->layers.push_back(c1);
    cur = c1.out_buf;
    cur_real = f;
  }
  for (auto& l : a->layers) {
    if (l.BN == 0) return fail(SCV_ERR_INVALID, "layer %s: N=%d has no supported tile width", l.name.c_str(), l.ntotal);
    if (l.cin_pad % l.KC != 0) return fail(SCV_ERR_INVALID, "layer %s: Cin pad %d vs KC %d", l.name.c_str(), l.cin_pad, l.KC);
  }
  if ((int)a->layers.size() > SCV_MAX_LAYERS) return fail(SCV_ERR_INVALID, "too many layers");
  return SCV_OK;
}

// ============================================================== engine
struct Plan {
  int B, H, W;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<void*> buf_ptr;
  std::vector<ConvLaunch> launches;
};

struct scv_engine {
  Arch arch;
  int device = 0;
  bool weights_set = false;
  std::vector<float> h_head_w, h_head_b;
  float* d_head_w = nullptr;
  float* d_head_b = nullptr;
  int* d_err = nullptr;
  cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;
  std::vector<std::unique_ptr<Plan>> plans;
  // host-buffer mosaic calls: two scene slots so that streamed scenes overlap (scv_stream_submit)
  struct Slot {
    void* d_scene = nullptr;
    size_t scene_bytes = 0;
    void* d_prob = nullptr;
    size_t prob_bytes = 0;
    uint8_t* d_mask = nullptr;
    size_t mask_bytes = 0;
    cudaEvent_t compute_done = nullptr, d2h_done = nullptr;
    std::vector<cudaEvent_t> up_ev, row_ev;  // per tile row: H2D landed / kernels done (reused across scenes)
    std::vector<void*> registered;           // host ranges page-locked for the scene in flight
    int ticket = -1;
    bool busy = false;
  };
  Slot slots[2];
  int next_ticket = 0;
  // scratch
  float* d_prob = nullptr;  // patch-list rasters
  size_t prob_bytes = 0;
  uint8_t* d_mask = nullptr;
  size_t mask_bytes = 0;
  int2* d_origins = nullptr;
  size_t origins_cap = 0;
  std::vector<int2> h_origins;
  long long origins_key[8] = {};
  bool origins_valid = false;
  float* d_tile_stats = nullptr;
  size_t tile_stats_cap = 0;
  void* d_stage = nullptr;  // raw tiles staging for predict_tiles / predict_patches
  size_t stage_bytes = 0;
  float* d_tile_probs = nullptr;
  size_t tile_probs_bytes = 0;
  int32_t* d_tile_classes = nullptr;
  size_t tile_classes_bytes = 0;
  // super-batch tensors shared by every plan: K1 output (layer-0 input) and head logits for up to tile_cap tiles
  __nv_bfloat16* d_x0_all = nullptr;
  float* d_logits_all = nullptr;
  int tile_cap = 0, tile_cap_side = 0;
  // options
  int opt_profile_layers = 0;
  int opt_stages = 0;
  int opt_watchdog_ms = 2000;
  int opt_host_super_tiles = 0;    // ... when the scene streams in from host memory (H2D / compute / D2H overlap);
                                   // 0 = one device batch per K1 / K4 launch: compute starts after ~3 tile rows of H2D
  int opt_host_first_row = 1;      // host-buffer mosaic calls: first super-batch = first tile row (short pipeline start)
  int opt_host_register = 0;       // page-lock pageable caller buffers for the duration of a host-buffer mosaic call
                                   // (measured: cudaHostRegister of a scene's 2.4 GB costs 1.3 s, staged copies 0.4 s -> off)
  int opt_super_tiles = 2048;  // target tiles per K1 / K4 launch (a full 10980^2 scene = 1764 chips: x0 4.2 GB + logits 1.0 GB)
  // timing
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct BatchEv {
    int e0, e1, e2, e3;
    std::vector<int> layer_ev;
    int ntiles;
  };
  std::vector<BatchEv> batch_ev;
  int n_launches = 0;
  int ev_host_begin = -1, ev_host_end = -1;  // host-buffer calls: first H2D enqueued (copy stream) / last D2H landed
  int last_side = 0;
  bool times_pending = false;
  scv_times times{};
};

static int dtype_bytes(int dt) {
  switch (dt) {
    case SCV_U8: return 1;
    case SCV_U16:
    case SCV_I16: return 2;
    case SCV_F32: return 4;
    case SCV_F64: return 8;
  }
  return 0;
}

static int ensure(void** p, size_t* cap, size_t need) {
  if (*cap >= need && *p) return SCV_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CUDA_TRY(cudaMalloc(p, need));
  *cap = need;
  return SCV_OK;
}

static int check_device_err(scv_engine* e) {
  int h = 0;
  CUDA_TRY(cudaMemcpy(&h, e->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (h) {
    cudaMemset(e->d_err, 0, sizeof(int));
    return fail(SCV_ERR_KERNEL, "device watchdog tripped: a conv pipeline stalled for > %d ms", e->opt_watchdog_ms);
  }
  return SCV_OK;
}

// ---- weights ---------------------------------------------------------------
static const float BN_EPS = 1e-3f;  // keras BatchNormalization default

static int upload_layer(LayerDef& l, const std::vector<float>& w /*[ntotal][ktotal]*/, const std::vector<float>& bias,
                        const std::vector<float>* skip_s, const std::vector<float>* skip_t) {
  const int ntaps = l.kind == L_CONV3 ? 9 : 1;
  const size_t ktotal = k_total(l.KC, ntaps, l.cin_pad);
  std::vector<__nv_bfloat16> wb(w.size());
  for (size_t i = 0; i < w.size(); ++i) wb[i] = __float2bfloat16_rn(w[i]);
  if (l.d_w) cudaFree(l.d_w);
  if (l.d_bias) cudaFree(l.d_bias);
  l.d_w = nullptr;
  l.d_bias = nullptr;
  CUDA_TRY(cudaMalloc(&l.d_w, wb.size() * 2));
  CUDA_TRY(cudaMemcpy(l.d_w, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&l.d_bias, bias.size() * 4));
  CUDA_TRY(cudaMemcpy(l.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  if (skip_s) {
    if (l.d_skip_s) cudaFree(l.d_skip_s);
    if (l.d_skip_t) cudaFree(l.d_skip_t);
    CUDA_TRY(cudaMalloc(&l.d_skip_s, skip_s->size() * 4));
    CUDA_TRY(cudaMalloc(&l.d_skip_t, skip_t->size() * 4));
    CUDA_TRY(cudaMemcpy(l.d_skip_s, skip_s->data(), skip_s->size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(l.d_skip_t, skip_t->data(), skip_t->size() * 4, cudaMemcpyHostToDevice));
  }
  if (l.KC == 8 && l.kind == L_CONV3) {
    std::vector<__nv_bfloat16> w16((size_t)l.ntotal * 144, __float2bfloat16_rn(0.f));
    for (int o = 0; o < l.ntotal; ++o)
      for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < 8; ++c) w16[(size_t)o * 144 + tap * 16 + c] = wb[(size_t)o * ktotal + tap * 8 + c];
    if (l.d_w16) cudaFree(l.d_w16);
    l.d_w16 = nullptr;
    CUDA_TRY(cudaMalloc(&l.d_w16, w16.size() * 2));
    CUDA_TRY(cudaMemcpy(l.d_w16, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
  }
  return make_w_tmap(&l.tmB, l.d_w, (int)ktotal, l.ntotal, l.KC, l.BN);
}

// BN fold: s = gamma / sqrt(var + eps), t = beta - mean * s  (fp32, like keras inference)
static void bn_affine(const scv_tensor* t, int gamma_idx, int offset, int n, std::vector<float>* s, std::vector<float>* sh) {
  s->resize(n);
  sh->resize(n);
  const float *g = t[gamma_idx].data + offset, *b = t[gamma_idx + 1].data + offset, *m = t[gamma_idx + 2].data + offset,
              *v = t[gamma_idx + 3].data + offset;
  for (int i = 0; i < n; ++i) {
    const float sc = g[i] / std::sqrt(v[i] + BN_EPS);
    (*s)[i] = sc;
    (*sh)[i] = b[i] - m[i] * sc;
  }
}

static int fold_and_upload(scv_engine* e, const scv_tensor* t) {
  Arch& a = e->arch;
  for (auto& l : a.layers) {
    const int ntaps = l.kind == L_CONV3 ? 9 : 1;
    const size_t ktotal = k_total(l.KC, ntaps, l.cin_pad);
    std::vector<float> w((size_t)l.ntotal * ktotal, 0.f), bias(l.ntotal, 0.f);
    const float* K = t[l.w_kernel].data;
    const float* Bv = t[l.w_bias].data;
    if (l.kind == L_CONV3 || l.kind == L_CONV1) {
      std::vector<float> s, sh;
      bn_affine(t, l.w_bn, 0, l.cout, &s, &sh);
      // keras HWIO (k,k,Cin,Cout) -> [o][tap*cin_pad + cin_off + c] with W' = W*s[o]
      for (int tap = 0; tap < ntaps; ++tap)
        for (int c = 0; c < l.cin_real; ++c) {
          const float* src = K + ((size_t)tap * l.cin_real + c) * l.cout;
          for (int o = 0; o < l.cout; ++o) w[(size_t)o * ktotal + (size_t)tap * l.cin_pad + l.cin_off + c] = src[o] * s[o];
        }
      for (int o = 0; o < l.cout; ++o) bias[o] = Bv[o] * s[o] + sh[o];  // (b - mean)*s + beta
      std::vector<float> ks, kt;
      if (l.epi == EPI_POOL_SKIP) {
        bn_affine(t, l.w_skip_bn, l.skip_off, l.cout, &ks, &kt);
        SCV_TRY(upload_layer(l, w, bias, &ks, &kt));
      } else {
        SCV_TRY(upload_layer(l, w, bias, nullptr, nullptr));
      }
    } else {
      // keras Conv2DTranspose kernel (2,2,F,Cin) -> rows (a*2+b)*F + o, K = Cin; the BN that follows the
      // concat is folded on the up half: W' = W*s2[o], b' = b*s2[o] + t2[o]
      const int F = l.cout;
      std::vector<float> s2, t2;
      bn_affine(t, l.w_up_bn, l.up_off >= 0 ? l.up_off : F, F, &s2, &t2);
      for (int ab = 0; ab < 4; ++ab)
        for (int o = 0; o < F; ++o) {
          const float* src = K + ((size_t)ab * F + o) * l.cin_real;
          float* dst = &w[((size_t)ab * F + o) * ktotal];
          for (int c = 0; c < l.cin_real; ++c) dst[c] = src[c] * s2[o];
          bias[(size_t)ab * F + o] = Bv[o] * s2[o] + t2[o];
        }
      SCV_TRY(upload_layer(l, w, bias, nullptr, nullptr));
    }
  }
  // 1x1 head: kernel (1,1,F0,ncls) -> [c][k] fp32
  const int hk = a.spec_index.at("head/kernel");
  const int F0 = a.cfg.filters[0], ncls = a.cfg.nclasses;
  e->h_head_w.assign(t[hk].data, t[hk].data + (size_t)F0 * ncls);
  e->h_head_b.assign(t[hk + 1].data, t[hk + 1].data + ncls);
  if (!e->d_head_w) CUDA_TRY(cudaMalloc(&e->d_head_w, (size_t)F0 * ncls * 4));
  if (!e->d_head_b) CUDA_TRY(cudaMalloc(&e->d_head_b, (size_t)ncls * 4));
  CUDA_TRY(cudaMemcpy(e->d_head_w, e->h_head_w.data(), (size_t)F0 * ncls * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(e->d_head_b, e->h_head_b.data(), (size_t)ncls * 4, cudaMemcpyHostToDevice));
  return SCV_OK;
}

// ---- plans -------------------------------------------------------------------
static int pick_stages(int KC, int BN, int iters, int override_stages) {
  if (override_stages > 0) return std::max(1, std::min(override_stages, 12));
  const int sb = conv_stage_bytes(KC, BN);
  int st = (100 * 1024) / sb;  // ~2 CTAs per SM
  st = std::max(2, std::min(st, 8));
  return std::max(1, std::min(st, iters));
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

static int sm_count() {  // of the CURRENT device (engines on different GPUs may share a process)
  static int cache[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

// Persistent weight-stationary slab kernel: used where the layer's weights (all taps) fit in shared
// memory next to a few halo slabs and there are enough 8x16 M tiles to keep every SM busy.
static const size_t kSlabSmemBudget = 222 * 1024;
static const size_t kSlabWeightBudget = 150 * 1024;

// Measured on B200 (tools/microbench/umma_rate.cu): an SS-operand 128 x N x 16 bf16 UMMA sustains one
// issue per ~{46, 54, 70, 134} cycles for N = {32, 64, 128, 256} (operand fetch ~115 B/clk from smem).
static int umma_cycles(int bn) { return bn <= 32 ? 46 : (bn <= 64 ? 54 : (bn <= 128 ? 70 : 134)); }

static bool plan_slab(const LayerDef& l, int B, int h, int w, int* bn_out, int* nslab, int* nacc_out) {
  if (!env_int("SCV_SLAB", 1) && l.KC != 8) return false;
  if (w % 8 || h % 16) return false;
  const int ntaps = l.kind == L_CONV3 ? 9 : 1;
  const long long m_tiles = (long long)B * (h / 16) * (w / 8);
  const bool force = env_int("SCV_SLAB_FORCE", 0) != 0 || l.KC == 8;  // the 8-channel first layer has no tile-kernel form
  const int slab_stride = slab_stride_bytes(l.KC, ntaps);
  const int chunks = l.cin_pad / l.KC;
  for (int bn : {256, 128, 64, 32}) {
    if (l.ntotal % bn) continue;
    if (l.epi == EPI_HEAD && bn != l.ntotal) continue;
    const int n_tiles_n = l.ntotal / bn;
    // persistent CTAs need enough tiles each, and re-reading A once per N tile must stay cheap
    if (!force && (n_tiles_n > 2 || m_tiles * n_tiles_n < 8LL * sm_count())) continue;
    const size_t wbytes = slab_weight_bytes(l.KC, bn, ntaps, l.cin_pad);
    if (wbytes > kSlabWeightBudget) continue;
    // epilogue budget: with few MMAs per tile the epilogue warps must turn tiles around fast -> 4 groups
    const long long mma_cycles = (long long)(k_total(l.KC, ntaps, l.cin_pad) / 16) * umma_cycles(bn);
    const int first = (mma_cycles < 3000 && slab_nacc_ok(bn, 4)) ? 4 : 2;
    for (int nacc : {first, first == 4 ? 2 : 4}) {
      if (!slab_nacc_ok(bn, nacc)) continue;
      const size_t fixed = wbytes + slab_stage_bytes(l.epi, nacc) + 6 * 1024;
      const int need = std::max(2, chunks);  // at least one whole tile's worth of slabs
      if (fixed + (size_t)need * slab_stride > kSlabSmemBudget) continue;
      int ns = std::min(8, (int)((kSlabSmemBudget - fixed) / slab_stride));
      if (const int o = env_int("SCV_SLAB_STAGES", 0)) ns = std::max(need, std::min(ns, o));
      // two issuer warps need every slab slot to come back to the same issuer (see conv_slab_kernel)
      if (ns >= 2 * chunks) ns = ns / (2 * chunks) * (2 * chunks);
      *bn_out = bn;
      *nslab = ns;
      *nacc_out = nacc;
      return true;
    }
  }
  return false;
}

// CTA-pair slab kernel (conv_slab2.cuh, tcgen05.mma.cta_group::2): the Cout = 64 3x3 layers that would run in the slab
// kernel with N = 64 (the 192 x 192 level).  Same (chunk, tap, k) summation order -> same bits as the other 8x16-tile
// kernels.  Measured per layer (full scene, tools/r02_exp31.sh): decoder_1/conv0 (Cin 128) 11.4 -> 7.8 ms -- the weights
// resident per SM halve, so the slab ring covers two tiles and there are four accumulators instead of two --,
// decoder_1/conv1 4.4 -> 3.95 ms; the K-short encoder_1/conv0 (KC 32) 2.5 -> 3.3 ms and the pooling encoder_1/conv1
// 4.8 -> 5.9 ms are SLOWER (the pair's tile turn-around is the slower of two epilogues plus a remote arrive), so by
// default only KC = 64 layers with the plain store epilogue take it.  SCV_SLAB2=0: off; =2: every eligible layer, any
// launch size (tests); =3: every eligible layer at the default size threshold (experiments).
static bool plan_slab2(const LayerDef& l, int B, int h, int w, int* nslab, int* nacc_out, int* pairs_out) {
  const int mode = env_int("SCV_SLAB2", 1);
  if (!mode) return false;
  if (l.kind != L_CONV3 || l.ntotal != kSlab2BN || l.cout != kSlab2BN) return false;
  if (l.KC != 32 && l.KC != 64) return false;
  if (l.epi != EPI_STORE && l.epi != EPI_POOL_SKIP) return false;
  if (mode == 1 && (l.KC != 64 || l.epi != EPI_STORE)) return false;
  if (w % 8 || h % 16) return false;
  const long long m_tiles = (long long)B * (h / 16) * (w / 8);
  if (m_tiles & 1) return false;
  if (mode != 2 && m_tiles < 8LL * sm_count()) return false;
  const int chunks = l.cin_pad / l.KC;
  const int slab_stride = slab_stride_bytes(l.KC, 9);
  for (int nacc : {4, 2}) {
    const size_t fixed = slab2_weight_bytes(l.cin_pad) + slab_stage_bytes(l.epi, nacc) + 6 * 1024;
    const int need = std::max(2, chunks);
    if (fixed + (size_t)need * slab_stride > kSlabSmemBudget) continue;
    int ns = std::min(8, (int)((kSlabSmemBudget - fixed) / slab_stride));
    if (const int o = env_int("SCV_SLAB_STAGES", 0)) ns = std::max(need, std::min(ns, o));
    if (ns >= 2 * chunks) ns = ns / (2 * chunks) * (2 * chunks);  // two issuers: see conv_slab_kernel
    const int pairs = conv_slab2_max_pairs(slab2_smem_bytes(l.KC, l.cin_pad, ns, l.epi, nacc), nacc);
    if (pairs < 8) return false;  // cluster launch not available
    *nslab = ns;
    *nacc_out = nacc;
    *pairs_out = pairs;
    return true;
  }
  return false;
}

// Row-streaming tap-packed kernel (conv_rows.cuh): 3x3 layers with Cout 32/64 on rows that split into
// 128-pixel strips.  The choice depends on the layer geometry only -- never on the batch size -- so a
// tile's result does not depend on how many tiles ran with it.
static bool plan_rows(const LayerDef& l, int h, int w, int ncls, int* nslab) {
  if (!env_int("SCV_ROWS", 1)) return false;
  if (l.kind != L_CONV3 || (h & 1) || (w & 1)) return false;
  if (w % kRowsPx) {
    // partial last strip (192-pixel rows = 1.5 strips; SCV_ROWS_PARTIAL=1): supported -- TMA zero-fills the loads
    // and clips the stores -- and bit-identical to the 8x16-tile kernels, but measured no faster on the Cout = 64
    // layers at 192 x 192 (enc1.c2 4.92 vs 4.93 ms, dec1.c2 4.42 vs 4.45 ms: 25 % of the M lanes idle), so off
    const int strips = (w + kRowsPx - 1) / kRowsPx;
    if (!env_int("SCV_ROWS_PARTIAL", 0) || w < kRowsPx || 10 * w < 7 * strips * kRowsPx) return false;
  }
  const bool first = l.KC == 8 && l.d_w16 != nullptr && env_int("SCV_ROWS_FIRST", 1);  // 8 stored channels read as 16
  if (l.KC != 32 && l.KC != 64 && !first) return false;
  if (l.ntotal != l.cout || (l.cout != 32 && l.cout != 64)) return false;
  if (l.epi != EPI_STORE && l.epi != EPI_POOL_SKIP && l.epi != EPI_HEAD) return false;
  const int kc = first ? 16 : l.KC, cin = first ? 16 : l.cin_pad;
  const int chunks = cin / kc;
  for (int ns = std::max(6, 4 * chunks); ns >= 2 * chunks && ns >= 3; --ns)  // slabs of two input rows each
    if (rows_smem_bytes(kc, l.cout, cin, ns, l.epi, ncls) <= kSlabSmemBudget) {
      *nslab = ns;
      return true;
    }
  return false;
}

static int fill_launch(scv_engine* e, const LayerDef& l, const void* in_ptr, int in_pitch, int B, int h, int w,
                       ConvLaunch* L, int n_in = 0) {
  if (n_in <= 0) n_in = B;  // images in the input tensor (>= B when the input is a slice of a larger tensor)
  memset(L, 0, sizeof *L);
  ConvParams& p = L->p;
  p.N = B;
  p.H = h;
  p.W = w;
  p.Cin = l.cin_pad;
  p.ntaps = l.kind == L_CONV3 ? 9 : 1;
  p.relu = 1;
  p.bias = l.d_bias;
  p.Cout = l.cout;
  p.skip_s = l.d_skip_s;
  p.skip_t = l.d_skip_t;
  p.err = e ? e->d_err : nullptr;
  p.watchdog_ns = (unsigned long long)(e ? e->opt_watchdog_ms : 2000) * 1000000ull;
  L->KC = l.KC;
  L->EPI = l.epi;
  int bn = 0, ns = 0, nacc = 2;
  p.dil = l.dil;
  // dilated 3x3 and 1x1 convolutions (ASPP) exist in the tile kernels only: a tap is a shifted TMA box there
  const bool generic_only = l.dil != 1 || l.kind == L_CONV1;
  if (!generic_only && plan_rows(l, h, w, e ? e->arch.cfg.nclasses : 1, &ns)) {
    const bool first = l.KC == 8;  // 8-channel input: boxes of 16 channels, the upper 8 zero-filled by TMA (out of bounds)
    const int kc = first ? 16 : l.KC, cin = first ? 16 : l.cin_pad;
    L->slab = 2;
    L->KC = kc;
    p.Cin = cin;
    p.dbg = env_int("SCV_ROWS_DBG", 0);
    L->BN = l.cout;
    L->nacc = rows_epi_groups(l.cout, l.epi);
    p.TW = kRowsPx, p.TH = 1, p.TN = 1;
    p.tiles_x = (w + kRowsPx - 1) / kRowsPx;
    p.tiles_y = h;
    p.tiles_n = B;
    p.n_tiles_n = 1;
    p.nslab = ns;
    // issuers in flight must not exceed the reuse distance (in row pairs) of a slab slot or an accumulator pair
    p.n_issuers = std::max(1, std::min({kRowsIssuers, ns / (cin / kc), 256 / l.cout,
                                        env_int("SCV_ROWS_ISSUERS", kRowsIssuers)}));
    const long long pairs = (long long)B * p.tiles_x * (h / 2);
    L->grid = (int)std::min<long long>(sm_count(), pairs);
    L->smem = rows_smem_bytes(kc, l.cout, cin, ns, l.epi, e ? e->arch.cfg.nclasses : 1);
    if (first) SCV_TRY(make_w_tmap(&L->tmB, l.d_w16, 144, l.ntotal, 16, l.cout));
    else if (l.BN != l.cout) SCV_TRY(make_w_tmap(&L->tmB, l.d_w, (int)k_total(l.KC, 9, l.cin_pad), l.ntotal, l.KC, l.cout));
    else L->tmB = l.tmB;
    SCV_TRY(make_act_tmap(&L->tmA, in_ptr, n_in, h, w, in_pitch, kc, kRowsSlabPx, 2, 1));
    return SCV_OK;
  }
  // weight-streaming halo-slab kernel: Cout = 128 layers whose 3x3 weights exceed shared memory (conv_slabw.cuh);
  // same summation order as the slab / tile kernels, so the choice may depend on the batch size
  const bool slabw_n64 = l.ntotal == 64 && l.cin_pad >= 128 && env_int("SCV_SLABW_N64", 0);  // measured slower (14.4 vs 12.0 ms): off
  if (!generic_only && env_int("SCV_SLABW", 1) && l.kind == L_CONV3 && l.KC == 64 && (l.ntotal == 128 || slabw_n64) &&
      l.cin_pad >= env_int("SCV_SLABW_MIN_CIN", 64) && (l.epi == EPI_STORE || l.epi == EPI_POOL_SKIP) && w % 8 == 0 && h % 16 == 0 &&
      ((long long)B * (h / 16) * (w / 8) >= 4LL * sm_count() || env_int("SCV_SLABW", 1) == 2)) {
    const int wn = l.ntotal;
    int nbr = wn == 64 ? 10 : 6;
    while (nbr > 2 && slabw_smem_bytes(64, wn, nbr, l.epi) > kSlabSmemBudget) --nbr;
    L->slab = 4;
    L->BN = wn;
    L->nacc = 4;
    p.TW = 8, p.TH = 16, p.TN = 1;
    p.tiles_x = w / 8;
    p.tiles_y = h / 16;
    p.tiles_n = B;
    p.num_m_tiles = p.tiles_x * p.tiles_y * B;
    p.n_tiles_n = 1;
    p.nstage = nbr;
    p.n_issuers = std::max(1, std::min(kSlabwIssuers, env_int("SCV_SLABW_ISSUERS", kSlabwIssuers)));
    L->grid = (int)std::min<long long>(sm_count(), p.num_m_tiles);
    L->smem = slabw_smem_bytes(64, wn, nbr, l.epi);
    if (l.BN != wn) SCV_TRY(make_w_tmap(&L->tmB, l.d_w, (int)k_total(l.KC, 9, l.cin_pad), l.ntotal, l.KC, wn));
    else L->tmB = l.tmB;
    SCV_TRY(make_act_tmap(&L->tmA, in_ptr, n_in, h, w, in_pitch, l.KC, 10, 18, 1));
    return SCV_OK;
  }
  int pairs = 0;
  if (!generic_only && plan_slab2(l, B, h, w, &ns, &nacc, &pairs)) {
    L->nacc = nacc;
    L->slab = 6;
    L->BN = kSlab2BN;
    p.TW = 8, p.TH = 16, p.TN = 1;
    p.tiles_x = w / 8;
    p.tiles_y = h / 16;
    p.tiles_n = B;
    p.num_m_tiles = p.tiles_x * p.tiles_y * B;
    p.n_tiles_n = 1;
    p.nslab = ns;
    p.n_issuers = (ns % (2 * (l.cin_pad / l.KC)) == 0) ? 2 : 1;
    L->grid = 2 * (int)std::min<long long>({(long long)pairs, (long long)sm_count() / 2, (long long)p.num_m_tiles / 2});
    L->smem = slab2_smem_bytes(l.KC, l.cin_pad, ns, l.epi, nacc);
    SCV_TRY(make_w_tmap(&L->tmB, l.d_w, (int)k_total(l.KC, 9, l.cin_pad), l.ntotal, l.KC, kSlab2BH));
    SCV_TRY(make_act_tmap(&L->tmA, in_ptr, n_in, h, w, in_pitch, l.KC, 10, 18, 1));
    return SCV_OK;
  }
  if (!generic_only && plan_slab(l, B, h, w, &bn, &ns, &nacc)) {
    L->nacc = nacc;
    const int sw = p.ntaps == 9 ? 10 : 8, sh = p.ntaps == 9 ? 18 : 16;
    L->slab = 1;
    L->BN = bn;
    p.TW = 8, p.TH = 16, p.TN = 1;
    p.tiles_x = w / 8;
    p.tiles_y = h / 16;
    p.tiles_n = B;
    p.num_m_tiles = p.tiles_x * p.tiles_y * B;
    p.n_tiles_n = l.ntotal / bn;
    p.nslab = ns;
    p.n_issuers = (ns % (2 * (l.cin_pad / l.KC)) == 0) ? 2 : 1;
    int grid = sm_count() / p.n_tiles_n * p.n_tiles_n;
    const long long work = (long long)p.num_m_tiles * p.n_tiles_n;
    if (grid > work) grid = (int)work;
    L->grid = std::max(grid, p.n_tiles_n);
    L->smem = slab_smem_bytes(l.KC, bn, p.ntaps, l.cin_pad, ns, l.epi, e ? e->arch.cfg.nclasses : 1, nacc);
    if (l.BN != bn) SCV_TRY(make_w_tmap(&L->tmB, l.d_w, (int)k_total(l.KC, p.ntaps, l.cin_pad), l.ntotal, l.KC, bn));
    else L->tmB = l.tmB;
    SCV_TRY(make_act_tmap(&L->tmA, in_ptr, n_in, h, w, in_pitch, l.KC, sw, sh, 1));
    return SCV_OK;
  }
  if (l.KC == 8) return fail(SCV_ERR_INVALID, "layer %s: 8-channel input needs tile sides that are multiples of 16", l.name.c_str());
  L->slab = 0;
  L->BN = l.BN;
  tile_box(w, &p.TW, &p.TH, &p.TN);
  p.tiles_x = (w + p.TW - 1) / p.TW;
  p.tiles_y = (h + p.TH - 1) / p.TH;
  p.tiles_n = (B + p.TN - 1) / p.TN;
  p.n_tiles_n = l.ntotal / l.BN;
  const int iters = p.ntaps * (l.cin_pad / l.KC);
  p.nstage = pick_stages(l.KC, l.BN, iters, e ? e->opt_stages : 0);
  L->grid = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles_n;
  L->tmB = l.tmB;
  L->smem = conv_smem_bytes(l.KC, l.BN, p.nstage, l.epi, e ? e->arch.cfg.nclasses : 1);
  // enough tiles for every SM to loop: persistent form (same K order per tile -> same bits)
  const int ncls = e ? e->arch.cfg.nclasses : 1;
  const int sms = sm_count() / p.n_tiles_n * p.n_tiles_n;
  if (env_int("SCV_PTILE", 1) && sms > 0 && (L->grid >= 2 * sms || env_int("SCV_PTILE", 1) == 2)) {
    int ns = 8;
    while (ns > 2 && ptile_smem_bytes(l.KC, l.BN, ns, l.epi, ncls) > kSlabSmemBudget) --ns;
    if (!(e && e->opt_stages > 0)) p.nstage = std::min(ns, std::max(2, iters * 2));
    else p.nstage = std::max(2, std::min(ns, e->opt_stages));
    p.ncls = ncls;
    p.n_issuers = std::max(1, std::min({kPtileIssuers, p.nstage, 2 * iters, env_int("SCV_PTILE_ISSUERS", kPtileIssuers)}));
    L->slab = 3;
    L->grid = std::min(L->grid, sms);
    L->smem = ptile_smem_bytes(l.KC, l.BN, p.nstage, l.epi, ncls);
  }
  SCV_TRY(make_act_tmap(&L->tmA, in_ptr, n_in, h, w, in_pitch, l.KC, p.TW, p.TH, p.TN));
  return SCV_OK;
}

// TMA-store maps of a slab launch's bf16 outputs (call once the output pointers are known).
static int finish_slab_maps(ConvLaunch* L, const LayerDef& l) {
  if (!L->slab || L->slab == 3 || l.epi == EPI_HEAD) return SCV_OK;
  const ConvParams& p = L->p;
  const int cb = 32;  // kStageRowB / 2 channels per store box
  if (L->slab == 2) {  // row kernel: one warp stores 32 pixels of two rows
    SCV_TRY(make_store_tmap(&L->tmOut, p.out, p.N, p.H, p.W, p.out_pitch, cb, 32, 2));
    if (l.epi == EPI_POOL_SKIP)
      SCV_TRY(make_store_tmap(&L->tmPool, p.pool_out, p.N, p.H / 2, p.W / 2, p.pool_pitch, cb, 16, 1));
    return SCV_OK;
  }
  if (l.epi == EPI_CONVT) return make_convt_store_tmap(&L->tmOut, p.out, p.N, p.H, p.W, p.out_pitch, cb);
  SCV_TRY(make_store_tmap(&L->tmOut, p.out, p.N, p.H, p.W, p.out_pitch, cb, 8, 4));
  if (l.epi == EPI_POOL_SKIP)
    SCV_TRY(make_store_tmap(&L->tmPool, p.pool_out, p.N, p.H / 2, p.W / 2, p.pool_pitch, cb, 4, 2));
  return SCV_OK;
}

static int get_plan(scv_engine* e, int B, int H, int W, Plan** out) {
  for (auto& pl : e->plans)
    if (pl->B == B && pl->H == H && pl->W == W) {
      *out = pl.get();
      return SCV_OK;
    }
  Arch& a = e->arch;
  const int L = a.cfg.nlevels;
  if (H % (1 << L) || W % (1 << L) || H <= 0 || W <= 0)
    return fail(SCV_ERR_INVALID, "tile %dx%d: H and W must be positive multiples of %d", H, W, 1 << L);
  // keep at most 3 plans alive (full batch, tail batch, one spare)
  while (e->plans.size() >= 3) {
    cudaStreamSynchronize(e->stream);
    cudaFree(e->plans.front()->arena);
    e->plans.erase(e->plans.begin());
  }
  auto pl = std::make_unique<Plan>();
  pl->B = B;
  pl->H = H;
  pl->W = W;
  std::vector<size_t> off(a.bufs.size());
  size_t total = 0;
  for (size_t i = 0; i < a.bufs.size(); ++i) {
    const BufDef& b = a.bufs[i];
    size_t bytes = (size_t)B * b.mult * (H >> b.level) * (W >> b.level) * b.channels * (b.f32 ? 4 : 2);
    if ((int)i == a.x0_buf || (int)i == a.logits_buf) bytes = 0;  // live in the engine's super-batch tensors
    off[i] = total;
    total += (bytes + 1023) & ~size_t(1023);
  }
  CUDA_TRY(cudaMalloc(&pl->arena, std::max<size_t>(total, 1024)));
  pl->arena_bytes = total;
  // channel padding of a concat tensor (siamese: 3 F -> a multiple of 64) is read by the next conv against zero weights:
  // it must hold finite values
  if (a.cfg.arch == SCV_ARCH_SIAMESE) CUDA_TRY(cudaMemset(pl->arena, 0, std::max<size_t>(total, 1024)));
  pl->buf_ptr.resize(a.bufs.size());
  for (size_t i = 0; i < a.bufs.size(); ++i) pl->buf_ptr[i] = pl->arena + off[i];
  if (!e->d_x0_all || e->tile_cap < B || e->tile_cap_side != H)
    return fail(SCV_ERR_STATE, "internal: super-batch tensors not sized before planning");
  pl->buf_ptr[a.x0_buf] = e->d_x0_all;
  pl->buf_ptr[a.logits_buf] = e->d_logits_all;
  for (auto& l : a.layers) {
    ConvLaunch Ln;
    const int h = H >> l.level, w = W >> l.level;
    // the layer that reads x0 sees the whole super-batch tensor (tile_cap images) and is offset per launch
    const int n_in = l.in_buf == a.x0_buf ? e->tile_cap : B * l.nmult;
    // half h of a two-images-per-tile buffer = images [h B, (h + 1) B)
    auto half_ptr = [&](int buf, int half) {
      const BufDef& b = a.bufs[buf];
      return (uint8_t*)pl->buf_ptr[buf] + (size_t)half * B * (H >> b.level) * (W >> b.level) * b.channels * (b.f32 ? 4 : 2);
    };
    SCV_TRY(fill_launch(e, l, half_ptr(l.in_buf, l.in_half), a.bufs[l.in_buf].channels, B * l.nmult, h, w, &Ln, n_in));
    ConvParams& p = Ln.p;
    if (l.epi == EPI_HEAD) {
      p.head_w = e->d_head_w;
      p.head_b = e->d_head_b;
      p.ncls = a.cfg.nclasses;
      p.logits = (float*)pl->buf_ptr[l.out_buf];
    } else {
      p.out = (__nv_bfloat16*)half_ptr(l.out_buf, l.out_half);
      p.out_pitch = a.bufs[l.out_buf].channels;
      p.out_choff = l.out_choff;
      if (l.epi == EPI_POOL_SKIP) {
        p.pool_out = (__nv_bfloat16*)half_ptr(l.pool_buf, l.pool_half);
        p.pool_pitch = a.bufs[l.pool_buf].channels;
      }
    }
    if (Ln.slab == 1 && l.epi == EPI_CONVT && env_int("SCV_LSU_CONVT", 0)) p.linear_out = 1;  // slab kernel, transposed conv (experiment: slower at 128-byte pitch)
    SCV_TRY(finish_slab_maps(&Ln, l));
    if (env_int("SCV_PLAN_DEBUG", 0))
      fprintf(stderr, "[scv plan B=%d] %-18s %dx%d Cin=%d N=%d  %s KC=%d BN=%d %s=%d nacc=%d grid=%d smem=%zu\n", B,
              l.name.c_str(), h, w, l.cin_pad, l.ntotal, Ln.slab == 6 ? "slab2" : Ln.slab == 4 ? "slabw" : (Ln.slab == 3 ? "ptile" : (Ln.slab == 2 ? "rows" : (Ln.slab ? "slab" : "tile"))), Ln.KC, Ln.BN,
              (Ln.slab == 1 || Ln.slab == 2 || Ln.slab == 6) ? "nslab" : "nstage", (Ln.slab == 1 || Ln.slab == 2 || Ln.slab == 6) ? Ln.p.nslab : Ln.p.nstage,
              Ln.slab == 3 ? 2 : (Ln.slab ? Ln.nacc : 1), Ln.grid, Ln.smem);
    pl->launches.push_back(Ln);
  }
  // Fuse the two 32-channel conv pairs of the 384-pixel level into one cluster launch each (conv_fused.cuh) when both
  // halves run in the row kernel: decoder_0/conv0 -> conv1 + head (bit 0 of SCV_FUSE) and encoder_0/conv0 -> conv1 +
  // pool + skip (bit 1).  The 32-channel intermediates never go to HBM.  Bit-identical to the two launches.  Full
  // scene: 125.0 ms unfused, 122.9-123.4 with the decoder tail fused, 121.6-121.9 with both (tools/r02_exp21.sh).
  const int fuse = env_int("SCV_FUSE", 3);
  for (size_t i = 0; i + 1 < a.layers.size() && fuse; ++i) {
    const LayerDef &l1 = a.layers[i], &l2 = a.layers[i + 1];
    ConvLaunch &A = pl->launches[i], &Bn = pl->launches[i + 1];
    if (A.slab != 2 || Bn.slab != 2 || l1.epi != EPI_STORE || l1.cout != 32 || l2.cout != 32 || l2.cin_pad != 32 ||
        l2.KC != 32 || l2.in_buf != l1.out_buf || l1.level != 0 || l2.level != 0 || W != kF2Cluster * kRowsPx || (H & 1))
      continue;
    const bool tail = l2.epi == EPI_HEAD && l1.KC == 64 && l1.cin_pad == 64 && (fuse & 1);
    const bool head = l2.epi == EPI_POOL_SKIP && l1.KC == 8 && A.KC == 16 && (fuse & 2);
    if (!tail && !head) continue;
    const size_t smem = fused_smem_bytes(A.KC, l2.epi, a.cfg.nclasses);
    if (smem > 227 * 1024) continue;
    const int ncl = conv_fused_max_clusters(smem, l2.epi);
    if (ncl < 8) continue;  // cluster launch not available / too few co-resident clusters: keep the two launches
    A.slab = 5;
    A.EPI = l2.epi;
    A.tmB2 = Bn.tmB;
    A.tmOut = Bn.tmOut;
    A.tmPool = Bn.tmPool;
    A.p.bias2 = l2.d_bias;
    A.p.head_w = Bn.p.head_w;
    A.p.head_b = Bn.p.head_b;
    A.p.ncls = Bn.p.ncls;
    A.p.logits = Bn.p.logits;
    A.p.out = Bn.p.out;
    A.p.out_pitch = Bn.p.out_pitch;
    A.p.out_choff = Bn.p.out_choff;
    A.p.pool_out = Bn.p.pool_out;
    A.p.pool_pitch = Bn.p.pool_pitch;
    A.p.skip_s = Bn.p.skip_s;
    A.p.skip_t = Bn.p.skip_t;
    A.smem = smem;
    A.grid = kF2Cluster * (int)std::min<long long>(ncl, (long long)B * (H / 2));
    Bn.slab = -1;
    if (env_int("SCV_PLAN_DEBUG", 0))
      fprintf(stderr, "[scv plan B=%d] %s + %s fused: %d clusters of %d CTAs, smem=%zu\n", B, l1.name.c_str(), l2.name.c_str(),
              A.grid / kF2Cluster, kF2Cluster, smem);
  }
  *out = pl.get();
  e->plans.push_back(std::move(pl));
  return SCV_OK;
}

// Timing events are pooled per engine and re-recorded call after call (creating / destroying ~100 events per call
// showed up as host time at 8 ranks per box).
static int new_event(scv_engine* e, cudaStream_t s) {
  if (e->ev_used == e->ev_pool.size()) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return -1;
    e->ev_pool.push_back(ev);
  }
  cudaEventRecord(e->ev_pool[e->ev_used], s);
  return (int)e->ev_used++;
}

static void reset_timing(scv_engine* e) {
  e->ev_used = 0;
  e->batch_ev.clear();
  e->ev_host_begin = e->ev_host_end = -1;
  e->n_launches = 0;
  e->times_pending = false;
}

// Launches the network for `pl->B` tiles whose K1 output starts at image `tile_off` of the super-batch tensors.
static int run_layers(scv_engine* e, Plan* pl, int tile_off, int side, cudaStream_t s, scv_engine::BatchEv* bev) {
  const Arch& a = e->arch;
  for (size_t i = 0; i < pl->launches.size(); ++i) {
    if (e->opt_profile_layers && bev) bev->layer_ev.push_back(new_event(e, s));
    ConvLaunch L = pl->launches[i];
    const LayerDef& l = a.layers[i];
    if (l.in_buf == a.x0_buf) L.p.n_in_off = tile_off;
    if (l.epi == EPI_HEAD || (L.slab == 5 && L.EPI == EPI_HEAD)) L.p.logits = e->d_logits_all + (size_t)tile_off * side * side * a.cfg.nclasses;
    if (L.slab < 0) continue;  // folded into the previous (fused) launch
    cudaError_t err = conv_launch(L, s);
    if (err != cudaSuccess)
      return fail(SCV_ERR_CUDA, "launch of layer %s failed: %s", l.name.c_str(), cudaGetErrorString(err));
    e->n_launches++;
  }
  if (e->opt_profile_layers && bev) bev->layer_ev.push_back(new_event(e, s));
  return SCV_OK;
}

static int finalize_times(scv_engine* e) {
  if (!e->times_pending) return SCV_OK;
  scv_times& t = e->times;
  memset(&t, 0, sizeof t);
  t.n_layers = (int)e->arch.layers.size();
  for (size_t i = 0; i < e->arch.layers.size(); ++i) {
    const LayerDef& l = e->arch.layers[i];
    const double px = (double)(e->last_side >> l.level) * (e->last_side >> l.level);
    t.layer_flops[i] = l.flops_per_tile_at_unit * px;
  }
  if (!e->batch_ev.empty()) {
    CUDA_TRY(cudaEventSynchronize(e->ev_pool[e->batch_ev.back().e3]));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev_pool[e->batch_ev.front().e0], e->ev_pool[e->batch_ev.back().e3]));
    t.total_ms = ms;
    for (auto& b : e->batch_ev) {
      cudaEventElapsedTime(&ms, e->ev_pool[b.e0], e->ev_pool[b.e1]);
      t.extract_ms += ms;
      cudaEventElapsedTime(&ms, e->ev_pool[b.e1], e->ev_pool[b.e2]);
      t.network_ms += ms;
      cudaEventElapsedTime(&ms, e->ev_pool[b.e2], e->ev_pool[b.e3]);
      t.stitch_ms += ms;
      t.n_tiles += b.ntiles;
      const size_t per = e->arch.layers.size() + 1;  // events per network pass
      for (size_t k = 0; k + per <= b.layer_ev.size(); k += per)
        for (size_t i = 0; i + 1 < per && i < SCV_MAX_LAYERS; ++i) {
          cudaEventElapsedTime(&ms, e->ev_pool[b.layer_ev[k + i]], e->ev_pool[b.layer_ev[k + i + 1]]);
          t.layer_ms[i] += ms;
        }
    }
    t.n_batches = (int)e->batch_ev.size();
    if (e->ev_host_begin >= 0 && e->ev_host_end >= 0) {
      CUDA_TRY(cudaEventSynchronize(e->ev_pool[e->ev_host_end]));
      cudaEventElapsedTime(&t.h2d_lead_ms, e->ev_pool[e->ev_host_begin], e->ev_pool[e->batch_ev.front().e0]);
      cudaEventElapsedTime(&t.d2h_tail_ms, e->ev_pool[e->batch_ev.back().e3], e->ev_pool[e->ev_host_end]);
    }
  }
  t.n_launches = e->n_launches;
  e->times_pending = false;
  return SCV_OK;
}

// ---- shared batch pipeline ------------------------------------------------------
struct TileJob {
  // source (mosaic or stacked tiles, device resident)
  const void* d_src;
  size_t src_bytes;
  int dtype, src_W, C, src_row0;
  int side;
  const int2* d_src_origins;  // n tiles
  const scv_norm* norm;
  int valid[4];               // y0, y1, x0, x1 of the valid window (y1 <= y0: none)
  int order_kernel, tiles_per_row, order_skip0;  // regular chip grid: K1 block ordering (order_kernel == 0: off)
  // destination: stitched raster ...
  const int2* d_dst_origins;  // n tiles (null -> per-tile outputs)
  int kernel_h, kernel_w, crop_y, crop_x, out_W, dst_row0, out_channel, force_scalar;
  void* d_prob;
  int prob_f64, accumulate;
  uint8_t* d_mask;
  // ... or whole tiles
  float* d_tile_probs;
  int32_t* d_tile_classes;
};

static bool tile_stats_mode(int m) {
  return m == SCV_NORM_TILE_ZSCORE || m == SCV_NORM_TILE_MINMAX || m == SCV_NORM_TILE_GLOBAL_ZSCORE ||
         m == SCV_NORM_TILE_GLOBAL_MINMAX;
}

static int norm_to_params(const scv_norm* n, int C, ExtractParams* ep) {
  ep->norm_mode = n ? n->mode : SCV_NORM_NONE;
  ep->ngroups = 0;
  for (int c = 0; c < SCV_MAX_BANDS; ++c) {
    ep->sub[c] = 0.f;
    ep->div[c] = 1.f;
    ep->group_end[c] = 0;
  }
  if (!n || n->mode == SCV_NORM_NONE) return SCV_OK;
  if (n->mode < 0 || n->mode > SCV_NORM_TILE_GLOBAL_ZSCORE) return fail(SCV_ERR_INVALID, "unknown norm mode %d", n->mode);
  if (n->mode == SCV_NORM_PER_BAND) {
    if (n->nbands != C) return fail(SCV_ERR_INVALID, "norm has %d bands, input has %d", n->nbands, C);
    for (int c = 0; c < C; ++c) {
      ep->sub[c] = n->sub[c];
      ep->div[c] = n->div[c];
    }
    return SCV_OK;
  }
  ep->div[0] = n->div[0];  // epsilon
  if (n->ngroups < 0 || n->ngroups > SCV_MAX_BANDS) return fail(SCV_ERR_INVALID, "norm: ngroups %d out of range", n->ngroups);
  int end = 0;
  for (int g = 0; g < n->ngroups; ++g) {
    if (n->group_size[g] <= 0) return fail(SCV_ERR_INVALID, "norm: group %d has size %d", g, n->group_size[g]);
    end += n->group_size[g];
    ep->group_end[g] = end;
  }
  if (end > C) return fail(SCV_ERR_INVALID, "norm: channel groups cover %d bands, input has %d", end, C);
  ep->ngroups = n->ngroups;
  return SCV_OK;
}

// (Re)sizes the super-batch tensors; plans bake their addresses into tensor maps, so growing drops the plans.
static int ensure_tile_tensors(scv_engine* e, int ntiles, int side) {
  if (e->d_x0_all && e->tile_cap >= ntiles && e->tile_cap_side == side) return SCV_OK;
  const int cap = std::max(ntiles, e->tile_cap_side == side ? e->tile_cap : 0);
  CUDA_TRY(cudaDeviceSynchronize());  // callers may be on a user stream
  for (auto& pl : e->plans) cudaFree(pl->arena);
  e->plans.clear();
  cudaFree(e->d_x0_all);
  cudaFree(e->d_logits_all);
  e->d_x0_all = nullptr;
  e->d_logits_all = nullptr;
  e->tile_cap = 0;
  const size_t px = (size_t)cap * side * side;
  CUDA_TRY(cudaMalloc(&e->d_x0_all, px * e->arch.c0pad * 2));
  CUDA_TRY(cudaMalloc(&e->d_logits_all, px * e->arch.cfg.nclasses * 4));
  e->tile_cap = cap;
  e->tile_cap_side = side;
  return SCV_OK;
}

// Runs tiles [t0, t0+n) of the job as ONE super-batch on `s`: one K1 launch for all of them, the network in
// device batches of `sizes`, one K4 launch for all of them (the HBM-bound kernels need launches long enough to
// reach bandwidth; the conv kernels need batches small enough for the activation arena).
static int run_super(scv_engine* e, const TileJob& job, int t0, const std::vector<int>& sizes, cudaStream_t s) {
  int n = 0;
  for (int nb : sizes) n += nb;
  if (n == 0) return SCV_OK;
  SCV_TRY(ensure_tile_tensors(e, n, job.side));
  e->last_side = job.side;
  Arch& a = e->arch;
  scv_engine::BatchEv bev{};
  bev.ntiles = n;
  bev.e0 = new_event(e, s);

  ExtractParams ep{};
  ep.src = (const uint8_t*)job.d_src;
  ep.src_bytes = job.src_bytes;
  ep.dtype = job.dtype;
  ep.W = job.src_W;
  ep.C = job.C;
  ep.src_row0 = job.src_row0;
  ep.origins = job.d_src_origins + t0;
  ep.n_tiles = n;
  ep.side = job.side;
  ep.cpad = a.c0pad;
  SCV_TRY(norm_to_params(job.norm, job.C, &ep));
  ep.valid_y0 = job.valid[0], ep.valid_y1 = job.valid[1], ep.valid_x0 = job.valid[2], ep.valid_x1 = job.valid[3];
  if (job.order_kernel > 0) {
    ep.order_kernel = job.order_kernel;
    ep.tiles_per_row = job.tiles_per_row;
    ep.order_skip = (job.order_skip0 + t0) % job.tiles_per_row;
  }
  const int row_bytes = job.side * job.C * dtype_bytes(job.dtype);
  ep.rows_per_block = std::max(1, std::min(8, (44 * 1024) / (row_bytes + 32)));
  ep.out = e->d_x0_all;
  if (tile_stats_mode(ep.norm_mode)) {
    if (ep.valid_y1 > ep.valid_y0)
      return fail(SCV_ERR_INVALID, "per-tile statistics cannot be combined with a valid window (the padding would enter the statistics)");
    SCV_TRY(ensure((void**)&e->d_tile_stats, &e->tile_stats_cap, (size_t)n * job.C * 2 * sizeof(float)));
    TileStatsParams sp{};
    sp.src = ep.src;
    sp.dtype = job.dtype;
    sp.W = job.src_W;
    sp.C = job.C;
    sp.src_row0 = job.src_row0;
    sp.origins = ep.origins;
    sp.side = job.side;
    sp.mode = ep.norm_mode;
    sp.eps = ep.div[0];
    sp.stats = e->d_tile_stats;
    sp.ngroups = ep.ngroups;
    for (int c = 0; c < SCV_MAX_BANDS; ++c) sp.group_end[c] = ep.group_end[c];
    CUDA_TRY(launch_tile_stats(sp, n, s));
    e->n_launches++;
    ep.tile_stats = e->d_tile_stats;
  }
  if (extract_smem_bytes(ep) > 160 * 1024) return fail(SCV_ERR_INVALID, "tile row of %d bytes too large for the extract kernel", row_bytes);
  CUDA_TRY(launch_extract(ep, s));
  e->n_launches++;
  bev.e1 = new_event(e, s);

  int off = 0;
  for (int nb : sizes) {
    Plan* pl = nullptr;
    SCV_TRY(get_plan(e, nb, job.side, job.side, &pl));
    SCV_TRY(run_layers(e, pl, off, job.side, s, &bev));
    off += nb;
  }
  bev.e2 = new_event(e, s);

  if (job.d_dst_origins) {
    StitchParams sp{};
    sp.logits = e->d_logits_all;
    sp.side = job.side;
    sp.ncls = a.cfg.nclasses;
    sp.head = a.cfg.head;
    sp.threshold = a.cfg.threshold;
    sp.out_channel = job.out_channel;
    sp.crop_y = job.crop_y;
    sp.crop_x = job.crop_x;
    sp.kernel_h = job.kernel_h;
    sp.kernel_w = job.kernel_w;
    sp.dst_origins = job.d_dst_origins + t0;
    sp.force_scalar = job.force_scalar;
    sp.dst_row0 = job.dst_row0;
    sp.out_W = job.out_W;
    sp.prob = job.d_prob;
    sp.prob_f64 = job.prob_f64;
    sp.accumulate = job.accumulate;
    sp.mask = job.d_mask;
    CUDA_TRY(launch_stitch(sp, n, s));
  } else {
    HeadTilesParams hp{};
    hp.logits = e->d_logits_all;
    hp.npix = (long long)n * job.side * job.side;
    hp.ncls = a.cfg.nclasses;
    hp.head = a.cfg.head;
    hp.threshold = a.cfg.threshold;
    const size_t tile_px = (size_t)job.side * job.side;
    hp.probs = job.d_tile_probs ? job.d_tile_probs + (size_t)t0 * tile_px * a.cfg.nclasses : nullptr;
    hp.classes = job.d_tile_classes ? job.d_tile_classes + (size_t)t0 * tile_px : nullptr;
    CUDA_TRY(launch_head_tiles(hp, s));
  }
  e->n_launches++;
  bev.e3 = new_event(e, s);
  e->batch_ev.push_back(bev);
  return SCV_OK;
}

static void balanced_batches(int n, int maxb, std::vector<int>* sizes) {
  sizes->clear();
  if (n <= 0) return;
  const int nb = (n + maxb - 1) / maxb;
  const int base = n / nb, extra = n % nb;
  for (int i = 0; i < nb; ++i) sizes->push_back(base + (i < extra ? 1 : 0));
}

// Groups the balanced device batches of `n` tiles into super-batches of roughly `target` tiles.
static void super_batches(int n, int maxb, int target, std::vector<std::vector<int>>* out) {
  std::vector<int> sizes;
  balanced_batches(n, maxb, &sizes);
  out->clear();
  if (sizes.empty()) return;
  const int per = std::max(1, (target + sizes[0] - 1) / sizes[0]);          // device batches per super-batch
  const int nsuper = ((int)sizes.size() + per - 1) / per;
  const int base = (int)sizes.size() / nsuper, extra = (int)sizes.size() % nsuper;
  size_t k = 0;
  for (int i = 0; i < nsuper; ++i) {
    const int cnt = base + (i < extra ? 1 : 0);
    out->emplace_back(sizes.begin() + k, sizes.begin() + k + cnt);
    k += cnt;
  }
}

static int check_tiling(const scv_engine* e, const scv_tiling* t) {
  if (!t || t->kernel <= 0 || t->buff < 0 || (t->buff & 1))
    return fail(SCV_ERR_INVALID, "tiling: kernel must be > 0 and buff a non-negative even number");
  const int side = t->kernel + t->buff;
  if (side % (1 << e->arch.cfg.nlevels))
    return fail(SCV_ERR_INVALID, "tile side kernel+buff=%d must be a multiple of %d", side, 1 << e->arch.cfg.nlevels);
  return SCV_OK;
}

// Chip list of a mosaic call: generate_chip_indices (utils/prediction_tools.py:87-109,
// y in range(buff/2, H-(buff+kernel), kernel)) restricted to the chip range [tb, te) of its row-major order.
struct MosaicGeom {
  std::vector<int> ys, xs;
  int half, side, K, ncols;
  int tb, te;          // chip range
  int r_first, r_last; // tile rows touched (inclusive)
  int n() const { return te - tb; }
  int c0(int r) const { return r == r_first ? tb % ncols : 0; }                 // first chip column of tile row r
  int c1(int r) const { return r == r_last ? (te - 1) % ncols + 1 : ncols; }    // one past the last
};

static int mosaic_geom(const scv_engine* e, int H, int W, const scv_tiling* t, const scv_mosaic_opts* o, MosaicGeom* g) {
  SCV_TRY(check_tiling(e, t));
  g->side = t->buff + t->kernel;
  g->half = t->buff / 2;
  g->K = t->kernel;
  g->ys.clear();
  g->xs.clear();
  for (int y = g->half; y < H - g->side; y += t->kernel) g->ys.push_back(y);
  for (int x = g->half; x < W - g->side; x += t->kernel) g->xs.push_back(x);
  g->ncols = (int)g->xs.size();
  const int total = (int)g->ys.size() * g->ncols;
  g->tb = o ? std::max(0, o->tile_begin) : 0;
  g->te = (o && o->tile_end > 0) ? std::min(o->tile_end, total) : total;
  if (g->tb >= g->te) {
    g->tb = g->te = 0;  // empty chip list: nothing predicted
    g->r_first = 0, g->r_last = -1;
    return SCV_OK;
  }
  g->r_first = g->tb / g->ncols;
  g->r_last = (g->te - 1) / g->ncols;
  return SCV_OK;
}

static int check_opts(const scv_engine* e, const scv_mosaic_opts* o) {
  if (!o) return SCV_OK;
  if (o->out_channel < 0 || o->out_channel >= e->arch.cfg.nclasses) return fail(SCV_ERR_INVALID, "out_channel %d out of range", o->out_channel);
  if (o->out_dtype != 0 && o->out_dtype != SCV_F32 && o->out_dtype != SCV_F64)
    return fail(SCV_ERR_INVALID, "out_dtype must be SCV_F32 or SCV_F64");
  return SCV_OK;
}

// Uploads the chip origins of `g` ([n source origins][n destination origins]) unless the same list is already
// resident (repeated calls on one geometry: no per-call upload, no synchronisation).
static int upload_origins(scv_engine* e, const MosaicGeom& g, int H, int W, cudaStream_t s, int* force_scalar) {
  const int n = g.n();
  long long key[8] = {H, W, g.K, g.half, g.tb, g.te, 1, 0};
  int fs = 0;
  for (int x : g.xs) fs |= (x & 3) ? 3 : ((x & 7) ? 2 : 0);
  *force_scalar = fs;
  if (e->origins_valid && memcmp(key, e->origins_key, sizeof key) == 0) return SCV_OK;
  e->h_origins.resize(2 * (size_t)n);
  for (int k = 0; k < n; ++k) {
    const int t = g.tb + k, r = t / g.ncols, c = t % g.ncols;
    e->h_origins[k] = make_int2(g.xs[c] - g.half, g.ys[r] - g.half);
    e->h_origins[n + k] = make_int2(g.xs[c], g.ys[r]);
  }
  e->origins_valid = false;
  SCV_TRY(ensure((void**)&e->d_origins, &e->origins_cap, e->h_origins.size() * sizeof(int2)));
  // pageable source: the runtime stages it before returning, so h_origins may be rewritten by the next call
  CUDA_TRY(cudaMemcpyAsync(e->d_origins, e->h_origins.data(), e->h_origins.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
  if (s == e->stream) {  // a caller-supplied stream gives no ordering against later calls on other streams
    memcpy(e->origins_key, key, sizeof key);
    e->origins_valid = true;
  }
  return SCV_OK;
}

static void fill_mosaic_job(TileJob* job, const MosaicGeom& g, int dtype, int W, int C, const scv_norm* norm,
                            const scv_mosaic_opts* o, int force_scalar, const int2* d_origins) {
  job->dtype = dtype;
  job->src_W = W;
  job->C = C;
  job->side = g.side;
  job->d_src_origins = d_origins;
  job->norm = norm;
  job->d_dst_origins = d_origins + g.n();
  job->kernel_h = job->kernel_w = g.K;
  job->crop_y = job->crop_x = g.half;
  job->out_W = W;
  job->out_channel = o ? o->out_channel : 0;
  job->force_scalar = force_scalar;
  job->prob_f64 = (o && o->out_dtype == SCV_F64) ? 1 : 0;
  job->accumulate = (o && o->accumulate) ? 1 : 0;
  if (o) memcpy(job->valid, o->valid, sizeof job->valid);
  job->order_kernel = g.K <= g.side ? g.K : 0;
  job->tiles_per_row = g.ncols;
  job->order_skip0 = g.tb % std::max(1, g.ncols);
}

// RAII: the API must not change the caller's (e.g. torch's) current device behind its back
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// =================================================================== C ABI
extern "C" {

const char* scv_version(void) { return "scv-b200 0.3 (sm_100a, tcgen05/TMEM/TMA)"; }
const char* scv_last_error(void) { return g_err.c_str(); }

int scv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int scv_num_weights(const scv_config* cfg) {
  if (validate_config(cfg) != SCV_OK) return SCV_ERR_INVALID;
  std::vector<WeightSpec> s;
  build_specs(cfg, &s);
  return (int)s.size();
}

int scv_weight_shape(const scv_config* cfg, int index, int* ndim, int64_t shape[4], char* name, int name_len) {
  SCV_TRY(validate_config(cfg));
  std::vector<WeightSpec> s;
  build_specs(cfg, &s);
  if (index < 0 || index >= (int)s.size()) return fail(SCV_ERR_INVALID, "weight index %d out of range", index);
  if (ndim) *ndim = s[index].ndim;
  if (shape)
    for (int i = 0; i < 4; ++i) shape[i] = s[index].shape[i];
  if (name && name_len > 0) snprintf(name, name_len, "%s", s[index].name.c_str());
  return SCV_OK;
}

int scv_engine_create(const scv_config* cfg, scv_engine** out) {
  if (!out) return fail(SCV_ERR_INVALID, "out is NULL");
  *out = nullptr;
  SCV_TRY(validate_config(cfg));
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SCV_ERR_CUDA, "no CUDA device available: this engine has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SCV_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(SCV_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
  auto e = std::make_unique<scv_engine>();
  SCV_TRY(build_arch(cfg, &e->arch));
  if (e->arch.cfg.max_batch <= 0) e->arch.cfg.max_batch = 64;
  e->device = cfg->device;
  e->opt_host_register = env_int("SCV_HOST_REGISTER", 0);
  e->opt_host_first_row = env_int("SCV_HOST_FIRST_ROW", 1);
  CUDA_TRY(conv_init_attributes());
  CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&e->h2d, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&e->d2h, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(&e->d_err, sizeof(int)));
  CUDA_TRY(cudaMemset(e->d_err, 0, sizeof(int)));
  for (auto& sl : e->slots) {
    CUDA_TRY(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
  }
  *out = e.release();
  return SCV_OK;
}

static void slot_release_host(scv_engine::Slot& sl) {
  for (void* p : sl.registered) cudaHostUnregister(p);
  sl.registered.clear();
}

void scv_engine_destroy(scv_engine* e) {
  if (!e) return;
  DeviceGuard guard;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  reset_timing(e);
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  for (auto& pl : e->plans) cudaFree(pl->arena);
  for (auto& l : e->arch.layers) {
    cudaFree(l.d_w);
    cudaFree(l.d_w16);
    cudaFree(l.d_bias);
    cudaFree(l.d_skip_s);
    cudaFree(l.d_skip_t);
  }
  for (auto& sl : e->slots) {
    slot_release_host(sl);
    cudaFree(sl.d_scene);
    cudaFree(sl.d_prob);
    cudaFree(sl.d_mask);
    for (auto ev : sl.up_ev) cudaEventDestroy(ev);
    for (auto ev : sl.row_ev) cudaEventDestroy(ev);
    if (sl.compute_done) cudaEventDestroy(sl.compute_done);
    if (sl.d2h_done) cudaEventDestroy(sl.d2h_done);
  }
  cudaFree(e->d_head_w);
  cudaFree(e->d_head_b);
  cudaFree(e->d_err);
  cudaFree(e->d_prob);
  cudaFree(e->d_mask);
  cudaFree(e->d_origins);
  cudaFree(e->d_tile_stats);
  cudaFree(e->d_stage);
  cudaFree(e->d_tile_probs);
  cudaFree(e->d_tile_classes);
  cudaFree(e->d_x0_all);
  cudaFree(e->d_logits_all);
  cudaStreamDestroy(e->stream);
  cudaStreamDestroy(e->h2d);
  cudaStreamDestroy(e->d2h);
  delete e;
}

int scv_engine_set_weights(scv_engine* e, const scv_tensor* tensors, int n) {
  if (!e || !tensors) return fail(SCV_ERR_INVALID, "NULL argument");
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(e->device));
  const auto& specs = e->arch.specs;
  if (n != (int)specs.size()) return fail(SCV_ERR_INVALID, "expected %d weight arrays, got %d", (int)specs.size(), n);
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].data) return fail(SCV_ERR_INVALID, "weight %d (%s) has NULL data", i, specs[i].name.c_str());
    if (tensors[i].ndim != specs[i].ndim) return fail(SCV_ERR_INVALID, "weight %d (%s): ndim %d, expected %d", i, specs[i].name.c_str(), tensors[i].ndim, specs[i].ndim);
    for (int d = 0; d < specs[i].ndim; ++d)
      if (tensors[i].shape[d] != specs[i].shape[d])
        return fail(SCV_ERR_INVALID, "weight %d (%s): dim %d is %lld, expected %lld", i, specs[i].name.c_str(), d, (long long)tensors[i].shape[d], (long long)specs[i].shape[d]);
  }
  CUDA_TRY(cudaDeviceSynchronize());
  SCV_TRY(fold_and_upload(e, tensors));
  // weight pointers / tensor maps are baked into the plans
  for (auto& pl : e->plans) cudaFree(pl->arena);
  e->plans.clear();
  e->weights_set = true;
  return SCV_OK;
}

int scv_set_option(scv_engine* e, const char* key, int value) {
  if (!e || !key) return fail(SCV_ERR_INVALID, "NULL argument");
  DeviceGuard guard;
  cudaSetDevice(e->device);
  const std::string k(key);
  if (k == "profile_layers") e->opt_profile_layers = value;
  else if (k == "stages") {
    e->opt_stages = value;
    cudaStreamSynchronize(e->stream);
    for (auto& pl : e->plans) cudaFree(pl->arena);
    e->plans.clear();
  } else if (k == "watchdog_ms") e->opt_watchdog_ms = value;
  else if (k == "max_batch") e->arch.cfg.max_batch = std::max(1, value);
  else if (k == "super_tiles") e->opt_super_tiles = std::max(1, value);
  else if (k == "host_super_tiles") e->opt_host_super_tiles = std::max(0, value);
  else if (k == "host_register") e->opt_host_register = value;
  else if (k == "host_first_row") e->opt_host_first_row = value;
  else return fail(SCV_ERR_INVALID, "unknown option '%s'", key);
  return SCV_OK;
}

int scv_get_times(scv_engine* e, scv_times* out) {
  if (!e || !out) return fail(SCV_ERR_INVALID, "NULL argument");
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(e->device));
  SCV_TRY(finalize_times(e));
  *out = e->times;
  return SCV_OK;
}

int scv_check(scv_engine* e) {
  if (!e) return fail(SCV_ERR_INVALID, "engine is NULL");
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaDeviceSynchronize());
  return check_device_err(e);
}

void* scv_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    g_err = "cudaHostAlloc failed";
    return nullptr;
  }
  return p;
}
void scv_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

static int precheck(scv_engine* e, int dtype, int C) {
  if (!e) return fail(SCV_ERR_INVALID, "engine is NULL");
  if (!e->weights_set) return fail(SCV_ERR_STATE, "predict called before scv_engine_set_weights");
  if (dtype_bytes(dtype) == 0) return fail(SCV_ERR_INVALID, "unknown dtype %d", dtype);
  if (C != input_channels(&e->arch.cfg))
    return fail(SCV_ERR_INVALID, "input has %d bands, model expects %d%s", C, input_channels(&e->arch.cfg),
                e->arch.cfg.arch == SCV_ARCH_SIAMESE ? " (the two images stacked along the channel axis)" : "");
  CUDA_TRY(cudaSetDevice(e->device));
  return SCV_OK;
}

int scv_predict_mosaic_device_ex(scv_engine* e, const void* d_hwc, int dtype, int H, int W, int C, int src_row0,
                                 const scv_tiling* tiling, const scv_norm* norm, const scv_mosaic_opts* opts,
                                 void* d_prob, uint8_t* d_mask, int dst_row0, void* stream) {
  DeviceGuard guard;
  SCV_TRY(precheck(e, dtype, C));
  SCV_TRY(check_opts(e, opts));
  if (!d_hwc) return fail(SCV_ERR_INVALID, "d_hwc is NULL");
  MosaicGeom g;
  SCV_TRY(mosaic_geom(e, H, W, tiling, opts, &g));
  reset_timing(e);
  e->times_pending = true;
  if (g.n() == 0) return SCV_OK;  // empty chip list: nothing predicted
  if (g.ys[g.r_first] - g.half < src_row0) return fail(SCV_ERR_INVALID, "src_row0=%d is below the first needed mosaic row %d", src_row0, g.ys[g.r_first] - g.half);
  if (g.ys[g.r_first] < dst_row0) return fail(SCV_ERR_INVALID, "dst_row0=%d is below the first written row %d", dst_row0, g.ys[g.r_first]);
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  int force_scalar = 0;
  SCV_TRY(upload_origins(e, g, H, W, s, &force_scalar));

  TileJob job{};
  fill_mosaic_job(&job, g, dtype, W, C, norm, opts, force_scalar, e->d_origins);
  job.d_src = d_hwc;
  const int last_row = g.ys[g.r_last] - g.half + g.side;  // exclusive
  job.src_bytes = (size_t)(last_row - src_row0) * W * C * dtype_bytes(dtype);
  job.src_row0 = src_row0;
  job.dst_row0 = dst_row0;
  job.d_prob = d_prob;
  job.d_mask = d_mask;
  std::vector<std::vector<int>> supers;
  super_batches(g.n(), e->arch.cfg.max_batch, e->opt_super_tiles, &supers);
  int t0 = 0;
  for (auto& sizes : supers) {
    SCV_TRY(run_super(e, job, t0, sizes, s));
    for (int nb : sizes) t0 += nb;
  }
  if (!stream) {
    CUDA_TRY(cudaStreamSynchronize(s));
    SCV_TRY(check_device_err(e));
  }
  return SCV_OK;
}

static void rows_to_opts(const scv_engine* e, int H, int W, const scv_tiling* t, int row_begin, int row_end, int out_channel,
                         scv_mosaic_opts* o) {
  memset(o, 0, sizeof *o);
  o->out_channel = out_channel;
  if (!t || t->kernel <= 0) return;
  const int side = t->buff + t->kernel, half = t->buff / 2;
  int nrows = 0, ncols = 0;
  for (int y = half; y < H - side; y += t->kernel) ++nrows;
  for (int x = half; x < W - side; x += t->kernel) ++ncols;
  if (row_end < 0 || row_end > nrows) row_end = nrows;
  row_begin = std::max(0, row_begin);
  o->tile_begin = row_begin * ncols;
  o->tile_end = row_end * ncols;
  if (o->tile_end <= o->tile_begin) o->tile_begin = o->tile_end = nrows * ncols + 1;  // empty range (tile_end <= 0 would mean "all")
  (void)e;
}

int scv_predict_mosaic_device(scv_engine* e, const void* d_hwc, int dtype, int H, int W, int C, int src_row0,
                              const scv_tiling* tiling, const scv_norm* norm, int tile_row_begin, int tile_row_end,
                              int out_channel, float* d_prob, uint8_t* d_mask, int dst_row0, void* stream) {
  scv_mosaic_opts o;
  rows_to_opts(e, H, W, tiling, tile_row_begin, tile_row_end, out_channel, &o);
  return scv_predict_mosaic_device_ex(e, d_hwc, dtype, H, W, C, src_row0, tiling, norm, &o, d_prob, d_mask, dst_row0, stream);
}

// Page-locks [p, p+bytes) for the duration of a scene unless it already is (cudaHostAlloc / registered memory):
// pageable buffers would turn every cudaMemcpyAsync below into a synchronous staged copy.
static void maybe_register(scv_engine* e, scv_engine::Slot& sl, const void* p, size_t bytes) {
  if (!e->opt_host_register || !p || bytes == 0) return;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  if (at.type != cudaMemoryTypeUnregistered) return;
  const uintptr_t lo = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(4095);
  const uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + bytes + 4095) & ~uintptr_t(4095);
  if (cudaHostRegister(reinterpret_cast<void*>(lo), hi - lo, cudaHostRegisterPortable) == cudaSuccess)
    sl.registered.push_back(reinterpret_cast<void*>(lo));
  else
    cudaGetLastError();  // not fatal: the copies fall back to staged transfers
}

static bool host_is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// Blocks until the scene that last used `sl` is complete in host memory, then releases its host registrations.
static int slot_wait(scv_engine* e, scv_engine::Slot& sl) {
  if (!sl.busy) return SCV_OK;
  cudaError_t err = cudaEventSynchronize(sl.d2h_done);
  cudaError_t err2 = cudaEventSynchronize(sl.compute_done);
  sl.busy = false;
  slot_release_host(sl);
  CUDA_TRY(err);
  CUDA_TRY(err2);
  return check_device_err(e);
}

static cudaEvent_t slot_event(std::vector<cudaEvent_t>& pool, size_t i) {
  while (pool.size() <= i) {
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    pool.push_back(ev);
  }
  return pool[i];
}

// Enqueues one scene on slot `sl` (H2D per tile row on the copy stream, K1 / network / K4 per super-batch on the
// compute stream, D2H of every completed tile row on the third stream) and returns without waiting.
static double host_now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int submit_host_mosaic(scv_engine* e, scv_engine::Slot& sl, const void* hwc, int dtype, int H, int W, int C,
                              const scv_tiling* tiling, const scv_norm* norm, const scv_mosaic_opts* opts,
                              void* out_prob, uint8_t* out_mask, bool may_register) {
  const bool trace = env_int("SCV_HOST_TRACE", 0) != 0;
  const double tr0 = trace ? host_now_ms() : 0.0;
  double tr1 = 0, tr2 = 0;
  MosaicGeom g;
  SCV_TRY(mosaic_geom(e, H, W, tiling, opts, &g));
  reset_timing(e);
  e->times_pending = true;
  if (g.n() == 0) return SCV_OK;
  const int es = dtype_bytes(dtype), K = g.K;
  const bool f64 = opts && opts->out_dtype == SCV_F64, acc = opts && opts->accumulate;
  const size_t osz = f64 ? 8 : 4;
  const size_t row_bytes = (size_t)W * C * es;
  const int src_row0 = g.ys[g.r_first] - g.half;
  const int src_row1 = g.ys[g.r_last] - g.half + g.side;
  const int dst_row0 = g.ys[g.r_first];
  const int dst_rows = (g.r_last - g.r_first + 1) * K;
  SCV_TRY(ensure(&sl.d_scene, &sl.scene_bytes, (size_t)(src_row1 - src_row0) * row_bytes));
  SCV_TRY(ensure(&sl.d_prob, &sl.prob_bytes, (size_t)dst_rows * W * osz));
  if (out_mask) SCV_TRY(ensure((void**)&sl.d_mask, &sl.mask_bytes, (size_t)dst_rows * W));
  if (may_register) {
    maybe_register(e, sl, (const uint8_t*)hwc + (size_t)src_row0 * row_bytes, (size_t)(src_row1 - src_row0) * row_bytes);
    maybe_register(e, sl, (uint8_t*)out_prob + (size_t)dst_row0 * W * osz, (size_t)dst_rows * W * osz);
    if (out_mask) maybe_register(e, sl, out_mask + (size_t)dst_row0 * W, (size_t)dst_rows * W);
  }
  int force_scalar = 0;
  SCV_TRY(upload_origins(e, g, H, W, e->stream, &force_scalar));

  // From here on work is in flight: every failure goes through `fail_rc` so the streams are drained before
  // the caller gets its buffers back.
  int rc = SCV_OK;
  cudaError_t ce = cudaSuccess;
#define HOST_TRY(expr)                                                                                        \
  do {                                                                                                        \
    if (rc == SCV_OK && (ce = (expr)) != cudaSuccess)                                                         \
      rc = fail(SCV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(ce), __FILE__, __LINE__);   \
  } while (0)
  auto col_x0 = [&](int c) { return g.xs[c] - g.half; };
  auto col_x1 = [&](int c) { return g.xs[c] - g.half + g.side; };

  // H2D: one chunk per tile row = the mosaic rows it adds to what the rows above already brought (the first
  // chunk also carries the top buffer); a chunk's rows that only this tile row reads are narrowed to its chip
  // columns, the rows it shares with the next tile row to the union of both.  With page-locked buffers all
  // chunks are enqueued up front on the copy stream, so PCIe runs back to back while compute starts on chunk 0;
  // with pageable buffers (every cudaMemcpyAsync is then a blocking staged copy) a chunk is uploaded just before
  // the first super-batch that needs it, and the download of super-batch k is issued after the kernels of k+1,
  // so the host blocks in copies while the GPU computes.
  const int ntr = g.r_last - g.r_first + 1;
  int uploaded = src_row0, rows_uploaded = 0;
  e->ev_host_begin = new_event(e, e->h2d);
  auto upload_tile_row = [&](int r) {
    const int need = g.ys[r] - g.half + g.side;  // exclusive
    const int share = r < g.r_last ? std::max(uploaded, g.ys[r + 1] - g.half) : need;  // rows >= share are also read by r+1
    auto copy_rows = [&](int y0, int y1, int xa, int xb) {
      if (y1 <= y0 || xb <= xa) return;
      const size_t off = (size_t)xa * C * es;
      HOST_TRY(cudaMemcpy2DAsync((uint8_t*)sl.d_scene + (size_t)(y0 - src_row0) * row_bytes + off, row_bytes,
                                 (const uint8_t*)hwc + (size_t)y0 * row_bytes + off, row_bytes, (size_t)(xb - xa) * C * es,
                                 y1 - y0, cudaMemcpyHostToDevice, e->h2d));
    };
    copy_rows(uploaded, share, col_x0(g.c0(r)), col_x1(g.c1(r) - 1));
    if (r < g.r_last)
      copy_rows(share, need, col_x0(std::min(g.c0(r), g.c0(r + 1))), col_x1(std::max(g.c1(r), g.c1(r + 1)) - 1));
    uploaded = need;
    if (acc) {  // the rows of the caller's raster this tile row accumulates into
      const size_t xo = (size_t)g.xs[g.c0(r)], wbytes = (size_t)(g.c1(r) - g.c0(r)) * K * osz;
      HOST_TRY(cudaMemcpy2DAsync((uint8_t*)sl.d_prob + ((size_t)(g.ys[r] - dst_row0) * W + xo) * osz, (size_t)W * osz,
                                 (const uint8_t*)out_prob + ((size_t)g.ys[r] * W + xo) * osz, (size_t)W * osz, wbytes, K,
                                 cudaMemcpyHostToDevice, e->h2d));
    }
    cudaEvent_t ev = slot_event(sl.up_ev, r - g.r_first);
    if (!ev) rc = fail(SCV_ERR_CUDA, "cudaEventCreate failed");
    else HOST_TRY(cudaEventRecord(ev, e->h2d));
  };
  if (trace) tr1 = host_now_ms();
  const bool pinned_in = host_is_pinned((const uint8_t*)hwc + (size_t)src_row0 * row_bytes);
  const bool pinned_out = host_is_pinned((uint8_t*)out_prob + (size_t)dst_row0 * W * osz);
  if (pinned_in && (!acc || pinned_out))
    for (; rows_uploaded < ntr && rc == SCV_OK; ++rows_uploaded) upload_tile_row(g.r_first + rows_uploaded);

  if (trace) tr2 = host_now_ms();
  TileJob job{};
  fill_mosaic_job(&job, g, dtype, W, C, norm, opts, force_scalar, e->d_origins);
  job.d_src = sl.d_scene;
  job.src_bytes = (size_t)(src_row1 - src_row0) * row_bytes;
  job.src_row0 = src_row0;
  job.dst_row0 = dst_row0;
  job.d_prob = sl.d_prob;
  job.d_mask = out_mask ? sl.d_mask : nullptr;

  // Super-batches: smaller than on the device-resident path so that compute starts after the first tile row(s)
  // of H2D and the D2H tail after the last K4 stays short.
  std::vector<std::vector<int>> supers;
  const int host_super = e->opt_host_super_tiles > 0 ? e->opt_host_super_tiles : e->arch.cfg.max_batch;
  {
    // the first super-batch is the (rest of the) first tile row only: the kernels start after ONE tile row of H2D
    // instead of the three a full device batch spans, and the exposed start of the pipeline shrinks accordingly
    const int first = e->opt_host_first_row ? std::min({g.n(), g.ncols - g.c0(g.r_first), e->arch.cfg.max_batch}) : 0;
    std::vector<std::vector<int>> rest;
    super_batches(g.n() - first, e->arch.cfg.max_batch, std::min(e->opt_super_tiles, host_super), &rest);
    if (first > 0 && g.n() - first >= e->arch.cfg.max_batch / 2) supers.push_back(std::vector<int>{first});
    else if (first > 0) super_batches(g.n(), e->arch.cfg.max_batch, std::min(e->opt_super_tiles, host_super), &rest);
    for (auto& r : rest) supers.push_back(r);
  }
  int t0 = 0, rows_downloaded = 0, rows_waited = 0, n_row_ev = 0;
  // D2H of tile rows [rows_downloaded, upto) (cores only), ordered after `ev` on the compute stream
  auto download_rows = [&](int upto, cudaEvent_t ev) {
    if (upto <= rows_downloaded) return;
    HOST_TRY(cudaStreamWaitEvent(e->d2h, ev, 0));
    for (int rr = rows_downloaded; rr < upto && rc == SCV_OK; ++rr) {
      const int r = g.r_first + rr;
      const size_t xo = (size_t)g.xs[g.c0(r)], wpx = (size_t)(g.c1(r) - g.c0(r)) * K;
      const size_t doff = (size_t)(g.ys[r] - dst_row0) * W + xo, hoff = (size_t)g.ys[r] * W + xo;
      HOST_TRY(cudaMemcpy2DAsync((uint8_t*)out_prob + hoff * osz, (size_t)W * osz, (uint8_t*)sl.d_prob + doff * osz,
                                 (size_t)W * osz, wpx * osz, K, cudaMemcpyDeviceToHost, e->d2h));
      if (out_mask)
        HOST_TRY(cudaMemcpy2DAsync(out_mask + hoff, (size_t)W, sl.d_mask + doff, (size_t)W, wpx, K, cudaMemcpyDeviceToHost,
                                   e->d2h));
    }
    rows_downloaded = upto;
  };
  int pending_rows = 0;            // tile rows complete once `pending_ev` has fired, not yet downloaded
  cudaEvent_t pending_ev = nullptr;
  for (auto& sizes : supers) {
    if (rc != SCV_OK) break;
    int nsup = 0;
    for (int nb : sizes) nsup += nb;
    const int last_tile_row = (g.tb + t0 + nsup - 1) / g.ncols - g.r_first;  // relative tile row this super-batch reaches
    for (; rows_uploaded <= last_tile_row && rc == SCV_OK; ++rows_uploaded) upload_tile_row(g.r_first + rows_uploaded);
    for (; rows_waited <= last_tile_row && rc == SCV_OK; ++rows_waited)
      HOST_TRY(cudaStreamWaitEvent(e->stream, sl.up_ev[rows_waited], 0));
    if (rc != SCV_OK) break;
    if ((rc = run_super(e, job, t0, sizes, e->stream)) != SCV_OK) break;
    t0 += nsup;
    cudaEvent_t ev = slot_event(sl.row_ev, n_row_ev++);
    if (!ev) {
      rc = fail(SCV_ERR_CUDA, "cudaEventCreate failed");
      break;
    }
    HOST_TRY(cudaEventRecord(ev, e->stream));
    // the previous super-batch's rows leave now: its kernels are done or running, this one's are queued behind
    if (pending_ev) download_rows(pending_rows, pending_ev);
    pending_rows = (g.tb + t0 == g.te) ? ntr : (g.tb + t0) / g.ncols - g.r_first;
    pending_ev = ev;
  }
  if (rc == SCV_OK && pending_ev) download_rows(pending_rows, pending_ev);
  e->ev_host_end = new_event(e, e->d2h);
#undef HOST_TRY
  if (trace)
    fprintf(stderr, "[scv host trace dev %d] setup %.3f ms, H2D enqueue %.3f ms, kernels + D2H enqueue %.3f ms\n", e->device,
            tr1 - tr0, tr2 - tr1, host_now_ms() - tr2);
  sl.busy = true;
  cudaEventRecord(sl.compute_done, e->stream);
  cudaEventRecord(sl.d2h_done, e->d2h);
  if (rc != SCV_OK) {  // drain everything that was enqueued against the caller's buffers before reporting
    const std::string msg = g_err;
    cudaStreamSynchronize(e->h2d);
    cudaStreamSynchronize(e->stream);
    cudaStreamSynchronize(e->d2h);
    sl.busy = false;
    slot_release_host(sl);
    g_err = msg;
  }
  return rc;
}

static int host_args_check(scv_engine* e, const void* hwc, int dtype, int C, const scv_mosaic_opts* opts, void* out_prob) {
  SCV_TRY(precheck(e, dtype, C));
  SCV_TRY(check_opts(e, opts));
  if (!hwc || !out_prob) return fail(SCV_ERR_INVALID, "NULL buffer");
  return SCV_OK;
}

int scv_stream_submit(scv_engine* e, const void* hwc, int dtype, int H, int W, int C, const scv_tiling* tiling,
                      const scv_norm* norm, const scv_mosaic_opts* opts, void* out_prob, uint8_t* out_mask, int* ticket) {
  DeviceGuard guard;
  SCV_TRY(host_args_check(e, hwc, dtype, C, opts, out_prob));
  const int t = e->next_ticket;
  scv_engine::Slot& sl = e->slots[t & 1];
  SCV_TRY(slot_wait(e, sl));  // at most two scenes in flight
  SCV_TRY(submit_host_mosaic(e, sl, hwc, dtype, H, W, C, tiling, norm, opts, out_prob, out_mask, false));
  sl.ticket = t;
  e->next_ticket = t + 1;
  if (ticket) *ticket = t;
  return SCV_OK;
}

int scv_stream_wait(scv_engine* e, int ticket) {
  if (!e) return fail(SCV_ERR_INVALID, "engine is NULL");
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(e->device));
  int rc = SCV_OK;
  // scenes complete in submission order: waiting for ticket t means waiting for every busy slot up to t
  for (int k = 0; k < 2; ++k) {
    scv_engine::Slot& sl = e->slots[(e->next_ticket + k) & 1];  // older slot first
    if (sl.busy && (ticket < 0 || sl.ticket <= ticket)) {
      const int r = slot_wait(e, sl);
      if (rc == SCV_OK) rc = r;
    }
  }
  return rc;
}

int scv_predict_mosaic_ex(scv_engine* e, const void* hwc, int dtype, int H, int W, int C, const scv_tiling* tiling,
                          const scv_norm* norm, const scv_mosaic_opts* opts, void* out_prob, uint8_t* out_mask) {
  DeviceGuard guard;
  const bool trace = env_int("SCV_HOST_TRACE", 0) != 0;
  const double t0 = trace ? host_now_ms() : 0.0;
  SCV_TRY(host_args_check(e, hwc, dtype, C, opts, out_prob));
  for (auto& sl : e->slots) SCV_TRY(slot_wait(e, sl));  // a synchronous call does not overtake streamed scenes
  scv_engine::Slot& sl = e->slots[0];
  const double t1 = trace ? host_now_ms() : 0.0;
  SCV_TRY(submit_host_mosaic(e, sl, hwc, dtype, H, W, C, tiling, norm, opts, out_prob, out_mask, true));
  sl.ticket = -1;
  const double t2 = trace ? host_now_ms() : 0.0;
  const int rc = slot_wait(e, sl);
  if (trace)
    fprintf(stderr, "[scv host trace dev %d] call: entry %.3f ms, submit %.3f ms, wait %.3f ms\n", e->device, t1 - t0, t2 - t1,
            host_now_ms() - t2);
  return rc;
}

int scv_predict_mosaic(scv_engine* e, const void* hwc, int dtype, int H, int W, int C, const scv_tiling* tiling,
                       const scv_norm* norm, int tile_row_begin, int tile_row_end, int out_channel, float* out_prob,
                       uint8_t* out_mask) {
  scv_mosaic_opts o;
  rows_to_opts(e, H, W, tiling, tile_row_begin, tile_row_end, out_channel, &o);
  return scv_predict_mosaic_ex(e, hwc, dtype, H, W, C, tiling, norm, &o, out_prob, out_mask);
}

// shared by predict_tiles / predict_patches: upload stacked patches batch by batch
static int run_stacked(scv_engine* e, const void* nhwc, int dtype, int N, int H, int W, int C, const scv_norm* norm,
                       TileJob job, float* h_probs, int32_t* h_classes) {
  if (H != W) return fail(SCV_ERR_INVALID, "patches must be square (got %dx%d)", H, W);
  const int es = dtype_bytes(dtype);
  const size_t tile_bytes = (size_t)H * W * C * es;
  const size_t tile_px = (size_t)H * W;
  const int ncls = e->arch.cfg.nclasses;
  std::vector<int> sizes;
  balanced_batches(N, e->arch.cfg.max_batch, &sizes);
  const int maxb = sizes.empty() ? 0 : sizes[0];
  SCV_TRY(ensure(&e->d_stage, &e->stage_bytes, (size_t)maxb * tile_bytes));
  // origins layout: [maxb stacked-source origins][N destination origins] (the latter filled by the caller)
  e->origins_valid = false;
  e->h_origins.resize((size_t)maxb);
  for (int i = 0; i < maxb; ++i) e->h_origins[i] = make_int2(0, i * H);
  CUDA_TRY(cudaMemcpyAsync(e->d_origins, e->h_origins.data(), e->h_origins.size() * sizeof(int2), cudaMemcpyHostToDevice, e->stream));
  job.d_src = e->d_stage;
  job.dtype = dtype;
  job.src_W = W;
  job.C = C;
  job.src_row0 = 0;
  job.side = H;
  job.norm = norm;
  int t0 = 0;
  for (int nb : sizes) {
    CUDA_TRY(cudaMemcpyAsync(e->d_stage, (const uint8_t*)nhwc + (size_t)t0 * tile_bytes, (size_t)nb * tile_bytes,
                             cudaMemcpyHostToDevice, e->stream));
    job.src_bytes = (size_t)nb * tile_bytes;
    // run_super offsets origin arrays and per-tile outputs by t0: compensate for the per-batch staging
    TileJob bj = job;
    bj.d_src_origins = e->d_origins - t0;
    SCV_TRY(run_super(e, bj, t0, std::vector<int>{nb}, e->stream));
    if (h_probs)
      CUDA_TRY(cudaMemcpyAsync(h_probs + (size_t)t0 * tile_px * ncls, e->d_tile_probs + (size_t)t0 * tile_px * ncls,
                               (size_t)nb * tile_px * ncls * 4, cudaMemcpyDeviceToHost, e->stream));
    if (h_classes)
      CUDA_TRY(cudaMemcpyAsync(h_classes + (size_t)t0 * tile_px, e->d_tile_classes + (size_t)t0 * tile_px,
                               (size_t)nb * tile_px * 4, cudaMemcpyDeviceToHost, e->stream));
    t0 += nb;
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  SCV_TRY(check_device_err(e));
  return SCV_OK;
}

int scv_predict_tiles(scv_engine* e, const void* nhwc, int dtype, int N, int H, int W, int C, const scv_norm* norm,
                      float* probs, int32_t* classes) {
  DeviceGuard guard;
  SCV_TRY(precheck(e, dtype, C));
  if (!nhwc) return fail(SCV_ERR_INVALID, "nhwc is NULL");
  if (N < 0) return fail(SCV_ERR_INVALID, "N < 0");
  for (auto& sl : e->slots) SCV_TRY(slot_wait(e, sl));
  reset_timing(e);
  e->times_pending = true;
  if (N == 0) return SCV_OK;
  const size_t tile_px = (size_t)H * W;
  const int ncls = e->arch.cfg.nclasses;
  if (probs) SCV_TRY(ensure((void**)&e->d_tile_probs, &e->tile_probs_bytes, (size_t)N * tile_px * ncls * 4));
  if (classes) SCV_TRY(ensure((void**)&e->d_tile_classes, &e->tile_classes_bytes, (size_t)N * tile_px * 4));
  SCV_TRY(ensure((void**)&e->d_origins, &e->origins_cap, (size_t)(e->arch.cfg.max_batch + 1) * sizeof(int2)));
  TileJob job{};
  job.d_dst_origins = nullptr;
  job.d_tile_probs = probs ? e->d_tile_probs : nullptr;
  job.d_tile_classes = classes ? e->d_tile_classes : nullptr;
  return run_stacked(e, nhwc, dtype, N, H, W, C, norm, job, probs, classes);
}

int scv_predict_patches_ex(scv_engine* e, const void* nhwc, int dtype, int N, int H, int W, int C, const scv_crop* crop,
                           const scv_norm* norm, int cols, int out_channel, float* out_prob, uint8_t* out_mask) {
  DeviceGuard guard;
  SCV_TRY(precheck(e, dtype, C));
  if (!nhwc || !out_prob || !crop) return fail(SCV_ERR_INVALID, "NULL buffer");
  if (H % (1 << e->arch.cfg.nlevels) || W % (1 << e->arch.cfg.nlevels))
    return fail(SCV_ERR_INVALID, "patch %dx%d: H and W must be multiples of %d", H, W, 1 << e->arch.cfg.nlevels);
  if (crop->h <= 0 || crop->w <= 0 || crop->y0 < 0 || crop->x0 < 0 || crop->y0 + crop->h > H || crop->x0 + crop->w > W)
    return fail(SCV_ERR_INVALID, "crop window [%d:%d, %d:%d] leaves the %dx%d patch", crop->y0, crop->y0 + crop->h, crop->x0,
                crop->x0 + crop->w, H, W);
  if (cols <= 0 || N % cols) return fail(SCV_ERR_INVALID, "N=%d patches do not fill rows of %d", N, cols);
  if (out_channel < 0 || out_channel >= e->arch.cfg.nclasses) return fail(SCV_ERR_INVALID, "out_channel %d out of range", out_channel);
  for (auto& sl : e->slots) SCV_TRY(slot_wait(e, sl));
  reset_timing(e);
  e->times_pending = true;
  if (N == 0) return SCV_OK;
  const int rows = N / cols;
  const size_t out_px = (size_t)rows * crop->h * cols * crop->w;
  SCV_TRY(ensure((void**)&e->d_prob, &e->prob_bytes, out_px * 4));
  if (out_mask) SCV_TRY(ensure((void**)&e->d_mask, &e->mask_bytes, out_px));
  const int maxb = e->arch.cfg.max_batch;
  SCV_TRY(ensure((void**)&e->d_origins, &e->origins_cap, (size_t)(maxb + N + 1) * sizeof(int2)));
  std::vector<int2> dst((size_t)N);
  for (int i = 0; i < N; ++i) dst[i] = make_int2((i % cols) * crop->w, (i / cols) * crop->h);
  // on the compute stream, like every kernel that reads it (a pageable source is staged before the call returns)
  CUDA_TRY(cudaMemcpyAsync(e->d_origins + maxb, dst.data(), dst.size() * sizeof(int2), cudaMemcpyHostToDevice, e->stream));
  TileJob job{};
  job.d_dst_origins = e->d_origins + maxb;
  job.kernel_h = crop->h;
  job.kernel_w = crop->w;
  job.crop_y = crop->y0;
  job.crop_x = crop->x0;
  job.out_W = cols * crop->w;
  job.dst_row0 = 0;
  job.out_channel = out_channel;
  job.force_scalar = (crop->w & 3) ? 3 : ((crop->w & 7) ? 2 : 0);
  job.d_prob = e->d_prob;
  job.d_mask = out_mask ? e->d_mask : nullptr;
  SCV_TRY(run_stacked(e, nhwc, dtype, N, H, W, C, norm, job, nullptr, nullptr));
  CUDA_TRY(cudaMemcpy(out_prob, e->d_prob, out_px * 4, cudaMemcpyDeviceToHost));
  if (out_mask) CUDA_TRY(cudaMemcpy(out_mask, e->d_mask, out_px, cudaMemcpyDeviceToHost));
  return SCV_OK;
}

int scv_predict_patches(scv_engine* e, const void* nhwc, int dtype, int N, int H, int W, int C,
                        const scv_tiling* tiling, const scv_norm* norm, int cols, int out_channel, float* out_prob,
                        uint8_t* out_mask) {
  if (!e) return fail(SCV_ERR_INVALID, "engine is NULL");
  SCV_TRY(check_tiling(e, tiling));
  const int side = tiling->kernel + tiling->buff;
  if (H != side || W != side) return fail(SCV_ERR_INVALID, "patches are %dx%d but kernel+buff=%d", H, W, side);
  const scv_crop crop = {tiling->buff / 2, tiling->buff / 2, tiling->kernel, tiling->kernel};
  return scv_predict_patches_ex(e, nhwc, dtype, N, H, W, C, &crop, norm, cols, out_channel, out_prob, out_mask);
}

// ------------------------------------------------------------------ debug entries
static int debug_common_begin(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SCV_ERR_CUDA, "no CUDA device available: this engine has no CPU fallback");
  }
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(conv_init_attributes());
  return SCV_OK;
}

static int debug_conv(int device, int kind, const float* x, int N, int H, int W, int Cin, const float* kernel,
                      const float* bias, int Cout, int relu, float* y, float* pooled) {
  SCV_TRY(debug_common_begin(device));
  LayerDef l{};
  l.name = "debug";
  l.kind = kind;
  l.epi = kind == L_CONVT ? EPI_CONVT : (pooled ? EPI_POOL_SKIP : EPI_STORE);
  l.cin_real = Cin;
  l.cin_pad = pad_channels(Cin);
  l.cout = Cout;
  l.ntotal = kind == L_CONVT ? 4 * Cout : Cout;
  l.KC = kc_for(l.cin_pad);
  l.BN = bn_for(l.ntotal);
  if (const char* s = getenv("SCV_DEBUG_BN")) {
    const int bn = atoi(s);
    if (bn > 0 && l.ntotal % bn == 0) l.BN = bn;
  }
  if (l.BN == 0) return fail(SCV_ERR_INVALID, "Cout=%d unsupported (needs a multiple of 32)", Cout);
  const int ntaps = kind == L_CONV3 ? 9 : 1;
  const size_t ktotal = k_total(l.KC, ntaps, l.cin_pad);
  std::vector<float> w((size_t)l.ntotal * ktotal, 0.f), b(l.ntotal, 0.f);
  if (kind == L_CONV3) {
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < Cin; ++c)
        for (int o = 0; o < Cout; ++o) w[(size_t)o * ktotal + (size_t)tap * l.cin_pad + c] = kernel[((size_t)tap * Cin + c) * Cout + o];
    for (int o = 0; o < Cout; ++o) b[o] = bias[o];
  } else {
    for (int ab = 0; ab < 4; ++ab)
      for (int o = 0; o < Cout; ++o) {
        for (int c = 0; c < Cin; ++c) w[((size_t)ab * Cout + o) * ktotal + c] = kernel[((size_t)ab * Cout + o) * Cin + c];
        b[(size_t)ab * Cout + o] = bias[o];
      }
  }
  std::vector<float> ones(l.ntotal, 1.f), zeros(l.ntotal, 0.f);
  int rc = upload_layer(l, w, b, pooled ? &ones : nullptr, pooled ? &zeros : nullptr);
  // input: pad channels, round to bf16
  const size_t npix = (size_t)N * H * W;
  std::vector<__nv_bfloat16> xin(npix * l.cin_pad, __float2bfloat16_rn(0.f));
  for (size_t p = 0; p < npix; ++p)
    for (int c = 0; c < Cin; ++c) xin[p * l.cin_pad + c] = __float2bfloat16_rn(x[p * Cin + c]);
  const int oh = kind == L_CONVT ? 2 * H : H, ow = kind == L_CONVT ? 2 * W : W;
  const size_t nout = (size_t)N * oh * ow * Cout;
  const size_t npool = pooled ? (size_t)N * (H / 2) * (W / 2) * Cout : 0;
  __nv_bfloat16 *d_x = nullptr, *d_y = nullptr, *d_p = nullptr;
  float* d_f = nullptr;
  int* d_err = nullptr;
  std::vector<__nv_bfloat16> hy(nout), hp(npool);
  ConvLaunch Ln;
  auto cleanup = [&]() {
    cudaFree(d_x);
    cudaFree(d_y);
    cudaFree(d_p);
    cudaFree(d_f);
    cudaFree(d_err);
    cudaFree(l.d_w);
    cudaFree(l.d_w16);
    cudaFree(l.d_bias);
    cudaFree(l.d_skip_s);
    cudaFree(l.d_skip_t);
  };
#define DBG_TRY(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      cleanup();                                                                              \
      return fail(SCV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                         \
  } while (0)
  if (rc != SCV_OK) {
    cleanup();
    return rc;
  }
  DBG_TRY(cudaMalloc(&d_x, xin.size() * 2));
  DBG_TRY(cudaMalloc(&d_y, nout * 2));
  DBG_TRY(cudaMemset(d_y, 0xff, nout * 2));  // NaN pattern: unwritten outputs are visible
  if (pooled) {
    DBG_TRY(cudaMalloc(&d_p, npool * 2));
    DBG_TRY(cudaMemset(d_p, 0xff, npool * 2));
  }
  DBG_TRY(cudaMalloc(&d_err, sizeof(int)));
  DBG_TRY(cudaMemset(d_err, 0, sizeof(int)));
  DBG_TRY(cudaMemcpy(d_x, xin.data(), xin.size() * 2, cudaMemcpyHostToDevice));
  rc = fill_launch(nullptr, l, d_x, l.cin_pad, N, H, W, &Ln);
  if (rc != SCV_OK) {
    cleanup();
    return rc;
  }
  if (const char* s = getenv("SCV_DEBUG_STAGES"))
    if (!Ln.slab) {
      Ln.p.nstage = std::max(1, atoi(s));
      Ln.smem = conv_smem_bytes(l.KC, l.BN, Ln.p.nstage, l.epi, 1);
    }
  if (const char* s = getenv("SCV_DEBUG_GRID"))  // persistent kernels: fewer CTAs -> longer streams per CTA
    if (Ln.slab && atoi(s) > 0) {
      Ln.grid = std::max(Ln.p.n_tiles_n, std::min(Ln.grid, atoi(s)) / Ln.p.n_tiles_n * Ln.p.n_tiles_n);
      if (Ln.slab == 6) Ln.grid = std::max(2, Ln.grid & ~1);
    }
  Ln.p.relu = relu;
  Ln.p.out = d_y;
  Ln.p.out_pitch = Cout;
  Ln.p.out_choff = 0;
  Ln.p.pool_out = d_p;
  Ln.p.pool_pitch = Cout;
  Ln.p.err = d_err;
  rc = finish_slab_maps(&Ln, l);
  if (rc != SCV_OK) {
    cleanup();
    return rc;
  }
  DBG_TRY(conv_launch(Ln, 0));
  DBG_TRY(cudaDeviceSynchronize());
  int herr = 0;
  DBG_TRY(cudaMemcpy(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  DBG_TRY(cudaMemcpy(hy.data(), d_y, nout * 2, cudaMemcpyDeviceToHost));
  if (pooled) DBG_TRY(cudaMemcpy(hp.data(), d_p, npool * 2, cudaMemcpyDeviceToHost));
  cleanup();
#undef DBG_TRY
  for (size_t i = 0; i < nout; ++i) y[i] = __bfloat162float(hy[i]);
  for (size_t i = 0; i < npool; ++i) pooled[i] = __bfloat162float(hp[i]);
  if (herr) return fail(SCV_ERR_KERNEL, "device watchdog tripped in debug conv (KC=%d BN=%d slab=%d)", l.KC, Ln.BN, Ln.slab);
  return SCV_OK;
}

int scv_debug_conv3x3(int device, const float* x, int N, int H, int W, int Cin, const float* kernel, const float* bias,
                      int Cout, int relu, float* y, float* pooled) {
  if (!x || !kernel || !bias || !y) return fail(SCV_ERR_INVALID, "NULL argument");
  if (pooled && ((H | W) & 1)) return fail(SCV_ERR_INVALID, "pooled output needs even H and W");
  return debug_conv(device, L_CONV3, x, N, H, W, Cin, kernel, bias, Cout, relu, y, pooled);
}

int scv_debug_convT2x2(int device, const float* x, int N, int H, int W, int Cin, const float* kernel,
                       const float* bias, int Cout, int relu, float* y) {
  if (!x || !kernel || !bias || !y) return fail(SCV_ERR_INVALID, "NULL argument");
  return debug_conv(device, L_CONVT, x, N, H, W, Cin, kernel, bias, Cout, relu, y, nullptr);
}

int scv_debug_extract(int device, const void* hwc, int dtype, int H, int W, int C, const scv_tiling* tiling,
                      const scv_norm* norm, const int32_t* indices_yx, int n_tiles, float* tiles_out, int* cpad_out) {
  if (!hwc || !tiling || !indices_yx || !tiles_out) return fail(SCV_ERR_INVALID, "NULL argument");
  if (dtype_bytes(dtype) == 0) return fail(SCV_ERR_INVALID, "unknown dtype %d", dtype);
  if (C < 1 || C > SCV_MAX_BANDS) return fail(SCV_ERR_INVALID, "C out of range");
  SCV_TRY(debug_common_begin(device));
  const int side = tiling->kernel + tiling->buff, half = tiling->buff / 2;
  const int cpad = pad_channels(C);
  if (cpad_out) *cpad_out = cpad;
  std::vector<int2> org(n_tiles);
  for (int i = 0; i < n_tiles; ++i) {
    org[i] = make_int2(indices_yx[2 * i + 1] - half, indices_yx[2 * i] - half);
    if (org[i].x < 0 || org[i].y < 0 || org[i].x + side > W || org[i].y + side > H)
      return fail(SCV_ERR_INVALID, "chip %d at (%d,%d) leaves the %dx%d mosaic", i, indices_yx[2 * i], indices_yx[2 * i + 1], H, W);
  }
  const size_t src_bytes = (size_t)H * W * C * dtype_bytes(dtype);
  const size_t nout = (size_t)n_tiles * side * side * cpad;
  void* d_src = nullptr;
  int2* d_org = nullptr;
  __nv_bfloat16* d_out = nullptr;
  float* d_stats = nullptr;
  std::vector<__nv_bfloat16> h(nout);
  auto cleanup = [&]() {
    cudaFree(d_src);
    cudaFree(d_org);
    cudaFree(d_out);
    cudaFree(d_stats);
  };
#define DBG_TRY(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      cleanup();                                                                              \
      return fail(SCV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                         \
  } while (0)
  DBG_TRY(cudaMalloc(&d_src, src_bytes));
  DBG_TRY(cudaMalloc(&d_org, org.size() * sizeof(int2)));
  DBG_TRY(cudaMalloc(&d_out, nout * 2));
  DBG_TRY(cudaMemcpy(d_src, hwc, src_bytes, cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemcpy(d_org, org.data(), org.size() * sizeof(int2), cudaMemcpyHostToDevice));
  ExtractParams ep{};
  ep.src = (const uint8_t*)d_src;
  ep.src_bytes = src_bytes;
  ep.dtype = dtype;
  ep.W = W;
  ep.C = C;
  ep.src_row0 = 0;
  ep.origins = d_org;
  ep.n_tiles = n_tiles;
  ep.side = side;
  ep.cpad = cpad;
  int rc = norm_to_params(norm, C, &ep);
  if (rc != SCV_OK) {
    cleanup();
    return rc;
  }
  const int row_bytes = side * C * dtype_bytes(dtype);
  ep.rows_per_block = std::max(1, std::min(8, (44 * 1024) / (row_bytes + 32)));
  ep.out = d_out;
  if (tile_stats_mode(ep.norm_mode)) {
    DBG_TRY(cudaMalloc(&d_stats, (size_t)n_tiles * C * 2 * sizeof(float)));
    TileStatsParams sp{};
    sp.src = ep.src;
    sp.dtype = dtype;
    sp.W = W;
    sp.C = C;
    sp.src_row0 = 0;
    sp.origins = d_org;
    sp.side = side;
    sp.mode = ep.norm_mode;
    sp.eps = ep.div[0];
    sp.stats = d_stats;
    sp.ngroups = ep.ngroups;
    for (int c = 0; c < SCV_MAX_BANDS; ++c) sp.group_end[c] = ep.group_end[c];
    DBG_TRY(launch_tile_stats(sp, n_tiles, 0));
    ep.tile_stats = d_stats;
  }
  DBG_TRY(launch_extract(ep, 0));
  DBG_TRY(cudaDeviceSynchronize());
  DBG_TRY(cudaMemcpy(h.data(), d_out, nout * 2, cudaMemcpyDeviceToHost));
  cleanup();
#undef DBG_TRY
  for (size_t i = 0; i < nout; ++i) tiles_out[i] = __bfloat162float(h[i]);
  return SCV_OK;
}

}  // extern "C"
