"""Oracle (test infrastructure): the reference U-Net forward pass on the CPU.

Restates ``utils/model_tools.py:174-454`` (variant "B", the model
``get_unet_model`` builds *as written*) and
``notebooks/UNET_G4G_2019_solar.ipynb:1162-1213`` (variant "A", the classic
2x(conv-BN-ReLU) block of the notebooks / the intent of ``binary_unet``) with
torch-CPU fp32 (or fp64) contractions and *un-folded* BatchNorm.

Keras semantics restated (TensorFlow is not installable here -- parity
unpinned, see ``oracle/__init__.py``):

* ``Conv2D(F,(3,3),padding='same')``: NHWC cross-correlation, kernel HWIO
  (3,3,Cin,F), zero pad 1, + bias.
* ``BatchNormalization()``: axis -1, eps 1e-3, inference:
  ``gamma*(x-mean)/sqrt(var+eps)+beta``; ``get_weights()`` order
  [gamma, beta, moving_mean, moving_variance].
* ``MaxPooling2D((2,2), strides=(2,2))``: 'valid'.
* ``Conv2DTranspose(F,(2,2),strides=(2,2),padding='same')``: kernel
  (2,2,F,Cin); ``out[2i+a,2j+b,o] = sum_c in[i,j,c]*K[a,b,o,c] + bias[o]``.
* ``concatenate([skip, up], -1)``: skip first (``model_tools.py:307``).
* heads: ``Conv2D(ncls,(1,1),softmax)`` + int32 argmax (``:405-406``; ties ->
  lowest index) or ``Conv2D(1,(1,1),sigmoid)`` + int32(p > thr) strict
  (``:443-445``).

Weight list order = Keras ``model.get_weights()``: layers in creation order,
each layer trainable weights then non-trainable.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3
DEFAULT_FILTERS = (32, 64, 128, 256, 512)


# --------------------------------------------------------------------------- specs
def weight_specs(variant='A', nchannels=6, nclasses=1, filters=DEFAULT_FILTERS):
    """[(name, shape)] in ``get_weights()`` order.

    variant 'A': two conv-BN per encoder/centre block (solar.ipynb:1162-1170).
    variant 'B': one conv-BN per encoder/centre block -- ``conv_block.call``
    applies ``cba1`` twice to ``inputs`` and never builds ``cba2``
    (model_tools.py:238-239).
    """
    assert variant in ('A', 'B')
    nconv = 2 if variant == 'A' else 1
    specs = []

    def conv(name, cin, cout, k=3):
        specs.append((f'{name}/kernel', (k, k, cin, cout)))
        specs.append((f'{name}/bias', (cout,)))

    def bn(name, c):
        for w in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            specs.append((f'{name}/{w}', (c,)))

    cin = nchannels
    for i, f in enumerate(filters):
        for j in range(nconv):
            conv(f'encoder_{i}/conv{j}', cin, f)
            bn(f'encoder_{i}/bn{j}', f)
            cin = f
    fc = filters[-1] * 2
    for j in range(nconv):
        conv(f'center/conv{j}', cin, fc)
        bn(f'center/bn{j}', fc)
        cin = fc
    for i in range(len(filters) - 1, -1, -1):
        f = filters[i]
        specs.append((f'decoder_{i}/up/kernel', (2, 2, f, cin)))
        specs.append((f'decoder_{i}/up/bias', (f,)))
        bn(f'decoder_{i}/bn_cat', 2 * f)
        conv(f'decoder_{i}/conv0', 2 * f, f)
        bn(f'decoder_{i}/bn0', f)
        conv(f'decoder_{i}/conv1', f, f)
        bn(f'decoder_{i}/bn1', f)
        cin = f
    conv('head', cin, nclasses, k=1)
    return specs


def init_weights(specs, seed=0, randomize_bn=True, head_bias=None, head_gain=1.0, dtype=np.float32):
    """Deterministic Keras-style weights: glorot-uniform kernels, zero biases.

    ``randomize_bn`` draws gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1),
    var~U(0.5,1.5) and small random conv biases so BN folding is exercised;
    otherwise BN is the Keras default (1, 0, 0, 1).  ``head_gain`` scales the 1x1 head
    kernel: with a plain glorot head every probability of a random-init net sits within
    ~1e-2 of 0.5 (hard part 1), which makes thresholded-mask agreement a
    coin-flip test of rounding noise; a gain spreads the logits the way training does.
    """
    rng = np.random.default_rng(seed)
    out = []
    for name, shape in specs:
        leaf = name.rsplit('/', 1)[1]
        if leaf == 'kernel':
            if 'up/' in name:  # Conv2DTranspose (kh,kw,out,in): fan_in = kh*kw*in ... Keras computes
                kh, kw, co, ci = shape  # fans from (kh,kw,out,in) as receptive*shape[-2], receptive*shape[-1]
                fan_in, fan_out = kh * kw * co, kh * kw * ci
            else:
                kh, kw, ci, co = shape
                fan_in, fan_out = kh * kw * ci, kh * kw * co
            limit = np.sqrt(6.0 / (fan_in + fan_out))
            w = rng.uniform(-limit, limit, shape)
            if name.startswith('head'):
                w = w * head_gain
        elif leaf == 'bias':
            if name.startswith('head') and head_bias is not None:
                w = np.full(shape, head_bias, dtype=np.float64)
            elif randomize_bn:
                w = rng.normal(0.0, 0.05, shape)
            else:
                w = np.zeros(shape)
        elif leaf == 'gamma':
            w = rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)
        elif leaf == 'beta':
            w = rng.normal(0.0, 0.1, shape) if randomize_bn else np.zeros(shape)
        elif leaf == 'moving_mean':
            w = rng.normal(0.0, 0.1, shape) if randomize_bn else np.zeros(shape)
        elif leaf == 'moving_variance':
            w = rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)
        else:
            raise ValueError(name)
        out.append(np.ascontiguousarray(w, dtype=dtype))
    return out


# ------------------------------------------------------------------ torch forward
class _Cursor:
    def __init__(self, weights):
        self.w = weights
        self.i = 0

    def take(self, n):
        out = self.w[self.i:self.i + n]
        self.i += n
        return out


class _Prepared:
    """Weights converted once to torch tensors in the layout torch wants (what a
    keras.Model holds resident between predict calls)."""

    def __init__(self, weights, dt):
        import torch
        self.dt = dt
        self.t = []
        for w in weights:
            w = np.asarray(w)
            t = torch.from_numpy(np.ascontiguousarray(w)).to(dt)
            if w.ndim == 4:  # HWIO conv kernel / (kh,kw,out,in) transposed-conv kernel -> torch layout
                t = t.permute(3, 2, 0, 1).contiguous()
            self.t.append(t)


def prepare(weights, precision='fp32'):
    import torch
    return _Prepared(weights, torch.float32 if precision == 'fp32' else torch.float64)


def forward(x_nhwc, weights, variant='A', filters=DEFAULT_FILTERS, head='sigmoid',
            threshold=0.5, precision='fp32', return_logits=False, num_threads=None):
    """Forward pass.  ``x_nhwc`` (N,H,W,C) float; ``weights`` is the keras-ordered list (or the result
    of :func:`prepare`).  Returns ``(probs, classes)``: probs (N,H,W,k) float32, classes int32 --
    (N,H,W,1) for the sigmoid head (``model_tools.py:445``), (N,H,W) for softmax/argmax (``:406``).
    """
    import torch
    import torch.nn.functional as F
    if num_threads:
        torch.set_num_threads(num_threads)
    prep = weights if isinstance(weights, _Prepared) else prepare(weights, precision)
    dt = prep.dt
    nconv = 2 if variant == 'A' else 1
    cur = _Cursor(prep.t)

    def bn(y, g, be, mu, var):
        g, be, mu, var = (v.view(1, -1, 1, 1) for v in (g, be, mu, var))
        return g * (y - mu) / torch.sqrt(var + BN_EPS) + be

    def conv_bn_relu(x):
        k, b = cur.take(2)
        y = F.conv2d(x, k, b, padding=k.shape[-1] // 2)
        return torch.relu(bn(y, *cur.take(4)))

    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x_nhwc))).to(dt).permute(0, 3, 1, 2)
        skips = []
        for _ in filters:
            for _ in range(nconv):
                x = conv_bn_relu(x)
            skips.append(x)
            x = F.max_pool2d(x, 2, 2)
        for _ in range(nconv):
            x = conv_bn_relu(x)
        for i in range(len(filters) - 1, -1, -1):
            k, b = cur.take(2)
            up = F.conv_transpose2d(x, k, b, stride=2)
            x = torch.cat([skips[i], up], dim=1)
            x = torch.relu(bn(x, *cur.take(4)))
            x = conv_bn_relu(x)
            x = conv_bn_relu(x)
        k, b = cur.take(2)
        logits = F.conv2d(x, k, b)
        assert cur.i == len(prep.t), 'weight list length does not match the architecture'
        logits = logits.permute(0, 2, 3, 1).contiguous()
        if head == 'sigmoid':
            probs = torch.sigmoid(logits)
            classes = (probs > threshold).to(torch.int32)
        else:
            probs = torch.softmax(logits, dim=-1)
            classes = torch.argmax(probs, dim=-1).to(torch.int32)
        if return_logits:
            return probs.float().numpy(), classes.numpy(), logits.float().numpy()
        return probs.float().numpy(), classes.numpy()


def make_predict_fn(weights, **kw):
    """A ``keras.Model.predict`` stand-in for ``oracle.tiling`` functions:
    batch (N,h,w,C) -> probs (N,h,w,k)."""
    prep = prepare(weights, kw.pop('precision', 'fp32'))

    def predict(batch, verbose=0, steps=None):
        return forward(np.asarray(batch, dtype=np.float32), prep, **kw)[0]
    return predict


# --------------------------------------------------- naive numpy pin (tiny shapes)
def naive_conv2d_same(x, kernel, bias):
    """Direct-loop NHWC 'same' cross-correlation, float64. x (H,W,Cin)."""
    H, W, _ = x.shape
    kh, kw, _, co = kernel.shape
    ph, pw = kh // 2, kw // 2
    out = np.zeros((H, W, co))
    for y in range(H):
        for xx in range(W):
            acc = bias.astype(np.float64).copy()
            for a in range(kh):
                for b in range(kw):
                    yy, xc = y + a - ph, xx + b - pw
                    if 0 <= yy < H and 0 <= xc < W:
                        acc += x[yy, xc, :].astype(np.float64) @ kernel[a, b].astype(np.float64)
            out[y, xx] = acc
    return out


def naive_bn(x, g, be, mu, var):
    return g * (x - mu) / np.sqrt(var + BN_EPS) + be


def naive_maxpool2(x):
    H, W, C = x.shape
    return x.reshape(H // 2, 2, W // 2, 2, C).max(axis=(1, 3))


def naive_conv2d_transpose2(x, kernel, bias):
    """Conv2DTranspose k=2 s=2, kernel (2,2,out,in). x (h,w,Cin) -> (2h,2w,out)."""
    h, w, _ = x.shape
    co = kernel.shape[2]
    out = np.zeros((2 * h, 2 * w, co))
    for i in range(h):
        for j in range(w):
            for a in range(2):
                for b in range(2):
                    out[2 * i + a, 2 * j + b] = kernel[a, b].astype(np.float64) @ x[i, j].astype(np.float64) + bias
    return out


def naive_forward(x_hwc, weights, variant='A', filters=DEFAULT_FILTERS, head='sigmoid'):
    """Float64 loop implementation of the same network (tiny inputs only)."""
    nconv = 2 if variant == 'A' else 1
    cur = _Cursor([np.asarray(w, dtype=np.float64) for w in weights])

    def cbr(x):
        k, b = cur.take(2)
        return np.maximum(naive_bn(naive_conv2d_same(x, k, b), *cur.take(4)), 0.0)

    x = np.asarray(x_hwc, dtype=np.float64)
    skips = []
    for _ in filters:
        for _ in range(nconv):
            x = cbr(x)
        skips.append(x)
        x = naive_maxpool2(x)
    for _ in range(nconv):
        x = cbr(x)
    for i in range(len(filters) - 1, -1, -1):
        k, b = cur.take(2)
        up = naive_conv2d_transpose2(x, k, b)
        x = np.concatenate([skips[i], up], axis=-1)
        x = np.maximum(naive_bn(x, *cur.take(4)), 0.0)
        x = cbr(x)
        x = cbr(x)
    k, b = cur.take(2)
    logits = x @ k[0, 0] + b
    if head == 'sigmoid':
        return 1.0 / (1.0 + np.exp(-logits))
    e = np.exp(logits - logits.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


# ----------------------------------------------------------------------- FLOPs
def flops_per_tile(h, w, variant='A', nchannels=6, nclasses=1, filters=DEFAULT_FILTERS):
    """Algorithmic FLOPs of one (h,w) tile (2*H*W*Cin*Cout*kh*kw
    per conv, 2*h*w*Cin*Cout*4 per ConvT; padding channels not counted)."""
    total = 0
    hh, ww = h, w
    dims = []
    for name, shape in weight_specs(variant, nchannels, nclasses, filters):
        if not name.endswith('kernel'):
            continue
        dims.append((name, shape))
    # walk resolutions
    level_hw = [(h >> i, w >> i) for i in range(len(filters) + 1)]
    for name, shape in dims:
        if name.startswith('encoder_'):
            lvl = int(name.split('/')[0].split('_')[1])
            hh, ww = level_hw[lvl]
            total += 2 * hh * ww * shape[2] * shape[3] * 9
        elif name.startswith('center'):
            hh, ww = level_hw[len(filters)]
            total += 2 * hh * ww * shape[2] * shape[3] * 9
        elif '/up/' in name:
            lvl = int(name.split('/')[0].split('_')[1])
            hh, ww = level_hw[lvl + 1]
            total += 2 * hh * ww * shape[3] * shape[2] * 4
        elif name.startswith('decoder_'):
            lvl = int(name.split('/')[0].split('_')[1])
            hh, ww = level_hw[lvl]
            total += 2 * hh * ww * shape[2] * shape[3] * 9
        else:  # head
            total += 2 * h * w * shape[2] * shape[3]
    return total
