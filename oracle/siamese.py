"""Oracle (test infrastructure): the reference's siamese U-Net forward pass on the CPU.

Restates ``make_siamese_unet`` / ``get_siamese_layers`` / ``DilatedSpatialPyramidPooling``
(``utils/model_tools.py:533-663``) with torch-CPU fp32 (or fp64) contractions and *un-folded* BatchNorm, plus an
independent float64 loop implementation for tiny inputs.  TensorFlow is not installable here: like ``oracle/unet.py``
this restates the published Keras semantics -- **parity unpinned** (``oracle/__init__.py``).

The network as the reference builds it:

* one shared ``encoder_block`` per level, applied to both images (``:600-609``); ``conv_block.call`` applies ``cba1``
  twice to its *input* and never builds ``cba2`` (``:238-239``), so an encoder block is ONE conv-BN-ReLU + 2x2 max-pool;
  the skip of level i is ``concat([encoded_b, encoded_a])`` (``:604, :609``);
* one shared ``DilatedSpatialPyramidPooling(filters[-1] * 2)`` on both pooled images (``:612-614``):
  ``cba`` 1x1, ``cba3_3 / cba3_6 / cba3_12`` 3x3 at dilation 3 / 6 / 12 ('same' padding), concat in that order, ``cba3``
  1x1 (``:561-573``; ``cba2`` and the pooling branch are never called);
* ``squeezed = concat([aspp_b, aspp_a])`` (``:620``);
* ``decoder_block`` per level (``:288-317``): Conv2DTranspose 2x2/2 -> ``concatenate([skip, up])`` -> BN -> ReLU ->
  2 x (conv 3x3 -> BN -> ReLU);
* ``Conv2D(1, (1,1), sigmoid)`` + ``int32(p > class_thresh)`` (``:659-660``).

Weight list order used here and by the engine (``scv_weight_shape``): per conv-BN unit (kernel, bias, gamma, beta,
moving_mean, moving_variance) in the order encoder_0.., ASPP/cba, ASPP/cba3, ASPP/cba3_3, _6, _12, decoder_{L-1}..0,
head.  ``tf.keras`` 2 lists a composite layer's trainable weights before its non-trainable ones, so
``model.get_weights()`` of the reference interleaves the ASPP layer differently; :func:`from_keras2_order` maps that
list onto this one.
"""
from __future__ import annotations

import numpy as np

from .unet import BN_EPS, _Cursor, _Prepared, naive_bn, naive_conv2d_transpose2, naive_maxpool2, prepare  # noqa: F401

DEFAULT_FILTERS = (32, 64, 128)
ASPP_RATES = (3, 6, 12)


def weight_specs(nchannels=3, filters=DEFAULT_FILTERS):
    specs = []

    def conv(name, cin, cout, k=3):
        specs.append((f'{name}/kernel', (k, k, cin, cout)))
        specs.append((f'{name}/bias', (cout,)))

    def bn(name, c):
        for w in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            specs.append((f'{name}/{w}', (c,)))

    cin = nchannels
    for i, f in enumerate(filters):
        conv(f'encoder_{i}/conv0', cin, f)
        bn(f'encoder_{i}/bn0', f)
        cin = f
    nf = 2 * filters[-1]
    conv('ASPP/cba/conv', cin, nf, 1)
    bn('ASPP/cba/bn', nf)
    conv('ASPP/cba3/conv', 4 * nf, nf, 1)
    bn('ASPP/cba3/bn', nf)
    for r in ASPP_RATES:
        conv(f'ASPP/cba3_{r}/conv', cin, nf, 3)
        bn(f'ASPP/cba3_{r}/bn', nf)
    cin = 2 * nf
    for i in range(len(filters) - 1, -1, -1):
        f = filters[i]
        specs.append((f'decoder_{i}/up/kernel', (2, 2, f, cin)))
        specs.append((f'decoder_{i}/up/bias', (f,)))
        bn(f'decoder_{i}/bn_cat', 3 * f)
        conv(f'decoder_{i}/conv0', 3 * f, f)
        bn(f'decoder_{i}/bn0', f)
        conv(f'decoder_{i}/conv1', f, f)
        bn(f'decoder_{i}/bn1', f)
        cin = f
    conv('head', cin, 1, k=1)
    return specs


def keras2_permutation(nlevels):
    """Index list p with ``engine_order[i] = keras2_order[p[i]]``: ``tf.keras`` 2 returns, for the composite ASPP layer,
    the trainable weights of its five conv-BN units (kernel, bias, gamma, beta each) and then their moving statistics
    (mean, variance each); every other layer of the model holds a single conv / BN, where both orders coincide."""
    n_enc = 6 * nlevels
    p = list(range(n_enc))
    base = n_enc
    for u in range(5):
        p += [base + 4 * u + j for j in range(4)] + [base + 20 + 2 * u + j for j in range(2)]
    rest = base + 30
    return p + list(range(rest, rest + 18 * nlevels + 2))


def from_keras2_order(weights, nlevels=len(DEFAULT_FILTERS)):
    return [weights[i] for i in keras2_permutation(nlevels)]


def forward(a_nhwc, b_nhwc, weights, filters=DEFAULT_FILTERS, threshold=0.5, precision='fp32', return_logits=False,
            num_threads=None):
    """``model.predict([input_a, input_b])``: two (N,H,W,C) arrays -> (probs (N,H,W,1) float32, classes (N,H,W,1) int32)."""
    import torch
    import torch.nn.functional as F
    if num_threads:
        torch.set_num_threads(num_threads)
    prep = weights if isinstance(weights, _Prepared) else prepare(weights, precision)
    dt = prep.dt
    cur = _Cursor(prep.t)

    def bn(y, g, be, mu, var):
        g, be, mu, var = (v.view(1, -1, 1, 1) for v in (g, be, mu, var))
        return g * (y - mu) / torch.sqrt(var + BN_EPS) + be

    def unit(dil=1):
        k, b = cur.take(2)
        bnw = cur.take(4)
        pad = dil * (k.shape[-1] // 2)
        return lambda x: torch.relu(bn(F.conv2d(x, k, b, padding=pad, dilation=dil), *bnw))

    with torch.no_grad():
        xa = torch.from_numpy(np.ascontiguousarray(np.asarray(a_nhwc))).to(dt).permute(0, 3, 1, 2)
        xb = torch.from_numpy(np.ascontiguousarray(np.asarray(b_nhwc))).to(dt).permute(0, 3, 1, 2)
        skips = []
        for _ in filters:
            enc = unit()
            ea, eb = enc(xa), enc(xb)
            skips.append(torch.cat([eb, ea], dim=1))
            xa, xb = F.max_pool2d(ea, 2, 2), F.max_pool2d(eb, 2, 2)
        cba, cba3 = unit(), unit()
        branches = [cba] + [unit(r) for r in ASPP_RATES]

        def aspp(x):
            return cba3(torch.cat([br(x) for br in branches], dim=1))
        x = torch.cat([aspp(xb), aspp(xa)], dim=1)
        for i in range(len(filters) - 1, -1, -1):
            k, b = cur.take(2)
            up = F.conv_transpose2d(x, k, b, stride=2)
            x = torch.cat([skips[i], up], dim=1)
            x = torch.relu(bn(x, *cur.take(4)))
            x = unit()(x)
            x = unit()(x)
        k, b = cur.take(2)
        logits = F.conv2d(x, k, b)
        assert cur.i == len(prep.t), 'weight list length does not match the architecture'
        logits = logits.permute(0, 2, 3, 1).contiguous()
        probs = torch.sigmoid(logits)
        classes = (probs > threshold).to(torch.int32)
        if return_logits:
            return probs.float().numpy(), classes.numpy(), logits.float().numpy()
        return probs.float().numpy(), classes.numpy()


def make_predict_fn(weights, nchannels, **kw):
    """``keras.Model.predict`` stand-in for ``oracle.tiling``: batch (N,h,w,2C) with the two images stacked along the
    channel axis ([a | b]) -> probs (N,h,w,1)."""
    prep = prepare(weights, kw.pop('precision', 'fp32'))

    def predict(batch, verbose=0, steps=None):
        batch = np.asarray(batch, dtype=np.float32)
        return forward(batch[..., :nchannels], batch[..., nchannels:], prep, **kw)[0]
    return predict


# --------------------------------------------------- naive float64 pin (tiny shapes)
def naive_conv2d_same_dilated(x, kernel, bias, dil=1):
    """Direct-loop NHWC 'same' cross-correlation with dilation, float64.  x (H,W,Cin)."""
    H, W, _ = x.shape
    kh, kw, _, co = kernel.shape
    out = np.zeros((H, W, co))
    for y in range(H):
        for xx in range(W):
            acc = bias.astype(np.float64).copy()
            for a in range(kh):
                for b in range(kw):
                    yy, xc = y + (a - kh // 2) * dil, xx + (b - kw // 2) * dil
                    if 0 <= yy < H and 0 <= xc < W:
                        acc += x[yy, xc, :].astype(np.float64) @ kernel[a, b].astype(np.float64)
            out[y, xx] = acc
    return out


def naive_forward(a_hwc, b_hwc, weights, filters=DEFAULT_FILTERS):
    cur = _Cursor([np.asarray(w, dtype=np.float64) for w in weights])

    def unit(dil=1):
        k, b = cur.take(2)
        bnw = cur.take(4)
        return lambda x: np.maximum(naive_bn(naive_conv2d_same_dilated(x, k, b, dil), *bnw), 0.0)

    xa, xb = np.asarray(a_hwc, dtype=np.float64), np.asarray(b_hwc, dtype=np.float64)
    skips = []
    for _ in filters:
        enc = unit()
        ea, eb = enc(xa), enc(xb)
        skips.append(np.concatenate([eb, ea], axis=-1))
        xa, xb = naive_maxpool2(ea), naive_maxpool2(eb)
    cba, cba3 = unit(), unit()
    branches = [cba] + [unit(r) for r in ASPP_RATES]

    def aspp(x):
        return cba3(np.concatenate([br(x) for br in branches], axis=-1))
    x = np.concatenate([aspp(xb), aspp(xa)], axis=-1)
    for i in range(len(filters) - 1, -1, -1):
        k, b = cur.take(2)
        up = naive_conv2d_transpose2(x, k, b)
        x = np.concatenate([skips[i], up], axis=-1)
        x = np.maximum(naive_bn(x, *cur.take(4)), 0.0)
        x = unit()(x)
        x = unit()(x)
    k, b = cur.take(2)
    return 1.0 / (1.0 + np.exp(-(x @ k[0, 0] + b)))
