"""Oracle (test infrastructure): pin ``oracle.unet`` against real Keras when TensorFlow is importable.

TensorFlow is not installable in the build image (no wheel, no network), which is why the float arithmetic of
the oracle is "parity unpinned" (``oracle/__init__.py``).  SURVEY 8(c) asks for a run-time probe on the GPU box:
if ``import tensorflow`` works there, this module builds the reference's layer stack with ``tf.keras`` --
``conv_batch_act`` / ``conv_block`` / ``encoder_block`` / ``decoder_block`` wired as ``build_unet_layers`` does
(``utils/model_tools.py:174-186, :211-240, :262-286, :288-318, :321-379``), heads as ``:405-406`` / ``:443-445``
and the notebook's two-conv block (``notebooks/UNET_G4G_2019_solar.ipynb:1162-1213``) -- loads the oracle's
deterministic weights with ``model.set_weights`` and compares one ``model.predict`` with ``oracle.unet.forward``.
Called from ``tests/test_gpu_baseline_configs.py`` (skipped without TensorFlow) and from
``bench.py --impl reference`` (reported in the JSON line).
"""
from __future__ import annotations

import numpy as np

from . import unet as ounet


def build_keras(tf, variant='A', nchannels=6, nclasses=1, filters=(32, 64), head='sigmoid'):
    """The reference network in tf.keras functional form; layers are created in the order of
    ``oracle.unet.weight_specs`` so that ``get_weights()`` lines up."""
    L = tf.keras.layers
    nconv = 2 if variant == 'A' else 1

    def cba(x, f):  # conv_batch_act, model_tools.py:178-186
        x = L.Conv2D(f, (3, 3), padding='same')(x)
        x = L.BatchNormalization()(x)
        return L.ReLU()(x)

    inp = L.Input((None, None, nchannels))
    x, skips = inp, []
    for f in filters:
        for _ in range(nconv):
            x = cba(x, f)
        skips.append(x)
        x = L.MaxPooling2D((2, 2), strides=(2, 2))(x)
    for _ in range(nconv):
        x = cba(x, filters[-1] * 2)
    for i in range(len(filters) - 1, -1, -1):  # decoder_block, :288-318
        f = filters[i]
        up = L.Conv2DTranspose(f, (2, 2), strides=(2, 2), padding='same')(x)
        x = L.concatenate([skips[i], up], axis=-1)
        x = L.BatchNormalization()(x)
        x = L.Activation('relu')(x)
        x = cba(x, f)
        x = cba(x, f)
    act = 'sigmoid' if head == 'sigmoid' else 'softmax'
    out = L.Conv2D(nclasses, (1, 1), activation=act)(x)
    return tf.keras.Model(inp, out)


def compare(tf, variant='A', filters=(32, 64), hw=64, seed=0):
    """max |keras.predict - oracle.forward| on one random batch (fp32 both)."""
    head = 'sigmoid' if variant == 'A' else 'softmax'
    ncls = 1 if variant == 'A' else 2
    specs = ounet.weight_specs(variant, 6, ncls, tuple(filters))
    w = ounet.init_weights(specs, seed=seed, randomize_bn=True)
    model = build_keras(tf, variant, 6, ncls, tuple(filters), head)
    kw = model.get_weights()
    assert [tuple(a.shape) for a in kw] == [tuple(s) for _, s in specs], 'keras get_weights() order differs from the oracle'
    model.set_weights(w)
    x = np.random.default_rng(seed + 1).random((2, hw, hw, 6)).astype(np.float32)
    want = model.predict(x, verbose=0)
    got, _ = ounet.forward(x, w, variant, tuple(filters), head=head)
    return float(np.abs(np.asarray(want) - got).max())


def build_keras_siamese(tf, nchannels=3, filters=(32, 64)):
    """``make_siamese_unet`` (``utils/model_tools.py:638-663``) with the reference's own layer classes restated:
    ``conv_batch_act`` (``:178-186``), ``conv_block`` whose ``call`` applies ``cba1`` twice to its input (``:211-240``),
    ``encoder_block`` (``:262-286``), ``DilatedSpatialPyramidPooling`` (``:533-574``), ``decoder_block`` (``:288-318``)."""
    L = tf.keras.layers

    class ConvBatchAct(L.Layer):
        def __init__(self, f, kernel_size=(3, 3), dilation_rate=1, **kw):
            super().__init__(**kw)
            self.conv_layer = L.Conv2D(f, kernel_size, padding='same', dilation_rate=dilation_rate)
            self.bn_layer = L.BatchNormalization()
            self.activation_layer = L.Activation('relu')

        def call(self, inputs):
            return self.activation_layer(self.bn_layer(self.conv_layer(inputs)))

    class ConvBlock(L.Layer):
        def __init__(self, f, **kw):
            super().__init__(**kw)
            self.cba1 = ConvBatchAct(f)
            self.cba2 = ConvBatchAct(f)  # never called, hence never built: no weights (model_tools.py:238-239)

        def call(self, inputs):
            y = self.cba1(inputs)
            y = self.cba1(inputs)
            return y

    class EncoderBlock(L.Layer):
        def __init__(self, f, **kw):
            super().__init__(**kw)
            self.encoder = ConvBlock(f)
            self.pooler = L.MaxPooling2D((2, 2), strides=(2, 2))

        def call(self, x):
            encoded = self.encoder(x)
            return self.pooler(encoded), encoded

    class ASPP(L.Layer):
        def __init__(self, f, **kw):
            super().__init__(**kw)
            self.cba = ConvBatchAct(f, (1, 1))
            self.cba2 = ConvBatchAct(f, (1, 1))  # never called
            self.cba3 = ConvBatchAct(f, (1, 1))
            self.cba3_3 = ConvBatchAct(f, (3, 3), 3)
            self.cba3_6 = ConvBatchAct(f, (3, 3), 6)
            self.cba3_12 = ConvBatchAct(f, (3, 3), 12)

        def call(self, x):
            return self.cba3(L.Concatenate(axis=-1)([self.cba(x), self.cba3_3(x), self.cba3_6(x), self.cba3_12(x)]))

    def decoder_block(x, skip, f):
        d = L.Conv2DTranspose(f, (2, 2), strides=(2, 2), padding='same')(x)
        d = L.concatenate([skip, d], axis=-1)
        d = L.Activation('relu')(L.BatchNormalization()(d))
        d = L.Activation('relu')(L.BatchNormalization()(L.Conv2D(f, (3, 3), padding='same')(d)))
        d = L.Activation('relu')(L.BatchNormalization()(L.Conv2D(f, (3, 3), padding='same')(d)))
        return d

    a, b = L.Input((None, None, nchannels)), L.Input((None, None, nchannels))
    pa, pb, net = a, b, []
    for f in filters:
        enc = EncoderBlock(f)
        pa, ea = enc(pa)
        pb, eb = enc(pb)
        net.append(L.Concatenate(axis=-1)([eb, ea]))
    aspp = ASPP(filters[-1] * 2)
    d = L.Concatenate(axis=-1)([aspp(pb), aspp(pa)])
    for j in range(len(filters) - 1, -1, -1):
        d = decoder_block(d, net[j], filters[j])
    probs = L.Conv2D(1, (1, 1), activation='sigmoid')(d)
    return tf.keras.Model([a, b], probs)


def compare_siamese(tf, filters=(32, 64), hw=64, seed=0):
    """max |keras.predict([a, b]) - oracle.siamese.forward| with the oracle's weights mapped onto whatever order
    ``get_weights()`` of the installed Keras uses for the composite ASPP layer (tf.keras 2: trainable first; Keras 3:
    per sub-layer) -- the order found is part of the result."""
    from . import siamese as osi
    specs = osi.weight_specs(3, tuple(filters))
    w = ounet.init_weights(specs, seed=seed, randomize_bn=True)
    model = build_keras_siamese(tf, 3, tuple(filters))
    shapes = [tuple(x.shape) for x in model.get_weights()]
    perm = osi.keras2_permutation(len(filters))
    keras2 = [None] * len(w)
    for i, src in enumerate(perm):
        keras2[src] = w[i]
    if shapes == [tuple(x.shape) for x in keras2]:
        order, wl = 'keras2', keras2
    elif shapes == [tuple(s) for _, s in specs]:
        order, wl = 'grouped', w
    else:
        raise AssertionError('keras get_weights() order of the siamese model matches neither known order')
    model.set_weights(wl)
    rng = np.random.default_rng(seed + 1)
    xa, xb = (rng.random((2, hw, hw, 3)).astype(np.float32) for _ in range(2))
    want = model.predict([xa, xb], verbose=0)
    got, _ = osi.forward(xa, xb, w, tuple(filters))
    return float(np.abs(np.asarray(want) - got).max()), order


def probe():
    """{'tensorflow': version | None, 'max_abs_vs_oracle': float | None} -- never raises."""
    try:
        import tensorflow as tf  # noqa: WPS433 (run-time probe, by design)
    except Exception as exc:  # ImportError, or a broken install
        return {'tensorflow': None, 'why': f'{type(exc).__name__}: {exc}'[:120], 'max_abs_vs_oracle': None}
    try:
        out = {'tensorflow': tf.__version__, 'max_abs_vs_oracle': max(compare(tf, 'A'), compare(tf, 'B'))}
        try:
            out['siamese_max_abs_vs_oracle'], out['siamese_weight_order'] = compare_siamese(tf)
        except Exception as exc:  # the U-Net pin stands on its own
            out['siamese_why'] = f'{type(exc).__name__}: {exc}'[:200]
        return out
    except Exception as exc:
        return {'tensorflow': tf.__version__, 'why': f'{type(exc).__name__}: {exc}'[:200], 'max_abs_vs_oracle': None}
