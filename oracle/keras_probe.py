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


def probe():
    """{'tensorflow': version | None, 'max_abs_vs_oracle': float | None} -- never raises."""
    try:
        import tensorflow as tf  # noqa: WPS433 (run-time probe, by design)
    except Exception as exc:  # ImportError, or a broken install
        return {'tensorflow': None, 'why': f'{type(exc).__name__}: {exc}'[:120], 'max_abs_vs_oracle': None}
    try:
        return {'tensorflow': tf.__version__, 'max_abs_vs_oracle': max(compare(tf, 'A'), compare(tf, 'B'))}
    except Exception as exc:
        return {'tensorflow': tf.__version__, 'why': f'{type(exc).__name__}: {exc}'[:200], 'max_abs_vs_oracle': None}
