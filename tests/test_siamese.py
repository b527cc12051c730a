"""Siamese U-Net + ASPP (make_siamese_unet, utils/model_tools.py:533-663), CPU side: the torch oracle against an
independent naive float64 loop implementation (tiny shapes), weight bookkeeping of oracle / engine / Python layer."""
import numpy as np
import pytest

from oracle import siamese as osi
from oracle import unet as ounet


def test_oracle_forward_matches_naive_loops():
    filters = (4, 8)
    specs = osi.weight_specs(2, filters)
    w = ounet.init_weights(specs, seed=5)
    rng = np.random.default_rng(1)
    a = rng.random((2, 16, 24, 2)).astype(np.float32)
    b = rng.random((2, 16, 24, 2)).astype(np.float32)
    probs, classes = osi.forward(a, b, w, filters)
    assert probs.shape == (2, 16, 24, 1) and classes.shape == (2, 16, 24, 1) and classes.dtype == np.int32
    for n in range(2):
        ref = osi.naive_forward(a[n], b[n], w, filters)
        assert np.abs(probs[n] - ref).max() < 2e-6
    assert np.array_equal(classes, (probs > 0.5).astype(np.int32))
    # the two inputs are not interchangeable (skip and ASPP concatenations are ordered [b, a])
    swapped, _ = osi.forward(b, a, w, filters)
    assert np.abs(swapped - probs).max() > 1e-4


def test_dilated_conv_restatement():
    """'same' dilated cross-correlation: torch vs the loop version, rate larger than the image included."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    x = rng.standard_normal((7, 9, 3))
    k = rng.standard_normal((3, 3, 3, 4))
    bias = rng.standard_normal(4)
    for dil in (1, 3, 6, 12):
        ref = osi.naive_conv2d_same_dilated(x, k, bias, dil)
        got = F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(k).permute(3, 2, 0, 1),
                       torch.from_numpy(bias), padding=dil, dilation=dil)[0].permute(1, 2, 0).numpy()
        assert np.abs(got - ref).max() < 1e-12


def test_weight_bookkeeping():
    specs = osi.weight_specs(3)
    assert len(specs) == 6 * 3 + 30 + 18 * 3 + 2
    assert sum(int(np.prod(s)) for _, s in specs) == 2_362_625
    p = osi.keras2_permutation(3)
    assert sorted(p) == list(range(len(specs)))
    # tf.keras 2 order: ASPP trainable weights (kernel, bias, gamma, beta per unit), then the moving statistics
    names = [n for n, _ in specs]
    keras2 = [None] * len(specs)
    for i, src in enumerate(p):
        keras2[src] = names[i]
    aspp = [n for n in keras2 if n.startswith('ASPP')]
    assert all(n.endswith(('kernel', 'bias', 'gamma', 'beta')) for n in aspp[:20])
    assert all(n.endswith(('moving_mean', 'moving_variance')) for n in aspp[20:])
    assert aspp[:4] == ['ASPP/cba/conv/kernel', 'ASPP/cba/conv/bias', 'ASPP/cba/bn/gamma', 'ASPP/cba/bn/beta']
    assert aspp[20:22] == ['ASPP/cba/bn/moving_mean', 'ASPP/cba/bn/moving_variance']


def test_engine_weight_list_equals_oracle():
    from satellite_computervision_b200 import model_tools
    for nch, filters in ((3, (32, 64, 128)), (6, (32, 64)), (4, (64, 128, 256))):
        m = model_tools.make_siamese_unet(nch, list(filters), [2] * len(filters))
        specs = osi.weight_specs(nch, filters)
        assert m.weight_names == [n for n, _ in specs]
        assert m.weight_shapes == [tuple(s) for _, s in specs]
        assert m.keras2_permutation() == osi.keras2_permutation(len(filters))
        w = ounet.init_weights(specs, seed=1)
        m.set_weights(w)
        keras2 = [None] * len(w)
        for i, src in enumerate(m.keras2_permutation()):
            keras2[src] = w[i]
        m2 = model_tools.make_siamese_unet(nch, list(filters), [2] * len(filters))
        m2.set_weights(keras2, order='keras2')
        assert all(np.array_equal(x, y) for x, y in zip(m.get_weights(), m2.get_weights()))
    with pytest.raises(Exception):
        model_tools.make_siamese_unet(9)  # 2 x 9 bands exceed the extract kernel's band limit
    with pytest.raises(ValueError):
        m.set_weights(w[:-1])
    with pytest.raises(NotImplementedError):
        m.load_weights('weights.h5')
    with pytest.raises(NotImplementedError):
        model_tools.make_siamese_unet(3, [32, 64], [2, 4])


def test_npz_round_trip(tmp_path):
    from satellite_computervision_b200 import model_tools
    m = model_tools.make_siamese_unet(3, seed=3)
    path = str(tmp_path / 'siamese.npz')
    m.save_weights(path)
    m2 = model_tools.make_siamese_unet(3, seed=4)
    assert not all(np.array_equal(x, y) for x, y in zip(m.get_weights(), m2.get_weights()))
    m2.load_weights(path)
    assert all(np.array_equal(x, y) for x, y in zip(m.get_weights(), m2.get_weights()))
    # a list in tf.keras 2 order goes through the permutation
    w = m.get_weights()
    keras2 = [None] * len(w)
    for i, src in enumerate(m.keras2_permutation()):
        keras2[src] = w[i]
    np.savez(str(tmp_path / 'k2.npz'), *keras2)
    m3 = model_tools.make_siamese_unet(3, seed=5)
    m3.load_weights(str(tmp_path / 'k2.npz'), order='keras2')
    assert all(np.array_equal(x, y) for x, y in zip(w, m3.get_weights()))


def test_real_keras_siamese_matches_oracle_when_tensorflow_is_present():
    """The siamese oracle restates Keras like the U-Net oracle does; when TensorFlow imports, build the reference's own
    layer classes with tf.keras and pin both the arithmetic and the get_weights() order of the composite ASPP layer."""
    tf = pytest.importorskip('tensorflow')
    from oracle import keras_probe
    err, order = keras_probe.compare_siamese(tf)
    print('siamese oracle vs tf.keras: max|dp|', err, 'weight order', order)
    assert err <= 1e-4
