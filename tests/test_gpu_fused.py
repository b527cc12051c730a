"""The fused 384-pixel conv pairs (conv_fused.cuh: encoder_0/conv0 -> conv1 + pool + skip, decoder_0/conv0 -> conv1 + head, each
one cluster launch with the 32-channel intermediate staying in shared memory) against the two-launch path:
bit-identical probabilities and classes."""
import numpy as np
import pytest

from oracle import unet as ounet
from satellite_computervision_b200 import model_tools

pytestmark = pytest.mark.gpu


def _predict(x, w, filters, fuse, monkeypatch, max_batch):
    monkeypatch.setenv('SCV_FUSE', str(fuse))
    m = model_tools.binary_unet(nchannels=6, filters=list(filters), max_batch=max_batch, outputs='both')
    m.set_weights(w)
    out = m.predict(x)
    t = m.times()
    m.close()
    return out, t['n_launches']


@pytest.mark.parametrize('N,max_batch', [(1, 8), (3, 2), (5, 8)])
def test_fused_decoder_tail_is_bit_identical(N, max_batch, monkeypatch):
    filters = (32, 64)
    w = ounet.init_weights(ounet.weight_specs('A', 6, 1, filters), seed=11, randomize_bn=True)
    # square tiles only (engine contract): 384 x 384; small N / batch sizes make segments start mid-image
    x = np.random.default_rng(12).random((N, 384, 384, 6)).astype(np.float32)
    (p0, c0), n0 = _predict(x, w, filters, 0, monkeypatch, max_batch)
    for fuse in (1, 2, 3):  # decoder tail, encoder pair, both
        (p1, c1), n1 = _predict(x, w, filters, fuse, monkeypatch, max_batch)
        assert n1 < n0, 'the fused plan must launch fewer kernels'
        assert np.array_equal(p1, p0) and np.array_equal(c1, c0), fuse
    ref_p, _ = ounet.forward(x[:1], w, 'A', filters)
    assert np.abs(p1[:1] - ref_p).max() <= 1e-2
