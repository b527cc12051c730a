"""The fused decoder tail (conv_fused.cuh: decoder_0/conv0 -> decoder_0/conv1 + head in one cluster launch, the
32-channel intermediate staying in shared memory) against the two-launch path: bit-identical logits / probabilities."""
import numpy as np
import pytest

from oracle import unet as ounet
from satellite_computervision_b200 import model_tools

pytestmark = pytest.mark.gpu


def _predict(x, w, filters, fuse, monkeypatch, max_batch):
    monkeypatch.setenv('SCV_FUSE', '1' if fuse else '0')
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
    (p1, c1), n1 = _predict(x, w, filters, True, monkeypatch, max_batch)
    (p0, c0), n0 = _predict(x, w, filters, False, monkeypatch, max_batch)
    assert n1 < n0, 'the fused plan must launch fewer kernels'
    assert np.array_equal(p1, p0) and np.array_equal(c1, c0)
    ref_p, _ = ounet.forward(x[:1], w, 'A', filters)
    assert np.abs(p1[:1] - ref_p).max() <= 1e-2
