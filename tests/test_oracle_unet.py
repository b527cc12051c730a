"""Pin oracle.unet's torch-CPU forward against an independent naive float64
loop implementation of the same Keras semantics (tiny shapes), and check the
architecture bookkeeping against SURVEY/BASELINE numbers."""
import numpy as np
import pytest

from oracle import unet


@pytest.mark.parametrize('variant,head,ncls', [('A', 'sigmoid', 1), ('B', 'softmax', 2), ('A', 'softmax', 3)])
def test_forward_matches_naive_loops(variant, head, ncls):
    filters = (4, 8)
    specs = unet.weight_specs(variant, 3, ncls, filters)
    w = unet.init_weights(specs, seed=3)
    x = np.random.default_rng(1).random((2, 8, 12, 3)).astype(np.float32)
    probs, classes = unet.forward(x, w, variant, filters, head=head)
    for n in range(2):
        ref = unet.naive_forward(x[n], w, variant, filters, head=head)
        assert np.abs(probs[n] - ref).max() < 2e-6
    if head == 'sigmoid':
        assert classes.shape == (2, 8, 12, 1) and classes.dtype == np.int32
        assert np.array_equal(classes, (probs > 0.5).astype(np.int32))
    else:
        assert classes.shape == (2, 8, 12)
        assert np.array_equal(classes, probs.argmax(-1))


def test_param_and_flop_counts_match_survey():
    a = unet.weight_specs('A', 6, 1)
    b = unet.weight_specs('B', 6, 2)
    assert sum(int(np.prod(s)) for _, s in a) == 31_127_361
    assert sum(int(np.prod(s)) for _, s in b) == 18_537_474
    assert abs(unet.flops_per_tile(384, 384, 'A') / 1e9 - 67.410) < 1e-3
    assert abs(unet.flops_per_tile(384, 384, 'B', nclasses=2) / 1e9 - 51.112) < 1e-3


def test_fp64_bounds_fp32_error():
    filters = (8, 16, 32)
    specs = unet.weight_specs('A', 6, 1, filters)
    w = unet.init_weights(specs, seed=0)
    x = np.random.default_rng(0).random((1, 32, 32, 6)).astype(np.float32)
    p32, _ = unet.forward(x, w, 'A', filters)
    p64, _ = unet.forward(x, w, 'A', filters, precision='fp64')
    assert np.abs(p32 - p64).max() < 1e-5
