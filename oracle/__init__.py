"""CPU oracle for the tiled U-Net predict path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy for the integer/index work, torch-CPU
fp32/fp64 for the float contractions), the algorithm of the reference's tiled
predict path (mjevans26/Satellite_ComputerVision):

* ``oracle.tiling``    -- ``utils/prediction_tools.py:87-156, :245-373, :475-520``
* ``oracle.normalize`` -- ``utils/processing.py:225-322``, ``utils/pc_tools.py:90-107``
* ``oracle.unet``      -- ``utils/model_tools.py:174-454`` and
  ``notebooks/UNET_G4G_2019_solar.ipynb:1162-1213``
* ``oracle.siamese``   -- ``utils/model_tools.py:533-663`` (siamese U-Net with the atrous pyramid; float arithmetic
  unpinned like the U-Net's, cross-checked against ``oracle.siamese.naive_*`` in ``tests/test_siamese.py``)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline.  Nothing under ``satellite_computervision_b200/`` imports it;
the product path has no CPU fallback.

PINNING STATUS
--------------
* Index / crop / stitch / normalise arithmetic: PINNED against the reference's
  own Python source, imported in the build container with TensorFlow,
  matplotlib and rasterio stubbed out (``tests/golden/make_golden.py`` ->
  ``tests/golden/*.npz``).
* Float network arithmetic (Conv2D / BatchNormalization / MaxPooling2D /
  Conv2DTranspose / softmax / sigmoid): **parity unpinned**.  All of it lives in
  third-party TensorFlow/Keras (unpinned version, not vendored, not installable
  here: no network, no wheel).  The reference has no tests, golden vectors or
  saved weights.  The oracle restates the published Keras semantics (HWIO
  cross-correlation with 'same' zero padding, BN eps=1e-3 with moving
  statistics, 'valid' 2x2 max-pool, Conv2DTranspose kernel (kh,kw,out,in),
  softmax over the last axis, argmax ties -> lowest index, strict ``greater``)
  and is cross-checked against an independent naive numpy-loop implementation
  (``oracle.unet.naive_*``) in ``tests/test_oracle_unet.py``.
"""
