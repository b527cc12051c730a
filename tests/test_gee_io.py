"""GEE patch-file formats and GeoTIFF output without TensorFlow / rasterio.
Independent checks: CRC-32C known answers, tf.train.Example bytes decoded by google.protobuf with a schema
built here from the published example.proto / feature.proto field numbers, GeoTIFFs decoded by PIL."""
import json
import os
import struct

import numpy as np
import pytest

from satellite_computervision_b200 import gee_io as g


def test_crc32c_known_answers_and_block_path():
    assert g.crc32c(b'') == 0
    assert g.crc32c(b'123456789') == 0xE3069283
    assert g.crc32c(b'\x00' * 32) == 0x8A9136AA          # RFC 3720 B.4
    assert g.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert g.crc32c(bytes(range(32))) == 0x46DD794E
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, 100_003, dtype=np.uint8).tobytes()   # lane-parallel path + ragged tail
    t = g._crc_table()
    crc = 0xFFFFFFFF
    for b in data:
        crc = int(t[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    assert g.crc32c(data) == crc ^ 0xFFFFFFFF


def _example_schema():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name='scv_test_example.proto', package='scvtest', syntax='proto3')
    def msg(name):
        m = fd.message_type.add(); m.name = name; return m
    def field(m, name, num, ftype, label=1, type_name=None, oneof=None):
        f = m.field.add(); f.name, f.number, f.type, f.label = name, num, ftype, label
        if type_name: f.type_name = type_name
        if oneof is not None: f.oneof_index = oneof
    T = descriptor_pb2.FieldDescriptorProto
    bl = msg('BytesList'); field(bl, 'value', 1, T.TYPE_BYTES, 3)
    fl = msg('FloatList'); field(fl, 'value', 1, T.TYPE_FLOAT, 3)
    il = msg('Int64List'); field(il, 'value', 1, T.TYPE_INT64, 3)
    fe = msg('Feature'); fe.oneof_decl.add().name = 'kind'
    field(fe, 'bytes_list', 1, T.TYPE_MESSAGE, 1, '.scvtest.BytesList', 0)
    field(fe, 'float_list', 2, T.TYPE_MESSAGE, 1, '.scvtest.FloatList', 0)
    field(fe, 'int64_list', 3, T.TYPE_MESSAGE, 1, '.scvtest.Int64List', 0)
    fs = msg('Features')
    entry = fs.nested_type.add(); entry.name = 'FeatureEntry'; entry.options.map_entry = True
    field(entry, 'key', 1, T.TYPE_STRING); field(entry, 'value', 2, T.TYPE_MESSAGE, 1, '.scvtest.Feature')
    field(fs, 'feature', 1, T.TYPE_MESSAGE, 3, '.scvtest.Features.FeatureEntry')
    ex = msg('Example'); field(ex, 'features', 1, T.TYPE_MESSAGE, 1, '.scvtest.Features')
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName('scvtest.Example'))


def test_example_bytes_agree_with_protobuf():
    Example = _example_schema()
    rng = np.random.default_rng(1)
    feats = {'B2': rng.random(50).astype(np.float32), 'B11': rng.random(7).astype(np.float32),
             'label': np.array([0, 1, -5, 2 ** 40], dtype=np.int64), 'name': [b'abc', b'']}
    rec = g.build_example(feats)
    ex = Example.FromString(rec)               # our writer, protobuf's reader
    assert np.array_equal(np.array(ex.features.feature['B2'].float_list.value, np.float32), feats['B2'])
    assert list(ex.features.feature['label'].int64_list.value) == feats['label'].tolist()
    assert list(ex.features.feature['name'].bytes_list.value) == feats['name']
    ex2 = Example()                            # protobuf's writer, our reader
    for k in ('B2', 'B11'):
        ex2.features.feature[k].float_list.value.extend(feats[k].tolist())
    ex2.features.feature['label'].int64_list.value.extend(feats['label'].tolist())
    ex2.features.feature['name'].bytes_list.value.extend(feats['name'])
    got = g.parse_example(ex2.SerializeToString())
    assert np.array_equal(got['B2'], feats['B2']) and np.array_equal(got['B11'], feats['B11'])
    assert np.array_equal(got['label'], feats['label']) and got['name'] == feats['name']


@pytest.mark.parametrize('compression', ['GZIP', ''])
def test_tfrecord_framing_round_trip(tmp_path, compression):
    recs = [b'', b'x', os.urandom(5000)]
    p = str(tmp_path / ('a.tfrecord.gz' if compression else 'a.tfrecords'))
    g.write_tfrecords(p, recs, compression)
    assert list(g.read_tfrecords(p, verify=True)) == recs
    if not compression:
        raw = open(p, 'rb').read()
        # first record: length 0, masked crc of the 8 length bytes, no data, masked crc of b''
        assert raw[:8] == struct.pack('<Q', 0) and struct.unpack('<I', raw[12:16])[0] == g.masked_crc32c(b'')
        bad = bytearray(raw); bad[-10] ^= 1
        open(p, 'wb').write(bytes(bad))
        with pytest.raises(ValueError):
            list(g.read_tfrecords(p, verify=True))


def test_patch_files_to_dataset_order_and_layout(tmp_path):
    """Files are consumed in sorted order, bands stacked in `features` order -> HWC (reference :175, :195-204)."""
    from satellite_computervision_b200 import prediction_tools as pt
    rng = np.random.default_rng(2)
    feats = ['B2', 'B3', 'B4']
    k, b = 8, 4
    patches = rng.random((5, k + b, k + b, 3)).astype(np.float32)
    g.write_patch_tfrecords(str(tmp_path / 'img00001.tfrecord.gz'), patches[3:], feats)
    g.write_patch_tfrecords(str(tmp_path / 'img00000.tfrecord.gz'), patches[:3], feats)
    files = [str(tmp_path / 'img00001.tfrecord.gz'), str(tmp_path / 'img00000.tfrecord.gz')]
    ds = pt.make_pred_dataset(files, feats, [k, k], [b, b], moments=[(0.0, 1.0)] * 3)
    got = list(ds)
    assert len(got) == 5 and all(t.raw.shape == (1, k + b, k + b, 3) for t in got)
    assert np.array_equal(np.concatenate([t.raw for t in got]), patches)
    assert got[0].norm.mode == 1 and np.allclose(got[0].norm.div, 1.0 + 1e-8)
    # one-hot + derived band are appended un-normalised (identity sub/div)
    lab = rng.integers(0, 3, (2, k + b, k + b)).astype(np.float32)
    recs = [g.build_example({'B2': patches[i, ..., 0].ravel(), 'B3': patches[i, ..., 1].ravel(), 'lc': lab[i].ravel()}) for i in range(2)]
    g.write_tfrecords(str(tmp_path / 'oh.tfrecord.gz'), recs)
    ds2 = pt.make_pred_dataset([str(tmp_path / 'oh.tfrecord.gz')], ['B2', 'B3', 'lc'], [k, k], [b, b], moments=[(0.0, 2.0)] * 2,
                               one_hot={'lc': 3}, ndvi=lambda d: d['B3'] - d['B2'])
    t = next(iter(ds2))
    assert t.raw.shape == (1, k + b, k + b, 2 + 1 + 3)
    assert np.array_equal(t.raw[0, ..., 2], patches[0, ..., 1] - patches[0, ..., 0])
    assert np.array_equal(t.raw[0, ..., 3:].argmax(-1), lab[0].astype(int))
    assert list(t.norm.div[2:]) == [1.0] * 4 and list(t.norm.sub[2:]) == [0.0] * 4


def test_prediction_tfrecords_match_reference_layout(tmp_path):
    from satellite_computervision_b200 import prediction_tools as pt
    rng = np.random.default_rng(3)
    k, b = 8, 4
    probs = rng.random((3, k + b, k + b, 2)).astype(np.float32)
    classes = rng.integers(0, 2, (3, k + b, k + b)).astype(np.int32)
    out = pt.write_tfrecord_predictions([probs, classes], str(tmp_path), 'pred', [k, k], [b, b])
    assert out.endswith('pred.tfrecords')
    recs = [g.parse_example(r) for r in g.read_tfrecords(out, verify=True)]
    assert len(recs) == 3 and sorted(recs[0]) == ['b1', 'b2', 'b3']
    for i, r in enumerate(recs):
        assert np.array_equal(r['b1'].reshape(k, k), probs[i, 2:10, 2:10, 0])
        assert np.array_equal(r['b3'].reshape(k, k), classes[i, 2:10, 2:10].astype(np.float32))


@pytest.mark.parametrize('dtype,bands', [('float32', 1), ('uint8', 1), ('float32', 3), ('uint16', 1)])
def test_geotiff_is_readable_by_pil_with_geo_tags(tmp_path, dtype, bands):
    from PIL import Image
    rng = np.random.default_rng(4)
    img = (rng.random((37, 53, bands)) * 200).astype(dtype)
    mixer = {'projection': {'crs': 'EPSG:32618', 'affine': {'doubleMatrix': [10.0, 0.0, 300000.0, 0.0, -10.0, 4500000.0]}},
             'patchesPerRow': 1, 'totalPatches': 1, 'patchDimensions': [8, 8]}
    jf = tmp_path / 'mixer.json'
    jf.write_text(json.dumps(mixer))
    from satellite_computervision_b200 import prediction_tools as pt
    path = pt.write_geotiff_prediction(img if bands > 1 else img[..., 0], str(jf), str(tmp_path / 'aoi'))
    if bands == 1:      # PIL: pixels + the geo tags
        with Image.open(path) as im:
            assert im.size == (53, 37)
            tags = im.tag_v2
            assert tuple(tags[33550]) == (10.0, 10.0, 0.0)
            assert tuple(tags[33922]) == (0.0, 0.0, 0.0, 300000.0, 4500000.0, 0.0)
            keys = tuple(tags[34735])
            assert keys[:4] == (1, 1, 0, 3) and keys[4:8] == (1024, 0, 1, 1) and keys[12:16] == (3072, 0, 1, 32618)
            assert np.array_equal(np.array(im), img[..., 0])
    else:               # multi-band float: libtiff through OpenCV (PIL has no 3 x float32 mode)
        import cv2
        arr = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        assert arr is not None and arr.shape == (37, 53, 3) and np.array_equal(arr[..., ::-1], img)


def test_geotiff_rotated_transform_and_many_strips(tmp_path):
    from PIL import Image
    img = np.arange(600 * 500, dtype=np.float32).reshape(600, 500)
    path = g.write_geotiff(str(tmp_path / 'r.tif'), img, (9.0, 1.0, 5.0, -1.0, -9.0, 7.0), 'EPSG:4326', rows_per_strip=7)
    with Image.open(path) as im:
        assert np.array_equal(np.array(im), img)
        m = tuple(im.tag_v2[34264])
        assert m[:4] == (9.0, 1.0, 0.0, 5.0) and m[4:8] == (-1.0, -9.0, 0.0, 7.0) and m[15] == 1.0
        assert tuple(im.tag_v2[34735])[4:8] == (1024, 0, 1, 2)


def test_bigtiff_container(tmp_path):
    """The 64-bit container (used above 4 GB) written for a small image: OpenCV's libtiff reads it back."""
    import cv2
    img = (np.random.default_rng(5).random((33, 47)) * 100).astype(np.float32)
    path = g.write_geotiff(str(tmp_path / 'big.tif'), img, (10.0, 0.0, 1.0, 0.0, -10.0, 2.0), 'EPSG:32618', rows_per_strip=5, bigtiff=True)
    raw = open(path, 'rb').read(16)
    assert raw[:4] == b'II\x2b\x00' and raw[4:8] == b'\x08\x00\x00\x00'
    arr = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert arr is not None and np.array_equal(arr, img)
