"""Keras weight files without h5py: reader + writer round trips over the legacy
full-model / weights-only layouts and the Keras 3 layout, container details the parser must survive
(many links -> several symbol nodes and B-tree levels, scalar / array string attributes, name lists split
over layer_names0.., empty weight lists) and the UNetModel.load_weights / save_weights surface.
No real h5py-written file exists in this image: "parity unpinned" for the container."""
import os
import zipfile

import numpy as np
import pytest

from satellite_computervision_b200 import keras_h5 as kh


def _layers(rng, n=3):
    out = []
    for i in range(n):
        c = f'conv2d_{i}' if i else 'conv2d'
        b = f'batch_normalization_{i}' if i else 'batch_normalization'
        out.append((c, [(f'{c}/kernel:0', rng.standard_normal((3, 3, 4 + i, 8)).astype(np.float32)),
                        (f'{c}/bias:0', rng.standard_normal(8).astype(np.float32))]))
        out.append((f'activation_{i}', []))
        out.append((b, [(f'{b}/{k}:0', rng.standard_normal(8).astype(np.float32))
                        for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')]))
    return out


@pytest.mark.parametrize('full_model', [False, True])
def test_legacy_round_trip(tmp_path, full_model):
    rng = np.random.default_rng(0)
    layers = _layers(rng)
    p = kh.write_weights_h5(str(tmp_path / 'w.h5'), layers, full_model=full_model,
                            root_attrs={'model_config': np.bytes_(b'{"class_name": "Functional"}')})
    named = kh.read_named_weights(p)
    want = [(f'{ln}/{wn}', a) for ln, ws in layers for wn, a in ws]
    assert [n for n, _ in named] == [n for n, _ in want]
    for (_, got), (_, ref) in zip(named, want):
        assert got.dtype == np.float32 and np.array_equal(got, ref)
    f = kh.H5File(p)
    root = f['model_weights'] if full_model else f
    assert bytes(f.attrs['model_config']).startswith(b'{"class_name"')
    assert [x.decode() for x in root.attrs['layer_names']] == [ln for ln, _ in layers]
    assert f.attrs['backend'] == b'tensorflow' if not full_model else f['model_weights'].attrs['backend'] == b'tensorflow'
    assert root['conv2d']['conv2d']['kernel:0'].shape == (3, 3, 4, 8)


def test_many_layers_span_symbol_nodes_and_btree_levels(tmp_path):
    rng = np.random.default_rng(1)
    layers = [(f'dense_{i}', [(f'dense_{i}/kernel:0', rng.standard_normal((2, 3)).astype(np.float32))]) for i in range(300)]
    p = kh.write_weights_h5(str(tmp_path / 'many.h5'), layers)
    named = kh.read_named_weights(p)
    assert len(named) == 300
    for (n, got), (ln, ws) in zip(named, layers):
        assert n == f'{ln}/{ws[0][0]}' and np.array_equal(got, ws[0][1])


def test_split_layer_name_attributes(tmp_path):
    """Keras writes layer_names0, layer_names1, ... when the list exceeds the 64 KB attribute limit."""
    rng = np.random.default_rng(2)
    layers = _layers(rng, 2)
    w = kh._Writer()
    links = {}
    for ln, ws in layers:
        sub = kh._build_tree(w, dict(ws))
        links[ln] = w.group(sub, {'weight_names': np.array([n.encode() for n, _ in ws], 'S') if ws else np.zeros((0,), 'S1')})[0]
    names = [ln.encode() for ln, _ in layers]
    root = w.group(links, {'layer_names0': np.array(names[:2], 'S'), 'layer_names1': np.array(names[2:], 'S')})
    data = w.finish(root)
    named = kh.read_named_weights(data)
    assert [n for n, _ in named] == [f'{ln}/{wn}' for ln, ws in layers for wn, _ in ws]


def test_keras3_layout_and_archive(tmp_path):
    rng = np.random.default_rng(3)
    layers = [(ln, ws) for ln, ws in _layers(rng, 12) if ws]   # conv2d..conv2d_11: natural order != alphabetical
    p = kh.write_weights_keras3(str(tmp_path / 'm.weights.h5'), layers)
    named = kh.read_named_weights(p)
    convs = [n for n, _ in named if n.startswith('conv2d')]
    assert convs[:6] == ['conv2d/vars/0', 'conv2d/vars/1', 'conv2d_1/vars/0', 'conv2d_1/vars/1', 'conv2d_2/vars/0', 'conv2d_2/vars/1']
    by = dict(named)
    for ln, ws in layers:
        for i, (_, a) in enumerate(ws):
            assert np.array_equal(by[f'{ln}/vars/{i}'], a)
    z = str(tmp_path / 'm.keras')
    with zipfile.ZipFile(z, 'w') as zf:
        zf.write(p, 'model.weights.h5')
        zf.writestr('config.json', '{}')
    assert len(kh.read_named_weights(z)) == len(named)


def test_not_hdf5_and_truncated(tmp_path):
    bad = tmp_path / 'x.h5'
    bad.write_bytes(b'not an hdf5 file at all' * 10)
    with pytest.raises(ValueError):
        kh.H5File(str(bad))


class _FakeModel:
    """The attributes keras_h5 uses of a UNetModel (no CUDA library needed for this test)."""

    def __init__(self, names, arrays):
        self.weight_names, self._w = names, arrays
        self.weight_shapes = [a.shape for a in arrays]

    def get_weights(self):
        return self._w


def _unet_like(rng):
    names, arrays = [], []
    def conv(p, cin, cout, k=3):
        names.extend([f'{p}/kernel', f'{p}/bias']); arrays.extend([rng.standard_normal((k, k, cin, cout)).astype(np.float32), rng.standard_normal(cout).astype(np.float32)])
    def bn(p, c):
        for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            names.append(f'{p}/{leaf}'); arrays.append(rng.standard_normal(c).astype(np.float32))
    conv('encoder_0/conv0', 6, 8); bn('encoder_0/bn0', 8); conv('encoder_0/conv1', 8, 8); bn('encoder_0/bn1', 8)
    conv('center/conv0', 8, 16); bn('center/bn0', 16)
    names.extend(['decoder_0/up/kernel', 'decoder_0/up/bias']); arrays.extend([rng.standard_normal((2, 2, 8, 16)).astype(np.float32), rng.standard_normal(8).astype(np.float32)])
    bn('decoder_0/bn_cat', 16); conv('decoder_0/conv0', 16, 8); bn('decoder_0/bn0', 8)
    conv('head', 8, 1, 1)
    return _FakeModel(names, arrays)


def test_model_weight_grouping_round_trip(tmp_path):
    """save_weights('.h5') groups the flat list into Keras-style layers; both file flavours load back in order."""
    m = _unet_like(np.random.default_rng(4))
    groups = kh.keras_layer_groups(m)
    assert [g for g, _ in groups] == ['conv2d', 'batch_normalization', 'conv2d_1', 'batch_normalization_1', 'conv2d_2',
                                      'batch_normalization_2', 'conv2d_transpose', 'batch_normalization_3', 'conv2d_3',
                                      'batch_normalization_4', 'conv2d_4']
    p = kh.write_weights_h5(str(tmp_path / 'unet.hdf5'), groups)
    got = kh.read_weights(p, m)
    assert len(got) == len(m.get_weights()) and all(np.array_equal(a, b) for a, b in zip(got, m.get_weights()))
    p3 = kh.write_weights_keras3(str(tmp_path / 'unet.weights.h5'), groups)
    got3 = kh.read_weights(p3, m)   # file order is alphabetical by layer; matched back by kind + index
    assert all(np.array_equal(a, b) for a, b in zip(got3, m.get_weights()))
    short = _FakeModel(m.weight_names[:-2], m.get_weights()[:-2])
    with pytest.raises(ValueError):
        kh.read_weights(p, short)


def test_keras3_nested_reference_blocks_are_matched_by_structure(tmp_path):
    """ADVICE: a Keras 3 file saved from the reference's own get_unet_model has nested encoder_i/encoder/cbaN/
    {conv_layer,bn_layer} groups (utils/model_tools.py:174-286) next to the auto-named functional decoder; the file
    order (alphabetical) is not get_weights() order, so tensors are matched by structure -- or the load fails."""
    m = _unet_like(np.random.default_rng(5))
    w = dict(zip(m.weight_names, m.get_weights()))

    def take(prefix, leaves):
        return [(f'{prefix}/{l}', w[f'{prefix}/{l}']) for l in leaves]
    cv, bnl = ['kernel', 'bias'], ['gamma', 'beta', 'moving_mean', 'moving_variance']
    layers = [('encoder_0/encoder/cba1/conv_layer', take('encoder_0/conv0', cv)), ('encoder_0/encoder/cba1/bn_layer', take('encoder_0/bn0', bnl)),
              ('encoder_0/encoder/cba2/conv_layer', take('encoder_0/conv1', cv)), ('encoder_0/encoder/cba2/bn_layer', take('encoder_0/bn1', bnl)),
              ('conv_block/cba1/conv_layer', take('center/conv0', cv)), ('conv_block/cba1/bn_layer', take('center/bn0', bnl)),
              ('conv2d_transpose', take('decoder_0/up', cv)), ('batch_normalization', take('decoder_0/bn_cat', bnl)),
              ('conv2d', take('decoder_0/conv0', cv)), ('batch_normalization_1', take('decoder_0/bn0', bnl)),
              ('probs', take('head', cv))]
    p = kh.write_weights_keras3(str(tmp_path / 'ref.weights.h5'), layers)
    got = kh.read_weights(p, m)
    assert len(got) == len(m.get_weights()) and all(np.array_equal(a, b) for a, b in zip(got, m.get_weights()))
    # an unknown layer group is an error, not a positional guess
    bad = kh.write_weights_keras3(str(tmp_path / 'bad.weights.h5'), layers[:-1] + [('my_custom_head', take('head', cv))])
    with pytest.raises(ValueError, match='cannot find the Keras layer|not matched'):
        kh.read_weights(bad, m)


def test_new_style_container_hand_assembled():
    """Superblock v2, 'OHDR' v2 object headers, compact groups made of link messages, v2 dataspaces, a v3
    attribute holding a variable-length UTF-8 string in a global heap -- assembled byte by byte here,
    independent of keras_h5._Writer (which only emits the old-style container)."""
    import struct
    buf = bytearray(b'\0' * 48)

    def alloc(b):
        while len(buf) % 8:
            buf.append(0)
        a = len(buf)
        buf.extend(b)
        return a

    def ohdr(msgs):
        body = b''.join(struct.pack('<BHB', t, len(m), 0) + m for t, m in msgs)
        return alloc(b'OHDR' + bytes([2, 0x01]) + struct.pack('<H', len(body)) + body + b'\0\0\0\0')  # flags 1: 2-byte chunk size

    def link(name, addr):
        n = name.encode()
        return (0x06, bytes([1, 0x00, len(n)]) + n + struct.pack('<Q', addr))

    f32 = struct.pack('<BBBBI', 0x11, 0x20, 31, 0, 4) + struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
    data = np.arange(6, dtype=np.float32).reshape(2, 3)
    daddr = alloc(data.tobytes())
    dset = ohdr([(0x01, bytes([2, 2, 0, 1]) + struct.pack('<QQ', 2, 3)), (0x03, f32),
                 (0x08, struct.pack('<BBQQ', 3, 1, daddr, data.nbytes))])
    text = 'größe'.encode('utf-8')
    gcol_body = struct.pack('<HHIQ', 1, 1, 0, len(text)) + text + b'\0' * (-len(text) % 8)
    gcol = alloc(b'GCOL' + bytes([1, 0, 0, 0]) + struct.pack('<Q', 16 + len(gcol_body) + 16) + gcol_body + struct.pack('<HHIQ', 0, 0, 0, 0))
    vlen = struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + struct.pack('<BBBBI', 0x13, 0x00, 0, 0, 1)
    name = b'note\0'
    attr = (bytes([3, 0]) + struct.pack('<HHH', len(name), len(vlen), 4) + bytes([1]) + name + vlen + bytes([2, 0, 0, 0]) +
            struct.pack('<IQI', len(text), gcol, 1))
    vars_g = ohdr([link('0', dset)])
    layer = ohdr([link('vars', vars_g)])
    layers = ohdr([link('dense', layer)])
    root = ohdr([link('layers', layers), (0x0C, attr)])
    sb = kh.SIGNATURE + bytes([2, 8, 8, 0]) + struct.pack('<QQQQ', 0, kh.UNDEF, len(buf), root) + b'\0\0\0\0'
    buf[0:len(sb)] = sb
    f = kh.H5File(bytes(buf))
    assert f.attrs['note'].decode('utf-8') == 'größe'
    d = f['layers/dense/vars/0']
    assert d.is_dataset and d.shape == (2, 3) and np.array_equal(d.read(), data)
    assert [n for n, _ in kh.read_named_weights(bytes(buf))] == ['dense/vars/0']


def test_chunked_gzip_shuffle_dataset_hand_assembled():
    """A chunked dataset (v1 B-tree of chunks, filter pipeline shuffle + deflate, edge chunks that overhang the
    dataspace) assembled byte by byte: what ``h5py.create_dataset(..., compression='gzip', shuffle=True)`` writes."""
    import struct
    import zlib
    data = np.arange(5 * 7, dtype=np.float32).reshape(5, 7) * 0.5
    cdims = (4, 4)
    buf = bytearray(b'\0' * 96)

    def alloc(b):
        while len(buf) % 8:
            buf.append(0)
        a = len(buf)
        buf.extend(b)
        return a

    # chunks: shuffle (byte transpose) then deflate, like the HDF5 filter pipeline applies them on write
    keys = []
    for oy in range(0, 5, 4):
        for ox in range(0, 7, 4):
            chunk = np.zeros(cdims, np.float32)
            part = data[oy:oy + 4, ox:ox + 4]
            chunk[:part.shape[0], :part.shape[1]] = part
            shuffled = np.frombuffer(chunk.tobytes(), np.uint8).reshape(-1, 4).T.tobytes()
            comp = zlib.compress(shuffled)
            keys.append((len(comp), (oy, ox), alloc(comp)))
    # one leaf B-tree node of type 1: key = chunk size, filter mask, offsets (rank + 1), then child address
    node = b'TREE' + struct.pack('<BBHQQ', 1, 0, len(keys), kh.UNDEF, kh.UNDEF)
    for size, (oy, ox), addr in keys:
        node += struct.pack('<IIQQQ', size, 0, oy, ox, 0) + struct.pack('<Q', addr)
    node += struct.pack('<IIQQQ', 0, 0, 8, 8, 0)  # final key
    btree = alloc(node)

    def msg(t, body, flags=0):
        body = body + b'\0' * (-len(body) % 8)
        return struct.pack('<HHB3x', t, len(body), flags) + body

    f32 = struct.pack('<BBBBI', 0x11, 0x20, 31, 0, 4) + struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
    space = struct.pack('<BBB5x', 1, 2, 0) + struct.pack('<QQ', 5, 7)
    layout = struct.pack('<BBB', 3, 2, 3) + struct.pack('<Q', btree) + struct.pack('<III', 4, 4, 4)
    # filter pipeline v1: shuffle (id 2, one client value = element size), deflate (id 1, level)
    def filt(fid, name, cvals):
        nm = name + b'\0' * (-len(name) % 8)
        b = struct.pack('<HHHH', fid, len(nm), 1, len(cvals)) + nm + b''.join(struct.pack('<I', v) for v in cvals)
        return b + (b'\0' * 4 if len(cvals) % 2 else b'')
    pipeline = struct.pack('<BB6x', 1, 2) + filt(2, b'shuffle\0', [4]) + filt(1, b'deflate\0', [4])
    msgs = [msg(0x01, space), msg(0x03, f32, 1), msg(0x08, layout), msg(0x0B, pipeline)]
    body = b''.join(msgs)
    dset = alloc(struct.pack('<BxHII4x', 1, len(msgs), 1, len(body)) + body)
    w = kh._Writer()
    w.buf = buf                      # reuse the old-style group writer for the root group around the raw dataset
    root = w.group({'kernel': dset})
    f = kh.H5File(w.finish(root))
    d = f['kernel']
    assert d.shape == (5, 7) and d.layout[0] == 'chunked' and [fid for fid, _ in d.filters] == [2, 1]
    assert np.array_equal(d.read(), data)
