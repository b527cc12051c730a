"""The file formats either side of the predict path in the GEE workflow (Appendix C),
without TensorFlow / rasterio:

* ``*.tfrecord.gz`` patch files written by ``Export.image.toCloudStorage(fileFormat='TFRecord')`` and read
  by ``make_pred_dataset`` (``utils/prediction_tools.py:159-226``): whole-file GZIP, TFRecord framing
  (``uint64 len | masked crc32c(len) | data | masked crc32c(data)``), one ``tf.train.Example`` per patch with
  one ``FloatList`` of (kernel+buffer)^2 values per band;
* prediction TFRecords (``write_tfrecord_predictions``, ``:375-445``): uncompressed, ``b1..bC`` FloatLists of
  the cropped patch;
* the mixer JSON (``:317-329, :448-455``);
* single-file GeoTIFF output (``write_geotiff_prediction(s)``, ``:447-536``): baseline striped TIFF
  (BigTIFF above 4 GB) with the ModelTransformation / GeoKey tags rasterio would write for ``transform=`` and
  ``crs='EPSG:xxxx'``.

Everything here is host-side byte shuffling (numpy + zlib); the arithmetic of the path stays on the GPU.
"""
from __future__ import annotations

import gzip
import json
import os
import struct

import numpy as np

# ------------------------------------------------------------------------------------------ crc32c
_CRC_TABLE = None


def _crc_table():
    """Byte-at-a-time table of the reflected Castagnoli polynomial."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        t = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (poly if c & 1 else 0)
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli).  Patch records are ~3.5 MB: long messages are split into 4096 lanes whose
    byte-serial CRCs advance together as one numpy recurrence and are then stitched with the GF(2)
    "append n zero bytes" operator (the zlib crc32_combine construction)."""
    mv = memoryview(data)
    if len(mv) >= 4096:
        return _crc32c_blocks(mv, 0xFFFFFFFF) ^ 0xFFFFFFFF
    return _crc32c_raw(mv, 0xFFFFFFFF) ^ 0xFFFFFFFF


def _gf2_times(mat, vec):
    out = 0
    i = 0
    while vec:
        if vec & 1:
            out ^= mat[i]
        vec >>= 1
        i += 1
    return out


def _gf2_square(mat):
    return [_gf2_times(mat, mat[i]) for i in range(32)]


_ZERO_OPS = {}


def _zeros_operator(nbytes):
    """32x32 GF(2) matrix advancing a (reflected) CRC-32C register over `nbytes` zero bytes."""
    if nbytes in _ZERO_OPS:
        return _ZERO_OPS[nbytes]
    odd = [0x82F63B78] + [1 << (i - 1) for i in range(1, 32)]  # one zero bit
    m = odd
    for _ in range(3):  # -> one zero byte (8 bits)
        m = _gf2_square(m)
    result, power, n = None, m, nbytes
    while n:
        if n & 1:
            result = power if result is None else [_gf2_times(power, result[i]) for i in range(32)]
        n >>= 1
        if n:
            power = _gf2_square(power)
    _ZERO_OPS[nbytes] = result
    return result


def _crc32c_raw(mv, crc):
    t = _crc_table()
    for b in mv:
        crc = int(t[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc


def _crc32c_blocks(mv, crc):
    """Split the message into K equal lanes, run the K byte-serial CRCs as one numpy recurrence (vectorised
    over lanes), then stitch the lane CRCs with the zero-extension operator."""
    t = _crc_table()
    n = len(mv)
    K = 4096
    L = n // K
    body = np.frombuffer(mv[:K * L], dtype=np.uint8).reshape(K, L)
    state = np.zeros(K, dtype=np.uint32)
    for j in range(L):
        state = t[(state ^ body[:, j]) & 0xFF] ^ (state >> 8)
    # lane k's raw register started from 0; the true CRC carries `crc` in through the first lane and every
    # lane's result must be advanced over the bytes that follow it: combine left to right
    op = _zeros_operator(L)
    acc = crc
    for k in range(K):
        acc = _gf2_times(op, acc) ^ int(state[k])
    return _crc32c_raw(mv[K * L:], acc)


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ TFRecord framing
def _open(path, mode, compression):
    if compression is None:
        compression = 'GZIP' if str(path).endswith('.gz') else ''
    return gzip.open(path, mode) if compression.upper() == 'GZIP' else open(path, mode)


def read_tfrecords(path, compression=None, verify=False):
    """Yield the raw record payloads of one TFRecord file (``compression`` 'GZIP', '' or None = by suffix)."""
    with _open(path, 'rb', compression) as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError(f'{path}: truncated record header')
            (n,), (hcrc,) = struct.unpack('<Q', head[:8]), struct.unpack('<I', head[8:])
            if verify and masked_crc32c(head[:8]) != hcrc:
                raise ValueError(f'{path}: corrupt record length')
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise ValueError(f'{path}: truncated record')
            if verify and masked_crc32c(data) != struct.unpack('<I', tail)[0]:
                raise ValueError(f'{path}: corrupt record payload')
            yield data


def write_tfrecords(path, records, compression=None):
    with _open(path, 'wb', compression) as f:
        for data in records:
            head = struct.pack('<Q', len(data))
            f.write(head + struct.pack('<I', masked_crc32c(head)) + data + struct.pack('<I', masked_crc32c(data)))
    return path


# ------------------------------------------------------------------------------------------ tf.train.Example
def _varint(buf, p):
    v, s = 0, 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << s
        if b < 0x80:
            return v, p
        s += 7


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _fields(buf):
    """(field number, wire type, value) of one protobuf message; length-delimited values as memoryviews."""
    p, n = 0, len(buf)
    while p < n:
        key, p = _varint(buf, p)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _varint(buf, p)
        elif wt == 2:
            ln, p = _varint(buf, p)
            v = buf[p:p + ln]
            p += ln
        elif wt == 5:
            v = buf[p:p + 4]
            p += 4
        elif wt == 1:
            v = buf[p:p + 8]
            p += 8
        else:
            raise ValueError(f'unsupported protobuf wire type {wt}')
        yield fn, wt, v


def _parse_feature(buf):
    for fn, wt, v in _fields(buf):
        if fn == 2:  # FloatList { repeated float value = 1 [packed] }
            chunks = []
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    chunks.append(np.frombuffer(v2, dtype='<f4'))
            return np.concatenate(chunks) if len(chunks) != 1 else chunks[0]
        if fn == 3:  # Int64List
            vals = []
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    q = 0
                    while q < len(v2):
                        x, q = _varint(v2, q)
                        vals.append(x - (1 << 64) if x >> 63 else x)
                elif f2 == 1:
                    vals.append(v2 - (1 << 64) if v2 >> 63 else v2)
            return np.array(vals, dtype=np.int64)
        if fn == 1:  # BytesList
            return [bytes(v2) for f2, w2, v2 in _fields(v) if f2 == 1]
    return np.zeros(0, np.float32)


def parse_example(record):
    """``tf.io.parse_single_example`` without a schema: {feature name: float32 / int64 array or [bytes]}."""
    buf = memoryview(record)
    out = {}
    for fn, wt, features in _fields(buf):
        if fn != 1:
            continue
        for f2, w2, entry in _fields(features):  # map<string, Feature> entries
            if f2 != 1:
                continue
            key, val = None, None
            for f3, w3, v3 in _fields(entry):
                if f3 == 1:
                    key = bytes(v3).decode('utf-8')
                elif f3 == 2:
                    val = _parse_feature(v3)
            if key is not None:
                out[key] = val
    return out


def _ld(fn, payload):
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def build_example(features):
    """``tf.train.Example(features=Features(feature={name: Feature(float_list=FloatList(value=...))}))``
    ``.SerializeToString()``; int64 arrays become Int64Lists, bytes / lists of bytes BytesLists."""
    entries = b''
    for name, val in features.items():
        if isinstance(val, (bytes, str)):
            val = [val]
        if isinstance(val, list) and val and isinstance(val[0], (bytes, str)):
            items = b''.join(_ld(1, v.encode() if isinstance(v, str) else v) for v in val)
            feat = _ld(1, items)
        else:
            arr = np.asarray(val)
            if arr.dtype.kind in 'iu':
                packed = b''.join(_enc_varint(int(x) & 0xFFFFFFFFFFFFFFFF) for x in arr.ravel())
                feat = _ld(3, _ld(1, packed))
            else:
                feat = _ld(2, _ld(1, np.ascontiguousarray(arr, dtype='<f4').tobytes()))
        entries += _ld(1, _ld(1, name.encode('utf-8')) + _ld(2, feat))
    return _ld(1, entries)


# ------------------------------------------------------------------------------------------ mixer
def read_mixer(path):
    """The ``*mixer.json`` sidecar of a GEE patch export (``utils/prediction_tools.py:317-329, :448-455``)."""
    with open(path) as f:
        m = json.load(f)
    for k in ('patchesPerRow', 'totalPatches'):
        if k not in m:
            raise ValueError(f'{path}: mixer has no {k}')
    return m


# ------------------------------------------------------------------------------------------ patch datasets
def iter_patches(file_list, features, kernel_shape=(256, 256), kernel_buffer=(128, 128), one_hot=None, derived=None):
    """Yield one (h, w, C) float32 patch per Example of the (lexicographically sorted, ``:175``) files:
    bands stacked in ``features`` order and transposed to HWC (``:195-204``); ``one_hot`` {name: depth}
    appends one-hot planes (``:198-200, :213-215``); ``derived`` callables get the feature dict and return an
    (h, w) band appended after the stack (the ``**kwargs`` hooks of ``:208-211``)."""
    h, w = kernel_shape[0] + kernel_buffer[0], kernel_shape[1] + kernel_buffer[1]
    one_hot = dict(one_hot or {})
    for path in sorted(file_list):
        for rec in read_tfrecords(path):
            ex = parse_example(rec)
            dic = {}
            for k in features:
                if k not in ex:
                    raise KeyError(f'{path}: feature {k!r} missing (have {sorted(ex)})')
                v = np.asarray(ex[k], dtype=np.float32)
                if v.size != h * w:
                    raise ValueError(f'{path}: feature {k!r} has {v.size} values, expected {h}x{w}')
                dic[k] = v.reshape(h, w)
            bands = np.stack([dic[k] for k in features if k not in one_hot], axis=-1)
            extra = [np.asarray(fxn(dic), np.float32)[..., None] for fxn in (derived or {}).values()]
            hot = [_one_hot(dic[k].astype(np.uint8), depth) for k, depth in one_hot.items()]
            yield bands, extra, hot


def _one_hot(idx, depth):
    """``tf.one_hot(tf.cast(x, tf.uint8), depth)`` (``utils/prediction_tools.py:198-200``): a category outside
    [0, depth) yields an all-zero vector instead of an error."""
    idx = np.asarray(idx).astype(np.int64)
    out = np.zeros(idx.shape + (int(depth),), dtype=np.float32)
    ok = (idx >= 0) & (idx < depth)
    np.put_along_axis(out, np.where(ok, idx, 0)[..., None], ok[..., None].astype(np.float32), axis=-1)
    return out


def write_patch_tfrecords(path, patches, features, compression='GZIP'):
    """Write (N, h, w, C) patches the way a GEE TFRecord export lays them out (tests, synthetic fixtures)."""
    patches = np.asarray(patches, dtype=np.float32)
    recs = (build_example({k: p[..., i].ravel() for i, k in enumerate(features)}) for p in patches)
    return write_tfrecords(path, recs, compression)


def write_prediction_tfrecords(predictions, out_image_file, kernel_shape=(256, 256), kernel_buffer=(128, 128)):
    """``write_tfrecord_predictions`` (``utils/prediction_tools.py:375-445``) minus the predict call: crop each
    (h, w, C) prediction to ``[y_buffer:y_size, x_buffer:x_size]`` and write features ``b1..bC``."""
    if isinstance(predictions, list):
        parts = [np.asarray(p) if np.asarray(p).ndim == 4 else np.asarray(p)[..., None] for p in predictions]
        predictions = np.concatenate([p.astype(np.float32) for p in parts], axis=3)
    predictions = np.asarray(predictions)
    nch = predictions.shape[-1]
    x_buffer, y_buffer = int(kernel_buffer[0] / 2), int(kernel_buffer[1] / 2)
    x_size, y_size = x_buffer + kernel_shape[1], y_buffer + kernel_shape[0]

    def recs():
        for prediction in predictions:
            patch = prediction[y_buffer:y_size, x_buffer:x_size, :]
            yield build_example({f'b{i + 1}': patch[:, :, i].ravel() for i in range(nch)})

    return write_tfrecords(out_image_file, recs(), compression='')


# ------------------------------------------------------------------------------------------ GeoTIFF
def _epsg(crs):
    if crs is None:
        return None
    if isinstance(crs, int):
        return crs
    s = str(crs).strip()
    if s.upper().startswith('EPSG:'):
        return int(s.split(':')[1])
    return None


_TIFF_TYPES = {np.dtype('uint8'): (1, 8), np.dtype('uint16'): (1, 16), np.dtype('int16'): (2, 16),
               np.dtype('uint32'): (1, 32), np.dtype('int32'): (2, 32), np.dtype('float32'): (3, 32),
               np.dtype('float64'): (3, 64)}


def write_geotiff(path, image, transform=None, crs=None, rows_per_strip=None, bigtiff=None):
    """Band-interleaved-by-pixel striped GeoTIFF of an (H, W) or (H, W, C) array.  ``transform`` = the six
    affine coefficients (a, b, c, d, e, f) of ``rio.Affine`` / the mixer's ``doubleMatrix``
    (x = a*col + b*row + c, y = d*col + e*row + f); ``crs`` = 'EPSG:n'; ``bigtiff`` forces (True) or forbids
    (False) the 64-bit container, default: by size."""
    img = np.asarray(image)
    if img.ndim == 2:
        img = img[..., None]
    if img.dtype not in _TIFF_TYPES:
        img = img.astype(np.float32)
    img = np.ascontiguousarray(img.astype(img.dtype.newbyteorder('<'), copy=False))
    H, W, Cn = img.shape
    fmt, bits = _TIFF_TYPES[np.dtype(img.dtype.name)]
    row_bytes = W * Cn * img.dtype.itemsize
    if rows_per_strip is None:
        rows_per_strip = max(1, min(H, (1 << 20) // max(row_bytes, 1)))
    nstrips = (H + rows_per_strip - 1) // rows_per_strip
    big = (H * row_bytes + 16 * nstrips + 4096 >= (1 << 32) - (1 << 20)) if bigtiff is None else bool(bigtiff)

    # tag -> (type, values); types: 3 SHORT, 4 LONG, 12 DOUBLE, 16 LONG8
    off_t = 16 if big else 4
    tags = {
        256: (4, [W]), 257: (4, [H]), 258: (3, [bits] * Cn), 259: (3, [1]),
        262: (3, [1]), 273: (off_t, None), 277: (3, [Cn]), 278: (4, [rows_per_strip]),
        279: (off_t, [min(rows_per_strip, H - s * rows_per_strip) * row_bytes for s in range(nstrips)]),
        284: (3, [1]), 339: (3, [fmt] * Cn),
    }
    if Cn > 1:
        tags[338] = (3, [0] * (Cn - 1))  # extra samples: unspecified
    if transform is not None:
        a, b, c, d, e, f = [float(v) for v in transform]
        if b == 0.0 and d == 0.0 and a > 0.0 and e < 0.0:  # north-up: pixel scale + tiepoint, like GDAL
            tags[33550] = (12, [a, -e, 0.0])                      # ModelPixelScale
            tags[33922] = (12, [0.0, 0.0, 0.0, c, f, 0.0])         # ModelTiepoint: raster (0,0) -> (c, f)
        else:
            tags[34264] = (12, [a, b, 0.0, c, d, e, 0.0, f, 0, 0, 0, 0, 0, 0, 0, 1.0])  # ModelTransformation
    epsg = _epsg(crs)
    if epsg is not None:
        geographic = epsg in (4326, 4269, 4267, 4258)
        keys = [(1024, 0, 1, 2 if geographic else 1), (1025, 0, 1, 1),
                (2048 if geographic else 3072, 0, 1, epsg)]
        tags[34735] = (3, [1, 1, 0, len(keys)] + [v for k in keys for v in k])  # GeoKeyDirectory

    sizes = {3: 2, 4: 4, 12: 8, 16: 8}
    codes = {3: 'H', 4: 'I', 12: 'd', 16: 'Q'}
    entry = 20 if big else 12
    inline = 8 if big else 4
    header = 16 if big else 8
    ntags = len(tags)
    ifd_size = (8 if big else 2) + ntags * entry + (8 if big else 4)
    # layout: header | IFD | out-of-line tag values | pixel data
    pos = header + ifd_size
    extra = {}
    for t in sorted(tags):
        ty, vals = tags[t]
        count = nstrips if vals is None else len(vals)
        nbytes = count * sizes[ty]
        if nbytes > inline:
            pos += pos & 1
            extra[t] = pos
            pos += nbytes
    data0 = (pos + 15) & ~15
    strip_offsets = [data0 + s * rows_per_strip * row_bytes for s in range(nstrips)]
    tags[273] = (off_t, strip_offsets)

    out = bytearray()
    if big:
        out += struct.pack('<2sHHHQ', b'II', 43, 8, 0, header)
        out += struct.pack('<Q', ntags)
    else:
        out += struct.pack('<2sHI', b'II', 42, header)
        out += struct.pack('<H', ntags)
    blobs = {}
    for t in sorted(tags):
        ty, vals = tags[t]
        raw = struct.pack('<' + codes[ty] * len(vals), *vals)
        if big:
            out += struct.pack('<HHQ', t, ty, len(vals))
        else:
            out += struct.pack('<HHI', t, ty, len(vals))
        if len(raw) <= inline:
            out += raw + b'\0' * (inline - len(raw))
        else:
            out += struct.pack('<Q' if big else '<I', extra[t])
            blobs[extra[t]] = raw
    out += struct.pack('<Q' if big else '<I', 0)
    for off in sorted(blobs):
        out += b'\0' * (off - len(out))
        out += blobs[off]
    out += b'\0' * (data0 - len(out))
    with open(path, 'wb') as fh:
        fh.write(out)
        fh.write(img.tobytes())
    return path


def geotiff_path(out_dir, base):
    return os.path.join(out_dir, f'{base}.tif')
