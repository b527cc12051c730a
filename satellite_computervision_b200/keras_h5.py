"""Keras-format weight files in, without h5py.

The reference loads weights with ``m.load_weights(path)`` / ``models.load_model(path)``
(``utils/model_tools.py:1155-1162, :1198-1200, :1224-1236``) from files written by
``m.save('UNET256.h5')`` and ``ModelCheckpoint('best_weights.hdf5')``
(``notebooks/UNET_G4G_2019_solar.ipynb:1228-1234, :1277``).  Those are HDF5 containers; h5py is not
available here, so this module carries a small read-only HDF5 parser for the subset libhdf5 writes
by default (superblock v0-v3, v1/v2 object headers, symbol-table and compact-link groups, contiguous /
compact / chunked(+deflate, shuffle) datasets, fixed- and variable-length string attributes) plus a
matching minimal writer used by the tests and by ``UNetModel.save_weights('x.h5')``.

Layouts understood by :func:`read_weights`:

* legacy full model   ``/model_weights/<layer>/<weight_name>`` with ``layer_names`` / ``weight_names`` attrs
* legacy weights only ``/<layer>/<weight_name>`` with ``layer_names`` on the root
* Keras 3             ``/layers/<layer>/vars/<i>`` (``.weights.h5``), nested sub-layers walked in name order

The result is the flat ``model.get_weights()`` list; ``UNetModel.set_weights`` checks count and shapes.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# ------------------------------------------------------------------------------------------ reader
class _Datatype:
    def __init__(self, cls, size, dtype=None, vlen_string=False, base=None, strpad=0):
        self.cls, self.size, self.dtype, self.vlen_string, self.base, self.strpad = cls, size, dtype, vlen_string, base, strpad


class H5Object:
    """A group or a dataset: ``attrs`` dict, ``links`` (name -> address) for groups, ``read()`` for datasets."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.attrs, self.links = {}, {}
        self.shape = self.dtype = self.layout = None
        self.filters = []
        f._parse_object_header(self)

    @property
    def is_dataset(self):
        return self.layout is not None

    def keys(self):
        return list(self.links)

    def __contains__(self, name):
        return name in self.links

    def __getitem__(self, path):
        obj = self
        for part in [p for p in path.split('/') if p]:
            if part not in obj.links:
                raise KeyError(f'{part!r} not found (have {sorted(obj.links)[:8]}...)')
            obj = H5Object(self.f, obj.links[part])
        return obj

    def read(self):
        return self.f._read_dataset(self)


class H5File(H5Object):
    """Read-only view of an HDF5 file held in memory."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, 'rb') as fh:
                self.buf = fh.read()
        base = 0
        while base < len(self.buf) and self.buf[base:base + 8] != SIGNATURE:
            base = 512 if base == 0 else base * 2  # the superblock may sit at 0, 512, 1024, ...
        if self.buf[base:base + 8] != SIGNATURE:
            raise H5Error('not an HDF5 file (signature not found)')
        ver = self.buf[base + 8]
        if ver in (0, 1):
            self.O, self.L = self.buf[base + 13], self.buf[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            self.base = self._u(p, self.O)
            p += 4 * self.O  # base, free-space, EOF, driver-info addresses
            root_addr = self._u(p + self.O, self.O)  # symbol-table entry: link-name offset, object header address
        elif ver in (2, 3):
            self.O, self.L = self.buf[base + 9], self.buf[base + 10]
            p = base + 12
            self.base = self._u(p, self.O)
            root_addr = self._u(p + 3 * self.O, self.O)
        else:
            raise H5Error(f'unsupported superblock version {ver}')
        if self.O != 8 or self.L != 8:
            raise H5Error('only 8-byte offsets/lengths are supported')
        super().__init__(self, root_addr)

    # ---- primitives
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], 'little')

    def _a(self, addr):  # file address -> buffer offset
        return addr + self.base

    # ---- object headers
    def _parse_object_header(self, obj):
        o = self._a(obj.addr)
        if self.buf[o:o + 4] == b'OHDR':
            self._parse_ohdr_v2(obj, o)
            return
        if self.buf[o] != 1:
            raise H5Error(f'unsupported object header version {self.buf[o]} at {obj.addr}')
        nmsg = self._u(o + 2, 2)
        size = self._u(o + 8, 4)
        blocks = [(o + 16, size)]
        seen = 0
        while blocks and seen < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and seen < nmsg:
                mtype, msize, _flags = self._u(p, 2), self._u(p + 2, 2), self.buf[p + 4]
                body = p + 8
                if mtype == 0x10:
                    blocks.append((self._a(self._u(body, 8)), self._u(body + 8, 8)))
                else:
                    self._message(obj, mtype, body, msize)
                p = body + msize
                seen += 1

    def _parse_ohdr_v2(self, obj, o):
        flags = self.buf[o + 5]
        p = o + 6
        if flags & 0x20:
            p += 16  # access, modification, change, birth times
        if flags & 0x10:
            p += 4   # max compact / min dense attributes
        csz = 1 << (flags & 3)
        chunk0 = self._u(p, csz)
        p += csz
        track = bool(flags & 0x04)
        blocks = [(p, chunk0)]
        while blocks:
            p, left = blocks.pop(0)
            end = p + left
            while p + 4 + (2 if track else 0) <= end:
                mtype, msize = self.buf[p], self._u(p + 1, 2)
                body = p + 4 + (2 if track else 0)
                if mtype == 0x10:
                    a, ln = self._a(self._u(body, 8)), self._u(body + 8, 8)
                    blocks.append((a + 4, ln - 8))  # skip 'OCHK', drop the checksum
                elif mtype != 0:
                    self._message(obj, mtype, body, msize)
                p = body + msize

    def _message(self, obj, mtype, p, size):
        if mtype == 0x01:
            obj.shape = self._dataspace(p)
        elif mtype == 0x03:
            obj.dtype = self._datatype(p)[0]
        elif mtype == 0x08:
            obj.layout = self._layout(p)
        elif mtype == 0x0B:
            obj.filters = self._filters(p)
        elif mtype == 0x0C:
            name, val = self._attribute(p)
            obj.attrs[name] = val
        elif mtype == 0x11:
            self._walk_group_btree(obj, self._u(p, 8), self._u(p + 8, 8))
        elif mtype == 0x06:
            name, addr = self._link(p)
            if addr is not None:
                obj.links[name] = addr
        elif mtype == 0x02:
            ver, flags = self.buf[p], self.buf[p + 1]
            q = p + 2 + (8 if flags & 1 else 0)
            if self._u(q, 8) != UNDEF:
                raise H5Error('dense link storage (fractal heap) is not supported; re-save with default h5py settings')

    # ---- message bodies
    def _dataspace(self, p):
        ver, rank = self.buf[p], self.buf[p + 1]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if self.buf[p + 3] == 2:
                return None  # null dataspace
            q = p + 4
        else:
            raise H5Error(f'dataspace version {ver}')
        return tuple(self._u(q + 8 * i, 8) for i in range(rank))

    def _datatype(self, p):
        cls, bits0, bits1 = self.buf[p] & 0x0F, self.buf[p + 1], self.buf[p + 2]
        size = self._u(p + 4, 4)
        end = p + 8
        order = '>' if bits0 & 1 else '<'
        if cls == 0:
            dt = np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
            return _Datatype(cls, size, dt), end + 4
        if cls == 1:
            return _Datatype(cls, size, np.dtype(f'{order}f{size}')), end + 12
        if cls == 3:
            return _Datatype(cls, size, np.dtype(f'S{size}'), strpad=bits0 & 0x0F), end
        if cls == 9:
            base, end2 = self._datatype(end)
            return _Datatype(cls, size, None, vlen_string=(bits0 & 0x0F) == 1, base=base), end2
        raise H5Error(f'unsupported datatype class {cls}')

    def _layout(self, p):
        ver = self.buf[p]
        if ver == 3:
            cls = self.buf[p + 1]
            if cls == 0:
                n = self._u(p + 2, 2)
                return ('compact', p + 4, n)
            if cls == 1:
                return ('contiguous', self._u(p + 2, 8), self._u(p + 10, 8))
            if cls == 2:
                nd = self.buf[p + 2]
                addr = self._u(p + 3, 8)
                dims = tuple(self._u(p + 11 + 4 * i, 4) for i in range(nd))
                return ('chunked', addr, dims)
        elif ver in (1, 2):
            nd, cls = self.buf[p + 1], self.buf[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = self._u(q, 8)
                q += 8
            dims = tuple(self._u(q + 4 * i, 4) for i in range(nd))
            q += 4 * nd
            if cls == 1:
                return ('contiguous', addr, None)
            if cls == 2:
                return ('chunked', addr, dims + (self._u(q, 4),))
            n = self._u(q, 4)
            return ('compact', q + 4, n)
        raise H5Error(f'unsupported data layout (version {ver})')

    def _filters(self, p):
        ver, n = self.buf[p], self.buf[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = self._u(q, 2)
            if ver == 1 or fid >= 256:
                nlen = self._u(q + 2, 2)
                q += 2
            else:
                nlen = 0
            ncli = self._u(q + 4, 2)
            q += 6
            q += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cli = [self._u(q + 4 * i, 4) for i in range(ncli)]
            q += 4 * ncli
            if ver == 1 and ncli % 2:
                q += 4
            out.append((fid, cli))
        return out

    def _attribute(self, p):
        ver = self.buf[p]
        nsz, tsz, ssz = self._u(p + 2, 2), self._u(p + 4, 2), self._u(p + 6, 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = self.buf[q:q + nsz].split(b'\0')[0].decode('utf-8')
        q += pad(nsz)
        dt, _ = self._datatype(q)
        q += pad(tsz)
        shape = self._dataspace(q)
        q += pad(ssz)
        return name, self._decode(dt, shape, q)

    def _decode(self, dt, shape, off):
        shape = () if shape is None else shape
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            if not dt.vlen_string:
                raise H5Error('variable-length sequences are not supported')
            vals = [self._global_heap(self._u(off + 16 * i + 4, 8), self._u(off + 16 * i + 12, 4),
                                      self._u(off + 16 * i, 4)) for i in range(n)]
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(self.buf, dtype=dt.dtype, count=n, offset=off).reshape(shape)
        if dt.cls == 3:
            arr = np.char.rstrip(arr, b'\0') if arr.shape else np.bytes_(bytes(arr).rstrip(b'\0'))
        elif not arr.shape:
            arr = arr[()]
        return arr

    def _global_heap(self, addr, index, length):
        o = self._a(addr)
        if self.buf[o:o + 4] != b'GCOL':
            raise H5Error('bad global heap collection')
        size = self._u(o + 8, 8)
        p, end = o + 16, o + size
        while p + 16 <= end:
            idx, osz = self._u(p, 2), self._u(p + 8, 8)
            if idx == index:
                return self.buf[p + 16:p + 16 + min(osz, length) if length else p + 16 + osz]
            if idx == 0:
                break
            p += 16 + (osz + 7) // 8 * 8
        raise H5Error('global heap object not found')

    def _link(self, p):
        flags = self.buf[p + 1]
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = self.buf[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        lsz = 1 << (flags & 3)
        nlen = self._u(q, lsz)
        q += lsz
        name = self.buf[q:q + nlen].decode('utf-8')
        q += nlen
        return name, (self._u(q, 8) if ltype == 0 else None)

    # ---- old-style groups: v1 B-tree of symbol nodes + local heap of names
    def _walk_group_btree(self, obj, btree, heap):
        h = self._a(heap)
        if self.buf[h:h + 4] != b'HEAP':
            raise H5Error('bad local heap')
        data = self._a(self._u(h + 24, 8))

        def name_at(off):
            e = self.buf.index(b'\0', data + off)
            return self.buf[data + off:e].decode('utf-8')

        def node(addr):
            o = self._a(addr)
            if self.buf[o:o + 4] == b'SNOD':
                for i in range(self._u(o + 6, 2)):
                    e = o + 8 + 40 * i
                    obj.links[name_at(self._u(e, 8))] = self._u(e + 8, 8)
                return
            if self.buf[o:o + 4] != b'TREE' or self.buf[o + 4] != 0:
                raise H5Error('bad group B-tree node')
            used = self._u(o + 6, 2)
            for i in range(used):
                node(self._u(o + 24 + 8 + 16 * i, 8))  # key_i (8) child_i (8) ...

        if btree != UNDEF:
            node(btree)

    # ---- dataset payloads
    def _read_dataset(self, obj):
        if obj.dtype is None or obj.shape is None:
            raise H5Error('not a dataset')
        kind = obj.layout[0]
        n = int(np.prod(obj.shape)) if obj.shape else 1
        if kind == 'contiguous':
            if obj.layout[1] == UNDEF:
                return np.zeros(obj.shape, obj.dtype.dtype)
            return self._decode(obj.dtype, obj.shape, self._a(obj.layout[1]))
        if kind == 'compact':
            return self._decode(obj.dtype, obj.shape, obj.layout[1])
        # chunked: v1 B-tree keyed by chunk offsets
        if obj.dtype.cls not in (0, 1):
            raise H5Error('chunked datasets of this type are not supported')
        cdims = obj.layout[2][:-1]
        out = np.zeros(obj.shape, obj.dtype.dtype)
        rank = len(cdims)

        def node(addr):
            o = self._a(addr)
            if self.buf[o:o + 4] != b'TREE' or self.buf[o + 4] != 1:
                raise H5Error('bad chunk B-tree node')
            level, used = self.buf[o + 5], self._u(o + 6, 2)
            ksz = 8 + 8 * (rank + 1)
            p = o + 24
            for _ in range(used):
                csize, mask = self._u(p, 4), self._u(p + 4, 4)
                offs = tuple(self._u(p + 8 + 8 * d, 8) for d in range(rank))
                child = self._u(p + ksz, 8)
                if level > 0:
                    node(child)
                else:
                    raw = self.buf[self._a(child):self._a(child) + csize]
                    for k, (fid, cli) in reversed(list(enumerate(obj.filters))):
                        if mask & (1 << k):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = cli[0] if cli else obj.dtype.size
                            raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                        else:
                            raise H5Error(f'unsupported filter {fid}')
                    chunk = np.frombuffer(raw, obj.dtype.dtype, count=int(np.prod(cdims))).reshape(cdims)
                    sl = tuple(slice(o_, min(o_ + c, s)) for o_, c, s in zip(offs, cdims, obj.shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
                p += ksz + 8

        if obj.layout[1] != UNDEF:
            node(obj.layout[1])
        return out


# ------------------------------------------------------------------------------------------ keras layouts
def _as_str_list(v):
    if v is None:
        return []
    arr = np.atleast_1d(v)
    return [x.decode('utf-8') if isinstance(x, (bytes, np.bytes_)) else str(x) for x in arr.tolist()]


def _chunked_attr(attrs, name):
    """Keras splits long name lists into name0, name1, ... (``save_attributes_to_hdf5_group``)."""
    if name in attrs:
        return _as_str_list(attrs[name])
    out, i = [], 0
    while f'{name}{i}' in attrs:
        out += _as_str_list(attrs[f'{name}{i}'])
        i += 1
    return out


def _open(path_or_bytes):
    """H5File from a path, bytes, or a Keras 3 ``.keras`` archive (zip holding model.weights.h5)."""
    if isinstance(path_or_bytes, H5File):
        return path_or_bytes
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        return H5File(path_or_bytes)
    with open(path_or_bytes, 'rb') as fh:
        head = fh.read(4)
    if head[:2] == b'PK':
        import zipfile
        with zipfile.ZipFile(path_or_bytes) as z:
            return H5File(z.read('model.weights.h5'))
    return H5File(path_or_bytes)


def _natural(name):
    base, _, suf = name.rpartition('_')
    return (base, int(suf)) if base and suf.isdigit() else (name, -1)


def read_named_weights(path_or_bytes):
    """[(name, ndarray)]: legacy files in ``model.get_weights()`` order (the file's ``layer_names`` /
    ``weight_names`` attrs say so); Keras 3 files per layer in natural name order (conv2d, conv2d_1, ...)."""
    f = _open(path_or_bytes)
    root = f['model_weights'] if 'model_weights' in f else f
    out = []
    if 'layer_names' in root.attrs or 'layer_names0' in root.attrs:  # legacy Keras HDF5
        for layer in _chunked_attr(root.attrs, 'layer_names'):
            g = root[layer]
            for wn in _chunked_attr(g.attrs, 'weight_names'):
                out.append((f'{layer}/{wn}', np.asarray(g[wn].read())))
        return out
    if 'layers' in root or 'vars' in root:  # Keras 3 .weights.h5: layers/<name>/vars/<i>, sub-layers nested
        def walk(g, prefix):
            if 'vars' in g:
                v = g['vars']
                for k in sorted(v.keys(), key=lambda s: int(s) if s.isdigit() else 0):
                    out.append((f'{prefix}/vars/{k}', np.asarray(v[k].read())))
            for k in sorted((k for k in g.keys() if k != 'vars'), key=_natural):
                child = g[k]
                if not child.is_dataset:
                    walk(child, f'{prefix}/{k}' if prefix else k)
        walk(root['layers'] if 'layers' in root else root, '')
        return out
    raise H5Error('no Keras weight layout found (expected layer_names attrs or a layers/ group)')


def _layer_kind(name):
    for kind in ('conv2d_transpose', 'batch_normalization', 'conv2d'):
        if name == kind or (name.startswith(kind + '_') and name[len(kind) + 1:].isdigit()):
            return kind
    return None


def _match_structural(by_layer, model, path):
    """Keras 3 file of the reference's own ``get_unet_model`` (subclassed blocks, ``utils/model_tools.py:174-286``):
    encoder_i / conv_block groups hold nested ``encoder/cba1/{conv_layer,bn_layer}`` sub-layers, the decoder is
    functional (auto-named conv2d_transpose / batch_normalization / conv2d, head ``probs`` / ``logits``).  Returns the
    tensors in the model's ``get_weights()`` order or raises: file order is NOT get_weights() order, so there is no
    positional fallback."""
    paths = list(by_layer)

    def nested(top_pred, tail):
        hits = [p for p in paths if top_pred(p.split('/', 1)[0]) and p.endswith(tail)]
        return hits[0] if len(hits) == 1 else None

    flat = {}
    for p in paths:
        k = _layer_kind(p)
        if k is not None:
            flat.setdefault(k, []).append(p)
    for v in flat.values():
        v.sort(key=_natural)
    heads = [p for p in paths if p in ('probs', 'logits') or p.endswith('probs')]
    ordered, used = [], set()
    prefixes = []
    for name in model.weight_names:
        pre = name.rsplit('/', 1)[0]
        if not prefixes or prefixes[-1] != pre:
            prefixes.append(pre)
    for pre in prefixes:
        block, _, leaf = pre.partition('/')
        hit = None
        if block.startswith('encoder_') or block == 'center':
            j = int(leaf[-1]) + 1
            tail = f'cba{j}/' + ('bn_layer' if leaf.startswith('bn') else 'conv_layer')
            if block == 'center':
                hit = nested(lambda t: t.startswith('conv_block'), tail)
            else:
                i = block.split('_')[1]
                hit = nested(lambda t: t in (f'encoder_{i}', 'encoder_block' if i == '0' else f'encoder_block_{i}'), tail)
        elif pre == 'head':
            hit = heads[0] if len(heads) == 1 else (flat.get('conv2d') or [None])[-1]
            if hit in flat.get('conv2d', []):
                flat['conv2d'].remove(hit)
        else:
            kind = 'conv2d_transpose' if leaf == 'up' else ('batch_normalization' if leaf.startswith('bn') else 'conv2d')
            cands = flat.get(kind, [])
            if kind == 'conv2d' and len(heads) != 1 and len(cands) == 1:
                cands = []  # the last auto-named conv2d is the head
            hit = cands.pop(0) if cands else None
        if hit is None or hit in used:
            raise ValueError(f'{path}: cannot find the Keras layer for {pre!r} among {sorted(paths)[:6]}... '
                             '(expected auto-named functional layers or the reference\'s nested encoder_i/encoder/cbaN groups)')
        used.add(hit)
        ordered += by_layer[hit]
    if len(used) != len(paths):
        raise ValueError(f'{path}: {len(paths) - len(used)} layer group(s) of the file were not matched: '
                         f'{sorted(set(paths) - used)[:6]}')
    return ordered


def read_weights(path, model=None):
    """Flat float32 weight list for ``UNetModel.set_weights``.  Legacy files are already in
    ``get_weights()`` order.  Keras 3 files are keyed by layer name only, so with a ``model`` the layers are
    matched by kind in creation order (conv2d, conv2d_1, ... / batch_normalization... / conv2d_transpose...),
    the way a functional U-Net numbers them, or -- for files written from the reference's own
    subclassed blocks -- by their nested structure; anything else is an error, never a positional guess."""
    named = read_named_weights(path)
    keras3 = bool(named) and '/vars/' in named[0][0]
    if keras3 and model is not None:
        by_layer = {}
        for n, wgt in named:
            by_layer.setdefault(n.split('/vars/')[0], []).append(wgt)
        kinds = {}
        for lname in by_layer:
            k = _layer_kind(lname.rsplit('/', 1)[-1]) if '/' not in lname else None
            if k is not None:
                kinds.setdefault(k, []).append(lname)
        if sum(len(v) for v in kinds.values()) == len(by_layer):
            for v in kinds.values():
                v.sort(key=lambda s: _natural(s.rsplit('/', 1)[-1]))
            ordered = []
            for lname, _ in keras_layer_groups(model):
                k = _layer_kind(lname)
                if not kinds.get(k):
                    raise ValueError(f'{path}: no {k} layer left for {lname}')
                ordered += by_layer[kinds[k].pop(0)]
        else:
            ordered = _match_structural(by_layer, model, path)
        named = [('', wgt) for wgt in ordered]
    ws = [np.ascontiguousarray(w, dtype=np.float32) for _, w in named]
    if model is not None and len(ws) != len(model.weight_shapes):
        raise ValueError(f'{path}: file holds {len(ws)} weight tensors, the model expects {len(model.weight_shapes)}')
    return ws


# ------------------------------------------------------------------------------------------ writer
class _Writer:
    """Minimal HDF5 writer: superblock v0, v1 object headers, symbol-table groups, contiguous datasets,
    fixed-length string / numeric attributes -- the container h5py produces with default settings."""

    LEAF_K, INT_K = 4, 16

    def __init__(self):
        self.buf = bytearray(b'\0' * 96)  # superblock (56 bytes + 40-byte root symbol-table entry)

    def _alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf += b'\0'
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _msg(mtype, body, flags=0):
        body = body + b'\0' * (-len(body) % 8)
        return struct.pack('<HHB3x', mtype, len(body), flags) + body

    @staticmethod
    def _dataspace(shape):
        return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', d) for d in shape)

    @staticmethod
    def _datatype(dt):
        dt = np.dtype(dt)
        if dt.kind == 'f':
            props = {4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023), 2: (0, 16, 10, 5, 0, 10, 15)}[dt.itemsize]
            sign = {4: 31, 8: 63, 2: 15}[dt.itemsize]
            return struct.pack('<BBBBI', 0x11, 0x20, sign, 0, dt.itemsize) + struct.pack('<HHBBBBI', *props)
        if dt.kind in 'iu':
            return struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)
        if dt.kind == 'S':
            return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, dt.itemsize)  # null-padded, ASCII (what h5py writes for numpy 'S')
        raise H5Error(f'cannot write dtype {dt}')

    def _attr(self, name, value):
        arr = np.asarray(value)
        if arr.dtype.kind == 'U':
            arr = np.char.encode(arr, 'utf-8')
        if arr.dtype.kind == 'S' and arr.dtype.itemsize == 0:
            arr = arr.astype('S1')
        nm = name.encode() + b'\0'
        pad = lambda b: b + b'\0' * (-len(b) % 8)
        dt, ds = self._datatype(arr.dtype), self._dataspace(arr.shape)
        body = struct.pack('<BxHHH', 1, len(nm), len(dt), len(ds)) + pad(nm) + pad(dt) + pad(ds) + arr.tobytes()
        if len(body) > 64000:
            raise H5Error(f'attribute {name} too large for a compact object header')
        return self._msg(0x0C, body)

    def _object_header(self, msgs):
        data = b''.join(msgs)
        hdr = struct.pack('<BxHII4x', 1, len(msgs), 1, len(data))
        return self._alloc(hdr + data)

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr)
        addr = self._alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack('<BBQQ', 3, 1, addr, arr.nbytes)
        return self._object_header([self._msg(0x01, self._dataspace(arr.shape)), self._msg(0x03, self._datatype(arr.dtype), 1),
                                    self._msg(0x08, layout)])

    def group(self, links, attrs=None):
        """links: {name: object header address}.  Returns (header address, btree address, heap address)."""
        names = sorted(links)
        heap_data = bytearray(b'\0' * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode() + b'\0'
            heap_data += b'\0' * (-len(heap_data) % 8)
        heap_data += b'\0' * 16
        data_addr = self._alloc(bytes(heap_data))
        heap = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), UNDEF & 0xFFFFFFFFFFFFFFFF, data_addr))
        # symbol nodes of up to 2*LEAF_K entries, one B-tree level (up to 2*INT_K nodes), more levels on demand
        per = 2 * self.LEAF_K
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), per):
            part = names[i:i + per]
            body = b''.join(struct.pack('<QQII16x', offs[n], links[n], 0, 0) for n in part)
            body += b'\0' * (40 * (per - len(part)))
            snods.append(self._alloc(b'SNOD' + struct.pack('<BxH', 1, len(part)) + body))
            keys.append(offs[part[-1]] if part else 0)

        def tree(level, children, ckeys):
            out, okeys = [], [ckeys[0]]
            width = 2 * self.INT_K
            for i in range(0, len(children), width):
                ch, ks = children[i:i + width], ckeys[i:i + width + 1]
                body = struct.pack('<Q', ks[0])
                for c, k in zip(ch, ks[1:]):
                    body += struct.pack('<QQ', c, k)
                body += b'\0' * (16 * (width - len(ch)))
                out.append(self._alloc(b'TREE' + struct.pack('<BBHQQ', 0, level, len(ch), UNDEF, UNDEF) + body))
                okeys.append(ks[-1])
            return out, okeys

        nodes, nkeys, level = snods, keys, 0
        while True:
            nodes, nkeys = tree(level, nodes, nkeys)
            level += 1
            if len(nodes) == 1:
                break
        btree = nodes[0]
        msgs = [self._msg(0x11, struct.pack('<QQ', btree, heap))]
        for k, v in (attrs or {}).items():
            msgs.append(self._attr(k, v))
        return self._object_header(msgs), btree, heap

    def finish(self, root):
        addr, btree, heap = root
        sb = SIGNATURE + struct.pack('<BBBxBBBxHHI', 0, 0, 0, 0, 8, 8, self.LEAF_K, self.INT_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, addr, 1, 0) + struct.pack('<QQ', btree, heap)
        self.buf[0:len(sb)] = sb
        return bytes(self.buf)


def _build_tree(w, entries):
    """entries: {'a/b/c': ndarray}; returns {name: address} of the top-level objects, creating sub-groups."""
    groups, leaves = {}, {}
    for path, arr in entries.items():
        head, _, rest = path.partition('/')
        if rest:
            groups.setdefault(head, {})[rest] = arr
        else:
            leaves[head] = arr
    links = {k: w.dataset(v) for k, v in leaves.items()}
    for g, sub in groups.items():
        links[g] = w.group(_build_tree(w, sub))[0]
    return links


def write_weights_h5(path, layers, full_model=False, root_attrs=None):
    """Write a legacy Keras HDF5 weight file.  ``layers``: [(layer_name, [(weight_name, ndarray), ...])], e.g.
    ``('conv2d', [('conv2d/kernel:0', k), ('conv2d/bias:0', b)])``.  ``full_model=True`` nests everything
    under ``model_weights`` like ``model.save('x.h5')``."""
    w = _Writer()
    links = {}
    for lname, weights in layers:
        sub = _build_tree(w, {wn: np.asarray(a, np.float32) for wn, a in weights})
        wnames = np.array([wn.encode() for wn, _ in weights] or [b''], dtype='S')
        attrs = {'weight_names': wnames} if weights else {'weight_names': np.zeros((0,), 'S1')}
        links[lname] = w.group(sub, attrs)[0]
    lnames = np.array([ln.encode() for ln, _ in layers], dtype='S')
    gattrs = {'layer_names': lnames, 'backend': np.bytes_(b'tensorflow'), 'keras_version': np.bytes_(b'2.11.0')}
    if full_model:
        mw = w.group(links, gattrs)[0]
        attrs = {'keras_version': np.bytes_(b'2.11.0'), 'backend': np.bytes_(b'tensorflow')}
        attrs.update(root_attrs or {})
        root = w.group({'model_weights': mw}, attrs)
    else:
        gattrs.update(root_attrs or {})
        root = w.group(links, gattrs)
    data = w.finish(root)
    with open(path, 'wb') as fh:
        fh.write(data)
    return path


def keras_layer_groups(model):
    """Group a UNetModel's flat weight list the way Keras names a functional U-Net's layers
    (SURVEY Appendix B): conv2d[_k], batch_normalization[_k], conv2d_transpose[_k], head layer last."""
    counters = {}

    def auto(base):
        k = counters.get(base, 0)
        counters[base] = k + 1
        return base if k == 0 else f'{base}_{k}'

    layers, cur_prefix, cur = [], None, None
    leafmap = {'kernel': 'kernel:0', 'bias': 'bias:0', 'gamma': 'gamma:0', 'beta': 'beta:0',
               'moving_mean': 'moving_mean:0', 'moving_variance': 'moving_variance:0'}
    for name, arr in zip(model.weight_names, model.get_weights()):
        prefix, leaf = name.rsplit('/', 1)
        if prefix != cur_prefix:
            kind = 'batch_normalization' if '/bn' in f'/{prefix}' else ('conv2d_transpose' if prefix.endswith('/up') else 'conv2d')
            cur = (auto(kind), [])
            layers.append(cur)
            cur_prefix = prefix
        cur[1].append((f'{cur[0]}/{leafmap[leaf]}', arr))
    return layers


def write_weights_keras3(path, layers):
    """Keras 3 ``.weights.h5`` layout: ``layers/<layer>[/<sub-layer>...]/vars/<i>`` (used by the tests); a layer name
    containing '/' becomes nested groups, the way sub-layers of a subclassed layer are stored."""
    w = _Writer()
    tree = {}
    for lname, weights in layers:
        node = tree
        for part in lname.split('/'):
            node = node.setdefault(part, {})
        node['vars'] = w.group({str(i): w.dataset(np.asarray(a, np.float32)) for i, (_, a) in enumerate(weights)})[0]

    def emit(node):
        return w.group({k: (v if k == 'vars' else emit(v)) for k, v in node.items()})[0]

    root = w.group({'layers': emit(tree), 'vars': w.group({})[0]})
    with open(path, 'wb') as fh:
        fh.write(w.finish(root))
    return path
