"""Minimal pure-Python HDF5 reader (and writer) for Keras model files.

The reference loads its networks with ``keras.models.load_model(path)``
(inaSpeechSegmenter/segmenter.py:129-131) from ``keras_*_cnn.hdf5`` release
assets; h5py / TensorFlow are not available on the target boxes, so this module
parses the subset of HDF5 that h5py writes for such files with its default
settings: superblock v0/v1, version-1 object headers (with continuation
blocks), old-style groups (v1 B-tree + symbol-table nodes + local heap),
contiguous and compact datasets of IEEE floats / integers, and compact
attributes holding fixed-length or variable-length (global heap) strings.
Anything outside that subset raises ``NotImplementedError`` naming the feature
(dense attribute storage, chunked/compressed datasets, new-style groups).

``load_keras_hdf5(path) -> (model_config dict, {'<layer>/<var>': ndarray})`` is
what ``models.load_model_file`` uses.  ``write_keras_hdf5`` emits the same
subset (one symbol-table node per group) so tests and ``tools/`` can produce
``.hdf5`` models without h5py; the reader is additionally checked against the
reference's own h5py-written fixture ``media/test.h5``.
"""
import json
import struct

import numpy as np

SIG = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


# =============================================================================== reader
class H5File:
    def __init__(self, path):
        with open(path, 'rb') as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != SIG:
            raise ValueError('%s is not an HDF5 file' % path)
        ver = b[8]
        if ver not in (0, 1):
            raise NotImplementedError('HDF5 superblock version %d (file written with libver="latest"?)' % ver)
        self.so, self.sl = b[13], b[14]
        if (self.so, self.sl) != (8, 8):
            raise NotImplementedError('HDF5 offset/length sizes %d/%d' % (self.so, self.sl))
        pos = 24 if ver == 0 else 28
        self.base, _, _, _ = struct.unpack_from('<QQQQ', b, pos)
        pos += 32
        # root group symbol table entry
        _, self.root_addr, cache, _ = struct.unpack_from('<QQII', b, pos)
        self.root = self._object(self.root_addr)

    # ---- low level -----------------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        ver = b[addr]
        if b[addr:addr + 4] == b'OHDR':
            raise NotImplementedError('version-2 object headers (file written with libver="latest")')
        if ver != 1:
            raise NotImplementedError('object header version %d' % ver)
        nmsg, _, hsize = struct.unpack_from('<HIi', b, addr + 2)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            start, size = blocks.pop(0)
            p = start
            while p + 8 <= start + size and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from('<HHB', b, p)
                body = p + 8
                if mtype == 0x10:
                    off, ln = struct.unpack_from('<QQ', b, body)
                    blocks.append((off, ln))
                out.append((mtype, body, msize, mflags))
                p = body + msize
        return out

    def _heap_string(self, heap_addr, off):
        b = self.buf
        assert b[heap_addr:heap_addr + 4] == b'HEAP'
        data_addr = struct.unpack_from('<Q', b, heap_addr + 24)[0]
        s = data_addr + off
        return b[s:b.index(b'\0', s)].decode()

    def _group_entries(self, btree, heap):
        b = self.buf
        out = {}

        def walk(node):
            assert b[node:node + 4] == b'TREE', 'bad B-tree node'
            level, used = struct.unpack_from('<BH', b, node + 5)
            p = node + 24
            for i in range(used):
                child = struct.unpack_from('<Q', b, p + 8)[0]
                p += 16
                if level > 0:
                    walk(child)
                else:
                    assert b[child:child + 4] == b'SNOD'
                    n = struct.unpack_from('<H', b, child + 6)[0]
                    q = child + 8
                    for _ in range(n):
                        name_off, obj = struct.unpack_from('<QQ', b, q)
                        out[self._heap_string(heap, name_off)] = obj
                        q += 40
        walk(btree)
        return out

    def _datatype(self, p):
        b = self.buf
        cls, ver = b[p] & 0x0F, b[p] >> 4
        bits0 = b[p + 1]
        size = struct.unpack_from('<I', b, p + 4)[0]
        if cls == 0:
            return ('int', '<>'[bits0 & 1] + ('i' if bits0 & 8 else 'u') + str(size), size)
        if cls == 1:
            return ('float', '<>'[bits0 & 1] + 'f' + str(size), size)
        if cls == 3:
            return ('string', 'S%d' % size, size)
        if cls == 9:
            is_str = (bits0 & 0x0F) == 1
            base = self._datatype(p + 8)
            return ('vlen_str' if is_str else 'vlen', base, 16)
        raise NotImplementedError('HDF5 datatype class %d' % cls)

    def _dataspace(self, p):
        b = self.buf
        ver, rank = b[p], b[p + 1]
        q = p + (8 if ver == 1 else 4)
        return tuple(struct.unpack_from('<%dQ' % rank, b, q)) if rank else ()

    def _global_heap_object(self, gcol, index):
        b = self.buf
        assert b[gcol:gcol + 4] == b'GCOL'
        size = struct.unpack_from('<Q', b, gcol + 8)[0]
        p = gcol + 16
        while p < gcol + size:
            idx, _, _, osize = struct.unpack_from('<HHIQ', b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + osize]
            p += 16 + (osize + 7) // 8 * 8
        raise KeyError('global heap object %d' % index)

    def _decode(self, dt, shape, raw):
        kind, code, size = dt
        n = int(np.prod(shape)) if shape else 1
        if kind in ('int', 'float'):
            return np.frombuffer(raw[:n * size], dtype=code).reshape(shape).copy()
        if kind == 'string':
            arr = [raw[i * size:(i + 1) * size].split(b'\0')[0].decode('utf-8', 'replace') for i in range(n)]
            return arr[0] if not shape else np.array(arr, dtype=object).reshape(shape)
        if kind == 'vlen_str':
            arr = []
            for i in range(n):
                ln, addr, idx = struct.unpack_from('<IQI', raw, i * 16)
                arr.append(self._global_heap_object(addr, idx)[:ln].decode('utf-8', 'replace') if ln else '')
            return arr[0] if not shape else np.array(arr, dtype=object).reshape(shape)
        raise NotImplementedError('decode %s' % kind)

    def _attribute(self, p):
        b = self.buf
        ver = b[p]
        if ver == 1:
            nsz, dsz, ssz = struct.unpack_from('<HHH', b, p + 2)
            q = p + 8
            name = b[q:q + nsz].split(b'\0')[0].decode(); q += (nsz + 7) // 8 * 8
            dt = self._datatype(q); q += (dsz + 7) // 8 * 8
            shape = self._dataspace(q); q += (ssz + 7) // 8 * 8
        elif ver in (2, 3):
            nsz, dsz, ssz = struct.unpack_from('<HHH', b, p + 2)
            q = p + (8 if ver == 2 else 9)
            name = b[q:q + nsz].split(b'\0')[0].decode(); q += nsz
            dt = self._datatype(q); q += dsz
            shape = self._dataspace(q); q += ssz
        else:
            raise NotImplementedError('attribute message version %d' % ver)
        n = int(np.prod(shape)) if shape else 1
        return name, self._decode(dt, shape, b[q:q + n * dt[2]])

    def _object(self, addr):
        """-> dict(kind='group'|'dataset', attrs={}, children={name: addr} | reader)"""
        obj = {'addr': addr, 'attrs': {}, 'kind': None}
        dt = shape = layout = None
        for mtype, body, msize, _ in self._messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from('<QQ', self.buf, body)
                obj['kind'] = 'group'
                obj['children'] = self._group_entries(btree, heap)
            elif mtype in (0x02, 0x06):
                raise NotImplementedError('new-style HDF5 groups (file written with libver="latest")')
            elif mtype == 0x0C:
                k, v = self._attribute(body)
                obj['attrs'][k] = v
            elif mtype == 0x15:
                fheap = struct.unpack_from('<Q', self.buf, body + (6 if self.buf[body + 1] & 1 else 2))[0]
                if fheap != UNDEF:
                    raise NotImplementedError('dense attribute storage (an attribute > 64 KB, e.g. a very large model_config)')
            elif mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dt = self._datatype(body)
            elif mtype == 0x08:
                layout = body
        if layout is not None and dt is not None:
            obj['kind'] = 'dataset'
            obj['shape'], obj['dtype'], obj['layout'] = shape, dt, layout
        return obj

    # ---- public ------------------------------------------------------------------
    def read_dataset(self, obj):
        b, p = self.buf, obj['layout']
        ver, cls = b[p], b[p + 1]
        if ver != 3:
            raise NotImplementedError('data layout message version %d' % ver)
        n = int(np.prod(obj['shape'])) if obj['shape'] else 1
        nbytes = n * obj['dtype'][2]
        if cls == 0:
            size = struct.unpack_from('<H', b, p + 2)[0]
            raw = b[p + 4:p + 4 + size]
        elif cls == 1:
            addr, size = struct.unpack_from('<QQ', b, p + 2)
            raw = b'' if addr == UNDEF else b[self.base + addr:self.base + addr + nbytes]
            if addr == UNDEF:
                raw = bytes(nbytes)
        else:
            raise NotImplementedError('chunked / compressed HDF5 datasets (re-save the model without compression)')
        return self._decode(obj['dtype'], obj['shape'], raw)

    def walk(self, obj=None, prefix=''):
        """Yields (path, object) depth-first."""
        obj = obj or self.root
        yield prefix or '/', obj
        if obj['kind'] == 'group':
            for name in sorted(obj['children']):
                child = self._object(obj['children'][name])
                yield from self.walk(child, prefix + '/' + name)

    def get(self, path):
        obj = self.root
        for part in [p for p in path.split('/') if p]:
            obj = self._object(obj['children'][part])
        return obj


def load_keras_hdf5(path):
    """Keras HDF5 model file -> (model_config dict, weights {'<layer>/<var>': float32 array})."""
    f = H5File(path)
    cfg = f.root['attrs'].get('model_config')
    if cfg is None:
        raise ValueError('%s has no model_config attribute (weights-only file?)' % path)
    if isinstance(cfg, bytes):
        cfg = cfg.decode()
    config = json.loads(cfg)
    top = f.get('model_weights') if 'model_weights' in f.root.get('children', {}) else f.root
    weights = {}
    for lname in top['children']:
        layer = f._object(top['children'][lname])
        for path_, obj in f.walk(layer, ''):
            if obj['kind'] == 'dataset':
                var = path_.rsplit('/', 1)[-1]
                var = var[:-2] if var.endswith(':0') else var
                weights['%s/%s' % (lname, var)] = np.asarray(f.read_dataset(obj), dtype=np.float32)
    return config, weights


# =============================================================================== writer
class _Writer:
    """Emits superblock v0 + v1 object headers + one symbol-table node per group."""
    LEAF_K = 512          # entries per symbol-table node = 2K: one SNOD per group is always enough here

    def __init__(self):
        self.buf = bytearray(b'\0' * 2048)          # superblock area, patched at the end

    def _alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _pad8(b):
        return b + b'\0' * ((-len(b)) % 8)

    @staticmethod
    def _dt_float32():
        return struct.pack('<BBBBI', 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dt_string(n):
        return struct.pack('<BBBBI', 0x13, 0x00, 0x00, 0x00, n)

    @staticmethod
    def _dataspace(shape):
        if not shape:
            return struct.pack('<BBBB4x', 1, 0, 0, 0)
        return struct.pack('<BBBB4x', 1, len(shape), 0, 0) + struct.pack('<%dQ' % len(shape), *shape)

    def _attr_msg(self, name, value):
        if isinstance(value, str):
            raw = value.encode() + b'\0'
            dt, ds, data = self._dt_string(len(raw)), self._dataspace(()), raw
        else:                                   # list of strings -> fixed-length string array
            items = [v.encode() for v in value]
            width = max([len(i) for i in items] + [1])
            dt, ds = self._dt_string(width), self._dataspace((len(items),))
            data = b''.join(i.ljust(width, b'\0') for i in items)
        nm = name.encode() + b'\0'
        body = struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(ds)) + self._pad8(nm) + self._pad8(dt) + self._pad8(ds) + data
        return 0x0C, body

    def _header(self, msgs):
        body = b''
        for mtype, data in msgs:
            data = self._pad8(data)
            if len(data) > 0xFFF8:
                raise ValueError('attribute too large for a compact object-header message')
            body += struct.pack('<HHB3x', mtype, len(data), 0) + data
        hdr = struct.pack('<BBHII4x', 1, 0, len(msgs), 1, len(body))
        return self._alloc(hdr + body)

    def dataset(self, arr, attrs=None):
        arr = np.ascontiguousarray(arr, dtype='<f4')
        data_addr = self._alloc(arr.tobytes())
        layout = struct.pack('<BBQQ', 3, 1, data_addr, arr.nbytes)
        msgs = [(0x01, self._dataspace(arr.shape)), (0x03, self._dt_float32()), (0x08, layout)]
        msgs += [self._attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs)

    def group(self, children, attrs=None):
        names = sorted(children)
        heap_data = bytearray(b'\0' * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode() + b'\0'
            while len(heap_data) % 8:
                heap_data.append(0)
        heap_data_addr = self._alloc(bytes(heap_data))
        heap = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), UNDEF, heap_data_addr))
        assert len(names) <= 2 * self.LEAF_K
        snod = b'SNOD' + struct.pack('<BBH', 1, 0, len(names))
        for n in names:
            snod += struct.pack('<QQII16x', offs[n], children[n], 0, 0)
        snod += b'\0' * (40 * (2 * self.LEAF_K - len(names)))
        snod_addr = self._alloc(snod)
        last = offs[names[-1]] if names else 0
        tree = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1 if names else 0, UNDEF, UNDEF) + struct.pack('<QQQ', 0, snod_addr, last)
        tree += b'\0' * (16 * 2 * 16)
        btree = self._alloc(tree)
        msgs = [(0x11, struct.pack('<QQ', btree, heap))] + [self._attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs), btree, heap

    def finish(self, root):
        root_addr, btree, heap = root
        sb = SIG + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, 16, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, root_addr, 1, 0) + struct.pack('<QQ', btree, heap)
        self.buf[:len(sb)] = sb
        return bytes(self.buf)


def write_keras_hdf5(path, config, weights, keras_version='2.3.1', backend='tensorflow'):
    """(model_config dict, {'<layer>/<var>': array}) -> a Keras-style whole-model HDF5 file:
    /model_weights/<layer>/<layer>/<var>:0 datasets, layer_names / weight_names / model_config attributes."""
    w = _Writer()
    layers = {}
    for key, arr in weights.items():
        lname, var = key.split('/', 1)
        layers.setdefault(lname, {})[var] = arr
    cfg_layers = config['config']['layers'] if isinstance(config.get('config'), dict) else config['config']
    order = [l['config']['name'] for l in cfg_layers]
    layer_groups = {}
    for lname in order:
        vars_ = layers.get(lname, {})
        inner = {v + ':0': w.dataset(a) for v, a in vars_.items()}
        names = ['%s/%s:0' % (lname, v) for v in vars_]
        if inner:
            inner_addr = w.group(inner)[0]
            layer_groups[lname] = w.group({lname: inner_addr}, {'weight_names': names})[0]
        else:
            layer_groups[lname] = w.group({}, {'weight_names': names} if names else None)[0]
    mw = w.group(layer_groups, {'layer_names': order, 'backend': backend, 'keras_version': keras_version})[0]
    root = w.group({'model_weights': mw}, {'model_config': json.dumps(config), 'keras_version': keras_version,
                                           'backend': backend})
    with open(path, 'wb') as f:
        f.write(w.finish(root))
