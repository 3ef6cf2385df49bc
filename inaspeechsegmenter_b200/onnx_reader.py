"""Dependency-free reader for the ResNet x-vector extractor shipped as ``final.onnx``
(the asset the reference's production backend loads: vbx_segmenter.py:249-266,
remote_utils.py:5,13-14).  No ``onnx`` / ``onnxruntime`` / protobuf runtime is
needed: the file is walked with a ~100-line protobuf wire-format decoder, the
graph is matched structurally (stem conv -> bottleneck blocks found from their
residual ``Add`` nodes -> embedding ``Gemm``), and the weights are laid out in the
blob ``iss_resnet_create`` expects -- the same layout ``resnet_blob_from_state``
builds from ``raw_81.pth``.

torch's exporter folds eval-mode BatchNorm into the preceding convolution
(weights scaled per output channel, BN shift as the conv bias); such convs are
stored as (W', scale = 1, shift = bias).  Un-folded ``BatchNormalization`` nodes
are handled too (scale/shift computed from the running statistics).

Field numbers follow onnx.proto3 (ModelProto.graph = 7; GraphProto.node = 1,
.initializer = 5, .input = 11, .output = 12; NodeProto.input = 1, .output = 2,
.name = 3, .op_type = 4, .attribute = 5; AttributeProto.name = 1, .f = 2, .i = 3,
.ints = 8; TensorProto.dims = 1, .data_type = 2, .float_data = 4, .int64_data = 7,
.name = 8, .raw_data = 9).
"""
import struct

import numpy as np


# ----------------------------------------------------------------------------- protobuf wire format
def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError('malformed varint')


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message; length-delimited values are memoryviews."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val = buf[pos:pos + n]; pos += n
        elif wt == 5:
            val = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d (field %d)' % (wt, fn))
        if pos > end:
            raise ValueError('truncated protobuf message')
        yield fn, wt, val


def _packed_varints(val, wt):
    if wt == 0:
        return [val]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(v)
    return out


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


_DTYPES = {1: np.float32, 7: np.int64, 6: np.int32, 11: np.float64, 10: np.float16}


def _tensor(buf):
    dims, dtype, name, raw, floats, int64s = [], 1, '', None, [], []
    for fn, wt, val in _fields(buf):
        if fn == 1:
            dims += [_signed(v) for v in _packed_varints(val, wt)]
        elif fn == 2:
            dtype = val
        elif fn == 8:
            name = bytes(val).decode()
        elif fn == 9:
            raw = val
        elif fn == 4:
            floats.append(np.frombuffer(val, dtype='<f4') if wt == 2 else np.frombuffer(val, dtype='<f4', count=1))
        elif fn == 7:
            int64s += [_signed(v) for v in _packed_varints(val, wt)]
        elif fn == 13 and len(val):
            raise NotImplementedError('ONNX external tensor data is not supported (%s)' % name)
    if dtype not in _DTYPES:
        raise NotImplementedError('ONNX tensor %s: data_type %d not supported' % (name, dtype))
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(_DTYPES[dtype]).newbyteorder('<'))
    elif floats:
        arr = np.concatenate(floats)
    else:
        arr = np.asarray(int64s, dtype=np.int64)
    return name, arr.astype(_DTYPES[dtype], copy=False).reshape(dims)


def _attribute(buf):
    name, out = '', {}
    ints = []
    for fn, wt, val in _fields(buf):
        if fn == 1:
            name = bytes(val).decode()
        elif fn == 2:
            out['f'] = struct.unpack('<f', val)[0]
        elif fn == 3:
            out['i'] = _signed(val)
        elif fn == 8:
            ints += [_signed(v) for v in _packed_varints(val, wt)]
        elif fn == 5:
            out['t'] = _tensor(val)[1]
    if ints:
        out['ints'] = ints
    return name, out


class Node:
    __slots__ = ('op', 'name', 'inputs', 'outputs', 'attrs')

    def __init__(self):
        self.op, self.name, self.inputs, self.outputs, self.attrs = '', '', [], [], {}

    def __repr__(self):
        return '%s(%s -> %s)' % (self.op, ','.join(self.inputs), ','.join(self.outputs))


def _node(buf):
    n = Node()
    for fn, wt, val in _fields(buf):
        if fn == 1:
            n.inputs.append(bytes(val).decode())
        elif fn == 2:
            n.outputs.append(bytes(val).decode())
        elif fn == 3:
            n.name = bytes(val).decode()
        elif fn == 4:
            n.op = bytes(val).decode()
        elif fn == 5:
            k, v = _attribute(val)
            n.attrs[k] = v
    return n


def _value_info_name(buf):
    for fn, wt, val in _fields(buf):
        if fn == 1:
            return bytes(val).decode()
    return ''


class OnnxGraph:
    """nodes (file order), initializers {name: ndarray}, graph inputs / outputs (names)."""

    def __init__(self, nodes, initializers, inputs, outputs):
        self.nodes, self.initializers, self.inputs, self.outputs = nodes, initializers, inputs, outputs
        self.producer = {}
        for nd in nodes:
            for o in nd.outputs:
                self.producer[o] = nd
        for nd in nodes:                                   # Constant nodes behave like initializers
            if nd.op == 'Constant' and 'value' in nd.attrs and 't' in nd.attrs['value']:
                self.initializers.setdefault(nd.outputs[0], nd.attrs['value']['t'])

    def const(self, name):
        """Value of a constant tensor: an initializer, possibly behind Identity nodes (the exporter
        de-duplicates equal initializers that way)."""
        seen = 0
        while name not in self.initializers:
            nd = self.producer.get(name)
            if nd is None or nd.op != 'Identity' or seen > 64:
                raise ValueError('ONNX tensor %s is not a constant' % name)
            name, seen = nd.inputs[0], seen + 1
        return self.initializers[name]


def load_onnx_graph(path_or_bytes):
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data = memoryview(path_or_bytes)
    else:
        with open(path_or_bytes, 'rb') as f:
            data = memoryview(f.read())
    graph = None
    for fn, wt, val in _fields(data):
        if fn == 7 and wt == 2:
            graph = val
    if graph is None:
        raise ValueError('not an ONNX ModelProto: no graph field')
    nodes, inits, inputs, outputs = [], {}, [], []
    for fn, wt, val in _fields(graph):
        if fn == 1:
            nodes.append(_node(val))
        elif fn == 5:
            k, v = _tensor(val)
            inits[k] = v
        elif fn == 11:
            inputs.append(_value_info_name(val))
        elif fn == 12:
            outputs.append(_value_info_name(val))
    inputs = [i for i in inputs if i not in inits]
    return OnnxGraph(nodes, inits, inputs, outputs)


# ----------------------------------------------------------------------------- ResNet structure matching
_PASS = ('Relu',)


class _ConvUnit:
    """One Conv2d + BatchNorm pair of resnet.py as (weight [cout,cin,kh,kw], scale[cout], shift[cout])."""

    def __init__(self, g, conv, bn):
        w = np.asarray(g.const(conv.inputs[1]), dtype=np.float32)
        cout = w.shape[0]
        bias = np.asarray(g.const(conv.inputs[2]), dtype=np.float32) if len(conv.inputs) > 2 and conv.inputs[2] else np.zeros(cout, np.float32)
        if conv.attrs.get('group', {}).get('i', 1) != 1:
            raise NotImplementedError('grouped convolution in the ONNX graph')
        if any(d != 1 for d in conv.attrs.get('dilations', {}).get('ints', [1, 1])):
            raise NotImplementedError('dilated convolution in the ONNX graph')
        self.w = w
        self.stride = tuple(conv.attrs.get('strides', {}).get('ints', [1, 1]))
        self.pads = tuple(conv.attrs.get('pads', {}).get('ints', [0, 0, 0, 0]))
        if bn is None:
            self.scale, self.shift = np.ones(cout, np.float32), bias
        else:
            gamma, beta, mean, var = (np.asarray(g.const(k), dtype=np.float32) for k in bn.inputs[1:5])
            eps = np.float32(bn.attrs.get('epsilon', {}).get('f', 1e-5))
            self.scale = (gamma / np.sqrt(var + eps)).astype(np.float32)
            self.shift = (beta - mean * self.scale + bias * self.scale).astype(np.float32)
        self.input = conv.inputs[0]

    def blob_parts(self):
        return [np.ascontiguousarray(self.w.transpose(2, 3, 1, 0)).ravel(), self.scale, self.shift]


def _unit_ending_at(g, tensor):
    """The conv(+bn) unit whose (post-BN, pre-ReLU) output is `tensor`; None if `tensor` is not produced by one."""
    nd = g.producer.get(tensor)
    bn = None
    if nd is not None and nd.op == 'BatchNormalization':
        bn, nd = nd, g.producer.get(nd.inputs[0])
    if nd is None or nd.op != 'Conv':
        return None
    return _ConvUnit(g, nd, bn)


def _skip_relu(g, tensor):
    nd = g.producer.get(tensor)
    while nd is not None and nd.op in _PASS:
        tensor = nd.inputs[0]
        nd = g.producer.get(tensor)
    return tensor


def resnet_blob_from_onnx(path_or_bytes):
    """(blob float32, m_channels, feat_dim, embed_dim, num_blocks) from a torch-exported ONNX file of
    resnet.py's Bottleneck ResNet (resnet.py:48-135).  Raises ValueError if the graph is not that network."""
    g = path_or_bytes if isinstance(path_or_bytes, OnnxGraph) else load_onnx_graph(path_or_bytes)
    adds = [nd for nd in g.nodes if nd.op == 'Add' and all(i in g.producer for i in nd.inputs)]
    blocks = []                                            # (block_input_tensor, c1, c2, c3, shortcut or None)
    for add in adds:
        chains = []
        for t in add.inputs:
            u3 = _unit_ending_at(g, t)
            chains.append((t, u3))
        main = None
        for idx, (t, u3) in enumerate(chains):
            if u3 is None:
                continue
            u2 = _unit_ending_at(g, _skip_relu(g, u3.input))
            if u2 is None or u2.w.shape[2:] != (3, 3):
                continue
            u1 = _unit_ending_at(g, _skip_relu(g, u2.input))
            if u1 is None:
                continue
            main = (idx, u1, u2, u3)
        if main is None:
            continue                                       # an Add that is not a residual join (e.g. the eps of the std pooling)
        idx, u1, u2, u3 = main
        other_t, other_u = chains[1 - idx]
        block_in = u1.input
        if other_u is not None and other_u.input == block_in:
            sc = other_u
        elif other_t == block_in:
            sc = None
        else:
            raise ValueError('residual Add %s: shortcut branch does not start at the block input' % add.name)
        blocks.append((block_in, u1, u2, u3, sc))
    if not blocks:
        raise ValueError('no bottleneck blocks found in the ONNX graph')
    stem = _unit_ending_at(g, _skip_relu(g, blocks[0][0]))
    if stem is None or stem.w.shape[1] != 1 or stem.w.shape[2:] != (3, 3):
        raise ValueError('stem convolution (1 -> m, 3x3) not found in front of the first block')
    m = stem.w.shape[0]
    parts = stem.blob_parts()
    num_blocks, planes_prev = [], None
    for block_in, u1, u2, u3, sc in blocks:
        planes = u1.w.shape[0]
        if u3.w.shape[0] != 4 * planes or u2.w.shape[:2] != (planes, planes):
            raise ValueError('block with planes=%d is not a resnet.py Bottleneck' % planes)
        if planes != planes_prev:
            num_blocks.append(0)
            planes_prev = planes
            if sc is None:
                raise ValueError('first block of a stage has no shortcut convolution')
        num_blocks[-1] += 1
        for u in (u1, u2, u3) + ((sc,) if sc is not None else ()):
            parts += u.blob_parts()
    if len(num_blocks) != 4 or [blocks[0][1].w.shape[0] * (1 << i) for i in range(4)] != \
            [p for p in dict.fromkeys(b[1].w.shape[0] for b in blocks)]:
        raise ValueError('expected 4 stages with planes m, 2m, 4m, 8m; found stages %r' % (num_blocks,))
    gemm = [nd for nd in g.nodes if nd.op == 'Gemm']
    if len(gemm) != 1:
        raise ValueError('expected exactly one Gemm (embedding) node, found %d' % len(gemm))
    gm = gemm[0]
    W = np.asarray(g.const(gm.inputs[1]), dtype=np.float32)
    if gm.attrs.get('transB', {}).get('i', 0):
        W = W.T                                            # -> [in][embed]
    if gm.attrs.get('transA', {}).get('i', 0) or gm.attrs.get('alpha', {}).get('f', 1.0) != 1.0 or \
            gm.attrs.get('beta', {}).get('f', 1.0) != 1.0:
        raise NotImplementedError('Gemm with transA / alpha / beta')
    b = np.asarray(g.const(gm.inputs[2]), dtype=np.float32) if len(gm.inputs) > 2 else np.zeros(W.shape[1], np.float32)
    parts += [np.ascontiguousarray(W).ravel(), b]
    embed_dim = W.shape[1]
    feat_dim = W.shape[0] // (2 * 8 * m * 4) * 8           # embedding in = (feat_dim / 8) * 16 m * expansion (resnet.py:103)
    blob = np.concatenate([np.asarray(p, dtype=np.float32).ravel() for p in parts])
    return blob, m, feat_dim, embed_dim, tuple(num_blocks)
