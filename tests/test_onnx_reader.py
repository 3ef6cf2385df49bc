"""final.onnx reader (inaspeechsegmenter_b200/onnx_reader.py) against a file written by torch's own exporter
from the REAL resnet.py (tests/golden/make_onnx_golden.py), and -- where the asset exists -- the reference's
known-answer test of the production x-vector backend (run_test.py:189-195)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')


def _eval_blob(blob, m, feat_dim, embed_dim, num_blocks, x):
    """Evaluates the iss_resnet_create blob layout with torch-CPU ops (test-side statement of the layout:
    per conv [kh][kw][cin][cout] weights, scale[cout], shift[cout]; blocks conv1, conv2, conv3, (shortcut);
    embedding [in][embed] + bias).  x: [n, feat_dim, T]."""
    pos = [0]

    def take(n):
        v = blob[pos[0]:pos[0] + n]
        pos[0] += n
        return v

    def conv_bn(inp, cin, cout, k, stride, pad):
        w = torch.from_numpy(take(k * k * cin * cout).reshape(k, k, cin, cout).transpose(3, 2, 0, 1).copy())
        scale, shift = torch.from_numpy(take(cout).copy()), torch.from_numpy(take(cout).copy())
        y = F.conv2d(inp, w, stride=stride, padding=pad)
        return y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)

    out = F.relu(conv_bn(x.unsqueeze(1), 1, m, 3, 1, 1))
    inp = m
    for planes, nb, stride in zip((m, 2 * m, 4 * m, 8 * m), num_blocks, (1, 2, 2, 2)):
        for b in range(nb):
            s = stride if b == 0 else 1
            y = F.relu(conv_bn(out, inp, planes, 1, 1, 0))
            y = F.relu(conv_bn(y, planes, planes, 3, s, 1))
            y = conv_bn(y, planes, 4 * planes, 1, 1, 0)
            sc = conv_bn(out, inp, 4 * planes, 1, s, 0) if (s != 1 or inp != 4 * planes) else out
            out = F.relu(y + sc)
            inp = 4 * planes
    mean = out.mean(-1)
    std = torch.sqrt((out * out).mean(-1) - mean ** 2 + 1e-10)
    feat = torch.cat((mean.flatten(1), std.flatten(1)), 1)
    d = feat.shape[1]
    W = torch.from_numpy(take(d * embed_dim).reshape(d, embed_dim).copy())
    bias = torch.from_numpy(take(embed_dim).copy())
    assert pos[0] == blob.size
    return (feat @ W + bias).numpy()


def test_wire_format_primitives():
    from inaspeechsegmenter_b200 import onnx_reader as R
    assert R._varint(bytes([0xAC, 0x02]), 0) == (300, 2)
    assert R._signed((1 << 64) - 1) == -1
    # TensorProto{dims: [2, 2], data_type: FLOAT, name: "w", raw_data: 4 floats}
    raw = np.arange(4, dtype='<f4').tobytes()
    msg = bytes([0x08, 2, 0x08, 2, 0x10, 1, 0x42, 1]) + b'w' + bytes([0x4A, len(raw)]) + raw
    name, arr = R._tensor(memoryview(msg))
    assert name == 'w' and arr.shape == (2, 2) and arr.dtype == np.float32 and arr[1, 1] == 3.0
    with pytest.raises(ValueError):
        R.load_onnx_graph(b'\x08\x01')


def test_reader_on_torch_exported_resnet():
    """Graph written by torch.onnx.export of the real resnet.py (small config, random BN statistics):
    structure is recovered, and the blob evaluates to the module's own output."""
    from inaspeechsegmenter_b200 import onnx_reader as R
    from inaspeechsegmenter_b200.vbx_segmenter import resnet_blob_from_state
    z = np.load(os.path.join(GOLD, 'resnet_small.npz'))
    g = R.load_onnx_graph(os.path.join(GOLD, 'resnet_small.onnx'))
    assert g.inputs == ['input'] and g.outputs == ['output']
    assert sum(nd.op == 'Conv' for nd in g.nodes) == 1 + 5 * 3 + 4          # stem + 5 blocks x 3 + 4 shortcuts
    blob, m, feat_dim, embed_dim, nb = R.resnet_blob_from_onnx(g)
    assert (m, feat_dim, embed_dim, nb) == (4, 16, 8, (2, 1, 1, 1))
    sd = {k[3:]: z[k] for k in z.files if k.startswith('sd:')}
    ref_blob = resnet_blob_from_state(sd, m=4, num_blocks=(2, 1, 1, 1))
    assert blob.size == ref_blob.size
    x, y = torch.from_numpy(z['x']), z['y']
    got = _eval_blob(blob, m, feat_dim, embed_dim, nb, x)
    got_sd = _eval_blob(ref_blob, m, feat_dim, embed_dim, nb, x)
    tol = 2e-5 * np.abs(y).max()
    assert np.abs(got - y).max() <= tol, np.abs(got - y).max()        # ONNX route == the real module
    assert np.abs(got_sd - y).max() <= tol                              # state_dict route too (same layout)


def test_reader_rejects_other_graphs():
    from inaspeechsegmenter_b200 import onnx_reader as R
    g = R.load_onnx_graph(os.path.join(GOLD, 'resnet_small.onnx'))
    g.nodes = [nd for nd in g.nodes if nd.op != 'Gemm']
    with pytest.raises(ValueError):
        R.resnet_blob_from_onnx(g)


@pytest.mark.gpu
def test_vbx_onnx_known_answer(media):
    """The reference's test_vbx_onnx (run_test.py:189-195): media/test.h5 lamartinemelbands -> lamartineonnx at
    4 decimals, with the DEFAULT engine.  Needs the release asset final.onnx (absent off-line => skipped)."""
    from inaspeechsegmenter_b200 import keras_hdf5, models, vbx_segmenter as vb
    path = models.find_model_file('final.onnx')
    if path is None:
        pytest.skip('final.onnx not available (release asset, remote_utils.py:5)')
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    h5 = keras_hdf5.H5File(os.path.join(media, 'test.h5'))
    feats = h5.read_dataset(h5.get('lamartinemelbands'))
    ref = h5.read_dataset(h5.get('lamartineonnx'))
    ext = vb.B200BackendExtractor(onnx_path=path)
    got = ext.get_embedding(np.asarray(feats, dtype=np.float32))
    np.testing.assert_almost_equal(ref, got, decimal=4)
