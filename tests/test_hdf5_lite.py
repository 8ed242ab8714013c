"""utils/hdf5_lite.py: the reader against a file written by the HDF5 library itself, the writer by round trip, and the
Keras weight-file layout on top (tf.keras hdf5_format: layer_names / weight_names attributes, <layer>/<weight name>)."""
import glob
import os
import struct

import numpy as np
import pytest

from cmr_landmark_detection_b200.utils import hdf5_lite as H


def _library_written_file():
    try:
        import scipy.io
    except Exception:
        return None
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), 'matlab', 'tests', 'data', 'testhdf5_7.4_GLNX86.mat'))
    return hits[0] if hits else None


def test_reader_on_a_library_written_file():
    """MATLAB v7.3 files are HDF5 files (512-byte user block, superblock 0, symbol-table groups, version-1 object
    headers): the one scipy ships holds testdouble = 0 : pi/4 : 2 pi as a 9 x 1 float64 dataset with a string attribute."""
    path = _library_written_file()
    if path is None:
        pytest.skip('scipy test data not installed')
    f = H.File(path)
    assert f.keys() == ['testdouble']
    d = f['testdouble']
    assert not d.is_group and d.shape == (9, 1)
    assert np.allclose(d.read().reshape(-1), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    assert bytes(d.attrs['MATLAB_class']) == b'double'
    with pytest.raises(KeyError):
        f['nope']


def _layers(rng):
    return [('conv2d', [('conv2d/kernel:0', rng.standard_normal((3, 3, 1, 32)).astype(np.float32)),
                        ('conv2d/bias:0', rng.standard_normal(32).astype(np.float32))]),
            ('batch_normalization', [('batch_normalization/%s:0' % n, rng.standard_normal(32).astype(np.float32))
                                     for n in ('gamma', 'beta', 'moving_mean', 'moving_variance')]),
            ('max_pooling2d', []),
            ('unet', [('unet/kernel:0', rng.standard_normal((1, 1, 32, 2)).astype(np.float32)),
                      ('unet/bias:0', np.zeros(2, np.float32))])]


def test_keras_layout_round_trip(tmp_path):
    layers = _layers(np.random.default_rng(0))
    path = str(tmp_path / 'model.h5')
    H.save_keras_weights(path, layers)
    f = H.File(path)
    assert [bytes(x).decode() for x in f.attrs['layer_names']] == [n for n, _ in layers]
    assert bytes(f.attrs['backend']) == b'tensorflow'
    assert sorted(f.keys()) == sorted(n for n, _ in layers)
    assert f['conv2d'].keys() == ['conv2d'] and sorted(f['conv2d/conv2d'].keys()) == ['bias:0', 'kernel:0']
    assert f['max_pooling2d'].keys() == [] and np.asarray(f['max_pooling2d'].attrs['weight_names']).size == 0
    back = H.load_keras_weights(path)
    assert [n for n, _ in back] == [n for n, w in layers if w]          # weightless layers are dropped, as Keras does
    for (n0, w0), (n1, w1) in zip([l for l in layers if l[1]], back):
        assert [a for a, _ in w0] == [a for a, _ in w1]
        for (_, a), (_, b) in zip(w0, w1):
            assert a.dtype == b.dtype == np.float32 and a.shape == b.shape and np.array_equal(a, b)


def test_written_file_structure(tmp_path):
    """The bytes the HDF5 library looks at first: signature, superblock 0 with 8-byte offsets and the library's default
    B-tree geometry (symbol nodes of 2 x 4 entries, tree nodes of 2 x 16 children), end-of-file address, root symbol-table
    entry -> version-1 object header whose first message is the symbol table.  45 members need six symbol nodes under one
    tree node; 300 members need 38 symbol nodes under two level-0 nodes under a level-1 root."""
    w = H.Writer()
    for i in range(45):
        w.dataset('/g/d%02d' % i, np.full((2, 3), i, np.float32))
    for i in range(300):
        w.dataset('/big/x%03d' % i, np.full((1,), i, np.float32))
    w.group('/empty')
    w.attr('/', 'note', b'hello')
    path = str(tmp_path / 'wide.h5')
    w.save(path)
    raw = open(path, 'rb').read()
    assert raw[:8] == H.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert struct.unpack_from('<HH', raw, 16) == (4, 16)
    base, free, eof, drv = struct.unpack_from('<QQQQ', raw, 24)
    assert base == 0 and eof == len(raw) and free == drv == H.UNDEF
    name_off, root_hdr, cache = struct.unpack_from('<QQI', raw, 56)
    assert cache == 1 and raw[root_hdr] == 1 and struct.unpack_from('<H', raw, root_hdr + 16)[0] == 0x0011
    f = H.File(path)
    assert bytes(f.attrs['note']) == b'hello'
    assert f.keys() == ['big', 'empty', 'g'] and f['empty'].keys() == []
    assert f['g'].keys() == ['d%02d' % i for i in range(45)]              # in name order, as the B-tree must be
    assert all(float(f['g/d%02d' % i].read()[1, 2]) == i for i in range(45))
    assert f['big'].keys() == ['x%03d' % i for i in range(300)]
    assert all(float(f['big/x%03d' % i].read()[0]) == i for i in (0, 7, 8, 255, 256, 299))
    # the big group's tree really has two levels, and every node's keys ascend (heap offsets follow name order here)
    bt = struct.unpack_from('<Q', [b for t, b in f['big']._msgs if t == 0x0011][0], 0)[0]
    assert raw[bt:bt + 4] == b'TREE' and raw[bt + 5] == 1 and struct.unpack_from('<H', raw, bt + 6)[0] == 2
    keys = [struct.unpack_from('<Q', raw, bt + 24 + 16 * i)[0] for i in range(3)]
    assert keys[0] == 0 and keys == sorted(keys)


def test_unsupported_files_fail_loudly(tmp_path):
    p = tmp_path / 'x.h5'
    p.write_bytes(b'not an hdf5 file' * 64)
    with pytest.raises(H.Hdf5Error):
        H.File(str(p))
    p.write_bytes(H.SIGNATURE + bytes([2]) + b'\x00' * 200)              # superblock version 2 (libver='latest')
    with pytest.raises(H.Hdf5Error):
        H.File(str(p))


def test_c2_model_sized_weight_file(tmp_path):
    """The whole C2 network's weight table (118 arrays, 8.64 M floats, names and shapes from the C plan) through the
    writer and back: what ModelCheckpoint('model.h5') writes and pred_fold's load_weights reads."""
    import ctypes as C
    from cmr_landmark_detection_b200.runtime import ffi
    L = ffi.lib()
    cfg = ffi.rvip_cfg(H=256, W=256, in_ch=1, classes=2, depth=4, filters=32, batch_norm=1, bn_first=0, use_upsample=1,
                       precision=1, dropout_mid=0.5, bn_momentum=0.99, bn_eps=1e-3)
    h = C.c_void_p()
    ffi.check(L.rvip_create(C.byref(cfg), C.byref(h)))
    rng = np.random.default_rng(3)
    layers, n_conv, n_bn = [], 0, 0
    for i in range(L.rvip_num_tensors(h)):
        nm = C.create_string_buffer(128)
        st, off, nd, dims = C.c_int(), C.c_longlong(), C.c_int(), (C.c_int * 4)()
        ffi.check(L.rvip_tensor_info(h, i, nm, 128, C.byref(st), C.byref(off), C.byref(nd), C.byref(dims)))
        name, shape = nm.value.decode(), tuple(dims[:nd.value])
        leaf = name.rsplit('/', 1)[-1]
        if leaf == 'kernel':
            lname = 'unet' if name.startswith('head/') else ('conv2d' if n_conv == 0 else 'conv2d_%d' % n_conv)
            n_conv += 0 if name.startswith('head/') else 1
            layers.append((lname, []))
        elif leaf == 'gamma':
            lname = 'batch_normalization' if n_bn == 0 else 'batch_normalization_%d' % n_bn
            n_bn += 1
            layers.append((lname, []))
        layers[-1][1].append(('%s/%s:0' % (layers[-1][0], leaf), rng.standard_normal(shape).astype(np.float32)))
    L.rvip_destroy(h)
    assert len(layers) == 41 and sum(len(w) for _, w in layers) == 118
    path = str(tmp_path / 'model.h5')
    H.save_keras_weights(path, layers)
    assert 8641730 * 4 < os.path.getsize(path) < 8641730 * 4 + 400000      # payload + ~2.6 KB of group structure per group
    back = H.load_keras_weights(path)
    assert [n for n, _ in back] == [n for n, _ in layers]
    for (_, w0), (_, w1) in zip(layers, back):
        for (n0, a), (n1, b) in zip(w0, w1):
            assert n0 == n1 and np.array_equal(a, b)


def test_convert_tool_round_trip(tmp_path):
    """tools/keras_h5_convert.py: HDF5 -> .npz -> HDF5 keeps every layer, name and array."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('keras_h5_convert', os.path.join(root, 'tools', 'keras_h5_convert.py'))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    layers = [l for l in _layers(np.random.default_rng(1)) if l[1]]
    a, z, b = str(tmp_path / 'a.h5'), str(tmp_path / 'w.npz'), str(tmp_path / 'b.h5')
    H.save_keras_weights(a, layers)
    tool.to_npz(a, z)
    tool.to_h5(z, b)
    back = H.load_keras_weights(b)
    assert [n for n, _ in back] == [n for n, _ in layers]
    for (_, w0), (_, w1) in zip(layers, back):
        assert [n for n, _ in w0] == [n for n, _ in w1]
        assert all(np.array_equal(x, y) for (_, x), (_, y) in zip(w0, w1))


def test_reader_variable_length_string_attribute(tmp_path):
    """h5py 3 stores str attributes (`backend`, `keras_version` of newer Keras) as variable-length UTF-8 strings: a
    version-3 attribute message whose 16-byte element points into a global heap collection.  Hand-assembled here from the
    format specification (the writer itself only emits fixed-length strings); an attribute of a datatype the reader does
    not know is skipped with a reason instead of making the file unreadable."""
    w = H.Writer()
    w.dataset('/d', np.arange(4, dtype=np.float32))
    path = str(tmp_path / 'v.h5')
    w.save(path)
    raw = bytearray(open(path, 'rb').read())
    # global heap collection: object 1 = b'tensorflow', then the free-space object 0
    text = b'tensorflow'
    gcol_addr = len(raw)
    obj = struct.pack('<HHIQ', 1, 0, 0, len(text)) + text + b'\x00' * (-len(text) % 8)
    body = obj + struct.pack('<HHIQ', 0, 0, 0, 0)
    raw += b'GCOL' + struct.pack('<B3xQ', 1, 16 + len(body)) + body
    # attribute message version 3: vlen string (class 9, type = string, UTF-8) of base type 1-byte string, scalar space v2
    base = struct.pack('<B3BI', 0x13, 0x00, 0, 0, 1)
    vtype = struct.pack('<B3BI', 0x19, 0x01 | (1 << 4), 0x01, 0, 16) + base
    space = struct.pack('<BBBB', 2, 0, 0, 0)
    name = b'backend\x00'
    att = struct.pack('<BBHHHB', 3, 0, len(name), len(vtype), len(space), 1) + name + vtype + space
    att += struct.pack('<IQI', len(text), gcol_addr, 1)
    weird = struct.pack('<BBHHHB', 3, 0, 6, 8, len(space), 0) + b'weird\x00' + struct.pack('<B3BI', 0x16, 0, 0, 0, 4) + space + b'\x00' * 4
    # a new object header for the root group: its symbol-table message copied over, plus the two attributes
    f = H.File(path)
    symtab = [b for t, b in f._msgs if t == 0x0011][0]
    msgs = [H._msg(0x0011, symtab, flags=1), H._msg(0x000C, att), H._msg(0x000C, weird)]
    blob = b''.join(msgs)
    raw += b'\x00' * (-len(raw) % 8)
    hdr_addr = len(raw)
    raw += struct.pack('<BxHII4x', 1, len(msgs), 1, len(blob)) + blob
    struct.pack_into('<Q', raw, 56 + 8, hdr_addr)                       # root symbol-table entry -> the new header
    struct.pack_into('<Q', raw, 24 + 16, len(raw))                      # end-of-file address
    open(path, 'wb').write(bytes(raw))
    g = H.File(path)
    assert bytes(g.attrs['backend']) == text
    assert 'weird' in g.unreadable_attrs and 'weird' not in g.attrs
    assert np.array_equal(g['d'].read(), np.arange(4, dtype=np.float32))


def test_reader_follows_object_header_continuations(tmp_path):
    """A root group whose attributes did not fit the first header chunk: the library appends a continuation message
    (0x0010: address, length) and stores the rest there.  Assembled by hand around a file the writer produced."""
    layers = [l for l in _layers(np.random.default_rng(2)) if l[1]]
    path = str(tmp_path / 'c.h5')
    H.save_keras_weights(path, layers)
    f = H.File(path)
    raw = bytearray(open(path, 'rb').read())
    symtab = [b for t, b in f._msgs if t == 0x0011][0]
    attrs = [H._attribute('layer_names', np.array([n.encode() for n, _ in layers])), H._attribute('backend', b'tensorflow'),
             H._attribute('keras_version', b'2.4.0')]
    raw += b'\x00' * (-len(raw) % 8)
    cont_addr = len(raw)
    cont = attrs[1] + attrs[2] + H._msg(0x0000, b'\x00' * 24)            # two attributes and a NIL filler
    raw += cont
    first = [H._msg(0x0011, symtab, flags=1), attrs[0], H._msg(0x0010, struct.pack('<QQ', cont_addr, len(cont)))]
    blob = b''.join(first)
    hdr_addr = len(raw)
    raw += struct.pack('<BxHII4x', 1, 6, 1, len(blob)) + blob
    struct.pack_into('<Q', raw, 56 + 8, hdr_addr)
    struct.pack_into('<Q', raw, 24 + 16, len(raw))
    open(path, 'wb').write(bytes(raw))
    g = H.File(path)
    assert bytes(g.attrs['keras_version']) == b'2.4.0' and bytes(g.attrs['backend']) == b'tensorflow'
    back = H.load_keras_weights(path)
    assert [n for n, _ in back] == [n for n, _ in layers]
    assert all(np.array_equal(a, b) for (_, w0), (_, w1) in zip(layers, back) for (_, a), (_, b) in zip(w0, w1))


def test_reader_unfiltered_chunked_dataset(tmp_path):
    """create_dataset(..., chunks=...) without compression: data layout class 2 -> version-1 B-tree of raw-data chunks.
    A 5 x 7 float32 array in 2 x 3 chunks (edge chunks stick out of the array), assembled by hand from the specification."""
    arr = np.arange(35, dtype=np.float32).reshape(5, 7)
    w = H.Writer()
    w.dataset('/plain', arr)
    path = str(tmp_path / 'k.h5')
    w.save(path)
    raw = bytearray(open(path, 'rb').read())
    ch = (2, 3)
    keys, kids = [], []
    for i in range(0, 5, ch[0]):
        for j in range(0, 7, ch[1]):
            block = np.zeros(ch, np.float32)
            part = arr[i:i + ch[0], j:j + ch[1]]
            block[:part.shape[0], :part.shape[1]] = part
            raw += b'\x00' * (-len(raw) % 8)
            kids.append(len(raw))
            raw += block.tobytes()
            keys.append(struct.pack('<II3Q', block.nbytes, 0, i, j, 0))
    raw += b'\x00' * (-len(raw) % 8)
    tree_addr = len(raw)
    node = b'TREE' + struct.pack('<BBHQQ', 1, 0, len(kids), H.UNDEF, H.UNDEF)
    for k, c in zip(keys, kids):
        node += k + struct.pack('<Q', c)
    node += struct.pack('<II3Q', 0, 0, 6, 9, 0)                         # final key: one chunk past the end
    raw += node
    layout = struct.pack('<BBBQ3I', 3, 2, 3, tree_addr, ch[0], ch[1], 4)
    msgs = [H._msg(0x0001, H._dataspace(arr.shape)), H._msg(0x0003, H._datatype_f32(), flags=1), H._msg(0x0008, layout)]
    blob = b''.join(msgs)
    raw += b'\x00' * (-len(raw) % 8)
    hdr = len(raw)
    raw += struct.pack('<BxHII4x', 1, len(msgs), 1, len(blob)) + blob
    # point the symbol-table entry of '/plain' at the chunked dataset's header
    f = H.File(path)
    old = f._links()['plain']
    snod = raw.find(b'SNOD')
    assert struct.unpack_from('<Q', raw, snod + 8 + 8)[0] == old
    struct.pack_into('<Q', raw, snod + 8 + 8, hdr)
    struct.pack_into('<Q', raw, 24 + 16, len(raw))
    open(path, 'wb').write(bytes(raw))
    assert np.array_equal(H.File(path)['plain'].read(), arr)


def test_golden_file_is_reproduced_byte_for_byte(tmp_path):
    """tests/golden/keras_weights_tiny.h5 is the file a maintainer can open with h5py (tools/keras_h5_convert.py verify);
    the writer must keep producing exactly those bytes, and the reader must return exactly the arrays of the .npz twin."""
    import importlib.util
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    spec = importlib.util.spec_from_file_location('make_hdf5_golden', os.path.join(here, 'make_hdf5_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = str(tmp_path / 't.h5')
    H.save_keras_weights(path, mod.layers())
    assert open(path, 'rb').read() == open(os.path.join(here, 'keras_weights_tiny.h5'), 'rb').read()
    z = np.load(os.path.join(here, 'keras_weights_tiny.npz'))
    back = H.load_keras_weights(os.path.join(here, 'keras_weights_tiny.h5'))
    assert [n for n, _ in back] == ['conv2d', 'batch_normalization', 'unet']
    for lname, ws in back:
        for wn, arr in ws:
            assert np.array_equal(arr, z['%s|%s' % (lname, wn)])
