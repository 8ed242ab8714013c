"""Minimal NRRD (Nearly Raw Raster Data, NRRD0004, raw little-endian encoding) writer / reader for label and image
volumes -- the container pred_fold writes through SimpleITK (src/models/predict_model.py:174-186,
sitk.GetImageFromArray + SetSpacing + WriteImage) and evaluate_cv reads back (evaluate_cv.py:400, sitk.ReadImage).
SimpleITK is not installed in this image; this covers exactly the case pred_fold's `orig_given == False` branch produces
(axis-aligned volume, spacing only).  Host-side file I/O, no compute."""
from __future__ import annotations

import numpy as np

_TYPES = {'uint8': 'uint8', 'int8': 'int8', 'uint16': 'uint16', 'int16': 'int16', 'uint32': 'uint32', 'int32': 'int32',
          'float32': 'float', 'float64': 'double'}
_RTYPES = {'uint8': np.uint8, 'uchar': np.uint8, 'unsigned char': np.uint8, 'int8': np.int8, 'uint16': np.uint16,
           'ushort': np.uint16, 'int16': np.int16, 'short': np.int16, 'uint32': np.uint32, 'int32': np.int32,
           'int': np.int32, 'float': np.float32, 'double': np.float64}


def write_nrrd(path: str, volume_zyx: np.ndarray, spacing_xyz=(1.0, 1.0, 1.0)) -> None:
    """volume [Z, Y, X] (numpy order, as handed to sitk.GetImageFromArray) -> NRRD with sizes X Y Z."""
    v = np.ascontiguousarray(volume_zyx)
    if v.ndim != 3 or v.dtype.name not in _TYPES:
        raise ValueError('write_nrrd: need a 3-D volume of %s, got %s %s' % (sorted(_TYPES), v.dtype, v.shape))
    sx, sy, sz = (float(s) for s in spacing_xyz)
    hdr = ['NRRD0004', 'type: %s' % _TYPES[v.dtype.name], 'dimension: 3', 'space: left-posterior-superior',
           'sizes: %d %d %d' % (v.shape[2], v.shape[1], v.shape[0]),
           'space directions: (%r,0,0) (0,%r,0) (0,0,%r)' % (sx, sy, sz), 'kinds: domain domain domain',
           'endian: little', 'encoding: raw', 'space origin: (0,0,0)', '', '']
    with open(path, 'wb') as f:
        f.write('\n'.join(hdr).encode('ascii'))
        f.write(v.astype(v.dtype.newbyteorder('<'), copy=False).tobytes())


def read_nrrd(path: str):
    """-> (volume [Z, Y, X], spacing (x, y, z)); raw encoding only."""
    with open(path, 'rb') as f:
        blob = f.read()
    end = blob.index(b'\n\n')
    fields = {}
    for line in blob[:end].decode('ascii', 'replace').split('\n')[1:]:
        if ':' in line and not line.startswith('#'):
            k, val = line.split(':', 1)
            fields[k.strip()] = val.strip().lstrip('=').strip()
    if fields.get('encoding', 'raw') != 'raw':
        raise NotImplementedError('read_nrrd: encoding %r (only raw is implemented)' % fields.get('encoding'))
    sizes = [int(s) for s in fields['sizes'].split()]
    dt = np.dtype(_RTYPES[fields['type']]).newbyteorder('<' if fields.get('endian', 'little') == 'little' else '>')
    vol = np.frombuffer(blob, dtype=dt, offset=end + 2, count=int(np.prod(sizes))).reshape(sizes[::-1])
    spacing = [1.0] * len(sizes)
    if 'space directions' in fields:
        vecs = [v for v in fields['space directions'].replace('none', '').split(')') if '(' in v]
        spacing = [float(np.linalg.norm([float(c) for c in v.split('(')[1].split(',')])) for v in vecs]
    elif 'spacings' in fields:
        spacing = [float(s) for s in fields['spacings'].split()]
    return vol.astype(dt.newbyteorder('=')), tuple(spacing)
