"""
Minimal NIfTI-1 single-file (.nii / .nii.gz) reader and writer.

nibabel, which the reference scripts use for I/O
(``/root/reference/scripts/asl_example.py:14,47-48``,
``gen_test_data.py:8,51-56``), is not available in this image, so the small
subset needed for the example data (``scripts/asldata_diff.nii.gz`` float32
4-D, ``asldata_mask.nii.gz`` int16 3-D) and for writing result maps is
implemented here with ``struct`` + ``gzip``.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
           256: np.int8, 512: np.uint16, 768: np.uint32}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class NiftiImage:
    def __init__(self, data, affine=None, pixdim=None):
        self.data = data
        self.affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        self.pixdim = pixdim

    def get_fdata(self):
        return self.data

    get_data = get_fdata

    @property
    def shape(self):
        return self.data.shape


def _open(fname, mode):
    return gzip.open(fname, mode) if str(fname).endswith(".gz") else open(fname, mode)


def load(fname):
    with _open(fname, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError("%s: too short for a NIfTI-1 header" % fname)
    end = "<"
    if struct.unpack("<i", raw[:4])[0] != 348:
        end = ">"
        if struct.unpack(">i", raw[:4])[0] != 348:
            raise ValueError("%s: not a NIfTI-1 file" % fname)
    if raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError("%s: bad NIfTI-1 magic" % fname)
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, = struct.unpack(end + "h", raw[70:72])
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(end + "3f", raw[108:120])
    sform_code, = struct.unpack(end + "h", raw[254:256])
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %i" % (fname, datatype))
    shape = [int(d) for d in dim[1:1 + dim[0]]]
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape))
    arr = np.frombuffer(raw, dtype=dt, count=n, offset=int(vox_offset)).reshape(shape, order="F")
    arr = arr.astype(dt.newbyteorder("="))
    if slope not in (0.0, 1.0) or (slope != 0.0 and inter != 0.0):
        if slope != 0.0:
            arr = arr.astype(np.float32) * np.float32(slope) + np.float32(inter)
    affine = np.eye(4)
    if sform_code > 0:
        affine[:3, :] = np.array(struct.unpack(end + "12f", raw[280:328])).reshape(3, 4)
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return NiftiImage(arr, affine, pixdim[1:1 + len(shape)])


def save(img_or_array, fname, affine=None):
    if isinstance(img_or_array, NiftiImage):
        data, affine = img_or_array.data, img_or_array.affine
    else:
        data = np.asarray(img_or_array)
    if data.dtype == np.float64 or data.dtype == np.int64 or data.dtype == bool:
        data = data.astype(np.float32)
    if data.dtype not in _CODES:
        raise ValueError("unsupported dtype for NIfTI output: %s" % data.dtype)
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, _CODES[data.dtype])
    struct.pack_into("<h", hdr, 72, data.dtype.itemsize * 8)
    vox = [float(np.linalg.norm(affine[:3, i])) or 1.0 for i in range(3)]
    struct.pack_into("<8f", hdr, 76, 1.0, *vox, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)
    struct.pack_into("<B", hdr, 123, 10)  # xyzt_units: mm + s
    struct.pack_into("<2h", hdr, 252, 0, 2)  # qform_code, sform_code (aligned)
    struct.pack_into("<12f", hdr, 280, *affine[:3, :].reshape(-1))
    hdr[344:348] = b"n+1\0"
    with _open(fname, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\0\0\0\0")
        f.write(np.asfortranarray(data).tobytes(order="F"))
