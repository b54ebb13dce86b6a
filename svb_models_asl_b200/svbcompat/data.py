"""
Volumetric data model: mirror of ``svb.DataModel`` as the reference drives it
(``/root/reference/scripts/gen_test_data.py:10,37-38``, ``aslnn.py:191-192``) and
of the attributes the plugins read (``aslrest.py:110-114,124,142,433-456``):
``n_nodes, n_tpts, shape, mask_vol, mask_flattened, data_flattened,
is_volumetric, node_labels, _get_data()``.

Voxel order is C-order over (x, y, z) of the mask (z fastest), because the
reference masks a ``[X,Y,Z,T]`` array with a boolean volume (``aslrest.py:443``).
Surface / hybrid (toblerone) projection is out of scope: ``is_volumetric`` is
always True.
"""
import numpy as np

from . import nifti
from .utils import LogBase


class DataModel(LogBase):
    is_volumetric = True
    is_hybrid = False

    def __init__(self, data, mask=None, **kwargs):
        LogBase.__init__(self)
        self.nii, vol = self._get_data(data)
        while vol.ndim < 4:
            vol = vol[np.newaxis, ...]
        self.data_vol = vol
        self.shape = list(vol.shape[:3])
        self.n_tpts = int(vol.shape[3])
        flat = vol.reshape(-1, self.n_tpts)
        if mask is not None:
            _nii, mask_vol = self._get_data(mask)
            self.mask_vol = np.asarray(mask_vol)
            while self.mask_vol.ndim < 3:
                self.mask_vol = self.mask_vol[np.newaxis, ...]
            if list(self.mask_vol.shape) != self.shape:
                raise ValueError("Mask shape %s does not match data shape %s"
                                 % (list(self.mask_vol.shape), self.shape))
        else:
            self.mask_vol = np.ones(self.shape, dtype=np.int32)
        self.mask_flattened = self.mask_vol.reshape(-1) > 0
        self.data_flattened = np.ascontiguousarray(flat[self.mask_flattened], dtype=np.float32)
        self.n_unmasked_voxels = self.n_nodes = int(self.data_flattened.shape[0])
        self.node_labels = [(slice(0, self.n_nodes), "GM")]
        if self.nii is not None:
            self.affine = self.nii.affine
        else:
            self.affine = np.eye(4)

    def _get_data(self, data):
        """-> (NiftiImage or None, ndarray).  Accepts a filename, array or scalar."""
        if isinstance(data, str):
            img = nifti.load(data)
            return img, np.asarray(img.data)
        if isinstance(data, nifti.NiftiImage):
            return data, np.asarray(data.data)
        return None, np.asarray(data)

    def voxel_coords(self):
        """Integer (x,y,z) of every masked voxel, in voxel order -> [W,3] int32"""
        idx = np.nonzero(self.mask_flattened)[0]
        X, Y, Z = self.shape
        return np.stack([idx // (Y * Z), (idx // Z) % Y, idx % Z], axis=1).astype(np.int32)

    def neighbour_table(self):
        """6-connected neighbours inside the mask -> [W,6] int32 voxel indices (-x,+x,-y,+y,-z,+z), -1 = none.
        The Laplacian structure of the spatial ("M") prior."""
        coords = self.voxel_coords()
        lut = -np.ones(self.shape, dtype=np.int64)
        lut[coords[:, 0], coords[:, 1], coords[:, 2]] = np.arange(len(coords))
        out = -np.ones((len(coords), 6), dtype=np.int32)
        k = 0
        for axis in range(3):
            for step in (-1, 1):
                c = coords.astype(np.int64).copy()
                c[:, axis] += step
                ok = (c[:, axis] >= 0) & (c[:, axis] < self.shape[axis])
                c[~ok] = 0
                out[:, k] = np.where(ok, lut[c[:, 0], c[:, 1], c[:, 2]], -1)
                k += 1
        return out

    def nifti_image(self, values):
        """Put per-voxel values [W] or [W,N] back in the volume -> NiftiImage"""
        values = np.asarray(values)
        tail = list(values.shape[1:])
        vol = np.zeros([int(np.prod(self.shape))] + tail, dtype=np.float32)
        vol[self.mask_flattened] = values
        return nifti.NiftiImage(vol.reshape(self.shape + tail), self.affine)

    # Surface-mode hooks the reference calls only when not volumetric (aslrest.py:447-454)
    def voxels_to_nodes_ts(self, t, **_kw):
        return t

    def nodes_to_voxels_ts(self, t, **_kw):
        return t

    def uncache_tensors(self):
        pass
