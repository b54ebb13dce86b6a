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

    @classmethod
    def header(cls, shape, n_tpts, mask=None):
        """A data model that knows the volume's geometry but holds no samples (`data_flattened` is None): for
        callers whose data already live on the device (bench.py's 10 M-voxel volume) and only need the plugin's
        option handling, time points and neighbour structure."""
        self = cls.__new__(cls)
        LogBase.__init__(self)
        self.nii, self.data_vol = None, None
        self.shape = [int(x) for x in shape]
        self.n_tpts = int(n_tpts)
        self.mask_vol = np.ones(self.shape, dtype=np.int32) if mask is None else np.asarray(mask)
        self.mask_flattened = self.mask_vol.reshape(-1) > 0
        self.data_flattened = None
        self.n_unmasked_voxels = self.n_nodes = int(self.mask_flattened.sum())
        self.node_labels = [(slice(0, self.n_nodes), "GM")]
        self.affine = np.eye(4)
        return self

    def voxel_coords(self, lo=None, hi=None):
        """Integer (x,y,z) of the masked voxels [lo, hi) (default all), in voxel order -> [n,3] int32"""
        if getattr(self, "_vox_index", None) is None:
            self._vox_index = np.nonzero(self.mask_flattened)[0]
        idx = self._vox_index[slice(lo, hi)]
        X, Y, Z = self.shape
        return np.stack([idx // (Y * Z), (idx // Z) % Y, idx % Z], axis=1).astype(np.int32)

    def neighbour_table(self, lo=None, hi=None):
        """6-connected neighbours inside the mask of the voxels [lo, hi) (default all) -> [n,6] int32 GLOBAL voxel
        indices (-x,+x,-y,+y,-z,+z), -1 = none.  The Laplacian structure of the spatial ("M") prior (SURVEY
        Appendix A.5); a rank of a sharded fit asks for its own range only."""
        coords = self.voxel_coords(lo, hi)
        if getattr(self, "_vox_lut", None) is None:
            lut = -np.ones(int(np.prod(self.shape)), dtype=np.int32)
            lut[self.mask_flattened] = np.arange(self.n_nodes, dtype=np.int32)
            self._vox_lut = lut.reshape(self.shape)
        lut = self._vox_lut
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
