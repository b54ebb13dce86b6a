"""
CPU tier: the data formats either side of the hot path - NIfTI-1 files (the reference reads/writes them through
nibabel: scripts/asl_example.py:14,45-48, gen_test_data.py:51-56), the DataModel a plugin is handed
(aslrest.py:110-114,433-456) and the neighbour structure of the spatial prior.
"""
import gzip
import os
import struct

import numpy as np
import pytest

from svb import DataModel
from svb_models_asl_b200.svbcompat import nifti


@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.uint8, np.int32, np.float64])
@pytest.mark.parametrize("shape", [(5, 4, 3), (5, 4, 3, 7), (1, 1, 9, 2)])
def test_nifti_round_trip_keeps_values_shape_and_affine(tmp_path, dtype, shape):
    rng = np.random.default_rng(3)
    data = (rng.normal(0, 50, shape)).astype(dtype)
    affine = np.array([[3.0, 0, 0, -90], [0, 3.0, 0, -126], [0, 0, 5.0, -72], [0, 0, 0, 1]])
    for ext in (".nii", ".nii.gz"):
        path = str(tmp_path / ("vol" + ext))
        nifti.save(nifti.NiftiImage(data, affine), path)
        img = nifti.load(path)
        want = data.astype(np.float32) if dtype == np.float64 else data
        np.testing.assert_array_equal(img.data, want)
        assert img.shape == tuple(shape)
        np.testing.assert_allclose(img.affine, affine)
        assert img.get_fdata() is img.data and img.get_data() is img.data        # the nibabel calls the scripts make


def test_nifti_reads_big_endian_and_scaled_files_and_rejects_garbage(tmp_path):
    data = np.arange(24, dtype=np.int16).reshape(2, 3, 4)
    hdr = bytearray(348)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, 2, 3, 4, 1, 1, 1, 1)
    struct.pack_into(">h", hdr, 70, 4)
    struct.pack_into(">h", hdr, 72, 16)
    struct.pack_into(">8f", hdr, 76, 1.0, 2.0, 2.0, 2.0, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into(">3f", hdr, 108, 352.0, 0.5, 10.0)                # scl_slope 0.5, scl_inter 10
    hdr[344:348] = b"n+1\0"
    path = str(tmp_path / "be.nii.gz")
    with gzip.open(path, "wb") as f:
        f.write(bytes(hdr) + b"\0\0\0\0" + data.astype(">i2").tobytes(order="F"))
    img = nifti.load(path)
    np.testing.assert_allclose(img.data, data * 0.5 + 10.0)
    np.testing.assert_allclose(np.diag(img.affine)[:3], [2.0, 2.0, 2.0])          # no sform: pixdim on the diagonal
    bad = str(tmp_path / "bad.nii")
    open(bad, "wb").write(b"\1" * 400)
    with pytest.raises(ValueError, match="not a NIfTI-1"):
        nifti.load(bad)
    open(bad, "wb").write(b"\1" * 10)
    with pytest.raises(ValueError, match="too short"):
        nifti.load(bad)
    with pytest.raises(ValueError, match="unsupported dtype"):
        nifti.save(np.zeros((2, 2, 2), dtype=np.complex64), str(tmp_path / "c.nii"))


def test_data_model_masks_in_c_order_and_restores_volumes(tmp_path):
    rng = np.random.default_rng(5)
    vol = rng.normal(size=(4, 3, 5, 6)).astype(np.float32)
    mask = (rng.uniform(size=(4, 3, 5)) > 0.4).astype(np.int16)
    dpath, mpath = str(tmp_path / "d.nii.gz"), str(tmp_path / "m.nii.gz")
    nifti.save(vol, dpath)
    nifti.save(mask, mpath)
    dm = DataModel(dpath, mask=mpath)
    assert dm.shape == [4, 3, 5] and dm.n_tpts == 6 and dm.is_volumetric
    assert dm.n_nodes == dm.n_unmasked_voxels == int(mask.sum())
    # voxel order = C order over the mask, z fastest (what aslrest.py:438-443 assumes for the slice timing)
    np.testing.assert_array_equal(dm.data_flattened, vol[mask > 0])
    coords = dm.voxel_coords()
    np.testing.assert_array_equal(coords, np.argwhere(mask > 0))
    # per-voxel results go back where they came from, zeros outside the mask
    img = dm.nifti_image(np.arange(dm.n_nodes, dtype=np.float32))
    assert img.shape == (4, 3, 5)
    np.testing.assert_array_equal(img.data[mask > 0], np.arange(dm.n_nodes))
    assert (img.data[mask == 0] == 0).all()
    assert dm.nifti_image(dm.data_flattened).shape == (4, 3, 5, 6)
    with pytest.raises(ValueError, match="Mask shape"):
        DataModel(vol, mask=np.ones((4, 3, 4)))
    # array inputs of lower rank (gen_test_data.py:38: DataModel(ndarray[n, T]))
    flat = DataModel(np.zeros((7, 6), dtype=np.float32))
    assert flat.n_nodes == 7 and flat.n_tpts == 6


def test_neighbour_table_is_the_six_connected_graph_inside_the_mask():
    rng = np.random.default_rng(8)
    mask = (rng.uniform(size=(5, 4, 6)) > 0.3).astype(np.int32)
    dm = DataModel(np.zeros((5, 4, 6, 2), dtype=np.float32), mask=mask)
    nb = dm.neighbour_table()
    coords = dm.voxel_coords()
    index = {tuple(c): i for i, c in enumerate(coords)}
    steps = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
    for w, c in enumerate(coords):
        for k, s in enumerate(steps):
            assert nb[w, k] == index.get(tuple(c + np.asarray(s)), -1)
    # symmetric: u is a neighbour of w  <=>  w is a neighbour of u (in the opposite slot)
    for w in range(len(coords)):
        for k in range(6):
            u = nb[w, k]
            if u >= 0:
                assert nb[u, k ^ 1] == w
    # z is the fastest axis: +z / -z neighbours are adjacent voxel indices
    zs = nb[:, 5]
    assert ((zs == -1) | (zs == np.arange(len(coords)) + 1)).all()
