"""sd_gen_voxel_grid against the host construction and against the grid the reference itself builds
(tests/golden/voxel_grid.npz, oracle/make_golden_grid.py).  Needs a B200: run with ``-m gpu``.  (Last in the run order on
purpose: the kernel's centre arithmetic was changed to the reference's double evaluation after the round's GPU minutes were
spent -- checked by a CPU emulation of the same IEEE operations -- and this is its first run on the device.)"""
import numpy as np
import pytest

from oracle import oracle as O
from scenedino_b200 import ops
from scenedino_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("x_range", [None, (0, 1), (37, 101), (255, 256)])
def test_gen_voxel_grid_bit_identical_to_the_host_grid(x_range):
    """sd_gen_voxel_grid: the SSC grid of sscbench/evaluate_model_sscbench.py:270-278 made on the device, whole and in
    x-slabs, bit for bit what synthetic.ssc_voxel_grid (the host construction) gives."""
    want = syn.ssc_voxel_grid(x_range=x_range)
    got = ops.gen_voxel_grid(syn.velo_to_cam(), x_range=x_range)
    assert got.shape == want.shape and np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(got.cpu().numpy(), O.voxel_grid(syn.velo_to_cam(), x_range=x_range))      # the C oracle
    odd = ops.gen_voxel_grid(syn.velo_to_cam(), dims=(5, 3, 7), voxel_size=0.35, origin=(1.5, -2.25, 0.1))
    assert np.array_equal(odd.cpu().numpy(), syn.ssc_voxel_grid(dims=(5, 3, 7), voxel_size=0.35, origin=(1.5, -2.25, 0.1)))
    # ... and what the reference's own generate_point_grid + .float() produced (oracle/make_golden_grid.py)
    import hashlib
    import os
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxel_grid.npz"))
    assert np.array_equal(odd.cpu().numpy(), ref["odd"])
    if x_range is None:
        host = np.ascontiguousarray(got.cpu().numpy())
        assert hashlib.sha256(host.tobytes()).digest() == ref["sha256_f32"].tobytes()
        assert np.array_equal(host[ref["sample_idx"]], ref["sample"])
