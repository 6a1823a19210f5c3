"""GPU parity tests written after the round's GPU budget was spent: their first run is the round-end run, so they are
non-strict expected failures until a round has seen them pass (then the marker goes).  The file sorts last on purpose:
nothing runs after it.  Needs a B200: ``-m gpu``."""
import numpy as np
import pytest

from helpers import TOL_FP32
from scenedino_b200 import ops
from test_gpu_parity import _check_pass, dev, g2n, scenes_from_golden

pytestmark = pytest.mark.gpu

UNSEEN = pytest.mark.xfail(strict=False, reason="first run happens at round end (GPU budget of the round was spent); "
                                                "the oracle side of the same fixture is green in test_oracle_golden.py")


@UNSEEN
def test_render_d768_vs_reference(golden):
    """SURVEY 8d cfg 3 in small, fp32 path: the 768-d head (d_out = 769) with four colour views, coarse pass and the
    96-sample fine pass on the reference's own merged depths, against the reference's outputs."""
    g = golden("render_d768")
    _, dsc, _, dmlp = scenes_from_golden(g)
    rays = g["rays"][0]
    z = ops.sample_coarse(dev(rays), dev(g["u_coarse"]), dev(g["lin"]), True)
    assert np.array_equal(g2n(z), g["coarse.z_samps"][0])
    for p in ("coarse.", "fine."):
        o = ops.render_pass(dsc, dmlp, dev(rays), dev(g[p + "z_samps"][0]), hard_alpha_cap=True, precision=ops.FP32)
        assert o["dino_features"].shape == (48, 768)
        _check_pass(o, g, p, TOL_FP32)
