"""Golden vector for the SSC voxel grid: runs the REFERENCE's own construction (sscbench/evaluate_model_sscbench.py:270-278:
generate_point_grid -> TSDFVolume.vox2world (numba) -> rigid_transform, then torch .float()) and stores digests + a strided
sample in tests/golden/voxel_grid.npz.  Test infrastructure; needs /root/reference (this container only).

    python oracle/make_golden_grid.py
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SD_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if "skimage" not in sys.modules:                    # fusion.py imports skimage.measure for its mesh export only
        try:
            import skimage  # noqa: F401
        except ImportError:
            m = types.ModuleType("skimage"); m.measure = types.ModuleType("skimage.measure")
            sys.modules["skimage"] = m; sys.modules["skimage.measure"] = m.measure
    sys.path.insert(0, os.path.join(REF, "sscbench")); sys.path.insert(0, REF)
    import point_utils as pu
    from fusion import TSDFVolume, rigid_transform

    T = pu.read_calib()["Tr"]
    pts, _ = pu.generate_point_grid(vox_origin=np.array([0, -25.6, -2]), scene_size=(51.2, 51.2, 6.4), voxel_size=0.2,
                                    cam_E=T, cam_k=pu.get_cam_k())
    full = torch.tensor(pts).reshape(1, -1, 3).float().numpy()[0]          # evaluate_model_sscbench.py:277
    idx = np.arange(0, len(full), 4099)
    # a second, odd grid straight through the two functions generate_point_grid calls
    dims, origin, vs = (5, 3, 7), np.array([1.5, -2.25, 0.1]), 0.35
    xv, yv, zv = np.meshgrid(range(dims[0]), range(dims[1]), range(dims[2]), indexing="ij")
    vox = np.concatenate([xv.reshape(1, -1), yv.reshape(1, -1), zv.reshape(1, -1)], axis=0).astype(int).T
    odd = rigid_transform(TSDFVolume.vox2world(origin, vox, vs), T)
    odd = torch.tensor(odd).float().numpy()
    out = dict(T=np.asarray(T, np.float64), dims=np.array([256, 256, 32]), origin=np.array([0, -25.6, -2.0]), voxel_size=np.float64(0.2),
               sha256_f32=np.frombuffer(hashlib.sha256(np.ascontiguousarray(full).tobytes()).digest(), np.uint8),
               sample_idx=idx, sample=full[idx],
               slab_37_101=full[37 * 256 * 32:37 * 256 * 32 + 64], odd_dims=np.array(dims), odd_origin=origin,
               odd_voxel_size=np.float64(vs), odd=odd)
    path = os.path.join(ROOT, "tests", "golden", "voxel_grid.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(full), "voxels")


if __name__ == "__main__":
    main()
