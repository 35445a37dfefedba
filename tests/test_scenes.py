"""Scene builders (host NumPy) vs masks produced by the reference's own builders."""
from __future__ import annotations

import hashlib
import json

import numpy as np
import pytest
from conftest import GOLDEN

from fs.boundary_condition import build_scene, get_boundary_condition


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("num", (1, 2, 3, 4, 5))
@pytest.mark.parametrize("res", (16, 24, 32, 40, 64))
def test_small_scenes_equal_reference(num, res, masks_small):
    const, mask = build_scene(num, 2 * res, res)
    assert mask.dtype == np.uint8 and const.dtype == np.float32
    np.testing.assert_array_equal(mask, masks_small[f"bc{num}_r{res}_mask"])
    np.testing.assert_array_equal(const, masks_small[f"bc{num}_r{res}_const"])


SHAS = json.loads((GOLDEN / "masks_sha256.json").read_text())


@pytest.mark.parametrize("key", [k for k in SHAS if int(k.split("_r")[1]) <= 2048])
def test_large_scenes_hash_equal_reference(key):
    num, res = int(key[2]), int(key.split("_r")[1])
    const, mask = build_scene(num, 2 * res, res)
    assert sha(mask) == SHAS[key]["mask"]
    assert sha(const) == SHAS[key]["const"]
    assert int((mask == 0).sum()) == SHAS[key]["fluid"]


def test_dye_scenes_equal_reference():
    sc = np.load(GOLDEN / "dye_scenes.npz")
    for num in (1, 2, 3, 4, 5):
        for res in (16, 20, 24, 40):
            const, mask, dye = build_scene(num, 2 * res, res, with_dye=True)
            np.testing.assert_array_equal(mask, sc[f"bc{num}_r{res}_mask"])
            np.testing.assert_array_equal(const, sc[f"bc{num}_r{res}_const"])
            np.testing.assert_array_equal(dye, sc[f"bc{num}_r{res}_dye"])


def test_unknown_scene_and_dye():
    with pytest.raises(NotImplementedError):
        get_boundary_condition(7, 16, enable_dye=False, device="cpu")
    with pytest.raises(NotImplementedError):
        build_scene(0, 32, 16)


def test_non_2to1_scene_has_same_structure():
    """weak-scaling grids: X is a free parameter of the same scene description."""
    const, mask = build_scene(2, 96, 32)
    assert mask.shape == (96, 32)
    assert (mask[:2, 32 // 3: 2 * 32 // 3] == 2).all() and (mask[-2:, 32 // 3: 2 * (32 // 3)] == 3).all()
    assert (mask[:, :2] == 1).all() and (mask[:, -2:] == 1).all()


def test_balanced_strips_follow_the_cost_model():
    """fs.boundary_condition.scene_row_cost + fs.distributed.balanced_bounds: strips of (nearly) equal estimated work, every
    row assigned exactly once, and bc5 -- a third of it wall -- gets taller strips where its slab is."""
    import numpy as np

    from fs.boundary_condition import scene_row_cost
    from fs.distributed import Partition, balanced_bounds

    X, Y = 4096, 2048
    w = scene_row_cost(5, X, Y, 200)
    assert w.shape == (X,) and (w > 0).all()
    for world in (2, 4, 8):
        b = balanced_bounds(w, world, min_rows=64)
        assert b[0] == 0 and b[-1] == X and all(b[k + 1] - b[k] >= 64 for k in range(world))
        loads = [w[b[k]:b[k + 1]].sum() for k in range(world)]
        equal = [w[k * X // world:(k + 1) * X // world].sum() for k in range(world)]
        assert max(loads) <= max(equal) + 1e-9 and max(loads) / (sum(loads) / world) < 1.12
        parts = [Partition(X, r, world, 9, b) for r in range(world)]
        assert [p.owned() for p in parts] == [(b[k], b[k + 1]) for k in range(world)]
    assert np.diff(balanced_bounds(w, 8, 64))[0] > X // 8          # the slab rows (0 .. 11 X / 30) are cheap: the first strip is taller
    with __import__("pytest").raises(ValueError):
        Partition(X, 0, 2, 9, (0, 5, X))                            # a strip thinner than the halo
