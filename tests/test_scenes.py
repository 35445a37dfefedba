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
