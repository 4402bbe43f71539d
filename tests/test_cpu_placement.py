"""Round-2 groundwork (CPU only): the sort-free placement formula for transposed cross-level Gram
entries (oracle/placement_proto.py) reproduces a brute-force sort exactly."""
import numpy as np
import pytest

from oracle import nksr_oracle as O
from oracle import placement_proto as P
from tests import clouds


@pytest.mark.parametrize("l,k,prune", [(0, 1, False), (0, 2, False), (1, 1, False), (0, 1, True), (0, 2, True)])
def test_structural_placement_equals_sorted_placement(l, k, prune):
    xyz, _ = clouds.sphere(600, radius=0.3, noise=0.01, seed=3)
    blob = np.random.default_rng(1).normal(0, 0.08, (150, 3)).astype(np.float32) + np.float32([0.9, 0.1, -0.2])
    svh = O.OracleSVH(0.06, 3).build_point_splatting(np.concatenate([xyz, blob]))
    if prune:       # an adaptive hierarchy: half of level 0 removed, so some level-1 voxels have no children
        keys = list(svh.keys)
        keys[0] = keys[0][O.key_to_ijk(keys[0], 0)[:, 0] >= 0]
        svh = O.OracleSVH(0.06, 3).build_from_keys(keys)
    by_formula, seg_len = P.placement_by_structure(svh, l, k)
    by_sort = P.placement_by_sort(svh, l, k)
    assert by_formula == by_sort
    # the positions of every coarse voxel are a permutation of 0..len-1 and the prefix total is the length
    per_c = {}
    for (c, _), pos in by_formula.items():
        per_c.setdefault(c, []).append(pos)
    for c, lst in per_c.items():
        assert sorted(lst) == list(range(len(lst))) and seg_len[c] == len(lst)
