"""Constant skeleton tables of the 17-joint H36M-style pose used by KASportsFormer.

These define results, so each table cites where the reference fixes it:
  * bone endpoints ........ reference model/KASportsFormer.py:46-47
  * limb groups ........... reference model/modules/bone_refusion.py:34-40
  * skeleton adjacency .... reference model/modules/graph.py:16-17
  * flip pairs ............ reference utils/utilities.py:128

The same tables are baked into `csrc/kasf_tables.cuh` (constant memory); `tests/test_tables.py`
checks the two copies against each other through the C-ABI.
"""

NUM_JOINTS = 17

# bone k = joints[BONE_CHILD[k]] - joints[BONE_PARENT[k]], k = 0..15
BONE_CHILD = (0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15)
BONE_PARENT = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16)

# 17 limb groups; entries index *joints* of the raw input (the reference feeds raw joints here)
LIMB_GROUPS = (
    (0, 1, 2), (3, 4, 5), (6, 7), (8, 9), (10, 11, 12), (13, 14, 15),
    (6, 7, 1, 2), (6, 7, 4, 5), (6, 7, 11, 12), (6, 7, 14, 15), (6, 7, 9),
    (14, 15, 11, 12), (1, 2, 4, 5),
    (14, 15, 4, 5), (11, 12, 4, 5),
    (10, 0), (13, 3),
)
LIMB_HIDDEN = 16
LIMB_CHANNEL_NAMES = ("mlp_dir_x", "mlp_dir_y", "mlp_len")  # applied to input channels x, y, conf

# undirected skeleton edges (no self loops): 16 edges = 32 directed entries
SKELETON_NEIGHBOURS = {
    0: (1, 7, 4), 1: (2, 0), 2: (3, 1), 3: (2,), 4: (5, 0), 5: (6, 4), 6: (5,),
    7: (0, 8), 8: (7, 9, 11, 14), 9: (8, 10), 10: (9,), 11: (12, 8), 12: (13, 11),
    13: (12,), 14: (15, 8), 15: (16, 14), 16: (15,),
}

FLIP_LEFT = (1, 2, 3, 14, 15, 16)
FLIP_RIGHT = (4, 5, 6, 11, 12, 13)


def skeleton_degrees():
    return tuple(len(SKELETON_NEIGHBOURS[i]) for i in range(NUM_JOINTS))


def flip_permutation():
    """perm[j] = source joint that lands on j after a left/right flip."""
    perm = list(range(NUM_JOINTS))
    for l, r in zip(FLIP_LEFT, FLIP_RIGHT):
        perm[l], perm[r] = r, l
    return tuple(perm)
