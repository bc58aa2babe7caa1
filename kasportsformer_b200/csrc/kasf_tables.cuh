// kasf_tables.cuh -- the fixed 17-joint skeleton tables, held in constant memory.
// Sources: bone endpoints reference model/KASportsFormer.py:46-47; limb groups
// model/modules/bone_refusion.py:34-40; adjacency model/modules/graph.py:16-17; flip pairs
// utils/utilities.py:128.  `kasf_table()` exposes them so tests can compare with the host copies.
#pragma once
#include <stdint.h>

namespace kasf {

#define KASF_BONE_CHILD  {0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15}
#define KASF_BONE_PARENT {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}
#define KASF_LIMB_SIZE   {3, 3, 2, 2, 3, 3, 4, 4, 4, 4, 3, 4, 4, 4, 4, 2, 2}
#define KASF_LIMB_MEMBER                                                                             \
    {0, 1, 2, -1,   3, 4, 5, -1,   6, 7, -1, -1,   8, 9, -1, -1,   10, 11, 12, -1,   13, 14, 15, -1, \
     6, 7, 1, 2,    6, 7, 4, 5,    6, 7, 11, 12,   6, 7, 14, 15,   6, 7, 9, -1,                      \
     14, 15, 11, 12,   1, 2, 4, 5,   14, 15, 4, 5,   11, 12, 4, 5,   10, 0, -1, -1,   13, 3, -1, -1}
// neighbour lists (padded with -1 to 4) and degrees of the undirected skeleton, no self loops
#define KASF_NBR                                                                                      \
    {1, 7, 4, -1,   2, 0, -1, -1,   3, 1, -1, -1,   2, -1, -1, -1,   5, 0, -1, -1,   6, 4, -1, -1,    \
     5, -1, -1, -1,   0, 8, -1, -1,   7, 9, 11, 14,   8, 10, -1, -1,   9, -1, -1, -1,   12, 8, -1, -1, \
     13, 11, -1, -1,   12, -1, -1, -1,   15, 8, -1, -1,   16, 14, -1, -1,   15, -1, -1, -1}
#define KASF_DEG {3, 2, 2, 1, 2, 2, 1, 2, 4, 2, 1, 2, 2, 1, 2, 2, 1}
// flip: joint j of the flipped pose is joint FLIP[j] of the original
#define KASF_FLIP {0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 14, 15, 16, 11, 12, 13}

static const int h_bone_child[16] = KASF_BONE_CHILD;
static const int h_bone_parent[16] = KASF_BONE_PARENT;
static const int h_limb_size[17] = KASF_LIMB_SIZE;
static const int h_limb_member[68] = KASF_LIMB_MEMBER;
static const int h_nbr[68] = KASF_NBR;
static const int h_deg[17] = KASF_DEG;
static const int h_flip[17] = KASF_FLIP;

}  // namespace kasf
