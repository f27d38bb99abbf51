"""fallingsand3d_b200 — B200-native voxel simulation step for FallingSand3D (hot path only).

Holds only what the path needs: csrc/ (hand-written sm_100a kernels + the C ABI of
include/fs3d.h) and the host-side mirror of the voxel-world interface.  See DESIGN.md.
"""
from .world import (EMPTY, SAND, WATER, STONE, GAS, OIL, HONEY, GRAVEL, SCENE_RANDOM8, SCENE_MIXED8, FLAG_MATERIALS8, FLAG_NO_FUSE4, SCENE_EMPTY, SCENE_SAND_BLOCK, SCENE_MIXED, SCENE_RANDOM,
                    SCENE_MIXED_NOISE, FLAG_SKIP_SETTLED, FLAG_NO_FUSE, FLAG_NO_PEER_PUSH, FLAG_PEER_PUSH_SHARED_DEVICE, FLAG_EXPORTABLE, RM_SDF_SPHERE, RM_VOXELS, RM_SRGB, RM_BRICKS, RM_NO_BRICKS, Fs3dError, VoxelWorld)

__all__ = ["EMPTY", "SAND", "WATER", "STONE", "GAS", "OIL", "HONEY", "GRAVEL", "SCENE_RANDOM8", "SCENE_MIXED8", "FLAG_MATERIALS8", "FLAG_NO_FUSE4", "SCENE_EMPTY", "SCENE_SAND_BLOCK", "SCENE_MIXED", "SCENE_RANDOM",
           "SCENE_MIXED_NOISE", "FLAG_SKIP_SETTLED", "FLAG_NO_FUSE", "FLAG_NO_PEER_PUSH", "FLAG_PEER_PUSH_SHARED_DEVICE", "FLAG_EXPORTABLE", "RM_SDF_SPHERE", "RM_VOXELS", "RM_SRGB", "RM_BRICKS", "RM_NO_BRICKS", "Fs3dError",
           "VoxelWorld"]
