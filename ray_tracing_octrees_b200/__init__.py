"""B200-native ray-casting core for the Ray_Tracing_Octrees hot path (BVH + octree traversal).

The product is the C-ABI library librto.so (include/rto_c.h, csrc/); this package is its Python binding and
the mirror of the reference's class interface used by tests and bench.py.
"""
from .api import (BVH, Camera, HostBVH, RayTracerBVH, Scene, VoxelGrid, city_block_grid, create_octree_from_voxel_grid,
                  generate_test_volume, marching_cubes_mesh, MISS_T)
from ._lib import (FLAG_NO_PRUNE, FLAG_SHADOWS, MEM_DEVICE, MEM_HOST, MODE_BVH, MODE_OCTREE_GLSL, MODE_OCTREE_SKIP,
                   RtoCamera, RtoError, RtoFrame, lib)
