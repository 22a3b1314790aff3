"""B200-native ray-casting core for the Ray_Tracing_Octrees hot path (BVH + octree traversal).

The product is the C-ABI library librto.so (include/rto_c.h, csrc/); this package is its Python binding and
the mirror of the reference's class interface used by tests and bench.py.
"""
from .api import (BVH, Camera, ExchangeBuffer, Group, codes_frame_words, HostBVH, RayTracerBVH, Scene, VoxelGrid, city_block_grid, create_octree_from_voxel_grid,
                  create_octree_on_device, dual_contouring_mesh, dual_contouring_mesh_with_normals, load_triangle_cache, save_triangle_cache, frustum_cull, generate_test_volume, skip_distance_from_probes, skip_probe_rays, view_proj, load_csv_data_into_voxel_grid, marching_cubes_mesh, marching_cubes_mesh_on_device, MISS_T)
from ._lib import (FLAG_NO_PRUNE, FLAG_SHADOWS, FLAG_SORT_RAYS, MEM_DEVICE, MEM_HOST, MODE_BVH, MODE_OCTREE_GLSL, MODE_OCTREE_SKIP,
                   RtoCamera, RtoError, RtoFrame, lib)
